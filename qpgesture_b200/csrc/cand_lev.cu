// Mode-B candidate distance: unit-cost Levenshtein distance between 11-token
// strings (token = g0*320+g1 of a vq-wav2vec frame; GestureKNN.py:44-67
// 'combine'), fused with the min-by-start-code reduction of
// search_audio_cands(mode='wavvq_feat') (GestureKNN.py:666-691).
//
// Bound: integer ALU (121 DP cells per pair, 48 bytes per window), not HBM.
// One thread per window keeps the 12-entry DP row in registers; queries sit in
// shared memory (uniform broadcast reads).  Exact integers.
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int L = QPG_LEV_TOKENS;   // 11
constexpr int LS = QPG_LEV_STRIDE;  // 12
constexpr int KB = QPG_CODEBOOK_SIZE;
constexpr int LQT = 4;              // queries per pass

__device__ __forceinline__ int lev11(const uint32_t* __restrict__ a /* query, shared */, const uint32_t (&b)[LS]) {
  int prev[L + 1], cur[L + 1];
#pragma unroll
  for (int j = 0; j <= L; ++j) prev[j] = j;
#pragma unroll
  for (int i = 1; i <= L; ++i) {
    const uint32_t ai = a[i - 1];
    cur[0] = i;
#pragma unroll
    for (int j = 1; j <= L; ++j) {
      const int sub = prev[j - 1] + (ai != b[j - 1] ? 1 : 0);
      const int del = prev[j] + 1;
      const int ins = cur[j - 1] + 1;
      cur[j] = min(min(sub, del), ins);
    }
#pragma unroll
    for (int j = 0; j <= L; ++j) prev[j] = cur[j];
  }
  return prev[L];
}

__device__ __forceinline__ void load_tokens(const uint32_t* __restrict__ p, uint32_t (&b)[LS]) {
  const uint4* v = reinterpret_cast<const uint4*>(p);
  const uint4 t0 = v[0], t1 = v[1], t2 = v[2];
  b[0] = t0.x; b[1] = t0.y; b[2] = t0.z; b[3] = t0.w;
  b[4] = t1.x; b[5] = t1.y; b[6] = t1.z; b[7] = t1.w;
  b[8] = t2.x; b[9] = t2.y; b[10] = t2.z; b[11] = t2.w;
}

__global__ void __launch_bounds__(256)
    cand_lev_kernel(const uint32_t* __restrict__ tokens, const int32_t* __restrict__ labels, int64_t W,
                    int64_t id_offset, const uint32_t* __restrict__ q_tokens, int nq, Pair* __restrict__ table) {
  __shared__ Pair tab[LQT * KB];
  __shared__ uint32_t qs[LQT * LS];
  const int tid = threadIdx.x;
  for (int i = tid; i < LQT * KB; i += blockDim.x) {
    Pair p;
    p.d = (unsigned long long)__double_as_longlong(kEmptyDist);
    p.id = ~0ull;
    tab[i] = p;
  }
  if (tid < LQT * LS) qs[tid] = (tid / LS) < nq ? q_tokens[tid] : 0u;
  __syncthreads();
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + tid; w < W; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t b[LS];
    load_tokens(tokens + w * LS, b);
    const int label = labels[w];
    if ((unsigned)label >= (unsigned)KB) continue;
#pragma unroll
    for (int qi = 0; qi < LQT; ++qi) {
      if (qi < nq) {
        const int d = lev11(qs + qi * LS, b);
        Pair mine;
        mine.d = (unsigned long long)__double_as_longlong((double)d);
        mine.id = (unsigned long long)(id_offset + w);
        pair_min_shared(&tab[qi * KB + label], mine);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nq * KB; i += blockDim.x) {
    const Pair e = tab[i];
    if (e.id != ~0ull) pair_min_global(&table[i], e);
  }
}

__global__ void lev_pairs_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n,
                                 int32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t bb[LS];
  load_tokens(b + i * LS, bb);
  out[i] = lev11(a + i * LS, bb);
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_cand_lev_minbycode(const uint32_t* tokens, const int32_t* labels, int64_t W, int64_t id_offset,
                                      const uint32_t* q_tokens, int Q, qpg_pair_t* table, void* stream) {
  QPG_CHECK_ARG(W >= 0 && Q >= 0, "W >= 0, Q >= 0");
  if (W == 0 || Q == 0) return QPG_OK;
  QPG_CHECK_ARG(tokens && labels && q_tokens && table, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(tokens) & 15) == 0, "tokens must be 16-byte aligned");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(table) & 15) == 0, "table must be 16-byte aligned");
  int64_t blocks = (W + 255) / 256;
  const int cap = sm_count() * 4;
  if (blocks > cap) blocks = cap;
  for (int q0 = 0; q0 < Q; q0 += LQT) {
    const int nq = (Q - q0) < LQT ? (Q - q0) : LQT;
    cand_lev_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        tokens, labels, W, id_offset, q_tokens + (size_t)q0 * LS, nq, reinterpret_cast<Pair*>(table) + (size_t)q0 * KB);
    QPG_LAUNCH_CHECK();
  }
  return QPG_OK;
}

extern "C" int qpg_lev_distance(const uint32_t* a_tokens, const uint32_t* b_tokens, int64_t n, int32_t* out,
                                void* stream) {
  QPG_CHECK_ARG(n >= 0, "n >= 0");
  if (n == 0) return QPG_OK;
  QPG_CHECK_ARG(a_tokens && b_tokens && out, "null pointer");
  QPG_CHECK_ARG(((reinterpret_cast<uintptr_t>(a_tokens) | reinterpret_cast<uintptr_t>(b_tokens)) & 15) == 0,
                "tokens must be 16-byte aligned");
  lev_pairs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a_tokens, b_tokens, n, out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
