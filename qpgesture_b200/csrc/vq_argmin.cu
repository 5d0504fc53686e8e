// VQ codebook L2 argmin (BottleneckBlock.quantise, codebook/models/bottleneck.py:120-126)
// and embedding lookup (dequantise, :128-130).
//
//   dist[m][k] = fl32( fl32(|x_m|^2 - 2*<x_m, c_k>) + |c_k|^2 ),   idx[m] = first argmin_k
//
// The reference evaluates this in float32 through an SGEMM whose summation order is
// library defined; here every inner product is accumulated in float64 and rounded once
// (the correctly rounded value any float32 summation order approximates), then the
// reference's two float32 roundings are applied.  The [M, K] distance matrix is never
// written to memory.  Bound: FP64 pipe (2*M*K*D flop); M is small on the matcher's path
// (30 latents per 4-s sequence).
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int TM = 16;        // latents per CTA
constexpr int THREADS = 128;  // one code per thread; the codebook is split over gridDim.y CTAs

// order-preserving map float -> uint32 (handles the slightly negative distances rounding can produce)
__device__ __forceinline__ uint32_t ordered_u32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered_f32(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void vq_key_init_kernel(unsigned long long* __restrict__ keys, int64_t M) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < M) keys[i] = ~0ull;
}

__global__ void vq_key_finish_kernel(int64_t* __restrict__ idx_io, float* __restrict__ min_out, int64_t M) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const unsigned long long key = (unsigned long long)idx_io[i];
  if (min_out) min_out[i] = unordered_f32((uint32_t)(key >> 32));
  idx_io[i] = (int64_t)(key & 0xffffffffull);
}

// keys[m] = min over codes of (ordered(dist) << 32 | k): lexicographic (distance, code index) -> first minimum.
// The latent tile is converted to float64 once in shared memory, so the inner loop is 1 F2F per 16 DFMA.
__global__ void __launch_bounds__(THREADS)
    vq_argmin_kernel(const float* __restrict__ x, const float* __restrict__ cb, int64_t M, int D, int K,
                     unsigned long long* __restrict__ keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);           // [TM][Dp] float64 copies of the latents
  float* xn = reinterpret_cast<float*>(xs + (size_t)TM * ((D + 3) & ~3));   // [TM] squared norms (float32)
  const int Dp = (D + 3) & ~3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * TM;

  for (int i = tid; i < TM * Dp; i += THREADS) {
    const int r = i / Dp, d = i - r * Dp;
    const int64_t m = m0 + r;
    xs[i] = (m < M && d < D) ? (double)x[m * D + d] : 0.0;
  }
  __syncthreads();
  for (int r = warp; r < TM; r += THREADS / 32) {
    double s = 0.0;
    for (int d = lane; d < Dp; d += 32) s = fma(xs[r * Dp + d], xs[r * Dp + d], s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) xn[r] = (float)s;
  }
  __syncthreads();

  const int k = blockIdx.y * THREADS + tid;
  if (k < K) {
    const float* row = cb + (size_t)k * D;
    double acc[TM], kn = 0.0;
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r] = 0.0;
    const bool vec = (D & 3) == 0;
    for (int d = 0; d < Dp; d += 4) {
      float4 cv;
      if (vec) {
        cv = *reinterpret_cast<const float4*>(row + d);
      } else {
        cv.x = d < D ? row[d] : 0.f;
        cv.y = d + 1 < D ? row[d + 1] : 0.f;
        cv.z = d + 2 < D ? row[d + 2] : 0.f;
        cv.w = d + 3 < D ? row[d + 3] : 0.f;
      }
      const double c0 = cv.x, c1 = cv.y, c2 = cv.z, c3 = cv.w;
      kn = fma(c0, c0, kn);
      kn = fma(c1, c1, kn);
      kn = fma(c2, c2, kn);
      kn = fma(c3, c3, kn);
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const double2 x01 = *reinterpret_cast<const double2*>(xs + (size_t)r * Dp + d);      // broadcast reads
        const double2 x23 = *reinterpret_cast<const double2*>(xs + (size_t)r * Dp + d + 2);
        double a = acc[r];
        a = fma(x01.x, c0, a);
        a = fma(x01.y, c1, a);
        a = fma(x23.x, c2, a);
        a = fma(x23.y, c3, a);
        acc[r] = a;
      }
    }
    const float knf = (float)kn;
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      if (m0 + r < M) {
        // the reference's float32 formula on a correctly rounded inner product (bottleneck.py:123)
        const float dist = __fadd_rn(__fsub_rn(xn[r], __fmul_rn(2.0f, (float)acc[r])), knf);
        const unsigned long long key = ((unsigned long long)ordered_u32(dist) << 32) | (unsigned long long)k;
        atomicMin(&keys[m0 + r], key);
      }
    }
  }
}

// ---- few latents (one 4-s sequence = 30): ONE launch, no key buffer ------------------------------------------
// A cluster of 8 CTAs owns SM_TM latents; each CTA takes 64 codes (one per thread), so a 30-latent call spreads
// over 120 CTAs instead of 8.  Block minima of the lexicographic (distance, code) key go to CTA 0 of the cluster
// through distributed shared memory, which writes idx / min directly: no init / finish kernels, no atomics.  The
// arithmetic per (latent, code) is vq_argmin_kernel's, operation for operation (same float64 accumulation order,
// same warp reduction for |x|^2), so the two kernels return identical bits.
constexpr int SM_TM = 2;
constexpr int SM_CODES = 64;     // codes (= threads) per CTA
constexpr int SM_CLUSTER = 8;    // CTAs per cluster: up to 512 codes
__global__ void __cluster_dims__(1, SM_CLUSTER, 1) __launch_bounds__(SM_CODES)
    vq_argmin_small_kernel(const float* __restrict__ x, const float* __restrict__ cb, int64_t M, int D, int K,
                           int64_t* __restrict__ idx_out, float* __restrict__ min_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Dp = (D + 3) & ~3;
  double* xs = reinterpret_cast<double*>(smem_raw);                 // [SM_TM][Dp]
  float* xn = reinterpret_cast<float*>(xs + (size_t)SM_TM * Dp);    // [SM_TM]
  unsigned long long* red = reinterpret_cast<unsigned long long*>(xn + 4);      // [SM_TM][2] warp minima
  unsigned long long* gather = red + SM_TM * 2;                                 // [SM_CLUSTER][SM_TM], used in CTA 0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * SM_TM;
  // every CTA of the cluster is running before anyone writes into CTA 0's shared memory (completed at the wait below)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  for (int i = tid; i < SM_TM * Dp; i += SM_CODES) {
    const int r = i / Dp, d = i - r * Dp;
    const int64_t m = m0 + r;
    xs[i] = (m < M && d < D) ? (double)x[m * D + d] : 0.0;
  }
  __syncthreads();
  for (int r = warp; r < SM_TM; r += SM_CODES / 32) {
    double s = 0.0;
    for (int d = lane; d < Dp; d += 32) s = fma(xs[r * Dp + d], xs[r * Dp + d], s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) xn[r] = (float)s;
  }
  __syncthreads();
  const int k = blockIdx.y * SM_CODES + tid;
  unsigned long long key[SM_TM];
#pragma unroll
  for (int r = 0; r < SM_TM; ++r) key[r] = ~0ull;
  if (k < K) {
    const float* row = cb + (size_t)k * D;
    double acc[SM_TM], kn = 0.0;
#pragma unroll
    for (int r = 0; r < SM_TM; ++r) acc[r] = 0.0;
    // d ascending, four elements per step, exactly as vq_argmin_kernel; the loads run one 128-byte line (8 float4) ahead
    auto step = [&](const float4 cv, int d) {
      const double c0 = cv.x, c1 = cv.y, c2 = cv.z, c3 = cv.w;
      kn = fma(c0, c0, kn);
      kn = fma(c1, c1, kn);
      kn = fma(c2, c2, kn);
      kn = fma(c3, c3, kn);
#pragma unroll
      for (int r = 0; r < SM_TM; ++r) {
        const double2 x01 = *reinterpret_cast<const double2*>(xs + (size_t)r * Dp + d);
        const double2 x23 = *reinterpret_cast<const double2*>(xs + (size_t)r * Dp + d + 2);
        double a = acc[r];
        a = fma(x01.x, c0, a);
        a = fma(x01.y, c1, a);
        a = fma(x23.x, c2, a);
        a = fma(x23.y, c3, a);
        acc[r] = a;
      }
    };
    if ((D & 31) == 0) {
      float4 cur[8], nxt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = __ldg(reinterpret_cast<const float4*>(row) + j);
      for (int d = 0; d < D; d += 32) {
        if (d + 32 < D) {
#pragma unroll
          for (int j = 0; j < 8; ++j) nxt[j] = __ldg(reinterpret_cast<const float4*>(row + d + 32) + j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) step(cur[j], d + 4 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      }
    } else {
      for (int d = 0; d < Dp; d += 4) {
        float4 cv;
        cv.x = d < D ? row[d] : 0.f;
        cv.y = d + 1 < D ? row[d + 1] : 0.f;
        cv.z = d + 2 < D ? row[d + 2] : 0.f;
        cv.w = d + 3 < D ? row[d + 3] : 0.f;
        step(cv, d);
      }
    }
    const float knf = (float)kn;
#pragma unroll
    for (int r = 0; r < SM_TM; ++r) {
      const float dist = __fadd_rn(__fsub_rn(xn[r], __fmul_rn(2.0f, (float)acc[r])), knf);
      key[r] = ((unsigned long long)ordered_u32(dist) << 32) | (unsigned long long)k;
    }
  }
#pragma unroll
  for (int r = 0; r < SM_TM; ++r) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[r], o);
      key[r] = other < key[r] ? other : key[r];
    }
    if (lane == 0) red[r * 2 + warp] = key[r];
  }
  __syncthreads();
  // block minimum -> slot [my rank] of CTA 0's gather array (distributed shared memory)
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  if (tid < SM_TM) {
    const unsigned long long v = red[tid * 2] < red[tid * 2 + 1] ? red[tid * 2] : red[tid * 2 + 1];
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(gather + rank * SM_TM + tid)), "r"(0u));
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"(v) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && tid < SM_TM && m0 + tid < M) {
    unsigned long long v = ~0ull;
#pragma unroll
    for (int c = 0; c < SM_CLUSTER; ++c) {
      const unsigned long long o = gather[c * SM_TM + tid];
      v = o < v ? o : v;
    }
    idx_out[m0 + tid] = (int64_t)(v & 0xffffffffull);
    if (min_out) min_out[m0 + tid] = unordered_f32((uint32_t)(v >> 32));
  }
}

constexpr int64_t kSmallM = 256;
inline bool small_ok(int64_t M, int D, int K) {
  return M <= kSmallM && K <= SM_CODES * SM_CLUSTER && (size_t)SM_TM * ((D + 3) & ~3) * sizeof(double) <= 160 * 1024;
}
int launch_small(const float* x, const float* cb, int64_t M, int D, int K, int64_t* idx_out, float* min_out,
                 cudaStream_t st) {
  const int Dp = (D + 3) & ~3;
  const size_t smem = (size_t)SM_TM * Dp * sizeof(double) + 4 * sizeof(float) +
                      (SM_TM * 2 + SM_CLUSTER * SM_TM) * sizeof(unsigned long long);
  QPG_CUDA(cudaFuncSetAttribute(vq_argmin_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((M + SM_TM - 1) / SM_TM), SM_CLUSTER);
  vq_argmin_small_kernel<<<grid, SM_CODES, smem, st>>>(x, cb, M, D, K, idx_out, min_out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

// ---- fast path: tensor-core filter + exact re-evaluation of the near-minimal codes --------------------------
// G = X * C^T comes from the tcgen05 TF32 tap-GEMM (conv1d_tc.cu, one tap).  TF32 truncates both operands to 11
// significant bits, so |G~ - G| <= ~2^-9 * sum|x_i c_i| <= 2^-9 |x||c|; with the margin below every code whose
// exact distance could be the minimum is re-evaluated with the SAME arithmetic as vq_argmin_kernel (float64
// inner product rounded once, then the reference's two float32 roundings) and the lexicographic (distance,
// index) minimum over those is taken - the result is the exact kernel's, at a fraction of its cost (typically
// 2-6 of the 512 codes per latent survive the filter).  One warp per latent.
__global__ void vq_code_norms_kernel(const float* __restrict__ cb, int D, int K, float* __restrict__ kn_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float* row = cb + (size_t)k * D;
  double kn = 0.0;                                                  // same ascending order as vq_argmin_kernel
  for (int d = 0; d < D; ++d) kn = fma((double)row[d], (double)row[d], kn);
  kn_out[k] = (float)kn;
}

__global__ void __launch_bounds__(256)
    vq_select_kernel(const float* __restrict__ x, const float* __restrict__ cb, const float* __restrict__ G,
                     const float* __restrict__ kn_in, int64_t M, int D, int K, int64_t* __restrict__ idx_out,
                     float* __restrict__ min_out) {
  extern __shared__ float s_kn[];                                   // [K] |c_k|^2 (float64 accumulated, rounded)
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_kn[k] = kn_in[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  const float* xr = x + (size_t)m * D;
  double xs = 0.0;
  for (int d = lane; d < D; d += 32) xs = fma((double)xr[d], (double)xr[d], xs);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) xs += __shfl_xor_sync(0xffffffffu, xs, o);
  // vq_argmin_kernel sums the squares of a latent in the same lane-strided order, so xn is bit-identical
  const float xn = (float)xs;
  const float xnorm = sqrtf(xn);
  // approximate distances (without the common |x|^2) and the acceptance bound per code
  float best = 3.0e38f;
  for (int k = lane; k < K; k += 32) {
    const float g = G[(size_t)m * K + k];
    const float kn = s_kn[k];
    const float hi = (kn - 2.0f * g) + 0.0079f * xnorm * sqrtf(kn) + 2e-6f * (xn + kn);     // upper bound of the exact value
    best = fminf(best, hi);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  unsigned long long key = ~0ull;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    bool cand = false;
    if (k < K) {
      const float g = G[(size_t)m * K + k];
      const float kn = s_kn[k];
      const float lo = (kn - 2.0f * g) - 0.0079f * xnorm * sqrtf(kn) - 2e-6f * (xn + kn);   // lower bound
      cand = lo <= best;
    }
    unsigned mask = __ballot_sync(0xffffffffu, cand);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const int kc = k0 + src;
      const float* row = cb + (size_t)kc * D;
      double acc = 0.0;
      for (int d = lane; d < D; d += 32) acc = fma((double)xr[d], (double)row[d], acc);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const float dist = __fadd_rn(__fsub_rn(xn, __fmul_rn(2.0f, (float)acc)), s_kn[kc]);
      const unsigned long long kk = ((unsigned long long)ordered_u32(dist) << 32) | (unsigned long long)kc;
      key = kk < key ? kk : key;
    }
  }
  if (lane == 0) {
    idx_out[m] = (int64_t)(key & 0xffffffffull);
    if (min_out) min_out[m] = unordered_f32((uint32_t)(key >> 32));
  }
}

__global__ void dequantise_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb, int64_t M, int D,
                                  int K, float* __restrict__ out) {
  const int64_t m = blockIdx.x;
  int64_t k = idx[m];
  if (k < 0) k = 0;
  if (k >= K) k = K - 1;
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[m * D + d] = cb[k * D + d];
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_vq_argmin_f32(const float* x, const float* codebook, int64_t M, int D, int K, int64_t* idx_out,
                                 float* min_out, void* stream) {
  QPG_CHECK_ARG(M >= 0 && D > 0 && K > 0, "M >= 0, D > 0, K > 0");
  if (M == 0) return QPG_OK;
  QPG_CHECK_ARG(x && codebook && idx_out, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, "codebook must be 16-byte aligned");
  const int Dp = (D + 3) & ~3;
  const size_t smem = (size_t)TM * Dp * sizeof(double) + TM * sizeof(float) + 64;
  QPG_CHECK_ARG(smem <= 200 * 1024, "D too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (small_ok(M, D, K)) return launch_small(x, codebook, M, D, K, idx_out, min_out, st);
  // idx_out doubles as the 64-bit key buffer (ordered distance << 32 | code) until the finish kernel
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(idx_out);
  vq_key_init_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(keys, M);
  QPG_LAUNCH_CHECK();
  // the attribute is per device: set it on every launch (a process may use several GPUs)
  QPG_CUDA(cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((K + THREADS - 1) / THREADS));
  vq_argmin_kernel<<<grid, THREADS, smem, st>>>(x, codebook, M, D, K, keys);
  QPG_LAUNCH_CHECK();
  vq_key_finish_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(idx_out, min_out, M);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_vq_dequantise_f32(const int64_t* idx, const float* codebook, int64_t M, int D, int K, float* out,
                                     void* stream) {
  QPG_CHECK_ARG(M >= 0 && D > 0 && K > 0, "M >= 0, D > 0, K > 0");
  if (M == 0) return QPG_OK;
  QPG_CHECK_ARG(idx && codebook && out, "null pointer");
  dequantise_kernel<<<(unsigned)M, 128, 0, (cudaStream_t)stream>>>(idx, codebook, M, D, K, out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_vq_argmin_fast(const float* x, const float* codebook, int64_t M, int D, int K, float* scratch,
                                  int64_t* idx_out, float* min_out, void* stream) {
  QPG_CHECK_ARG(M >= 0 && D > 0 && K > 0, "M >= 0, D > 0, K > 0");
  if (M == 0) return QPG_OK;
  QPG_CHECK_ARG(x && codebook && idx_out && scratch, "null pointer");
  QPG_CHECK_ARG(D % 4 == 0 && K % 16 == 0 && K <= 12288 && M < (1ll << 31), "needs D % 4 == 0, K % 16 == 0");
  // a handful of latents: the one-launch exact kernel beats three launches (GEMM, code norms, selection)
  if (small_ok(M, D, K)) return launch_small(x, codebook, M, D, K, idx_out, min_out, (cudaStream_t)stream);
  // G[M, K] = X * C^T on the tensor cores: the latents as a one-item sequence of M frames, the codebook as the
  // (K-major) weights of a single tap
  qpg_conv_tc_desc_t d;
  d.B = 1;
  d.T_view = (int)M;
  d.C_view = D;
  d.n_out = (int)M;
  d.C_in = D;
  d.C_out = K;
  d.K_pad = D;
  d.BN = K % 256 == 0 ? 256 : (K % 128 == 0 ? 128 : (K % 64 == 0 ? 64 : 16));
  d.N_pad = K;
  d.n_taps = 1;
  for (int i = 0; i < 4; ++i) d.row_offset[i] = d.chan_offset[i] = 0;
  d.out_rows_per_item = (int)M;
  d.out_ld = K;
  d.out_chan_offset = 0;
  const int rc = qpg_conv1d_taps_tf32(&d, x, codebook, nullptr, nullptr, scratch, nullptr, stream);
  if (rc != QPG_OK) return rc;
  float* kn = scratch + (size_t)M * K;
  vq_code_norms_kernel<<<(unsigned)((K + 127) / 128), 128, 0, (cudaStream_t)stream>>>(codebook, D, K, kn);
  QPG_LAUNCH_CHECK();
  const size_t smem = (size_t)K * sizeof(float);
  vq_select_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, smem, (cudaStream_t)stream>>>(x, codebook, scratch, kn, M, D,
                                                                                         K, idx_out, min_out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
