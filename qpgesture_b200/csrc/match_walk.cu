// Sequential tail of CodeKNN.search_code_knn (GestureKNN.py:528-660), split so that almost nothing is
// left on the dependent chain:
//
//   lookup (parallel, state independent): for every query step and EVERY possible previous code `last`
//       c_a = argmin_c (pos_rank[last][c] + 0.05*freq_rank[c]) + rank_a[c]      (:540-555,:574-576)
//       c_t = same with the text ranks
//     and everything that hangs off the two chosen windows (window id, the code that becomes `last` if
//     that candidate wins, its phase frame) is gathered into one 32-byte entry.  512 x 512 integer
//     work per step, spread over the whole chip.
//   walk (one CTA per clip): per step ONE entry load (both possible next entries are prefetched while
//     the phase distance is computed), the two 128-d phase-manifold cosines (:627-644), the pick
//     (audio wins ties, :646) - the same float64 arithmetic as match_tail_kernel.
//
// Tie handling (the reference ranks with NumPy's unstable argsort, so exact ties are platform defined
// there): this implementation is stable (lower code first) and REPORTS when a tie could have mattered:
// status bit 1 (value 2) is set when a visited step had an exact tie at the arg-min, an exact distance
// tie between non-empty bins, or an empty bin that could win under some ordering of the sentinel ties.
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int KB = QPG_CODEBOOK_SIZE;
constexpr int WIN = 26;
constexpr int NCODE = 30;
constexpr int NFRM = 240;
constexpr int PC = 16;

struct __align__(16) Entry {
  long long w[2];        // audio / text candidate window (global id), -1 = the chosen bin is empty
  short nl[2];           // code that becomes `last` when that candidate wins (payload[3])
  short nls[2];          // ... at the end of a segment (payload[1], GestureKNN.py:800)
  short frame[2];        // phase frame of the window start, int(k/398*240)
  int flags;             // bit 0: tie-dependent choice
};
static_assert(sizeof(Entry) == 32, "entry is 32 bytes");

// grid (Q, 4), 128 threads: thread = one value of `last`
__global__ void __launch_bounds__(128)
    match_lookup_kernel(const Pair* __restrict__ aud_table, const Pair* __restrict__ txt_table,
                        const int32_t* __restrict__ aud_rank, const int32_t* __restrict__ txt_rank,
                        const int16_t* __restrict__ pos_rank_t, const int32_t* __restrict__ freq_rank,
                        const int32_t* __restrict__ code, int64_t n_seq, const int32_t* __restrict__ aud_frame,
                        const int32_t* __restrict__ txt_frame, const int32_t* __restrict__ qflags_a,
                        const int32_t* __restrict__ qflags_t, Entry* __restrict__ entries) {
  constexpr int EMPTY = 1 << 30;
  __shared__ int s_key[2][KB];          // 20*rank + freq_rank per table; empty bins carry the EMPTY bit
  __shared__ int s_fr[KB];
  __shared__ int s_ne[2];
  const int q = blockIdx.x, tid = threadIdx.x;
  const int last = blockIdx.y * 128 + tid;
  if (tid < 2) s_ne[tid] = 0;
  __syncthreads();
  for (int c = tid; c < KB; c += 128) {
    const int fr = freq_rank[c];
    const long long ida = (long long)aud_table[(size_t)q * KB + c].id, idt = (long long)txt_table[(size_t)q * KB + c].id;
    s_fr[c] = fr;
    s_key[0][c] = (20 * aud_rank[(size_t)q * KB + c] + fr) | (ida < 0 ? EMPTY : 0);
    s_key[1][c] = (20 * txt_rank[(size_t)q * KB + c] + fr) | (idt < 0 ? EMPTY : 0);
    if (ida >= 0) atomicAdd(&s_ne[0], 1);
    if (idt >= 0) atomicAdd(&s_ne[1], 1);
  }
  __syncthreads();
  // integer keys 20*(pos + rank) + freq order exactly like NumPy's float64 (pos + freq*0.05) + rank whenever
  // they differ (the float64 rounding is ~1e-13, distinct keys are >= 0.05 apart)
  int best[2] = {0x7fffffff, 0x7fffffff}, arg[2] = {0, 0}, ties[2] = {0, 0};
  int best_ne[2] = {0x7fffffff, 0x7fffffff};       // best key over non-empty bins
  int lb_e[2] = {0x3fffffff, 0x3fffffff};          // min over empty bins of 20*pos + freq
#pragma unroll 4
  for (int c = 0; c < KB; ++c) {
    const int p20 = 20 * (int)pos_rank_t[(size_t)c * KB + last];
    const int fr = s_fr[c];
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int sk = s_key[x][c];
      const int key = p20 + (sk & ~EMPTY);
      if (key < best[x]) {
        best[x] = key;
        arg[x] = c;
        ties[x] = 1;
      } else if (key == best[x]) {
        ++ties[x];
      }
      if (sk & EMPTY) lb_e[x] = min(lb_e[x], p20 + fr);
      else best_ne[x] = min(best_ne[x], key);
    }
  }
  Entry e;
  e.flags = ((qflags_a && qflags_a[q]) || (qflags_t && qflags_t[q])) ? 1 : 0;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int32_t* rank = x == 0 ? aud_rank : txt_rank;
    if (ties[x] > 1) {
      // equal integer keys: NumPy compares the float64 values, which may differ in the last bits between
      // different (pos, rank) splits -> evaluate exactly those in float64 with NumPy's operation order
      double bv = 1e300;
      int bc = 0, nt = 0;
      for (int c = 0; c < KB; ++c) {
        const int pos = (int)pos_rank_t[(size_t)c * KB + last];
        if (20 * pos + (s_key[x][c] & ~EMPTY) != best[x]) continue;
        const int rk = rank[(size_t)q * KB + c];
        const double v = __dadd_rn(__dadd_rn((double)pos, __dmul_rn((double)s_fr[c], 0.05)), (double)rk);
        if (v < bv) {
          bv = v;
          bc = c;
          nt = 1;
        } else if (v == bv) {
          ++nt;
        }
      }
      arg[x] = bc;
      if (nt > 1) e.flags |= 1;                    // a true tie at the arg-min: NumPy's order is platform defined
    }
    // empty bins all hold the sentinel 1e3, their mutual rank order is NumPy's business: the lowest rank any of
    // them can get is the number of non-empty bins.  If that could reach the best non-empty key, say so.
    const int ne = s_ne[x];
    if (ne < KB && lb_e[x] + 20 * ne <= best_ne[x]) e.flags |= 1;
    const Pair* table = x == 0 ? aud_table : txt_table;
    const long long w = (long long)table[(size_t)q * KB + arg[x]].id;
    e.w[x] = -1;
    e.nl[x] = e.nls[x] = 0;
    e.frame[x] = 0;
    if (w >= 0 && w < n_seq * WIN) {
      const long long j = w / WIN;
      const int m = (int)(w - j * WIN);
      e.w[x] = w;
      e.nl[x] = (short)code[(size_t)j * NCODE + m + 3];
      e.nls[x] = (short)code[(size_t)j * NCODE + m + 1];
      e.frame[x] = (short)(x == 0 ? aud_frame : txt_frame)[m];
    }
  }
  entries[(size_t)q * KB + last] = e;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void bar64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ Entry load_entry(const Entry* p) {
  const int4* q = reinterpret_cast<const int4*>(p);
  const int4 a = __ldg(q), b = __ldg(q + 1);
  Entry e;
  e.w[0] = ((long long)(unsigned)a.y << 32) | (unsigned)a.x;
  e.w[1] = ((long long)(unsigned)a.w << 32) | (unsigned)a.z;
  e.nl[0] = (short)(b.x & 0xffff);
  e.nl[1] = (short)((unsigned)b.x >> 16);
  e.nls[0] = (short)(b.y & 0xffff);
  e.nls[1] = (short)((unsigned)b.y >> 16);
  e.frame[0] = (short)(b.z & 0xffff);
  e.frame[1] = (short)((unsigned)b.z >> 16);
  e.flags = b.w;
  return e;
}

// one CTA (2 warps) per clip: warp 0 scores the audio candidate, warp 1 the text candidate
__global__ void __launch_bounds__(64)
    match_walk_kernel(const Entry* __restrict__ entries, const int32_t* __restrict__ code,
                      const float* __restrict__ phase_amp, const int32_t* __restrict__ seed_code,
                      const float* __restrict__ seed_phase, int n_seg, int64_t* __restrict__ codes_out,
                      int32_t* __restrict__ vote_out, float* __restrict__ phase_out, int32_t* __restrict__ status_out) {
  __shared__ float prev[8 * PC];
  __shared__ double s_dist[2];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_steps = n_seg * 8;
  const size_t q0 = (size_t)b * n_steps;
  for (int e = threadIdx.x; e < 8 * PC; e += 64) prev[e] = seed_phase[(size_t)b * 8 * PC + e];
  for (int i = threadIdx.x; i < n_seg * NCODE; i += 64) codes_out[(size_t)b * n_seg * NCODE + i] = -1;
  int last = seed_code[b];
  int status = 0;
  if ((unsigned)last >= (unsigned)KB) {
    if (threadIdx.x == 0) status_out[b] = 1;
    return;
  }
  Entry cur = load_entry(entries + q0 * KB + last);
  __syncthreads();
  for (int st = 0; st < n_steps; ++st) {
    const int g = st >> 3, s = st & 7;
    const size_t q = q0 + st;
    status |= (cur.flags & 1) ? 2 : 0;
    if (cur.w[0] < 0 || cur.w[1] < 0) {             // IndexError at GestureKNN.py:631
      status |= 1;
      break;
    }
    // both possible next entries (state independent addresses once `cur` is known)
    Entry nx0 = cur, nx1 = cur;
    if (st + 1 < n_steps) {
      const int l0 = s == 7 ? cur.nls[0] : cur.nl[0], l1 = s == 7 ? cur.nls[1] : cur.nl[1];
      nx0 = load_entry(entries + (q + 1) * KB + l0);
      nx1 = load_entry(entries + (q + 1) * KB + l1);
    }
    const long long w = warp == 0 ? cur.w[0] : cur.w[1];
    const long long j = w / WIN;
    const int m = (int)(w - j * WIN);
    const int f = warp == 0 ? cur.frame[0] : cur.frame[1];
    const float* head = phase_amp + ((size_t)j * NFRM + f) * PC;      // rows f..f+7; rows f+24..f+31 = next prev
    float tl[4];
    double av[4], bv[4], sa = 0.0, sb = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e2 = lane + 32 * k, row = e2 >> 4, col = e2 & 15;
      const float fa = row < 5 ? prev[(3 + row) * PC + col] : head[(row - 5) * PC + col];
      const float fb = row < 3 ? prev[(5 + row) * PC + col] : head[(row - 3) * PC + col];
      tl[k] = head[24 * PC + e2];
      av[k] = (double)fa;
      bv[k] = (double)fb;
      sa = fma(av[k], av[k], sa);
      sb = fma(bv[k], bv[k], sb);
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    const double na = sa > 0.0 ? sqrt(sa) : 1.0, nb = sb > 0.0 ? sqrt(sb) : 1.0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double d = av[k] / na - bv[k] / nb;
      acc = fma(d, d, acc);
    }
    const double dist = 0.5 * warp_sum(acc);
    if (lane == 0) s_dist[warp] = dist;
    bar64(1);
    const int win = (s_dist[0] <= s_dist[1]) ? 0 : 1;                 // audio wins ties (:646)
    if (warp == win) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        prev[lane + 32 * k] = tl[k];
        if (phase_out) phase_out[(q * 8) * PC + lane + 32 * k] = tl[k];
      }
      if (lane < 4) {
        const int p = s * 4 + lane;
        if (p < NCODE) codes_out[((size_t)b * n_seg + g) * NCODE + p] = code[(size_t)j * NCODE + m + lane];
      }
      if (lane == 0) vote_out[q] = win;
    }
    cur = win == 0 ? nx0 : nx1;
    bar64(2);
  }
  if (threadIdx.x == 0) status_out[b] = status;
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_match_lookup(const qpg_pair_t* aud_table, const qpg_pair_t* txt_table, const int32_t* aud_rank,
                                const int32_t* txt_rank, const int16_t* pos_rank_t, const int32_t* freq_rank,
                                const int32_t* code, int64_t n_seq, const int32_t* aud_frame, const int32_t* txt_frame,
                                const int32_t* qflags_a, const int32_t* qflags_t, int Q, void* entries, void* stream) {
  QPG_CHECK_ARG(Q >= 0 && n_seq >= 0, "negative size");
  if (Q == 0) return QPG_OK;
  QPG_CHECK_ARG(aud_table && txt_table && aud_rank && txt_rank && pos_rank_t && freq_rank && code && aud_frame &&
                    txt_frame && entries,
                "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(entries) & 15) == 0, "entries must be 16-byte aligned");
  match_lookup_kernel<<<dim3((unsigned)Q, 4), 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const Pair*>(aud_table), reinterpret_cast<const Pair*>(txt_table), aud_rank, txt_rank, pos_rank_t,
      freq_rank, code, n_seq, aud_frame, txt_frame, qflags_a, qflags_t, reinterpret_cast<Entry*>(entries));
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_match_walk(const void* entries, const int32_t* code, const float* phase_amp,
                              const int32_t* seed_code, const float* seed_phase, int n_clips, int n_seg,
                              int64_t* codes_out, int32_t* vote_out, float* phase_out, int32_t* status_out,
                              void* stream) {
  QPG_CHECK_ARG(n_clips >= 0 && n_seg >= 0, "negative size");
  if (n_clips == 0 || n_seg == 0) return QPG_OK;
  QPG_CHECK_ARG(entries && code && phase_amp && seed_code && seed_phase && codes_out && vote_out && status_out,
                "null pointer");
  match_walk_kernel<<<n_clips, 64, 0, (cudaStream_t)stream>>>(reinterpret_cast<const Entry*>(entries), code, phase_amp,
                                                              seed_code, seed_phase, n_seg, codes_out, vote_out,
                                                              phase_out, status_out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
