// Table merge, rank transform and the sequential tail of
// CodeKNN.search_code_knn (GestureKNN.py:528-660) on the device.
//
// None of this is bandwidth- or compute-heavy (512-element work per step); it
// lives on the GPU so that a whole batch of clips is matched without a
// host round trip per step.
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int KB = QPG_CODEBOOK_SIZE;
constexpr int WIN = 26;     // windows per database sequence (30 - STEP_SZ)
constexpr int NCODE = 30;   // codes per sequence
constexpr int NFRM = 240;   // phase frames per sequence
constexpr int PC = 16;      // phase(8) | amplitude(8)

__global__ void table_merge_kernel(const Pair* __restrict__ parts, int n_parts, int64_t n, Pair* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Pair best = parts[i];
  for (int p = 1; p < n_parts; ++p) {
    const Pair e = parts[(int64_t)p * n + i];
    if (pair_less(e, best)) best = e;
  }
  out[i] = best;
}

// ranks[q][c] = #{c' : d[c'] < d[c] or (d[c'] == d[c] and c' < c)}
// qflags[q] (optional) = 1 when two NON-EMPTY bins hold exactly the same distance (the reference's rank of those
// bins is whatever NumPy's unstable argsort does with ties)
__global__ void __launch_bounds__(KB) rank512_kernel(const Pair* __restrict__ table, int32_t* __restrict__ ranks,
                                                     int32_t* __restrict__ qflags) {
  __shared__ unsigned long long d[KB];
  __shared__ int s_tie;
  const int q = blockIdx.x, c = threadIdx.x;
  const Pair e = table[(size_t)q * KB + c];
  const unsigned long long mine = e.d;
  const bool nonempty = (long long)e.id >= 0;
  d[c] = mine;
  if (c == 0) s_tie = 0;
  __syncthreads();
  int r = 0, tie = 0;
#pragma unroll 8
  for (int j = 0; j < KB; ++j) {
    const unsigned long long o = d[j];
    r += (o < mine) || (o == mine && j < c);
    tie |= (o == mine && j != c);
  }
  ranks[(size_t)q * KB + c] = r;
  if (qflags) {
    if (tie && nonempty) s_tie = 1;
    __syncthreads();
    if (c == 0) qflags[q] = s_tie;
  }
}

// pull a buffer into L2 (prefetch.global.L2 per 128-byte line): the streaming scans evict the small tables the
// sequential tail reads with dependent loads, so they are re-warmed right before it
__global__ void l2_prefetch_kernel(const char* __restrict__ p, size_t bytes) {
  const size_t lines = (bytes + 127) / 128;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < lines; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i * 128));
}

struct ArgMin {
  double v;
  int i;
};
__device__ __forceinline__ ArgMin argmin_combine(ArgMin a, ArgMin b) {
  return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ ArgMin warp_argmin(ArgMin a) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    ArgMin b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = argmin_combine(a, b);
  }
  return a;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two warps per clip: warp 0 follows the audio table, warp 1 the text table.  Lane l owns codes
// 16*l .. 16*l+15 in registers, so both 512-way arg-mins are register scans + 5 shuffle steps; the two
// warps exchange (phase distance, 4 codes, next phase) through shared memory at two 64-thread barriers
// per step.  Ranks and window ids of step s+1 do not depend on the state and are prefetched during step s.
constexpr int EPL = KB / 32;  // entries per lane

__device__ __forceinline__ void load16(const int32_t* __restrict__ p, int (&v)[EPL]) {
  const int4* q = reinterpret_cast<const int4*>(p);
#pragma unroll
  for (int k = 0; k < EPL / 4; ++k) {
    const int4 t = q[k];
    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
}

__device__ __forceinline__ void bar64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(64)
    match_tail_kernel(const Pair* __restrict__ aud_table, const Pair* __restrict__ txt_table,
                      const int32_t* __restrict__ aud_rank, const int32_t* __restrict__ txt_rank,
                      const int32_t* __restrict__ pos_rank, const int32_t* __restrict__ freq_rank,
                      const int32_t* __restrict__ code, int64_t n_seq, const float* __restrict__ phase_amp,
                      const int32_t* __restrict__ aud_frame, const int32_t* __restrict__ txt_frame,
                      const int32_t* __restrict__ seed_code, const float* __restrict__ seed_phase, int n_seg,
                      int seg_begin, int seg_count, float* __restrict__ state,
                      int64_t* __restrict__ codes_out, int32_t* __restrict__ vote_out,
                      float* __restrict__ phase_out, int32_t* __restrict__ status_out) {
  __shared__ float prev[8 * PC];
  __shared__ double s_dist[2];
  __shared__ int s_code[2][4];
  __shared__ int s_bad[2];

  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // warp 0: audio, 1: text
  const Pair* table = warp == 0 ? aud_table : txt_table;
  const int32_t* rank = warp == 0 ? aud_rank : txt_rank;
  const int32_t* frame = warp == 0 ? aud_frame : txt_frame;

  float* st_b = state ? state + (size_t)b * (8 * PC + 4) : nullptr;
  const bool resume = seg_begin > 0 && st_b != nullptr;
  for (int e = threadIdx.x; e < 8 * PC; e += 64) prev[e] = resume ? st_b[4 + e] : seed_phase[(size_t)b * 8 * PC + e];
  int last = resume ? __float_as_int(st_b[0]) : seed_code[b];
  if (resume && __float_as_int(st_b[1]) != 0) {   // an earlier segment already failed
    if (threadIdx.x == 0) status_out[b] = 1;
    return;
  }
  double freq_term[EPL];
  {
    int fr[EPL];
    load16(freq_rank + lane * EPL, fr);
#pragma unroll
    for (int k = 0; k < EPL; ++k) freq_term[k] = __dmul_rn((double)fr[k], 0.05);
  }
  const int n_steps = (seg_begin + seg_count) * 8;
  const int first_step = seg_begin * 8;
  const size_t q0 = (size_t)b * n_seg * 8;

  int rk_n[EPL];
  long long id_n[EPL];
  {
    const size_t base = (q0 + first_step) * KB + lane * EPL;
    load16(rank + base, rk_n);
#pragma unroll
    for (int k = 0; k < EPL; ++k) id_n[k] = (long long)table[base + k].id;
  }
  __syncthreads();

  int code29 = last;
  for (int st = first_step; st < n_steps; ++st) {
    const int g = st >> 3, s = st & 7;
    const size_t q = q0 + st;
    int pr[EPL], rk[EPL];
    long long id[EPL];
    load16(pos_rank + (size_t)last * KB + lane * EPL, pr);           // depends on the state: round trip 1
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      rk[k] = rk_n[k];
      id[k] = id_n[k];
    }
    if (st + 1 < n_steps) {                                           // state independent: overlaps this step
      const size_t base = (q + 1) * KB + lane * EPL;
      load16(rank + base, rk_n);
#pragma unroll
      for (int k = 0; k < EPL; ++k) id_n[k] = (long long)table[base + k].id;
    }
    // combined = (pos_score + freq_score*0.05) + rank, same IEEE operations as NumPy (GestureKNN.py:545,554,575)
    ArgMin best{1e300, KB};
    long long best_id = -1;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const double v = __dadd_rn(__dadd_rn((double)pr[k], freq_term[k]), (double)rk[k]);
      if (v < best.v) {            // ascending k: the first minimum of this lane wins
        best.v = v;
        best.i = lane * EPL + k;
        best_id = id[k];
      }
    }
    {
      // lexicographic (value, code) arg-min across lanes; carry the window id along
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
        const long long oid = __shfl_xor_sync(0xffffffffu, best_id, o);
        if (ov < best.v || (ov == best.v && oi < best.i)) {
          best.v = ov;
          best.i = oi;
          best_id = oid;
        }
      }
    }
    // phase continuity of this warp's candidate (GestureKNN.py:627-644): round trip 2
    const long long w = best_id;
    const bool bad = w < 0 || w >= n_seq * WIN;
    float tl[4] = {0.f, 0.f, 0.f, 0.f};
    double dist = 0.0;
    if (!bad) {
      const long long j = w / WIN;
      const int m = (int)(w - j * WIN);
      const int f = frame[m];
      const float* head = phase_amp + ((size_t)j * NFRM + f) * PC;   // rows f .. f+7 ; rows f+24 .. f+31 = next prev
      int my_code = 0;
      if (lane < 4) my_code = code[(size_t)j * NCODE + m + lane];
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
      dist = 0.5 * warp_sum(acc);
      if (lane < 4) s_code[warp][lane] = my_code;
    }
    if (lane == 0) {
      s_dist[warp] = dist;
      s_bad[warp] = bad ? 1 : 0;
    }
    bar64(1);                                                         // both candidates scored
    if (s_bad[0] | s_bad[1]) {
      if (threadIdx.x == 0) {
        status_out[b] = 1;
        if (st_b) st_b[1] = __int_as_float(1);
      }
      return;
    }
    const int final_idx = (s_dist[0] <= s_dist[1]) ? 0 : 1;           // tmp_distance.index(min(...)): audio wins ties
    const int p1 = s_code[final_idx][1], p3 = s_code[final_idx][3];
    if (warp == final_idx) {                                          // the winner installs the next phase
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        prev[lane + 32 * k] = tl[k];
        if (phase_out) phase_out[(q * 8) * PC + lane + 32 * k] = tl[k];
      }
      if (lane < 4) {
        const int p = s * 4 + lane;
        if (p < NCODE) codes_out[((size_t)b * n_seg + g) * NCODE + p] = s_code[final_idx][lane];
      }
      if (lane == 0) vote_out[q] = final_idx;
    }
    if (s == 7) code29 = p1;   // produced code #30 seeds the next segment (GestureKNN.py:800)
    last = (s == 7) ? code29 : p3;
    bar64(2);                                                         // prev / s_* may be overwritten
  }
  if (threadIdx.x == 0) status_out[b] = 0;
  if (st_b) {
    if (threadIdx.x == 0) {
      st_b[0] = __int_as_float(last);
      st_b[1] = __int_as_float(0);
    }
    for (int e = threadIdx.x; e < 8 * PC; e += 64) st_b[4 + e] = prev[e];
  }
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_l2_prefetch(const void* ptr, size_t bytes, void* stream) {
  if (bytes == 0) return QPG_OK;
  QPG_CHECK_ARG(ptr != nullptr, "null pointer");
  const size_t lines = (bytes + 127) / 128;
  size_t blocks = (lines + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  l2_prefetch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const char*>(ptr), bytes);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_table_merge(const qpg_pair_t* parts, int n_parts, int64_t n_entries, qpg_pair_t* out,
                               void* stream) {
  QPG_CHECK_ARG(n_parts >= 1 && n_entries >= 0, "n_parts >= 1, n_entries >= 0");
  if (n_entries == 0) return QPG_OK;
  QPG_CHECK_ARG(parts && out, "null pointer");
  table_merge_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const Pair*>(parts), n_parts, n_entries, reinterpret_cast<Pair*>(out));
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_rank512(const qpg_pair_t* table, int Q, int32_t* ranks, void* stream) {
  QPG_CHECK_ARG(Q >= 0, "Q >= 0");
  if (Q == 0) return QPG_OK;
  QPG_CHECK_ARG(table && ranks, "null pointer");
  rank512_kernel<<<Q, KB, 0, (cudaStream_t)stream>>>(reinterpret_cast<const Pair*>(table), ranks, nullptr);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_rank512_ties(const qpg_pair_t* table, int Q, int32_t* ranks, int32_t* qflags, void* stream) {
  QPG_CHECK_ARG(Q >= 0, "Q >= 0");
  if (Q == 0) return QPG_OK;
  QPG_CHECK_ARG(table && ranks && qflags, "null pointer");
  rank512_kernel<<<Q, KB, 0, (cudaStream_t)stream>>>(reinterpret_cast<const Pair*>(table), ranks, qflags);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

static int match_tail_launch(const qpg_pair_t* aud_table, const qpg_pair_t* txt_table, const int32_t* aud_rank,
                             const int32_t* txt_rank, const int32_t* pos_rank, const int32_t* freq_rank,
                             const int32_t* code, int64_t n_seq, const float* phase_amp, const int32_t* aud_frame,
                             const int32_t* txt_frame, const int32_t* seed_code, const float* seed_phase, int n_clips,
                             int n_seg, int seg_begin, int seg_count, float* state, int64_t* codes_out,
                             int32_t* vote_out, float* phase_out, int32_t* status_out, void* stream) {
  QPG_CHECK_ARG(n_clips >= 0 && n_seg >= 0 && n_seq >= 0, "negative size");
  QPG_CHECK_ARG(seg_begin >= 0 && seg_count >= 0 && seg_begin + seg_count <= n_seg, "segment range");
  QPG_CHECK_ARG(seg_begin == 0 || state != nullptr, "resuming needs a state buffer");
  if (n_clips == 0 || seg_count == 0) return QPG_OK;
  QPG_CHECK_ARG(aud_table && txt_table && aud_rank && txt_rank && pos_rank && freq_rank && code && phase_amp &&
                    aud_frame && txt_frame && seed_code && seed_phase && codes_out && vote_out && status_out,
                "null pointer");
  match_tail_kernel<<<n_clips, 64, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const Pair*>(aud_table), reinterpret_cast<const Pair*>(txt_table), aud_rank, txt_rank,
      pos_rank, freq_rank, code, n_seq, phase_amp, aud_frame, txt_frame, seed_code, seed_phase, n_seg, seg_begin,
      seg_count, state, codes_out, vote_out, phase_out, status_out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_match_tail(const qpg_pair_t* aud_table, const qpg_pair_t* txt_table, const int32_t* aud_rank,
                              const int32_t* txt_rank, const int32_t* pos_rank, const int32_t* freq_rank,
                              const int32_t* code, int64_t n_seq, const float* phase_amp, const int32_t* aud_frame,
                              const int32_t* txt_frame, const int32_t* seed_code, const float* seed_phase,
                              int n_clips, int n_seg, int64_t* codes_out, int32_t* vote_out, float* phase_out,
                              int32_t* status_out, void* stream) {
  return match_tail_launch(aud_table, txt_table, aud_rank, txt_rank, pos_rank, freq_rank, code, n_seq, phase_amp,
                           aud_frame, txt_frame, seed_code, seed_phase, n_clips, n_seg, 0, n_seg, nullptr, codes_out,
                           vote_out, phase_out, status_out, stream);
}

extern "C" int qpg_match_tail_segments(const qpg_pair_t* aud_table, const qpg_pair_t* txt_table,
                                       const int32_t* aud_rank, const int32_t* txt_rank, const int32_t* pos_rank,
                                       const int32_t* freq_rank, const int32_t* code, int64_t n_seq,
                                       const float* phase_amp, const int32_t* aud_frame, const int32_t* txt_frame,
                                       const int32_t* seed_code, const float* seed_phase, int n_clips, int n_seg,
                                       int seg_begin, int seg_count, float* state, int64_t* codes_out,
                                       int32_t* vote_out, float* phase_out, int32_t* status_out, void* stream) {
  return match_tail_launch(aud_table, txt_table, aud_rank, txt_rank, pos_rank, freq_rank, code, n_seq, phase_amp,
                           aud_frame, txt_frame, seed_code, seed_phase, n_clips, n_seg, seg_begin, seg_count, state,
                           codes_out, vote_out, phase_out, status_out, stream);
}
