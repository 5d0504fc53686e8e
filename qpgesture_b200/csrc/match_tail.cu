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
__global__ void __launch_bounds__(KB) rank512_kernel(const Pair* __restrict__ table, int32_t* __restrict__ ranks) {
  __shared__ unsigned long long d[KB];
  const int q = blockIdx.x, c = threadIdx.x;
  const unsigned long long mine = table[(size_t)q * KB + c].d;
  d[c] = mine;
  __syncthreads();
  int r = 0;
#pragma unroll 8
  for (int j = 0; j < KB; ++j) {
    const unsigned long long o = d[j];
    r += (o < mine) || (o == mine && j < c);
  }
  ranks[(size_t)q * KB + c] = r;
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

__global__ void __launch_bounds__(KB)
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
  __shared__ ArgMin red_a[KB / 32], red_t[KB / 32];
  __shared__ int s_choice[2];
  __shared__ double s_dist[2];
  __shared__ long long s_win[2];
  __shared__ int s_frame[2];
  __shared__ int s_code[2][4];
  __shared__ float s_tail[2][8 * PC];
  __shared__ int s_fail;

  const int b = blockIdx.x, c = threadIdx.x, lane = c & 31, warp = c >> 5;
  // chained launches (one per segment) hand (last code, previous phase) over through `state`
  float* st_b = state ? state + (size_t)b * (8 * PC + 4) : nullptr;
  const bool resume = seg_begin > 0 && st_b != nullptr;
  if (c < 8 * PC) prev[c] = resume ? st_b[4 + c] : seed_phase[(size_t)b * 8 * PC + c];
  if (c == 0) s_fail = 0;
  int last = resume ? __float_as_int(st_b[0]) : seed_code[b];
  if (resume && __float_as_int(st_b[1]) != 0) {   // an earlier segment already failed
    if (c == 0) status_out[b] = 1;
    return;
  }
  const double freq_term = __dmul_rn((double)freq_rank[c], 0.05);
  const int n_steps = (seg_begin + seg_count) * 8;
  const int first_step = seg_begin * 8;
  const size_t q0 = (size_t)b * n_seg * 8;
  // ranks and window ids of a step do not depend on the sequential state: keep one step in flight
  int ra_n = aud_rank[(q0 + first_step) * KB + c], rt_n = txt_rank[(q0 + first_step) * KB + c];
  long long ida_n = (long long)aud_table[(q0 + first_step) * KB + c].id;
  long long idt_n = (long long)txt_table[(q0 + first_step) * KB + c].id;
  __syncthreads();

  int code29 = last;
  for (int st = first_step; st < n_steps; ++st) {
    const int g = st >> 3, s = st & 7;
    const size_t q = q0 + st;
    const int ra = ra_n, rt = rt_n;
    const long long ida = ida_n, idt = idt_n;
    // pos_score + freq_score*0.05, then + rank (same IEEE operations as NumPy, GestureKNN.py:545,554,575)
    const int pr = pos_rank[(size_t)last * KB + c];
    if (st + 1 < n_steps) {
      ra_n = aud_rank[(q + 1) * KB + c];
      rt_n = txt_rank[(q + 1) * KB + c];
      ida_n = (long long)aud_table[(q + 1) * KB + c].id;
      idt_n = (long long)txt_table[(q + 1) * KB + c].id;
    }
    const double base = __dadd_rn((double)pr, freq_term);
    ArgMin a{__dadd_rn(base, (double)ra), c};
    ArgMin t{__dadd_rn(base, (double)rt), c};
    a = warp_argmin(a);
    t = warp_argmin(t);
    if (lane == 0) {
      red_a[warp] = a;
      red_t[warp] = t;
    }
    __syncthreads();
    {
      // every warp reduces the 16 partials redundantly: no second hand-off through shared memory
      ArgMin x = lane < KB / 32 ? red_a[lane] : ArgMin{1e300, KB};
      ArgMin y = lane < KB / 32 ? red_t[lane] : ArgMin{1e300, KB};
      x = warp_argmin(x);
      y = warp_argmin(y);
      if (c == x.i) s_win[0] = ida;
      if (c == y.i) s_win[1] = idt;
    }
    __syncthreads();
    // warps 0 / 1 score the audio / text candidate by phase continuity (GestureKNN.py:627-644)
    if (warp < 2) {
      const long long w = s_win[warp];
      if (w < 0 || w >= n_seq * WIN) {
        if (lane == 0) s_fail = 1;
      } else {
        const long long j = w / WIN;
        const int m = (int)(w - j * WIN);
        const int f = (warp == 0 ? aud_frame : txt_frame)[m];
        const float* head = phase_amp + ((size_t)j * NFRM + f) * PC;  // rows f .. f+7
        if (lane < 4) s_code[warp][lane] = code[(size_t)j * NCODE + m + lane];
        float tl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) tl[k] = head[24 * PC + lane + 32 * k];   // window frames 24..31 (next prev)
        double av[4], bv[4], sa = 0.0, sb = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e2 = lane + 32 * k, row = e2 >> 4, col = e2 & 15;
          const float fa = row < 5 ? prev[(3 + row) * PC + col] : head[(row - 5) * PC + col];
          const float fb = row < 3 ? prev[(5 + row) * PC + col] : head[(row - 3) * PC + col];
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
        acc = 0.5 * warp_sum(acc);
#pragma unroll
        for (int k = 0; k < 4; ++k) s_tail[warp][lane + 32 * k] = tl[k];
        if (lane == 0) {
          s_dist[warp] = acc;
          s_frame[warp] = f;
        }
      }
    }
    __syncthreads();
    if (s_fail) {
      if (c == 0) {
        status_out[b] = 1;
        if (st_b) st_b[1] = __int_as_float(1);
      }
      return;
    }
    const int final_idx = (s_dist[0] <= s_dist[1]) ? 0 : 1;  // tmp_distance.index(min(...)): audio wins ties
    float new_prev = 0.f;
    if (c < 8 * PC) new_prev = s_tail[final_idx][c];
    const int p1 = s_code[final_idx][1], p3 = s_code[final_idx][3];
    if (c < 4) {
      const int p = s * 4 + c;
      if (p < NCODE) codes_out[((size_t)b * n_seg + g) * NCODE + p] = s_code[final_idx][c];
    }
    if (c == 0) vote_out[q] = final_idx;
    if (s == 7) code29 = p1;  // produced code #30 seeds the next segment (GestureKNN.py:800)
    last = (s == 7) ? code29 : p3;
    __syncthreads();  // everyone has read prev / s_* of this step
    if (c < 8 * PC) {
      prev[c] = new_prev;
      if (phase_out) phase_out[(q * 8) * PC + c] = new_prev;
    }
    __syncthreads();
  }
  if (c == 0) status_out[b] = 0;
  if (st_b) {
    if (c == 0) {
      st_b[0] = __int_as_float(last);
      st_b[1] = __int_as_float(0);
    }
    if (c < 8 * PC) st_b[4 + c] = prev[c];
  }
}

}  // namespace
}  // namespace qpg

using namespace qpg;

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
  rank512_kernel<<<Q, KB, 0, (cudaStream_t)stream>>>(reinterpret_cast<const Pair*>(table), ranks);
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
  match_tail_kernel<<<n_clips, KB, 0, (cudaStream_t)stream>>>(
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
