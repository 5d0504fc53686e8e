// Legacy pose-feature matcher (the `GestureKNN` class, codebook/Speech2GestureMatching/GestureKNN.py:70-284), sm_100a.
//
// Per 8-frame step the reference walks EVERY database sequence: distance of the query feature to each of its frames
// (np.linalg.norm of the 96-d pose part, or sklearn's cosine on the 112-d audio part in the "fake" variant), argsort,
// first frame that is not an exact zero, not within step_sz of the end and whose control mask is set at both ends;
// then rank(pose distance) + rank(audio cosine distance) over the sequences and the desired_k-th candidate wins
// (:135-144).  "185 seqs takes 1h 58min" (:408) on the host.
//
// Here one step of a BATCH of clips is four launches over a float64 copy of the feature table:
//   legacy_frame_cands_kernel  one CTA per (database sequence, clip): all frame distances, the reference's walk as
//                              two lexicographic reductions, and the audio cosine distance of the chosen frame
//   legacy_rank_kernel         rank transforms (argsort().argsort()) as counts of lexicographically smaller elements
//   legacy_pick_kernel         the candidate whose position in the order by (rank sum, sequence) is desired_k (the
//                              host shim can make this one choice with NumPy instead, see the tie note below)
//   legacy_gather_kernel       writes the 8 motion frames and feeds the pose feature of the last one back
// The clips of a batch advance together step by step (the feedback is per clip), so the sequential depth is
// n_frames / 8 steps whatever the batch size.
//
// Ties: NumPy's argsort is unstable, so exact ties are platform defined in the reference - and rank SUMS are small
// integers, so ties at the pick are common (:139).  The device pick is stable (lower sequence first) and sets bit 1
// of the clip's status when the picked position was tied (or any float distance tied exactly); the Python shim's
// default hands the n_seq rank sums of each step to np.argsort on the host instead, which is the reference's order.
// Float64 sums run in index order; np.linalg.norm's BLAS dot may associate differently, so distances can differ in
// the last bit - which changes a result only at such ties.
#include <math.h>

#include "qpg_common.cuh"

namespace qpg {
namespace {

struct LegacyCand {      // per (clip, database sequence)
  double pose_d;         // distance of the chosen frame (L2 of the pose part, or cosine of the audio part)
  double aud_d;          // audio cosine distance of the chosen frame (search_motion only)
  int frame;             // -1: this sequence offers no candidate
  int tie;               // the walk met an exact tie at the chosen distance
};

// sklearn paired cosine: 0.5 * || x/|x| - y/|y| ||^2, rows with a norm below 10*eps stay unscaled (normalize());
// identical vectors give exactly 0 (the reference relies on that, GestureKNN.py:125-126, :184)
__device__ __forceinline__ double cosine_sklearn(const double* __restrict__ x, const double* __restrict__ y, int n) {
  double sx = 0.0, sy = 0.0;
  for (int i = 0; i < n; ++i) {
    sx = fma(x[i], x[i], sx);
    sy = fma(y[i], y[i], sy);
  }
  const double nx = sx > kTinySq ? sqrt(sx) : 1.0, ny = sy > kTinySq ? sqrt(sy) : 1.0;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    const double d = x[i] / nx - y[i] / ny;
    acc = fma(d, d, acc);
  }
  return 0.5 * acc;
}

constexpr int LEG_MAX_FRAMES = 1024;

__global__ void __launch_bounds__(128)
legacy_frame_cands_kernel(const double* __restrict__ feat, const int32_t* __restrict__ mask, int n_frames, int F,
                          const double* __restrict__ query, int q_ld, int lo, int dim, int metric,
                          const double* __restrict__ aud_query, int n_aud, int step_sz, LegacyCand* __restrict__ cands,
                          int n_seq) {
  __shared__ double s_d[LEG_MAX_FRAMES];
  __shared__ double s_q[256];
  __shared__ int s_masksum;
  const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const double* rows = feat + (size_t)k * n_frames * F;
  const int32_t* mk = mask + (size_t)k * n_frames;
  if (tid == 0) s_masksum = 0;
  for (int i = tid; i < dim; i += blockDim.x) s_q[i] = query[(size_t)b * q_ld + i];
  __syncthreads();
  int msum = 0;
  for (int l = tid; l < n_frames; l += blockDim.x) {
    msum += mk[l] != 0;
    const double* x = rows + (size_t)l * F + lo;
    double d;
    if (metric == 0) {                       // np.linalg.norm(query - row)
      double acc = 0.0;
      for (int i = 0; i < dim; ++i) {
        const double t = s_q[i] - x[i];
        acc = fma(t, t, acc);
      }
      d = sqrt(acc);
    } else {
      d = cosine_sklearn(s_q, x, dim);
    }
    s_d[l] = d;
  }
  if (msum) atomicAdd(&s_masksum, msum);
  __syncthreads();
  if (tid >= 32) return;
  // the walk of :173-199 as two reductions by warp 0.  The LAST element of the ascending order is never examined
  // (:177): with a stable sort that is the highest frame among the maxima.
  const int lane = tid;
  double dmax = -1.0;
  int imax = -1;
  for (int l = lane; l < n_frames; l += 32)
    if (s_d[l] > dmax || (s_d[l] == dmax && l > imax)) {
      dmax = s_d[l];
      imax = l;
    }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, dmax, o);
    const int oi = __shfl_xor_sync(0xffffffffu, imax, o);
    if (od > dmax || (od == dmax && oi > imax)) {
      dmax = od;
      imax = oi;
    }
  }
  auto acceptable = [&](int l) {
    return l != imax && s_d[l] != 0.0 && l <= n_frames - step_sz && mk[l] + mk[min(l + step_sz - 1, n_frames - 1)] == 2;
  };
  double best = 1e300;
  int arg = -1;
  for (int l = lane; l < n_frames; l += 32)
    if (acceptable(l) && s_d[l] < best) {       // ascending l: the lowest frame among equal distances stays
      best = s_d[l];
      arg = l;
    }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
    if (oi >= 0 && (arg < 0 || od < best || (od == best && oi < arg))) {
      best = od;
      arg = oi;
    }
  }
  int ties = 0;
  for (int l = lane; l < n_frames; l += 32) ties |= (acceptable(l) && l != arg && s_d[l] == best) ? 1 : 0;
  ties = __any_sync(0xffffffffu, ties) ? 1 : 0;
  // a maximum shared with the excluded last element makes the exclusion itself order dependent
  int max_tie = 0;
  for (int l = lane; l < n_frames; l += 32) max_tie |= (l != imax && s_d[l] == dmax) ? 1 : 0;
  max_tie = __any_sync(0xffffffffu, max_tie);
  if (lane == 0) {
    LegacyCand c;
    c.frame = s_masksum == 0 ? -1 : arg;
    c.pose_d = best;
    c.aud_d = 0.0;
    c.tie = ties | ((max_tie && arg >= 0 && best == dmax) ? 1 : 0);
    if (c.frame >= 0 && aud_query)
      c.aud_d = cosine_sklearn(aud_query + (size_t)b * n_aud, rows + (size_t)c.frame * F, n_aud);
    cands[(size_t)b * n_seq + k] = c;
  }
}

// comb[b][i] = rank of pose_d + rank of aud_d among the sequences that offer a candidate (stable: lower sequence first)
__global__ void __launch_bounds__(256)
legacy_rank_kernel(const LegacyCand* __restrict__ cands, int n_seq, int use_aud, int32_t* __restrict__ comb,
                   int32_t* __restrict__ tie_flag) {
  __shared__ double s_p[256], s_a[256];
  __shared__ int s_f[256];
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  const LegacyCand* cb = cands + (size_t)b * n_seq;
  LegacyCand me;
  me.frame = -1;
  me.pose_d = me.aud_d = 0.0;
  me.tie = 0;
  if (i < n_seq) me = cb[i];
  int rp = 0, ra = 0, tie = 0;
  for (int j0 = 0; j0 < n_seq; j0 += 256) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    if (j < n_seq) {
      const LegacyCand o = cb[j];
      s_p[threadIdx.x] = o.pose_d;
      s_a[threadIdx.x] = o.aud_d;
      s_f[threadIdx.x] = o.frame;
    } else {
      s_f[threadIdx.x] = -1;
    }
    __syncthreads();
    if (me.frame < 0) continue;
    for (int t = 0; t < 256; ++t) {
      if (s_f[t] < 0) continue;
      const int jj = j0 + t;
      if (jj == i) continue;
      rp += (s_p[t] < me.pose_d) || (s_p[t] == me.pose_d && jj < i);
      ra += (s_a[t] < me.aud_d) || (s_a[t] == me.aud_d && jj < i);
      tie |= (s_p[t] == me.pose_d) || (use_aud && s_a[t] == me.aud_d);
    }
  }
  if (i < n_seq) {
    comb[(size_t)b * n_seq + i] = me.frame < 0 ? -1 : rp + (use_aud ? ra : 0);
    tie_flag[(size_t)b * n_seq + i] = me.frame < 0 ? 0 : (tie | me.tie);
  }
}

// chosen[b] = (sequence, frame) of the candidate at position desired_k[b] of the order by (comb, sequence).
// status[b]: bit 0 = fewer than desired_k + 1 candidates (IndexError in the reference, :144), bit 1 = a tie among
// the rank sums or in a rank transform / frame walk that could reorder the winner.
__global__ void __launch_bounds__(256)
legacy_pick_kernel(const LegacyCand* __restrict__ cands, const int32_t* __restrict__ comb,
                   const int32_t* __restrict__ tie_flag, int n_seq, const int32_t* __restrict__ desired_k,
                   int32_t* __restrict__ chosen, int32_t* __restrict__ status, int* __restrict__ n_found) {
  __shared__ int s_c[256], s_t[256];
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  const int32_t* cb = comb + (size_t)b * n_seq;
  const int mine = i < n_seq ? cb[i] : -1;
  int pos = 0, any_tie = 0, found = 0;
  for (int j0 = 0; j0 < n_seq; j0 += 256) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    s_c[threadIdx.x] = j < n_seq ? cb[j] : -1;
    s_t[threadIdx.x] = j < n_seq ? tie_flag[(size_t)b * n_seq + j] : 0;
    __syncthreads();
    for (int t = 0; t < 256; ++t) {
      const int c = s_c[t];
      if (c < 0) continue;
      ++found;
      any_tie |= s_t[t];                      // an exact FLOAT tie in any rank transform can move rank sums around
      const int jj = j0 + t;
      if (mine < 0 || jj == i) continue;
      pos += c < mine || (c == mine && jj < i);
      any_tie |= c == mine;
    }
  }
  if (i == 0) n_found[b] = found;
  if (mine >= 0 && pos == desired_k[b]) {
    chosen[2 * b + 0] = i;
    chosen[2 * b + 1] = cands[(size_t)b * n_seq + i].frame;
    if (any_tie || tie_flag[(size_t)b * n_seq + i]) atomicOr(&status[b], 2);
  }
}

// pred_motion[b][:, j0 : j0 + step] = motion[k, f : f + step, :]^T ; next pose query = feat[k, f + step - 1, n_aud:]
__global__ void __launch_bounds__(256)
legacy_gather_kernel(const double* __restrict__ feat, const double* __restrict__ motion, int n_frames, int F, int J,
                     int n_aud, int n_body, int step_sz, const int32_t* __restrict__ chosen,
                     const int32_t* __restrict__ n_found, const int32_t* __restrict__ desired_k, int j0, int out_frames,
                     double* __restrict__ pred, double* __restrict__ next_pose, int32_t* __restrict__ chosen_log,
                     int step_idx, int n_steps, int32_t* __restrict__ status) {
  const int b = blockIdx.x;
  if (n_found[b] <= desired_k[b] || (status[b] & 1)) {          // IndexError: the clip stops here
    if (threadIdx.x == 0) atomicOr(&status[b], 1);
    return;
  }
  const int k = chosen[2 * b], f = chosen[2 * b + 1];
  if (threadIdx.x == 0 && chosen_log) {
    chosen_log[((size_t)b * n_steps + step_idx) * 2 + 0] = k;
    chosen_log[((size_t)b * n_steps + step_idx) * 2 + 1] = f;
  }
  for (int idx = threadIdx.x; idx < J * step_sz; idx += blockDim.x) {
    const int jn = idx / step_sz, t = idx - jn * step_sz;
    if (j0 + t < out_frames && f + t < n_frames)
      pred[((size_t)b * J + jn) * out_frames + j0 + t] = motion[((size_t)k * n_frames + f + t) * J + jn];
  }
  if (next_pose)
    for (int i = threadIdx.x; i < n_body; i += blockDim.x)
      next_pose[(size_t)b * n_body + i] = feat[((size_t)k * n_frames + min(f + step_sz - 1, n_frames - 1)) * F + n_aud + i];
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" size_t qpg_legacy_cand_bytes(void) { return sizeof(LegacyCand); }

extern "C" int qpg_legacy_candidates(const double* feat, const int32_t* mask, int n_seq, int n_frames, int F, int n_aud,
                                     int n_body, int step_sz, int n_clips, int metric, const double* query, int q_ld,
                                     const double* aud_query, void* cands, int32_t* comb, int32_t* tie_flag,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QPG_CHECK_ARG(feat && mask && query && cands && comb && tie_flag, "null pointer");
  QPG_CHECK_ARG(n_seq > 0 && n_frames > 0 && n_frames <= LEG_MAX_FRAMES && step_sz >= 1 && step_sz <= n_frames,
                "1 <= step_sz <= n_frames <= 1024");
  QPG_CHECK_ARG(n_aud > 0 && n_body >= 0 && F >= n_aud + n_body && n_aud <= 256 && n_body <= 256,
                "feature layout: F >= n_aud + n_body, parts of at most 256 values");
  QPG_CHECK_ARG(metric == 0 || metric == 1, "metric 0 (L2 on the pose part) or 1 (cosine on the audio part)");
  QPG_CHECK_ARG(metric == 1 || aud_query, "metric 0 needs the audio query");
  QPG_CHECK_ARG(n_clips <= 65535, "at most 65535 clips per call");
  if (n_clips <= 0) return QPG_OK;
  const int lo = metric == 0 ? n_aud : 0, dim = metric == 0 ? n_body : n_aud;
  LegacyCand* c = reinterpret_cast<LegacyCand*>(cands);
  legacy_frame_cands_kernel<<<dim3((unsigned)n_seq, (unsigned)n_clips), 128, 0, stream>>>(
      feat, mask, n_frames, F, query, q_ld, lo, dim, metric, metric == 0 ? aud_query : nullptr, n_aud, step_sz, c, n_seq);
  QPG_LAUNCH_CHECK();
  const dim3 g((unsigned)((n_seq + 255) / 256), (unsigned)n_clips);
  legacy_rank_kernel<<<g, 256, 0, stream>>>(c, n_seq, metric == 0 ? 1 : 0, comb, tie_flag);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_legacy_pick(const void* cands, const int32_t* comb, const int32_t* tie_flag, int n_seq, int n_clips,
                               const int32_t* desired_k, int32_t* chosen, int32_t* n_found, int32_t* status,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QPG_CHECK_ARG(cands && comb && tie_flag && desired_k && chosen && n_found && status, "null pointer");
  QPG_CHECK_ARG(n_seq > 0 && n_clips <= 65535, "bad size");
  if (n_clips <= 0) return QPG_OK;
  const dim3 g((unsigned)((n_seq + 255) / 256), (unsigned)n_clips);
  legacy_pick_kernel<<<g, 256, 0, stream>>>(reinterpret_cast<const LegacyCand*>(cands), comb, tie_flag, n_seq, desired_k,
                                            chosen, status, n_found);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_legacy_gather(const double* feat, const double* motion, int n_seq, int n_frames, int F, int J,
                                 int n_aud, int n_body, int step_sz, int n_clips, const int32_t* chosen,
                                 const int32_t* n_found, const int32_t* desired_k, int j0, int out_frames, int step_idx,
                                 int n_steps, double* pred, double* next_pose, int32_t* chosen_log, int32_t* status,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QPG_CHECK_ARG(feat && motion && chosen && n_found && desired_k && pred && status, "null pointer");
  QPG_CHECK_ARG(n_seq > 0 && n_frames > 0 && J > 0 && step_sz >= 1 && F >= n_aud + n_body, "bad shape");
  if (n_clips <= 0) return QPG_OK;
  legacy_gather_kernel<<<(unsigned)n_clips, 256, 0, stream>>>(feat, motion, n_frames, F, J, n_aud, n_body, step_sz, chosen,
                                                             n_found, desired_k, j0, out_frames, pred, next_pose,
                                                             chosen_log, step_idx, n_steps, status);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
