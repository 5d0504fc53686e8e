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

// grid (Q, 16), 256 threads: lane = one of 32 values of `last`, warp = one 64-code slice of the scan
__global__ void __launch_bounds__(256, 6)
    match_lookup_kernel(const Pair* __restrict__ aud_table, const Pair* __restrict__ txt_table,
                        const int32_t* __restrict__ aud_rank, const int32_t* __restrict__ txt_rank,
                        const int16_t* __restrict__ pos_rank_t, const int32_t* __restrict__ freq_rank,
                        const int32_t* __restrict__ code, int64_t n_seq, const int32_t* __restrict__ aud_frame,
                        const int32_t* __restrict__ txt_frame, const int32_t* __restrict__ qflags_a,
                        const int32_t* __restrict__ qflags_t, Entry* __restrict__ entries) {
  constexpr int EMPTY = 1 << 30;
  constexpr int NSL = 8, CPS = KB / NSL;        // slices, codes per slice
  __shared__ int s_key[2][KB];                  // 20*rank + freq_rank per table; empty bins carry the EMPTY bit
  __shared__ int s_fr[KB];
  __shared__ int s_ne[2];
  __shared__ int p_best[NSL][2][32], p_arg[NSL][2][32], p_ties[NSL][2][32], p_bne[NSL][2][32], p_lbe[NSL][2][32];
  const int q = blockIdx.x, tid = threadIdx.x, l = tid & 31, sl = tid >> 5;
  const int last = blockIdx.y * 32 + l;
  int ne_a = 0, ne_t = 0;
  for (int c = tid; c < KB; c += 256) {
    const int fr = freq_rank[c];
    const long long ida = (long long)aud_table[(size_t)q * KB + c].id, idt = (long long)txt_table[(size_t)q * KB + c].id;
    s_fr[c] = fr;
    s_key[0][c] = (20 * aud_rank[(size_t)q * KB + c] + fr) | (ida < 0 ? EMPTY : 0);
    s_key[1][c] = (20 * txt_rank[(size_t)q * KB + c] + fr) | (idt < 0 ? EMPTY : 0);
    ne_a += ida >= 0;
    ne_t += idt >= 0;
  }
  // every thread staged exactly two codes: count the non-empty bins with two block-wide popcounts
  const int na1 = __syncthreads_count(ne_a >= 1), na2 = __syncthreads_count(ne_a >= 2);
  const int nt1 = __syncthreads_count(ne_t >= 1), nt2 = __syncthreads_count(ne_t >= 2);
  if (tid == 0) {
    s_ne[0] = na1 + na2;
    s_ne[1] = nt1 + nt2;
  }
  __syncthreads();
  // integer keys 20*(pos + rank) + freq order exactly like NumPy's float64 (pos + freq*0.05) + rank whenever
  // they differ (the float64 rounding is ~1e-13, distinct keys are >= 0.05 apart).  On EQUAL keys NumPy compares
  // the float64 values, which may differ in the last bits between different (pos, rank) splits: those are
  // evaluated with NumPy's operation order right here; equal float64 values are a true tie (flagged).
  auto f64_value = [&](int pos, int c, int x) {
    const int fr = s_fr[c];
    const int rk = ((s_key[x][c] & ~EMPTY) - fr) / 20;
    return __dadd_rn(__dadd_rn((double)pos, __dmul_rn((double)fr, 0.05)), (double)rk);
  };
  int best[2] = {0x7fffffff, 0x7fffffff}, arg[2] = {0, 0}, bpos[2] = {0, 0}, tie[2] = {0, 0};
  int best_ne[2] = {0x7fffffff, 0x7fffffff};       // best key over non-empty bins
  int lb_e[2] = {0x3fffffff, 0x3fffffff};          // min over empty bins of 20*pos + freq
  auto offer = [&](int x, int key, int c, int pos, int t) {
    if (key < best[x]) {
      best[x] = key;
      arg[x] = c;
      bpos[x] = pos;
      tie[x] = t;
    } else if (key == best[x]) {
      const double vb = f64_value(bpos[x], arg[x], x), vc = f64_value(pos, c, x);
      if (vc < vb) {
        arg[x] = c;
        bpos[x] = pos;
        tie[x] = t;
      } else if (vc == vb) {
        tie[x] = 1;                                // the lower code (offered first) stays
      }
    }
  };
  const bool any_empty = s_ne[0] < KB || s_ne[1] < KB;          // block uniform; false for any sizeable database
  int m1[2] = {0x7fffffff, 0x7fffffff}, m2[2] = {0x7fffffff, 0x7fffffff};
  for (int c0 = sl * CPS; c0 < (sl + 1) * CPS; c0 += 8) {      // eight loads in flight per round
    int pr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pr[i] = (int)pos_rank_t[(size_t)(c0 + i) * KB + last];
    if (!any_empty) {
      // fast path: (key << 9 | code) packed in one int; smallest and second smallest value per table with three
      // min/max each - the smallest carries the lowest code among equal keys
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        const int v = pr[i] * (20 << 9);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int val = v + ((s_key[x][c] << 9) | c);
          const int t = max(m1[x], val);
          m1[x] = min(m1[x], val);
          m2[x] = min(m2[x], t);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        const int p20 = 20 * pr[i];
        const int fr = s_fr[c];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int sk = s_key[x][c];
          const int key = p20 + (sk & ~EMPTY);
          offer(x, key, c, pr[i], 0);
          if (sk & EMPTY) lb_e[x] = min(lb_e[x], p20 + fr);
          else best_ne[x] = min(best_ne[x], key);
        }
      }
    }
  }
  if (!any_empty) {
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int key = m1[x] >> 9, c1 = m1[x] & (KB - 1);
      if (((m1[x] ^ m2[x]) >> 9) != 0) {           // the slice's best key is unique
        best[x] = key;
        arg[x] = c1;
        bpos[x] = (key - s_key[x][c1]) / 20;
        tie[x] = 0;
      } else {                                     // equal keys inside the slice: let the float64 values decide
        for (int c = sl * CPS; c < (sl + 1) * CPS; ++c) {
          const int pos = (int)pos_rank_t[(size_t)c * KB + last];
          if (20 * pos + s_key[x][c] == key) offer(x, key, c, pos, 0);
        }
      }
    }
  }
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    p_best[sl][x][l] = best[x];
    p_arg[sl][x][l] = arg[x];
    p_ties[sl][x][l] = (bpos[x] << 1) | tie[x];
    p_bne[sl][x][l] = best_ne[x];
    p_lbe[sl][x][l] = lb_e[x];
  }
  __syncthreads();
  if (sl != 0) return;
  // combine the slices in ascending code order: the first minimum keeps the arg
  for (int s2 = 1; s2 < NSL; ++s2) {
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int pt = p_ties[s2][x][l];
      offer(x, p_best[s2][x][l], p_arg[s2][x][l], pt >> 1, pt & 1);
      best_ne[x] = min(best_ne[x], p_bne[s2][x][l]);
      lb_e[x] = min(lb_e[x], p_lbe[s2][x][l]);
    }
  }
  Entry e;
  e.flags = ((qflags_a && qflags_a[q]) || (qflags_t && qflags_t[q])) ? 1 : 0;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    if (tie[x]) e.flags |= 1;                      // a true tie at the arg-min: NumPy's order is platform defined
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

// ---- lookup, pruned: one thread per (step, last), codes visited in RANK order --------------------------------
// key(c) = 20*(rank[c] + pos_rank[last][c]) + freq_rank[c] >= 20*rank[c], so once 20*r exceeds the best key found
// among the r lowest-ranked codes nothing further can win or tie: with two (near-)random permutations the minimum
// of rank + pos is ~sqrt(2*512), i.e. ~40 of the 512 codes are visited instead of all of them (8.5 M -> ~1 M warp
// instructions per 48 steps).  All lanes of a warp walk the same code sequence (the order depends on the step
// only), so the pos_rank_t[c][last] loads stay coalesced.  The arg-min rule is match_lookup_kernel's: integer keys,
// float64 comparison with NumPy's operation order on equal keys, the lower code on exact ties (flagged).  Steps
// with empty bins take the full loop in code order (sentinel rule, see match_lookup_kernel).
__global__ void __launch_bounds__(512)
    match_lookup_pruned_kernel(const Pair* __restrict__ aud_table, const Pair* __restrict__ txt_table,
                               const int32_t* __restrict__ aud_rank, const int32_t* __restrict__ txt_rank,
                               const int16_t* __restrict__ pos_rank_t, const int32_t* __restrict__ freq_rank,
                               const int32_t* __restrict__ code, int64_t n_seq, const int32_t* __restrict__ aud_frame,
                               const int32_t* __restrict__ txt_frame, const int32_t* __restrict__ qflags_a,
                               const int32_t* __restrict__ qflags_t, Entry* __restrict__ entries) {
  constexpr int EMPTY = 1 << 30;
  __shared__ int s_key[2][KB];                  // 20*rank + freq_rank per table; empty bins carry the EMPTY bit
  __shared__ int s_fr[KB];
  __shared__ short s_inv[2][KB];                // code by rank
  __shared__ int s_ne[2];
  __shared__ long long s_w[256];                // the text half of the entry, handed to the audio thread
  __shared__ int s_half[256];                   // nl | nls << 9 | frame << 18 | flag << 27
  const int q = blockIdx.x, tid = threadIdx.x;
  // threads 0-255 search the audio table, 256-511 the text table, for the same 256 values of `last`: the two
  // searches are independent chains of ~9 dependent gather batches each, so they run side by side
  const int x = tid >> 8, lt = tid & 255;
  const int last = blockIdx.y * 256 + lt;
  s_inv[0][tid] = 0;                            // ranks are a permutation of 0..511 when they come from this library;
  s_inv[1][tid] = 0;                            // anything else must still index inside the tables
  __syncthreads();
  {
    const int c = tid;                          // 512 threads stage one code each
    const int fr = freq_rank[c];
    const int ra = aud_rank[(size_t)q * KB + c], rt = txt_rank[(size_t)q * KB + c];
    const long long ida = (long long)aud_table[(size_t)q * KB + c].id, idt = (long long)txt_table[(size_t)q * KB + c].id;
    s_fr[c] = fr;
    s_key[0][c] = (20 * ra + fr) | (ida < 0 ? EMPTY : 0);
    s_key[1][c] = (20 * rt + fr) | (idt < 0 ? EMPTY : 0);
    s_inv[0][ra & (KB - 1)] = (short)c;
    s_inv[1][rt & (KB - 1)] = (short)c;
    const int na = __syncthreads_count(ida >= 0), nt = __syncthreads_count(idt >= 0);
    if (tid == 0) {
      s_ne[0] = na;
      s_ne[1] = nt;
    }
  }
  __syncthreads();
  auto f64_value = [&](int pos, int c) {
    const int fr = s_fr[c];
    const int rk = ((s_key[x][c] & ~EMPTY) - fr) / 20;
    return __dadd_rn(__dadd_rn((double)pos, __dmul_rn((double)fr, 0.05)), (double)rk);
  };
  int best = 0x7fffffff, arg = 0, bpos = 0, tie = 0;
  int best_ne = 0x7fffffff;                        // best key over non-empty bins
  int lb_e = 0x3fffffff;                           // min over empty bins of 20*pos + freq
  auto offer = [&](int key, int c, int pos) {      // any visiting order: the lower code stays on exact ties
    if (key < best) {
      best = key;
      arg = c;
      bpos = pos;
      tie = 0;
    } else if (key == best) {
      const double vb = f64_value(bpos, arg), vc = f64_value(pos, c);
      if (vc < vb) {
        arg = c;
        bpos = pos;
        tie = 0;
      } else if (vc == vb) {
        tie = 1;
        if (c < arg) {
          arg = c;
          bpos = pos;
        }
      }
    }
  };
  if (s_ne[x] == KB) {                             // warp uniform (x is): no empty bin in this table
    for (int r0 = 0; r0 < KB && 20 * r0 <= best; r0 += 8) {       // eight loads in flight per round trip
      int cc[8], pr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        cc[i] = s_inv[x][r0 + i];
        pr[i] = (int)pos_rank_t[(size_t)cc[i] * KB + last];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) offer(20 * pr[i] + s_key[x][cc[i]], cc[i], pr[i]);
    }
  } else {                                         // empty bins: every code, in code order (sentinel rule)
    for (int c0 = 0; c0 < KB; c0 += 8) {
      int pr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pr[i] = (int)pos_rank_t[(size_t)(c0 + i) * KB + last];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i, p20 = 20 * pr[i], sk = s_key[x][c];
        const int key = p20 + (sk & ~EMPTY);
        offer(key, c, pr[i]);
        if (sk & EMPTY) lb_e = min(lb_e, p20 + s_fr[c]);
        else best_ne = min(best_ne, key);
      }
    }
  }
  // this thread's half of the entry
  int flag = tie;                                  // a true tie at the arg-min: NumPy's order is platform defined
  {
    // empty bins all hold the sentinel 1e3, their mutual rank order is NumPy's business: the lowest rank any of
    // them can get is the number of non-empty bins.  If that could reach the best non-empty key, say so.
    const int ne = s_ne[x];
    if (ne < KB && lb_e + 20 * ne <= best_ne) flag = 1;
  }
  const Pair* table = x == 0 ? aud_table : txt_table;
  const long long w = (long long)table[(size_t)q * KB + arg].id;
  long long ew = -1;
  int nl = 0, nls = 0, frame = 0;
  if (w >= 0 && w < n_seq * WIN) {
    const long long j = w / WIN;
    const int m = (int)(w - j * WIN);
    ew = w;
    nl = code[(size_t)j * NCODE + m + 3];
    nls = code[(size_t)j * NCODE + m + 1];
    frame = (x == 0 ? aud_frame : txt_frame)[m];
  }
  if (x == 1) {
    s_w[lt] = ew;
    s_half[lt] = (nl & 511) | ((nls & 511) << 9) | ((frame & 511) << 18) | (flag << 27);
  }
  __syncthreads();
  if (x == 0) {
    const int h = s_half[lt];
    Entry e;
    e.flags = (((qflags_a && qflags_a[q]) || (qflags_t && qflags_t[q])) ? 1 : 0) | flag | ((h >> 27) & 1);
    e.w[0] = ew;
    e.w[1] = s_w[lt];
    e.nl[0] = (short)nl;
    e.nl[1] = (short)(h & 511);
    e.nls[0] = (short)nls;
    e.nls[1] = (short)((h >> 9) & 511);
    e.frame[0] = (short)frame;
    e.frame[1] = (short)((h >> 18) & 511);
    entries[(size_t)q * KB + last] = e;
  }
}

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

// The pick of GestureKNN.py:627-646 by ONE warp: lanes 0-15 score the audio candidate, lanes 16-31 the text
// candidate; lane (hl = lane & 15) owns column hl of the 8 x 16 phase|amplitude block:
//   a = [prev[3..7]; head[0..2]],  b = [prev[5..7]; head[0..4]]   (rows; :636,:644)
//   dist = 0.5 * || a/|a| - b/|b| ||^2   (sklearn paired cosine distance, float64, all-zero vector -> norm 1)
// prev5[k] = prev[3 + k][hl] (k < 5) is passed in registers, head rows come from `head` (row stride PC).
// Fixed summation order (8 terms per lane, then a 16-lane butterfly); returns 0 (audio, also on ties) or 1.
__device__ __forceinline__ int phase_pick(const float (&prev5)[5], const float* __restrict__ head, int lane) {
  const int hl = lane & 15;
  float h[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) h[k] = head[k * PC + hl];
  // float32 filter: both distances with ~1e-5 absolute error (vectors of 128 normalised components); decided
  // when they are further apart than 1e-3.  Float64 division and square root cost ~20x more on this chip.
  {
    float sa = 0.f, sb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float a = k < 5 ? prev5[k] : h[k - 5], b = k < 3 ? prev5[k + 2] : h[k - 3];
      sa = fmaf(a, a, sa);
      sb = fmaf(b, b, sb);
      ab = fmaf(a, b, ab);
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
      ab += __shfl_xor_sync(0xffffffffu, ab, o);
    }
    // 0.5*|a^ - b^|^2 = 0.5*(|a^|^2 + |b^|^2) - a^.b^  with |v^| = 1 unless v == 0 (then v^ = 0)
    const float ia = sa > 0.f ? rsqrtf(sa) : 0.f, ib = sb > 0.f ? rsqrtf(sb) : 0.f;
    const float d32 = 0.5f * ((sa > 0.f ? 1.f : 0.f) + (sb > 0.f ? 1.f : 0.f)) - ab * ia * ib;
    const float da = __shfl_sync(0xffffffffu, d32, 0), dt = __shfl_sync(0xffffffffu, d32, 16);
    // magnitudes: the float32 sums are only trusted when no component over/underflowed their squares
    const bool sane = sa < 1e30f && sb < 1e30f && (sa == 0.f || sa > 1e-30f) && (sb == 0.f || sb > 1e-30f);
    const bool all_sane = __all_sync(0xffffffffu, sane);
    if (all_sane && fabsf(da - dt) > 1e-3f) return da < dt ? 0 : 1;
  }
  double av[8], bv[8], sa = 0.0, sb = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    av[k] = (double)(k < 5 ? prev5[k] : h[k - 5]);
    bv[k] = (double)(k < 3 ? prev5[k + 2] : h[k - 3]);
    sa = fma(av[k], av[k], sa);
    sb = fma(bv[k], bv[k], sb);
  }
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  const double na = sa > 0.0 ? sqrt(sa) : 1.0, nb = sb > 0.0 ? sqrt(sb) : 1.0;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double d = av[k] / na - bv[k] / nb;
    acc = fma(d, d, acc);
  }
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const double dist = 0.5 * acc;
  const double d_a = __shfl_sync(0xffffffffu, dist, 0), d_t = __shfl_sync(0xffffffffu, dist, 16);
  return d_a <= d_t ? 0 : 1;                                  // tmp_distance.index(min(...)): audio wins ties (:646)
}

// ---- direct walk: one warp per clip, two dependent loads per step (many clips: latency hidden across clips) ----
__global__ void __launch_bounds__(32)
    match_walk_kernel(const Entry* __restrict__ entries, const int32_t* __restrict__ code,
                      const float* __restrict__ phase_amp, const int32_t* __restrict__ seed_code,
                      const float* __restrict__ seed_phase, int n_seg, int64_t* __restrict__ codes_out,
                      int32_t* __restrict__ vote_out, float* __restrict__ phase_out, int32_t* __restrict__ status_out) {
  const int b = blockIdx.x, lane = threadIdx.x, hw = lane >> 4, hl = lane & 15;
  const int n_steps = n_seg * 8;
  const size_t q0 = (size_t)b * n_steps;
  for (int i = lane; i < n_seg * NCODE; i += 32) codes_out[(size_t)b * n_seg * NCODE + i] = -1;
  const int last0 = seed_code[b];
  if ((unsigned)last0 >= (unsigned)KB) {
    if (lane == 0) status_out[b] = 1;
    return;
  }
  float prev5[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) prev5[k] = seed_phase[(size_t)b * 8 * PC + (3 + k) * PC + hl];
  Entry cur = load_entry(entries + q0 * KB + last0);
  int status = 0;
  __syncwarp();
  for (int st = 0; st < n_steps; ++st) {
    const int g = st >> 3, s = st & 7;
    const size_t q = q0 + st;
    status |= (cur.flags & 1) ? 2 : 0;
    if (cur.w[0] < 0 || cur.w[1] < 0) {             // IndexError at GestureKNN.py:631
      status |= 1;
      break;
    }
    Entry nx0 = cur, nx1 = cur;                      // both possible next entries, fetched while the pick is computed
    if (st + 1 < n_steps) {
      const int l0 = s == 7 ? cur.nls[0] : cur.nl[0], l1 = s == 7 ? cur.nls[1] : cur.nl[1];
      nx0 = load_entry(entries + (q + 1) * KB + l0);
      nx1 = load_entry(entries + (q + 1) * KB + l1);
    }
    const long long w = hw == 0 ? cur.w[0] : cur.w[1];
    const long long j = w / WIN;
    const int f = hw == 0 ? cur.frame[0] : cur.frame[1];
    const float* head = phase_amp + ((size_t)j * NFRM + f) * PC;      // rows f..f+7; rows f+24..f+31 = next prev
    float tl[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) tl[k] = head[(24 + k) * PC + hl];
    const int win = phase_pick(prev5, head, lane);
    // the winner's tail rows become prev: lanes of the other half fetch them by shuffle
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = __shfl_sync(0xffffffffu, tl[k], win * 16 + hl);
      if (k >= 3) prev5[k - 3] = t;
      if (phase_out && hw == 0) phase_out[(q * 8 + k) * PC + hl] = t;
    }
    const long long ww = win == 0 ? cur.w[0] : cur.w[1];
    const long long jw = ww / WIN;
    const int mw = (int)(ww - jw * WIN);
    if (lane < 4) {
      const int p = s * 4 + lane;
      if (p < NCODE) codes_out[((size_t)b * n_seg + g) * NCODE + p] = code[(size_t)jw * NCODE + mw + lane];
    }
    if (lane == 0) vote_out[q] = win;
    cur = win == 0 ? nx0 : nx1;
  }
  if (lane == 0) status_out[b] = status;
}

// ---- transitions: the pick for EVERY reachable state, in parallel (few clips: nothing left on the chain) ----
// state entering step st = (last, which) of the winner of step st-1, i.e. window entries[q-1][last].w[which]
// (step 0: the seed, state 0).  trans[q][state] = next state | tie flag << 10;  -1 = IndexError at this step
// (-3: and the choice that led there was tie dependent), -2 = unreachable.  One warp per (step, state).
__global__ void __launch_bounds__(256, 6)
    match_transition_kernel(const Entry* __restrict__ entries, const float* __restrict__ phase_amp,
                            const int32_t* __restrict__ seed_code, const float* __restrict__ seed_phase, int n_steps,
                            long long n_warps, int16_t* __restrict__ trans) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= n_warps) return;
  const int lane = threadIdx.x & 31, hw = lane >> 4, hl = lane & 15;
  const long long q = wid >> 10;
  const int state = (int)(wid & 1023);
  const int st = (int)(q % n_steps);
  const long long b = q / n_steps;
  int last;
  float prev5[5];
  if (st == 0) {
    if (state != 0) {
      if (lane == 0) trans[wid] = -2;
      return;
    }
    last = seed_code[b];
    if ((unsigned)last >= (unsigned)KB) {
      if (lane == 0) trans[wid] = -1;
      return;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) prev5[k] = seed_phase[(size_t)b * 8 * PC + (3 + k) * PC + hl];
  } else {
    const Entry ep = load_entry(entries + (q - 1) * KB + (state >> 1));
    const int xp = state & 1;
    const long long wp = xp == 0 ? ep.w[0] : ep.w[1];
    if (ep.w[0] < 0 || ep.w[1] < 0) {                // the previous step raised: nothing continues from it
      if (lane == 0) trans[wid] = -2;
      return;
    }
    last = (st & 7) == 0 ? (xp == 0 ? ep.nls[0] : ep.nls[1]) : (xp == 0 ? ep.nl[0] : ep.nl[1]);
    const long long jp = wp / WIN;
    const int fp = xp == 0 ? ep.frame[0] : ep.frame[1];
    const float* tailp = phase_amp + ((size_t)jp * NFRM + fp + 24) * PC;
#pragma unroll
    for (int k = 0; k < 5; ++k) prev5[k] = tailp[(3 + k) * PC + hl];
  }
  const Entry e = load_entry(entries + q * KB + last);
  if (e.w[0] < 0 || e.w[1] < 0) {
    if (lane == 0) trans[wid] = (e.flags & 1) ? -3 : -1;       // IndexError (tie dependent: -3)
    return;
  }
  const long long w = hw == 0 ? e.w[0] : e.w[1];
  const long long j = w / WIN;
  const int f = hw == 0 ? e.frame[0] : e.frame[1];
  const int win = phase_pick(prev5, phase_amp + ((size_t)j * NFRM + f) * PC, lane);
  if (lane == 0) trans[wid] = (int16_t)((last << 1) | win | ((e.flags & 1) << 10));
}

// one CTA per clip: the clip's transition table goes to shared memory, one thread follows it, then all threads
// write the outputs of the visited states
// ---- transitions from per-window phase statistics: four states per warp ------------------------------------
// The two vectors of the pick are a = [prev rows 3..7; head rows 0..2] and b = [prev rows 5..7; head rows 0..4]
// (prev = rows 24..31 behind the previous winner's phase frame, head = rows behind the candidate's), so
//   |a|^2 = T5(prev) + H3(cand),  |b|^2 = T3(prev) + H5(cand),
//   a.b   = TT(prev) + [prev row 6 . head row 0 + prev row 7 . head row 1] + HH(cand)
// where T5, T3, TT depend on the previous window only and H3, H5, HH on the candidate only.  qpg_phase_stats
// tabulates them once per (window, table) together with the four rows of the cross term (72 floats), so the
// float32 filter of a state is 2 x 32 multiply-adds and two scalar loads: eight lanes per state instead of a warp,
// no phase_amp access, no window-id division.  States whose two distances are within 1e-3 (or whose sums are not
// trustworthy in float32) take the float64 pick of match_transition_kernel, a warp at a time.
constexpr int PSTAT = 72;                 // floats per (window, table): 8 scalars, head rows 0-1, tail rows 30-31

__global__ void __launch_bounds__(256)
    phase_stats_kernel(const float* __restrict__ phase_amp, long long n_windows, const int32_t* __restrict__ aud_frame,
                       const int32_t* __restrict__ txt_frame, float* __restrict__ stats) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;   // (window, table)
  if (wid >= n_windows * 2) return;
  const int lane = threadIdx.x & 31, hl = lane & 15, x = (int)(wid & 1);
  const long long w = wid >> 1, j = w / WIN;
  const int m = (int)(w - j * WIN);
  const int f = (x == 0 ? aud_frame : txt_frame)[m];
  const float* rows = phase_amp + ((size_t)j * NFRM + f) * PC;
  float* out = stats + (size_t)wid * PSTAT;
  // lanes 0-15: head rows 0..4, lanes 16-31: tail rows 27..31 (= prev rows 3..7); lane hl owns column hl
  const int r0 = lane < 16 ? 0 : 27;
  float v[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) v[k] = (f + r0 + k < NFRM) ? rows[(r0 + k) * PC + hl] : 0.f;
  // head: H3 = |rows 0..2|^2, H5 = |rows 0..4|^2, HH = r0.r2 + r1.r3 + r2.r4
  // tail: T5 = |rows 27..31|^2, T3 = |rows 29..31|^2, TT = r27.r29 + r28.r30 + r29.r31
  float s3, s5, sd;
  if (lane < 16) {
    s3 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    s5 = s3 + v[3] * v[3] + v[4] * v[4];
  } else {
    s3 = v[2] * v[2] + v[3] * v[3] + v[4] * v[4];
    s5 = s3 + v[0] * v[0] + v[1] * v[1];
  }
  sd = v[0] * v[2] + v[1] * v[3] + v[2] * v[4];
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    s5 += __shfl_xor_sync(0xffffffffu, s5, o);
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
  }
  if (lane == 0) {
    out[0] = s3;
    out[1] = s5;
    out[2] = sd;
    out[3] = 0.f;
  }
  if (lane == 16) {
    out[4] = s5;      // T5
    out[5] = s3;      // T3
    out[6] = sd;      // TT
    out[7] = 0.f;
  }
  if (lane < 16) {
    out[8 + hl] = v[0];
    out[24 + hl] = v[1];
  } else {
    out[40 + hl] = v[3];      // row 30 = prev row 6
    out[56 + hl] = v[4];      // row 31 = prev row 7
  }
}

__global__ void __launch_bounds__(256, 4)
    match_transition_fast_kernel(const Entry* __restrict__ entries, const float* __restrict__ stats,
                                 const float* __restrict__ phase_amp, const int32_t* __restrict__ seed_code,
                                 const float* __restrict__ seed_phase, int n_steps, long long n_states,
                                 int16_t* __restrict__ trans) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
  const long long sid = wid * 4 + grp;                 // (step, state) of this group of eight lanes
  if (wid * 4 >= n_states) return;
  const bool live = sid < n_states;
  const long long q = (live ? sid : n_states - 1) >> 10;
  const int state = (int)((live ? sid : n_states - 1) & 1023);
  const int st = (int)(q % n_steps);
  const long long b = q / n_steps;
  // result: >= 0 decided, -1/-2/-3 as match_transition_kernel, kSlow = needs the float64 pick
  constexpr int kSlow = -100;
  int result = kSlow, last = 0, eflag = 0, fp = 0, f0 = 0, f1 = 0;
  long long wp = -1, w0 = -1, w1 = -1;
  if (st == 0) {
    if (state != 0) result = -2;
    else {
      last = seed_code[b];
      if ((unsigned)last >= (unsigned)KB) result = -1;
    }
  } else {
    const Entry ep = load_entry(entries + (q - 1) * KB + (state >> 1));
    const int xp = state & 1;
    if (ep.w[0] < 0 || ep.w[1] < 0) result = -2;      // the previous step raised: nothing continues from it
    else {
      wp = xp == 0 ? ep.w[0] : ep.w[1];
      fp = xp == 0 ? ep.frame[0] : ep.frame[1];
      last = (st & 7) == 0 ? (xp == 0 ? ep.nls[0] : ep.nls[1]) : (xp == 0 ? ep.nl[0] : ep.nl[1]);
    }
  }
  if (result == kSlow) {
    const Entry e = load_entry(entries + q * KB + last);
    eflag = e.flags & 1;
    if (e.w[0] < 0 || e.w[1] < 0) result = eflag ? -3 : -1;       // IndexError (tie dependent: -3)
    else {
      w0 = e.w[0];
      w1 = e.w[1];
      f0 = e.frame[0];
      f1 = e.frame[1];
    }
  }
  {
    // float32 filter from the tables: lanes 0-3 of the group score the audio candidate, lanes 4-7 the text one.
    // Executed by every lane (groups without a live pick read window 0) so that the shuffles stay warp-uniform.
    const bool use = result == kSlow && st != 0;
    const int x = gl >> 2, c4 = (gl & 3) * 4;
    const float* P = stats + ((size_t)(use ? wp : 0) * 2 + (state & 1)) * PSTAT;
    const float* C = stats + ((size_t)(use ? (x == 0 ? w0 : w1) : 0) * 2 + x) * PSTAT;
    const float4 t6 = __ldg(reinterpret_cast<const float4*>(P + 40 + c4)), t7 = __ldg(reinterpret_cast<const float4*>(P + 56 + c4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(C + 8 + c4)), h1 = __ldg(reinterpret_cast<const float4*>(C + 24 + c4));
    const float4 ps = __ldg(reinterpret_cast<const float4*>(P + 4)), cs = __ldg(reinterpret_cast<const float4*>(C));
    float xs = t6.x * h0.x;
    xs = fmaf(t6.y, h0.y, xs);
    xs = fmaf(t6.z, h0.z, xs);
    xs = fmaf(t6.w, h0.w, xs);
    xs = fmaf(t7.x, h1.x, xs);
    xs = fmaf(t7.y, h1.y, xs);
    xs = fmaf(t7.z, h1.z, xs);
    xs = fmaf(t7.w, h1.w, xs);
    xs += __shfl_xor_sync(0xffffffffu, xs, 1);
    xs += __shfl_xor_sync(0xffffffffu, xs, 2);
    const float sa = ps.x + cs.x, sb = ps.y + cs.y, ab = ps.z + xs + cs.z;
    const float ia = sa > 0.f ? rsqrtf(sa) : 0.f, ib = sb > 0.f ? rsqrtf(sb) : 0.f;
    const float d32 = 0.5f * ((sa > 0.f ? 1.f : 0.f) + (sb > 0.f ? 1.f : 0.f)) - ab * ia * ib;
    const int sane = (sa < 1e30f && sb < 1e30f && (sa == 0.f || sa > 1e-30f) && (sb == 0.f || sb > 1e-30f)) ? 1 : 0;
    const float da = __shfl_sync(0xffffffffu, d32, lane & ~7), dt = __shfl_sync(0xffffffffu, d32, (lane & ~7) + 4);
    const int sane_a = __shfl_sync(0xffffffffu, sane, lane & ~7), sane_t = __shfl_sync(0xffffffffu, sane, (lane & ~7) + 4);
    if (use && sane_a && sane_t && fabsf(da - dt) > 1e-3f) result = (last << 1) | (da < dt ? 0 : 1) | (eflag << 10);
  }
  // undecided states: the float64 pick, one state at a time by the whole warp
  unsigned need = __ballot_sync(0xffffffffu, live && result == kSlow && gl == 0);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    const int s_st = __shfl_sync(0xffffffffu, st, src), s_last = __shfl_sync(0xffffffffu, last, src);
    const int s_flag = __shfl_sync(0xffffffffu, eflag, src);
    const long long s_b = __shfl_sync(0xffffffffu, b, src), s_wp = __shfl_sync(0xffffffffu, wp, src);
    const long long s_w0 = __shfl_sync(0xffffffffu, w0, src), s_w1 = __shfl_sync(0xffffffffu, w1, src);
    const int s_fp = __shfl_sync(0xffffffffu, fp, src), s_f0 = __shfl_sync(0xffffffffu, f0, src);
    const int s_f1 = __shfl_sync(0xffffffffu, f1, src);
    const int hw = lane >> 4, hl = lane & 15;
    float prev5[5];
    if (s_st == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) prev5[k] = seed_phase[(size_t)s_b * 8 * PC + (3 + k) * PC + hl];
    } else {
      const float* tailp = phase_amp + ((size_t)(s_wp / WIN) * NFRM + s_fp + 24) * PC;
#pragma unroll
      for (int k = 0; k < 5; ++k) prev5[k] = tailp[(3 + k) * PC + hl];
    }
    const long long w = hw == 0 ? s_w0 : s_w1;
    const int f = hw == 0 ? s_f0 : s_f1;
    const int win = phase_pick(prev5, phase_amp + ((size_t)(w / WIN) * NFRM + f) * PC, lane);
    if (lane == src) result = (s_last << 1) | win | (s_flag << 10);
  }
  if (live && gl == 0) trans[sid] = (int16_t)result;
}

constexpr int WALK_MAX_STEPS = 104;                  // 104 * 2 KiB = 208 KiB of shared memory
__global__ void __launch_bounds__(256)
    match_table_walk_kernel(const int16_t* __restrict__ trans, const Entry* __restrict__ entries,
                            const int32_t* __restrict__ code, const float* __restrict__ phase_amp, int n_seg,
                            int64_t* __restrict__ codes_out, int32_t* __restrict__ vote_out,
                            float* __restrict__ phase_out, int32_t* __restrict__ status_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int16_t* tab = reinterpret_cast<int16_t*>(smem_raw);                     // [n_steps][1024]
  __shared__ int16_t s_state[WALK_MAX_STEPS];                              // (last << 1 | win) per visited step
  __shared__ int s_done, s_status;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n_steps = n_seg * 8;
  const size_t q0 = (size_t)b * n_steps;
  {
    const int4* src = reinterpret_cast<const int4*>(trans + q0 * 1024);
    int4* dst = reinterpret_cast<int4*>(tab);
    for (int i = tid; i < n_steps * 128; i += 256) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < n_seg * NCODE; i += 256) codes_out[(size_t)b * n_seg * NCODE + i] = -1;
  __syncthreads();
  if (tid == 0) {
    int state = 0, status = 0, st = 0;
    for (; st < n_steps; ++st) {
      const int v = tab[st * 1024 + state];
      if (v < 0) {
        status |= v == -3 ? 3 : 1;
        break;
      }
      status |= (v >> 10) & 1 ? 2 : 0;
      state = v & 1023;
      s_state[st] = (int16_t)state;
    }
    s_done = st;
    s_status = status;
  }
  __syncthreads();
  const int done = s_done;
  // outputs: 4 codes + vote + (optionally) the 8 x 16 phase block per visited step
  for (int i = tid; i < done * 4; i += 256) {
    const int st = i >> 2, k = i & 3, s = st & 7, g = st >> 3;
    const int state = s_state[st];
    const Entry e = load_entry(entries + (q0 + st) * KB + (state >> 1));
    const long long w = (state & 1) == 0 ? e.w[0] : e.w[1];
    const long long j = w / WIN;
    const int m = (int)(w - j * WIN);
    const int p = s * 4 + k;
    if (p < NCODE) codes_out[((size_t)b * n_seg + g) * NCODE + p] = code[(size_t)j * NCODE + m + k];
    if (k == 0) vote_out[q0 + st] = state & 1;
  }
  if (phase_out) {
    for (int i = tid; i < done * 8 * PC; i += 256) {
      const int st = i / (8 * PC), r = i - st * 8 * PC;
      const int state = s_state[st];
      const Entry e = load_entry(entries + (q0 + st) * KB + (state >> 1));
      const long long w = (state & 1) == 0 ? e.w[0] : e.w[1];
      const int f = (state & 1) == 0 ? e.frame[0] : e.frame[1];
      phase_out[(q0 + st) * 8 * PC + r] = phase_amp[((size_t)(w / WIN) * NFRM + f + 24) * PC + r];
    }
  }
  if (tid == 0) status_out[b] = s_status;
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
  match_lookup_pruned_kernel<<<dim3((unsigned)Q, 2), 512, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const Pair*>(aud_table), reinterpret_cast<const Pair*>(txt_table), aud_rank, txt_rank, pos_rank_t,
      freq_rank, code, n_seq, aud_frame, txt_frame, qflags_a, qflags_t, reinterpret_cast<Entry*>(entries));
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

static int walk_impl(const void* entries, const int32_t* code, const float* phase_amp, const float* phase_stats,
                     const int32_t* seed_code, const float* seed_phase, int n_clips, int n_seg, int16_t* trans,
                     int64_t* codes_out, int32_t* vote_out, float* phase_out, int32_t* status_out, void* stream) {
  QPG_CHECK_ARG(n_clips >= 0 && n_seg >= 0, "negative size");
  if (n_clips == 0 || n_seg == 0) return QPG_OK;
  QPG_CHECK_ARG(entries && code && phase_amp && seed_code && seed_phase && codes_out && vote_out && status_out,
                "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_steps = n_seg * 8;
  if (trans != nullptr && n_steps <= WALK_MAX_STEPS) {
    QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(trans) & 15) == 0, "trans must be 16-byte aligned");
    const long long n_states = (long long)n_clips * n_steps * 1024;
    if (phase_stats) {
      QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(phase_stats) & 15) == 0, "phase_stats must be 16-byte aligned");
      const long long n_warps = (n_states + 3) / 4;
      match_transition_fast_kernel<<<(unsigned)((n_warps * 32 + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const Entry*>(entries), phase_stats, phase_amp, seed_code, seed_phase, n_steps, n_states, trans);
    } else {
      match_transition_kernel<<<(unsigned)((n_states * 32 + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const Entry*>(entries), phase_amp, seed_code, seed_phase, n_steps, n_states, trans);
    }
    QPG_LAUNCH_CHECK();
    const size_t smem = (size_t)n_steps * 2048;
    QPG_CUDA(cudaFuncSetAttribute(match_table_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_table_walk_kernel<<<n_clips, 256, smem, st>>>(trans, reinterpret_cast<const Entry*>(entries), code, phase_amp,
                                                        n_seg, codes_out, vote_out, phase_out, status_out);
    QPG_LAUNCH_CHECK();
    return QPG_OK;
  }
  match_walk_kernel<<<n_clips, 32, 0, st>>>(reinterpret_cast<const Entry*>(entries), code, phase_amp, seed_code,
                                            seed_phase, n_seg, codes_out, vote_out, phase_out, status_out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_match_walk(const void* entries, const int32_t* code, const float* phase_amp,
                              const int32_t* seed_code, const float* seed_phase, int n_clips, int n_seg,
                              int16_t* trans, int64_t* codes_out, int32_t* vote_out, float* phase_out,
                              int32_t* status_out, void* stream) {
  return walk_impl(entries, code, phase_amp, nullptr, seed_code, seed_phase, n_clips, n_seg, trans, codes_out, vote_out,
                   phase_out, status_out, stream);
}

extern "C" int qpg_match_walk_stats(const void* entries, const int32_t* code, const float* phase_amp,
                                    const float* phase_stats, const int32_t* seed_code, const float* seed_phase,
                                    int n_clips, int n_seg, int16_t* trans, int64_t* codes_out, int32_t* vote_out,
                                    float* phase_out, int32_t* status_out, void* stream) {
  return walk_impl(entries, code, phase_amp, phase_stats, seed_code, seed_phase, n_clips, n_seg, trans, codes_out,
                   vote_out, phase_out, status_out, stream);
}

extern "C" size_t qpg_phase_stats_floats(void) { return PSTAT; }

extern "C" int qpg_phase_stats(const float* phase_amp, int64_t n_seq, const int32_t* aud_frame, const int32_t* txt_frame,
                               float* stats, void* stream) {
  QPG_CHECK_ARG(n_seq >= 0, "negative size");
  if (n_seq == 0) return QPG_OK;
  QPG_CHECK_ARG(phase_amp && aud_frame && txt_frame && stats, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(stats) & 15) == 0, "stats must be 16-byte aligned");
  const long long n_windows = (long long)n_seq * WIN;
  phase_stats_kernel<<<(unsigned)((n_windows * 2 * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      phase_amp, n_windows, aud_frame, txt_frame, stats);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
