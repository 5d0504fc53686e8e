// One-pass candidate scan for up to 64 query steps: exact integer dot products on the tensor cores
// (tcgen05.mma kind::i8, int32 accumulators in TMEM) over an int8-"sliced" copy of the window table,
// plus the interval logic that turns them into the exact (distance, window id) tables of
// CodeKNN.search_audio_cands / search_text_cands (GestureKNN.py:666-691, :708-721).
//
// Why: one query step is 0.5 flop/B (HBM bound), but a 24-s clip is 48 steps that all read the same
// table.  48 steps per pass are 24 flop/B - past the FFMA and far past the FP64 ridge - so the only way
// to stay on the HBM roofline with ONE pass per clip is the tensor pipe.  Floating-point tensor-core
// accumulation has no documented rounding, so the table is stored as integers instead:
//
//   x' = x * 2^-colexp[k]                       (optional per-column power of two, exact)
//   X  = rint(x' * 2^(30-ex)),  |x'| < 2^ex     (31-bit fixed point relative to the row maximum)
//   X  = d0*2^24 + d1*2^16 + d2*2^8 + d3        (balanced base-256 digits = four int8 slices)
//
// and the queries likewise (Y, digits e_t).  The kernel accumulates the ten digit products with
// s+t <= 3, P_j = sum_{s+t=j} sum_k d_s[k] e_t[k], EXACTLY (integers), so
//   dot(x,q) = 2^(ex+eq-36) * v + R,   v = P0*2^24 + P1*2^16 + P2*2^8 + P3  (int64),
// (issued as four instructions per 32-column k-step: slice s of the rows against the stacked query slices
//  0..3-s, so that products of equal weight s+t accumulate in the same TMEM columns)
//   |R| <= 2^(ex+eq-60) * (L1(X)/2 + L1(Y)/2 + K/4 + 128*65793*sum_k(|e1|+|e2|+|e3|))
// (quantisation |dX|,|dY| <= 1/2 plus the six dropped products, |d_s| <= 128).  That is a cosine error
// of ~1e-7 on Gaussian data with a rigorous, per-(row,query) bound: every distance is an interval
// [lo, hi], a bin's winner is decided when only one interval can hold the minimum, the rank transform is
// decided when bin intervals do not overlap, and the few undecided (query, bin) pairs (~0.15 %) are
// re-evaluated in float64 from the float32 table (`exact_distance`, same arithmetic as cand_cosine.cu).
// Integer partial sums are associative, so the stream-K split over CTAs with 64-bit global atomics is
// bit-reproducible.
//
// Layouts (all tiles are pre-swizzled images of what the UMMA descriptor expects: K-major, 128-byte rows,
// SWIZZLE_128B, 8-row groups 1024 bytes apart - one plain cp.async.bulk per tile, no tensor map):
//   database slices  [RT = ceil(W/128)][NKB = ceil(D/128)][4 slices][128 rows x 128 B]   (rows in bin order)
//   query slices     [NKB][4 slices][n_pad rows x 128 B],  n_pad in {16,32,48,64}
//   sacc             int64 [n_pad][Wpad]   v per (query, sorted row position)
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int NS = 4;                 // int8 slices per value
constexpr int TM = 128;               // rows per tile (UMMA M, TMEM lanes)
constexpr int KBW = 128;              // columns (bytes) per k-block = one swizzle row
constexpr int SLICE_BYTES = TM * KBW; // 16 KiB
constexpr int STAGES = 2;
constexpr int MAX_NPAD = 64;
constexpr int KB = QPG_CODEBOOK_SIZE;
constexpr double DROP_C = 128.0 * 65793.0;   // bound of the dropped digit products per unit of sum(|e1|+|e2|+|e3|)
constexpr double EPS_SLACK = 1e-12;          // float64 rounding of both evaluations (<= 5e-13 at D = 6144)

// byte offset of element (row r, column byte kbyte) inside a [rows x 128 B] SWIZZLE_128B tile
__host__ __device__ __forceinline__ uint32_t swz_offset(uint32_t r, uint32_t kbyte) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((kbyte >> 4) ^ (r & 7u)) & 7u) << 4) + (kbyte & 15u);
}

// ------------------------------------------------------------------ slicing (database rows and queries)
struct RowInfo {   // per sorted database row
  double r1;       // 2^(ex-60) / |x|        (0 for an all-zero row)
  double r2;       // 0.5 * L1(X) * r1
};

// One CTA per output row `pos`.  DB mode (n_rows_tile = 128): tile (pos/128, kb, s); query mode: one tile
// column of n_pad rows, row = pos.  sign = -1 applies 2^-colexp (database), +1 applies 2^+colexp (queries).
struct SliceJob {
  const float* rows;
  const int8_t* col_exp;
  int8_t* out;
  qpg_qinfo_t* q_info;
  int64_t ld;
  int D, nkb;
};
struct SliceJobs {
  SliceJob j[2];
};

// 2^e as a double, |e| < 1000 (exact; multiplying by it is exact unless the result is subnormal)
__device__ __forceinline__ double pow2(int e) { return __hiloint2double((1023 + e) << 20, 0); }

constexpr int SLICE_THREADS = 256;
template <bool kQuery>
__global__ void __launch_bounds__(SLICE_THREADS)
    slice_kernel(const SliceJobs jobs, const int32_t* __restrict__ order, int n_pad,
                 const double* __restrict__ sqnorm_in, RowInfo* __restrict__ row_info) {
  constexpr int NW = SLICE_THREADS / 32;
  const SliceJob job = blockIdx.y == 0 ? jobs.j[0] : jobs.j[1];
  const float* __restrict__ rows = job.rows;
  const int8_t* __restrict__ col_exp = job.col_exp;
  int8_t* __restrict__ out = job.out;
  qpg_qinfo_t* __restrict__ q_info = job.q_info;
  const int64_t ld = job.ld;
  const int D = job.D, nkb = job.nkb;
  __shared__ double s_red[NW];
  __shared__ unsigned long long s_l1[NW], s_el[NW];
  __shared__ int s_ex;
  const int64_t pos = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t src = order ? (int64_t)order[pos] : pos;
  const float* x = rows + src * ld;
  const int Dp = nkb * KBW;

  // pass 1: maximum magnitude (after the column scaling) and, for queries, the squared norm
  double mx = 0.0, sq = 0.0;
  for (int k = tid; k < D; k += SLICE_THREADS) {
    double v = (double)x[k];
    sq = fma(v, v, sq);
    if (col_exp) v *= pow2(kQuery ? (int)col_exp[k] : -(int)col_exp[k]);
    mx = fmax(mx, fabs(v));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    double m = s_red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = fmax(m, s_red[w]);
    s_ex = m > 0.0 ? ilogb(m) + 1 : 0;         // |x'| < 2^ex
  }
  __syncthreads();
  const int ex = s_ex;
  if (kQuery) {   // squared norm in a fixed order (thread-strided partials, shuffle tree, warps in order)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) s_red[warp] = sq;           // s_red reuse is safe: everyone passed the barrier above
  }

  // pass 2: digits.  Thread t owns columns 4t..4t+3 of every 1024-column stripe -> one 32-bit store per slice
  unsigned long long l1 = 0;
  unsigned int el = 0;
  const int64_t tile_row = kQuery ? pos : (pos % TM);
  const int64_t rt = kQuery ? 0 : pos / TM;
  const double scale = pow2(30 - ex);           // ex in [-148, 129] for float32 data
  for (int k0 = 4 * tid; k0 < Dp; k0 += 4 * SLICE_THREADS) {
    uint32_t w[NS] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + i;
      int X = 0;
      if (k < D) {
        double v = (double)x[k];
        if (col_exp) v *= pow2(kQuery ? (int)col_exp[k] : -(int)col_exp[k]);
        X = __double2int_rn(v * scale);          // |X| <= 2^30
      }
      l1 += (unsigned long long)(X < 0 ? -X : X);
      int r = X;
      int dig[NS];
#pragma unroll
      for (int s = NS - 1; s >= 1; --s) {
        const int d = ((r + 128) & 255) - 128;
        dig[s] = d;
        r = (r - d) >> 8;
        el += (unsigned int)(d < 0 ? -d : d);
      }
      dig[0] = r;                                // |r| <= 64 because |X| <= 2^30
#pragma unroll
      for (int s = 0; s < NS; ++s) w[s] |= ((uint32_t)dig[s] & 0xffu) << (8 * i);
    }
    const int kb = k0 / KBW, kbyte = k0 % KBW;
    const size_t tile_bytes = kQuery ? (size_t)n_pad * KBW : (size_t)SLICE_BYTES;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const size_t base = (((size_t)rt * nkb + kb) * NS + s) * tile_bytes;
      *reinterpret_cast<uint32_t*>(out + base + swz_offset((uint32_t)tile_row, (uint32_t)kbyte)) = w[s];
    }
  }
  unsigned long long el64 = el;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    el64 += __shfl_xor_sync(0xffffffffu, el64, o);
  }
  if (lane == 0) {
    s_l1[warp] = l1;
    s_el[warp] = el64;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long t1 = 0, t2 = 0;
    double sqq = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      t1 += s_l1[w];
      t2 += s_el[w];
      if (kQuery) sqq += s_red[w];
    }
    const double L1 = (double)t1, EL = (double)t2;
    if (kQuery) {
      qpg_qinfo_t qi;
      qi.sq = sqq;
      qi.g = sqq > kTinySq ? pow2(ex) / sqrt(sqq) : 0.0;
      qi.h = 0.5 * L1 + DROP_C * EL + 0.25 * (double)D;
      qi.ex = ex;
      qi.pad = 0;
      q_info[pos] = qi;
    } else {
      const double sqx = sqnorm_in[src];
      RowInfo ri;
      ri.r1 = sqx > kTinySq ? pow2(ex - 60) / sqrt(sqx) : 0.0;
      ri.r2 = 0.5 * L1 * ri.r1;
      row_info[pos] = ri;
    }
  }
}

// ------------------------------------------------------------------ tcgen05 / TMEM wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_128B operand tile with 128-byte rows (same descriptor as conv1d_tc.cu, proven on B200)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct SegParam {
  const int8_t* A;            // database slices of this feature block
  const int8_t* B;            // query slices
  unsigned long long* sacc;   // [n_pad][Wpad]
  int nkb;
};
struct ScanParams {
  SegParam seg[2];
  int nseg, nkb_total, n_pad, nq;
  long long RT, W, Wpad, total_units;
};

// a "run" = consecutive units of one (row tile, feature block): one TMEM accumulator stage
struct UnitPos {
  long long rt;
  int seg, kb;
};
__device__ __forceinline__ UnitPos unit_pos(const ScanParams& p, long long u) {
  UnitPos r;
  r.rt = u / p.nkb_total;
  const int kbu = (int)(u - r.rt * p.nkb_total);
  r.seg = kbu < p.seg[0].nkb ? 0 : 1;
  r.kb = r.seg == 0 ? kbu : kbu - p.seg[0].nkb;
  return r;
}

__global__ void __launch_bounds__(256, 1) sliced_scan_kernel(const ScanParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t b_slice_bytes = (uint32_t)p.n_pad * KBW;
  const uint32_t stage_bytes = NS * SLICE_BYTES + NS * b_slice_bytes;       // 88 KiB at 48 query steps
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u0 = (long long)blockIdx.x * p.total_units / gridDim.x;
  const long long u1 = (long long)(blockIdx.x + 1) * p.total_units / gridDim.x;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);       // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer: one bulk copy per slice tile, query slices in one copy =====
    if (lane == 0) {
      int it = 0;
      for (long long u = u0; u < u1; ++u, ++it) {
        const UnitPos up = unit_pos(p, u);
        const int8_t* seg_a = up.seg == 0 ? p.seg[0].A : p.seg[1].A;
        const int8_t* seg_b = up.seg == 0 ? p.seg[0].B : p.seg[1].B;
        const int seg_nkb = up.seg == 0 ? p.seg[0].nkb : p.seg[1].nkb;
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], NS * SLICE_BYTES + NS * b_slice_bytes);
        const int8_t* a = seg_a + ((size_t)up.rt * seg_nkb + up.kb) * (size_t)(NS * SLICE_BYTES);
#pragma unroll
        for (int sl = 0; sl < NS; ++sl)
          bulk_g2s(st + sl * SLICE_BYTES, a + (size_t)sl * SLICE_BYTES, SLICE_BYTES, &full[s]);
        bulk_g2s(st + NS * SLICE_BYTES, seg_b + (size_t)up.kb * NS * b_slice_bytes, NS * b_slice_bytes, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: 4 k-steps x 4 MMAs per unit =====
    // The ten digit products with s + t <= 3 are issued as FOUR instructions per k-step: A slice s against the
    // query slices 0..3-s stacked along N (they are contiguous 8-row-aligned tiles in shared memory, so
    // [B_0; ..; B_{3-s}] is itself a valid K-major SWIZZLE_128B operand with N = n_pad*(4-s) rows), written at
    // column offset s*n_pad: product (s, t) lands in column block s + t, i.e. products of equal weight share an
    // accumulator, and every A slice is fetched from shared memory once per k-step instead of up to four times
    // (operand fetch, not the tensor pipe, was the limiter: 96 cycles per N = 48 instruction measured).
    if (lane == 0) {
      // D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TM >> 4) << 24);
      int it = 0, run = 0;
      bool first = true;
      for (long long u = u0; u < u1; ++u, ++it) {
        const UnitPos up = unit_pos(p, u);
        const int as = run & 1;
        if (first && run >= 2) {
          mbar_wait(&acc_empty[as], ((run >> 1) - 1) & 1);
          tc_fence_after();
        }
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + NS * SLICE_BYTES;
        const uint32_t acc0 = tmem_base + (uint32_t)(as * NS * p.n_pad);
        const uint64_t b_desc0 = umma_desc_sw128(b_addr);
#pragma unroll
        for (int k4 = 0; k4 < KBW / 32; ++k4) {
#pragma unroll
          for (int sa = 0; sa < NS; ++sa) {
            const uint32_t n_rows = (uint32_t)(p.n_pad * (NS - sa));
            const uint64_t a_desc = umma_desc_sw128(a_addr + sa * SLICE_BYTES) + 2u * k4;
            tc_mma_i8(acc0 + (uint32_t)(sa * p.n_pad), a_desc, b_desc0 + 2u * k4, idesc0 | ((n_rows >> 3) << 17),
                      (first && k4 == 0 && sa == 0) ? 0u : 1u);
          }
        }
        tc_commit(&empty[s]);
        first = false;
        bool run_ends = (u + 1 == u1);
        if (!run_ends) {
          const UnitPos nx = unit_pos(p, u + 1);
          run_ends = nx.rt != up.rt || nx.seg != up.seg;
        }
        if (run_ends) {
          tc_commit(&acc_full[as]);
          ++run;
          first = true;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> int64 v -> global accumulate =====
    const int quarter = warp & 3;
    int run = 0;
    long long u = u0;
    while (u < u1) {
      const UnitPos up = unit_pos(p, u);
      // extent of this run inside [u0, u1)
      const long long seg_first = up.rt * p.nkb_total + (up.seg == 0 ? 0 : p.seg[0].nkb);
      const long long seg_end = seg_first + (up.seg == 0 ? p.seg[0].nkb : p.seg[1].nkb);
      const long long run_end = seg_end < u1 ? seg_end : u1;
      const bool whole = (u == seg_first) && (run_end == seg_end);     // this CTA owns the full K range
      const int as = run & 1;
      mbar_wait(&acc_full[as], (run >> 1) & 1);
      tc_fence_after();
      const long long row = up.rt * TM + quarter * 32 + lane;
      unsigned long long* dst = (up.seg == 0 ? p.seg[0].sacc : p.seg[1].sacc) + row;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * NS * p.n_pad);
      for (int qc = 0; qc < p.n_pad; qc += 16) {
        uint32_t a0[16], a1[16], a2[16], a3[16];
        tmem_ld_x16(tbase + (uint32_t)(0 * p.n_pad + qc), a0);
        tmem_ld_x16(tbase + (uint32_t)(1 * p.n_pad + qc), a1);
        tmem_ld_x16(tbase + (uint32_t)(2 * p.n_pad + qc), a2);
        tmem_ld_x16(tbase + (uint32_t)(3 * p.n_pad + qc), a3);
        tmem_ld_wait();
        if (row < p.W) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (qc + i < p.nq) {
              const long long v = ((((long long)(int)a0[i] * 256 + (long long)(int)a1[i]) * 256 + (long long)(int)a2[i]) * 256) +
                                  (long long)(int)a3[i];
              unsigned long long* d = dst + (size_t)(qc + i) * p.Wpad;
              if (whole) *d = (unsigned long long)v;
              else red_add_u64(d, (unsigned long long)v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      ++run;
      u = run_end;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Plain CUDA-core evaluation of the same integer sums from the same tile images (test / debug only):
// validates the tile layout and the tensor-core kernel independently of each other.
__global__ void sliced_scan_ref_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int nkb, int n_pad,
                                       int nq, int q_stride, long long W, long long Wpad,
                                       long long* __restrict__ sacc) {
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int q = blockIdx.y * q_stride;
  if (row >= W || q >= nq) return;
  const long long rt = row / TM;
  const uint32_t r = (uint32_t)(row % TM);
  long long P[NS] = {0, 0, 0, 0};
  for (int kb = 0; kb < nkb; ++kb) {
    for (int kbyte = 0; kbyte < KBW; ++kbyte) {
      int d[NS], e[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        d[s] = A[(((size_t)rt * nkb + kb) * NS + s) * SLICE_BYTES + swz_offset(r, kbyte)];
        e[s] = B[((size_t)kb * NS + s) * (size_t)n_pad * KBW + swz_offset((uint32_t)q, kbyte)];
      }
#pragma unroll
      for (int j = 0; j < NS; ++j)
#pragma unroll
        for (int sa = 0; sa <= j; ++sa) P[j] += (long long)(d[sa] * e[j - sa]);
    }
  }
  sacc[(size_t)q * Wpad + row] = ((P[0] * 256 + P[1]) * 256 + P[2]) * 256 + P[3];
}

// ------------------------------------------------------------------ exact float64 re-evaluation
// distance of query `q` (float32 [D], squared norm sqq) to row w of a float32 table in the 4 KiB tile
// layout of qpg_pack_rows_f32; warp-cooperative, fixed summation order, result on every lane.
__device__ __noinline__ double exact_distance(const float* __restrict__ packed, int NC, int64_t w,
                                                 const float* __restrict__ q, int D, double sqq, double sqx, int lane) {
  const float* base = packed + (((size_t)(w >> 3) * NC) * 8 + (w & 7)) * 128;
  double acc = 0.0;
  // chunks that lie completely inside the D columns and can be read as float4 (row pointer 16-byte aligned)
  const int full = (reinterpret_cast<uintptr_t>(q) & 15) == 0 ? D / 128 : 0;
  int c = 0;
  // eight chunks of loads in flight per round trip (the row is read once, from HBM): the FMA order is the
  // plain ascending one, so the result does not depend on the unrolling
  for (; c + 8 <= full; c += 8) {
    float4 xv[8], qv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      xv[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(c + u) * 1024 + 4 * lane));
      qv[u] = __ldg(reinterpret_cast<const float4*>(q + (c + u) * 128 + 4 * lane));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc = fma((double)xv[u].x, (double)qv[u].x, acc);
      acc = fma((double)xv[u].y, (double)qv[u].y, acc);
      acc = fma((double)xv[u].z, (double)qv[u].z, acc);
      acc = fma((double)xv[u].w, (double)qv[u].w, acc);
    }
  }
  for (; c < NC; ++c) {
    const float4 xv = *reinterpret_cast<const float4*>(base + (size_t)c * 1024 + 4 * lane);
    const int k = c * 128 + 4 * lane;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < full) qv = *reinterpret_cast<const float4*>(q + k);
    else {
      if (k < D) qv.x = q[k];
      if (k + 1 < D) qv.y = q[k + 1];
      if (k + 2 < D) qv.z = q[k + 2];
      if (k + 3 < D) qv.w = q[k + 3];
    }
    acc = fma((double)xv.x, (double)qv.x, acc);
    acc = fma((double)xv.y, (double)qv.y, acc);
    acc = fma((double)xv.z, (double)qv.z, acc);
    acc = fma((double)xv.w, (double)qv.w, acc);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const double a = sqq > kTinySq ? 1.0 : 0.0, b = sqx > kTinySq ? 1.0 : 0.0;
  double cs = 0.0;
  if (sqq > kTinySq && sqx > kTinySq) cs = acc / (sqrt(sqq) * sqrt(sqx));
  const double d = 0.5 * (a + b) - cs;
  return d < 0.0 ? 0.0 : d;
}

struct Interval {
  double lo, hi;
};
struct QConst {          // what filter_interval needs of a query: 2^eq/|q| and the bound terms premultiplied
  double g, gh, g1;      // g, g*h*(1+1e-9), g*(1+1e-9)
};
__device__ __forceinline__ QConst make_qconst(const qpg_qinfo_t& qi) {
  QConst c;
  c.g = qi.g;
  c.gh = qi.g * qi.h * (1.0 + 1e-9);
  c.g1 = qi.g * (1.0 + 1e-9);
  return c;
}
// distance interval of (query, sorted row) from the exact integer v
__device__ __forceinline__ Interval filter_interval(long long v, const RowInfo ri, const QConst& qc) {
  const double a = qc.g > 0.0 ? 1.0 : 0.0, b = ri.r1 > 0.0 ? 1.0 : 0.0;
  const double c = (double)v * (ri.r1 * 16777216.0) * qc.g;                 // 2^(ex+eq-36) v / (|x||q|)
  const double eps = fma(ri.r1, qc.gh, ri.r2 * qc.g1) + EPS_SLACK;          // >= g*(r1*h + r2), see header
  const double d = 0.5 * (a + b) - c;
  Interval iv;
  iv.lo = d - eps;
  iv.hi = d + eps;
  if (iv.lo < 0.0) iv.lo = 0.0;
  if (iv.hi < 0.0) iv.hi = 0.0;
  return iv;
}

// ------------------------------------------------------------------ per-bin records
struct TableParams {         // device copy of qpg_sliced_table_t
  const float* packed;
  const double* sqnorm;
  const float* q;
  const qpg_qinfo_t* q_info;
  long long ldq;
  int D, NC;
  long long* sacc;
  const int32_t* bin_start;
  const RowInfo* row_info;
  const int32_t* order;
  qpg_bin_t* bins;
  long long bins_qstride;   // records between two queries (512 when the table's records are contiguous)
  Pair* table;
  int32_t* ranks;
  int32_t* qflags;
};
struct TablePair {
  TableParams t[2];
};

// One warp per (table, start code, group of BG queries): U = min hi over the bin's rows; candidates = rows whose
// interval reaches below U.  One candidate -> record (lo, hi, id).  Several -> float64 re-evaluation of exactly
// those rows (lexicographic (d, id) minimum) -> exact record lo = hi = d.  With `consume` the entries of sacc are
// zeroed once read, so the next pass needs no memset.
//
// The kernel is latency bound (three dependent round trips: bin bounds -> rows -> record), so it is organised for
// work per round trip and residency: the BG = 8 sacc loads of a lane are issued together, a bin of up to 32 rows
// (nearly all of them at 26 windows per sequence) is settled from registers in ONE trip over its rows, the
// per-query constants sit in shared memory, and the 64-bit minimum over the warp is two REDUX instead of ten
// shuffles.  9.3 M warp instructions instead of 13.7 M (ncu), 80 registers instead of 128; 17 us alone (first version
// 28.8) once the 33-64-row bins took the one-trip path too.
constexpr int BG = 8;

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32);
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? (unsigned)v : 0xffffffffu);
  return ((unsigned long long)mh << 32) | ml;
}

// several candidates in one bin: exact distances of exactly those rows; `m` = ballot of the candidate lanes of the
// 32 rows starting at `base`.  Updates the running lexicographic (d, id) minimum.
struct BestPair {
  double d;
  long long id;
};
__device__ __noinline__ BestPair bins_verify(const float* __restrict__ packed, const double* __restrict__ sqnorm,
                                             const int32_t* __restrict__ order, int NC, int D, unsigned m, int base,
                                             const float* qrow, double sqq, int64_t id_offset, int64_t row_base,
                                             int lane, BestPair best) {
  while (m) {
    const int src_lane = __ffs(m) - 1;
    m &= m - 1;
    const long long w = order[base + src_lane];
    const double d = exact_distance(packed, NC, w + row_base, qrow, D, sqq, sqnorm[w + row_base], lane);
    const long long id = id_offset + w;
    if (d < best.d || (d == best.d && id < best.id)) {
      best.d = d;
      best.id = id;
    }
  }
  return best;
}

// T refers to kernel-parameter space with a STATIC index (see the kernel below): its fields are read from the
// constant bank where they are used instead of occupying ~26 registers for the whole kernel
__device__ __forceinline__ void bins_body(const TableParams& T, QConst* s_qc, long long* s_v, long long W, long long Wpad, int nq,
                                          int64_t id_offset, int64_t row_base, int consume,
                                          unsigned long long* __restrict__ stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.y * BG, c = blockIdx.x * 8 + warp;
  if (threadIdx.x < BG) s_qc[threadIdx.x] = make_qconst(T.q_info[min(q0 + (int)threadIdx.x, nq - 1)]);
  const int b0 = T.bin_start[c], b1 = T.bin_start[c + 1];
  __syncthreads();
  const int n = b1 - b0;
  const int ng = min(BG, nq - q0);
  const unsigned long long kInf = ~0ull;
  if (n <= 64) {
    // ---- the whole bin in one trip: lane = rows `lane` and `lane + 32`.  The BG accumulators of a row go global ->
    // shared by cp.async (sixteen copies in flight per lane without holding 32 registers), then one query at a time.
    // (Bins of 33-64 rows used to take the sequential path below: a tenth of the bins at 26 windows per sequence,
    // and 17 % of the kernel's stall samples - each of those warps ran ~30 dependent round trips.)
    const bool valid0 = lane < n, valid1 = lane + 32 < n;
    const int pos0 = b0 + lane, pos1 = pos0 + 32;
    RowInfo ri0, ri1;
    ri0.r1 = ri0.r2 = ri1.r1 = ri1.r2 = 0.0;
    long long* my_v = s_v + (warp * BG) * 64 + lane;
    if (valid0) {
#pragma unroll
      for (int g = 0; g < BG; ++g) {
        const long long* src = T.sacc + (size_t)min(q0 + g, nq - 1) * Wpad + pos0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(my_v + g * 64)), "l"(src) : "memory");
      }
      ri0 = T.row_info[pos0];
    }
    if (valid1) {
#pragma unroll
      for (int g = 0; g < BG; ++g) {
        const long long* src = T.sacc + (size_t)min(q0 + g, nq - 1) * Wpad + pos1;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(my_v + g * 64 + 32)), "l"(src) : "memory");
      }
      ri1 = T.row_info[pos1];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
      const int qi_ = q0 + g;
      qpg_bin_t rec;
      rec.lo = kEmptyDist;
      rec.hi = kEmptyDist;
      rec.id = -1;
      rec.n = 0;
      rec.flags = 1;                             // empty bins are exact (sentinel)
      if (n > 0) {
        Interval iv0, iv1;
        iv0.lo = iv0.hi = iv1.lo = iv1.hi = 0.0;
        if (valid0) {
          iv0 = filter_interval(my_v[g * 64], ri0, s_qc[g]);
          if (consume) T.sacc[(size_t)qi_ * Wpad + pos0] = 0;
        }
        if (valid1) {
          iv1 = filter_interval(my_v[g * 64 + 32], ri1, s_qc[g]);
          if (consume) T.sacc[(size_t)qi_ * Wpad + pos1] = 0;
        }
        // non-negative doubles order like their bit patterns
        const unsigned long long h0 = valid0 ? (unsigned long long)__double_as_longlong(iv0.hi) : kInf;
        const unsigned long long h1 = valid1 ? (unsigned long long)__double_as_longlong(iv1.hi) : kInf;
        const unsigned long long U = warp_min_u64(h0 < h1 ? h0 : h1);
        const bool cand0 = valid0 && (unsigned long long)__double_as_longlong(iv0.lo) <= U;
        const bool cand1 = valid1 && (unsigned long long)__double_as_longlong(iv1.lo) <= U;
        const unsigned m0 = __ballot_sync(0xffffffffu, cand0), m1 = __ballot_sync(0xffffffffu, cand1);
        const int cnt = __popc(m0) + __popc(m1);
        if (cnt == 1) {
          const int src = __ffs(m0 | m1) - 1;
          rec.lo = __shfl_sync(0xffffffffu, m0 ? iv0.lo : iv1.lo, src);
          rec.hi = __longlong_as_double((long long)U);
          rec.id = id_offset + T.order[b0 + src + (m0 ? 0 : 32)];
          rec.n = 1;
          rec.flags = 0;
        } else {
          BestPair best;
          best.d = 1e300;
          best.id = -1;
          const float* qrow = T.q + (size_t)qi_ * T.ldq;
          const double sqq = T.q_info[qi_].sq;
          if (m0) best = bins_verify(T.packed, T.sqnorm, T.order, T.NC, T.D, m0, b0, qrow, sqq, id_offset, row_base, lane, best);
          if (m1) best = bins_verify(T.packed, T.sqnorm, T.order, T.NC, T.D, m1, b0 + 32, qrow, sqq, id_offset, row_base, lane, best);
          rec.lo = best.d;
          rec.hi = best.d;
          rec.id = best.id;
          rec.n = cnt;                           // how many rows were re-evaluated (diagnostics)
          rec.flags = 1;                         // exact
          if (lane == 0 && stats) atomicAdd(&stats[0], (unsigned long long)rec.n);
        }
      }
      if (lane == 0) T.bins[(size_t)qi_ * T.bins_qstride + c] = rec;
    }
  } else {
    // ---- more than 64 rows in a bin of a small table (rare): two trips per query, 32 rows at a time ----
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
      const int qi_ = q0 + g;
      long long* sv = T.sacc + (size_t)qi_ * Wpad;
      const QConst qg = s_qc[g];
      unsigned long long Ul = kInf;
      for (int pos = b0 + lane; pos < b1; pos += 32)
        Ul = min(Ul, (unsigned long long)__double_as_longlong(filter_interval(sv[pos], T.row_info[pos], qg).hi));
      const unsigned long long U = warp_min_u64(Ul);
      int cnt = 0;
      double best_lo = 0.0;
      BestPair best;
      best.d = 1e300;
      best.id = -1;
      long long single_pos = -1;
      const float* qrow = T.q + (size_t)qi_ * T.ldq;
      const double sqq = T.q_info[qi_].sq;
      for (int base = b0; base < b1; base += 32) {
        const int pos = base + lane;
        bool cand = false;
        double lo = 0.0;
        if (pos < b1) {
          lo = filter_interval(sv[pos], T.row_info[pos], qg).lo;
          cand = (unsigned long long)__double_as_longlong(lo) <= U;
          if (consume) sv[pos] = 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        const int k = __popc(m);
        if (k == 0) continue;
        if (cnt == 0 && k == 1) {                // remember the first lone candidate; verified only if another shows up
          const int src = __ffs(m) - 1;
          single_pos = base + src;
          best_lo = __shfl_sync(0xffffffffu, lo, src);
          cnt = 1;
          continue;
        }
        if (single_pos >= 0) {                   // a second candidate: the remembered one needs its exact distance too
          const long long w = T.order[single_pos];
          best.d = exact_distance(T.packed, T.NC, w + row_base, qrow, T.D, sqq, T.sqnorm[w + row_base], lane);
          best.id = id_offset + w;
          single_pos = -1;
        }
        best = bins_verify(T.packed, T.sqnorm, T.order, T.NC, T.D, m, base, qrow, sqq, id_offset, row_base, lane, best);
        cnt += k;
      }
      qpg_bin_t rec;
      if (cnt == 1 && single_pos >= 0) {
        rec.lo = best_lo;
        rec.hi = __longlong_as_double((long long)U);
        rec.id = id_offset + T.order[single_pos];
        rec.n = 1;
        rec.flags = 0;
      } else {
        rec.lo = best.d;
        rec.hi = best.d;
        rec.id = best.id;
        rec.n = cnt;
        rec.flags = 1;
        if (lane == 0 && stats) atomicAdd(&stats[0], (unsigned long long)cnt);
      }
      if (lane == 0) T.bins[(size_t)qi_ * T.bins_qstride + c] = rec;
    }
  }
  if (consume && c == KB - 1)                    // rows with a label outside [0, 512) belong to no bin
    for (int g = 0; g < ng; ++g)
      for (long long pos = b1 + lane; pos < W; pos += 32) T.sacc[(size_t)(q0 + g) * Wpad + pos] = 0;
}

__global__ void __launch_bounds__(256, 3)
    sliced_bins_kernel(const __grid_constant__ TablePair tp, long long W, long long Wpad, int nq, int64_t id_offset,
                       int64_t row_base, int consume, unsigned long long* __restrict__ stats) {
  __shared__ QConst s_qc[BG];
  __shared__ long long s_v[8 * BG * 64];         // [warp][query][row] accumulators of the warp's bin (32 KiB)
  if (blockIdx.z == 0) bins_body(tp.t[0], s_qc, s_v, W, Wpad, nq, id_offset, row_base, consume, stats);
  else bins_body(tp.t[1], s_qc, s_v, W, Wpad, nq, id_offset, row_base, consume, stats);
}

// Long bins, CTA-wide: one CTA per (table, start code, four queries), its 256 threads striding the bin's rows, so a
// 1700-row bin has ~26 independent sacc loads per thread and trip in flight (the warp-per-bin kernel below keeps
// four: it was bound by bytes in flight, 1.6 ms per 64-step pass of the all-speaker table).  Trip 1: U = min hi
// (block reduction).  Trip 2: candidates lo <= U into a small shared list.  Then one warp per query writes the
// record (one candidate) or settles the listed candidates in float64; more than CAND_MAX candidates (degenerate:
// e.g. an all-zero query ties every row) fall back to re-walking the bin.  sacc is zeroed last.
constexpr int LQ = 4;            // queries per CTA
constexpr int CAND_MAX = 32;
__device__ __forceinline__ void bins_cta_body(const TableParams& T, long long W, long long Wpad, int nq,
                                              int64_t id_offset, int64_t row_base, int consume,
                                              unsigned long long* __restrict__ stats) {
  __shared__ QConst s_qc[LQ];
  __shared__ unsigned long long s_wmin[LQ][8], s_U[LQ], s_lo1[LQ];
  __shared__ int s_cnt[LQ], s_cand[LQ][CAND_MAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x, q0 = blockIdx.y * LQ;
  const int ng = min(LQ, nq - q0);
  if (tid < LQ) {
    s_qc[tid] = make_qconst(T.q_info[min(q0 + tid, nq - 1)]);
    s_cnt[tid] = 0;
    s_lo1[tid] = ~0ull;
  }
  const int b0 = T.bin_start[c], b1 = T.bin_start[c + 1];
  __syncthreads();
  if (b1 > b0) {
    // trip 1: U per query
    unsigned long long um[LQ];
#pragma unroll
    for (int g = 0; g < LQ; ++g) um[g] = ~0ull;
    for (int pos = b0 + tid; pos < b1; pos += 256) {
      const RowInfo ri = T.row_info[pos];
      long long v[LQ];
#pragma unroll
      for (int g = 0; g < LQ; ++g) v[g] = T.sacc[(size_t)min(q0 + g, nq - 1) * Wpad + pos];
#pragma unroll
      for (int g = 0; g < LQ; ++g)
        um[g] = min(um[g], (unsigned long long)__double_as_longlong(filter_interval(v[g], ri, s_qc[g]).hi));
    }
#pragma unroll
    for (int g = 0; g < LQ; ++g) {
      const unsigned long long w = warp_min_u64(um[g]);
      if (lane == 0) s_wmin[g][warp] = w;
    }
    __syncthreads();
    if (tid < LQ) {
      unsigned long long u = s_wmin[tid][0];
#pragma unroll
      for (int w = 1; w < 8; ++w) u = min(u, s_wmin[tid][w]);
      s_U[tid] = u;
    }
    __syncthreads();
    // trip 2: candidates
    for (int pos = b0 + tid; pos < b1; pos += 256) {
      const RowInfo ri = T.row_info[pos];
      long long v[LQ];
#pragma unroll
      for (int g = 0; g < LQ; ++g) v[g] = T.sacc[(size_t)min(q0 + g, nq - 1) * Wpad + pos];
#pragma unroll
      for (int g = 0; g < LQ; ++g) {
        const unsigned long long lo = (unsigned long long)__double_as_longlong(filter_interval(v[g], ri, s_qc[g]).lo);
        if (g < ng && lo <= s_U[g]) {
          const int slot = atomicAdd(&s_cnt[g], 1);
          if (slot < CAND_MAX) s_cand[g][slot] = pos;
          atomicMin(&s_lo1[g], lo);          // with a single candidate: its lower bound
        }
      }
    }
    __syncthreads();
  }
  // records: warp g settles query q0 + g
  if (warp < ng) {
    const int g = warp, qi_ = q0 + g;
    qpg_bin_t rec;
    rec.lo = kEmptyDist;
    rec.hi = kEmptyDist;
    rec.id = -1;
    rec.n = 0;
    rec.flags = 1;                               // empty bins are exact (sentinel)
    if (b1 > b0) {
      const int cnt = s_cnt[g];
      if (cnt == 1) {
        rec.lo = __longlong_as_double((long long)s_lo1[g]);
        rec.hi = __longlong_as_double((long long)s_U[g]);
        rec.id = id_offset + T.order[s_cand[g][0]];
        rec.n = 1;
        rec.flags = 0;
      } else {
        const float* qrow = T.q + (size_t)qi_ * T.ldq;
        const double sqq = T.q_info[qi_].sq;
        BestPair best;
        best.d = 1e300;
        best.id = -1;
        if (cnt <= CAND_MAX) {
          for (int i = 0; i < cnt; ++i) {
            const long long w = T.order[s_cand[g][i]];
            const double d = exact_distance(T.packed, T.NC, w + row_base, qrow, T.D, sqq, T.sqnorm[w + row_base], lane);
            const long long id = id_offset + w;
            if (d < best.d || (d == best.d && id < best.id)) {
              best.d = d;
              best.id = id;
            }
          }
        } else {                                 // degenerate: walk the bin again, 32 rows at a time
          const QConst qg = s_qc[g];
          const unsigned long long U = s_U[g];
          const long long* sv = T.sacc + (size_t)qi_ * Wpad;
          for (int base = b0; base < b1; base += 32) {
            const int pos = base + lane;
            const bool cand = pos < b1 &&
                              (unsigned long long)__double_as_longlong(filter_interval(sv[pos], T.row_info[pos], qg).lo) <= U;
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (m) best = bins_verify(T.packed, T.sqnorm, T.order, T.NC, T.D, m, base, qrow, sqq, id_offset, row_base, lane, best);
          }
        }
        rec.lo = best.d;
        rec.hi = best.d;
        rec.id = best.id;
        rec.n = cnt;
        rec.flags = 1;
        if (lane == 0 && stats) atomicAdd(&stats[0], (unsigned long long)cnt);
      }
    }
    if (lane == 0) T.bins[(size_t)qi_ * T.bins_qstride + c] = rec;
  }
  if (consume) {
    __syncthreads();                             // the degenerate path above still reads sacc
    const long long end = c == KB - 1 ? W : b1;  // rows with a label outside [0, 512) belong to no bin
    for (long long pos = b0 + tid; pos < end; pos += 256)
      for (int g = 0; g < ng; ++g) T.sacc[(size_t)(q0 + g) * Wpad + pos] = 0;
  }
}

__global__ void __launch_bounds__(256, 3)
    sliced_bins_cta_kernel(const __grid_constant__ TablePair tp, long long W, long long Wpad, int nq, int64_t id_offset,
                           int64_t row_base, int consume, unsigned long long* __restrict__ stats) {
  if (blockIdx.z == 0) bins_cta_body(tp.t[0], W, Wpad, nq, id_offset, row_base, consume, stats);
  else bins_cta_body(tp.t[1], W, Wpad, nq, id_offset, row_base, consume, stats);
}

// Long bins (all-speaker tables and the synthetic sweeps: hundreds to thousands of rows per start code): the first
// version of the kernel, one warp per (table, start code, four queries) with the four sacc loads of a lane in flight
// together on both trips over the bin - the one-trip kernel above is sequential per query on such bins.
constexpr int BGL = 4;
__global__ void __launch_bounds__(256, 2)
    sliced_bins_long_kernel(const __grid_constant__ TablePair tp, long long W, long long Wpad, int nq, int64_t id_offset, int64_t row_base,
                       int consume, unsigned long long* __restrict__ stats) {
  const TableParams T = blockIdx.y == 0 ? tp.t[0] : tp.t[1];
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int n_groups = (nq + BGL - 1) / BGL;
  if (wid >= (long long)n_groups * KB) return;
  const int q0 = (int)(wid / KB) * BGL, c = (int)(wid % KB);
  const int b0 = T.bin_start[c], b1 = T.bin_start[c + 1];
  QConst qi[BGL];
#pragma unroll
  for (int g = 0; g < BGL; ++g) qi[g] = make_qconst(T.q_info[min(q0 + g, nq - 1)]);
  double U[BGL], lo0[BGL], hi0[BGL];
#pragma unroll
  for (int g = 0; g < BGL; ++g) {
    U[g] = 1e300;
    lo0[g] = hi0[g] = 0.0;
  }
  // pass 1: U per query; the intervals of the first 32 rows (most bins have no more) stay in registers
  for (int pos = b0 + lane; pos < b1; pos += 32) {
    const RowInfo ri = T.row_info[pos];
    long long v[BGL];
#pragma unroll
    for (int g = 0; g < BGL; ++g) v[g] = T.sacc[(size_t)min(q0 + g, nq - 1) * Wpad + pos];
#pragma unroll
    for (int g = 0; g < BGL; ++g) {
      const Interval iv = filter_interval(v[g], ri, qi[g]);
      if (pos < b0 + 32) {
        lo0[g] = iv.lo;
        hi0[g] = iv.hi;
      }
      U[g] = fmin(U[g], iv.hi);
    }
  }
#pragma unroll
  for (int g = 0; g < BGL; ++g)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) U[g] = fmin(U[g], __shfl_xor_sync(0xffffffffu, U[g], o));

#pragma unroll 1
  for (int g = 0; g < BGL; ++g) {
    const int qi_ = q0 + g;
    if (qi_ >= nq) break;
    long long* sv = T.sacc + (size_t)qi_ * Wpad;
    const float* qrow = T.q + (size_t)qi_ * T.ldq;
    const QConst qg = g == 0 ? qi[0] : g == 1 ? qi[1] : g == 2 ? qi[2] : qi[3];
    const double sqq = T.q_info[qi_].sq;
    const double Ug = g == 0 ? U[0] : g == 1 ? U[1] : g == 2 ? U[2] : U[3];
    const double l0 = g == 0 ? lo0[0] : g == 1 ? lo0[1] : g == 2 ? lo0[2] : lo0[3];
    qpg_bin_t rec;
    rec.lo = kEmptyDist;
    rec.hi = kEmptyDist;
    rec.id = -1;
    rec.n = 0;
    rec.flags = 1;                               // empty bins are exact (sentinel)
    if (b1 > b0) {
      int n = 0;
      double best_lo = 1e300, best_d = 1e300;
      long long best_id = -1, single_pos = -1;
      for (int base = b0; base < b1; base += 32) {
        const int pos = base + lane;
        bool cand = false;
        double lo = 0.0;
        if (pos < b1) {
          lo = base == b0 ? l0 : filter_interval(sv[pos], T.row_info[pos], qg).lo;
          cand = lo <= Ug;
          if (consume) sv[pos] = 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, cand);
        const int cnt = __popc(m);
        if (cnt == 0) continue;
        if (n == 0 && cnt == 1) {          // remember the first lone candidate; verified only if another shows up
          const int src_lane = __ffs(m) - 1;
          single_pos = base + src_lane;
          best_lo = __shfl_sync(0xffffffffu, lo, src_lane);
          n = 1;
          continue;
        }
        // more than one candidate so far: evaluate exactly (including the remembered one)
        if (n == 1 && single_pos >= 0) {
          const long long w = T.order[single_pos];
          best_d = exact_distance(T.packed, T.NC, w + row_base, qrow, T.D, sqq, T.sqnorm[w + row_base], lane);
          best_id = id_offset + w;
          single_pos = -1;
        }
        unsigned mm = m;
        while (mm) {
          const int src_lane = __ffs(mm) - 1;
          mm &= mm - 1;
          const long long w = T.order[base + src_lane];
          const double d = exact_distance(T.packed, T.NC, w + row_base, qrow, T.D, sqq, T.sqnorm[w + row_base], lane);
          const long long id = id_offset + w;
          if (d < best_d || (d == best_d && id < best_id)) {
            best_d = d;
            best_id = id;
          }
        }
        n += cnt;
      }
      if (n == 1 && single_pos >= 0) {
        rec.lo = best_lo;
        rec.hi = Ug;
        rec.id = id_offset + T.order[single_pos];
        rec.n = 1;
        rec.flags = 0;
      } else {
        rec.lo = best_d;
        rec.hi = best_d;
        rec.id = best_id;
        rec.n = n;                           // how many rows were re-evaluated (diagnostics)
        rec.flags = 1;                       // exact
        if (lane == 0 && stats) atomicAdd(&stats[0], (unsigned long long)n);
      }
    }
    if (lane == 0) T.bins[(size_t)qi_ * T.bins_qstride + c] = rec;
    if (consume && c == KB - 1)                  // rows with a label outside [0, 512) belong to no bin
      for (long long pos = b1 + lane; pos < W; pos += 32) sv[pos] = 0;
  }
}

// ------------------------------------------------------------------ resolve: merge parts, decide, verify, rank
// One CTA (512 threads, thread = start code) per (table, query).  T.bins points at part 0; part p is
// part_stride records further (P = 1 on a single GPU, the all-gathered per-rank records when database rows are
// sharded).  T.packed / T.sqnorm address rows by GLOBAL window id - first_id (a replicated float32 copy of the
// whole table), so any rank can re-evaluate any candidate.  Output: table [nq][512] (distance exact where it had
// to be decided, else the centre of the filter interval), ranks [nq][512] (stable, as qpg_rank512),
// qflags [nq] bit 0 = exact tie between two non-empty bins.
// ascending bitonic sort of 512 (key, val) pairs in shared memory by 512 threads; vals are distinct, so the order
// is total: equal keys keep the lower val first.  Partners closer than a warp are exchanged with shuffles, so only
// 10 of the 45 compare-exchange stages need a block barrier.
__device__ __forceinline__ void bitonic_sort_512(unsigned long long* key, int* val, int tid) {
  unsigned long long kk = key[tid];
  int vv = val[tid];
  for (int k = 2; k <= KB; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int other = tid ^ j;
      unsigned long long ok;
      int ov;
      if (j >= 32) {
        __syncthreads();
        key[tid] = kk;
        val[tid] = vv;
        __syncthreads();
        ok = key[other];
        ov = val[other];
      } else {
        ok = __shfl_xor_sync(0xffffffffu, kk, j);
        ov = __shfl_xor_sync(0xffffffffu, vv, j);
      }
      const bool want_min = ((tid & k) == 0) == (tid < other);
      const bool other_less = ok < kk || (ok == kk && ov < vv);
      if (other_less == want_min) {
        kk = ok;
        vv = ov;
      }
    }
  }
  __syncthreads();
  key[tid] = kk;
  val[tid] = vv;
  __syncthreads();
}

__global__ void __launch_bounds__(KB, 2)
    sliced_resolve_kernel(const TablePair tp, int P, long long part_stride, int64_t first_id,
                          unsigned long long* __restrict__ stats) {
  __shared__ unsigned long long s_key[KB];       // interval starts (bit patterns), sorted
  __shared__ int s_val[KB];
  __shared__ unsigned long long s_hi[KB];        // hi bits by code, then running maximum by sorted position
  __shared__ unsigned long long s_pm[KB];
  __shared__ unsigned long long s_d[KB];
  __shared__ long long s_id[KB];
  __shared__ int s_ov[KB], s_list[KB], s_link[KB];
  __shared__ unsigned long long s_wmax[KB / 32];
  __shared__ int s_n, s_tie;
  const TableParams T = blockIdx.y == 0 ? tp.t[0] : tp.t[1];
  const qpg_bin_t* __restrict__ parts = T.bins;
  const int qi_ = blockIdx.x, c = threadIdx.x, lane = c & 31, warp = c >> 5;
  const unsigned long long empty_bits = (unsigned long long)__double_as_longlong(kEmptyDist);
  if (c == 0) {
    s_n = 0;
    s_tie = 0;
  }
  // merge: U* = min hi; candidates = parts with lo <= U*
  double U = 1e300;
  for (int p = 0; p < P; ++p) {
    const qpg_bin_t r = parts[(size_t)p * part_stride + (size_t)qi_ * T.bins_qstride + c];
    if (r.id >= 0) U = fmin(U, r.hi);
  }
  int ncand = 0;
  double lo = kEmptyDist, hi = kEmptyDist;
  long long id = -1;
  bool exact = true;
  for (int p = 0; p < P; ++p) {
    const qpg_bin_t r = parts[(size_t)p * part_stride + (size_t)qi_ * T.bins_qstride + c];
    if (r.id >= 0 && r.lo <= U) {
      if (ncand == 0) {
        lo = r.lo;
        id = r.id;
        exact = (r.flags & 1) != 0;
      } else {
        if (r.lo < lo) lo = r.lo;
        exact = false;
      }
      ++ncand;
    }
  }
  if (ncand > 0) hi = U;
  // non-negative doubles order like their bit patterns
  const unsigned long long lo_b = (unsigned long long)__double_as_longlong(lo);
  const unsigned long long hi_b = (unsigned long long)__double_as_longlong(hi);
  s_key[c] = lo_b;
  s_val[c] = c;
  s_hi[c] = hi_b;
  __syncthreads();
  bitonic_sort_512(s_key, s_val, c);             // by interval start
  // running maximum of the interval ends in sorted order: shuffle scan inside each warp, then the warp totals
  unsigned long long pm = s_hi[s_val[c]];
  const unsigned long long my_hi_sorted = pm;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, pm, o);
    if (lane >= o && prev > pm) pm = prev;
  }
  if (lane == 31) s_wmax[warp] = pm;
  __syncthreads();
  {
    unsigned long long before = 0ull;
    for (int w = 0; w < warp; ++w) before = s_wmax[w] > before ? s_wmax[w] : before;
    if (before > pm) pm = before;
  }
  s_pm[c] = pm;
  __syncthreads();
  {
    // sorted position c: does this bin's interval touch another one?  (empty bins sort last and touch nothing)
    const bool nonempty = my_hi_sorted < empty_bits;
    const bool right = c + 1 < KB && s_key[c + 1] <= my_hi_sorted;
    const bool left = c > 0 && s_pm[c - 1] >= s_key[c];
    s_ov[s_val[c]] = nonempty && (left || right);
    // positions c and c + 1 belong to one cluster of (transitively) overlapping intervals
    s_link[c] = nonempty && c + 1 < KB && s_key[c + 1] < empty_bits && s_pm[c] >= s_key[c + 1];
  }
  __syncthreads();
  // a bin needs a float64 decision when several shards could hold its winner, or when its interval overlaps
  // another bin's (the rank transform would be undecided); exact points need nothing
  const bool need = ncand > 1 || (ncand == 1 && s_ov[c] && !(exact && lo == hi));
  if (need) s_list[atomicAdd(&s_n, 1)] = c;
  s_id[c] = id;
  s_d[c] = (unsigned long long)__double_as_longlong(ncand >= 1 ? 0.5 * (lo + hi) : kEmptyDist);
  s_hi[c] = hi_b;                                 // by code again (the verify loop reads U of a bin)
  __syncthreads();
  const int n_list = s_n;
  const qpg_qinfo_t qi = T.q_info[qi_];
  const float* qrow = T.q + (size_t)qi_ * T.ldq;
  for (int it = warp; it < n_list; it += KB / 32) {
    const int cc = s_list[it];
    const double Uc = __longlong_as_double((long long)s_hi[cc]);
    double bd = 1e300;
    long long bid = -1;
    for (int p = 0; p < P; ++p) {
      const qpg_bin_t r = parts[(size_t)p * part_stride + (size_t)qi_ * T.bins_qstride + cc];
      if (r.id >= 0 && r.lo <= Uc) {
        double d;
        if (r.flags & 1) d = r.lo;
        else {
          const int64_t w = r.id - first_id;
          d = exact_distance(T.packed, T.NC, w, qrow, T.D, qi.sq, T.sqnorm[w], lane);
        }
        if (d < bd || (d == bd && r.id < bid)) {
          bd = d;
          bid = r.id;
        }
      }
    }
    if (lane == 0) {
      s_d[cc] = (unsigned long long)__double_as_longlong(bd);
      s_id[cc] = bid;
    }
  }
  __syncthreads();
  if (c == 0 && stats) atomicAdd(&stats[1], (unsigned long long)n_list);
  // stable rank (ties -> lower code first) = position in the order by (distance, code).  Intervals that touch
  // nothing are already in that order (sorted by interval start; the true distance lies inside the interval), so
  // their rank is their sorted position: no second sort.  A cluster of touching intervals is a contiguous run of
  // sorted positions whose members all carry exact distances now; they are ranked among themselves.  Equal
  // distances (the tie flag) can only occur inside a cluster.
  {
    const int code = s_val[c];
    int rank = c;
    const bool in_cluster = s_link[c] || (c > 0 && s_link[c - 1]);
    if (in_cluster) {
      const unsigned long long dp = s_d[code];
      int start = c;
      while (start > 0 && s_link[start - 1]) --start;
      int smaller = 0, tie = 0;
      for (int m = start;; ++m) {
        if (m != c) {
          const int cm = s_val[m];
          const unsigned long long dm = s_d[cm];
          smaller += (dm < dp) || (dm == dp && cm < code);
          tie |= dm == dp;
        }
        if (m + 1 >= KB || !s_link[m]) break;
      }
      rank = start + smaller;
      if (tie) s_tie = 1;
    }
    T.ranks[(size_t)qi_ * KB + code] = rank;
  }
  const unsigned long long mine = s_d[c];
  Pair out;
  out.d = mine;
  out.id = (unsigned long long)s_id[c];
  T.table[(size_t)qi_ * KB + c] = out;
  __syncthreads();
  if (c == 0 && T.qflags) T.qflags[qi_] = s_tie;
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" size_t qpg_sliced_bytes(int64_t n_rows, int D) {
  if (n_rows <= 0 || D <= 0) return 0;
  const size_t rt = (size_t)((n_rows + TM - 1) / TM), nkb = (size_t)((D + KBW - 1) / KBW);
  return rt * nkb * NS * SLICE_BYTES;
}

extern "C" size_t qpg_sliced_query_bytes(int D, int n_pad) {
  if (D <= 0 || n_pad <= 0) return 0;
  return (size_t)((D + KBW - 1) / KBW) * NS * (size_t)n_pad * KBW;
}

extern "C" int qpg_slice_rows_i8(const float* rows, int64_t W, int D, const int32_t* order, const int8_t* col_exp,
                                 const double* row_sqnorm, int8_t* slices, void* row_info, void* stream) {
  QPG_CHECK_ARG(W >= 0 && D > 0 && D <= 32768, "W >= 0, 0 < D <= 32768 (int32 accumulators)");
  if (W == 0) return QPG_OK;
  QPG_CHECK_ARG(rows && row_sqnorm && slices && row_info, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(slices) & 1023) == 0, "slices must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  QPG_CUDA(cudaMemsetAsync(slices, 0, qpg_sliced_bytes(W, D), st));
  SliceJobs jobs;
  jobs.j[0].rows = rows;
  jobs.j[0].col_exp = col_exp;
  jobs.j[0].out = slices;
  jobs.j[0].q_info = nullptr;
  jobs.j[0].ld = D;
  jobs.j[0].D = D;
  jobs.j[0].nkb = (D + KBW - 1) / KBW;
  jobs.j[1] = jobs.j[0];
  slice_kernel<false><<<dim3((unsigned)W, 1), SLICE_THREADS, 0, st>>>(jobs, order, 0, row_sqnorm,
                                                           reinterpret_cast<RowInfo*>(row_info));
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_slice_queries_i8(const qpg_slice_job_t* jobs_in, int n_jobs, int Q, int n_pad, void* stream) {
  QPG_CHECK_ARG(jobs_in && (n_jobs == 1 || n_jobs == 2), "1 or 2 feature blocks");
  QPG_CHECK_ARG(Q >= 0 && n_pad >= 16 && n_pad <= MAX_NPAD && n_pad % 16 == 0 && Q <= n_pad,
                "n_pad in {16,32,48,64}, Q <= n_pad");
  if (Q == 0) return QPG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SliceJobs jobs;
  for (int i = 0; i < n_jobs; ++i) {
    const qpg_slice_job_t& j = jobs_in[i];
    QPG_CHECK_ARG(j.D > 0 && j.D <= 32768 && j.ldq >= j.D, "0 < D <= 32768, ldq >= D");
    QPG_CHECK_ARG(j.q && j.q_slices && j.q_info, "null pointer");
    QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(j.q_slices) & 1023) == 0, "q_slices must be 1024-byte aligned");
    if (Q < n_pad || j.D % KBW != 0) QPG_CUDA(cudaMemsetAsync(j.q_slices, 0, qpg_sliced_query_bytes(j.D, n_pad), st));
    jobs.j[i].rows = j.q;
    jobs.j[i].col_exp = j.col_exp;
    jobs.j[i].out = j.q_slices;
    jobs.j[i].q_info = j.q_info;
    jobs.j[i].ld = j.ldq;
    jobs.j[i].D = j.D;
    jobs.j[i].nkb = (j.D + KBW - 1) / KBW;
  }
  if (n_jobs == 1) jobs.j[1] = jobs.j[0];
  slice_kernel<true><<<dim3((unsigned)Q, (unsigned)n_jobs), SLICE_THREADS, 0, st>>>(jobs, nullptr, n_pad, nullptr, nullptr);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_sliced_scan_i8(const qpg_sliced_seg_t* segs, int n_segs, int64_t W, int n_pad, int nq,
                                  void* stream) {
  QPG_CHECK_ARG(segs && (n_segs == 1 || n_segs == 2), "1 or 2 feature blocks");
  QPG_CHECK_ARG(W >= 0 && n_pad >= 16 && n_pad <= MAX_NPAD && n_pad % 16 == 0 && nq >= 0 && nq <= n_pad,
                "n_pad in {16,32,48,64}, nq <= n_pad");
  if (W == 0 || nq == 0) return QPG_OK;
  ScanParams p;
  p.nseg = n_segs;
  p.nkb_total = 0;
  for (int i = 0; i < 2; ++i) {
    p.seg[i].A = nullptr;
    p.seg[i].B = nullptr;
    p.seg[i].sacc = nullptr;
    p.seg[i].nkb = 0;
  }
  for (int i = 0; i < n_segs; ++i) {
    QPG_CHECK_ARG(segs[i].db_slices && segs[i].q_slices && segs[i].sacc && segs[i].n_kblocks > 0, "null block");
    QPG_CHECK_ARG(segs[i].n_kblocks <= 256, "D <= 32768");
    QPG_CHECK_ARG(((reinterpret_cast<uintptr_t>(segs[i].db_slices) | reinterpret_cast<uintptr_t>(segs[i].q_slices)) & 127) == 0,
                  "slices must be 128-byte aligned");
    p.seg[i].A = segs[i].db_slices;
    p.seg[i].B = segs[i].q_slices;
    p.seg[i].sacc = reinterpret_cast<unsigned long long*>(segs[i].sacc);
    p.seg[i].nkb = segs[i].n_kblocks;
    p.nkb_total += segs[i].n_kblocks;
  }
  p.n_pad = n_pad;
  p.nq = nq;
  p.W = W;
  p.RT = (W + TM - 1) / TM;
  p.Wpad = p.RT * TM;
  p.total_units = p.RT * p.nkb_total;
  int grid = sm_count();
  if ((long long)grid > p.total_units) grid = (int)p.total_units;
  // sized for this pass's query count: what it leaves free lets the small kernels of other pipeline lanes co-reside
  const size_t smem = (size_t)STAGES * (NS * SLICE_BYTES + NS * n_pad * KBW) + 8 * sizeof(uint64_t) + 16;
  QPG_CUDA(cudaFuncSetAttribute(sliced_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sliced_scan_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_sliced_scan_ref(const int8_t* db_slices, const int8_t* q_slices, int n_kblocks, int64_t W, int n_pad,
                                   int nq, int q_stride, int64_t* sacc, void* stream) {
  QPG_CHECK_ARG(db_slices && q_slices && sacc && n_kblocks > 0 && W > 0 && nq > 0 && nq <= n_pad && q_stride >= 1,
                "bad argument");
  const long long Wpad = (W + TM - 1) / TM * TM;
  dim3 grid((unsigned)((W + 127) / 128), (unsigned)((nq + q_stride - 1) / q_stride));
  sliced_scan_ref_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(db_slices, q_slices, n_kblocks, n_pad, nq, q_stride, W, Wpad,
                                                                 reinterpret_cast<long long*>(sacc));
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

static int fill_tables(const qpg_sliced_table_t* tabs, int n_tabs, bool for_bins, TablePair* tp) {
  QPG_CHECK_ARG(tabs && (n_tabs == 1 || n_tabs == 2), "1 or 2 tables");
  for (int i = 0; i < n_tabs; ++i) {
    const qpg_sliced_table_t& t = tabs[i];
    QPG_CHECK_ARG(t.D > 0 && t.ldq >= t.D, "D > 0, ldq >= D");
    QPG_CHECK_ARG(t.packed && t.row_sqnorm && t.q && t.q_info && t.bins, "null pointer");
    if (for_bins) QPG_CHECK_ARG(t.sacc && t.bin_start && t.row_info && t.order, "null pointer (bins inputs)");
    else QPG_CHECK_ARG(t.table && t.ranks, "null pointer (resolve outputs)");
    TableParams& o = tp->t[i];
    o.packed = t.packed;
    o.sqnorm = t.row_sqnorm;
    o.q = t.q;
    o.q_info = t.q_info;
    o.ldq = t.ldq;
    o.D = t.D;
    o.NC = (t.D + 127) / 128;
    o.sacc = reinterpret_cast<long long*>(t.sacc);
    o.bin_start = t.bin_start;
    o.row_info = reinterpret_cast<const RowInfo*>(t.row_info);
    o.order = t.order;
    o.bins = t.bins;
    o.bins_qstride = t.bins_qstride > 0 ? t.bins_qstride : KB;
    o.table = reinterpret_cast<Pair*>(t.table);
    o.ranks = t.ranks;
    o.qflags = t.qflags;
  }
  if (n_tabs == 1) tp->t[1] = tp->t[0];
  return QPG_OK;
}

extern "C" int qpg_sliced_bins(const qpg_sliced_table_t* tabs, int n_tabs, int64_t W, int nq, int64_t id_offset,
                               int64_t row_base, int consume, uint64_t* stats, void* stream) {
  QPG_CHECK_ARG(W >= 0 && nq >= 0, "bad size");
  if (nq == 0) return QPG_OK;
  TablePair tp;
  const int rc = fill_tables(tabs, n_tabs, true, &tp);
  if (rc != QPG_OK) return rc;
  const long long Wpad = (W + TM - 1) / TM * TM;
  if (W <= 32 * KB) {                          // bins of about one warp of rows: settled in one trip
    const dim3 grid(KB / 8, (unsigned)((nq + BG - 1) / BG), (unsigned)n_tabs);    // CTA = 8 start codes x 8 queries
    sliced_bins_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tp, W, Wpad, nq, id_offset, row_base, consume,
                                                               reinterpret_cast<unsigned long long*>(stats));
  } else if (W >= 128 * KB) {                  // hundreds of rows per start code: a CTA per bin
    const dim3 grid(KB, (unsigned)((nq + LQ - 1) / LQ), (unsigned)n_tabs);
    sliced_bins_cta_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tp, W, Wpad, nq, id_offset, row_base, consume,
                                                                   reinterpret_cast<unsigned long long*>(stats));
  } else {
    const long long warps = (long long)((nq + BGL - 1) / BGL) * KB;
    const dim3 grid((unsigned)((warps * 32 + 255) / 256), (unsigned)n_tabs);
    sliced_bins_long_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tp, W, Wpad, nq, id_offset, row_base, consume,
                                                                    reinterpret_cast<unsigned long long*>(stats));
  }
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_sliced_resolve(const qpg_sliced_table_t* tabs, int n_tabs, int n_parts, int64_t part_stride, int nq,
                                  int64_t first_id, uint64_t* stats, void* stream) {
  QPG_CHECK_ARG(n_parts >= 1 && nq >= 0, "bad size");
  QPG_CHECK_ARG(n_parts == 1 || part_stride >= (int64_t)nq * KB, "part_stride smaller than one part");
  if (nq == 0) return QPG_OK;
  TablePair tp;
  const int rc = fill_tables(tabs, n_tabs, false, &tp);
  if (rc != QPG_OK) return rc;
  sliced_resolve_kernel<<<dim3((unsigned)nq, (unsigned)n_tabs), KB, 0, (cudaStream_t)stream>>>(
      tp, n_parts, (long long)part_stride, first_id, reinterpret_cast<unsigned long long*>(stats));
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
