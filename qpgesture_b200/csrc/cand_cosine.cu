// Candidate-distance kernel: cosine distance of up to 8 queries against every
// database window, fused with the reference's "best window per start-code"
// reduction (GestureKNN.py:666-691 mode 'wavlm_feat', :708-721).
//
// Roofline: HBM bandwidth.  One launch = one pass over the packed window table
// (W * 4 * ceil128(D) bytes) for QT queries; arithmetic is 1 F2F + QT DFMA per
// database element (float64 accumulation of float32 data so that the selected
// window ids equal the reference's float64 sklearn path).
//
// Data movement: the table is stored as 4 KiB tiles [8 rows][128 floats]
// (qpg_pack_rows_f32); every warp owns a contiguous run of row groups, streams
// its tiles through a private NS-deep shared-memory ring with TMA bulk copies
// (cp.async.bulk -> UBLKCP) completing on mbarriers, and refills a slot as soon
// as it has consumed it.  Queries live in shared memory for the whole launch.
// Per row group the 8xQT partial dot products are transposed-reduced with warp
// shuffles so that each lane ends up owning one (row, query) pair; it computes
// the distance and does a lexicographic (distance, window id) atomic min
// (128-bit CAS) into the CTA's [QT][512] table, which is merged into the global
// table at the end of the launch.
#include "qpg_common.cuh"
#include "cosine_common.cuh"

namespace qpg {
namespace {


// ------------------------------------------------------------------ pack ----
// One CTA (8 warps) per row group; warp r owns row 8g+r.
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ rows, int64_t W, int D,
                                                        int NC, float* __restrict__ packed,
                                                        double* __restrict__ row_sqnorm) {
  const int64_t g = blockIdx.x;
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = g * R + r;
  const bool live = row < W;
  const float* src = rows + row * (int64_t)D;
  double acc = 0.0;
  for (int c = 0; c < NC; ++c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int d0 = c * DC + 4 * lane;
    if (live) {
      if (d0 + 0 < D) v.x = src[d0 + 0];
      if (d0 + 1 < D) v.y = src[d0 + 1];
      if (d0 + 2 < D) v.z = src[d0 + 2];
      if (d0 + 3 < D) v.w = src[d0 + 3];
    }
    acc = fma((double)v.x, (double)v.x, acc);
    acc = fma((double)v.y, (double)v.y, acc);
    acc = fma((double)v.z, (double)v.z, acc);
    acc = fma((double)v.w, (double)v.w, acc);
    float4* dst = reinterpret_cast<float4*>(packed + ((g * NC + c) * R + r) * (int64_t)DC) + lane;
    *dst = v;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && live) row_sqnorm[row] = acc;
}

__global__ void table_init_kernel(Pair* t, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    Pair p;
    p.d = (unsigned long long)__double_as_longlong(kEmptyDist);
    p.id = ~0ull;
    t[i] = p;
  }
}

// ------------------------------------------------------------ main kernel ----
#ifndef QPG_QCVT_ALU
#define QPG_QCVT_ALU 0  // measured: ALU-side conversion is slower (66.0 vs 61.5 us at D=6144, 511 vs 364 us on 1Mx512)
#endif
#if QPG_QCVT_ALU
#define QPG_QCVT(x) f32_to_f64_alu(x)
#else
#define QPG_QCVT(x) ((double)(x))
#endif
constexpr int MAXNS = 4;  // ring depth cap per warp (mbarriers are allocated for this many)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// S = team size: the chunks of one row group are split over S consecutive warps
// (member m takes chunks [m*NC/S, (m+1)*NC/S)); member 0 adds the members' partial
// dot products in member order and runs the distance / min-by-code epilogue.
// S > 1 is chosen by the host when there are fewer row groups than warps.
template <int QT, int NCW>
__global__ void __launch_bounds__(NCW * 32, 1)
    cand_cosine_kernel(const float* __restrict__ packed, const double* __restrict__ row_sqnorm,
                       const int32_t* __restrict__ labels, int64_t W, int D, int NC, int64_t G,
                       int64_t id_offset, const float* __restrict__ q, int nq, Pair* __restrict__ table,
                       int pool_tiles, int S, int reverse) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int PER_LANE = (R * QT >= 32) ? (R * QT) / 32 : 1;
  constexpr int DUP = (R * QT >= 32) ? 1 : 32 / (R * QT);  // lanes holding the same total
  const int Dp = NC * DC;
  float* ring = reinterpret_cast<float*>(smem_raw);                        // [pool_tiles][1024]
  float* qs = ring + (size_t)pool_tiles * TILE_FLOATS;                     // [QT][Dp]
  Pair* tab = reinterpret_cast<Pair*>(qs + (size_t)QT * Dp);               // [QT][512]
  double* part = reinterpret_cast<double*>(tab + QT * KB);                 // [NCW/S][S-1][PER_LANE*32] team partials
  double* qn = part + (size_t)(NCW / S) * (S - 1) * (PER_LANE * 32);       // [QT] squared norms
  uint64_t* bars = reinterpret_cast<uint64_t*>(qn + QT);                   // [NCW][MAXNS] + 1
  int* busy = reinterpret_cast<int*>(bars + NCW * MAXNS + 1);              // [NCW]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = NCW * 32;

  // ---- work split: teams of S warps own contiguous runs of row groups
  const int team = warp / S, member = warp - team * S;
  const int teams_per_cta = NCW / S;
  const int64_t tslot = (int64_t)blockIdx.x * teams_per_cta + team;
  const int64_t nteams = (int64_t)gridDim.x * teams_per_cta;
  const int64_t g0 = tslot * G / nteams, g1 = (tslot + 1) * G / nteams;
  const int c_lo = (int)((int64_t)member * NC / S), c_hi = (int)((int64_t)(member + 1) * NC / S);
  const int ncs = c_hi - c_lo;
  const int64_t n_groups = g1 - g0;
  const int64_t n_tiles = n_groups * ncs;

  // ---- prologue: barriers; ring depth = pool / busy warps of this CTA
  uint64_t* qbar = bars + NCW * MAXNS;
  if (tid < NCW * MAXNS) mbar_init(&bars[tid], 1);
  if (tid == 0) mbar_init(qbar, 1);
  if (lane == 0) busy[warp] = n_tiles > 0 ? 1 : 0;
  fence_mbar_init();
  __syncthreads();
  int n_busy = 0, busy_rank = 0;
#pragma unroll
  for (int w = 0; w < NCW; ++w) {
    n_busy += busy[w];
    busy_rank += (w < warp) ? busy[w] : 0;
  }
  int NS = n_busy > 0 ? pool_tiles / n_busy : 1;
  NS = NS > MAXNS ? MAXNS : NS;
  float* my_ring = ring + (size_t)busy_rank * NS * TILE_FLOATS;
  uint64_t* my_bars = bars + warp * MAXNS;

  // prefetch cursor (pg, pc) runs NS tiles ahead of the consume cursor (g, c).  With `reverse` the
  // team walks its row groups from last to first (chunk order inside a group never changes, so the
  // summation order of a dot product is the same): consecutive passes alternate direction and the
  // second one starts on the part of the table the first one left in L2.
  const int64_t gstep = reverse ? -1 : 1;
  const int64_t gfirst = reverse ? g1 - 1 : g0;
  int64_t pg = gfirst;
  int pc = c_lo;
  int64_t issued = 0;
  if (lane == 0) {
    for (int s = 0; s < NS && issued < n_tiles; ++s, ++issued) {
      mbar_arrive_expect_tx(&my_bars[s], TILE_BYTES);
      bulk_g2s(my_ring + (size_t)s * TILE_FLOATS, packed + (pg * NC + pc) * (int64_t)TILE_FLOATS, TILE_BYTES,
               &my_bars[s]);
      if (++pc == c_hi) {
        pc = c_lo;
        pg += gstep;
      }
    }
  }

  // queries: one TMA bulk copy when rows are unpadded; table init meanwhile
  const bool q_bulk = (D == Dp) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
  if (q_bulk) {
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)nq * (uint32_t)Dp * 4u;
      mbar_arrive_expect_tx(qbar, bytes);
      bulk_g2s(qs, q, bytes, qbar);
    }
    for (int i = nq * Dp + tid; i < QT * Dp; i += nthreads) qs[i] = 0.f;
  } else {
    for (int i = tid; i < QT * Dp; i += nthreads) {
      const int qi = i / Dp, d = i - qi * Dp;
      qs[i] = (qi < nq && d < D) ? q[(size_t)qi * D + d] : 0.f;
    }
  }
  for (int i = tid; i < QT * KB; i += nthreads) {
    Pair p;
    p.d = (unsigned long long)__double_as_longlong(kEmptyDist);
    p.id = ~0ull;
    tab[i] = p;
  }
  if (q_bulk) mbar_wait(qbar, 0);
  __syncthreads();
  if (warp < QT) {
    double s = 0.0;
    for (int d = lane; d < Dp; d += 32) {
      const double v = (double)qs[warp * Dp + d];
      s = fma(v, v, s);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) qn[warp] = s;
  }
  __syncthreads();

  double acc[R * QT];
#pragma unroll
  for (int i = 0; i < R * QT; ++i) acc[i] = 0.0;

  const int bar_a = 1 + 2 * team, bar_b = 2 + 2 * team;  // named barriers of this team
  // partial-sum slots of this team: one per member 1..S-1
  double* team_part = part + (size_t)team * (S - 1) * (PER_LANE * 32);
  int s = 0, c = c_lo;
  uint32_t parity = 0;
  int64_t g = gfirst;
  int64_t g_done = 0;
  // a member without chunks (NC < S) still takes part in the team protocol once per group
  const int64_t n_iter = ncs > 0 ? n_tiles : n_groups;
  for (int64_t it = 0; it < n_iter; ++it) {
    if (ncs > 0) {
      mbar_wait(&my_bars[s], parity);
      const float4* tile = reinterpret_cast<const float4*>(my_ring + (size_t)s * TILE_FLOATS) + lane;
      float4 x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = tile[r * (DC / 4)];
      float4 qv[QT];
#pragma unroll
      for (int qi = 0; qi < QT; ++qi)
        qv[qi] = *reinterpret_cast<const float4*>(qs + (size_t)qi * Dp + c * DC + 4 * lane);

#pragma unroll
      for (int comp = 0; comp < 4; ++comp) {
        double qd[QT];
#pragma unroll
        for (int qi = 0; qi < QT; ++qi)
          qd[qi] = QPG_QCVT(comp == 0 ? qv[qi].x : comp == 1 ? qv[qi].y : comp == 2 ? qv[qi].z : qv[qi].w);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const double xd = (double)(comp == 0 ? x[r].x : comp == 1 ? x[r].y : comp == 2 ? x[r].z : x[r].w);
#pragma unroll
          for (int qi = 0; qi < QT; ++qi) acc[r * QT + qi] = fma(xd, qd[qi], acc[r * QT + qi]);
        }
      }

      // slot consumed by every lane -> refill it with the next tile of this warp's stream
      __syncwarp();
      if (lane == 0 && issued < n_tiles) {
        fence_proxy_async();
        mbar_arrive_expect_tx(&my_bars[s], TILE_BYTES);
        bulk_g2s(my_ring + (size_t)s * TILE_FLOATS, packed + (pg * NC + pc) * (int64_t)TILE_FLOATS, TILE_BYTES,
                 &my_bars[s]);
        ++issued;
        if (++pc == c_hi) {
          pc = c_lo;
          pg += gstep;
        }
      }
      if (++s == NS) {
        s = 0;
        parity ^= 1u;
      }
      if (++c < c_hi) continue;
      c = c_lo;
    }

    // ---- this warp's share of row group g is complete
    TransposeReduce<R * QT, 16>::run(acc, lane);
    if (S > 1) {
      if (member > 0) {
        if (g_done > 0) named_bar_sync(bar_b, S * 32);  // member 0 has consumed the previous partials
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) team_part[(size_t)(member - 1) * (PER_LANE * 32) + j * 32 + lane] = acc[j];
        named_bar_arrive(bar_a, S * 32);
      } else {
        named_bar_sync(bar_a, S * 32);
        for (int m = 1; m < S; ++m) {
#pragma unroll
          for (int j = 0; j < PER_LANE; ++j) acc[j] += team_part[(size_t)(m - 1) * (PER_LANE * 32) + j * 32 + lane];
        }
        if (g_done + 1 < n_groups) named_bar_arrive(bar_b, S * 32);
      }
    }
    if (member == 0) {
      const int r = lane >> 2;
      const int64_t row = g * R + r;
      if (row < W && (lane % DUP) == 0) {
        const double sqx = row_sqnorm[row];
        const int label = labels[row];
        if ((unsigned)label < (unsigned)KB) {
#pragma unroll
          for (int j = 0; j < PER_LANE; ++j) {
            const int idx = (R * QT >= 32) ? lane * PER_LANE + j : lane / DUP;
            const int qi = idx % QT;
            if (qi < nq) {
              const double dist = cosine_distance(acc[j], qn[qi], sqx);
              if (dist < kEmptyDist) {
                Pair mine;
                mine.d = (unsigned long long)__double_as_longlong(dist);
                mine.id = (unsigned long long)(id_offset + row);
                pair_min_shared(&tab[qi * KB + label], mine);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R * QT; ++i) acc[i] = 0.0;
    g += gstep;
    ++g_done;
  }

  // ---- merge the CTA table into the global one
  __syncthreads();
  for (int i = tid; i < nq * KB; i += nthreads) {
    const Pair e = tab[i];
    if (e.id != ~0ull) pair_min_global(&table[i], e);
  }
}

struct Tuning {
  int ncw = 0;  // 0 = auto
  int ns = 0;
  int grid = 0;
  int team = 0;
  int alternate = 1;  // alternate the walk direction between consecutive passes
};
Tuning g_tuning;

size_t fixed_smem(int QT, int D, int ncw, int S) {
  const int NC = (D + DC - 1) / DC;
  const int per_lane = (R * QT >= 32) ? (R * QT) / 32 : 1;
  return (size_t)QT * NC * DC * 4 + (size_t)QT * KB * sizeof(Pair) +
         (size_t)(ncw / S) * (S - 1) * per_lane * 32 * sizeof(double) + (size_t)QT * sizeof(double) +
         ((size_t)ncw * MAXNS + 1) * sizeof(uint64_t) + (size_t)ncw * sizeof(int);
}

template <int QT, int NCW>
int launch_cosine(const float* packed, const double* row_sqnorm, const int32_t* labels, int64_t W, int D,
                  int64_t id_offset, const float* q, int nq, Pair* table, int pool_tiles, int S, int grid,
                  int reverse, cudaStream_t st) {
  const int NC = (D + DC - 1) / DC;
  const int64_t G = (W + R - 1) / R;
  const size_t smem = (size_t)pool_tiles * TILE_BYTES + fixed_smem(QT, D, NCW, S);
  auto kern = cand_cosine_kernel<QT, NCW>;
  QPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NCW * 32, smem, st>>>(packed, row_sqnorm, labels, W, D, NC, G, id_offset, q, nq, table, pool_tiles,
                                     S, reverse);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

}  // namespace

int cosine_set_alternate(int on) {
  g_tuning.alternate = on ? 1 : 0;
  return QPG_OK;
}

int cosine_set_tuning(int ncw, int ns, int grid, int team) {
  g_tuning.ncw = ncw;
  g_tuning.ns = ns;
  g_tuning.grid = grid;
  g_tuning.team = team;
  return QPG_OK;
}

}  // namespace qpg

using namespace qpg;

extern "C" size_t qpg_packed_bytes(int64_t W, int D) {
  if (W < 0 || D <= 0) return 0;
  const int64_t G = (W + R - 1) / R;
  const int64_t NC = (D + DC - 1) / DC;
  return (size_t)(G * NC) * TILE_BYTES;
}

extern "C" int qpg_pack_rows_f32(const float* rows, int64_t W, int D, float* packed, double* row_sqnorm,
                                 void* stream) {
  QPG_CHECK_ARG(W >= 0 && D > 0, "W >= 0 and D > 0");
  if (W == 0) return QPG_OK;
  QPG_CHECK_ARG(rows && packed && row_sqnorm, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "packed must be 128-byte aligned");
  const int64_t G = (W + R - 1) / R;
  const int NC = (D + DC - 1) / DC;
  QPG_CHECK_ARG(G < (1ll << 31), "too many row groups");
  pack_rows_kernel<<<(unsigned)G, 256, 0, (cudaStream_t)stream>>>(rows, W, D, NC, packed, row_sqnorm);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_table_init(qpg_pair_t* table, int64_t n_entries, void* stream) {
  QPG_CHECK_ARG(n_entries >= 0, "n_entries >= 0");
  if (n_entries == 0) return QPG_OK;
  QPG_CHECK_ARG(table != nullptr, "null table");
  const int64_t blocks = (n_entries + 255) / 256;
  table_init_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<Pair*>(table), n_entries);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

static int cosine_scan(const float* packed, const double* row_sqnorm, const int32_t* labels, int64_t W, int D,
                       int64_t id_offset, const float* q, int Q, qpg_pair_t* table, int queries_per_pass,
                       int team_size, void* stream) {
  QPG_CHECK_ARG(W >= 0 && D > 0 && Q >= 0, "W >= 0, D > 0, Q >= 0");
  if (W == 0 || Q == 0) return QPG_OK;
  QPG_CHECK_ARG(packed && row_sqnorm && labels && q && table, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "packed must be 128-byte aligned");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(table) & 15) == 0, "table must be 16-byte aligned");
  QPG_CHECK_ARG(queries_per_pass >= 0 && queries_per_pass <= 8, "queries_per_pass in 0..8");
  QPG_CHECK_ARG(team_size >= 0 && team_size <= 6 && team_size != 5, "team_size in {0,1,2,3,4,6}");
  cudaStream_t st = (cudaStream_t)stream;
  const int NC = (D + DC - 1) / DC;
  const int64_t G = (W + R - 1) / R;

  // queries per pass: as many as fit next to a 2-deep ring for 12 warps, capped by Q
  int qt = queries_per_pass ? queries_per_pass : 8;
  while (qt > 1 && (qt / 2) >= Q) qt /= 2;  // do not carry idle query lanes
  if (qt != 1 && qt != 2 && qt != 4 && qt != 8) qt = qt > 4 ? 4 : (qt > 2 ? 2 : 1);
  while (qt > 1 && fixed_smem(qt, D, 12, 1) + 24 * TILE_BYTES > kSmemLimit) qt /= 2;
  int ncw = g_tuning.ncw ? g_tuning.ncw : (qt == 8 ? 8 : 12);
  if (qt == 8) ncw = 8;
  if (ncw != 8 && ncw != 12) ncw = 8;
  int grid = g_tuning.grid ? g_tuning.grid : sm_count();
  // team size: split row groups over S warps when there are fewer groups than warps
  int S = team_size ? team_size : g_tuning.team;
  if (S == 0) {
    // cost ~ (rounds of row groups per team) x (chunks per member); ties -> smaller team
    const int cand[5] = {1, 2, 3, 4, 6};
    int64_t best = -1;
    for (int i = 0; i < 5; ++i) {
      const int c = cand[i];
      if (ncw % c != 0 || c > NC) continue;
      const int64_t nteams = (int64_t)grid * (ncw / c);
      const int64_t cost = ((G + nteams - 1) / nteams) * ((NC + c - 1) / c);
      if (best < 0 || cost < best) {
        best = cost;
        S = c;
      }
    }
    if (S == 0) S = 1;
  }
  if (ncw % S != 0 || S > NC) S = 1;
  // the ring must stay at least 2 deep for every warp; shrink the team before giving that up
  while (S > 1 && fixed_smem(qt, D, ncw, S) + (size_t)2 * ncw * TILE_BYTES > kSmemLimit) {
    do { --S; } while (S > 1 && (ncw % S != 0));
  }
  if (fixed_smem(qt, D, ncw, S) + (size_t)ncw * TILE_BYTES > kSmemLimit) {
    set_error("D=%d too large for the shared-memory query tile", D);
    return QPG_E_UNSUPPORTED;
  }
  const int tiles_fit = (int)((kSmemLimit - fixed_smem(qt, D, ncw, S)) / TILE_BYTES);
  int ns = g_tuning.ns ? g_tuning.ns : 3;
  int pool = ns * ncw;
  if (pool > tiles_fit) pool = tiles_fit;
  // keep every SM busy: row groups are spread evenly over all CTAs (some warps idle) rather than
  // packed into fewer, fuller CTAs; only shrink the grid when there are fewer groups than CTAs
  if (grid > G) grid = (int)G;
  if (grid < 1) grid = 1;

  Pair* tab = reinterpret_cast<Pair*>(table);
  for (int q0 = 0; q0 < Q; q0 += qt) {
    const int nq = (Q - q0) < qt ? (Q - q0) : qt;
    const float* qp = q + (size_t)q0 * D;
    Pair* tp = tab + (size_t)q0 * KB;
    const int rev = (g_tuning.alternate && ((q0 / qt) & 1)) ? 1 : 0;
    int rc;
#define QPG_DISPATCH(QT_, NCW_)                                                                              \
  rc = launch_cosine<QT_, NCW_>(packed, row_sqnorm, labels, W, D, id_offset, qp, nq, tp, pool, S, grid, rev, st)
    if (qt == 8) QPG_DISPATCH(8, 8);
    else if (qt == 4 && ncw == 12) QPG_DISPATCH(4, 12);
    else if (qt == 4) QPG_DISPATCH(4, 8);
    else if (qt == 2 && ncw == 12) QPG_DISPATCH(2, 12);
    else if (qt == 2) QPG_DISPATCH(2, 8);
    else if (ncw == 12) QPG_DISPATCH(1, 12);
    else QPG_DISPATCH(1, 8);
#undef QPG_DISPATCH
    if (rc != QPG_OK) return rc;
  }
  return QPG_OK;
}

extern "C" int qpg_cand_cosine_minbycode(const float* packed, const double* row_sqnorm, const int32_t* labels,
                                         int64_t W, int D, int64_t id_offset, const float* q, int Q,
                                         qpg_pair_t* table, int queries_per_pass, void* stream) {
  return cosine_scan(packed, row_sqnorm, labels, W, D, id_offset, q, Q, table, queries_per_pass, 0, stream);
}

extern "C" int qpg_cand_cosine_minbycode_team(const float* packed, const double* row_sqnorm, const int32_t* labels,
                                              int64_t W, int D, int64_t id_offset, const float* q, int Q,
                                              qpg_pair_t* table, int queries_per_pass, int team_size, void* stream) {
  return cosine_scan(packed, row_sqnorm, labels, W, D, id_offset, q, Q, table, queries_per_pass, team_size, stream);
}
