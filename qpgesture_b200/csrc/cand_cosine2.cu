// Fused audio + text candidate scan: one pass over a table whose rows carry BOTH feature blocks
// ([D1 floats audio | D2 floats text], D1 a multiple of 128), producing both min-by-start-code
// tables (search_audio_cands :666-691 and search_text_cands :708-721 in a single sweep).
//
// Same streaming structure as cand_cosine_kernel (4 KiB tiles, per-warp TMA bulk-copy ring,
// float64 accumulation, transposed warp reduce), with two differences:
//   * a row group is finished twice: after chunk NC1 (audio dot products -> audio table) the
//     accumulators are reset and reused for chunks NC1..NC (text dot products -> text table);
//   * there is no per-CTA shared-memory table (the extra query columns need that space): results go
//     straight to the global tables with the 128-bit CAS lexicographic min.
// The text block adds D2/D1 = 6 % bytes to the pass and removes every separate text launch.
#include "cosine_common.cuh"

namespace qpg {
namespace {

constexpr int MAXNS2 = 4;

template <int QT, int NCW>
__global__ void __launch_bounds__(NCW * 32, 1)
    cand_cosine2_kernel(const float* __restrict__ packed, const double* __restrict__ sqnorm1,
                        const double* __restrict__ sqnorm2, const int32_t* __restrict__ labels, int64_t W, int NC1,
                        int NC, int64_t G, int64_t id_offset, const float* __restrict__ q, int nq,
                        Pair* __restrict__ table1, Pair* __restrict__ table2, int pool_tiles, int reverse) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int PER_LANE = (R * QT >= 32) ? (R * QT) / 32 : 1;
  constexpr int DUP = (R * QT >= 32) ? 1 : 32 / (R * QT);
  const int Dp = NC * DC;
  float* ring = reinterpret_cast<float*>(smem_raw);                        // [pool_tiles][1024]
  float* qs = ring + (size_t)pool_tiles * TILE_FLOATS;                     // [QT][Dp]
  double* qn = reinterpret_cast<double*>(qs + (size_t)QT * Dp);            // [2][QT] squared norms per block
  uint64_t* bars = reinterpret_cast<uint64_t*>(qn + 2 * QT);               // [NCW][MAXNS2] + 1
  int* busy = reinterpret_cast<int*>(bars + NCW * MAXNS2 + 1);             // [NCW]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = NCW * 32;

  const int64_t slot = (int64_t)blockIdx.x * NCW + warp;
  const int64_t nslots = (int64_t)gridDim.x * NCW;
  const int64_t g0 = slot * G / nslots, g1 = (slot + 1) * G / nslots;
  const int64_t n_groups = g1 - g0;
  const int64_t n_tiles = n_groups * NC;

  uint64_t* qbar = bars + NCW * MAXNS2;
  if (tid < NCW * MAXNS2) mbar_init(&bars[tid], 1);
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
  NS = NS > MAXNS2 ? MAXNS2 : NS;
  float* my_ring = ring + (size_t)busy_rank * NS * TILE_FLOATS;
  uint64_t* my_bars = bars + warp * MAXNS2;

  const int64_t gstep = reverse ? -1 : 1;
  const int64_t gfirst = reverse ? g1 - 1 : g0;
  int64_t pg = gfirst;
  int pc = 0;
  int64_t issued = 0;
  if (lane == 0) {
    for (int s = 0; s < NS && issued < n_tiles; ++s, ++issued) {
      mbar_arrive_expect_tx(&my_bars[s], TILE_BYTES);
      bulk_g2s(my_ring + (size_t)s * TILE_FLOATS, packed + (pg * NC + pc) * (int64_t)TILE_FLOATS, TILE_BYTES,
               &my_bars[s]);
      if (++pc == NC) {
        pc = 0;
        pg += gstep;
      }
    }
  }

  // queries [nq][Dp] (audio | text, unpadded because D1 % 128 == 0 and Dp == D1 + D2 rounded by the host)
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)nq * (uint32_t)Dp * 4u;
    mbar_arrive_expect_tx(qbar, bytes);
    bulk_g2s(qs, q, bytes, qbar);
  }
  for (int i = nq * Dp + tid; i < QT * Dp; i += nthreads) qs[i] = 0.f;
  mbar_wait(qbar, 0);
  __syncthreads();
  if (warp < 2 * QT) {
    const int blk = warp / QT, qi = warp - blk * QT;
    const int d_lo = blk == 0 ? 0 : NC1 * DC, d_hi = blk == 0 ? NC1 * DC : Dp;
    double s = 0.0;
    for (int d = d_lo + lane; d < d_hi; d += 32) {
      const double v = (double)qs[qi * Dp + d];
      s = fma(v, v, s);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) qn[blk * QT + qi] = s;
  }
  __syncthreads();

  double acc[R * QT];
#pragma unroll
  for (int i = 0; i < R * QT; ++i) acc[i] = 0.0;

  int s = 0, c = 0;
  uint32_t parity = 0;
  int64_t g = gfirst;
  for (int64_t it = 0; it < n_tiles; ++it) {
    mbar_wait(&my_bars[s], parity);
    const float4* tile = reinterpret_cast<const float4*>(my_ring + (size_t)s * TILE_FLOATS) + lane;
    float4 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = tile[r * (DC / 4)];
    float4 qv[QT];
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) qv[qi] = *reinterpret_cast<const float4*>(qs + (size_t)qi * Dp + c * DC + 4 * lane);
#pragma unroll
    for (int comp = 0; comp < 4; ++comp) {
      double qd[QT];
#pragma unroll
      for (int qi = 0; qi < QT; ++qi)
        qd[qi] = (double)(comp == 0 ? qv[qi].x : comp == 1 ? qv[qi].y : comp == 2 ? qv[qi].z : qv[qi].w);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double xd = (double)(comp == 0 ? x[r].x : comp == 1 ? x[r].y : comp == 2 ? x[r].z : x[r].w);
#pragma unroll
        for (int qi = 0; qi < QT; ++qi) acc[r * QT + qi] = fma(xd, qd[qi], acc[r * QT + qi]);
      }
    }
    __syncwarp();
    if (lane == 0 && issued < n_tiles) {
      fence_proxy_async();
      mbar_arrive_expect_tx(&my_bars[s], TILE_BYTES);
      bulk_g2s(my_ring + (size_t)s * TILE_FLOATS, packed + (pg * NC + pc) * (int64_t)TILE_FLOATS, TILE_BYTES,
               &my_bars[s]);
      ++issued;
      if (++pc == NC) {
        pc = 0;
        pg += gstep;
      }
    }
    if (++s == NS) {
      s = 0;
      parity ^= 1u;
    }
    ++c;
    if (c != NC1 && c != NC) continue;

    // ---- one feature block of row group g is complete (c == NC1: audio, c == NC: text)
    const int blk = (c == NC1) ? 0 : 1;
    TransposeReduce<R * QT, 16>::run(acc, lane);
    {
      const int r = lane >> 2;
      const int64_t row = g * R + r;
      if (row < W && (lane % DUP) == 0) {
        const double sqx = (blk == 0 ? sqnorm1 : sqnorm2)[row];
        const int label = labels[row];
        Pair* table = blk == 0 ? table1 : table2;
        if ((unsigned)label < (unsigned)KB) {
#pragma unroll
          for (int j = 0; j < PER_LANE; ++j) {
            const int idx = (R * QT >= 32) ? lane * PER_LANE + j : lane / DUP;
            const int qi = idx % QT;
            if (qi < nq) {
              const double dist = cosine_distance(acc[j], qn[blk * QT + qi], sqx);
              if (dist < kEmptyDist) {
                Pair mine;
                mine.d = (unsigned long long)__double_as_longlong(dist);
                mine.id = (unsigned long long)(id_offset + row);
                pair_min_global(&table[qi * KB + label], mine);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R * QT; ++i) acc[i] = 0.0;
    if (c == NC) {
      c = 0;
      g += gstep;
    }
  }
}

size_t fixed_smem2(int QT, int Dp, int ncw) {
  return (size_t)QT * Dp * 4 + (size_t)2 * QT * sizeof(double) + ((size_t)ncw * MAXNS2 + 1) * sizeof(uint64_t) +
         (size_t)ncw * sizeof(int);
}

template <int QT, int NCW>
int launch2(const float* packed, const double* sq1, const double* sq2, const int32_t* labels, int64_t W, int NC1,
            int NC, int64_t id_offset, const float* q, int nq, Pair* t1, Pair* t2, int pool, int grid, int reverse,
            cudaStream_t st) {
  const int64_t G = (W + R - 1) / R;
  const size_t smem = (size_t)pool * TILE_BYTES + fixed_smem2(QT, NC * DC, NCW);
  auto kern = cand_cosine2_kernel<QT, NCW>;
  QPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NCW * 32, smem, st>>>(packed, sq1, sq2, labels, W, NC1, NC, G, id_offset, q, nq, t1, t2, pool, reverse);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_cand_cosine2_minbycode(const float* packed, const double* row_sqnorm1, const double* row_sqnorm2,
                                          const int32_t* labels, int64_t W, int D1, int D2, int64_t id_offset,
                                          const float* q, int Q, qpg_pair_t* table1, qpg_pair_t* table2,
                                          void* stream) {
  QPG_CHECK_ARG(W >= 0 && D1 > 0 && D2 > 0 && Q >= 0, "W >= 0, D1 > 0, D2 > 0, Q >= 0");
  QPG_CHECK_ARG(D1 % DC == 0 && D2 % DC == 0, "both feature blocks must be multiples of 128 floats");
  if (W == 0 || Q == 0) return QPG_OK;
  QPG_CHECK_ARG(packed && row_sqnorm1 && row_sqnorm2 && labels && q && table1 && table2, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0,
                "packed must be 128-byte aligned, q 16-byte aligned");
  QPG_CHECK_ARG(((reinterpret_cast<uintptr_t>(table1) | reinterpret_cast<uintptr_t>(table2)) & 15) == 0,
                "tables must be 16-byte aligned");
  const int NC1 = D1 / DC, NC = (D1 + D2) / DC, Dp = NC * DC;
  const int ncw = 12;
  int qt = 4;
  while (qt > 1 && (qt / 2) >= Q) qt /= 2;
  while (qt > 1 && fixed_smem2(qt, Dp, ncw) + (size_t)2 * ncw * TILE_BYTES > kSmemLimit) qt /= 2;
  if (fixed_smem2(qt, Dp, ncw) + (size_t)2 * ncw * TILE_BYTES > kSmemLimit) {
    set_error("D1+D2=%d too large for the shared-memory query tile", Dp);
    return QPG_E_UNSUPPORTED;
  }
  int pool = (int)((kSmemLimit - fixed_smem2(qt, Dp, ncw)) / TILE_BYTES);
  if (pool > 3 * ncw) pool = 3 * ncw;
  const int64_t G = (W + R - 1) / R;
  int grid = sm_count();
  if (grid > G) grid = (int)G;
  Pair* t1 = reinterpret_cast<Pair*>(table1);
  Pair* t2 = reinterpret_cast<Pair*>(table2);
  cudaStream_t st = (cudaStream_t)stream;
  for (int q0 = 0; q0 < Q; q0 += qt) {
    const int nq = (Q - q0) < qt ? (Q - q0) : qt;
    const int rev = ((q0 / qt) & 1) ? 1 : 0;
    const float* qp = q + (size_t)q0 * Dp;
    int rc;
    if (qt == 4) rc = launch2<4, 12>(packed, row_sqnorm1, row_sqnorm2, labels, W, NC1, NC, id_offset, qp, nq,
                                     t1 + (size_t)q0 * KB, t2 + (size_t)q0 * KB, pool, grid, rev, st);
    else if (qt == 2) rc = launch2<2, 12>(packed, row_sqnorm1, row_sqnorm2, labels, W, NC1, NC, id_offset, qp, nq,
                                          t1 + (size_t)q0 * KB, t2 + (size_t)q0 * KB, pool, grid, rev, st);
    else rc = launch2<1, 12>(packed, row_sqnorm1, row_sqnorm2, labels, W, NC1, NC, id_offset, qp, nq,
                             t1 + (size_t)q0 * KB, t2 + (size_t)q0 * KB, pool, grid, rev, st);
    if (rc != QPG_OK) return rc;
  }
  return QPG_OK;
}
