// Tensor-core path of the VQ-VAE convolution stacks: implicit GEMM on tcgen05
// (5th-gen tensor cores, kind::tf32, FP32 accumulation in TMEM) fed by TMA.
//
//   out[b, t, n] = bias[n] + residual[b, t, n]
//                + sum_{tap} sum_{k < C_in} X[b, t + row_off[tap], chan_off[tap] + k] * W[tap][n][k]
//
// X is a *view* [B, T_view, C_view] of a channels-last float32 activation: stride-2
// convolutions and the two output phases of ConvTranspose1d(k4,s2,p1) are expressed on
// the paired-frame view [B, T/2, 2C] through per-tap channel offsets, so every operand
// tile is a dense TMA box and the conv zero padding is TMA's out-of-bounds zero fill
// (3-D tensor map, box [Bbox items][Tbox frames][32 floats], SWIZZLE_128B).
//
// Persistent CTAs (one per SM) walk the 128 x BN output tiles (BN <= 256, a multiple of 16) round-robin:
//   warp 0    TMA producer      cp.async.bulk.tensor (A: 3-D, B: 2-D) -> 4-stage smem ring, mbarrier full/empty,
//                               running ahead across tile boundaries
//   warp 1    MMA issuer        one lane: 4 x tcgen05.mma.kind::tf32 (K = 8) per stage, tcgen05.commit frees the stage
//   warp 2    TMEM allocator    TWO accumulators of 128 lanes x BN float32 columns: the epilogue of tile i drains one
//                               while the MMAs of tile i + 1 fill the other
//   warps 4-11 epilogue         tcgen05.ld 32x32b -> shared-memory transpose -> +bias, +residual (prefetched one chunk
//                               ahead) -> raw and/or ReLU'd copy in full 128-byte lines (the ReLU that ResConv1DBlock
//                               applies on load is applied by the PRODUCER of an activation)
// Replaces nn.Conv1d / nn.ConvTranspose1d at encdec.py:20,24,39,45,113 and resnet.py:33-36 in
// "fast" mode (TF32 operands, ~1e-3 relative); the float32 FFMA path of conv1d.cu stays the parity mode.
#include <cuda.h>

#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int BM = 128;          // output rows per CTA (TMEM lanes)
constexpr int BK = 32;           // floats per k-block = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int MAX_BN = 256;
constexpr int A_BYTES = BM * 128;       // 16 KiB
constexpr int B_BYTES = MAX_BN * 128;   // 32 KiB
constexpr int UMMA_K = 8;               // tf32

struct TcParams {
  int B, n_out, C_in, C_out, n_taps;
  int row_off[4], chan_off[4];
  int Tbox, Bbox, tiles_t;
  int BN, N_pad, kblocks;
  int out_rows_per_item, out_ld, out_chan_off;
};

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile whose rows are 128 bytes: 8-row groups are 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused with 128B swizzle), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row groups, bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100), bits [46,48)
  d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ===== epilogue: TMEM -> registers -> shared-memory transpose -> global =====
// tcgen05.ld 32x32b hands lane l row l of the warp's 32-row TMEM quarter.  Stored straight from there every float4
// access of a warp touches 32 different 128-byte lines, and the LSU spends one tag cycle per line: for the 1x1
// convolutions of the residual blocks (read residual, write raw + ReLU'd copy: 12 KiB per 32-column chunk and warp)
// that alone was 26 us of a 52 us kernel (ncu: 32 sectors per request, l1tex 48 % busy, lg_throttle stalls, tensor
// pipe 18 %).  So each 32 x 32 chunk goes through a padded shared-memory tile and leaves transposed: 8 lanes cover
// one row's 128 bytes, a warp instruction covers four full lines (8x fewer tag cycles), the residual is read the
// same way, and it is fetched one chunk AHEAD so that its latency sits under the stores of the current chunk.
// The tile's column chunks are dealt round-robin to `n_groups` groups of four warps.
constexpr int STG_LD = 33;                       // floats per staged row: conflict-free both ways
constexpr int STG_FLOATS_PER_WARP = 32 * STG_LD;
__device__ __forceinline__ void conv_epilogue(const TcParams& p, uint32_t acc_tmem, uint64_t* accum_full,
                                              uint32_t parity, float* __restrict__ stg, int group, int n_groups,
                                              int warp, int lane, int b0, int t0, int n0,
                                              const float* __restrict__ bias, const float* __restrict__ residual,
                                              float* __restrict__ out, float* __restrict__ out_relu) {
  const int q = warp & 3;                       // TMEM lane quarter this warp may access
  const bool vec = (p.out_ld & 3) == 0 && (p.out_chan_off & 3) == 0;
  const int step = 32 * n_groups;
  const int c4 = (lane & 7) * 4;
  // transposed pass: this lane handles columns c4..c4+3 of rows i * 4 + (lane >> 3), i = 0..7
  size_t trow_base[8];
  unsigned tok = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = q * 32 + i * 4 + (lane >> 3);
    const int b_local = r / p.Tbox, t_local = r - b_local * p.Tbox;
    const bool ok = r < p.Bbox * p.Tbox && (b0 + b_local) < p.B && (t0 + t_local) < p.n_out;
    tok |= ok ? (1u << i) : 0u;
    trow_base[i] = ((size_t)(b0 + b_local) * p.out_rows_per_item + (t0 + t_local)) * (size_t)p.out_ld + p.out_chan_off;
  }
  // the row this lane reads from TMEM (the scalar path for ragged channel counts stores it directly)
  const int r = q * 32 + lane;
  const int b_local = r / p.Tbox, t_local = r - b_local * p.Tbox;
  const bool row_ok = r < p.Bbox * p.Tbox && (b0 + b_local) < p.B && (t0 + t_local) < p.n_out;
  const size_t row_base = ((size_t)(b0 + b_local) * p.out_rows_per_item + (t0 + t_local)) * (size_t)p.out_ld +
                          p.out_chan_off;
  float4 rr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch_residual = [&](int cc) {
    const int ncol = n0 + cc;
    if (residual && cc < p.BN && vec && ncol + 32 <= p.C_out) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if ((tok >> i) & 1u) rr[i] = *reinterpret_cast<const float4*>(residual + trow_base[i] + ncol + c4);
    }
  };
  fetch_residual(32 * group);                   // independent of the accumulator: issued before the wait
  mbar_wait(accum_full, parity);
  tc_fence_after();
  for (int cc = 32 * group; cc < p.BN; cc += step) {
    uint32_t v[32];
    __syncwarp();                               // the previous chunk has left the staging tile
    tmem_ld_x32(acc_tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
    const int ncol = n0 + cc;                   // first output channel of this chunk
    if (vec && ncol + 32 <= p.C_out) {
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * STG_LD + j] = __uint_as_float(v[j]);
      __syncwarp();
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) bb = *reinterpret_cast<const float4*>(bias + ncol + c4);
      float4 o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* sr = stg + (i * 4 + (lane >> 3)) * STG_LD + c4;
        o[i] = make_float4(sr[0] + bb.x + rr[i].x, sr[1] + bb.y + rr[i].y, sr[2] + bb.z + rr[i].z,
                           sr[3] + bb.w + rr[i].w);
      }
      fetch_residual(cc + step);                // next chunk's residual is in flight while this one is stored
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if ((tok >> i) & 1u) {
          const size_t at = trow_base[i] + ncol + c4;
          if (out) *reinterpret_cast<float4*>(out + at) = o[i];
          if (out_relu)
            *reinterpret_cast<float4*>(out_relu + at) =
                make_float4(fmaxf(o[i].x, 0.f), fmaxf(o[i].y, 0.f), fmaxf(o[i].z, 0.f), fmaxf(o[i].w, 0.f));
        }
      }
    } else {
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = ncol + j;
          if (n < p.C_out) {
            float o = __uint_as_float(v[j]);
            if (bias) o += bias[n];
            if (residual) o += residual[row_base + n];
            if (out) out[row_base + n] = o;
            if (out_relu) out_relu[row_base + n] = fmaxf(o, 0.f);
          }
        }
      }
      fetch_residual(cc + step);
    }
  }
}

// ---- kernel -----------------------------------------------------------------------------------
// Persistent: CTA c takes tiles c, c + gridDim.x, ... (n fastest, so CTAs working at the same time share the
// activation rows in L2).  Two TMEM accumulators: the epilogue of tile i (TMEM -> +bias, +residual -> global, the
// memory-bound half of a 1x1 convolution) runs under the MMAs of tile i + 1; the TMA ring keeps running across tiles.
__global__ void __launch_bounds__(384, 1)
    conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcParams p,
                   int tiles_n, int n_tiles, const float* __restrict__ bias, const float* __restrict__ residual,
                   float* __restrict__ out, float* __restrict__ out_relu) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* smem_a = smem;                                  // [STAGES][16 KiB]
  unsigned char* smem_b = smem + STAGES * A_BYTES;               // [STAGES][32 KiB]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;                           // [2]
  uint64_t* acc_empty = acc_full + 2;                            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* staging = reinterpret_cast<float*>(tmem_slot + 4);      // [8 warps][32][STG_LD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_iter = p.n_taps * p.kblocks;
  const uint32_t acc_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
  const uint32_t tmem_cols = 2 * acc_cols;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);           // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)p.Bbox * p.Tbox * 128u, b_bytes = (uint32_t)p.BN * 128u;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tile_m = tile / tiles_n, n0 = (tile - tile_m * tiles_n) * p.BN;
        const int tb = tile_m / p.tiles_t, tt = tile_m - tb * p.tiles_t;
        const int b0 = tb * p.Bbox, t0 = tt * p.Tbox;
        for (int tap = 0; tap < p.n_taps; ++tap) {
          for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
            mbar_arrive_expect_tx(&full[s], a_bytes + b_bytes);
            tma_load_3d(smem_a + s * A_BYTES, &map_a, &full[s], p.chan_off[tap] + kb * BK, t0 + p.row_off[tap], b0);
            tma_load_2d(smem_b + s * B_BYTES, &map_b, &full[s], kb * BK, tap * p.N_pad + n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;
        if (lt >= 2) {                         // the epilogue of the tile two back has drained this accumulator
          mbar_wait(&acc_empty[as], ((lt >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + (uint32_t)as * acc_cols;
        for (int i = 0; i < n_iter; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128(smem_u32(smem_a + s * A_BYTES));
          const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            tc_mma_tf32(acc, a_desc + 2u * k, b_desc + 2u * k, idesc, (i | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty[s]);      // arrives when the MMAs above have finished reading this stage
        }
        tc_commit(&acc_full[as]);    // accumulator complete
      }
    }
  } else if (warp >= 4) {
    int lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int tile_m = tile / tiles_n, n0 = (tile - tile_m * tiles_n) * p.BN;
      const int tb = tile_m / p.tiles_t, tt = tile_m - tb * p.tiles_t;
      const int as = lt & 1;
      conv_epilogue(p, tmem_base + (uint32_t)as * acc_cols, &acc_full[as], (uint32_t)((lt >> 1) & 1),
                    staging + (warp - 4) * STG_FLOATS_PER_WARP, (warp - 4) >> 2, 2, warp, lane, tb * p.Bbox, tt * p.Tbox,
                    n0, bias, residual, out, out_relu);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- 3xTF32: float32-accurate products on the TF32 tensor cores ------------------------------------------
// x*w ~ x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with x_hi = x truncated to TF32 (11 significant bits) and x_lo = x - x_hi
// (exact in float32; its own truncation to TF32 and the dropped x_lo*w_lo are both ~2^-22 relative).  The weights
// arrive pre-split from the host (w_hi, w_lo: two tensor maps); the activation tile is split IN SHARED MEMORY by
// four extra warps (8-11) between the TMA arrival and the MMAs: A is masked in place to its TF32 part and
// A - A_hi goes to a second buffer, so activations stay single float32 arrays in HBM and the epilogue is
// unchanged.  Three MMAs per k-step into the same accumulator.  Stage = A, A_lo, B_hi, B_lo (16 KiB each, BN <= 128),
// three stages.
constexpr int SPLIT_STAGES = 3;
constexpr int SPLIT_BN = 128;
constexpr int SPLIT_STAGE_BYTES = 4 * A_BYTES;

__global__ void __launch_bounds__(384, 1)
    conv_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_b_lo, TcParams p, int tiles_n, int n_tiles,
                    const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ out,
                    float* __restrict__ out_relu) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SPLIT_STAGES * SPLIT_STAGE_BYTES);
  uint64_t* empty = full + SPLIT_STAGES;
  uint64_t* split = empty + SPLIT_STAGES;
  uint64_t* acc_full = split + SPLIT_STAGES;                     // [2]
  uint64_t* acc_empty = acc_full + 2;                            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* staging = reinterpret_cast<float*>(tmem_slot + 2);      // [4 warps][32][STG_LD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_iter = p.n_taps * p.kblocks;
  const uint32_t acc_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : 128;
  const uint32_t tmem_cols = 2 * acc_cols;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SPLIT_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&split[s], 4);              // one arrive per splitter warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);          // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // persistent like conv_tc_kernel: tiles blockIdx.x, + gridDim.x, ...; the epilogue of one tile runs under the
  // (three times longer) main loop of the next
  if (warp == 0) {
    if (lane == 0) {                                       // ===== TMA producer =====
      const uint32_t a_bytes = (uint32_t)p.Bbox * p.Tbox * 128u, b_bytes = (uint32_t)p.BN * 128u;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tile_m = tile / tiles_n, n0 = (tile - tile_m * tiles_n) * p.BN;
        const int tb = tile_m / p.tiles_t, tt = tile_m - tb * p.tiles_t;
        const int b0 = tb * p.Bbox, t0 = tt * p.Tbox;
        for (int tap = 0; tap < p.n_taps; ++tap) {
          for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
            const int s = it % SPLIT_STAGES;
            if (it >= SPLIT_STAGES) mbar_wait(&empty[s], ((it / SPLIT_STAGES) - 1) & 1);
            unsigned char* st = smem + s * SPLIT_STAGE_BYTES;
            mbar_arrive_expect_tx(&full[s], a_bytes + 2 * b_bytes);
            tma_load_3d(st, &map_a, &full[s], p.chan_off[tap] + kb * BK, t0 + p.row_off[tap], b0);
            tma_load_2d(st + 2 * A_BYTES, &map_b, &full[s], kb * BK, tap * p.N_pad + n0);
            tma_load_2d(st + 3 * A_BYTES, &map_b_lo, &full[s], kb * BK, tap * p.N_pad + n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                       // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;
        if (lt >= 2) {
          mbar_wait(&acc_empty[as], ((lt >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + (uint32_t)as * acc_cols;
        for (int i = 0; i < n_iter; ++i, ++it) {
          const int s = it % SPLIT_STAGES;
          mbar_wait(&split[s], (it / SPLIT_STAGES) & 1);    // TMA landed AND the activation tile is split
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * SPLIT_STAGE_BYTES);
          const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + A_BYTES);
          const uint64_t b_hi = umma_desc_sw128(st + 2 * A_BYTES), b_lo = umma_desc_sw128(st + 3 * A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            tc_mma_tf32(acc, a_lo + 2u * k, b_hi + 2u * k, idesc, (i | k) != 0 ? 1u : 0u);   // small terms first
            tc_mma_tf32(acc, a_hi + 2u * k, b_lo + 2u * k, idesc, 1u);
            tc_mma_tf32(acc, a_hi + 2u * k, b_hi + 2u * k, idesc, 1u);
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full[as]);
      }
    }
  } else if (warp >= 8) {
    // ===== splitter: A -> (A_hi in place, A_lo) for the whole 16 KiB tile; elementwise, so the swizzle is irrelevant
    const int t = threadIdx.x - 256;                       // 0..127
    int n_local = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_local;
    const int total = n_local * n_iter;
    for (int it = 0; it < total; ++it) {
      const int s = it % SPLIT_STAGES;
      mbar_wait(&full[s], (it / SPLIT_STAGES) & 1);
      float4* a = reinterpret_cast<float4*>(smem + s * SPLIT_STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(smem + s * SPLIT_STAGE_BYTES + A_BYTES);
#pragma unroll
      for (int j = 0; j < A_BYTES / 16 / 128; ++j) {
        const int i = j * 128 + t;
        const float4 v = a[i];
        float4 h;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        a[i] = h;
        lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
      fence_proxy_async();                                 // generic-proxy writes -> visible to the tensor core reads
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&split[s])) : "memory");
      }
    }
  } else if (warp >= 4) {
    int lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int tile_m = tile / tiles_n, n0 = (tile - tile_m * tiles_n) * p.BN;
      const int tb = tile_m / p.tiles_t, tt = tile_m - tb * p.tiles_t;
      const int as = lt & 1;
      conv_epilogue(p, tmem_base + (uint32_t)as * acc_cols, &acc_full[as], (uint32_t)((lt >> 1) & 1),
                    staging + (warp - 4) * STG_FLOATS_PER_WARP, 0, 1, warp, lane, tb * p.Bbox, tt * p.Tbox, n0, bias,
                    residual, out, out_relu);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- host: tensor maps --------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

}  // namespace
}  // namespace qpg

using namespace qpg;

static int conv_tc_launch(const qpg_conv_tc_desc_t* d, const float* in, const float* w, const float* w_lo,
                          const float* bias, const float* residual, float* out, float* out_relu, void* stream) {
  const bool split = w_lo != nullptr;
  QPG_CHECK_ARG(d != nullptr, "null descriptor");
  QPG_CHECK_ARG(d->B >= 0 && d->T_view > 0 && d->C_view > 0 && d->C_in > 0 && d->C_out > 0, "bad shape");
  QPG_CHECK_ARG(d->n_taps >= 1 && d->n_taps <= 4 && d->n_out >= 0, "n_taps in 1..4");
  QPG_CHECK_ARG((d->C_view & 3) == 0 && (d->K_pad & 3) == 0 && d->K_pad >= d->C_in,
                "C_view and K_pad must be multiples of 4 floats (16-byte TMA strides)");
  QPG_CHECK_ARG(d->BN >= 16 && d->BN <= (split ? SPLIT_BN : MAX_BN) && (d->BN & 15) == 0 && d->N_pad % d->BN == 0 &&
                    d->N_pad >= d->C_out,
                "BN multiple of 16 <= 256 (<= 128 for 3xTF32), N_pad multiple of BN");
  if (d->B == 0 || d->n_out == 0) return QPG_OK;
  QPG_CHECK_ARG(in && w && (out || out_relu), "null pointer");
  QPG_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(w_lo)) & 15) == 0,
                "16-byte alignment");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return QPG_E_CUDA;
  }
  TcParams p;
  p.B = d->B; p.n_out = d->n_out; p.C_in = d->C_in; p.C_out = d->C_out; p.n_taps = d->n_taps;
  for (int i = 0; i < 4; ++i) {
    p.row_off[i] = d->row_offset[i];
    p.chan_off[i] = d->chan_offset[i];
  }
  p.Tbox = d->n_out < BM ? d->n_out : BM;
  p.Bbox = BM / p.Tbox;
  if (p.Bbox > d->B) p.Bbox = d->B;
  if (p.Bbox > 256) p.Bbox = 256;
  p.tiles_t = (d->n_out + p.Tbox - 1) / p.Tbox;
  p.BN = d->BN; p.N_pad = d->N_pad;
  p.kblocks = (d->C_in + BK - 1) / BK;
  p.out_rows_per_item = d->out_rows_per_item; p.out_ld = d->out_ld; p.out_chan_off = d->out_chan_offset;

  CUtensorMap map_a, map_b, map_b_lo;
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->C_view, (cuuint64_t)d->T_view, (cuuint64_t)d->B};
    cuuint64_t strides[2] = {(cuuint64_t)d->C_view * 4, (cuuint64_t)d->C_view * 4 * (cuuint64_t)d->T_view};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)p.Tbox, (cuuint32_t)p.Bbox};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)rc);
      return QPG_E_CUDA;
    }
  }
  for (int which = 0; which < (split ? 2 : 1); ++which) {
    cuuint64_t dims[2] = {(cuuint64_t)d->K_pad, (cuuint64_t)d->n_taps * (cuuint64_t)d->N_pad};
    cuuint64_t strides[1] = {(cuuint64_t)d->K_pad * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)d->BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(which == 0 ? &map_b : &map_b_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<float*>(which == 0 ? w : w_lo), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)rc);
      return QPG_E_CUDA;
    }
  }
  const int tiles_b = (d->B + p.Bbox - 1) / p.Bbox;
  const int tiles_n = d->N_pad / d->BN, n_tiles = tiles_b * p.tiles_t * tiles_n;
  const int ctas = n_tiles < sm_count() ? n_tiles : sm_count();          // persistent CTAs, round-robin over the tiles
  // the shared-memory attribute is per device: set it on every launch (a process may use several GPUs)
  if (split) {
    const size_t smem = (size_t)SPLIT_STAGES * SPLIT_STAGE_BYTES + (3 * SPLIT_STAGES + 4) * sizeof(uint64_t) + 16 +
                        4 * STG_FLOATS_PER_WARP * sizeof(float);
    QPG_CUDA(cudaFuncSetAttribute(conv_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc3_kernel<<<ctas, 384, smem, (cudaStream_t)stream>>>(map_a, map_b, map_b_lo, p, tiles_n, n_tiles, bias, residual,
                                                              out, out_relu);
  } else {
    const size_t smem = (size_t)STAGES * (A_BYTES + B_BYTES) + (2 * STAGES + 4) * sizeof(uint64_t) + 16 +
                        8 * STG_FLOATS_PER_WARP * sizeof(float);
    QPG_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<<<ctas, 384, smem, (cudaStream_t)stream>>>(map_a, map_b, p, tiles_n, n_tiles, bias, residual, out,
                                                             out_relu);
  }
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_conv1d_taps_tf32(const qpg_conv_tc_desc_t* d, const float* in, const float* w, const float* bias,
                                    const float* residual, float* out, float* out_relu, void* stream) {
  return conv_tc_launch(d, in, w, nullptr, bias, residual, out, out_relu, stream);
}

extern "C" int qpg_conv1d_taps_3xtf32(const qpg_conv_tc_desc_t* d, const float* in, const float* w_hi, const float* w_lo,
                                      const float* bias, const float* residual, float* out, float* out_relu,
                                      void* stream) {
  if (w_lo == nullptr) {
    set_error("bad argument: w_lo is null");
    return QPG_E_BADARG;
  }
  return conv_tc_launch(d, in, w_hi, w_lo, bias, residual, out, out_relu, stream);
}
