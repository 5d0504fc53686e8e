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
// One CTA computes a 128 x BN output tile (BN <= 256, a multiple of 16):
//   warp 0   TMA producer      cp.async.bulk.tensor (A: 3-D, B: 2-D) -> 4-stage smem ring, mbarrier full/empty
//   warp 1   MMA issuer        one lane: 4 x tcgen05.mma.kind::tf32 (K = 8) per stage, tcgen05.commit frees the stage
//   warp 2   TMEM allocator    128 lanes x BN columns of float32 accumulators
//   warps 4-7 epilogue         tcgen05.ld 32x32b -> +bias, +residual -> raw and/or ReLU'd copy (the ReLU that
//                              ResConv1DBlock applies on load is applied by the PRODUCER of an activation)
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

// ===== epilogue: TMEM -> registers -> global (warps 4-7 of either kernel) =====
__device__ __forceinline__ void conv_epilogue(const TcParams& p, uint32_t tmem_base, uint64_t* accum_full, int warp,
                                              int lane, int b0, int t0, int n0, const float* __restrict__ bias,
                                              const float* __restrict__ residual, float* __restrict__ out,
                                              float* __restrict__ out_relu) {
  const int q = warp & 3;                       // TMEM lane quarter this warp may access
  const int r = q * 32 + lane;                  // row of the tile
  const int b_local = r / p.Tbox, t_local = r - b_local * p.Tbox;
  const bool row_ok = r < p.Bbox * p.Tbox && (b0 + b_local) < p.B && (t0 + t_local) < p.n_out;
  const size_t row_base = ((size_t)(b0 + b_local) * p.out_rows_per_item + (t0 + t_local)) * (size_t)p.out_ld +
                          p.out_chan_off;
  const bool vec = (p.out_ld & 3) == 0 && (p.out_chan_off & 3) == 0;
  mbar_wait(accum_full, 0);
  tc_fence_after();
  for (int cc = 0; cc < p.BN; cc += 32) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
    if (!row_ok) continue;
    const int ncol = n0 + cc;                   // first output channel of this chunk
    if (vec && ncol + 32 <= p.C_out) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3]));
        if (bias) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + ncol + j);
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        if (residual) {
          const float4 rr = *reinterpret_cast<const float4*>(residual + row_base + ncol + j);
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (out) *reinterpret_cast<float4*>(out + row_base + ncol + j) = o;
        if (out_relu)
          *reinterpret_cast<float4*>(out_relu + row_base + ncol + j) =
              make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
      }
    } else {
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
  }
}

// ---- kernel -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
    conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcParams p,
                   const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ out,
                   float* __restrict__ out_relu) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* smem_a = smem;                                  // [STAGES][16 KiB]
  unsigned char* smem_b = smem + STAGES * A_BYTES;               // [STAGES][32 KiB]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, n0 = blockIdx.y * p.BN;
  const int tb = tile_m / p.tiles_t, tt = tile_m - tb * p.tiles_t;
  const int b0 = tb * p.Bbox, t0 = tt * p.Tbox;
  const int n_iter = p.n_taps * p.kblocks;
  const uint32_t tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
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
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint64_t a_desc = umma_desc_sw128(smem_u32(smem_a + s * A_BYTES));
        const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          tc_mma_tf32(tmem_base, a_desc + 2u * k, b_desc + 2u * k, idesc, (it | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty[s]);      // arrives when the MMAs above have finished reading this stage
      }
      tc_commit(accum_full);       // accumulator complete
    }
  } else if (warp >= 4) {
    conv_epilogue(p, tmem_base, accum_full, warp, lane, b0, t0, n0, bias, residual, out, out_relu);
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
                    const __grid_constant__ CUtensorMap map_b_lo, TcParams p, const float* __restrict__ bias,
                    const float* __restrict__ residual, float* __restrict__ out, float* __restrict__ out_relu) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SPLIT_STAGES * SPLIT_STAGE_BYTES);
  uint64_t* empty = full + SPLIT_STAGES;
  uint64_t* split = empty + SPLIT_STAGES;
  uint64_t* accum_full = split + SPLIT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, n0 = blockIdx.y * p.BN;
  const int tb = tile_m / p.tiles_t, tt = tile_m - tb * p.tiles_t;
  const int b0 = tb * p.Bbox, t0 = tt * p.Tbox;
  const int n_iter = p.n_taps * p.kblocks;
  const uint32_t tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : 128;

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
    mbar_init(accum_full, 1);
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
    if (lane == 0) {                                       // ===== TMA producer =====
      const uint32_t a_bytes = (uint32_t)p.Bbox * p.Tbox * 128u, b_bytes = (uint32_t)p.BN * 128u;
      int it = 0;
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
  } else if (warp == 1) {
    if (lane == 0) {                                       // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % SPLIT_STAGES;
        mbar_wait(&split[s], (it / SPLIT_STAGES) & 1);      // TMA landed AND the activation tile is split
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * SPLIT_STAGE_BYTES);
        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + A_BYTES);
        const uint64_t b_hi = umma_desc_sw128(st + 2 * A_BYTES), b_lo = umma_desc_sw128(st + 3 * A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          tc_mma_tf32(tmem_base, a_lo + 2u * k, b_hi + 2u * k, idesc, (it | k) != 0 ? 1u : 0u);   // small terms first
          tc_mma_tf32(tmem_base, a_hi + 2u * k, b_lo + 2u * k, idesc, 1u);
          tc_mma_tf32(tmem_base, a_hi + 2u * k, b_hi + 2u * k, idesc, 1u);
        }
        tc_commit(&empty[s]);
      }
      tc_commit(accum_full);
    }
  } else if (warp >= 8) {
    // ===== splitter: A -> (A_hi in place, A_lo) for the whole 16 KiB tile; elementwise, so the swizzle is irrelevant
    const int t = threadIdx.x - 256;                       // 0..127
    for (int it = 0; it < n_iter; ++it) {
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
    conv_epilogue(p, tmem_base, accum_full, warp, lane, b0, t0, n0, bias, residual, out, out_relu);
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
  dim3 grid((unsigned)(tiles_b * p.tiles_t), (unsigned)(d->N_pad / d->BN));
  // the shared-memory attribute is per device: set it on every launch (a process may use several GPUs)
  if (split) {
    const size_t smem = (size_t)SPLIT_STAGES * SPLIT_STAGE_BYTES + (3 * SPLIT_STAGES + 1) * sizeof(uint64_t) + 16;
    QPG_CUDA(cudaFuncSetAttribute(conv_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc3_kernel<<<grid, 384, smem, (cudaStream_t)stream>>>(map_a, map_b, map_b_lo, p, bias, residual, out, out_relu);
  } else {
    const size_t smem = (size_t)STAGES * (A_BYTES + B_BYTES) + (2 * STAGES + 1) * sizeof(uint64_t) + 16;
    QPG_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(map_a, map_b, p, bias, residual, out, out_relu);
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
