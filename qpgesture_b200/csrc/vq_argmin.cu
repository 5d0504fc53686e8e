// VQ codebook L2 argmin (BottleneckBlock.quantise, codebook/models/bottleneck.py:120-126)
// and embedding lookup (dequantise, :128-130).
//
//   dist[m][k] = fl32( fl32(|x_m|^2 - 2*<x_m, c_k>) + |c_k|^2 ),   idx[m] = first argmin_k
//
// The reference evaluates this in float32 through an SGEMM whose summation order is
// library defined; here every inner product is accumulated in float64 and rounded once
// (the correctly rounded value any float32 summation order approximates), then the
// reference's two float32 roundings are applied.  The [M, K] distance matrix is never
// written to memory.  Bound: FP64 pipe (2*M*K*D flop); M is small on the matcher's path
// (30 latents per 4-s sequence).
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int TM = 16;        // latents per CTA
constexpr int THREADS = 128;  // one code per thread; the codebook is split over gridDim.y CTAs

// order-preserving map float -> uint32 (handles the slightly negative distances rounding can produce)
__device__ __forceinline__ uint32_t ordered_u32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered_f32(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void vq_key_init_kernel(unsigned long long* __restrict__ keys, int64_t M) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < M) keys[i] = ~0ull;
}

__global__ void vq_key_finish_kernel(int64_t* __restrict__ idx_io, float* __restrict__ min_out, int64_t M) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const unsigned long long key = (unsigned long long)idx_io[i];
  if (min_out) min_out[i] = unordered_f32((uint32_t)(key >> 32));
  idx_io[i] = (int64_t)(key & 0xffffffffull);
}

// keys[m] = min over codes of (ordered(dist) << 32 | k): lexicographic (distance, code index) -> first minimum.
// The latent tile is converted to float64 once in shared memory, so the inner loop is 1 F2F per 16 DFMA.
__global__ void __launch_bounds__(THREADS)
    vq_argmin_kernel(const float* __restrict__ x, const float* __restrict__ cb, int64_t M, int D, int K,
                     unsigned long long* __restrict__ keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw);           // [TM][Dp] float64 copies of the latents
  float* xn = reinterpret_cast<float*>(xs + (size_t)TM * ((D + 3) & ~3));   // [TM] squared norms (float32)
  const int Dp = (D + 3) & ~3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * TM;

  for (int i = tid; i < TM * Dp; i += THREADS) {
    const int r = i / Dp, d = i - r * Dp;
    const int64_t m = m0 + r;
    xs[i] = (m < M && d < D) ? (double)x[m * D + d] : 0.0;
  }
  __syncthreads();
  for (int r = warp; r < TM; r += THREADS / 32) {
    double s = 0.0;
    for (int d = lane; d < Dp; d += 32) s = fma(xs[r * Dp + d], xs[r * Dp + d], s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) xn[r] = (float)s;
  }
  __syncthreads();

  const int k = blockIdx.y * THREADS + tid;
  if (k < K) {
    const float* row = cb + (size_t)k * D;
    double acc[TM], kn = 0.0;
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r] = 0.0;
    const bool vec = (D & 3) == 0;
    for (int d = 0; d < Dp; d += 4) {
      float4 cv;
      if (vec) {
        cv = *reinterpret_cast<const float4*>(row + d);
      } else {
        cv.x = d < D ? row[d] : 0.f;
        cv.y = d + 1 < D ? row[d + 1] : 0.f;
        cv.z = d + 2 < D ? row[d + 2] : 0.f;
        cv.w = d + 3 < D ? row[d + 3] : 0.f;
      }
      const double c0 = cv.x, c1 = cv.y, c2 = cv.z, c3 = cv.w;
      kn = fma(c0, c0, kn);
      kn = fma(c1, c1, kn);
      kn = fma(c2, c2, kn);
      kn = fma(c3, c3, kn);
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const double2 x01 = *reinterpret_cast<const double2*>(xs + (size_t)r * Dp + d);      // broadcast reads
        const double2 x23 = *reinterpret_cast<const double2*>(xs + (size_t)r * Dp + d + 2);
        double a = acc[r];
        a = fma(x01.x, c0, a);
        a = fma(x01.y, c1, a);
        a = fma(x23.x, c2, a);
        a = fma(x23.y, c3, a);
        acc[r] = a;
      }
    }
    const float knf = (float)kn;
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      if (m0 + r < M) {
        // the reference's float32 formula on a correctly rounded inner product (bottleneck.py:123)
        const float dist = __fadd_rn(__fsub_rn(xn[r], __fmul_rn(2.0f, (float)acc[r])), knf);
        const unsigned long long key = ((unsigned long long)ordered_u32(dist) << 32) | (unsigned long long)k;
        atomicMin(&keys[m0 + r], key);
      }
    }
  }
}

__global__ void dequantise_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb, int64_t M, int D,
                                  int K, float* __restrict__ out) {
  const int64_t m = blockIdx.x;
  int64_t k = idx[m];
  if (k < 0) k = 0;
  if (k >= K) k = K - 1;
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[m * D + d] = cb[k * D + d];
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_vq_argmin_f32(const float* x, const float* codebook, int64_t M, int D, int K, int64_t* idx_out,
                                 float* min_out, void* stream) {
  QPG_CHECK_ARG(M >= 0 && D > 0 && K > 0, "M >= 0, D > 0, K > 0");
  if (M == 0) return QPG_OK;
  QPG_CHECK_ARG(x && codebook && idx_out, "null pointer");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, "codebook must be 16-byte aligned");
  const int Dp = (D + 3) & ~3;
  const size_t smem = (size_t)TM * Dp * sizeof(double) + TM * sizeof(float) + 64;
  QPG_CHECK_ARG(smem <= 200 * 1024, "D too large");
  cudaStream_t st = (cudaStream_t)stream;
  // idx_out doubles as the 64-bit key buffer (ordered distance << 32 | code) until the finish kernel
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(idx_out);
  vq_key_init_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(keys, M);
  QPG_LAUNCH_CHECK();
  // the attribute is per device: set it on every launch (a process may use several GPUs)
  QPG_CUDA(cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((K + THREADS - 1) / THREADS));
  vq_argmin_kernel<<<grid, THREADS, smem, st>>>(x, codebook, M, D, K, keys);
  QPG_LAUNCH_CHECK();
  vq_key_finish_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(idx_out, min_out, M);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_vq_dequantise_f32(const int64_t* idx, const float* codebook, int64_t M, int D, int K, float* out,
                                     void* stream) {
  QPG_CHECK_ARG(M >= 0 && D > 0 && K > 0, "M >= 0, D > 0, K > 0");
  if (M == 0) return QPG_OK;
  QPG_CHECK_ARG(idx && codebook && out, "null pointer");
  dequantise_kernel<<<(unsigned)M, 128, 0, (cudaStream_t)stream>>>(idx, codebook, M, D, K, out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
