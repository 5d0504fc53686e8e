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
constexpr int THREADS = 128;  // each thread owns K/THREADS codes

__global__ void sqnorm_rows_kernel(const float* __restrict__ a, int64_t n, int D, float* __restrict__ out) {
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int d = lane; d < D; d += 32) {
    const double v = (double)a[row * D + d];
    s = fma(v, v, s);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = (float)s;
}

template <int CPT>  // codes per thread
__global__ void __launch_bounds__(THREADS)
    vq_argmin_kernel(const float* __restrict__ x, const float* __restrict__ cb, int64_t M, int D, int K,
                     int64_t* __restrict__ idx_out, float* __restrict__ min_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);             // [TM][D]
  float* xn = xs + (size_t)TM * D;                            // [TM] squared norms (float32)
  float* red_v = xn + TM;                                     // [TM][THREADS/32]
  int* red_i = reinterpret_cast<int*>(red_v + TM * (THREADS / 32));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * TM;

  for (int i = tid; i < TM * D; i += THREADS) {
    const int64_t m = m0 + i / D;
    xs[i] = m < M ? x[m * D + (i % D)] : 0.f;
  }
  __syncthreads();
  for (int r = warp; r < TM; r += THREADS / 32) {
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double v = (double)xs[r * D + d];
      s = fma(v, v, s);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) xn[r] = (float)s;
  }
  __syncthreads();

  float best_v[TM];
  int best_i[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    best_v[r] = INFINITY;
    best_i[r] = 0x7fffffff;
  }
  // codes k = tid + c*THREADS (ascending in c, so a strict < keeps the first minimum per thread)
  for (int c = 0; c < CPT; ++c) {
    const int k = tid + c * THREADS;
    if (k >= K) break;
    const float* row = cb + (size_t)k * D;
    double acc[TM], kn = 0.0;
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r] = 0.0;
    for (int d = 0; d < D; d += 4) {
      float4 cv;
      if (d + 3 < D && ((D & 3) == 0)) {
        cv = *reinterpret_cast<const float4*>(row + d);
      } else {
        cv.x = row[d];
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
        const float* xr = xs + r * D + d;
        double a = acc[r];
        a = fma((double)xr[0], c0, a);
        if (d + 1 < D) a = fma((double)xr[1], c1, a);
        if (d + 2 < D) a = fma((double)xr[2], c2, a);
        if (d + 3 < D) a = fma((double)xr[3], c3, a);
        acc[r] = a;
      }
    }
    const float knf = (float)kn;
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const float dotf = (float)acc[r];
      const float dist = __fadd_rn(__fsub_rn(xn[r], __fmul_rn(2.0f, dotf)), knf);
      if (dist < best_v[r]) {
        best_v[r] = dist;
        best_i[r] = k;
      }
    }
  }
  // block argmin per latent: (value, index) lexicographic -> first minimum overall
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    float v = best_v[r];
    int i = best_i[r];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (ov < v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
      }
    }
    if (lane == 0) {
      red_v[r * (THREADS / 32) + warp] = v;
      red_i[r * (THREADS / 32) + warp] = i;
    }
  }
  __syncthreads();
  if (tid < TM && m0 + tid < M) {
    float v = red_v[tid * (THREADS / 32)];
    int i = red_i[tid * (THREADS / 32)];
    for (int w = 1; w < THREADS / 32; ++w) {
      const float ov = red_v[tid * (THREADS / 32) + w];
      const int oi = red_i[tid * (THREADS / 32) + w];
      if (ov < v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
      }
    }
    idx_out[m0 + tid] = i;
    if (min_out) min_out[m0 + tid] = v;
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
  QPG_CHECK_ARG(K <= 8 * THREADS, "K <= 1024");
  QPG_CHECK_ARG((reinterpret_cast<uintptr_t>(codebook) & 15) == 0, "codebook must be 16-byte aligned");
  const size_t smem = (size_t)TM * D * 4 + TM * 4 + TM * (THREADS / 32) * 8 + 64;
  QPG_CHECK_ARG(smem <= 200 * 1024, "D too large");
  const int64_t blocks = (M + TM - 1) / TM;
  const int cpt = (K + THREADS - 1) / THREADS;
  cudaStream_t st = (cudaStream_t)stream;
  if (cpt <= 4) {
    QPG_CUDA(cudaFuncSetAttribute(vq_argmin_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_argmin_kernel<4><<<(unsigned)blocks, THREADS, smem, st>>>(x, codebook, M, D, K, idx_out, min_out);
  } else {
    QPG_CUDA(cudaFuncSetAttribute(vq_argmin_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_argmin_kernel<8><<<(unsigned)blocks, THREADS, smem, st>>>(x, codebook, M, D, K, idx_out, min_out);
  }
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
