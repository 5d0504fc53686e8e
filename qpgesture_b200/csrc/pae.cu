// Periodic auto-encoder inference (the producer of the match database's phase column), sm_100a.
//
// Reference: codebook/PAE.py:50-162 (Model.forward, eval mode) and :477-508 (pose2phase: one 240-frame window per
// frame of a sequence, batch 1, ~250 MFLOP each).
//
// pose2phase's windows are shifted views of ONE padded velocity sequence V [T + 238, C].  Writing the first
// convolution (kernel 240, zeros padding 120, first frame of every window forced to zero) for window i at output
// position t gives
//     y1_i[o, t] = sum_c sum_{k = klo(t)}^{khi(t)} w[o, c, k] * V[s + k, c],      s = i + t - 121,
//     klo(t) = max(0, 121 - t),  khi(t) = min(239, 359 - t),
// i.e. a k-RANGE of the untruncated correlation on "diagonal" s.  All (i, t) with the same s share the 240 terms
//     G[o, k] = sum_c w[o, c, k] * V[s + k, c],
// so one CTA per diagonal computes them once (float64), prefix-sums them over k, and emits every (i, t) on the
// diagonal as a difference of two prefix values: 0.49 MFLOP per diagonal instead of 117 MFLOP per window, about 240x
// less arithmetic than the per-window convolution, and more accurate than a float32 sum of 32 400 products.  The
// second convolution's input differs per window (tanh is applied per window), so it runs as a plain batched
// convolution; the FFT-derived parameters and the phase come from a third kernel (float64 DFT of 240 samples).
#include <math.h>

#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int PAE_K = 240;       // PAE.py:27 `frames`
constexpr int PAE_O = 15;        // PAE.py:67 intermediate channels (input_channels / 9)
constexpr int PAE_CMAX = 135;    // PAE.py:32
constexpr int V_STRIDE = PAE_CMAX + 2;     // odd stride: conflict-free column reads

__global__ void __launch_bounds__(256, 1)
pae_sliding_conv1_kernel(const float* __restrict__ vel, const float* __restrict__ w, const float* __restrict__ scale,
                         const float* __restrict__ shift, int T, int C, float* __restrict__ h1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ds = reinterpret_cast<double*>(smem_raw);                       // [PAE_O][PAE_K + 1] prefix sums
  float* Vs = reinterpret_cast<float*>(Ds + PAE_O * (PAE_K + 1));         // [PAE_K][V_STRIDE]
  const int s = (int)blockIdx.x - (PAE_K / 2 + 1);
  const int Tp = T + PAE_K - 2;
  for (int idx = threadIdx.x; idx < PAE_K * C; idx += blockDim.x) {
    const int r = idx / C, c = idx - r * C;
    const int g = s + r;
    Vs[r * V_STRIDE + c] = (g >= 0 && g < Tp) ? vel[(size_t)g * C + c] : 0.f;
  }
  __syncthreads();
  const int k = threadIdx.x;
  if (k < PAE_K) {
    double acc[PAE_O];
#pragma unroll
    for (int o = 0; o < PAE_O; ++o) acc[o] = 0.0;
    const float* wk = w + k;
    for (int c = 0; c < C; ++c) {
      const double v = (double)Vs[k * V_STRIDE + c];
#pragma unroll
      for (int o = 0; o < PAE_O; ++o) acc[o] = fma((double)__ldg(wk + (size_t)(o * C + c) * PAE_K), v, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < PAE_O; ++o) Ds[o * (PAE_K + 1) + k + 1] = acc[o];
  }
  __syncthreads();
  // inclusive scan over k, one warp per output channel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < PAE_O; o += 8) {
    double carry = 0.0;
    double* d = Ds + o * (PAE_K + 1);
    if (lane == 0) d[0] = 0.0;
    for (int j = 0; j < (PAE_K + 31) / 32; ++j) {
      const int idx = j * 32 + lane;
      double x = idx < PAE_K ? d[idx + 1] : 0.0;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
      }
      x += carry;
      if (idx < PAE_K) d[idx + 1] = x;
      carry = __shfl_sync(0xffffffffu, x, 31);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < PAE_O * (PAE_K + 1); idx += blockDim.x) {
    const int o = idx / (PAE_K + 1), t = idx - o * (PAE_K + 1);
    const int i = s + (PAE_K / 2 + 1) - t;
    if (i < 0 || i >= T) continue;
    const int klo = max(0, PAE_K / 2 + 1 - t), khi = min(PAE_K - 1, PAE_K + PAE_K / 2 - 1 - t);
    const double val = Ds[o * (PAE_K + 1) + khi + 1] - Ds[o * (PAE_K + 1) + klo];
    h1[((size_t)i * PAE_O + o) * (PAE_K + 1) + t] = (float)tanh((double)scale[o] * val + (double)shift[o]);
  }
}

// ---- plain batched convolution, float32 FFMA (second conv of pose2phase; all four convs of Model.forward) ------
// One CTA = 8 output channels x 256 positions x TWO windows: the staged weights (the larger operand) serve both
// windows, which halves the weight staging and the shared-memory reads per multiply-add.
constexpr int CONV_OG = 8;       // output channels per CTA
constexpr int CONV_CC = 15;      // input channels staged per chunk
constexpr int CONV_TT = 256;     // output positions per CTA
constexpr int CONV_NW = 2;       // windows per CTA

__global__ void __launch_bounds__(256, 1)
pae_conv1d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ shift, int B, int Ci, int Lin, int Co, int K, int pad, int Lo, int act,
                  float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* ws = reinterpret_cast<float*>(smem_raw);                 // [CONV_OG][CONV_CC][K]
  float* xs = ws + CONV_OG * CONV_CC * K;                         // [CONV_NW][CONV_CC][CONV_TT + K - 1]
  const int xw = CONV_TT + K - 1;
  const int t0 = blockIdx.x * CONV_TT, o0 = blockIdx.y * CONV_OG, b0 = blockIdx.z * CONV_NW;
  const int t = threadIdx.x;
  float acc[CONV_NW][CONV_OG];
#pragma unroll
  for (int n = 0; n < CONV_NW; ++n)
#pragma unroll
    for (int o = 0; o < CONV_OG; ++o) acc[n][o] = 0.f;
  for (int c0 = 0; c0 < Ci; c0 += CONV_CC) {
    const int cc = min(CONV_CC, Ci - c0);
    for (int idx = threadIdx.x; idx < CONV_OG * cc * K; idx += blockDim.x) {
      const int o = idx / (cc * K), r = idx - o * (cc * K);
      const int c = r / K, kk = r - c * K;
      ws[(o * CONV_CC + c) * K + kk] = (o0 + o < Co) ? w[((size_t)(o0 + o) * Ci + c0 + c) * K + kk] : 0.f;
    }
    for (int idx = threadIdx.x; idx < CONV_NW * cc * xw; idx += blockDim.x) {
      const int n = idx / (cc * xw), r = idx - n * (cc * xw);
      const int c = r / xw, j = r - c * xw;
      const int pos = t0 + j - pad;
      xs[(n * CONV_CC + c) * xw + j] =
          (b0 + n < B && pos >= 0 && pos < Lin) ? x[((size_t)(b0 + n) * Ci + c0 + c) * Lin + pos] : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < cc; ++c) {
      const float* xc0 = xs + c * xw + t;
      const float* xc1 = xs + (CONV_CC + c) * xw + t;
      for (int kk = 0; kk < K; kk += 4) {
        const float a0 = xc0[kk], a1 = xc0[kk + 1], a2 = xc0[kk + 2], a3 = xc0[kk + 3];
        const float b0v = xc1[kk], b1v = xc1[kk + 1], b2v = xc1[kk + 2], b3v = xc1[kk + 3];
#pragma unroll
        for (int o = 0; o < CONV_OG; ++o) {
          const float4 wv = *reinterpret_cast<const float4*>(ws + (o * CONV_CC + c) * K + kk);
          acc[0][o] = fmaf(wv.x, a0, acc[0][o]);
          acc[0][o] = fmaf(wv.y, a1, acc[0][o]);
          acc[0][o] = fmaf(wv.z, a2, acc[0][o]);
          acc[0][o] = fmaf(wv.w, a3, acc[0][o]);
          acc[1][o] = fmaf(wv.x, b0v, acc[1][o]);
          acc[1][o] = fmaf(wv.y, b1v, acc[1][o]);
          acc[1][o] = fmaf(wv.z, b2v, acc[1][o]);
          acc[1][o] = fmaf(wv.w, b3v, acc[1][o]);
        }
      }
    }
    __syncthreads();
  }
  if (t0 + t < Lo) {
#pragma unroll
    for (int n = 0; n < CONV_NW; ++n)
      if (b0 + n < B) {
#pragma unroll
        for (int o = 0; o < CONV_OG; ++o)
          if (o0 + o < Co) {
            float v = fmaf(scale[o0 + o], acc[n][o], shift[o0 + o]);
            if (act) v = tanhf(v);
            out[((size_t)(b0 + n) * Co + o0 + o) * Lo + t0 + t] = v;
          }
      }
  }
}

// ---- (phase, frequency, amplitude, offset) of every latent channel: PAE.py:99-114, 132-136 --------------------
constexpr int PAR_TMAX = 256;
constexpr int PAR_EMAX = 8;

__global__ void __launch_bounds__(256)
pae_params_kernel(const float* __restrict__ latent, const float* __restrict__ fcw, const float* __restrict__ fc_scale,
                  const float* __restrict__ fc_shift, const float* __restrict__ freqs, float time_scale, int E, int T,
                  float* __restrict__ params) {
  __shared__ double cs[PAR_TMAX], sn[PAR_TMAX];
  __shared__ double ys[PAR_EMAX][PAR_TMAX];
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < T; j += blockDim.x) sincospi(2.0 * j / T, &sn[j], &cs[j]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e0 = 0; e0 < E; e0 += PAR_EMAX) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < min(PAR_EMAX, E - e0) * T; idx += blockDim.x)
      ys[idx / T][idx % T] = (double)latent[((size_t)b * E + e0) * T + idx];
    __syncthreads();
    const int e = e0 + warp;
    if (e >= E) continue;
    double* y = ys[warp];
    // offset and the two Linear(T, 2) outputs first; then the row is centred in place, so that the DFT below sees
    // y - mean (same bins k >= 1 mathematically, no large DC part in the sums, and an exactly constant latent gives
    // an exactly zero spectrum -> 0 / 0 = NaN frequency, the value of PAE.py:106 for zero power)
    double s_y = 0.0, v0 = 0.0, v1 = 0.0;
    for (int t = lane; t < T; t += 32) {
      s_y += y[t];
      v0 = fma((double)fcw[((size_t)e * 2 + 0) * T + t], y[t], v0);
      v1 = fma((double)fcw[((size_t)e * 2 + 1) * T + t], y[t], v1);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      s_y += __shfl_xor_sync(0xffffffffu, s_y, off);
      v0 += __shfl_xor_sync(0xffffffffu, v0, off);
      v1 += __shfl_xor_sync(0xffffffffu, v1, off);
    }
    const double mean = s_y / T;
    for (int t = lane; t < T; t += 32) y[t] -= mean;
    __syncwarp();
    // power spectrum without the DC bin, bins spread over the lanes
    double s_pow = 0.0, s_fpow = 0.0;
    for (int kb = 1 + lane; kb <= T / 2; kb += 32) {
      double re = 0.0, im = 0.0;
      int idx = 0;
      for (int t = 0; t < T; ++t) {
        re = fma(y[t], cs[idx], re);
        im = fma(y[t], sn[idx], im);
        idx += kb;
        if (idx >= T) idx -= T;
      }
      const double pw = re * re + im * im;
      s_pow += pw;
      s_fpow = fma((double)freqs[kb - 1], pw, s_fpow);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      s_pow += __shfl_xor_sync(0xffffffffu, s_pow, off);
      s_fpow += __shfl_xor_sync(0xffffffffu, s_fpow, off);
    }
    if (lane == 0) {
      const double vx = (double)fc_scale[e * 2 + 0] * v0 + (double)fc_shift[e * 2 + 0];
      const double vy = (double)fc_scale[e * 2 + 1] * v1 + (double)fc_shift[e * 2 + 1];
      // the model's own atan2 (PAE.py:92-97): atan(y / x), moved by half a turn when x < 0; nothing special at x == 0
      const double kPi = 3.14159265358979323846;
      double ph = atan(vy / vx);
      if (vx < 0.0 && vy >= 0.0) ph += kPi;
      if (vx < 0.0 && vy < 0.0) ph -= kPi;
      float* out = params + (size_t)b * 4 * E;
      out[0 * E + e] = (float)(ph / (2.0 * kPi));
      out[1 * E + e] = (float)(s_fpow / s_pow / (double)time_scale);       // 0 / 0 -> NaN, as in the reference
      out[2 * E + e] = (float)(2.0 * sqrt(s_pow) / T);
      out[3 * E + e] = (float)mean;
    }
  }
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_pae_sliding_conv1(const float* vel_pad, const float* w, const float* scale, const float* shift,
                                     int T, int C, int O, int K, float* h1, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QPG_CHECK_ARG(vel_pad && w && scale && shift && h1, "null pointer");
  QPG_CHECK_ARG(K == PAE_K && O == PAE_O, "built for kernel 240 and 15 intermediate channels (PAE.py:27,67)");
  QPG_CHECK_ARG(C >= 1 && C <= PAE_CMAX, "channels must be 1..135");
  if (T <= 0) return QPG_OK;
  const size_t smem = sizeof(double) * PAE_O * (PAE_K + 1) + sizeof(float) * PAE_K * V_STRIDE;
  QPG_CUDA(cudaFuncSetAttribute(pae_sliding_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pae_sliding_conv1_kernel<<<T + PAE_K, 256, smem, stream>>>(vel_pad, w, scale, shift, T, C, h1);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_pae_conv1d(const float* x, const float* w, const float* scale, const float* shift, int B, int Ci,
                              int Lin, int Co, int K, int pad, int act, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QPG_CHECK_ARG(x && w && scale && shift && out, "null pointer");
  QPG_CHECK_ARG(Ci >= 1 && Co >= 1 && Lin >= 1 && pad >= 0, "bad shape");
  QPG_CHECK_ARG(K >= 4 && K % 4 == 0 && K <= 256, "kernel width must be a multiple of 4, at most 256");
  const int Lo = Lin + 2 * pad - K + 1;
  QPG_CHECK_ARG(Lo >= 1, "kernel wider than the padded input");
  QPG_CHECK_ARG(B <= 65535 * CONV_NW, "at most 131070 windows per call");
  if (B <= 0) return QPG_OK;
  const size_t smem = sizeof(float) * ((size_t)CONV_OG * CONV_CC * K + (size_t)CONV_NW * CONV_CC * (CONV_TT + K - 1));
  QPG_CUDA(cudaFuncSetAttribute(pae_conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((Lo + CONV_TT - 1) / CONV_TT, (Co + CONV_OG - 1) / CONV_OG, (B + CONV_NW - 1) / CONV_NW);
  pae_conv1d_kernel<<<grid, 256, smem, stream>>>(x, w, scale, shift, B, Ci, Lin, Co, K, pad, Lo, act, out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_pae_params(const float* latent, const float* fcw, const float* fc_scale, const float* fc_shift,
                              const float* freqs, float time_scale, int B, int E, int T, float* params,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QPG_CHECK_ARG(latent && fcw && fc_scale && fc_shift && freqs && params, "null pointer");
  QPG_CHECK_ARG(T >= 2 && T <= PAR_TMAX && T % 2 == 0, "time range must be even, at most 256");
  QPG_CHECK_ARG(E >= 1 && time_scale > 0.f, "bad shape");
  if (B <= 0) return QPG_OK;
  pae_params_kernel<<<B, 256, 0, stream>>>(latent, fcw, fc_scale, fc_shift, freqs, time_scale, E, T, params);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
