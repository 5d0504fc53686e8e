// 1-D convolution stacks of the VQ-VAE as one generic "tap GEMM" over
// channels-last activations (see include/qpg.h, qpg_conv1d_taps_f32):
//
//   out[b, t*os + oo, co] = bias[co] + residual[b, t*os + oo, co]
//       + sum_{tap, ci} act(in[b, t*is + off[tap], ci]) * w[tap][ci][co]
//
// This file holds the float32 FFMA path (precision 0, parity mode): a classic
// 64x64x16 shared-memory tiled SGEMM with a 4x4 register tile whose A operand
// is gathered on the fly (implicit im2col: frame shift per tap, zero padding,
// optional ReLU on load), bias / residual fused in the epilogue.
// Replaces nn.Conv1d / nn.ConvTranspose1d at encdec.py:20,24,39,45,113 and
// resnet.py:33-36.
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct ConvParams {
  int B, T_in, T_out_total, C_in, C_out, n_taps;
  int off[4];
  int in_stride, out_stride, out_offset, n_out, relu_in;
};

__global__ void __launch_bounds__(256)
    conv_taps_f32_kernel(ConvParams p, const float* __restrict__ in, const float* __restrict__ w,
                         const float* __restrict__ bias, const float* __restrict__ residual,
                         float* __restrict__ out) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a 4x4 output tile
  const int64_t M = (int64_t)p.B * p.n_out;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load mapping: thread loads 4 consecutive channels of one row
  const int a_row = tid >> 2;         // 0..63
  const int a_col = (tid & 3) * 4;    // 0,4,8,12
  const int64_t am = m0 + a_row;
  const bool a_live = am < M;
  const int ab = a_live ? (int)(am / p.n_out) : 0;
  const int at = a_live ? (int)(am - (int64_t)ab * p.n_out) : 0;
  // B-load mapping: thread loads 4 consecutive output channels of one k row
  const int b_row = tid >> 4;         // 0..15
  const int b_col = (tid & 15) * 4;   // 0..60

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < p.n_taps; ++tap) {
    const int f = at * p.in_stride + p.off[tap];
    const bool f_ok = a_live && f >= 0 && f < p.T_in;
    const float* a_src = in + ((int64_t)ab * p.T_in + (f_ok ? f : 0)) * p.C_in;
    const float* w_tap = w + (int64_t)tap * p.C_in * p.C_out;
    for (int c0 = 0; c0 < p.C_in; c0 += BK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = c0 + a_col + j;
        float v = (f_ok && ci < p.C_in) ? a_src[ci] : 0.f;
        if (p.relu_in) v = fmaxf(v, 0.f);
        As[a_col + j][a_row] = v;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = c0 + b_row, co = n0 + b_col + j;
        Bs[b_row][b_col + j] = (ci < p.C_in && co < p.C_out) ? w_tap[(int64_t)ci * p.C_out + co] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = (int)(m / p.n_out);
    const int t = (int)(m - (int64_t)b * p.n_out);
    const int64_t orow = ((int64_t)b * p.T_out_total + (int64_t)t * p.out_stride + p.out_offset) * p.C_out;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.C_out) continue;
      float v = acc[i][j];
      if (bias) v += bias[co];
      if (residual) v += residual[orow + co];
      out[orow + co] = v;
    }
  }
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_conv1d_taps_f32(const qpg_conv_desc_t* d, const float* in, const float* w, const float* bias,
                                   const float* residual, float* out, void* stream) {
  QPG_CHECK_ARG(d != nullptr, "null descriptor");
  QPG_CHECK_ARG(d->B >= 0 && d->T_in > 0 && d->T_out_total > 0 && d->C_in > 0 && d->C_out > 0, "bad shape");
  QPG_CHECK_ARG(d->n_taps >= 1 && d->n_taps <= 4, "n_taps in 1..4");
  QPG_CHECK_ARG(d->n_out >= 0 && d->out_stride >= 1 && d->in_stride >= 1 && d->out_offset >= 0, "bad strides");
  QPG_CHECK_ARG(d->n_out == 0 || (int64_t)(d->n_out - 1) * d->out_stride + d->out_offset < d->T_out_total,
                "output index range exceeds T_out_total");
  if (d->B == 0 || d->n_out == 0) return QPG_OK;
  QPG_CHECK_ARG(in && w && out, "null pointer");
  if (d->precision != 0) {
    set_error("precision %d not built yet (0 = float32 FFMA)", d->precision);
    return QPG_E_UNSUPPORTED;
  }
  ConvParams p;
  p.B = d->B; p.T_in = d->T_in; p.T_out_total = d->T_out_total; p.C_in = d->C_in; p.C_out = d->C_out;
  p.n_taps = d->n_taps;
  for (int i = 0; i < 4; ++i) p.off[i] = d->tap_offset[i];
  p.in_stride = d->in_stride; p.out_stride = d->out_stride; p.out_offset = d->out_offset; p.n_out = d->n_out;
  p.relu_in = d->relu_in;
  const int64_t M = (int64_t)d->B * d->n_out;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d->C_out + BN - 1) / BN));
  conv_taps_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, in, w, bias, residual, out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
