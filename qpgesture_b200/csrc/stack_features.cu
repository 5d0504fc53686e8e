// Device-side feature stacking: raw WavLM frames -> the rows the matcher scans.
//
// Replaces, for the rows that are actually read, the host pipeline of
// codebook/Speech2GestureMatching/data_processing.py:255-274:
//   F.interpolate(wavlm^T, size=T_out, mode='linear', align_corners=True)   [n, T_in, C] -> [n, T_out, C]
//   feat[n, t, i, :] = interp[n, t + 2 i, :]  (i < 6, zero past the end)
// and the row selection of GestureKNN.py:671-690 (database windows t = 6 m, m < 26) or :528,:565
// (query steps t = 24 s, s < 8).  out[(n * n_rows + r) * 6C + i*C + c] = interp[n, row_step*r + 2i, c].
//
// The interpolation reproduces ATen's CPU kernel bit for bit: scale = (T_in-1)/(T_out-1) in float32,
// src = scale * t, i0 = floor(src), w1 = src - i0, w0 = 1 - w1, value = fma(x[i0], w0, fl(x[i1] * w1))
// (established against torch 2.11 in the build container, checked again by tests/test_matcher_gpu.py).
#include "qpg_common.cuh"

namespace qpg {
namespace {

__global__ void stack_wavlm_rows_kernel(const float* __restrict__ wavlm, int64_t n, int T_in, int C, int T_out,
                                        int n_rows, int row_step, int n_taps, int tap_step,
                                        float* __restrict__ out) {
  const int64_t total = n * n_rows * (int64_t)n_taps * C;
  const float scale = T_out > 1 ? __fdiv_rn((float)(T_in - 1), (float)(T_out - 1)) : 0.f;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    int64_t rest = e / C;
    const int i = (int)(rest % n_taps);
    rest /= n_taps;
    const int r = (int)(rest % n_rows);
    const int64_t seq = rest / n_rows;
    const int t = row_step * r + tap_step * i;
    float v = 0.f;
    if (t < T_out) {
      const float src = __fmul_rn(scale, (float)t);
      int i0 = (int)floorf(src);
      if (i0 > T_in - 1) i0 = T_in - 1;
      const int i1 = i0 + 1 < T_in ? i0 + 1 : T_in - 1;
      float w1 = __fsub_rn(src, (float)i0);
      w1 = fminf(fmaxf(w1, 0.f), 1.f);
      const float w0 = __fsub_rn(1.f, w1);
      const float* base = wavlm + seq * (int64_t)T_in * C;
      v = __fmaf_rn(base[(int64_t)i0 * C + c], w0, __fmul_rn(base[(int64_t)i1 * C + c], w1));
    }
    out[e] = v;
  }
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_stack_wavlm_rows(const float* wavlm, int64_t n_seq, int T_in, int C, int T_out, int n_rows,
                                    int row_step, float* out, void* stream) {
  QPG_CHECK_ARG(n_seq >= 0 && T_in > 0 && C > 0 && T_out > 0 && n_rows > 0 && row_step > 0, "bad shape");
  if (n_seq == 0) return QPG_OK;
  QPG_CHECK_ARG(wavlm && out, "null pointer");
  const int64_t total = n_seq * n_rows * 6ll * C;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  stack_wavlm_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(wavlm, n_seq, T_in, C, T_out, n_rows,
                                                                             row_step, 6, 2, out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
