// Post-decode pose processing on the device: the numeric half of make_bvh_GENEA2020_BT
// (process/process_bvh.py:57-77 of the reference): Savitzky-Golay smoothing (window 15, order 2, scipy's "interp"
// edges) of every pose channel over time, then rotation matrix -> intrinsic ZXY Euler angles in degrees per joint
// and frame with scipy's semantics (Rotation.from_matrix(...).as_euler('ZXY', degrees=True)): a matrix that is not
// orthogonal to 1e-12 is replaced by its orthogonal polar factor U V^T first, a non-positive determinant is an
// error.  The pymo pipeline / BVH writer that follow need the unshipped data_pipe_60_rotation.sav and stay on the host.
#include <math.h>

#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int SG_W = 15, SG_H = 7;

// out[t][c]: interior = correlation with the 15 symmetric coefficients; the first / last 7 frames = the quadratic
// fitted to the first / last 15 frames, evaluated at the frame (edge[7][15] rows, built on the host)
__global__ void __launch_bounds__(256)
savgol15_kernel(const float* __restrict__ x, int T, int C, const double* __restrict__ coef,
                const double* __restrict__ edge_first, const double* __restrict__ edge_last, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)T * C) return;
  const int t = (int)(i / C), c = (int)(i - (long long)t * C);
  double acc = 0.0;
  if (t < SG_H) {
    for (int k = 0; k < SG_W; ++k) acc = fma(edge_first[t * SG_W + k], (double)x[(size_t)k * C + c], acc);
  } else if (t >= T - SG_H) {
    const int r = t - (T - SG_H);
    for (int k = 0; k < SG_W; ++k) acc = fma(edge_last[r * SG_W + k], (double)x[(size_t)(T - SG_W + k) * C + c], acc);
  } else {
    for (int k = 0; k < SG_W; ++k) acc = fma(coef[k], (double)x[(size_t)(t - SG_H + k) * C + c], acc);
  }
  out[i] = acc;
}

__device__ __forceinline__ double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// flags: bit 0 = non-positive determinant (scipy raises ValueError), bit 1 = gimbal lock (third angle set to 0)
__global__ void __launch_bounds__(256)
rotmat_euler_zxy_kernel(const double* __restrict__ mats, long long N, double* __restrict__ euler,
                        int32_t* __restrict__ flags) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  double m[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = mats[i * 9 + k];
  int flag = 0;
  if (!(det3(m) > 0.0)) flag |= 1;
  // orthogonal to 1e-12 (atol, as scipy's isclose against the identity with rtol 1e-5 on ones: diag within ~1e-5)?
  // scipy: isclose(M M^T, I, atol=1e-12) with the default rtol = 1e-5 -> |g - e| <= 1e-12 + 1e-5 |e|
  bool orth = true;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double g = m[r * 3] * m[c * 3] + m[r * 3 + 1] * m[c * 3 + 1] + m[r * 3 + 2] * m[c * 3 + 2];
      const double e = r == c ? 1.0 : 0.0;
      if (!(fabs(g - e) <= 1e-12 + 1e-5 * e)) orth = false;
    }
  if (!orth && !(flag & 1)) {
    // orthogonal polar factor U V^T by scaled Newton iteration X <- (s X + X^-T / s) / 2 (quadratic convergence)
    for (int it = 0; it < 40; ++it) {
      const double d = det3(m);
      double inv_t[9];                               // X^-T = cofactor matrix / det
      inv_t[0] = (m[4] * m[8] - m[5] * m[7]) / d;
      inv_t[1] = (m[5] * m[6] - m[3] * m[8]) / d;
      inv_t[2] = (m[3] * m[7] - m[4] * m[6]) / d;
      inv_t[3] = (m[2] * m[7] - m[1] * m[8]) / d;
      inv_t[4] = (m[0] * m[8] - m[2] * m[6]) / d;
      inv_t[5] = (m[1] * m[6] - m[0] * m[7]) / d;
      inv_t[6] = (m[1] * m[5] - m[2] * m[4]) / d;
      inv_t[7] = (m[2] * m[3] - m[0] * m[5]) / d;
      inv_t[8] = (m[0] * m[4] - m[1] * m[3]) / d;
      double nx = 0.0, ni = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        nx = fma(m[k], m[k], nx);
        ni = fma(inv_t[k], inv_t[k], ni);
      }
      const double s = sqrt(sqrt(ni / nx));          // Frobenius scaling
      double diff = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const double nk = 0.5 * (s * m[k] + inv_t[k] / s);
        diff = fmax(diff, fabs(nk - m[k]));
        m[k] = nk;
      }
      if (diff < 1e-15) break;
    }
  }
  // R = Rz(a) Rx(b) Ry(c):  R21 = sin b,  R01 = -sin a cos b,  R11 = cos a cos b,  R20 = -cos b sin c,  R22 = cos b cos c
  const double kDeg = 57.29577951308232;
  double a, b, c;
  const double sb = fmin(1.0, fmax(-1.0, m[7]));
  b = asin(sb);
  if (hypot(m[1], m[4]) < 1e-7) {                    // cos b ~ 0: gimbal lock, scipy sets the third angle to zero
    flag |= 2;
    c = 0.0;
    a = atan2(m[3], m[0]);
  } else {
    a = atan2(-m[1], m[4]);
    c = atan2(-m[6], m[8]);
  }
  euler[i * 3 + 0] = a * kDeg;
  euler[i * 3 + 1] = b * kDeg;
  euler[i * 3 + 2] = c * kDeg;
  if (flags) flags[i] = flag;
}

}  // namespace
}  // namespace qpg

using namespace qpg;

extern "C" int qpg_savgol15_f64(const float* x, int T, int C, const double* coef, const double* edge_first,
                                const double* edge_last, double* out, void* stream) {
  QPG_CHECK_ARG(x && coef && edge_first && edge_last && out, "null pointer");
  QPG_CHECK_ARG(T >= 15 && C >= 1, "the window (15) must not be longer than the sequence");
  const long long n = (long long)T * C;
  savgol15_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, T, C, coef, edge_first, edge_last, out);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}

extern "C" int qpg_rotmat_to_euler_zxy(const double* mats, int64_t N, double* euler_deg, int32_t* flags, void* stream) {
  QPG_CHECK_ARG(N >= 0, "negative size");
  if (N == 0) return QPG_OK;
  QPG_CHECK_ARG(mats && euler_deg, "null pointer");
  rotmat_euler_zxy_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mats, N, euler_deg, flags);
  QPG_LAUNCH_CHECK();
  return QPG_OK;
}
