// Helpers shared by the cosine candidate-distance kernels (cand_cosine.cu, cand_cosine2.cu).
#pragma once
#include "qpg_common.cuh"

namespace qpg {
namespace {

constexpr int R = QPG_ROWS_PER_GROUP;  // 8 rows per tile
constexpr int DC = QPG_CHUNK;          // 128 floats per tile row
constexpr int TILE_FLOATS = R * DC;    // 1024
constexpr int TILE_BYTES = TILE_FLOATS * 4;
constexpr int KB = QPG_CODEBOOK_SIZE;
constexpr size_t kSmemLimit = 232448;  // 227 KiB opt-in limit per CTA on sm_100

// ------------------------------------------------- transposed warp reduce ----
// V partial sums per lane -> after log2 steps lane l owns total #(l*V/32 ...):
// V=64: totals 2l,2l+1 ; V=32: total l ; V=16: total l>>1 ; V=8: total l>>2.
template <int V, int O>
struct TransposeReduce {
  static __device__ __forceinline__ void run(double* v, int lane) {
    if constexpr (O >= 1) {
      if constexpr (V > 1) {
        constexpr int half = V / 2;
        const bool upper = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const double send = upper ? v[i] : v[i + half];
          const double keep = upper ? v[i + half] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
        }
        TransposeReduce<half, O / 2>::run(v, lane);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], O);
        TransposeReduce<1, O / 2>::run(v, lane);
      }
    }
  }
};

// Exact float32 -> float64 conversion on the integer ALU (normal numbers: re-bias the exponent, move the
// mantissa); zero / denormal / inf / nan take the F2F path.  Used for the query operands so that the XU pipe
// (F2F.F64.F32 runs there at 1/8 rate) only converts the database elements.
__device__ __forceinline__ double f32_to_f64_alu(float f) {
  const uint32_t u = __float_as_uint(f);
  const uint32_t e = u & 0x7f800000u;
  if (e == 0u || e == 0x7f800000u) return (double)f;
  const uint32_t hi = (u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u);
  return __hiloint2double((int)hi, (int)(u << 29));
}

__device__ __forceinline__ double cosine_distance(double dot, double sqq, double sqx) {
  // sklearn: normalize() leaves an all-zero vector at zero (norm 0 -> 1), then
  // 0.5*||a-b||^2 = 0.5*(|a|^2+|b|^2) - <a,b> with |a|,|b| in {0,1}.
  const double a = sqq > kTinySq ? 1.0 : 0.0;
  const double b = sqx > kTinySq ? 1.0 : 0.0;
  double c = 0.0;
  if (sqq > kTinySq && sqx > kTinySq) c = dot / (sqrt(sqq) * sqrt(sqx));
  double d = 0.5 * (a + b) - c;
  return d < 0.0 ? 0.0 : d;
}


}  // namespace
}  // namespace qpg
