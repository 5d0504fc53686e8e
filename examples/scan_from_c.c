/* Minimal plain-C user of the C ABI (include/qpg.h): pack a window table, scan it for two queries and read the
 * (best distance, best window) tables back.  Build:
 *   gcc -std=c99 -Wall -pedantic -Iinclude examples/scan_from_c.c -Lqpgesture_b200 -lqpg_sm100 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/qpgesture_b200 -o scan_from_c
 * (device memory is managed by the caller; here through the CUDA runtime's C API). */
#include <stdio.h>
#include <stdlib.h>

#include "qpg.h"

/* the few CUDA runtime entry points this example needs, declared by hand to stay free of CUDA headers */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
enum { H2D = 1, D2H = 2 };

#define CHECK(x)                                                        \
  do {                                                                  \
    int rc__ = (x);                                                     \
    if (rc__ != 0) {                                                    \
      fprintf(stderr, "%s -> %d (%s)\n", #x, rc__, qpg_last_error());   \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(void) {
  const int64_t W = 1000;
  const int D = 384, Q = 2;
  float *rows_h = malloc(sizeof(float) * W * D), *q_h = malloc(sizeof(float) * Q * D);
  int32_t* lab_h = malloc(sizeof(int32_t) * W);
  qpg_pair_t* tab_h = malloc(sizeof(qpg_pair_t) * Q * QPG_CODEBOOK_SIZE);
  void *rows_d, *packed_d, *sq_d, *lab_d, *q_d, *tab_d;
  int64_t i;
  unsigned s = 1u;
  for (i = 0; i < W * D; ++i) { s = s * 1664525u + 1013904223u; rows_h[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
  for (i = 0; i < Q * D; ++i) { s = s * 1664525u + 1013904223u; q_h[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
  for (i = 0; i < W; ++i) lab_h[i] = (int32_t)(i % QPG_CODEBOOK_SIZE);
  printf("libqpg version %d, packed bytes %zu\n", qpg_version(), qpg_packed_bytes(W, D));
  CHECK(cudaMalloc(&rows_d, sizeof(float) * W * D));
  CHECK(cudaMalloc(&packed_d, qpg_packed_bytes(W, D)));
  CHECK(cudaMalloc(&sq_d, sizeof(double) * W));
  CHECK(cudaMalloc(&lab_d, sizeof(int32_t) * W));
  CHECK(cudaMalloc(&q_d, sizeof(float) * Q * D));
  CHECK(cudaMalloc(&tab_d, sizeof(qpg_pair_t) * Q * QPG_CODEBOOK_SIZE));
  CHECK(cudaMemcpy(rows_d, rows_h, sizeof(float) * W * D, H2D));
  CHECK(cudaMemcpy(lab_d, lab_h, sizeof(int32_t) * W, H2D));
  CHECK(cudaMemcpy(q_d, q_h, sizeof(float) * Q * D, H2D));
  CHECK(qpg_pack_rows_f32(rows_d, W, D, packed_d, sq_d, NULL));
  CHECK(qpg_table_init(tab_d, (int64_t)Q * QPG_CODEBOOK_SIZE, NULL));
  CHECK(qpg_cand_cosine_minbycode(packed_d, sq_d, lab_d, W, D, 0, q_d, Q, tab_d, 0, NULL));
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(tab_h, tab_d, sizeof(qpg_pair_t) * Q * QPG_CODEBOOK_SIZE, D2H));
  printf("query 0, start-code 7: best distance %.6f at window %lld\n", tab_h[7].d, (long long)tab_h[7].id);
  cudaFree(rows_d); cudaFree(packed_d); cudaFree(sq_d); cudaFree(lab_d); cudaFree(q_d); cudaFree(tab_d);
  free(rows_h); free(q_h); free(lab_h); free(tab_h);
  return 0;
}
