// Shared device/host helpers for libqpg_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qpg.h"

namespace qpg {

// ---- host-side error plumbing ------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define QPG_CHECK_ARG(cond, msg)                   \
  do {                                             \
    if (!(cond)) {                                 \
      qpg::set_error("bad argument: %s", msg);     \
      return QPG_E_BADARG;                         \
    }                                              \
  } while (0)

#define QPG_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      qpg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                         \
      return QPG_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

#define QPG_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      qpg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),       \
                     __FILE__, __LINE__);                                               \
      return QPG_E_CUDA;                                                                \
    }                                                                                   \
    qpg::count_launch();                                                                \
  } while (0)

int sm_count();

constexpr double kEmptyDist = 1e3;  // GestureKNN.py:668,709
// sklearn's normalize() (called by paired_cosine_distances) leaves a row unscaled when its norm is below
// 10 * eps(float64): such a row counts as all-zero.  Squared threshold.
constexpr double kTinySq = 4.930380657631324e-30;

// ---- (distance, id) pairs ordered lexicographically ---------------------------
// Distances are non-negative doubles, so their bit patterns order like
// unsigned integers; id -1 (empty) is the largest unsigned id.
struct __align__(16) Pair {
  unsigned long long d;   // bit pattern of a non-negative double
  unsigned long long id;  // global window id
};

__device__ __forceinline__ bool pair_less(const Pair& a, const Pair& b) {
  return a.d < b.d || (a.d == b.d && a.id < b.id);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 128-bit compare-and-swap (SASS: ATOMS.CAS.128 / ATOMG.E.CAS.128)
__device__ __forceinline__ Pair cas128_shared(Pair* addr, Pair cmp, Pair val) {
  Pair old;
  asm volatile(
      "{\n\t.reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.shared.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old.d), "=l"(old.id)
      : "l"(cmp.d), "l"(cmp.id), "l"(val.d), "l"(val.id), "r"(smem_u32(addr))
      : "memory");
  return old;
}

__device__ __forceinline__ Pair cas128_global(Pair* addr, Pair cmp, Pair val) {
  Pair old;
  asm volatile(
      "{\n\t.reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.global.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old.d), "=l"(old.id)
      : "l"(cmp.d), "l"(cmp.id), "l"(val.d), "l"(val.id), "l"(addr)
      : "memory");
  return old;
}

__device__ __forceinline__ Pair ld128_shared(const Pair* addr) {
  Pair v;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.d), "=l"(v.id) : "r"(smem_u32(addr)) : "memory");
  return v;
}

__device__ __forceinline__ Pair ld128_global_volatile(const Pair* addr) {
  Pair v;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.d), "=l"(v.id) : "l"(addr) : "memory");
  return v;
}

// lexicographic atomic min of a Pair slot
__device__ __forceinline__ void pair_min_shared(Pair* slot, Pair mine) {
  Pair cur = ld128_shared(slot);
  while (pair_less(mine, cur)) {
    Pair old = cas128_shared(slot, cur, mine);
    if (old.d == cur.d && old.id == cur.id) break;
    cur = old;
  }
}

__device__ __forceinline__ void pair_min_global(Pair* slot, Pair mine) {
  Pair cur = ld128_global_volatile(slot);
  while (pair_less(mine, cur)) {
    Pair old = cas128_global(slot, cur, mine);
    if (old.d == cur.d && old.id == cur.id) break;
    cur = old;
  }
}

// ---- mbarrier + TMA bulk copy (SASS: SYNCS.*, UBLKCP) ---------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy completing on an mbarrier (bytes multiple of 16, 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace qpg
