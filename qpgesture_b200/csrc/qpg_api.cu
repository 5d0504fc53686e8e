// Library-level plumbing: version, thread-local error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "qpg_common.cuh"

namespace qpg {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int cosine_set_tuning(int ncw, int ns, int grid, int team);
int cosine_set_alternate(int on);

}  // namespace qpg

extern "C" int qpg_version(void) { return 100; }  // 0.1.0

extern "C" const char* qpg_last_error(void) { return qpg::g_err; }

extern "C" uint64_t qpg_launch_count(void) { return qpg::g_launches.load(std::memory_order_relaxed); }

// Tuning hook used by bench.py / profiles sweeps (not part of the reference surface).
extern "C" int qpg_tune_cosine(int compute_warps, int stages, int grid, int team) {
  return qpg::cosine_set_tuning(compute_warps, stages, grid, team);
}

extern "C" int qpg_tune_cosine_alternate(int enable) { return qpg::cosine_set_alternate(enable); }
