// Error reporting + launch accounting for libfgcolor.so
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace fgc {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FGC_ECUDA;
  }
  return FGC_OK;
}
}  // namespace fgc

extern "C" {
const char* fgc_last_error(void) { return fgc::g_err; }
int fgc_version(void) { return 100; }
long long fgc_launch_count(void) { return fgc::g_launches.load(); }
}
