#include "status.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace ed2 {
namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return ED2_OK;
}
}  // namespace ed2
