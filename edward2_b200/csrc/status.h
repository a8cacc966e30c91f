// Status codes, thread-local error text and the launch counter shared by all translation units.
#pragma once
#include <cuda_runtime.h>

#define ED2_OK 0
#define ED2_ERR_BAD_SHAPE (-1)
#define ED2_ERR_BAD_DTYPE (-2)
#define ED2_ERR_BAD_ARG (-3)
#define ED2_ERR_UNSUPPORTED_ARCH (-4)
#define ED2_ERR_CUDA (-5)
#define ED2_ERR_WORKSPACE (-6)

namespace ed2 {
// printf-style; stores the message for ed2_last_error() and returns `code`.
int set_error(int code, const char* fmt, ...);
const char* last_error();
// every kernel launch of this library bumps a process-wide counter (bench.py reports it as gpu_launches)
void count_launch(int n = 1);
long long launch_count();
// check the (sticky-free) launch status of the kernel just enqueued
int check_launch(const char* what);
}  // namespace ed2
