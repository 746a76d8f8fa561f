// Shared host-side plumbing for the C-ABI: per-thread error string, launch counter, launch checking.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace mm {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int32_t check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 2;
  }
  count_launch();
  return 0;
}

#define MM_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      mm::set_error(__VA_ARGS__);  \
      return 1;                    \
    }                              \
  } while (0)

using bf16 = __nv_bfloat16;

}  // namespace mm
