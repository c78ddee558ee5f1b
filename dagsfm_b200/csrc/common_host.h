// Host-side helpers shared by the C-ABI translation units: thread-local error
// text, CUDA status mapping, the kernel-launch counter.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace b2 {

int set_error(int code, const char* msg);          // stores msg, returns code
int set_cuda_error(cudaError_t e, const char* what, const char* file, int line);
void count_launches(uint64_t n);

}  // namespace b2

#define B2_CUDA(expr)                                                         \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return ::b2::set_cuda_error(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define B2_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)
