// Shared host/device helpers for the mpformer_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpformer_b200.h"

namespace mpf {

// Thread-local error text returned by mpf_last_error().
void set_error(const char* fmt, ...);
void clear_error();
// Counts kernels launched by this library (bench.py's gpu_launches claim).
void count_launch(int n = 1);
// cudaGetLastError() -> return code (+ message) for the C ABI.
int finish_launch(const char* what);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mpf

#define MPF_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      mpf::set_error(__VA_ARGS__);    \
      return MPF_ERR_BAD_ARG;         \
    }                                 \
  } while (0)

#define MPF_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      mpf::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                \
      return static_cast<int>(_e);                                                   \
    }                                                                                \
  } while (0)
