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

// True exactly once per (call site, device): function attributes such as the opt-in dynamic shared-memory size are
// per device, so a process that drives several GPUs must set them on each (`seen`: one static mask per call site).
inline bool first_use_on_this_device(unsigned long long& seen) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (seen & bit) return false;
  seen |= bit;          // benign race: the attribute is set again at worst
  return true;
}

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
