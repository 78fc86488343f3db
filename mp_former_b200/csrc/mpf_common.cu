// Error bookkeeping + ABI housekeeping entry points.
#include "mpf_common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace mpf {

static thread_local char g_err[512] = {0};
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void clear_error() { g_err[0] = 0; }

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

int finish_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return MPF_OK;
}

}  // namespace mpf

extern "C" {

int mpf_abi_version(void) { return MPF_ABI_VERSION; }

const char* mpf_last_error(void) { return mpf::g_err; }

uint64_t mpf_launch_count(void) { return mpf::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
