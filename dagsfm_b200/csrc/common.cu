// Library-wide pieces of the C ABI: error text, version, launch counter.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "../../include/dagsfm_b200.h"
#include "common_host.h"

namespace b2 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
  return code;
}
int set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof g_err, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e),
           base ? base + 1 : file, line, what);
  cudaGetLastError();  // clear the sticky-less error state
  return B2_ERR_CUDA;
}
void count_launches(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace b2

extern "C" {
const char* b2_last_error(void) { return b2::g_err; }
const char* b2_version(void) { return "dagsfm_b200 0.1 (sm_100a)"; }
uint64_t b2_kernel_launch_count(void) { return b2::g_launches.load(std::memory_order_relaxed); }
}
