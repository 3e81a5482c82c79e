// Error slot, launch counter and version of libb2s (the C ABI declared in include/b2s.h).
#include <atomic>
#include <string.h>
#include "b2s_internal.h"

namespace b2s {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return code;
}

int set_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
  return B2S_ERR_CUDA;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, what);
  return B2S_OK;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace b2s

extern "C" const char* b2s_last_error(void) { return b2s::g_err; }
extern "C" long long b2s_launch_count(void) { return b2s::g_launches.load(std::memory_order_relaxed); }
extern "C" int b2s_version(void) { return 100; }
