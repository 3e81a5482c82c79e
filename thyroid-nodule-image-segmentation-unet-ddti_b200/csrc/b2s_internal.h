// Internal helpers shared by the translation units of libb2s: error slot, launch check, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <atomic>
#include "../../include/b2s.h"

namespace b2s {

int set_error(int code, const char* msg);             // stores msg in the thread-local slot, returns code
int set_cuda_error(cudaError_t e, const char* what);  // B2S_ERR_CUDA with cudaGetErrorString
int check_launch(const char* what);                   // cudaGetLastError() -> status
void count_launch();                                  // bumps the kernel-launch counter (b2s_launch_count)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: one process may drive several
// GPUs (nn.DataParallel runs one host thread per GPU, reference utils/trainer.py:28-30), so it is set once per
// (kernel, device). `mask` is the launch site's function-local static (one bit per device ordinal).
template <typename K>
static inline cudaError_t allow_dynamic_smem(K kernel, int bytes, std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
  return e;
}

}  // namespace b2s
