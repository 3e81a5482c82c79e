// Internal helpers shared by the translation units of libb2s: error slot, launch check, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../include/b2s.h"

namespace b2s {

int set_error(int code, const char* msg);             // stores msg in the thread-local slot, returns code
int set_cuda_error(cudaError_t e, const char* what);  // B2S_ERR_CUDA with cudaGetErrorString
int check_launch(const char* what);                   // cudaGetLastError() -> status
void count_launch();                                  // bumps the kernel-launch counter (b2s_launch_count)
bool pdl_enabled();                                   // false when the environment has B2S_PDL=0 (A/B measurements)

// kernel<<<grid, block, smem, stream>>>(args...) with programmatic dependent launch allowed: the kernel may be
// scheduled while its predecessor in the stream is still running. Every kernel launched through here calls pdl_wait()
// (ptx.cuh) before it touches global memory.
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through check_launch()
}

}  // namespace b2s
