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

}  // namespace b2s
