// Helpers shared by the bandwidth-kernel translation units (elementwise.cu, vnet_ops.cu).
#pragma once
#include "ptx.cuh"
#include "b2s_internal.h"

namespace b2s {

constexpr int kThreads = 256;
constexpr int kSMs = 148;
constexpr int kEwBlocks = kSMs * 4;  // rows of every element-wise partial buffer

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// elementwise.cu
int reduce_to_small(const float* in, int rows, int K, float* scratch, const float** out_ptr, int* out_rows,
                    cudaStream_t stream);

static inline int ew_grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  if (blocks > kEwBlocks) blocks = kEwBlocks;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}
static inline bool ew_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
// channel counts the vectorised element-wise kernels accept: C/8 must be a power of two dividing 256
static inline bool ew_channels_supported(int C) { return C >= 8 && C % 8 == 0 && ew_pow2(C / 8) && C / 8 <= kThreads; }

}  // namespace b2s
