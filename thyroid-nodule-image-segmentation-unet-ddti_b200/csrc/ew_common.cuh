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

// ---- per-thread asynchronous prefetch ring ---------------------------------------------------------------------
// The streaming kernels are bound by the bytes they keep in flight: HBM3e at ~6.5 TB/s with ~1 us of loaded latency
// needs >= 44 KB outstanding per SM, and register loads at 64-160 registers per thread reach 16-32 KB (measured:
// read bandwidth == bytes in flight / 1 us for every kernel of this file). Each thread therefore keeps DEPTH work
// items of NV 16-byte global->shared async copies (LDGSTS: no destination registers) in flight and later reads back
// only the slots it filled itself. No block-level synchronisation is involved and the thread <-> element mapping of
// a kernel (hence its summation order and results) is the same as with direct loads.
// Layout [DEPTH][NV][kThreads] x 16 B: consecutive threads hit consecutive 16-byte slots (conflict-free).
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

template <int NV, int DEPTH>
struct PrefetchRing {
  static constexpr int kBytes = DEPTH * NV * kThreads * 16;
  uint32_t base;   // shared-space address of this thread's slot of (stage 0, vector 0)
  __device__ __forceinline__ explicit PrefetchRing(const void* smem) : base(smem_u32(smem) + threadIdx.x * 16u) {}
  __device__ __forceinline__ uint32_t addr(int stage, int v) const {
    return base + static_cast<uint32_t>(stage * NV + v) * (kThreads * 16u);
  }
  __device__ __forceinline__ void fetch(int stage, int v, const void* gmem) const { cp_async16(addr(stage, v), gmem); }
  __device__ __forceinline__ uint4 get(int stage, int v) const {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr(stage, v)) : "memory");
    return r;
  }
};

// Grid of a grid-stride streaming kernel: one full wave of resident blocks (occupancy with `smem` dynamic bytes, at
// most kEwBlocks / kSMs per SM) ...
template <typename K>
static inline int ew_wave_blocks(K kernel, int smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) != cudaSuccess || occ < 1) occ = 1;
  if (occ > kEwBlocks / kSMs) occ = kEwBlocks / kSMs;
  return kSMs * occ;
}
// ... or fewer when the tensor has fewer than `per_block` items per block of that wave.
static inline int ew_clamp_grid(int wave_blocks, long long work_items, int per_block) {
  const long long need = (work_items + per_block - 1) / per_block;
  if (need < 1) return 1;
  return need < wave_blocks ? static_cast<int>(need) : wave_blocks;
}
// opt a kernel into more than 48 KB of dynamic shared memory (once per kernel and process)
template <typename K>
static inline bool ew_allow_smem(K kernel, int smem) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
}

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
