// Bandwidth kernels of the V-Net variant (reference models/vnet.py): BatchNorm -> ReLU -> Dropout (+ residual)
// apply and its two-pass backward, squeeze-and-excitation (global pool, the two tiny FC layers, channel scale) forward
// and backward, per-channel sums (bias gradients), zero-insertion 2x upsampling (backward of the stride-2 conv).
// NHWC bf16 activations, 16-byte vectors (8 channels per thread), deterministic two-stage reductions.
#include "ew_common.cuh"

namespace b2s {

// ---- counter-based dropout mask: depends only on (seed, logical NHWC element index) so backward re-derives it ----
// Eight 16-bit uniforms per group of eight consecutive channels: keep iff u16 >= p * 65536. The first 32 bits are a
// full-avalanche hash of (seed, group index); the other three words are cheap bijective multiply-xorshift steps of
// it (the kernels that draw the mask are instruction-issue bound, so the mask has to cost few instructions).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
struct DropMask {
  uint32_t base0;     // mix32(seed): the per-launch key of indices below 2^35 (the usual case)
  uint32_t seed, thresh16;
  float keep_scale;   // 1 / (1 - p), p quantised to 1/65536
  __device__ __forceinline__ DropMask(uint32_t seed_, uint32_t thresh16_, float keep_scale_)
      : base0(mix32(seed_)), seed(seed_), thresh16(thresh16_), keep_scale(keep_scale_) {}
  // f[k] = keep_scale if element idx8 + k is kept, else 0 (idx8: a multiple of 8)
  __device__ __forceinline__ void factors(unsigned long long idx8, float (&f)[8]) const {
    const uint32_t lo = static_cast<uint32_t>(idx8 >> 3), hi = static_cast<uint32_t>(idx8 >> 35);
    const uint32_t base = hi ? mix32(hi + seed) : base0;
    uint32_t h[4];
    h[0] = mix32(lo ^ base);
    h[1] = h[0] * 0x9e3779b1U; h[1] ^= h[1] >> 15;
    h[2] = h[1] * 0x85ebca77U; h[2] ^= h[2] >> 13;
    h[3] = h[2] * 0xc2b2ae3dU; h[3] ^= h[3] >> 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[2 * j] = (h[j] & 0xFFFFu) >= thresh16 ? keep_scale : 0.f;
      f[2 * j + 1] = (h[j] >> 16) >= thresh16 ? keep_scale : 0.f;
    }
  }
};

constexpr int kBnActDepth = 6;   // work items each thread keeps in flight (PrefetchRing, ew_common.cuh)

// RELU == 1: out = dropout(relu(z * scale + shift)) + res      (models/vnet.py:51-59)
// RELU == 2: out = relu(z * scale + shift + res)                (ResidualBlock, models/mod.py:84);  RELU == 0: no ReLU
template <int RELU, bool DROP, bool HAS_RES>
__global__ void __launch_bounds__(kThreads)
bn_act_apply_kernel(const __nv_bfloat16* __restrict__ z, int z_cs, const float* __restrict__ scale,
                    const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res, int res_cs,
                    __nv_bfloat16* __restrict__ out, int out_cs, long long npix, int C, uint32_t drop_thresh,
                    float drop_scale, uint32_t seed, const long long* __restrict__ step_counter) {
  extern __shared__ uint4 ring_smem[];
  constexpr int NV = HAS_RES ? 2 : 1, DEPTH = kBnActDepth;
  const PrefetchRing<NV, DEPTH> ring(ring_smem);   // vector 0: z, 1: res
  if (DROP && step_counter) seed ^= mix32(static_cast<uint32_t>(*step_counter) + 0x632be5abU);
  const DropMask mask(seed, drop_thresh, drop_scale);
  const int groups = C / 8;
  const int gshift = __ffs(groups) - 1;
  const long long tid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  const int cg = static_cast<int>(tid & (groups - 1));
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale ? scale[cg * 8 + k] : 1.f; sh[k] = shift ? shift[cg * 8 + k] : 0.f; }
  const long long total = npix * groups;
  auto fetch = [&](int stage, long long i) {
    const long long pix = i >> gshift;
    ring.fetch(stage, 0, z + pix * z_cs + cg * 8);
    if (HAS_RES) ring.fetch(stage, 1, res + pix * res_cs + cg * 8);
  };
  long long inext = tid;
#pragma unroll
  for (int s = 0; s < DEPTH; ++s) {
    if (inext < total) fetch(s, inext);
    cp_async_commit();
    inext += stride;
  }
  int stage = 0;
  for (long long i = tid; i < total; i += stride) {
    cp_async_wait<DEPTH - 1>();
    float v[8], rr[8], f[8];
    unpack8(ring.get(stage, 0), v);
    if (HAS_RES) unpack8(ring.get(stage, 1), rr);
    if (inext < total) fetch(stage, inext);     // the item is in registers: refill its slots
    cp_async_commit();
    inext += stride;
    stage = stage + 1 == DEPTH ? 0 : stage + 1;
    const long long pix = i >> gshift;
    if (DROP) mask.factors(static_cast<unsigned long long>(pix) * C + cg * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = fmaf(v[k], sc[k], sh[k]);
      if (RELU == 1) a = fmaxf(a, 0.f);
      if (DROP) a *= f[k];
      if (HAS_RES) a += rr[k];
      if (RELU == 2) a = fmaxf(a, 0.f);
      v[k] = a;
    }
    stg16(out + pix * out_cs + cg * 8, pack8(v));
  }
}

// Backward of a = dropout(relu(bn(z))): dy = da * keep/(1-p) * (bn(z) > 0)   (RELU: the ReLU sits before the dropout).
// APPLY = false: partial [grid][2][C] = sum dy, sum dy * xhat;   APPLY = true: dz = c0 (dy - c1 - xhat c2), partial
// [grid][C] = sum dz (gradient of the conv bias). Rows beyond the grid are zero-filled (kEwBlocks rows in total).
template <bool APPLY, bool RELU, bool DROP>
__global__ void __launch_bounds__(kThreads)
bn_act_bwd_kernel(const __nv_bfloat16* __restrict__ da, int da_cs, const __nv_bfloat16* __restrict__ z, int z_cs,
                  const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ coef, __nv_bfloat16* __restrict__ dz,
                  int dz_cs, float* __restrict__ partial, long long npix, int C, uint32_t drop_thresh,
                  float drop_scale, uint32_t seed, const long long* __restrict__ step_counter) {
  extern __shared__ uint4 ring_smem[];      // prefetch ring; reused for the block reduction after the loop
  constexpr int DEPTH = kBnActDepth;
  const PrefetchRing<2, DEPTH> ring(ring_smem);    // vector 0: z, 1: da
  if (DROP && step_counter) seed ^= mix32(static_cast<uint32_t>(*step_counter) + 0x632be5abU);
  const DropMask mask(seed, drop_thresh, drop_scale);
  const int groups = C / 8;
  const int gshift = __ffs(groups) - 1;
  const long long tid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  const int cg = static_cast<int>(tid & (groups - 1));
  float sc[8], sh[8], mu[8], is[8], c0[8], c1[8], c2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cg * 8 + k;
    mu[k] = mean[c]; is[k] = invstd[c];
    if (RELU) { sc[k] = scale[c]; sh[k] = shift[c]; }
    if (APPLY) { c0[k] = coef[c]; c1[k] = coef[C + c]; c2[k] = coef[2 * C + c]; }
  }
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  const long long total = npix * groups;
  auto fetch = [&](int stage, long long i) {
    const long long pix = i >> gshift;
    ring.fetch(stage, 0, z + pix * z_cs + cg * 8);
    ring.fetch(stage, 1, da + pix * da_cs + cg * 8);
  };
  long long inext = tid;
#pragma unroll
  for (int s = 0; s < DEPTH; ++s) {
    if (inext < total) fetch(s, inext);
    cp_async_commit();
    inext += stride;
  }
  int stage = 0;
  for (long long i = tid; i < total; i += stride) {
    cp_async_wait<DEPTH - 1>();
    float zv[8], g[8], outv[8], f[8];
    unpack8(ring.get(stage, 0), zv);
    unpack8(ring.get(stage, 1), g);
    if (inext < total) fetch(stage, inext);     // the item is in registers: refill its slots
    cp_async_commit();
    inext += stride;
    stage = stage + 1 == DEPTH ? 0 : stage + 1;
    const long long pix = i >> gshift;
    if (DROP) mask.factors(static_cast<unsigned long long>(pix) * C + cg * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float d = g[k];
      if (DROP) d *= f[k];
      if (RELU && !(fmaf(zv[k], sc[k], sh[k]) > 0.f)) d = 0.f;
      const float xh = (zv[k] - mu[k]) * is[k];
      if (!APPLY) {
        acc[k] += d;
        acc[8 + k] = fmaf(d, xh, acc[8 + k]);
      } else {
        const float o = bf16_round(c0[k] * (d - c1[k] - xh * c2[k]));
        outv[k] = o;
        acc[k] += o;
      }
    }
    if (APPLY) stg16(dz + pix * dz_cs + cg * 8, pack8(outv));
  }
  float* red = reinterpret_cast<float*>(ring_smem);
  cp_async_wait<0>();
  constexpr int NV = APPLY ? 8 : 16;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) red[threadIdx.x * NV + k] = acc[k];
  __syncthreads();
  const int per_group = kThreads / groups;
  const int nout = APPLY ? C : 2 * C;
  float* row = partial + static_cast<size_t>(blockIdx.x) * nout;
  for (int o = threadIdx.x; o < nout; o += kThreads) {
    const int which = o / C, ch = o - which * C;
    const int gi = ch / 8, k = ch % 8;
    float s = 0.f;
    for (int t = 0; t < per_group; ++t) s += red[(t * groups + gi) * NV + which * 8 + k];
    row[o] = s;
  }
  for (int rr = blockIdx.x + gridDim.x; rr < kEwBlocks; rr += gridDim.x)
    for (int o = threadIdx.x; o < nout; o += kThreads) partial[static_cast<size_t>(rr) * nout + o] = 0.f;
}

// partial [kEwBlocks][C] = per-block sums over pixels of x (bias gradients of convs that are not followed by BN)
constexpr int kChannelSumsDepth = 8;

__global__ void __launch_bounds__(kThreads)
channel_sums_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, float* __restrict__ partial, long long npix, int C) {
  extern __shared__ uint4 ring_smem[];      // prefetch ring; reused for the block reduction after the loop
  constexpr int DEPTH = kChannelSumsDepth;
  const PrefetchRing<1, DEPTH> ring(ring_smem);
  const int groups = C / 8;
  const int gshift = __ffs(groups) - 1;
  const long long tid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  const int cg = static_cast<int>(tid & (groups - 1));
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const long long total = npix * groups;
  long long inext = tid;
#pragma unroll
  for (int s = 0; s < DEPTH; ++s) {
    if (inext < total) ring.fetch(s, 0, x + (inext >> gshift) * x_cs + cg * 8);
    cp_async_commit();
    inext += stride;
  }
  int stage = 0;
  for (long long i = tid; i < total; i += stride) {
    cp_async_wait<DEPTH - 1>();
    float v[8];
    unpack8(ring.get(stage, 0), v);
    if (inext < total) ring.fetch(stage, 0, x + (inext >> gshift) * x_cs + cg * 8);
    cp_async_commit();
    inext += stride;
    stage = stage + 1 == DEPTH ? 0 : stage + 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
  }
  float* red = reinterpret_cast<float*>(ring_smem);
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  const int per_group = kThreads / groups;
  float* row = partial + static_cast<size_t>(blockIdx.x) * C;
  for (int o = threadIdx.x; o < C; o += kThreads) {
    const int gi = o / 8, k = o % 8;
    float s = 0.f;
    for (int t = 0; t < per_group; ++t) s += red[(t * groups + gi) * 8 + k];
    row[o] = s;
  }
}

// dst [N,2Hs,2Ws,C]: dst[n,2i,2j] = src[n,i,j], zero elsewhere (zero insertion: stride-2 conv backward)
__global__ void __launch_bounds__(kThreads)
upsample_zero2x_kernel(const __nv_bfloat16* __restrict__ src, int src_cs, __nv_bfloat16* __restrict__ dst, int dst_cs,
                       int N, int Hs, int Ws, int C) {
  const int groups = C / 8;
  const int Hd = 2 * Hs, Wd = 2 * Ws;
  const long long total = static_cast<long long>(N) * Hd * Wd * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int cg = static_cast<int>(i % groups);
    const long long pix = i / groups;
    const int w = static_cast<int>(pix % Wd);
    const long long t = pix / Wd;
    const int h = static_cast<int>(t % Hd);
    const long long n = t / Hd;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(h & 1) && !(w & 1)) v = ldg16(src + ((n * Hs + (h >> 1)) * Ws + (w >> 1)) * src_cs + cg * 8);
    stg16(dst + pix * dst_cs + cg * 8, v);
  }
}

// dx = dy * (y > 0): gradient through a ReLU given its OUTPUT (ResidualBlock's ReLU after the add, models/mod.py:84)
__global__ void __launch_bounds__(kThreads)
relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_cs, const __nv_bfloat16* __restrict__ y, int y_cs,
                __nv_bfloat16* __restrict__ dx, int dx_cs, long long npix, int C) {
  const int groups = C / 8;
  const long long total = npix * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int cg = static_cast<int>(i % groups);
    const long long pix = i / groups;
    float g[8], v[8];
    unpack8(ldg16(dy + pix * dy_cs + cg * 8), g);
    unpack8(ldg16(y + pix * y_cs + cg * 8), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : 0.f;
    stg16(dx + pix * dx_cs + cg * 8, pack8(g));
  }
}

// F.max_pool2d(x, 2) backward: the whole gradient goes to the FIRST maximum of each 2x2 window in row-major order
// (torch semantics); dx is written densely (zeros elsewhere).
__global__ void __launch_bounds__(kThreads)
maxpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, const __nv_bfloat16* __restrict__ dpool, int dp_cs,
                      __nv_bfloat16* __restrict__ dx, int dx_cs, int N, int H, int W, int C) {
  const int groups = C / 8;
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Ho * Wo * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int cg = static_cast<int>(i % groups);
    const long long pp = i / groups;
    const int wo = static_cast<int>(pp % Wo);
    const long long t = pp / Wo;
    const int ho = static_cast<int>(t % Ho);
    const long long n = t / Ho;
    const long long p00 = (n * H + 2 * ho) * W + 2 * wo;
    float v[4][8], dp[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) unpack8(ldg16(x + (p00 + (q >> 1) * W + (q & 1)) * x_cs + cg * 8), v[q]);
    unpack8(ldg16(dpool + pp * dp_cs + cg * 8), dp);
    float o[4][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int arg = 0;
      float best = v[0][k];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q][k] > best) { best = v[q][k]; arg = q; }
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][k] = arg == q ? dp[k] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) stg16(dx + (p00 + (q >> 1) * W + (q & 1)) * dx_cs + cg * 8, pack8(o[q]));
  }
}

// ---- squeeze-and-excitation (models/vnet.py:5-26) ------------------------------------------------------------
// Pixels of one sample reduced by one block: 256 at the deep levels (few pixels per sample), more at high resolution
// so that a block streams >= 0.5 MB instead of paying a launch + block reduction per 32 KB.
__host__ __device__ constexpr int se_pix_per_block(long long HW) {
  return HW >= (1ll << 18) ? 4096 : HW >= (1ll << 16) ? 1024 : HW >= (1ll << 12) ? 256 : 64;   // >= 16 blocks per sample
}

// partial[n][chunk][C] = sum over the chunk's pixels of x (DOT: of x * y). The loads go through the per-thread
// prefetch ring: with plain register loads ptxas interleaved each load with its adds (35 registers, one or two loads in
// flight per thread: 3.4 TB/s at 512^2 and 0.5 TB/s at 32^2, where only 64 blocks exist).
constexpr int kSePoolDepth = 8;

template <bool DOT>
__global__ void __launch_bounds__(kThreads)
se_pool_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, const __nv_bfloat16* __restrict__ y, int y_cs,
               float* __restrict__ partial, long long HW, int C, int chunks) {
  extern __shared__ uint4 ring_smem[];      // prefetch ring; reused for the block reduction after the loop
  constexpr int NV = DOT ? 2 : 1, DEPTH = kSePoolDepth;
  const PrefetchRing<NV, DEPTH> ring(ring_smem);
  const int groups = C / 8;                 // power of two <= 256 (host-checked)
  const int cg = threadIdx.x % groups;
  const int pl = threadIdx.x / groups;
  const int ppi = kThreads / groups;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int ppb = se_pix_per_block(HW);
  const long long p0 = static_cast<long long>(chunk) * ppb;
  const long long p1 = min(p0 + ppb, HW);
  const long long base = static_cast<long long>(n) * HW;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  auto fetch = [&](int stage, long long p) {
    ring.fetch(stage, 0, x + (base + p) * x_cs + cg * 8);
    if (DOT) ring.fetch(stage, 1, y + (base + p) * y_cs + cg * 8);
  };
  long long pnext = p0 + pl;
#pragma unroll
  for (int s = 0; s < DEPTH; ++s) {
    if (pnext < p1) fetch(s, pnext);
    cp_async_commit();
    pnext += ppi;
  }
  int stage = 0;
  for (long long pcur = p0 + pl; pcur < p1; pcur += ppi) {
    cp_async_wait<DEPTH - 1>();
    float v[8], w8[8];
    unpack8(ring.get(stage, 0), v);
    if (DOT) unpack8(ring.get(stage, 1), w8);
    if (pnext < p1) fetch(stage, pnext);     // the pixel is in registers: refill its slot
    cp_async_commit();
    pnext += ppi;
    stage = stage + 1 == DEPTH ? 0 : stage + 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = DOT ? fmaf(v[k], w8[k], acc[k]) : acc[k] + v[k];
  }
  float* red = reinterpret_cast<float*>(ring_smem);
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  float* row = partial + (static_cast<size_t>(n) * chunks + chunk) * C;
  for (int o = threadIdx.x; o < C; o += kThreads) {
    const int gi = o / 8, k = o % 8;
    float s = 0.f;
    for (int t = 0; t < ppi; ++t) s += red[(t * groups + gi) * 8 + k];
    row[o] = s;
  }
}

// ---- the two tiny FC layers of the SE block ---------------------------------------------------------------------
// One block per sample walking W1 and W2 (2 MB at C = 1024) through a single SM took 73 us per call; the three phases
// are separate launches spread over the whole chip instead: chunk sums -> mean (or the gate gradient), then one warp
// per output row (coalesced over the contraction index, shuffle reduction) for W v, and lanes over the output index
// with the contraction split across the warps of a block for W^T v.

// out[n][c] = (sum_k partial[n][k][c]) * mult, where mult = scale (gate == nullptr) or g (1 - g) with g = gate[n][c]
__global__ void __launch_bounds__(kThreads)
se_chunk_sum_kernel(const float* __restrict__ partial, int chunks, int C, float scale, const float* __restrict__ gate,
                    float* __restrict__ out) {
  const int c = blockIdx.x * kThreads + threadIdx.x, n = blockIdx.y;
  if (c >= C) return;
  const float* src = partial + static_cast<size_t>(n) * chunks * C + c;
  float s = 0.f;
  int k = 0;
  for (; k + 8 <= chunks; k += 8) {     // eight independent loads, added in chunk order
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(src + static_cast<size_t>(k + j) * C);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  for (; k < chunks; ++k) s += __ldg(src + static_cast<size_t>(k) * C);
  float mult = scale;
  if (gate) { const float g = gate[static_cast<size_t>(n) * C + c]; mult = g * (1.f - g); }
  out[static_cast<size_t>(n) * C + c] = s * mult;
}

// out[n][r] = act(sum_k W[r][k] in[n][k] + bias[r]); W [R][K] row-major. One warp per (n, r). ACT 1: ReLU, 2: sigmoid.
template <int ACT>
__global__ void __launch_bounds__(kThreads)
se_matvec_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ in,
                 float* __restrict__ out, int R, int K) {
  const int r = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* wr = W + static_cast<size_t>(r) * K;
  const float* v = in + static_cast<size_t>(n) * K;
  float a = 0.f;
  for (int k = lane; k < K; k += 32) a = fmaf(__ldg(wr + k), __ldg(v + k), a);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
  if (lane == 0) {
    a += bias[r];
    out[static_cast<size_t>(n) * R + r] = ACT == 1 ? fmaxf(a, 0.f) : 1.f / (1.f + expf(-a));
  }
}

// out[n][j] = mask(sum_k W[k][j] in[n][k]); W [K][J] row-major (the transposed product). Block = 32 output columns x 8
// slices of k; mask_src != nullptr: zero where mask_src[n][j] <= 0 (gradient through the hidden layer's ReLU).
__global__ void __launch_bounds__(kThreads)
se_matvec_t_kernel(const float* __restrict__ W, const float* __restrict__ in, const float* __restrict__ mask_src,
                   float* __restrict__ out, int K, int J) {
  __shared__ float red[kThreads / 32][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane, n = blockIdx.y;
  const float* v = in + static_cast<size_t>(n) * K;
  float a = 0.f;
  if (j < J)
    for (int k = slice; k < K; k += kThreads / 32) a = fmaf(__ldg(W + static_cast<size_t>(k) * J + j), __ldg(v + k), a);
  red[slice][lane] = a;
  __syncthreads();
  if (slice == 0 && j < J) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < kThreads / 32; ++t) s += red[t][lane];
    if (mask_src && !(mask_src[static_cast<size_t>(n) * J + j] > 0.f)) s = 0.f;
    out[static_cast<size_t>(n) * J + j] = s;
  }
}

// y = x * gate[n][c]
__global__ void __launch_bounds__(kThreads)
se_scale_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, const float* __restrict__ gate,
                const float* __restrict__ add, float add_scale, __nv_bfloat16* __restrict__ y, int y_cs, long long HW,
                int C, long long npix) {
  // forward: y = x * gate;  backward (add != nullptr): dx = dy * gate + add[n][c] * add_scale, with x := dy
  const int groups = C / 8;
  const bool pow2 = (groups & (groups - 1)) == 0;
  const int gshift = __ffs(groups) - 1;
  const long long total = npix * groups;
  const long long tid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  constexpr int U = 4;   // independent 16-byte loads in flight per thread
  for (long long i0 = tid; i0 < total; i0 += stride * U) {
    uint4 raw[U];
    long long pixs[U];
    int cgs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      const long long ic = i < total ? i : i0;
      pixs[u] = pow2 ? ic >> gshift : ic / groups;
      cgs[u] = static_cast<int>(pow2 ? ic & (groups - 1) : ic - pixs[u] * groups);
      raw[u] = ldg16(x + pixs[u] * x_cs + cgs[u] * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= total) break;
      const long long n = static_cast<unsigned long long>(pixs[u]) / static_cast<unsigned long long>(HW);
      float v[8];
      unpack8(raw[u], v);
      const float4* gp = reinterpret_cast<const float4*>(gate + n * C + cgs[u] * 8);
      const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
      v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w; v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
      if (add) {
        const float4* ap = reinterpret_cast<const float4*>(add + n * C + cgs[u] * 8);
        const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
        v[0] = fmaf(a0.x, add_scale, v[0]); v[1] = fmaf(a0.y, add_scale, v[1]); v[2] = fmaf(a0.z, add_scale, v[2]);
        v[3] = fmaf(a0.w, add_scale, v[3]); v[4] = fmaf(a1.x, add_scale, v[4]); v[5] = fmaf(a1.y, add_scale, v[5]);
        v[6] = fmaf(a1.z, add_scale, v[6]); v[7] = fmaf(a1.w, add_scale, v[7]);
      }
      stg16(y + pixs[u] * y_cs + cgs[u] * 8, pack8(v));
    }
  }
}

// out[r][k] = sum_n a[n][r] * b[n][k]  (batch-summed outer products: the SE weight gradients), out [R][K];
// bias gradients: out_bias[r] = sum_n a[n][r]
__global__ void __launch_bounds__(kThreads)
outer_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                 float* __restrict__ out_bias, int N, int R, int K) {
  const long long total = static_cast<long long>(R) * K;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int r = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<long long>(r) * K);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(a[static_cast<size_t>(n) * R + r], b[static_cast<size_t>(n) * K + k], s);
    out[i] = s;
    if (k == 0 && out_bias) {
      float sb = 0.f;
      for (int n = 0; n < N; ++n) sb += a[static_cast<size_t>(n) * R + r];
      out_bias[r] = sb;
    }
  }
}

}  // namespace b2s

using namespace b2s;
#define STREAM(s) static_cast<cudaStream_t>(s)

// 16-bit threshold: P(drop) = thresh / 65536 (p is quantised to 1/65536; the kept values are scaled by the exact
// reciprocal of the quantised keep probability so the mask stays unbiased)
static uint32_t drop_threshold(float p) {
  if (p <= 0.f) return 0u;
  double t = static_cast<double>(p) * 65536.0 + 0.5;
  if (t < 1.0) t = 1.0;
  if (t > 65535.0) t = 65535.0;
  return static_cast<uint32_t>(t);
}
static float drop_scale_of(float p) {
  const uint32_t t = drop_threshold(p);
  return t ? static_cast<float>(65536.0 / (65536.0 - t)) : 1.f;
}

// one launch of a templated streaming kernel with its prefetch ring: opt into the shared memory, one resident wave
template <typename K, typename... Args>
static void launch_ring(K kernel, int smem, long long items, int per_block, cudaStream_t stream, Args... args) {
  ew_allow_smem(kernel, smem);
  // occupancy of each template instance, looked up once per host thread (instances of one signature share this function)
  thread_local const void* seen[32];
  thread_local int waves[32];
  thread_local int nseen = 0;
  int wave = 0;
  for (int i = 0; i < nseen; ++i)
    if (seen[i] == reinterpret_cast<const void*>(kernel)) wave = waves[i];
  if (!wave) {
    wave = ew_wave_blocks(kernel, smem);
    if (nseen < 32) { seen[nseen] = reinterpret_cast<const void*>(kernel); waves[nseen] = wave; ++nseen; }
  }
  const int grid = items < 0 ? wave : ew_clamp_grid(wave, items, per_block);
  kernel<<<grid, kThreads, smem, stream>>>(args...);
}

template <int RELU, bool DROP>
static void launch_bn_act_apply(bool has_res, long long items, cudaStream_t st, const __nv_bfloat16* z, int z_cs,
                                const float* scale, const float* shift, const __nv_bfloat16* res, int res_cs,
                                __nv_bfloat16* out, int out_cs, long long npix, int C, uint32_t thresh, float dscale,
                                uint32_t seed, const long long* step_counter) {
  if (has_res)
    launch_ring(bn_act_apply_kernel<RELU, DROP, true>, PrefetchRing<2, kBnActDepth>::kBytes, items, kThreads * 4, st, z,
                z_cs, scale, shift, res, res_cs, out, out_cs, npix, C, thresh, dscale, seed, step_counter);
  else
    launch_ring(bn_act_apply_kernel<RELU, DROP, false>, PrefetchRing<1, kBnActDepth>::kBytes, items, kThreads * 4, st, z,
                z_cs, scale, shift, res, res_cs, out, out_cs, npix, C, thresh, dscale, seed, step_counter);
}

extern "C" int b2s_bn_act_apply(const void* z, int z_cstride, const float* scale, const float* shift, const void* res,
                                int res_cstride, void* out, int out_cstride, long long npix, int C, int relu,
                                float dropout_p, unsigned seed, const long long* step_counter, void* stream) {
  if (!z || !out) return set_error(B2S_ERR_ARG, "b2s_bn_act_apply: null pointer");
  if (!ew_channels_supported(C)) return set_error(B2S_ERR_ARG, "b2s_bn_act_apply: C/8 must be a power of two <= 256");
  if (z_cstride % 8 || out_cstride % 8 || (res && res_cstride % 8))
    return set_error(B2S_ERR_ARG, "b2s_bn_act_apply: strides must be multiples of 8");
  if (dropout_p < 0.f || dropout_p >= 1.f) return set_error(B2S_ERR_ARG, "b2s_bn_act_apply: dropout_p in [0,1)");
  if (relu < 0 || relu > 2) return set_error(B2S_ERR_ARG, "b2s_bn_act_apply: relu mode must be 0, 1 or 2");
  count_launch();
  const uint32_t thresh = drop_threshold(dropout_p);
  const float dscale = drop_scale_of(dropout_p);
  const auto* zp = static_cast<const __nv_bfloat16*>(z);
  const auto* rp = static_cast<const __nv_bfloat16*>(res);
  auto* op = static_cast<__nv_bfloat16*>(out);
  const long long items = npix * (C / 8);
  cudaStream_t st = STREAM(stream);
#define B2S_APPLY(R, D) launch_bn_act_apply<R, D>(res != nullptr, items, st, zp, z_cstride, scale, shift, rp, res_cstride, \
                                                  op, out_cstride, npix, C, thresh, dscale, seed, step_counter)
  if (thresh) { if (relu == 1) B2S_APPLY(1, true); else if (relu == 2) B2S_APPLY(2, true); else B2S_APPLY(0, true); }
  else        { if (relu == 1) B2S_APPLY(1, false); else if (relu == 2) B2S_APPLY(2, false); else B2S_APPLY(0, false); }
#undef B2S_APPLY
  return check_launch("bn_act_apply_kernel");
}

template <bool APPLY>
static void launch_bn_act_bwd(bool relu, bool drop, cudaStream_t st, const __nv_bfloat16* da, int da_cs,
                              const __nv_bfloat16* z, int z_cs, const float* scale, const float* shift, const float* mean,
                              const float* invstd, const float* coef, __nv_bfloat16* dz, int dz_cs, float* partial,
                              long long npix, int C, uint32_t thresh, float dscale, uint32_t seed,
                              const long long* step_counter) {
  constexpr int smem = PrefetchRing<2, kBnActDepth>::kBytes;
#define B2S_BWD(R, D) launch_ring(bn_act_bwd_kernel<APPLY, R, D>, smem, -1, 1, st, da, da_cs, z, z_cs, scale, shift, mean, \
                                  invstd, coef, dz, dz_cs, partial, npix, C, thresh, dscale, seed, step_counter)
  if (relu) { if (drop) B2S_BWD(true, true); else B2S_BWD(true, false); }
  else      { if (drop) B2S_BWD(false, true); else B2S_BWD(false, false); }
#undef B2S_BWD
}

extern "C" int b2s_bn_act_bwd_reduce(const void* da, int da_cstride, const void* z, int z_cstride, const float* scale,
                                     const float* shift, const float* mean, const float* invstd, float* partial,
                                     long long npix, int C, int relu, float dropout_p, unsigned seed,
                                     const long long* step_counter, void* stream) {
  if (!da || !z || !scale || !shift || !mean || !invstd || !partial)
    return set_error(B2S_ERR_ARG, "b2s_bn_act_bwd_reduce: null pointer");
  if (!ew_channels_supported(C)) return set_error(B2S_ERR_ARG, "b2s_bn_act_bwd_reduce: unsupported C");
  if (da_cstride % 8 || z_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_bn_act_bwd_reduce: strides must be multiples of 8");
  count_launch();
  const uint32_t thresh = drop_threshold(dropout_p);
  launch_bn_act_bwd<false>(relu == 1, thresh != 0, STREAM(stream), static_cast<const __nv_bfloat16*>(da), da_cstride,
                           static_cast<const __nv_bfloat16*>(z), z_cstride, scale, shift, mean, invstd, nullptr, nullptr, 0,
                           partial, npix, C, thresh, drop_scale_of(dropout_p), seed, step_counter);
  return check_launch("bn_act_bwd_kernel<reduce>");
}

extern "C" int b2s_bn_act_bwd_apply(const void* da, int da_cstride, const void* z, int z_cstride, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, const float* coef,
                                    void* dz, int dz_cstride, float* dbias_partial, long long npix, int C, int relu,
                                    float dropout_p, unsigned seed, const long long* step_counter, void* stream) {
  if (!da || !z || !scale || !shift || !mean || !invstd || !coef || !dz || !dbias_partial)
    return set_error(B2S_ERR_ARG, "b2s_bn_act_bwd_apply: null pointer");
  if (!ew_channels_supported(C)) return set_error(B2S_ERR_ARG, "b2s_bn_act_bwd_apply: unsupported C");
  if (da_cstride % 8 || z_cstride % 8 || dz_cstride % 8)
    return set_error(B2S_ERR_ARG, "b2s_bn_act_bwd_apply: strides must be multiples of 8");
  count_launch();
  const uint32_t thresh = drop_threshold(dropout_p);
  launch_bn_act_bwd<true>(relu == 1, thresh != 0, STREAM(stream), static_cast<const __nv_bfloat16*>(da), da_cstride,
                          static_cast<const __nv_bfloat16*>(z), z_cstride, scale, shift, mean, invstd, coef,
                          static_cast<__nv_bfloat16*>(dz), dz_cstride, dbias_partial, npix, C, thresh,
                          drop_scale_of(dropout_p), seed, step_counter);
  return check_launch("bn_act_bwd_kernel<apply>");
}

extern "C" int b2s_relu_bwd(const void* dy, int dy_cstride, const void* y, int y_cstride, void* dx, int dx_cstride,
                            long long npix, int C, void* stream) {
  if (!dy || !y || !dx) return set_error(B2S_ERR_ARG, "b2s_relu_bwd: null pointer");
  if (C % 8 || dy_cstride % 8 || y_cstride % 8 || dx_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_relu_bwd: need multiples of 8");
  count_launch();
  relu_bwd_kernel<<<ew_grid_for(npix * (C / 8), kThreads * 4), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), dy_cstride, static_cast<const __nv_bfloat16*>(y), y_cstride,
      static_cast<__nv_bfloat16*>(dx), dx_cstride, npix, C);
  return check_launch("relu_bwd_kernel");
}

extern "C" int b2s_maxpool2x2_bwd(const void* x, int x_cstride, const void* dpool, int dpool_cstride, void* dx,
                                  int dx_cstride, int N, int H, int W, int C, void* stream) {
  if (!x || !dpool || !dx) return set_error(B2S_ERR_ARG, "b2s_maxpool2x2_bwd: null pointer");
  // odd H / W: the last row / column receives no gradient; the caller zero-fills dx in that case
  if (C % 8 || x_cstride % 8 || dpool_cstride % 8 || dx_cstride % 8 || H < 2 || W < 2)
    return set_error(B2S_ERR_ARG, "b2s_maxpool2x2_bwd: need C % 8 == 0, H, W >= 2");
  const long long items = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  count_launch();
  maxpool2x2_bwd_kernel<<<ew_grid_for(items, kThreads * 2), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cstride, static_cast<const __nv_bfloat16*>(dpool), dpool_cstride,
      static_cast<__nv_bfloat16*>(dx), dx_cstride, N, H, W, C);
  return check_launch("maxpool2x2_bwd_kernel");
}

extern "C" int b2s_channel_sums(const void* x, int x_cstride, float* partial, long long npix, int C, void* stream) {
  if (!x || !partial) return set_error(B2S_ERR_ARG, "b2s_channel_sums: null pointer");
  if (!ew_channels_supported(C)) return set_error(B2S_ERR_ARG, "b2s_channel_sums: unsupported C");
  if (x_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_channel_sums: stride must be a multiple of 8");
  count_launch();
  constexpr int smem = PrefetchRing<1, kChannelSumsDepth>::kBytes;
  channel_sums_kernel<<<kEwBlocks, kThreads, smem, STREAM(stream)>>>(static_cast<const __nv_bfloat16*>(x), x_cstride,
                                                                    partial, npix, C);
  return check_launch("channel_sums_kernel");
}

extern "C" int b2s_upsample_zero2x(const void* src, int src_cstride, void* dst, int dst_cstride, int N, int Hs, int Ws,
                                   int C, void* stream) {
  if (!src || !dst) return set_error(B2S_ERR_ARG, "b2s_upsample_zero2x: null pointer");
  if (C % 8 || src_cstride % 8 || dst_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_upsample_zero2x: need multiples of 8");
  const long long items = static_cast<long long>(N) * 4 * Hs * Ws * (C / 8);
  count_launch();
  upsample_zero2x_kernel<<<ew_grid_for(items, kThreads * 4), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), src_cstride, static_cast<__nv_bfloat16*>(dst), dst_cstride, N, Hs, Ws,
      C);
  return check_launch("upsample_zero2x_kernel");
}

extern "C" int b2s_se_chunks(long long HW) {
  const int ppb = se_pix_per_block(HW);
  return static_cast<int>((HW + ppb - 1) / ppb);
}

// partial [N][b2s_se_chunks(HW)][C]: sums of x over H*W (y == NULL) or of x*y (the gate gradient) per sample, channel
extern "C" int b2s_se_pool(const void* x, int x_cstride, const void* y, int y_cstride, float* partial, int N,
                           long long HW, int C, void* stream) {
  if (!x || !partial) return set_error(B2S_ERR_ARG, "b2s_se_pool: null pointer");
  if (!ew_channels_supported(C)) return set_error(B2S_ERR_ARG, "b2s_se_pool: unsupported C");
  if (x_cstride % 8 || (y && y_cstride % 8)) return set_error(B2S_ERR_ARG, "b2s_se_pool: strides must be multiples of 8");
  if (N <= 0 || N > 65535) return set_error(B2S_ERR_ARG, "b2s_se_pool: bad batch");
  const int chunks = b2s_se_chunks(HW);
  dim3 grid(chunks, N);
  count_launch();
  constexpr int smem_dot = PrefetchRing<2, kSePoolDepth>::kBytes, smem_sum = PrefetchRing<1, kSePoolDepth>::kBytes;
  if (y) {
    ew_allow_smem(se_pool_kernel<true>, smem_dot);
    se_pool_kernel<true><<<grid, kThreads, smem_dot, STREAM(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), x_cstride, static_cast<const __nv_bfloat16*>(y), y_cstride, partial, HW, C,
        chunks);
  } else {
    ew_allow_smem(se_pool_kernel<false>, smem_sum);
    se_pool_kernel<false><<<grid, kThreads, smem_sum, STREAM(stream)>>>(static_cast<const __nv_bfloat16*>(x), x_cstride,
                                                                        nullptr, 0, partial, HW, C, chunks);
  }
  return check_launch("se_pool_kernel");
}

extern "C" int b2s_se_fc_fwd(const float* partial, int chunks, long long HW, const float* w1, const float* b1,
                             const float* w2, const float* b2, float* mean, float* hidden, float* gate, int N, int C,
                             int Cr, void* stream) {
  if (!partial || !w1 || !b1 || !w2 || !b2 || !mean || !hidden || !gate)
    return set_error(B2S_ERR_ARG, "b2s_se_fc_fwd: null pointer");
  if (C > 4096 || Cr < 1 || Cr > C) return set_error(B2S_ERR_ARG, "b2s_se_fc_fwd: unsupported channel counts");
  // mean = pooled / HW; hidden = relu(W1 mean + b1); gate = sigmoid(W2 hidden + b2)      (models/vnet.py:20-25)
  cudaStream_t st = STREAM(stream);
  count_launch();
  se_chunk_sum_kernel<<<dim3((C + kThreads - 1) / kThreads, N), kThreads, 0, st>>>(
      partial, chunks, C, 1.f / static_cast<float>(HW), nullptr, mean);
  count_launch();
  se_matvec_kernel<1><<<dim3((Cr + 7) / 8, N), kThreads, 0, st>>>(w1, b1, mean, hidden, Cr, C);
  count_launch();
  se_matvec_kernel<2><<<dim3((C + 7) / 8, N), kThreads, 0, st>>>(w2, b2, hidden, gate, C, Cr);
  return check_launch("se_fc_fwd kernels");
}

extern "C" int b2s_se_scale(const void* x, int x_cstride, const float* gate, const float* add, float add_scale, void* y,
                            int y_cstride, int N, long long HW, int C, void* stream) {
  if (!x || !gate || !y) return set_error(B2S_ERR_ARG, "b2s_se_scale: null pointer");
  if (C % 8 || x_cstride % 8 || y_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_se_scale: need multiples of 8");
  const long long npix = static_cast<long long>(N) * HW;
  count_launch();
  se_scale_kernel<<<ew_grid_for(npix * (C / 8), kThreads * 4), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cstride, gate, add, add_scale, static_cast<__nv_bfloat16*>(y),
      y_cstride, HW, C, npix);
  return check_launch("se_scale_kernel");
}

extern "C" int b2s_se_fc_bwd(const float* partial, int chunks, const float* gate, const float* hidden, const float* mean,
                             const float* w1, const float* w2, float* ds, float* dh, float* dmean, float* dw1, float* db1,
                             float* dw2, float* db2, int N, int C, int Cr, void* stream) {
  if (!partial || !gate || !hidden || !mean || !w1 || !w2 || !ds || !dh || !dmean || !dw1 || !db1 || !dw2 || !db2)
    return set_error(B2S_ERR_ARG, "b2s_se_fc_bwd: null pointer");
  if (C > 4096 || Cr < 1 || Cr > C) return set_error(B2S_ERR_ARG, "b2s_se_fc_bwd: unsupported channel counts");
  // ds = dgate g (1 - g) with dgate = the pooled dy * x;  dh = (h > 0) W2^T ds;  dmean = W1^T dh
  cudaStream_t st = STREAM(stream);
  count_launch();
  se_chunk_sum_kernel<<<dim3((C + kThreads - 1) / kThreads, N), kThreads, 0, st>>>(partial, chunks, C, 1.f, gate, ds);
  count_launch();
  se_matvec_t_kernel<<<dim3((Cr + 31) / 32, N), kThreads, 0, st>>>(w2, ds, hidden, dh, C, Cr);
  count_launch();
  se_matvec_t_kernel<<<dim3((C + 31) / 32, N), kThreads, 0, st>>>(w1, dh, nullptr, dmean, Cr, C);
  int rc = check_launch("se_fc_bwd kernels");
  if (rc) return rc;
  // dW2 [C][Cr] = sum_n ds (x) hidden, db2 = sum_n ds;  dW1 [Cr][C] = sum_n dh (x) mean, db1 = sum_n dh
  count_launch();
  outer_sum_kernel<<<ew_grid_for(static_cast<long long>(C) * Cr, kThreads), kThreads, 0, STREAM(stream)>>>(
      ds, hidden, dw2, db2, N, C, Cr);
  count_launch();
  outer_sum_kernel<<<ew_grid_for(static_cast<long long>(C) * Cr, kThreads), kThreads, 0, STREAM(stream)>>>(
      dh, mean, dw1, db1, N, Cr, C);
  return check_launch("outer_sum_kernel");
}
