// Bandwidth-bound kernels of the UNet hot path (sm_100a): first conv (Cin=1), BatchNorm finalize / apply /
// backward, 2x2 max-pool (fused into BN apply and BN backward), 1x1 head + threshold, Dice+BCE(+FocalTversky)
// loss and gradient, weight packing, wgrad second-stage reduce, AdamW. All activations NHWC bf16, 16-byte
// vectorised (8 channels per thread), grids sized in multiples of the SM count.
#include "ew_common.cuh"

namespace b2s {

// ------------------------------------------------------------------------------------------------
// generic deterministic row reduction: out[k] = sum_r in[r][k]
// ------------------------------------------------------------------------------------------------
constexpr int kRedY = 32;  // row groups per block of the small reduction kernels (block = 32 x 32 threads)

// Rows a thread fetches before it starts adding: the loads of one batch are independent, so a thread waits for one
// memory round trip per kRedBatch rows instead of one per row (these kernels are latency-, not bandwidth-bound).
constexpr int kRedBatch = 8;

__global__ void __launch_bounds__(32 * kRedY)
reduce_rows_kernel(const float* __restrict__ in, int rows, int K, int rows_per_slice, float* __restrict__ out) {
  // grid (ceil(K/32), slices); block (32, kRedY): lane = column (coalesced), threadIdx.y strides over the rows
  __shared__ double red[kRedY][33];
  const int k = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_slice;
  const int r1 = min(r0 + rows_per_slice, rows);
  double acc = 0.0;
  if (k < K)
    for (int r = r0 + threadIdx.y; r < r1; r += kRedY * kRedBatch) {
      float v[kRedBatch];
#pragma unroll
      for (int j = 0; j < kRedBatch; ++j) {
        const int rr = r + j * kRedY;          // rows past the end re-read the last row (unconditional load) and add 0
        const float x = __ldg(in + static_cast<size_t>(min(rr, r1 - 1)) * K + k);
        v[j] = rr < r1 ? x : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kRedBatch; ++j) acc += static_cast<double>(v[j]);   // same order as row by row
    }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && k < K) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kRedY; ++j) s += red[j][threadIdx.x];
    out[static_cast<size_t>(blockIdx.y) * K + k] = static_cast<float>(s);
  }
}

// Sum of partial[r][which][c] over r for which in {0,1}; block (32, kRedY), channel c = blockIdx.x*32 + threadIdx.x.
// Result valid in threads with threadIdx.y == 0.
__device__ __forceinline__ void block_sum_pairs(const float* __restrict__ partial, int rows, int C, int c, double& s,
                                                double& q) {
  __shared__ double red[2][kRedY][33];
  double a0 = 0.0, a1 = 0.0;
  if (c < C)
    for (int r = threadIdx.y; r < rows; r += kRedY * kRedBatch) {
      float v0[kRedBatch], v1[kRedBatch];
#pragma unroll
      for (int j = 0; j < kRedBatch; ++j) {
        const int rr = r + j * kRedY;          // rows past the end re-read the last row (unconditional load) and add 0
        const float* row = partial + (static_cast<size_t>(min(rr, rows - 1)) * 2) * C + c;
        const float x0 = __ldg(row), x1 = __ldg(row + C);
        v0[j] = rr < rows ? x0 : 0.f;
        v1[j] = rr < rows ? x1 : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kRedBatch; ++j) { a0 += static_cast<double>(v0[j]); a1 += static_cast<double>(v1[j]); }
    }
  red[0][threadIdx.y][threadIdx.x] = a0;
  red[1][threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  s = 0.0; q = 0.0;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < kRedY; ++j) { s += red[0][j][threadIdx.x]; q += red[1][j][threadIdx.x]; }
  }
}

// Reduce [rows][K] down to at most 128 rows (in scratch) when rows > 1024; returns pointer/rows to finalize from.
int reduce_to_small(const float* in, int rows, int K, float* scratch, const float** out_ptr, int* out_rows,
                           cudaStream_t stream) {
  if (rows <= 1024) { *out_ptr = in; *out_rows = rows; return B2S_OK; }
  if (!scratch) return set_error(B2S_ERR_ARG, "scratch buffer required for rows > 1024");
  const int slices = 128;
  const int rps = (rows + slices - 1) / slices;
  const int used = (rows + rps - 1) / rps;
  dim3 grid((K + 31) / 32, used), block(32, kRedY);
  count_launch();
  reduce_rows_kernel<<<grid, block, 0, stream>>>(in, rows, K, rps, scratch);
  int rc = check_launch("reduce_rows_kernel");
  *out_ptr = scratch; *out_rows = used;
  return rc;
}

// ------------------------------------------------------------------------------------------------
// first conv: Cin = 1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
conv3x3_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      __nv_bfloat16* __restrict__ r, float* __restrict__ stats_partial, int N, int H, int W, int Cout,
                      int flags, const float* __restrict__ post_scale, const float* __restrict__ post_shift) {
  __shared__ float red[kThreads * 16];
  const int groups = Cout / 8;              // threads per pixel
  const int ppi = kThreads / groups;        // pixels per block iteration
  const int cg = threadIdx.x % groups;
  const int pl = threadIdx.x / groups;
  float wr[9][8], br[8], ps[8], pt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int co = cg * 8 + k;
    br[k] = bias ? bias[co] : 0.f;
    ps[k] = post_scale ? post_scale[co] : 1.f;
    pt[k] = post_scale ? post_shift[co] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][k] = w[co * 9 + t];
  }
  float st[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) st[k] = 0.f;
  const unsigned npix = static_cast<unsigned>(N) * H * W;   // host guarantees N*H*W < 2^31
  const bool relu = flags & B2S_FLAG_RELU;
  for (unsigned p0 = blockIdx.x * ppi; p0 < npix; p0 += gridDim.x * ppi) {
    const unsigned p = p0 + pl;
    if (p < npix) {
      const unsigned row = p / W;
      const int wq = static_cast<int>(p - row * W);
      const int hq = static_cast<int>(row % H);
      const float* xi = x + (p - wq - static_cast<unsigned>(hq) * W);  // image base
      float xv[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = hq + t / 3 - 1, ww = wq + t % 3 - 1;
        xv[t] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + hh * W + ww) : 0.f;
      }
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a = br[k];
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(xv[t], wr[t][k], a);
        if (relu) a = fmaxf(a, 0.f);
        if (post_scale) a = fmaf(a, ps[k], pt[k]);
        a = bf16_round(a);
        acc[k] = a;
        st[k] += a;
        st[8 + k] = fmaf(a, a, st[8 + k]);
      }
      stg16(r + static_cast<size_t>(p) * Cout + cg * 8, pack8(acc));
    }
  }
  if (stats_partial) {
    // out row layout [2][Cout]: thread-group g holds channels g*8..g*8+7 -> sum at [g*8+k], sumsq at [Cout+g*8+k]
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) red[threadIdx.x * 16 + k] = st[k];
    __syncthreads();
    float* row = stats_partial + static_cast<size_t>(blockIdx.x) * 2 * Cout;
    for (int o = threadIdx.x; o < 2 * Cout; o += kThreads) {
      const int which = o / Cout, ch = o - which * Cout;
      const int g = ch / 8, k = ch % 8;
      float acc = 0.f;
      for (int t = 0; t < ppi; ++t) acc += red[(t * groups + g) * 16 + which * 8 + k];
      row[o] = acc;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
conv3x3_c1_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dz,
                        float* __restrict__ partial, int N, int H, int W, int Cout) {
  __shared__ float red[kThreads * 8];
  const int groups = Cout / 8;
  const int ppi = kThreads / groups;
  const int cg = threadIdx.x % groups;
  const int pl = threadIdx.x / groups;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  const unsigned npix = static_cast<unsigned>(N) * H * W;
  for (unsigned p0 = blockIdx.x * ppi; p0 < npix; p0 += gridDim.x * ppi) {
    const unsigned p = p0 + pl;
    if (p < npix) {
      const unsigned row = p / W;
      const int wq = static_cast<int>(p - row * W);
      const int hq = static_cast<int>(row % H);
      const float* xi = x + (p - wq - static_cast<unsigned>(hq) * W);
      float g[8];
      unpack8(ldg16(dz + static_cast<size_t>(p) * Cout + cg * 8), g);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = hq + t / 3 - 1, ww = wq + t % 3 - 1;
        const float xv = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + hh * W + ww) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(xv, g[k], acc[t][k]);
      }
    }
  }
  float* row = partial + static_cast<size_t>(blockIdx.x) * Cout * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[t][k];
    __syncthreads();
    for (int o = threadIdx.x; o < Cout; o += kThreads) {
      const int g = o / 8, k = o % 8;
      float s = 0.f;
      for (int tt = 0; tt < ppi; ++tt) s += red[(tt * groups + g) * 8 + k];
      row[o * 9 + t] = s;
    }
  }
}

// Four-pixels-per-thread variants (W % 4 == 0): each thread owns 8 output channels of 4 consecutive pixels of one
// image row, so the 3x6 input window is read once (three aligned float4 + six scalars) for 288 FMAs.
__device__ __forceinline__ void c1_load_window(const float* __restrict__ xi, int hq, int wq, int H, int W,
                                               float (&xv)[3][6]) {
#pragma unroll
  for (int dr = 0; dr < 3; ++dr) {
    const int hh = hq + dr - 1;
    if (hh >= 0 && hh < H) {
      const float* rowp = xi + static_cast<long long>(hh) * W + wq;
      const float4 c = __ldg(reinterpret_cast<const float4*>(rowp));
      xv[dr][0] = wq > 0 ? __ldg(rowp - 1) : 0.f;
      xv[dr][1] = c.x; xv[dr][2] = c.y; xv[dr][3] = c.z; xv[dr][4] = c.w;
      xv[dr][5] = wq + 4 < W ? __ldg(rowp + 4) : 0.f;
    } else {
#pragma unroll
      for (int dc = 0; dc < 6; ++dc) xv[dr][dc] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2)
conv3x3_c1_fwd4_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       __nv_bfloat16* __restrict__ r, float* __restrict__ stats_partial, int N, int H, int W, int Cout,
                       int flags, const float* __restrict__ post_scale, const float* __restrict__ post_shift) {
  // weights live in shared memory ([tap][Cout], read as two broadcast float4 per tap): keeping all 72 of a thread's
  // weights in registers cost 191 registers and one resident block per SM
  __shared__ float red[kThreads * 16];
  __shared__ __align__(16) float wsm[9 * 128];
  const int groups = Cout / 8;
  const int qpi = kThreads / groups;        // pixel quads per block iteration
  const int cg = threadIdx.x % groups;
  const int ql = threadIdx.x / groups;
  for (int i = threadIdx.x; i < 9 * Cout; i += kThreads) {
    const int t = i / Cout, co = i - t * Cout;
    wsm[t * Cout + co] = w[co * 9 + t];
  }
  float br[8], ps[8], pt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    br[k] = bias ? bias[cg * 8 + k] : 0.f;
    ps[k] = post_scale ? post_scale[cg * 8 + k] : 1.f;
    pt[k] = post_scale ? post_shift[cg * 8 + k] : 0.f;
  }
  __syncthreads();
  float st[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) st[k] = 0.f;
  const unsigned nquads = static_cast<unsigned>(N) * H * W / 4;
  const bool relu = flags & B2S_FLAG_RELU;
  for (unsigned q0 = blockIdx.x * qpi; q0 < nquads; q0 += gridDim.x * qpi) {
    const unsigned qd = q0 + ql;
    if (qd < nquads) {
      const unsigned p = qd * 4;
      const unsigned row = p / W;
      const int wq = static_cast<int>(p - row * W);
      const int hq = static_cast<int>(row % H);
      const float* xi = x + (p - wq - static_cast<unsigned>(hq) * W);  // image base
      float xv[3][6];
      c1_load_window(xi, hq, wq, H, W, xv);
      // the 288 FMAs of the work item as 144 packed FFMA2 (one issue slot per channel pair; bit-identical to scalar
      // FMAs: 244 -> 221 us at batch 64 @256^2, the kernel is instruction-issue bound)
      float acc[4][8];
      f32x2 acc2[4][4];
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) acc2[px][kp] = f2_pack(br[2 * kp], br[2 * kp + 1]);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(&wsm[t * Cout + cg * 8]);      // channels 0..3
        const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(&wsm[t * Cout + cg * 8 + 4]);  // channels 4..7
        const f32x2 wp[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float v = xv[t / 3][px + t % 3];
          const f32x2 vv = f2_pack(v, v);
#pragma unroll
          for (int kp = 0; kp < 4; ++kp) acc2[px][kp] = f2_fma(vv, wp[kp], acc2[px][kp]);
        }
      }
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) f2_unpack(acc2[px][kp], acc[px][2 * kp], acc[px][2 * kp + 1]);
#pragma unroll
      for (int px = 0; px < 4; ++px) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float a = acc[px][k];
          if (relu) a = fmaxf(a, 0.f);
          if (post_scale) a = fmaf(a, ps[k], pt[k]);
          a = bf16_round(a);
          acc[px][k] = a;
          st[k] += a;
          st[8 + k] = fmaf(a, a, st[8 + k]);
        }
        stg16(r + static_cast<size_t>(p + px) * Cout + cg * 8, pack8(acc[px]));
      }
    }
  }
  if (stats_partial) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) red[threadIdx.x * 16 + k] = st[k];
    __syncthreads();
    float* row = stats_partial + static_cast<size_t>(blockIdx.x) * 2 * Cout;
    for (int o = threadIdx.x; o < 2 * Cout; o += kThreads) {
      const int which = o / Cout, ch = o - which * Cout;
      const int g = ch / 8, k = ch % 8;
      float acc = 0.f;
      for (int t = 0; t < qpi; ++t) acc += red[(t * groups + g) * 16 + which * 8 + k];
      row[o] = acc;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
conv3x3_c1_wgrad4_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dz,
                         float* __restrict__ partial, int N, int H, int W, int Cout) {
  __shared__ float red[kThreads * 8];
  const int groups = Cout / 8;
  const int qpi = kThreads / groups;
  const int cg = threadIdx.x % groups;
  const int ql = threadIdx.x / groups;
  f32x2 acc2[9][4];    // [tap][channel pair]: the weight-gradient sums as packed fp32 pairs
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) acc2[t][kp] = 0ull;
  const unsigned nquads = static_cast<unsigned>(N) * H * W / 4;
  for (unsigned q0 = blockIdx.x * qpi; q0 < nquads; q0 += gridDim.x * qpi) {
    const unsigned qd = q0 + ql;
    if (qd < nquads) {
      const unsigned p = qd * 4;
      const unsigned row = p / W;
      const int wq = static_cast<int>(p - row * W);
      const int hq = static_cast<int>(row % H);
      const float* xi = x + (p - wq - static_cast<unsigned>(hq) * W);
      uint4 graw[4];
#pragma unroll
      for (int px = 0; px < 4; ++px) graw[px] = ldg16(dz + static_cast<size_t>(p + px) * Cout + cg * 8);
      float xv[3][6];
      c1_load_window(xi, hq, wq, H, W, xv);
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        float g[8];
        unpack8(graw[px], g);
        const f32x2 g2[4] = {f2_pack(g[0], g[1]), f2_pack(g[2], g[3]), f2_pack(g[4], g[5]), f2_pack(g[6], g[7])};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float v = xv[t / 3][px + t % 3];
          const f32x2 vv = f2_pack(v, v);
#pragma unroll
          for (int kp = 0; kp < 4; ++kp) acc2[t][kp] = f2_fma(vv, g2[kp], acc2[t][kp]);   // FFMA2: 190 -> 162 us
        }
      }
    }
  }
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) f2_unpack(acc2[t][kp], acc[t][2 * kp], acc[t][2 * kp + 1]);
  float* row = partial + static_cast<size_t>(blockIdx.x) * Cout * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[t][k];
    __syncthreads();
    for (int o = threadIdx.x; o < Cout; o += kThreads) {
      const int g = o / 8, k = o % 8;
      float s = 0.f;
      for (int tt = 0; tt < qpi; ++tt) s += red[(tt * groups + g) * 8 + k];
      row[o * 9 + t] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_var, long long* nbt, float momentum, float eps,
                                   float* scale, float* shift, float* mean_out, float* invstd_out) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  block_sum_pairs(partial, rows, C, c, s, q);
  if (threadIdx.y != 0) return;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - static_cast<float>(mean) * sc;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

__global__ void bn_eval_affine_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                      float eps, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(rv[c] + eps);
  const float sc = (gamma ? gamma[c] : 1.f) * invstd;
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
}

// ------------------------------------------------------------------------------------------------
// BN apply (+ fused 2x2 max-pool)
// ------------------------------------------------------------------------------------------------
template <bool POOL> constexpr int bn_apply_depth() { return POOL ? 4 : 8; }
template <bool POOL> constexpr int bn_apply_nv() { return POOL ? 4 : 1; }

template <bool POOL>
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const __nv_bfloat16* __restrict__ r, int r_cs, const float* __restrict__ scale,
                const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, int y_cs,
                __nv_bfloat16* __restrict__ pooled, int N, int H, int W, int C) {
  extern __shared__ uint4 ring_smem[];
  constexpr int Q = bn_apply_nv<POOL>(), DEPTH = bn_apply_depth<POOL>();
  const PrefetchRing<Q, DEPTH> ring(ring_smem);
  const int groups = C / 8;               // power of two (host-checked)
  const int gshift = __ffs(groups) - 1;
  const long long tid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;  // multiple of groups
  const int cg = static_cast<int>(tid & (groups - 1));
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
  const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
  const long long total = static_cast<long long>(N) * Ho * Wo * groups;
  // first pixel of work item i (a pixel, or the top-left pixel of a 2x2 pooling window)
  auto first_pixel = [&](long long i) -> long long {
    const unsigned pp = static_cast<unsigned>(i >> gshift);   // (pooled) pixel index < 2^31
    if (!POOL) return pp;
    const unsigned prow = pp / Wo;
    const int wo = static_cast<int>(pp - prow * Wo);
    const unsigned n = prow / Ho;
    const int ho = static_cast<int>(prow - n * Ho);
    return (static_cast<long long>(n) * H + 2 * ho) * W + 2 * wo;
  };
  auto fetch = [&](int stage, long long i) {
    const long long p00 = first_pixel(i);
#pragma unroll
    for (int q = 0; q < Q; ++q) ring.fetch(stage, q, r + (p00 + (q >> 1) * W + (q & 1)) * r_cs + cg * 8);
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
    uint4 raw[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) raw[q] = ring.get(stage, q);
    if (inext < total) fetch(stage, inext);     // the item is in registers: refill its slots
    cp_async_commit();
    inext += stride;
    stage = stage + 1 == DEPTH ? 0 : stage + 1;
    const long long p00 = first_pixel(i);
    if (!POOL) {
      float v[8];
      unpack8(raw[0], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
      stg16(y + p00 * y_cs + cg * 8, pack8(v));
    } else {
      float m[8];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const long long pix = p00 + (q >> 1) * W + (q & 1);
        float v[8];
        unpack8(raw[q], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = bf16_round(fmaf(v[k], sc[k], sh[k]));
        stg16(y + pix * y_cs + cg * 8, pack8(v));
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = q == 0 ? v[k] : fmaxf(m[k], v[k]);
      }
      stg16(pooled + static_cast<size_t>(i >> gshift) * C + cg * 8, pack8(m));
    }
  }
}

// plain F.max_pool2d(x, 2) (inference path, where BN is already applied by the producing conv's epilogue)
__global__ void __launch_bounds__(kThreads)
maxpool2x2_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, __nv_bfloat16* __restrict__ pooled, int N, int H, int W,
                  int C) {
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
    float m[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
      unpack8(ldg16(x + (p00 + (q >> 1) * W + (q & 1)) * x_cs + cg * 8), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = q == 0 ? v[k] : fmaxf(m[k], v[k]);
    }
    stg16(pooled + pp * C + cg * 8, pack8(m));
  }
}

// ------------------------------------------------------------------------------------------------
// BN (+ReLU, + max-pool routing) backward
// ------------------------------------------------------------------------------------------------
template <bool POOL, bool APPLY> constexpr int bn_bwd_depth() { return POOL ? (APPLY ? 4 : 2) : 6; }
template <bool POOL> constexpr int bn_bwd_nv() { return POOL ? 9 : 2; }

template <bool POOL, bool APPLY>
__global__ void __launch_bounds__(kThreads, POOL ? (APPLY ? 1 : 2) : (APPLY ? 3 : 4))
bn_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_cs, const __nv_bfloat16* __restrict__ dpool,
              const __nv_bfloat16* __restrict__ r, int r_cs, const float* __restrict__ scale,
              const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
              const float* __restrict__ coef, __nv_bfloat16* __restrict__ dz, int dz_cs, float* __restrict__ partial,
              int N, int H, int W, int C) {
  extern __shared__ uint4 ring_smem[];      // prefetch ring; reused for the block reduction after the loop
  constexpr int Q = POOL ? 4 : 1;           // pixels per work item (one 2x2 pooling window, or one pixel)
  constexpr int NV = bn_bwd_nv<POOL>(), DEPTH = bn_bwd_depth<POOL, APPLY>();
  static_assert(PrefetchRing<NV, DEPTH>::kBytes >= kThreads * 16 * 4, "ring too small for the block reduction");
  const PrefetchRing<NV, DEPTH> ring(ring_smem);   // vectors [0,Q): r, [Q,2Q): dy, 2Q: dpool
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
    if (POOL) { sc[k] = scale[c]; sh[k] = shift[c]; }
    if (APPLY) { c0[k] = coef[c]; c1[k] = coef[C + c]; c2[k] = coef[2 * C + c]; }
  }
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;

  const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
  const long long total = static_cast<long long>(N) * Ho * Wo * groups;
  auto first_pixel = [&](long long i) -> long long {
    const unsigned pp = static_cast<unsigned>(i >> gshift);
    if (!POOL) return pp;
    const unsigned prow = pp / Wo;
    const int wo = static_cast<int>(pp - prow * Wo);
    const unsigned n = prow / Ho;
    const int ho = static_cast<int>(prow - n * Ho);
    return (static_cast<long long>(n) * H + 2 * ho) * W + 2 * wo;
  };
  auto fetch = [&](int stage, long long i) {
    const long long p00 = first_pixel(i);
    if (POOL) ring.fetch(stage, 2 * Q, dpool + static_cast<size_t>(i >> gshift) * C + cg * 8);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const long long pix = p00 + (q >> 1) * W + (q & 1);
      ring.fetch(stage, q, r + pix * r_cs + cg * 8);
      ring.fetch(stage, Q + q, dy + pix * dy_cs + cg * 8);
    }
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
    float rv[Q][8], g[Q][8], dp[8];
#pragma unroll
    for (int q = 0; q < Q; ++q) { unpack8(ring.get(stage, q), rv[q]); unpack8(ring.get(stage, Q + q), g[q]); }
    if (POOL) unpack8(ring.get(stage, 2 * Q), dp);
    if (inext < total) fetch(stage, inext);     // the item is in registers: refill its slots
    cp_async_commit();
    inext += stride;
    stage = stage + 1 == DEPTH ? 0 : stage + 1;
    const long long p00 = first_pixel(i);
    // With POOL, dpool goes to the FIRST maximum of y = bf16(r*scale+shift) in row-major window order (torch
    // max_pool2d backward semantics).
    if (POOL) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float best = bf16_round(fmaf(rv[0][k], sc[k], sh[k]));
        int arg = 0;
#pragma unroll
        for (int q = 1; q < Q; ++q) {
          const float yv = bf16_round(fmaf(rv[q][k], sc[k], sh[k]));
          if (yv > best) { best = yv; arg = q; }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) g[q][k] += (arg == q) ? dp[k] : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      float out[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (rv[q][k] - mu[k]) * is[k];
        if (!APPLY) {
          acc[k] += g[q][k];
          acc[8 + k] = fmaf(g[q][k], xh, acc[8 + k]);
        } else {
          float d = c0[k] * (g[q][k] - c1[k] - xh * c2[k]);
          d = rv[q][k] > 0.f ? d : 0.f;
          d = bf16_round(d);
          out[k] = d;
          acc[k] += d;
        }
      }
      if (APPLY) stg16(dz + (p00 + (q >> 1) * W + (q & 1)) * dz_cs + cg * 8, pack8(out));
    }
  }
  // block partials: !APPLY -> row [2][C] (sum dy, sum dy*xhat); APPLY -> row [C] (sum dz = conv bias gradient)
  constexpr int NV_RED = APPLY ? 8 : 16;
  float* red = reinterpret_cast<float*>(ring_smem);
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV_RED; ++k) red[threadIdx.x * NV_RED + k] = acc[k];
  __syncthreads();
  const int per_group = kThreads / groups;
  const int nout = APPLY ? C : 2 * C;
  float* row = partial + static_cast<size_t>(blockIdx.x) * nout;
  // NOTE: all threads of a block share tid % groups == threadIdx.x % groups because kThreads % groups == 0
  for (int o = threadIdx.x; o < nout; o += kThreads) {
    const int which = o / C, ch = o - which * C;
    const int gi = ch / 8, k = ch % 8;
    float s = 0.f;
    for (int t = 0; t < per_group; ++t) s += red[(t * groups + gi) * NV_RED + which * 8 + k];
    row[o] = s;
  }
  // the partial buffer always has kEwBlocks rows; rows beyond the (occupancy-sized) grid are zero-filled
  for (int rr = blockIdx.x + gridDim.x; rr < kEwBlocks; rr += gridDim.x)
    for (int o = threadIdx.x; o < nout; o += kThreads) partial[static_cast<size_t>(rr) * nout + o] = 0.f;
}

// mean != nullptr: the second sum is the RAW sum(dy * r) of the fused dgrad epilogue (b2s_conv_dgrad_bnred), turned
// into sum(dy * xhat) = invstd * (sum(dy * r) - mean * sum(dy)) here
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ invstd,
                                       float* dgamma, float* dbeta, float* coef, const float* __restrict__ mean = nullptr) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  block_sum_pairs(partial, rows, C, c, s, q);
  if (threadIdx.y != 0 || c >= C) return;
  if (mean) q = static_cast<double>(invstd[c]) * (q - static_cast<double>(mean[c]) * s);
  if (dgamma) dgamma[c] = static_cast<float>(q);
  if (dbeta) dbeta[c] = static_cast<float>(s);
  coef[c] = (gamma ? gamma[c] : 1.f) * invstd[c];
  coef[C + c] = static_cast<float>(s / count);
  coef[2 * C + c] = static_cast<float>(q / count);
}

// ------------------------------------------------------------------------------------------------
// head: folded BN + 1x1 conv (+ threshold); backward
// ------------------------------------------------------------------------------------------------
// 8 lanes cooperate on one pixel (C == 64): lane j holds channels 8j..8j+7.
template <bool O1>   // O1: one output channel (the reference's UNet(out_channels=1)): weights in registers, logits[p]
__global__ void __launch_bounds__(kThreads)
head_fwd_kernel(const __nv_bfloat16* __restrict__ r, int r_cs, const float* __restrict__ scale,
                const float* __restrict__ shift, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ logits, unsigned char* __restrict__ mask, int N, long long HW, int C, int O) {
  extern __shared__ float wf[];  // [O][C] folded weights, then [O] folded bias
  for (int i = threadIdx.x; i < O * C; i += kThreads) {
    const int c = i % C;
    wf[i] = w[i] * (scale ? scale[c] : 1.f);
  }
  for (int o = threadIdx.x; o < O; o += kThreads) {
    float acc = b ? b[o] : 0.f;
    if (shift)
      for (int c = 0; c < C; ++c) acc = fmaf(w[o * C + c], shift[c], acc);
    wf[O * C + o] = acc;
  }
  __syncthreads();
  const int groups = C / 8;  // lanes per pixel (power of two <= 32)
  const int cg = threadIdx.x % groups;
  const long long npix = static_cast<long long>(N) * HW;
  const long long ppi = kThreads / groups;
  constexpr int U = 4;       // pixels per lane group and loop trip: four independent 16-byte loads in flight
  float wr[8], fb = 0.f;
  if constexpr (O1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[k] = wf[cg * 8 + k];
    fb = wf[C];
  }
  for (long long p0 = static_cast<long long>(blockIdx.x) * ppi * U; p0 < npix;
       p0 += static_cast<long long>(gridDim.x) * ppi * U) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * ppi + threadIdx.x / groups;
      raw[u] = p < npix ? ldg16(r + p * r_cs + cg * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * ppi + threadIdx.x / groups;
      float v[8];
      unpack8(raw[u], v);
      if constexpr (O1) {      // same FMA chain and lane reduction as the general path; no shared-memory reads, no division
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(v[k], wr[k], acc);
        for (int off = groups >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (cg == 0 && p < npix) {
          const float logit = acc + fb;
          logits[p] = logit;
          if (mask) mask[p] = (1.f / (1.f + expf(-logit))) > 0.5f ? 1 : 0;
        }
      } else {
        for (int o = 0; o < O; ++o) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc = fmaf(v[k], wf[o * C + cg * 8 + k], acc);
          for (int off = groups >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
          if (cg == 0 && p < npix) {
            const float logit = acc + wf[O * C + o];
            const long long n = static_cast<unsigned>(p) / static_cast<unsigned>(HW), hw = p - n * HW;
            const long long oi = (n * O + o) * HW + hw;
            logits[oi] = logit;
            if (mask) mask[oi] = (1.f / (1.f + expf(-logit))) > 0.5f ? 1 : 0;
          }
        }
      }
    }
  }
}

template <bool O1>   // O1: one output channel: weights in registers, dlogits[p], two pixels in flight per thread
__global__ void __launch_bounds__(kThreads)
head_bwd_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ r, int r_cs,
                const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ w,
                __nv_bfloat16* __restrict__ dy, int dy_cs, float* __restrict__ partial, int N, long long HW, int C,
                int O) {
  __shared__ float red[kThreads * 9];
  const int groups = C / 8;
  const int cg = threadIdx.x % groups;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = scale ? scale[cg * 8 + k] : 1.f;
    sh[k] = shift ? shift[cg * 8 + k] : 0.f;
  }
  const long long npix = static_cast<long long>(N) * HW;
  const long long ppi = kThreads / groups;
  float* row = partial + static_cast<size_t>(blockIdx.x) * (O * C + O);
  // pass over output channels one at a time for dW/db; dy accumulates over o in registers when O == 1,
  // otherwise dy is recomputed (O is 1 for the reference's UNet(in_channels=1,out_channels=1)).
  for (int o = 0; o < O; ++o) {
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    if constexpr (O1) {
      float wr[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) wr[k] = w[cg * 8 + k];
      constexpr int U = 2;
      const long long step = static_cast<long long>(gridDim.x) * ppi;
      for (long long p0 = static_cast<long long>(blockIdx.x) * ppi + threadIdx.x / groups; p0 < npix; p0 += step * U) {
        uint4 raw[U];
        float g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long p = p0 + u * step;
          const bool on = p < npix;
          raw[u] = on ? ldg16(r + p * r_cs + cg * 8) : make_uint4(0, 0, 0, 0);
          g[u] = on ? __ldg(dlogits + p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long p = p0 + u * step;
          if (p >= npix) break;
          float v[8], d[8];
          unpack8(raw[u], v);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            acc[k] = fmaf(g[u], bf16_round(fmaf(v[k], sc[k], sh[k])), acc[k]);
            d[k] = g[u] * wr[k];
          }
          if (cg == 0) acc[8] += g[u];
          stg16(dy + p * dy_cs + cg * 8, pack8(d));
        }
      }
    } else {
      for (long long p0 = static_cast<long long>(blockIdx.x) * ppi; p0 < npix;
           p0 += static_cast<long long>(gridDim.x) * ppi) {
        const long long p = p0 + threadIdx.x / groups;
        if (p < npix) {
          const long long n = static_cast<unsigned>(p) / static_cast<unsigned>(HW), hw = p - n * HW;
          const float g = dlogits[(n * O + o) * HW + hw];
          float v[8];
          unpack8(ldg16(r + p * r_cs + cg * 8), v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(g, bf16_round(fmaf(v[k], sc[k], sh[k])), acc[k]);
          if (cg == 0) acc[8] += g;
          if (o == O - 1) {
            float d[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = 0.f;
            for (int oo = 0; oo < O; ++oo) {
              const float go = dlogits[(n * O + oo) * HW + hw];
#pragma unroll
              for (int k = 0; k < 8; ++k) d[k] = fmaf(go, w[oo * C + cg * 8 + k], d[k]);
            }
            stg16(dy + p * dy_cs + cg * 8, pack8(d));
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 9; ++k) red[threadIdx.x * 9 + k] = acc[k];
    __syncthreads();
    const int per_group = kThreads / groups;
    for (int c = threadIdx.x; c < C; c += kThreads) {
      const int gi = c / 8, k = c % 8;
      float s = 0.f;
      for (int t = 0; t < per_group; ++t) s += red[(t * groups + gi) * 9 + k];
      row[o * C + c] = s;
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int t = 0; t < per_group; ++t) s += red[(t * groups) * 9 + 8];
      row[O * C + o] = s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// loss
// ------------------------------------------------------------------------------------------------
constexpr long long kLossChunk = 16384;  // elements per block

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__global__ void __launch_bounds__(kThreads)
seg_loss_partial_kernel(const float* __restrict__ logits, const float* __restrict__ targets, long long per_sample,
                        int chunks, float* __restrict__ partial) {
  __shared__ float red[4][kThreads / 32];
  const int b = blockIdx.y, ch = blockIdx.x;
  const long long base = static_cast<long long>(b) * per_sample;
  const long long i0 = static_cast<long long>(ch) * kLossChunk;
  const long long i1 = min(i0 + kLossChunk, per_sample);
  float s_pt = 0.f, s_p = 0.f, s_t = 0.f, s_b = 0.f;
  const bool vec = ((per_sample & 3) == 0);
  if (vec) {
    for (long long i = i0 + threadIdx.x * 4; i < i1; i += kThreads * 4) {
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(logits + base + i));
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(targets + base + i));
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ts[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = xs[k], t = ts[k];
        const float e = expf(-fabsf(x));
        const float p = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
        s_pt = fmaf(p, t, s_pt); s_p += p; s_t += t;
        s_b += fmaxf(x, 0.f) - x * t + log1pf(e);
      }
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += kThreads) {
      const float x = logits[base + i], t = targets[base + i];
      const float e = expf(-fabsf(x));
      const float p = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      s_pt = fmaf(p, t, s_pt); s_p += p; s_t += t;
      s_b += fmaxf(x, 0.f) - x * t + log1pf(e);
    }
  }
  s_pt = warp_sum(s_pt); s_p = warp_sum(s_p); s_t = warp_sum(s_t); s_b = warp_sum(s_b);
  const int wi = threadIdx.x >> 5, li = threadIdx.x & 31;
  if (li == 0) { red[0][wi] = s_pt; red[1][wi] = s_p; red[2][wi] = s_t; red[3][wi] = s_b; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kThreads / 32; ++j) s += red[threadIdx.x][j];
    partial[(static_cast<size_t>(b) * chunks + ch) * 4 + threadIdx.x] = s;
  }
}

__global__ void seg_loss_finalize_kernel(const float* __restrict__ partial, int B, int chunks, long long per_sample,
                                         float* __restrict__ sums, float* __restrict__ out, float dice_smooth,
                                         float w_bce, float w_dice, float w_ft, float ft_alpha, float ft_beta,
                                         float ft_gamma, float ft_smooth) {
  // single block; thread b < B reduces its sample
  __shared__ double sh[4][kThreads];
  double acc[4] = {0, 0, 0, 0};
  double dice_sum = 0.0;
  for (int b = threadIdx.x; b < B; b += kThreads) {
    double s[4] = {0, 0, 0, 0};
    for (int c = 0; c < chunks; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += static_cast<double>(partial[(static_cast<size_t>(b) * chunks + c) * 4 + k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) { sums[b * 4 + k] = static_cast<float>(s[k]); acc[k] += s[k]; }
    dice_sum += (2.0 * s[0] + dice_smooth) / (s[1] + s[2] + dice_smooth);
  }
  // reuse sh: 0..2 = TP, sum p, sum t ; 3 = bce ; dice via second pass
#pragma unroll
  for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  __shared__ double dsum[kThreads];
  dsum[threadIdx.x] = dice_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0}, d = 0.0;
    for (int j = 0; j < kThreads; ++j) {
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] += sh[k][j];
      d += dsum[j];
    }
    const double bce = t[3] / (static_cast<double>(B) * static_cast<double>(per_sample));
    const double dice = 1.0 - d / B;
    const double TP = t[0], FP = t[1] - t[0], FN = t[2] - t[0];
    const double ti = (TP + ft_smooth) / (TP + ft_alpha * FP + ft_beta * FN + ft_smooth);
    const double ft = pow(fmax(1.0 - ti, 0.0), static_cast<double>(ft_gamma));
    out[0] = static_cast<float>(w_bce * bce + w_dice * dice + w_ft * ft);
    out[1] = static_cast<float>(bce);
    out[2] = static_cast<float>(dice);
    out[3] = static_cast<float>(ft);
    out[4] = static_cast<float>(t[0]);
    out[5] = static_cast<float>(t[1]);
    out[6] = static_cast<float>(t[2]);
  }
}

__global__ void __launch_bounds__(kThreads)
seg_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ targets,
                    const float* __restrict__ sums, const float* __restrict__ ft_tot, int B, long long per_sample,
                    long long bce_count, int dice_batch, const float* __restrict__ grad_out,
                    float* __restrict__ dlogits, float dice_smooth, float w_bce, float w_dice, float w_ft,
                    float ft_alpha, float ft_beta, float ft_gamma, float ft_smooth) {
  const int b = blockIdx.y;
  const float go = grad_out ? *grad_out : 1.f;
  const float I = sums[b * 4 + 0], U = sums[b * 4 + 1] + sums[b * 4 + 2];
  const float den = U + dice_smooth;
  // d dice_b / d p_i = (2 t_i den - (2I+s)) / den^2 ; loss = 1 - mean_b dice_b
  const float kd_a = -w_dice / dice_batch * 2.f / den;
  const float kd_b = w_dice / dice_batch * (2.f * I + dice_smooth) / (den * den);
  const float kb = w_bce / static_cast<float>(bce_count);
  float kf_t = 0.f, kf_1 = 0.f;
  if (w_ft != 0.f) {
    float TP, SP, ST;
    if (ft_tot) { TP = ft_tot[0]; SP = ft_tot[1]; ST = ft_tot[2]; }
    else {
      TP = SP = ST = 0.f;
      for (int j = 0; j < B; ++j) { TP += sums[j * 4]; SP += sums[j * 4 + 1]; ST += sums[j * 4 + 2]; }
    }
    const float FP = SP - TP, FN = ST - TP;
    const float D = TP + ft_alpha * FP + ft_beta * FN + ft_smooth;
    const float ti = (TP + ft_smooth) / D;
    const float base = fmaxf(1.f - ti, 0.f);
    const float dL_dti = base > 0.f ? -ft_gamma * powf(base, ft_gamma - 1.f) : 0.f;
    // d ti / d p_i = [t_i D - (TP+s)(t_i + alpha (1 - t_i))] / D^2   (FN' = -t_i cancels with beta term below)
    // D' = t_i + alpha (1 - t_i) - beta t_i
    // => d ti/d p_i = t_i * [D - (TP+s)(1 - alpha - beta)] / D^2 - (TP+s) alpha / D^2
    kf_t = w_ft * dL_dti * (D - (TP + ft_smooth) * (1.f - ft_alpha - ft_beta)) / (D * D);
    kf_1 = -w_ft * dL_dti * (TP + ft_smooth) * ft_alpha / (D * D);
  }
  const long long base_i = static_cast<long long>(b) * per_sample;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < per_sample;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float x = logits[base_i + i], t = targets[base_i + i];
    const float e = expf(-fabsf(x));
    const float p = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
    const float dp = p * (1.f - p);
    float g = kb * (p - t);
    g += (kd_a * t + kd_b + kf_t * t + kf_1) * dp;
    dlogits[base_i + i] = go * g;
  }
}

// Segmentation metrics counters (utils/utils.py:225-251, utils/trainer.py:101-107,236-250): pred = sigmoid(x) > 0.5 in
// fp32; the reference compares it with target.astype(int) (truncation) for accuracy / precision / recall and with
// target.astype(bool) (non-zero) for IoU. partial [blocks][6] = {TP, FP, FN, TN (int targets), intersection, union
// (bool targets)} per block; metrics_accumulate adds them to the running int64 counters [6] (+ element count in [6]).
__global__ void __launch_bounds__(kThreads)
seg_metrics_partial_kernel(const float* __restrict__ logits, const float* __restrict__ targets, long long n,
                           unsigned int* __restrict__ partial) {
  __shared__ unsigned int red[6][kThreads / 32];
  unsigned int c[6] = {0, 0, 0, 0, 0, 0};
  const long long i0 = static_cast<long long>(blockIdx.x) * kLossChunk;
  const long long i1 = min(i0 + kLossChunk, n);
  for (long long i = i0 + threadIdx.x; i < i1; i += kThreads) {
    const float x = logits[i], t = targets[i];
    const bool pred = (1.f / (1.f + expf(-x))) > 0.5f;
    const int ti = static_cast<int>(t);          // numpy astype(int): truncation toward zero
    const bool tb = t != 0.f;                    // numpy astype(bool)
    c[0] += (pred && ti == 1); c[1] += (pred && ti == 0); c[2] += (!pred && ti == 1); c[3] += (!pred && ti == 0);
    c[4] += (pred && tb); c[5] += (pred || tb);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    unsigned int v = c[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    unsigned int s = 0;
#pragma unroll
    for (int j = 0; j < kThreads / 32; ++j) s += red[threadIdx.x][j];
    partial[static_cast<size_t>(blockIdx.x) * 6 + threadIdx.x] = s;
  }
}

__global__ void seg_metrics_accumulate_kernel(const unsigned int* __restrict__ partial, int blocks, long long n,
                                              long long* __restrict__ counters) {
  __shared__ long long red[6][kThreads];
  long long c[6] = {0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < blocks; b += kThreads)
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] += partial[static_cast<size_t>(b) * 6 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) red[k][threadIdx.x] = c[k];
  __syncthreads();
  if (threadIdx.x < 6) {
    long long s = 0;
    for (int j = 0; j < kThreads; ++j) s += red[threadIdx.x][j];
    counters[threadIdx.x] += s;
  }
  if (threadIdx.x == 6) counters[6] += n;
}

// ------------------------------------------------------------------------------------------------
// weight packing / wgrad reduce / AdamW / copy
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                        __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int taps) {
  // one block per (32 co x 32 ci) tile: coalesced fp32 reads, both transposed bf16 layouts written in 64-B runs
  __shared__ float tile[32][32 * 9 + 1];
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int nci = min(32, Cin - ci0), nco = min(32, Cout - co0);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int col = ty; col < nco; col += 8) {
    const float* src = w + (static_cast<size_t>(co0 + col) * Cin + ci0) * taps;
    for (int i = tx; i < nci * taps; i += 32) tile[col][i] = src[i];
  }
  __syncthreads();
  for (int t = 0; t < taps; ++t) {
    if (wf)
      for (int col = ty; col < nco; col += 8)
        if (tx < nci)
          wf[(static_cast<size_t>(t) * Cout + co0 + col) * Cin + ci0 + tx] = __float2bfloat16_rn(tile[col][tx * taps + t]);
    if (wd)
      for (int cil = ty; cil < nci; cil += 8)
        if (tx < nco)
          wd[(static_cast<size_t>(taps - 1 - t) * Cin + ci0 + cil) * Cout + co0 + tx] =
              __float2bfloat16_rn(tile[tx][cil * taps + t]);
  }
}

__global__ void pack_convt_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                         __nv_bfloat16* __restrict__ wd, int Cin, int Cout) {
  // w [Cin][Cout][4]; wf [(ab)*Cout+co][Cin]; wd [(ab)*Cin+ci][Cout]
  const long long total = static_cast<long long>(Cin) * Cout * 4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ab = static_cast<int>(i & 3);
    const long long cc = i >> 2;
    const int co = static_cast<int>(cc % Cout);
    const int ci = static_cast<int>(cc / Cout);
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wf) wf[(static_cast<size_t>(ab) * Cout + co) * Cin + ci] = v;
    if (wd) wd[(static_cast<size_t>(ab) * Cin + ci) * Cout + co] = v;
  }
}

// All conv / transposed-conv weights of a network in ONE launch: block -> (tensor, 32x32 tile) through a small table in
// kernel-parameter space. Same tile transposes as the two kernels above.
constexpr int kPackMaxTensors = 40;
struct PackEntry {
  const float* w;
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;
  int d0, d1, taps;   // kind 0: w [d0=Cout][d1=Cin][taps];  kind 1: w [d0=Cin][d1=Cout][4]
  int kind;
  int tile_begin;     // first block of this tensor
};
struct PackTable {
  PackEntry e[kPackMaxTensors];
  int n;
};

__global__ void __launch_bounds__(kThreads)
pack_all_kernel(const __grid_constant__ PackTable tab) {
  __shared__ float tile[32][32 * 9 + 1];
  int ti = 0;
  while (ti + 1 < tab.n && static_cast<int>(blockIdx.x) >= tab.e[ti + 1].tile_begin) ++ti;
  const PackEntry& E = tab.e[ti];
  const int local = blockIdx.x - E.tile_begin;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int taps = E.taps;
  // tiles over (d1 / 32, d0 / 32): tile rows = 32 consecutive d0 indices, each row = 32 d1 indices x taps (contiguous)
  const int tiles1 = (E.d1 + 31) / 32;
  const int i1 = (local % tiles1) * 32, i0 = (local / tiles1) * 32;
  const int n1 = min(32, E.d1 - i1), n0 = min(32, E.d0 - i0);
  for (int rr = ty; rr < n0; rr += 8) {
    const float* src = E.w + (static_cast<size_t>(i0 + rr) * E.d1 + i1) * taps;
    for (int i = tx; i < n1 * taps; i += 32) tile[rr][i] = src[i];
  }
  __syncthreads();
  if (E.kind == 0) {          // rows = co, cols = ci
    const int Cout = E.d0, Cin = E.d1;
    for (int t = 0; t < taps; ++t) {
      if (E.wf)
        for (int col = ty; col < n0; col += 8)
          if (tx < n1)
            E.wf[(static_cast<size_t>(t) * Cout + i0 + col) * Cin + i1 + tx] = __float2bfloat16_rn(tile[col][tx * taps + t]);
      if (E.wd)
        for (int cil = ty; cil < n1; cil += 8)
          if (tx < n0)
            E.wd[(static_cast<size_t>(taps - 1 - t) * Cin + i1 + cil) * Cout + i0 + tx] =
                __float2bfloat16_rn(tile[tx][cil * taps + t]);
    }
  } else {                    // rows = ci, cols = co; wf [(t*Cout+co)][Cin], wd [(t*Cin+ci)][Cout]
    const int Cin = E.d0, Cout = E.d1;
    for (int t = 0; t < 4; ++t) {
      if (E.wf)
        for (int col = ty; col < n1; col += 8)
          if (tx < n0)
            E.wf[(static_cast<size_t>(t) * Cout + i1 + col) * Cin + i0 + tx] = __float2bfloat16_rn(tile[tx][col * 4 + t]);
      if (E.wd)
        for (int row = ty; row < n0; row += 8)
          if (tx < n1)
            E.wd[(static_cast<size_t>(t) * Cin + i0 + row) * Cout + i1 + tx] = __float2bfloat16_rn(tile[row][tx * 4 + t]);
    }
  }
}

// ws [splits][TAPS][Cin][Cout] -> layout 0: dw[co][ci][t]; layout 1: dw[ci][co][t].
// Block = 8 warps x (32 lanes = 128 co as float4). The warps are split into (8/sgroups) input channels x sgroups
// split groups, so that shallow layers (tiny K, many K-splits) and deep layers (huge K, 1-2 splits) both stream with
// TAPS independent float4 loads in flight per thread. TAPS and sgroups are compile-time / powers of two: the index
// arithmetic of the transposing write-out was two runtime divisions per element and made the kernel instruction-bound
// (1024 -> 1024, one split: 47 us for 75 MB).
template <int TAPS>
__global__ void __launch_bounds__(kThreads)
wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int Cin, int Cout, float* __restrict__ dw, int layout,
                    int sgroups) {
  __shared__ float tile[TAPS][8][129];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int sshift = __ffs(sgroups) - 1;             // sgroups in {1, 2, 4, 8}
  const int cpb = 8 >> sshift;                       // input channels per block
  const int cil = ty >> sshift, sg = ty & (sgroups - 1);
  const int ci0 = blockIdx.x * cpb, co0 = blockIdx.y * 128;
  const int ci = ci0 + cil, co = co0 + tx * 4;
  const size_t tap_stride = static_cast<size_t>(Cin) * Cout;
  const size_t split_stride = static_cast<size_t>(TAPS) * tap_stride;
  float4 acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ci < Cin && co < Cout) {
    const float* base = ws + static_cast<size_t>(ci) * Cout + co;
    for (int sp = sg; sp < splits; sp += sgroups) {
      float4 v[TAPS];
#pragma unroll
      for (int t = 0; t < TAPS; ++t) v[t] = __ldg(reinterpret_cast<const float4*>(base + sp * split_stride + t * tap_stride));
#pragma unroll
      for (int t = 0; t < TAPS; ++t) { acc[t].x += v[t].x; acc[t].y += v[t].y; acc[t].z += v[t].z; acc[t].w += v[t].w; }
    }
  }
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    tile[t][ty][tx * 4 + 0] = acc[t].x; tile[t][ty][tx * 4 + 1] = acc[t].y;
    tile[t][ty][tx * 4 + 2] = acc[t].z; tile[t][ty][tx * 4 + 3] = acc[t].w;
  }
  __syncthreads();
  const int nci = min(cpb, Cin - ci0), nco = min(128, Cout - co0);
  const int cshift = 3 - sshift;                     // log2(cpb)
  // e enumerates (outer, c, t) with t fastest: outer = output channel (layout 0) or nothing (layout 1, see below)
  if (layout == 0) {
    // dw[co][ci0 .. ci0+nci)[t]: nci*TAPS contiguous floats per output channel; thread <-> (col, c, t)
    for (int e = threadIdx.x; e < (nco * TAPS) << cshift; e += kThreads) {
      const int ct = e / TAPS, t = e - ct * TAPS;     // division by a constant
      const int col = ct >> cshift, c = ct & (cpb - 1);
      if (c >= nci) continue;
      float v = 0.f;
      for (int g = 0; g < sgroups; ++g) v += tile[t][(c << sshift) + g][col];
      dw[(static_cast<size_t>(co0 + col) * Cin + ci0 + c) * TAPS + t] = v;
    }
  } else {
    // dw[ci][co0 .. co0+nco)[t]: nco*TAPS contiguous floats per input channel; thread <-> (c, col, t)
    for (int e = threadIdx.x; e < nci * 128 * TAPS; e += kThreads) {
      const int cc = e / TAPS, t = e - cc * TAPS;
      const int c = cc >> 7, col = cc & 127;
      if (col >= nco) continue;
      float v = 0.f;
      for (int g = 0; g < sgroups; ++g) v += tile[t][(c << sshift) + g][col];
      dw[(static_cast<size_t>(ci0 + c) * Cout + co0 + col) * TAPS + t] = v;
    }
  }
}

// One AdamW update (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments).
__device__ __forceinline__ void adamw_update(float& p, float g, float& m, float& v, float lr, float beta1, float beta2,
                                             float eps, float wd, float step_size, float bc2_sqrt, float grad_scale) {
  const float gi = g * grad_scale;
  float pi = p * (1.f - lr * wd);
  const float mi = beta1 * m + (1.f - beta1) * gi;
  const float vi = beta2 * v + (1.f - beta2) * gi * gi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= step_size * (mi / denom);
  p = pi; m = mi; v = vi;
}

// Four parameters per thread and trip (four 16-byte loads in flight per thread); the last n % 4 elements -- or all of
// them when a buffer is not 16-byte aligned (vec == false) -- go one by one.
__device__ __forceinline__ void adamw_span(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                           float* __restrict__ v, long long n, bool vec, float lr, float beta1,
                                           float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                           float grad_scale) {
  const float step_size = lr / bc1;
  const long long n4 = vec ? n >> 2 : 0;
  const long long tid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = tid; i < n4; i += stride) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    adamw_update(pp.x, gg.x, mm.x, vv.x, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale);
    adamw_update(pp.y, gg.y, mm.y, vv.y, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale);
    adamw_update(pp.z, gg.z, mm.z, vv.z, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale);
    adamw_update(pp.w, gg.w, mm.w, vv.w, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += stride)
    adamw_update(p[i], g[i], m[i], v[i], lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale);
}

__global__ void __launch_bounds__(kThreads)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, bool vec, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
             float grad_scale) {
  adamw_span(p, g, m, v, n, vec, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale);
}

// Same update with the hyper-parameters read from device memory, so that a CUDA graph of the whole training step can
// be replayed while lr and the bias corrections change per step. hyper = {lr, beta1, beta2, eps, wd, bc1, sqrt(bc2),
// grad_scale}.
__global__ void __launch_bounds__(kThreads)
adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, bool vec, const float* __restrict__ hyper) {
  adamw_span(p, g, m, v, n, vec, hyper[0], hyper[1], hyper[2], hyper[3], hyper[4], hyper[5], hyper[6], hyper[7]);
}

__global__ void __launch_bounds__(kThreads)
copy_channels_kernel(const __nv_bfloat16* __restrict__ src, int src_cs, __nv_bfloat16* __restrict__ dst, int dst_cs,
                     long long npix, int C) {
  const int groups = C / 8;
  const bool pow2 = (groups & (groups - 1)) == 0;
  const int gshift = __ffs(groups) - 1;
  const long long total = npix * groups;
  const long long tid = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * kThreads;
  constexpr int U = 4;   // independent 16-byte loads in flight per thread
  for (long long i0 = tid; i0 < total; i0 += stride * U) {
    uint4 raw[U];
    long long off[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      const long long ic = i < total ? i : i0;
      const long long pix = pow2 ? ic >> gshift : ic / groups;
      const int cg = static_cast<int>(ic - pix * groups);
      raw[u] = ldg16(src + pix * src_cs + cg * 8);
      off[u] = pix * dst_cs + cg * 8;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < total) stg16(dst + off[u], raw[u]);
  }
}

static int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  if (blocks > kEwBlocks) blocks = kEwBlocks;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}
static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
// channel counts the vectorised element-wise kernels accept: C/8 must divide 256
static bool ew_channels_ok(int C) { return C >= 8 && C % 8 == 0 && pow2(C / 8) && C / 8 <= kThreads; }

}  // namespace b2s

using namespace b2s;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" int b2s_ew_rows(void) { return kEwBlocks; }

extern "C" int b2s_reduce_rows(const float* in, int rows, int K, float* scratch, float* out, void* stream) {
  if (!in || !out || rows <= 0 || K <= 0) return set_error(B2S_ERR_ARG, "b2s_reduce_rows: bad argument");
  const float* src; int r;
  int rc = reduce_to_small(in, rows, K, scratch, &src, &r, STREAM(stream));
  if (rc) return rc;
  dim3 grid((K + 31) / 32, 1), block(32, kRedY);
  count_launch();
  reduce_rows_kernel<<<grid, block, 0, STREAM(stream)>>>(src, r, K, r, out);
  return check_launch("reduce_rows_kernel");
}

extern "C" int b2s_c1_rows(int N, int H, int W) {
  const long long npix = static_cast<long long>(N) * H * W;
  return grid_for(npix, 32 * 16);
}

static int c1_fwd_impl(const float* x, const float* w, const float* bias, const float* post_scale,
                       const float* post_shift, void* r, float* stats_partial, int N, int H, int W, int Cout, int flags,
                       void* stream) {
  if (!x || !w || !r) return set_error(B2S_ERR_ARG, "b2s_conv3x3_c1_fwd: null pointer");
  if (!ew_channels_ok(Cout) || Cout > 128) return set_error(B2S_ERR_ARG, "b2s_conv3x3_c1_fwd: unsupported Cout");
  if ((flags & B2S_FLAG_STATS) && !stats_partial) return set_error(B2S_ERR_ARG, "b2s_conv3x3_c1_fwd: stats missing");
  if (static_cast<long long>(N) * H * W >= (1ll << 31)) return set_error(B2S_ERR_ARG, "b2s_conv3x3_c1_fwd: N*H*W >= 2^31");
  const int grid = b2s_c1_rows(N, H, W);
  count_launch();
  auto kfn = (W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
                 ? conv3x3_c1_fwd4_kernel : conv3x3_c1_fwd_kernel;
  kfn<<<grid, kThreads, 0, STREAM(stream)>>>(x, w, bias, static_cast<__nv_bfloat16*>(r),
                                             (flags & B2S_FLAG_STATS) ? stats_partial : nullptr, N, H, W, Cout, flags,
                                             post_scale, post_shift);
  return check_launch("conv3x3_c1_fwd_kernel");
}

extern "C" int b2s_conv3x3_c1_fwd(const float* x, const float* w, const float* bias, void* r, float* stats_partial,
                                  int N, int H, int W, int Cout, int flags, void* stream) {
  return c1_fwd_impl(x, w, bias, nullptr, nullptr, r, stats_partial, N, H, W, Cout, flags, stream);
}

// inference variant: eval-mode BatchNorm affine applied after the ReLU (see b2s_conv_fwd_affine)
extern "C" int b2s_conv3x3_c1_fwd_affine(const float* x, const float* w, const float* bias, const float* post_scale,
                                         const float* post_shift, void* y, int N, int H, int W, int Cout, int flags,
                                         void* stream) {
  if (!post_scale || !post_shift) return set_error(B2S_ERR_ARG, "b2s_conv3x3_c1_fwd_affine: null pointer");
  return c1_fwd_impl(x, w, bias, post_scale, post_shift, y, nullptr, N, H, W, Cout, flags & ~B2S_FLAG_STATS, stream);
}

extern "C" int b2s_conv3x3_c1_wgrad(const float* x, const void* dz, float* partial, int N, int H, int W, int Cout,
                                    void* stream) {
  if (!x || !dz || !partial) return set_error(B2S_ERR_ARG, "b2s_conv3x3_c1_wgrad: null pointer");
  if (!ew_channels_ok(Cout) || Cout > 128) return set_error(B2S_ERR_ARG, "b2s_conv3x3_c1_wgrad: unsupported Cout");
  const int grid = b2s_c1_rows(N, H, W);
  count_launch();
  auto kfn = (W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
                 ? conv3x3_c1_wgrad4_kernel : conv3x3_c1_wgrad_kernel;
  kfn<<<grid, kThreads, 0, STREAM(stream)>>>(x, static_cast<const __nv_bfloat16*>(dz), partial, N, H, W, Cout);
  return check_launch("conv3x3_c1_wgrad_kernel");
}

extern "C" int b2s_bn_finalize(const float* partial, int rows, int C, double count, const float* gamma,
                               const float* beta, float* running_mean, float* running_var,
                               long long* num_batches_tracked, float momentum, float eps, float* scale, float* shift,
                               float* mean, float* invstd, float* scratch, void* stream) {
  if (!partial || !scale || !shift || !mean || !invstd) return set_error(B2S_ERR_ARG, "b2s_bn_finalize: null pointer");
  if (rows <= 0 || C <= 0 || count <= 0) return set_error(B2S_ERR_ARG, "b2s_bn_finalize: empty");
  if ((running_mean == nullptr) != (running_var == nullptr))
    return set_error(B2S_ERR_ARG, "b2s_bn_finalize: running_mean/var must both be given");
  const float* src; int r;
  int rc = reduce_to_small(partial, rows, 2 * C, scratch, &src, &r, STREAM(stream));
  if (rc) return rc;
  count_launch();
  bn_finalize_kernel<<<(C + 31) / 32, dim3(32, kRedY), 0, STREAM(stream)>>>(src, r, C, count, gamma, beta, running_mean,
                                                                  running_var, num_batches_tracked, momentum, eps,
                                                                  scale, shift, mean, invstd);
  return check_launch("bn_finalize_kernel");
}

extern "C" int b2s_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, float* scale, float* shift, int C,
                                  void* stream) {
  if (!running_mean || !running_var || !scale || !shift) return set_error(B2S_ERR_ARG, "b2s_bn_eval_affine: null");
  count_launch();
  bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, STREAM(stream)>>>(gamma, beta, running_mean, running_var, eps, scale,
                                                                     shift, C);
  return check_launch("bn_eval_affine_kernel");
}

extern "C" int b2s_bn_apply(const void* r, int r_cstride, const float* scale, const float* shift, void* y,
                            int y_cstride, void* pooled, int N, int H, int W, int C, void* stream) {
  if (!r || !scale || !shift || !y) return set_error(B2S_ERR_ARG, "b2s_bn_apply: null pointer");
  if (!ew_channels_ok(C)) return set_error(B2S_ERR_ARG, "b2s_bn_apply: C/8 must be a power of two <= 256");
  if (r_cstride % 8 || y_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_bn_apply: strides must be multiples of 8");
  if (static_cast<long long>(N) * H * W >= (1ll << 31)) return set_error(B2S_ERR_ARG, "b2s_bn_apply: N*H*W >= 2^31");
  const auto* rp = static_cast<const __nv_bfloat16*>(r);
  auto* yp = static_cast<__nv_bfloat16*>(y);
  count_launch();
  if (pooled) {
    if (H % 2 || W % 2) return set_error(B2S_ERR_ARG, "b2s_bn_apply: pooling needs even H and W");
    constexpr int smem = PrefetchRing<bn_apply_nv<true>(), bn_apply_depth<true>()>::kBytes;
    ew_allow_smem(bn_apply_kernel<true>, smem);
    const long long items = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
    static const int wave = ew_wave_blocks(bn_apply_kernel<true>, smem);
    bn_apply_kernel<true><<<ew_clamp_grid(wave, items, kThreads * 2), kThreads, smem,
                            STREAM(stream)>>>(rp, r_cstride, scale, shift, yp, y_cstride,
                                              static_cast<__nv_bfloat16*>(pooled), N, H, W, C);
  } else {
    constexpr int smem = PrefetchRing<bn_apply_nv<false>(), bn_apply_depth<false>()>::kBytes;
    ew_allow_smem(bn_apply_kernel<false>, smem);
    const long long items = static_cast<long long>(N) * H * W * (C / 8);
    static const int wave = ew_wave_blocks(bn_apply_kernel<false>, smem);
    bn_apply_kernel<false><<<ew_clamp_grid(wave, items, kThreads * 4), kThreads, smem,
                             STREAM(stream)>>>(rp, r_cstride, scale, shift, yp, y_cstride, nullptr, N, H, W, C);
  }
  return check_launch("bn_apply_kernel");
}

extern "C" int b2s_maxpool2x2(const void* x, int x_cstride, void* pooled, int N, int H, int W, int C, void* stream) {
  if (!x || !pooled) return set_error(B2S_ERR_ARG, "b2s_maxpool2x2: null pointer");
  // odd H / W: floor semantics of F.max_pool2d (the last row / column is not pooled)
  if (C % 8 || x_cstride % 8 || H < 2 || W < 2) return set_error(B2S_ERR_ARG, "b2s_maxpool2x2: need C % 8 == 0, H, W >= 2");
  const long long items = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  count_launch();
  maxpool2x2_kernel<<<grid_for(items, kThreads * 2), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cstride, static_cast<__nv_bfloat16*>(pooled), N, H, W, C);
  return check_launch("maxpool2x2_kernel");
}

extern "C" int b2s_bn_bwd_reduce(const void* dy, int dy_cstride, const void* dpool, const void* r, int r_cstride,
                                 const float* scale, const float* shift, const float* mean, const float* invstd,
                                 float* partial, int N, int H, int W, int C, void* stream) {
  if (!dy || !r || !mean || !invstd || !partial) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_reduce: null pointer");
  if (!ew_channels_ok(C)) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_reduce: unsupported C");
  if (dy_cstride % 8 || r_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_reduce: strides must be multiples of 8");
  if (dpool && (!scale || !shift || H % 2 || W % 2)) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_reduce: pool args");
  if (static_cast<long long>(N) * H * W >= (1ll << 31)) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_reduce: N*H*W >= 2^31");
  const auto* dyp = static_cast<const __nv_bfloat16*>(dy);
  const auto* rp = static_cast<const __nv_bfloat16*>(r);
  count_launch();
  // one full wave of blocks; the kernel zero-fills the partial rows beyond its grid
  constexpr int smem_pool = PrefetchRing<bn_bwd_nv<true>(), bn_bwd_depth<true, false>()>::kBytes;
  constexpr int smem_flat = PrefetchRing<bn_bwd_nv<false>(), bn_bwd_depth<false, false>()>::kBytes;
  if (dpool) {
    ew_allow_smem(bn_bwd_kernel<true, false>, smem_pool);
    static const int wave = ew_wave_blocks(bn_bwd_kernel<true, false>, smem_pool);
    bn_bwd_kernel<true, false><<<wave, kThreads, smem_pool,
                                 STREAM(stream)>>>(
        dyp, dy_cstride, static_cast<const __nv_bfloat16*>(dpool), rp, r_cstride, scale, shift, mean, invstd, nullptr,
        nullptr, 0, partial, N, H, W, C);
  } else {
    ew_allow_smem(bn_bwd_kernel<false, false>, smem_flat);
    static const int wave = ew_wave_blocks(bn_bwd_kernel<false, false>, smem_flat);
    bn_bwd_kernel<false, false><<<wave, kThreads, smem_flat,
                                  STREAM(stream)>>>(
        dyp, dy_cstride, nullptr, rp, r_cstride, scale, shift, mean, invstd, nullptr, nullptr, 0, partial, N, H, W, C);
  }
  return check_launch("bn_bwd_kernel<reduce>");
}

extern "C" int b2s_bn_bwd_finalize(const float* partial, int rows, int C, double count, const float* gamma,
                                   const float* invstd, float* dgamma, float* dbeta, float* coef, float* scratch,
                                   void* stream) {
  if (!partial || !invstd || !coef) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_finalize: null pointer");
  const float* src; int r;
  int rc = reduce_to_small(partial, rows, 2 * C, scratch, &src, &r, STREAM(stream));
  if (rc) return rc;
  count_launch();
  bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, kRedY), 0, STREAM(stream)>>>(src, r, C, count, gamma, invstd, dgamma, dbeta,
                                                                      coef);
  return check_launch("bn_bwd_finalize_kernel");
}

extern "C" int b2s_bn_bwd_finalize_raw(const float* partial, int rows, int C, double count, const float* gamma,
                                       const float* mean, const float* invstd, float* dgamma, float* dbeta, float* coef,
                                       float* scratch, void* stream) {
  if (!partial || !invstd || !mean || !coef) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_finalize_raw: null pointer");
  const float* src; int r;
  int rc = reduce_to_small(partial, rows, 2 * C, scratch, &src, &r, STREAM(stream));
  if (rc) return rc;
  count_launch();
  bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, kRedY), 0, STREAM(stream)>>>(src, r, C, count, gamma, invstd, dgamma, dbeta,
                                                                      coef, mean);
  return check_launch("bn_bwd_finalize_kernel<raw>");
}

extern "C" int b2s_bn_bwd_apply(const void* dy, int dy_cstride, const void* dpool, const void* r, int r_cstride,
                                const float* scale, const float* shift, const float* mean, const float* invstd,
                                const float* coef, void* dz, int dz_cstride, float* dbias_partial, int N, int H, int W,
                                int C, void* stream) {
  if (!dy || !r || !mean || !invstd || !coef || !dz || !dbias_partial)
    return set_error(B2S_ERR_ARG, "b2s_bn_bwd_apply: null pointer");
  if (!ew_channels_ok(C)) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_apply: unsupported C");
  if (dy_cstride % 8 || r_cstride % 8 || dz_cstride % 8)
    return set_error(B2S_ERR_ARG, "b2s_bn_bwd_apply: strides must be multiples of 8");
  if (dpool && (!scale || !shift || H % 2 || W % 2)) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_apply: pool args");
  if (static_cast<long long>(N) * H * W >= (1ll << 31)) return set_error(B2S_ERR_ARG, "b2s_bn_bwd_apply: N*H*W >= 2^31");
  const auto* dyp = static_cast<const __nv_bfloat16*>(dy);
  const auto* rp = static_cast<const __nv_bfloat16*>(r);
  auto* dzp = static_cast<__nv_bfloat16*>(dz);
  count_launch();
  constexpr int smem_pool = PrefetchRing<bn_bwd_nv<true>(), bn_bwd_depth<true, true>()>::kBytes;
  constexpr int smem_flat = PrefetchRing<bn_bwd_nv<false>(), bn_bwd_depth<false, true>()>::kBytes;
  if (dpool) {
    ew_allow_smem(bn_bwd_kernel<true, true>, smem_pool);
    static const int wave = ew_wave_blocks(bn_bwd_kernel<true, true>, smem_pool);
    bn_bwd_kernel<true, true><<<wave, kThreads, smem_pool,
                                STREAM(stream)>>>(
        dyp, dy_cstride, static_cast<const __nv_bfloat16*>(dpool), rp, r_cstride, scale, shift, mean, invstd, coef, dzp,
        dz_cstride, dbias_partial, N, H, W, C);
  } else {
    ew_allow_smem(bn_bwd_kernel<false, true>, smem_flat);
    static const int wave = ew_wave_blocks(bn_bwd_kernel<false, true>, smem_flat);
    bn_bwd_kernel<false, true><<<wave, kThreads, smem_flat,
                                 STREAM(stream)>>>(
        dyp, dy_cstride, nullptr, rp, r_cstride, scale, shift, mean, invstd, coef, dzp, dz_cstride, dbias_partial, N, H,
        W, C);
  }
  return check_launch("bn_bwd_kernel<apply>");
}

extern "C" int b2s_head_fwd(const void* r, int r_cstride, const float* scale, const float* shift, const float* w,
                            const float* b, float* logits, unsigned char* mask, int N, long long HW, int C, int O,
                            void* stream) {
  if (!r || !w || !logits) return set_error(B2S_ERR_ARG, "b2s_head_fwd: null pointer");
  if (C % 8 || !pow2(C / 8) || C / 8 > 32) return set_error(B2S_ERR_ARG, "b2s_head_fwd: C/8 must be a power of 2 <= 32");
  if (O < 1 || O > 64) return set_error(B2S_ERR_ARG, "b2s_head_fwd: unsupported out_channels");
  if (static_cast<long long>(N) * HW >= (1ll << 31)) return set_error(B2S_ERR_ARG, "b2s_head_fwd: N*H*W >= 2^31");
  const long long npix = static_cast<long long>(N) * HW;
  const int ppi = kThreads / (C / 8);
  count_launch();
  auto kfn = O == 1 ? head_fwd_kernel<true> : head_fwd_kernel<false>;
  kfn<<<grid_for(npix, ppi * 16), kThreads, (O * C + O) * sizeof(float), STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(r), r_cstride, scale, shift, w, b, logits, mask, N, HW, C, O);
  return check_launch("head_fwd_kernel");
}

extern "C" int b2s_head_bwd(const float* dlogits, const void* r, int r_cstride, const float* scale, const float* shift,
                            const float* w, void* dy, int dy_cstride, float* partial, int N, long long HW, int C, int O,
                            void* stream) {
  if (!dlogits || !r || !w || !dy || !partial) return set_error(B2S_ERR_ARG, "b2s_head_bwd: null pointer");
  if (C % 8 || !pow2(C / 8) || C / 8 > 32) return set_error(B2S_ERR_ARG, "b2s_head_bwd: C/8 must be a power of 2 <= 32");
  if (O < 1 || O > 64) return set_error(B2S_ERR_ARG, "b2s_head_bwd: unsupported out_channels");
  if (static_cast<long long>(N) * HW >= (1ll << 31)) return set_error(B2S_ERR_ARG, "b2s_head_bwd: N*H*W >= 2^31");
  count_launch();
  auto kfn = O == 1 ? head_bwd_kernel<true> : head_bwd_kernel<false>;
  kfn<<<kEwBlocks, kThreads, 0, STREAM(stream)>>>(dlogits, static_cast<const __nv_bfloat16*>(r), r_cstride, scale, shift, w,
                                                  static_cast<__nv_bfloat16*>(dy), dy_cstride, partial, N, HW, C, O);
  return check_launch("head_bwd_kernel");
}

extern "C" int b2s_loss_chunks(long long per_sample) {
  return static_cast<int>((per_sample + kLossChunk - 1) / kLossChunk);
}

extern "C" int b2s_seg_loss_fwd(const float* logits, const float* targets, int B, long long per_sample, float* partial,
                                float* sums, float* out, float dice_smooth, float w_bce, float w_dice, float w_ft,
                                float ft_alpha, float ft_beta, float ft_gamma, float ft_smooth, void* stream) {
  if (!logits || !targets || !partial || !sums || !out) return set_error(B2S_ERR_ARG, "b2s_seg_loss_fwd: null pointer");
  if (B <= 0 || per_sample <= 0 || B > 65535) return set_error(B2S_ERR_ARG, "b2s_seg_loss_fwd: bad batch");
  const int chunks = b2s_loss_chunks(per_sample);
  count_launch();
  seg_loss_partial_kernel<<<dim3(chunks, B), kThreads, 0, STREAM(stream)>>>(logits, targets, per_sample, chunks,
                                                                           partial);
  int rc = check_launch("seg_loss_partial_kernel");
  if (rc) return rc;
  count_launch();
  seg_loss_finalize_kernel<<<1, kThreads, 0, STREAM(stream)>>>(partial, B, chunks, per_sample, sums, out, dice_smooth,
                                                              w_bce, w_dice, w_ft, ft_alpha, ft_beta, ft_gamma,
                                                              ft_smooth);
  return check_launch("seg_loss_finalize_kernel");
}

extern "C" int b2s_seg_loss_bwd(const float* logits, const float* targets, const float* sums, const float* ft_tot,
                                int B, long long per_sample, long long bce_count, int dice_batch,
                                const float* grad_out, float* dlogits, float dice_smooth, float w_bce, float w_dice,
                                float w_ft, float ft_alpha, float ft_beta, float ft_gamma, float ft_smooth,
                                void* stream) {
  if (!logits || !targets || !sums || !dlogits) return set_error(B2S_ERR_ARG, "b2s_seg_loss_bwd: null pointer");
  if (B <= 0 || per_sample <= 0 || B > 65535) return set_error(B2S_ERR_ARG, "b2s_seg_loss_bwd: bad batch");
  int gx = static_cast<int>((per_sample + kThreads * 4 - 1) / (kThreads * 4));
  if (gx > 1024) gx = 1024;
  count_launch();
  seg_loss_bwd_kernel<<<dim3(gx, B), kThreads, 0, STREAM(stream)>>>(logits, targets, sums, ft_tot, B, per_sample,
                                                                   bce_count, dice_batch, grad_out, dlogits,
                                                                   dice_smooth, w_bce, w_dice, w_ft, ft_alpha, ft_beta,
                                                                   ft_gamma, ft_smooth);
  return check_launch("seg_loss_bwd_kernel");
}

extern "C" int b2s_metrics_blocks(long long n) { return static_cast<int>((n + kLossChunk - 1) / kLossChunk); }

extern "C" int b2s_seg_metrics(const float* logits, const float* targets, long long n, unsigned int* partial,
                               long long* counters, void* stream) {
  if (!logits || !targets || !partial || !counters) return set_error(B2S_ERR_ARG, "b2s_seg_metrics: null pointer");
  if (n <= 0) return set_error(B2S_ERR_ARG, "b2s_seg_metrics: empty input");
  const int blocks = b2s_metrics_blocks(n);
  count_launch();
  seg_metrics_partial_kernel<<<blocks, kThreads, 0, STREAM(stream)>>>(logits, targets, n, partial);
  int rc = check_launch("seg_metrics_partial_kernel");
  if (rc) return rc;
  count_launch();
  seg_metrics_accumulate_kernel<<<1, kThreads, 0, STREAM(stream)>>>(partial, blocks, n, counters);
  return check_launch("seg_metrics_accumulate_kernel");
}

extern "C" int b2s_pack_conv_weight(const float* w, void* w_fwd, void* w_dgrad, int Cout, int Cin, int ksize,
                                    void* stream) {
  if (!w || (!w_fwd && !w_dgrad)) return set_error(B2S_ERR_ARG, "b2s_pack_conv_weight: null pointer");
  if (ksize != 1 && ksize != 3) return set_error(B2S_ERR_ARG, "b2s_pack_conv_weight: ksize must be 1 or 3");
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32);
  count_launch();
  pack_conv_weight_kernel<<<grid, kThreads, 0, STREAM(stream)>>>(w, static_cast<__nv_bfloat16*>(w_fwd),
                                                           static_cast<__nv_bfloat16*>(w_dgrad), Cout, Cin,
                                                           ksize * ksize);
  return check_launch("pack_conv_weight_kernel");
}

extern "C" int b2s_pack_convt_weight(const float* w, void* w_fwd, void* w_dgrad, int Cin, int Cout, void* stream) {
  if (!w || (!w_fwd && !w_dgrad)) return set_error(B2S_ERR_ARG, "b2s_pack_convt_weight: null pointer");
  const long long total = static_cast<long long>(Cin) * Cout * 4;
  count_launch();
  pack_convt_weight_kernel<<<grid_for(total, kThreads * 4), kThreads, 0, STREAM(stream)>>>(
      w, static_cast<__nv_bfloat16*>(w_fwd), static_cast<__nv_bfloat16*>(w_dgrad), Cin, Cout);
  return check_launch("pack_convt_weight_kernel");
}

// n tensors described by parallel arrays (host memory): kind[i] 0 = Conv2d weight [Cout][Cin][k][k] (d0 = Cout, d1 = Cin,
// taps = k*k), 1 = ConvTranspose2d weight [Cin][Cout][2][2] (d0 = Cin, d1 = Cout). wf / wd entries may be NULL.
extern "C" int b2s_pack_weights_all(int n, const void* const* w, void* const* wf, void* const* wd, const int* d0,
                                    const int* d1, const int* taps, const int* kind, void* stream) {
  if (n < 1 || n > kPackMaxTensors || !w || !wf || !wd || !d0 || !d1 || !taps || !kind)
    return set_error(B2S_ERR_ARG, "b2s_pack_weights_all: bad argument (1..40 tensors)");
  PackTable tab;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    if (!w[i] || (kind[i] != 0 && kind[i] != 1) || (kind[i] == 0 && taps[i] != 1 && taps[i] != 9) ||
        (kind[i] == 1 && taps[i] != 4) || d0[i] < 1 || d1[i] < 1)
      return set_error(B2S_ERR_ARG, "b2s_pack_weights_all: bad tensor description");
    tab.e[i] = PackEntry{static_cast<const float*>(w[i]), static_cast<__nv_bfloat16*>(wf[i]),
                         static_cast<__nv_bfloat16*>(wd[i]), d0[i], d1[i], taps[i], kind[i], blocks};
    blocks += ((d0[i] + 31) / 32) * ((d1[i] + 31) / 32);
  }
  tab.n = n;
  count_launch();
  pack_all_kernel<<<blocks, kThreads, 0, STREAM(stream)>>>(tab);
  return check_launch("pack_all_kernel");
}

extern "C" int b2s_wgrad_reduce(const float* ws, int splits, int taps, int Cin, int Cout, float* dw, int layout,
                                void* stream) {
  if (!ws || !dw) return set_error(B2S_ERR_ARG, "b2s_wgrad_reduce: null pointer");
  if (taps < 1 || taps > 9 || splits < 1) return set_error(B2S_ERR_ARG, "b2s_wgrad_reduce: bad taps/splits");
  if (Cout % 4) return set_error(B2S_ERR_ARG, "b2s_wgrad_reduce: Cout must be a multiple of 4");
  const int sgroups = splits >= 8 ? 8 : splits >= 4 ? 4 : splits >= 2 ? 2 : 1;
  const int cpb = 8 / sgroups;
  dim3 grid((Cin + cpb - 1) / cpb, (Cout + 127) / 128);
  count_launch();
  switch (taps) {
    case 9: wgrad_reduce_kernel<9><<<grid, kThreads, 0, STREAM(stream)>>>(ws, splits, Cin, Cout, dw, layout, sgroups); break;
    case 4: wgrad_reduce_kernel<4><<<grid, kThreads, 0, STREAM(stream)>>>(ws, splits, Cin, Cout, dw, layout, sgroups); break;
    case 1: wgrad_reduce_kernel<1><<<grid, kThreads, 0, STREAM(stream)>>>(ws, splits, Cin, Cout, dw, layout, sgroups); break;
    default: return set_error(B2S_ERR_ARG, "b2s_wgrad_reduce: taps must be 9 (3x3), 4 (transposed 2x2) or 1 (1x1)");
  }
  return check_launch("wgrad_reduce_kernel");
}

static bool adamw_aligned(const float* p, const float* g, const float* m, const float* v) {
  return ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
           reinterpret_cast<uintptr_t>(v)) & 15) == 0;
}

extern "C" int b2s_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  if (!p || !g || !m || !v) return set_error(B2S_ERR_ARG, "b2s_adamw_step: null pointer");
  if (step < 1) return set_error(B2S_ERR_ARG, "b2s_adamw_step: step is 1-based");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  count_launch();
  adamw_kernel<<<grid_for(n, kThreads * 8), kThreads, 0, STREAM(stream)>>>(
      p, g, m, v, n, adamw_aligned(p, g, m, v), lr, beta1, beta2, eps, weight_decay, static_cast<float>(bc1), static_cast<float>(sqrt(bc2)),
      grad_scale);
  return check_launch("adamw_kernel");
}

extern "C" int b2s_adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper,
                                  void* stream) {
  if (!p || !g || !m || !v || !hyper) return set_error(B2S_ERR_ARG, "b2s_adamw_step_dev: null pointer");
  count_launch();
  adamw_dev_kernel<<<grid_for(n, kThreads * 8), kThreads, 0, STREAM(stream)>>>(p, g, m, v, n,
                                                                              adamw_aligned(p, g, m, v), hyper);
  return check_launch("adamw_dev_kernel");
}

extern "C" int b2s_copy_channels(const void* src, int src_cstride, void* dst, int dst_cstride, long long npix, int C,
                                 void* stream) {
  if (!src || !dst) return set_error(B2S_ERR_ARG, "b2s_copy_channels: null pointer");
  if (C % 8 || src_cstride % 8 || dst_cstride % 8) return set_error(B2S_ERR_ARG, "b2s_copy_channels: need multiples of 8");
  count_launch();
  copy_channels_kernel<<<grid_for(npix * (C / 8), kThreads * 4), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), src_cstride, static_cast<__nv_bfloat16*>(dst), dst_cstride, npix, C);
  return check_launch("copy_channels_kernel");
}
