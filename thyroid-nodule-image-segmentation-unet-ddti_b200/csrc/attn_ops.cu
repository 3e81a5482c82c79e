// Bandwidth kernels of the models/mod.py attention-gate variant and of multi-channel inputs (sm_100a).
//
// Reference call sites:
//   AttentionGate.forward            models/mod.py:229-234   psi = sigmoid(BN1(conv1x1(relu(g1 + x1)))); return x * psi
//     - the one-channel BatchNorm + sigmoid over the fp32 map of the F_int -> 1 conv: psi_stats / psi_fwd /
//       psi_bwd_reduce / psi_bwd_apply (two-stage deterministic reductions; finalised by b2s_bn_finalize /
//       b2s_bn_bwd_finalize with C = 1)
//     - x * psi (psi broadcast over channels) and its backward (dx = dy * psi, dpsi = sum_c dy * x): pixel_scale
//   F.interpolate(x, size, mode='bilinear', align_corners=False)   models/mod.py:61-62,126-127,289-290 (odd sizes)
//   first conv of a net built with in_channels > 1   models/model.py:10, models/mod.py:25: the fp32 NCHW image is
//       converted once to NHWC bf16, zero-padded to 64 channels, and then takes the tensor-core conv path.
#include "ew_common.cuh"

namespace b2s {

// ---- image [N,C,H,W] fp32 -> [N,H,W,Cpad] bf16 (channels >= C are zero) --------------------------------------------
__global__ void __launch_bounds__(kThreads)
image_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, long long HW, long long npix,
                     int Cpad) {
  const int groups = Cpad / 8;
  const long long items = npix * groups;
  for (long long it = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * kThreads) {
    const long long p = it / groups;
    const int g = static_cast<int>(it - p * groups);
    const long long n = p / HW, hw = p - n * HW;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      v[k] = c < C ? __ldg(x + (n * C + c) * HW + hw) : 0.f;
    }
    stg16(y + p * Cpad + g * 8, pack8(v));
  }
}

// ---- one-channel BatchNorm + sigmoid over the sum of up to four fp32 maps -------------------------------------------
struct PsiMaps {
  const float* m[4];
  int count;
};
__device__ __forceinline__ float psi_in(const PsiMaps& a, long long i) {
  float v = __ldg(a.m[0] + i);
  for (int k = 1; k < a.count; ++k) v += __ldg(a.m[k] + i);
  return v;
}

// deterministic block reduction of two values; result valid in thread 0
__device__ __forceinline__ void block_reduce2(float& a, float& b) {
  __shared__ float red[2][kThreads / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = a; red[1][w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.f; b = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) { a += red[0][i]; b += red[1][i]; }
  }
}

// partial [gridDim.x][2] = {sum v, sum v^2}
__global__ void __launch_bounds__(kThreads)
psi_stats_kernel(PsiMaps a, long long n, float* __restrict__ partial) {
  float s = 0.f, q = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float v = psi_in(a, i);
    s += v;
    q = fmaf(v, v, q);
  }
  block_reduce2(s, q);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s; partial[2 * blockIdx.x + 1] = q; }
}

__global__ void __launch_bounds__(kThreads)
psi_fwd_kernel(PsiMaps a, const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ psi,
               long long n) {
  const float sc = scale[0], sh = shift[0];
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float z = fmaf(psi_in(a, i), sc, sh);
    psi[i] = 1.f / (1.f + expf(-z));
  }
}

// g = dpsi * psi * (1 - psi) is the gradient at the BatchNorm output; partial [gridDim.x][2] = {sum g, sum g * xhat}
__global__ void __launch_bounds__(kThreads)
psi_bwd_reduce_kernel(PsiMaps a, const float* __restrict__ psi, const float* __restrict__ dpsi,
                      const float* __restrict__ mean, const float* __restrict__ invstd, long long n,
                      float* __restrict__ partial) {
  const float mu = mean[0], is = invstd[0];
  float s = 0.f, q = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float p = __ldg(psi + i);
    const float g = __ldg(dpsi + i) * p * (1.f - p);
    const float xh = (psi_in(a, i) - mu) * is;
    s += g;
    q = fmaf(g, xh, q);
  }
  block_reduce2(s, q);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s; partial[2 * blockIdx.x + 1] = q; }
}

// dv = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)); coef = {gamma * invstd, mean(g), mean(g * xhat)}
__global__ void __launch_bounds__(kThreads)
psi_bwd_apply_kernel(PsiMaps a, const float* __restrict__ psi, const float* __restrict__ dpsi,
                     const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ coef,
                     float* __restrict__ dv, long long n) {
  const float mu = mean[0], is = invstd[0], c0 = coef[0], c1 = coef[1], c2 = coef[2];
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float p = __ldg(psi + i);
    const float g = __ldg(dpsi + i) * p * (1.f - p);
    const float xh = (psi_in(a, i) - mu) * is;
    dv[i] = c0 * (g - c1 - xh * c2);
  }
}

// ---- out[p, c] = x[p, c] * psi[p]; backward: dx = dy * psi, dpsi[p] = sum_c dy[p, c] * x[p, c] -----------------------
// L = min(C / 8, 32) lanes cooperate on one pixel; lane l owns channels (j * L + l) * 8 .. + 7 for j < C / (8 L).
template <bool BWD>
__global__ void __launch_bounds__(kThreads)
pixel_scale_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, const float* __restrict__ psi,
                   const __nv_bfloat16* __restrict__ dy, int dy_cs, __nv_bfloat16* __restrict__ out, int out_cs,
                   float* __restrict__ dpsi, long long npix, int C) {
  const int L = (C / 8) < 32 ? (C / 8) : 32;
  const int chunks = C / (8 * L);
  const int lane = threadIdx.x % L;
  const int ppb = kThreads / L;
  for (long long p = static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / L; ; p += static_cast<long long>(gridDim.x) * ppb) {
    // all lanes of a warp leave together: npix is tested on the warp's first pixel (L divides 32)
    const long long p_first = p - (threadIdx.x % 32) / L;
    if (p_first >= npix) break;
    const bool on = p < npix;
    const float s = on ? __ldg(psi + p) : 0.f;
    float acc = 0.f;
    for (int j = 0; j < chunks; ++j) {
      const int c = (j * L + lane) * 8;
      if (on) {
        float xv[8], o[8];
        unpack8(ldg16(x + p * x_cs + c), xv);
        if (BWD) {
          float dv[8];
          unpack8(ldg16(dy + p * dy_cs + c), dv);
#pragma unroll
          for (int k = 0; k < 8; ++k) { o[k] = dv[k] * s; acc = fmaf(dv[k], xv[k], acc); }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = xv[k] * s;
        }
        stg16(out + p * out_cs + c, pack8(o));
      }
    }
    if (BWD) {
      for (int off = L >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (on && lane == 0) dpsi[p] = acc;
    }
  }
}

// ---- bilinear resize (align_corners = False) of NHWC bf16 -----------------------------------------------------------
// Source coordinate of output index o: max(0, (o + 0.5) * in / out - 0.5); taps i0 = floor, i1 = min(i0 + 1, in - 1).
__device__ __forceinline__ void bilinear_taps(int o, int in, int out, int& i0, int& i1, float& w1) {
  float src = (static_cast<float>(o) + 0.5f) * (static_cast<float>(in) / static_cast<float>(out)) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = static_cast<int>(src);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + 1 < in ? i0 + 1 : in - 1;
  w1 = src - static_cast<float>(i0);
}

__global__ void __launch_bounds__(kThreads)
bilinear_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, __nv_bfloat16* __restrict__ y, int y_cs, int N,
                    int Hi, int Wi, int Ho, int Wo, int C) {
  const int groups = C / 8;
  const long long items = static_cast<long long>(N) * Ho * Wo * groups;
  for (long long it = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * kThreads) {
    const int g = static_cast<int>(it % groups);
    const long long p = it / groups;
    const int wo = static_cast<int>(p % Wo);
    const int ho = static_cast<int>((p / Wo) % Ho);
    const long long n = p / (static_cast<long long>(Wo) * Ho);
    int h0, h1, w0, w1;
    float ah, aw;
    bilinear_taps(ho, Hi, Ho, h0, h1, ah);
    bilinear_taps(wo, Wi, Wo, w0, w1, aw);
    const __nv_bfloat16* base = x + n * Hi * Wi * x_cs + g * 8;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(ldg16(base + (static_cast<long long>(h0) * Wi + w0) * x_cs), a);
    unpack8(ldg16(base + (static_cast<long long>(h0) * Wi + w1) * x_cs), b);
    unpack8(ldg16(base + (static_cast<long long>(h1) * Wi + w0) * x_cs), c);
    unpack8(ldg16(base + (static_cast<long long>(h1) * Wi + w1) * x_cs), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float top = a[k] + aw * (b[k] - a[k]);
      const float bot = c[k] + aw * (d[k] - c[k]);
      o[k] = top + ah * (bot - top);
    }
    stg16(y + p * y_cs + g * 8, pack8(o));
  }
}

// Backward as a gather (deterministic): input pixel (hi, wi) collects from the output rows / columns whose taps touch
// it. For an up-scaling resize those are at most a few; the candidate range is bounded by the inverse map +- 2.
__device__ __forceinline__ float tap_weight(int o, int i, int in, int out) {
  int i0, i1;
  float w1;
  bilinear_taps(o, in, out, i0, i1, w1);
  float w = 0.f;
  if (i0 == i) w += 1.f - w1;
  if (i1 == i) w += w1;
  return w;
}

__global__ void __launch_bounds__(kThreads)
bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_cs, __nv_bfloat16* __restrict__ dx, int dx_cs, int N,
                    int Hi, int Wi, int Ho, int Wo, int C) {
  const int groups = C / 8;
  const long long items = static_cast<long long>(N) * Hi * Wi * groups;
  const float rh = static_cast<float>(Ho) / static_cast<float>(Hi), rw = static_cast<float>(Wo) / static_cast<float>(Wi);
  for (long long it = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * kThreads) {
    const int g = static_cast<int>(it % groups);
    const long long p = it / groups;
    const int wi = static_cast<int>(p % Wi);
    const int hi = static_cast<int>((p / Wi) % Hi);
    const long long n = p / (static_cast<long long>(Wi) * Hi);
    // outputs o with a tap on input i satisfy (i - 1) < src(o) < (i + 1)  =>  o in ((i - 0.5) r - 0.5, (i + 1.5) r - 0.5)
    int ho_lo = static_cast<int>(floorf((hi - 0.5f) * rh - 0.5f)) - 1, ho_hi = static_cast<int>(ceilf((hi + 1.5f) * rh - 0.5f)) + 1;
    int wo_lo = static_cast<int>(floorf((wi - 0.5f) * rw - 0.5f)) - 1, wo_hi = static_cast<int>(ceilf((wi + 1.5f) * rw - 0.5f)) + 1;
    if (ho_lo < 0) ho_lo = 0;
    if (wo_lo < 0) wo_lo = 0;
    if (ho_hi > Ho - 1) ho_hi = Ho - 1;
    if (wo_hi > Wo - 1) wo_hi = Wo - 1;
    if (hi == 0) ho_lo = 0;                    // clamped sources (src < 0) all land on row / column 0
    if (wi == 0) wo_lo = 0;
    if (hi == Hi - 1) ho_hi = Ho - 1;          // and i1 is clamped to the last row / column
    if (wi == Wi - 1) wo_hi = Wo - 1;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const float wh = tap_weight(ho, hi, Hi, Ho);
      if (wh == 0.f) continue;
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const float ww = tap_weight(wo, wi, Wi, Wo);
        if (ww == 0.f) continue;
        float v[8];
        unpack8(ldg16(dy + ((n * Ho + ho) * Wo + wo) * dy_cs + g * 8), v);
        const float w = wh * ww;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, v[k], acc[k]);
      }
    }
    stg16(dx + p * dx_cs + g * 8, pack8(acc));
  }
}

static int maps_from(const float* const* maps, int n_maps, PsiMaps* out) {
  if (!maps || n_maps < 1 || n_maps > 4) return set_error(B2S_ERR_ARG, "psi: 1..4 input maps");
  out->count = n_maps;
  for (int i = 0; i < 4; ++i) out->m[i] = i < n_maps ? maps[i] : nullptr;
  for (int i = 0; i < n_maps; ++i)
    if (!maps[i]) return set_error(B2S_ERR_ARG, "psi: null map");
  return B2S_OK;
}

}  // namespace b2s

using namespace b2s;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" int b2s_image_to_nhwc(const float* x, void* y, int N, int C, long long HW, int Cpad, void* stream) {
  if (!x || !y) return set_error(B2S_ERR_ARG, "b2s_image_to_nhwc: null pointer");
  if (C < 1 || Cpad < C || Cpad % 8) return set_error(B2S_ERR_ARG, "b2s_image_to_nhwc: need 1 <= C <= Cpad, Cpad % 8 == 0");
  const long long npix = static_cast<long long>(N) * HW;
  count_launch();
  image_to_nhwc_kernel<<<ew_grid_for(npix * (Cpad / 8), kThreads * 4), kThreads, 0, STREAM(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), C, HW, npix, Cpad);
  return check_launch("image_to_nhwc_kernel");
}

// rows of the [rows][2] partial buffers of b2s_psi_stats / b2s_psi_bwd_reduce (reduce with b2s_bn_finalize /
// b2s_bn_bwd_finalize, C = 1)
extern "C" int b2s_psi_rows(long long n) { return ew_grid_for(n, kThreads * 8); }

extern "C" int b2s_psi_stats(const float* const* maps, int n_maps, long long n, float* partial, void* stream) {
  PsiMaps a;
  if (int rc = maps_from(maps, n_maps, &a)) return rc;
  if (!partial || n <= 0) return set_error(B2S_ERR_ARG, "b2s_psi_stats: bad argument");
  count_launch();
  psi_stats_kernel<<<b2s_psi_rows(n), kThreads, 0, STREAM(stream)>>>(a, n, partial);
  return check_launch("psi_stats_kernel");
}

extern "C" int b2s_psi_fwd(const float* const* maps, int n_maps, const float* scale, const float* shift, float* psi,
                           long long n, void* stream) {
  PsiMaps a;
  if (int rc = maps_from(maps, n_maps, &a)) return rc;
  if (!scale || !shift || !psi || n <= 0) return set_error(B2S_ERR_ARG, "b2s_psi_fwd: bad argument");
  count_launch();
  psi_fwd_kernel<<<ew_grid_for(n, kThreads * 4), kThreads, 0, STREAM(stream)>>>(a, scale, shift, psi, n);
  return check_launch("psi_fwd_kernel");
}

extern "C" int b2s_psi_bwd_reduce(const float* const* maps, int n_maps, const float* psi, const float* dpsi,
                                  const float* mean, const float* invstd, long long n, float* partial, void* stream) {
  PsiMaps a;
  if (int rc = maps_from(maps, n_maps, &a)) return rc;
  if (!psi || !dpsi || !mean || !invstd || !partial || n <= 0) return set_error(B2S_ERR_ARG, "b2s_psi_bwd_reduce: bad argument");
  count_launch();
  psi_bwd_reduce_kernel<<<b2s_psi_rows(n), kThreads, 0, STREAM(stream)>>>(a, psi, dpsi, mean, invstd, n, partial);
  return check_launch("psi_bwd_reduce_kernel");
}

extern "C" int b2s_psi_bwd_apply(const float* const* maps, int n_maps, const float* psi, const float* dpsi,
                                 const float* mean, const float* invstd, const float* coef, float* dv, long long n,
                                 void* stream) {
  PsiMaps a;
  if (int rc = maps_from(maps, n_maps, &a)) return rc;
  if (!psi || !dpsi || !mean || !invstd || !coef || !dv || n <= 0) return set_error(B2S_ERR_ARG, "b2s_psi_bwd_apply: bad argument");
  count_launch();
  psi_bwd_apply_kernel<<<ew_grid_for(n, kThreads * 4), kThreads, 0, STREAM(stream)>>>(a, psi, dpsi, mean, invstd, coef, dv, n);
  return check_launch("psi_bwd_apply_kernel");
}

static bool pixel_scale_channels_ok(int C) { return C >= 64 && C % 64 == 0 && ew_pow2(C / 8) && C <= 2048; }

extern "C" int b2s_pixel_scale_fwd(const void* x, int x_cstride, const float* psi, void* out, int out_cstride,
                                   long long npix, int C, void* stream) {
  if (!x || !psi || !out) return set_error(B2S_ERR_ARG, "b2s_pixel_scale_fwd: null pointer");
  if (!pixel_scale_channels_ok(C) || x_cstride % 8 || out_cstride % 8)
    return set_error(B2S_ERR_ARG, "b2s_pixel_scale_fwd: C must be 64 * 2^k, strides multiples of 8");
  const int L = (C / 8) < 32 ? (C / 8) : 32;
  count_launch();
  pixel_scale_kernel<false><<<ew_grid_for(npix, kThreads / L), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cstride, psi, nullptr, 0, static_cast<__nv_bfloat16*>(out), out_cstride,
      nullptr, npix, C);
  return check_launch("pixel_scale_kernel<fwd>");
}

extern "C" int b2s_pixel_scale_bwd(const void* x, int x_cstride, const float* psi, const void* dy, int dy_cstride,
                                   void* dx, int dx_cstride, float* dpsi, long long npix, int C, void* stream) {
  if (!x || !psi || !dy || !dx || !dpsi) return set_error(B2S_ERR_ARG, "b2s_pixel_scale_bwd: null pointer");
  if (!pixel_scale_channels_ok(C) || x_cstride % 8 || dy_cstride % 8 || dx_cstride % 8)
    return set_error(B2S_ERR_ARG, "b2s_pixel_scale_bwd: C must be 64 * 2^k, strides multiples of 8");
  const int L = (C / 8) < 32 ? (C / 8) : 32;
  count_launch();
  pixel_scale_kernel<true><<<ew_grid_for(npix, kThreads / L), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cstride, psi, static_cast<const __nv_bfloat16*>(dy), dy_cstride,
      static_cast<__nv_bfloat16*>(dx), dx_cstride, dpsi, npix, C);
  return check_launch("pixel_scale_kernel<bwd>");
}

extern "C" int b2s_bilinear_fwd(const void* x, int x_cstride, void* y, int y_cstride, int N, int Hi, int Wi, int Ho,
                                int Wo, int C, void* stream) {
  if (!x || !y) return set_error(B2S_ERR_ARG, "b2s_bilinear_fwd: null pointer");
  if (C % 8 || x_cstride % 8 || y_cstride % 8 || N < 1 || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1)
    return set_error(B2S_ERR_ARG, "b2s_bilinear_fwd: bad shape");
  count_launch();
  bilinear_fwd_kernel<<<ew_grid_for(static_cast<long long>(N) * Ho * Wo * (C / 8), kThreads * 2), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cstride, static_cast<__nv_bfloat16*>(y), y_cstride, N, Hi, Wi, Ho, Wo, C);
  return check_launch("bilinear_fwd_kernel");
}

extern "C" int b2s_bilinear_bwd(const void* dy, int dy_cstride, void* dx, int dx_cstride, int N, int Hi, int Wi, int Ho,
                                int Wo, int C, void* stream) {
  if (!dy || !dx) return set_error(B2S_ERR_ARG, "b2s_bilinear_bwd: null pointer");
  if (C % 8 || dy_cstride % 8 || dx_cstride % 8 || N < 1 || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1)
    return set_error(B2S_ERR_ARG, "b2s_bilinear_bwd: bad shape");
  count_launch();
  bilinear_bwd_kernel<<<ew_grid_for(static_cast<long long>(N) * Hi * Wi * (C / 8), kThreads * 2), kThreads, 0, STREAM(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), dy_cstride, static_cast<__nv_bfloat16*>(dx), dx_cstride, N, Hi, Wi, Ho, Wo, C);
  return check_launch("bilinear_bwd_kernel");
}
