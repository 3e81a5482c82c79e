// Host-side helpers shared by the tcgen05 conv translation units: TMA tensor-map encoding, tile-box selection,
// persistent-grid sizing.
#pragma once
#include "ptx.cuh"
#include "b2s_internal.h"

namespace b2s {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB

enum AMode : int { A_CONV3 = 0, A_1X1 = 1, A_CONVT_DGRAD = 2, A_CONV3_S2 = 3, A_TAPLIST = 4 };
enum OutMode : int { OUT_4D = 0, OUT_CONVT_5D = 1, OUT_SUB_5D = 2 };

struct ConvTcParams {
  int bw, bh, bn;                  // pixel box of one M tile, bw*bh*bn == 128
  int tiles_w, tiles_h, tiles_n;   // tiles over (W, H, N)
  int tiles_m;                     // tiles_w * tiles_h * tiles_n
  int W, H, N;                     // pixel space of the GEMM rows
  int num_taps, k_chunks;          // K loop = taps x (Cin/64)
  int a_mode, out_mode;
  int n_total;                     // GEMM N (= Cout; 4*Cout for convT forward)
  int cout_sub;                    // convT forward: Cout per (a,b) sub-position; else n_total
  int tiles_nn;                    // n_total / BLOCK_N
  int flags;                       // B2S_FLAG_*
  int tap_dh[4], tap_dw[4], tap_w[4];  // A_TAPLIST: pixel offsets of tap t and its row block in the packed weights
  int sub_a, sub_b;                // OUT_SUB_5D: sub-lattice (row, column parity) of the 2x up-sampled output
  int a_cstride;                   // A_CONV3_S2: pixel stride (elements) of the input, the offset of the odd-column parity
  const float* bias;               // [cout_sub] or nullptr
  const float* post_scale;         // eval-mode BatchNorm folded into the epilogue: y = act(acc + bias) * post_scale +
  const float* post_shift;         // post_shift per output channel (both nullptr in training)
  int items_m;                     // work items of the tile-pair / halo kernels (see conv_plan)
  float* stats;                    // [2 * gridDim.x / tiles_nn][2][n_total] partial column sums, or nullptr
  // B2S_FLAG_BNRED (input-gradient launches whose output dy feeds a train-mode BatchNorm backward): the second partial
  // row holds sum(dy * r) instead of sum(dy^2); r = that BatchNorm's saved input, same pixel space and channels as dy
  const __nv_bfloat16* red_r;
  int red_cs;                      // pixel stride of r in elements
};

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// bf16 tensor map, SWIZZLE_128B, inner box = 64 elements. dims/strides innermost first; strides in BYTES for
// dims 1..rank-1.
static inline int make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                     const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(B2S_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_b[i];
  if (reinterpret_cast<uintptr_t>(base) & 15) return set_error(B2S_ERR_ARG, "tensor base not 16-B aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] & 15) return set_error(B2S_ERR_ARG, "tensor stride not a multiple of 16 B");
  // L2 promotion: a 128-byte box row (64 channels) is promoted to a 256-byte DRAM fetch only when the neighbouring
  // 128 bytes belong to the same tensor view (the view covers whole pixel records, so another k-chunk / tap of the same
  // kernel reads them next). For a channel SLICE of a wider buffer (the up-conv half of a concat gradient: 128 B of every
  // 256-B record) the promoted half is never used: ncu showed 2.0x the algorithmic DRAM reads in the transposed-conv
  // dgrad / wgrad (1.05 GB for 537 MB).
  const bool whole_records = rank < 2 || dims[0] * 2 == strides_b[0];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  whole_records ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu %llu box %u %u %u %u)",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1],
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bdim[0],
             bdim[1], rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0);
    return set_error(B2S_ERR_CUDA, msg);
  }
  return B2S_OK;
}

static inline int pow2_ceil(int x) { int p = 1; while (p < x) p *= 2; return p; }

// Split `total` (power of two) pixels of one tile over (w, h, n).
static inline void pick_box(int W, int H, int N, int total, int* bw, int* bh, int* bn) {
  int w = pow2_ceil(W); if (w > total) w = total;
  int h = pow2_ceil(H); if (h > total / w) h = total / w;
  int n = total / (w * h);
  (void)N;
  *bw = w; *bh = h; *bn = n;
}

// NHWC activation map (C, W, H, N) over a channel slice of a buffer whose pixel stride is cstride elements.
static inline int make_act_map4(CUtensorMap* m, const void* base, int C, int W, int H, int N, int cstride, int bw, int bh,
                         int bn) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)cstride * 2, (uint64_t)W * cstride * 2, (uint64_t)H * W * cstride * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  return make_tmap(m, base, 4, dims, str, box);
}
// Stride-2 sampling without element strides (a box with elementStrides = 2 makes TMA fetch the skipped pixels as
// well: 182 TFLOP/s for the 64 -> 128 conv at 512^2). The NHWC tensor is viewed as (pc, w2, ph, h2, n) with
// w = 2*w2 + pw, h = 2*h2 + ph and pc = pw * cstride + c: a pixel pair is one row of cstride + C elements, so the
// parity pw is an offset in the innermost coordinate. One parity class of an image is then a dense box
// {64, bw, 1, bh, bn}; w2 = -1 / h2 = -1 (the conv's zero padding) are out of bounds of their own dimensions.
static inline int make_act_map5_parity(CUtensorMap* m, const void* base, int C, int W, int H, int N, int cstride, int bw,
                                       int bh, int bn) {
  uint64_t dims[5] = {(uint64_t)cstride + C, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)N};
  uint64_t str[4] = {2ull * cstride * 2, (uint64_t)W * cstride * 2, 2ull * W * cstride * 2,
                     (uint64_t)H * W * cstride * 2};
  uint32_t box[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bn};
  return make_tmap(m, base, 5, dims, str, box);
}
// 2x-upsampled NHWC tensor [N, 2Hi, 2Wi, C] viewed as (C, b, j, a, i*N) so that one (a,b) sub-lattice is a box.
static inline int make_up_map5(CUtensorMap* m, const void* base, int C, int Wi, int Hi, int N, int cstride, int bw,
                        int bhn) {
  const uint64_t Wo = 2ull * Wi;
  uint64_t dims[5] = {(uint64_t)C, 2, (uint64_t)Wi, 2, (uint64_t)Hi * N};
  uint64_t str[4] = {(uint64_t)cstride * 2, 2ull * cstride * 2, Wo * cstride * 2, 2ull * Wo * cstride * 2};
  uint32_t box[5] = {64, 1, (uint32_t)bw, 1, (uint32_t)bhn};
  return make_tmap(m, base, 5, dims, str, box);
}

static inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// Persistent grid: one CTA per SM, rounded down to a multiple of the column-tile count.
static inline int conv_grid(int tiles_m, int tiles_nn) {
  long long total = static_cast<long long>(tiles_m) * tiles_nn;
  int grid = static_cast<int>(total < num_sms() ? total : num_sms());
  grid = grid / tiles_nn * tiles_nn;
  if (grid < tiles_nn) grid = tiles_nn;
  return grid;
}


static inline int auto_block_n(int n_total, int cout_sub, int tile_n) {
  if (tile_n == 0) tile_n = (cout_sub >= 128) ? 128 : 64;
  if (tile_n > cout_sub) tile_n = cout_sub;
  (void)n_total;
  return tile_n;
}

// ---- kernel-variant planning --------------------------------------------------------------------
// The `tile_n` argument of the conv entry points: low 10 bits = BLOCK_N (0 = auto); the bits above force a variant.
constexpr int kTileNMask = 1023;
constexpr int kVarPair = 1 << 10;       // tile-pair kernel: two 128-pixel M tiles share every B (weight) stage
constexpr int kVarHalo = 1 << 11;       // row-pair halo kernel: 4 input rows staged once, 9 taps by shifted descriptors
constexpr int kVarLegacy = 1 << 12;     // one M tile per stage (conv_tc_kernel)

enum ConvKind : int { CONV_LEGACY = 0, CONV_PAIR = 1, CONV_HALO = 2 };

struct ConvPlan {
  int kind, block_n, tiles_nn, items_m, grid, stats_rows;
  int bw, bh, bn, tiles_w, tiles_h, tiles_n, tiles_m;
};

// Chooses kernel variant, tile width and persistent grid for a conv over pixel space (N,H,W) with GEMM N = n_total.
static inline int conv_plan(int N, int H, int W, int n_total, int cout_sub, int a_mode, int out_mode, int tile_n_arg,
                            ConvPlan* pl) {
  int tile_n = tile_n_arg & kTileNMask;
  const int force = tile_n_arg & ~kTileNMask;
  pick_box(W, H, N, kBlockM, &pl->bw, &pl->bh, &pl->bn);
  pl->tiles_w = (W + pl->bw - 1) / pl->bw; pl->tiles_h = (H + pl->bh - 1) / pl->bh; pl->tiles_n = (N + pl->bn - 1) / pl->bn;
  pl->tiles_m = pl->tiles_w * pl->tiles_h * pl->tiles_n;
  const bool halo_ok = a_mode == A_CONV3 && out_mode == OUT_4D && W % 128 == 0 && H % 2 == 0;
  int kind;
  if (force & kVarLegacy) kind = CONV_LEGACY;
  else if (force & kVarHalo) { if (!halo_ok) return -1; kind = CONV_HALO; }
  else if (force & kVarPair) kind = CONV_PAIR;
  else kind = halo_ok ? CONV_HALO : (pl->tiles_m >= 2 ? CONV_PAIR : CONV_LEGACY);
  if (kind == CONV_HALO) {
    if (tile_n == 0) tile_n = cout_sub >= 128 ? 128 : 64;
    if (tile_n != 64 && tile_n != 128) return -1;
    pl->bw = 128; pl->bh = 1; pl->bn = 1;
    pl->tiles_w = W / 128; pl->tiles_h = H; pl->tiles_n = N; pl->tiles_m = pl->tiles_w * H * N;
    pl->items_m = N * (H / 2) * (W / 128);
  } else if (kind == CONV_PAIR) {
    if (tile_n == 0) tile_n = cout_sub >= 256 ? 256 : cout_sub >= 128 ? 128 : 64;   // (convT tiles spanning sub-positions: no gain measured)
    pl->items_m = (pl->tiles_m + 1) / 2;
  } else {
    tile_n = auto_block_n(n_total, cout_sub, tile_n);
    pl->items_m = pl->tiles_m;
  }
  // A tile wider than cout_sub is allowed when it covers whole (a,b) sub-blocks of a transposed-conv forward
  // (n_total = 4 * cout_sub): the epilogue maps every 64-column chunk to its sub-position and bias on its own.
  const bool multi_sub = out_mode == OUT_CONVT_5D && kind == CONV_PAIR && tile_n > cout_sub && tile_n % cout_sub == 0 &&
                         n_total % tile_n == 0;
  if (tile_n > cout_sub && !multi_sub) tile_n = cout_sub;
  if ((tile_n != 64 && tile_n != 128 && tile_n != 256) || (!multi_sub && cout_sub % tile_n)) return -1;
  pl->kind = kind;
  pl->block_n = tile_n;
  pl->tiles_nn = n_total / tile_n;
  pl->grid = conv_grid(pl->items_m, pl->tiles_nn);
  pl->stats_rows = 2 * (pl->grid / pl->tiles_nn);
  return 0;
}

// wgrad2_tc.cu: row-halo weight-gradient kernel (W % 128 == 0, Cin and Cout in {64, 128})
bool wgrad_halo_eligible(int N, int H, int W, int Cin, int Cout);
int wgrad_halo_splits(int N, int H, int W, int Cin, int Cout, int splits_req);
int launch_wgrad_halo(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H, int W,
                      int Cin, int Cout, int splits_req, cudaStream_t stream);

// conv2_tc.cu: tile-pair and halo kernels
int launch_conv2(const ConvPlan& pl, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                 const ConvTcParams& p, cudaStream_t stream);

}  // namespace b2s
