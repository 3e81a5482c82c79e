// Row-halo weight-gradient kernel for wide images (W % 128 == 0, Cin and Cout <= 128): tcgen05 + TMEM + TMA.
//
//   dW[(tap, ci), co] = sum_pix x[pix + tap, ci] * dz[pix, co]            (autograd of models/model.py:36,39)
//
// wgrad_tc_kernel loads one shifted 64-pixel x box per (tap, ci block) and k step: every activation byte crosses
// L2 -> shared memory nine times and is used for only BLOCK_N columns, which pins the Cout = 64 layers at ~20 % of
// the tensor pipe. Here a CTA stages, per output row of 128 pixels, the (up to three) x rows that row needs ONCE
// (130 pixels each, zero-filled by TMA outside the image) plus the dz row, and runs ALL of its taps from that one
// stage: a tap is just a different start address (whole 128-byte pixel rows) of the MN-major A descriptor. The
// accumulators of all (tap, ci) row blocks of the CTA live in TMEM for the whole kernel (split-K over row tiles
// only); fp32 partials go to the same ws[split][tap*Cin+ci][co] workspace b2s_wgrad_reduce consumes.
//
// M tiles (128 accumulator lanes): Cin = 128 -> one tap x 128 input channels (two 64-channel panels, LBO = chunk
// stride); Cin = 64 -> two taps x 64 channels (LBO = byte distance between the two taps' start pixels).
// When tiles x BLOCK_N exceeds the 512 TMEM columns the taps are split over 2-3 CTA groups.
#include "conv_common.cuh"

namespace b2s {

constexpr int kWg2Threads = 192;            // warp0 producer (+TMEM alloc), warp1 MMA issuer, warps 2-5 epilogue
constexpr int kWgHaloRowBytes = 17 * 1024;  // 130 px x 128 B, padded to a multiple of 1024
constexpr int kWgMaxGroups = 4;
constexpr int kWgMaxTiles = 20;

struct WgHaloParams {
  int strips, H;                  // 128-pixel strips per image row, rows per image
  int rows_total, rows_per_split; // row tiles = N * H * strips, split over `splits` CTAs per group
  int chunks;                     // Cin / 64
  int groups;
  int g_tile0[kWgMaxGroups], g_ntiles[kWgMaxGroups];   // first global M tile and tile count of each CTA group
  int g_th_lo[kWgMaxGroups], g_nth[kWgMaxGroups];      // first tap row (0..2) and input rows staged per output row
  uint32_t tile_lo[kWgMaxTiles];  // per global M tile: (offset in the stage >> 4) | (LBO >> 4) << 16
  int rows_m;                     // 9 * Cin: valid rows of the (tap, ci) space
  int cout;
  float* ws;                      // [splits][9 * Cin][cout] fp32
};

template <int BLOCK_N, int A_SLOTS, int STAGES>
struct Wg2Cfg {
  static constexpr int kABytes = A_SLOTS * kWgHaloRowBytes;
  static constexpr int kBBytes = (BLOCK_N / 64) * 16384;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * STAGES + 1) * 8;
  static constexpr int kTotal = kTmemPtrOffset + 16;
  static constexpr int kDynBytes = kTotal + 1024;
  static_assert(kDynBytes <= 227 * 1024, "exceeds the 227 KB of shared memory a CTA may use");
};

template <int BLOCK_N, int A_SLOTS, int STAGES, int MAXT>
__global__ void __launch_bounds__(kWg2Threads, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDz,
                  const WgHaloParams p) {
  using L = Wg2Cfg<BLOCK_N, A_SLOTS, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x % p.groups;
  const int split = blockIdx.x / p.groups;
  const int row_begin = split * p.rows_per_split;
  const int row_end = min(row_begin + p.rows_per_split, p.rows_total);
  const int my_rows = max(row_end - row_begin, 0);
  const int tile0 = p.g_tile0[group], ntiles = p.g_ntiles[group];
  const int th_lo = p.g_th_lo[group], nth = p.g_nth[group];

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmDz);
    }
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===== producer: per output row tile, nth x-rows of 130 pixels per 64-channel chunk + the dz row =====
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = nth * p.chunks * 130 * 128 + (BLOCK_N / 64) * 16384;
      for (int rt = row_begin; rt < row_end; ++rt) {
        const int ws_ = rt % p.strips;
        const int h = (rt / p.strips) % p.H;
        const int n = rt / (p.strips * p.H);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        uint8_t* a_dst = smem + stage * L::kStageBytes;
        uint8_t* b_dst = a_dst + L::kABytes;
        for (int r = 0; r < nth; ++r)
          for (int c = 0; c < p.chunks; ++c)
            tma_load_4d(&tmX, &full_bar[stage], a_dst + (r * p.chunks + c) * kWgHaloRowBytes, c * 64, ws_ * 128 - 1,
                        h + th_lo + r - 1, n);
#pragma unroll
        for (int c = 0; c < BLOCK_N / 64; ++c)
          tma_load_4d(&tmDz, &full_bar[stage], b_dst + c * 16384, c * 64, ws_ * 128, h, n);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp walks the ring (uniform registers), one elected lane issues =====
    if (my_rows > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 1, 1);   // both operands MN-major
      const uint32_t tbase = __reduce_or_sync(0xffffffffu, tmem_base);
      const bool leader = elect_one();
      const uint64_t proto = umma_smem_desc_sw128(0, 0, 1024);   // SBO = 1024 B between 8-pixel groups
      const uint32_t desc_hi = static_cast<uint32_t>(proto >> 32);
      const uint32_t b_lo0 = static_cast<uint32_t>(proto) | ((16384u >> 4) << 16);   // dz: 64-channel panels 16 KB apart
      const uint32_t smem_lo = smem_u32(smem) >> 4;
      // MAXT = M tiles a CTA group of this instantiation can own. Kept tight: at BLOCK_N = 64 an MMA lasts 32 clocks
      // and every predicated slot of the unrolled issue loop costs issue time.
      constexpr int kMaxT = MAXT;
      static_assert(MAXT * BLOCK_N <= 512, "accumulators exceed TMEM");
      uint32_t tl[kMaxT];
#pragma unroll
      for (int t = 0; t < kMaxT; ++t) tl[t] = p.tile_lo[min(tile0 + t, kWgMaxTiles - 1)];
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_rows; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = smem_lo + stage * (L::kStageBytes >> 4);
        const uint32_t b_lo = b_lo0 + a_lo + (L::kABytes >> 4);
        if (leader) {
#pragma unroll 2
          for (int j = 0; j < 8; ++j) {        // 16 pixels (K rows of 128 B) per MMA
#pragma unroll
            for (int t = 0; t < kMaxT; ++t)
              if (t < ntiles)
                umma_bf16_lohi(tbase + t * BLOCK_N, a_lo + tl[t] + j * 128, desc_hi, b_lo + j * 128, desc_hi, idesc,
                               (it | j) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(tmem_full_bar);
      __syncwarp();
    }
  } else {
    // ===== epilogue: fp32 accumulators -> ws[split][tile * 128 + row][co] =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (my_rows > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    for (int t = 0; t < ntiles; ++t) {
      const int grow = (tile0 + t) * 128 + row;
      const bool valid = grow < p.rows_m;
      float* dst = p.ws + (static_cast<size_t>(split) * p.rows_m + grow) * p.cout;
      if (my_rows > 0) {
#pragma unroll 1
        for (int s = 0; s < BLOCK_N / 32; ++s) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * BLOCK_N + s * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(dst + s * 32 + c * 4) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
        }
      } else if (valid) {
        for (int c = 0; c < BLOCK_N / 4; ++c) *reinterpret_cast<uint4*>(dst + c * 4) = make_uint4(0, 0, 0, 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BLOCK_N, int A_SLOTS, int STAGES, int MAXT>
static int launch_wg2(int grid, const CUtensorMap& tmX, const CUtensorMap& tmDz, const WgHaloParams& p,
                      cudaStream_t stream) {
  using L = Wg2Cfg<BLOCK_N, A_SLOTS, STAGES>;
  auto kfn = wgrad_halo_kernel<BLOCK_N, A_SLOTS, STAGES, MAXT>;
  static std::atomic<unsigned long long> attr_devices{0};
  {
    cudaError_t e = allow_dynamic_smem(kfn, L::kDynBytes, attr_devices);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(wgrad_halo_kernel)");
  }
  kfn<<<grid, kWg2Threads, L::kDynBytes, stream>>>(tmX, tmDz, p);
  return check_launch("wgrad_halo_kernel");
}

bool wgrad_halo_eligible(int N, int H, int W, int Cin, int Cout) {
  if (W % 128 != 0 || N <= 0 || H <= 0) return false;
  if ((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128)) return true;
  return Cin == 256 && Cout == 64;   // 18 tiles of (tap, 128-channel half): three CTA groups of one tap row each
}

static int wg_num_tiles(int Cin) { return Cin == 64 ? 5 : 9 * (Cin / 128); }

// Splits (CTAs per tap group) the halo kernel uses: one CTA per SM in total.
int wgrad_halo_splits(int N, int H, int W, int Cin, int Cout, int splits_req) {
  const int tiles = wg_num_tiles(Cin);
  const int max_tiles = 512 / Cout;
  const int groups = (tiles + max_tiles - 1) / max_tiles;
  const int rows_total = N * H * (W / 128);
  int splits = splits_req > 0 ? splits_req : num_sms() / groups;
  if (splits > rows_total) splits = rows_total;
  if (splits < 1) splits = 1;
  const int rps = (rows_total + splits - 1) / splits;
  return (rows_total + rps - 1) / rps;
}

int launch_wgrad_halo(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H, int W,
                      int Cin, int Cout, int splits_req, cudaStream_t stream) {
  if (!wgrad_halo_eligible(N, H, W, Cin, Cout)) return set_error(B2S_ERR_ARG, "wgrad halo kernel: unsupported shape");
  WgHaloParams p{};
  const int chunks = Cin / 64;
  const int tiles = wg_num_tiles(Cin);          // M tiles of 128 (tap, ci) rows; Cin = 64 pairs two taps per tile
  const int halves = Cin >= 128 ? Cin / 128 : 1;   // 128-channel halves per tap (Cin >= 128)
  const int max_tiles = 512 / Cout;             // TMEM columns / BLOCK_N
  const int groups = (tiles + max_tiles - 1) / max_tiles;
  const int per = (tiles + groups - 1) / groups;
  p.strips = W / 128; p.H = H;
  p.rows_total = N * H * p.strips;
  const int splits = wgrad_halo_splits(N, H, W, Cin, Cout, splits_req);
  p.rows_per_split = (p.rows_total + splits - 1) / splits;
  p.chunks = chunks; p.groups = groups; p.rows_m = 9 * Cin; p.cout = Cout; p.ws = ws;
  const int RS = chunks * kWgHaloRowBytes;      // byte stride between consecutive staged input rows
  int max_slots = 0;
  for (int g = 0; g < groups; ++g) {
    const int a = g * per, b = (a + per < tiles) ? a + per : tiles;
    const int tap_first = Cin == 64 ? 2 * a : a / halves;
    const int tap_last = Cin == 64 ? (2 * b - 1 < 8 ? 2 * b - 1 : 8) : (b - 1) / halves;
    const int th_lo = tap_first / 3, th_hi = tap_last / 3;
    p.g_tile0[g] = a; p.g_ntiles[g] = b - a; p.g_th_lo[g] = th_lo; p.g_nth[g] = th_hi - th_lo + 1;
    if (p.g_nth[g] * chunks > max_slots) max_slots = p.g_nth[g] * chunks;
    for (int m = a; m < b; ++m) {
      const int t0 = Cin == 64 ? 2 * m : m / halves;
      // tile m of a Cin >= 128 layer = (tap m / halves, channels 128 * (m % halves) ...): two chunk slots further per half
      const int off0 = (t0 / 3 - th_lo) * RS + (t0 % 3) * 128 + (Cin >= 128 ? (m % halves) * 2 * kWgHaloRowBytes : 0);
      int lbo;
      if (Cin == 64) {
        const int t1 = t0 + 1;   // second 64-channel panel = the next tap (tile 4: a discarded duplicate)
        lbo = t1 <= 8 ? ((t1 / 3 - th_lo) * RS + (t1 % 3) * 128) - off0 : 128;
      } else {
        lbo = kWgHaloRowBytes;   // second panel = channels 64..127 of the same tap (next chunk slot)
      }
      p.tile_lo[m] = (static_cast<uint32_t>(off0) >> 4) | ((static_cast<uint32_t>(lbo) >> 4) << 16);
    }
  }
  CUtensorMap tmX, tmDz;
  int rc;
  if ((rc = make_act_map4(&tmX, x, Cin, W, H, N, x_cstride, 130, 1, 1))) return rc;
  if ((rc = make_act_map4(&tmDz, dz, Cout, W, H, N, dz_cstride, 128, 1, 1))) return rc;
  const int grid = groups * splits;
  int max_tiles_group = 0;
  for (int g = 0; g < groups; ++g)
    if (p.g_ntiles[g] > max_tiles_group) max_tiles_group = p.g_ntiles[g];
  count_launch();
  if (Cout == 64 && max_slots <= 3 && max_tiles_group <= 5) return launch_wg2<64, 3, 3, 5>(grid, tmX, tmDz, p, stream);
  if (Cout == 64 && max_slots <= 4 && max_tiles_group <= 5) return launch_wg2<64, 4, 2, 5>(grid, tmX, tmDz, p, stream);
  if (Cout == 64 && max_slots <= 4 && max_tiles_group <= 6) return launch_wg2<64, 4, 2, 6>(grid, tmX, tmDz, p, stream);
  if (Cout == 128 && max_slots <= 2 && max_tiles_group <= 3) return launch_wg2<128, 2, 3, 3>(grid, tmX, tmDz, p, stream);
  return set_error(B2S_ERR_ARG, "wgrad halo kernel: no instantiation for this shape");
}

}  // namespace b2s
