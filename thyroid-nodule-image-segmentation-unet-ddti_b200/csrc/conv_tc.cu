// Implicit-GEMM convolution kernels on tcgen05 tensor cores (sm_100a), TMA-fed, accumulators in TMEM.
//
//   conv_tc_kernel   : out[pix, n] = sum_tap sum_c A_tap[pix, c] * B[tap][n][c]   (K-major operands)
//                      covers 3x3 conv forward, 3x3 conv dgrad (rotated weights), 1x1 conv,
//                      ConvTranspose2d(k2,s2) forward (pixel-shuffle TMA store) and its dgrad.
//                      Replaces torch.nn.Conv2d / ConvTranspose2d calls of reference models/model.py:36,39,49.
//   wgrad_tc_kernel  : dW[(tap,ci), co] = sum_pix x[pix+tap, ci] * dz[pix, co]        (MN-major operands)
//                      split-K over pixels, fp32 partials to a workspace, deterministic second-stage reduce.
//
// conv_tc_kernel is persistent (one CTA per SM, 320 threads): warp0 = TMA producer (+TMEM alloc), warp1 = MMA
// issuer, warps 2-5 / 6-9 = two epilogue groups draining the two TMEM accumulators alternately.
// wgrad_tc_kernel runs one long split-K tile per CTA (192 threads, one epilogue group).
#include "conv_common.cuh"

namespace b2s {

constexpr int kNumThreads = 192;   // wgrad kernel: warp0 TMA, warp1 MMA, warps 2-5 epilogue


constexpr int kConvThreads = 320;   // warp0 TMA, warp1 MMA, warps 2-5 epilogue group 0, warps 6-9 epilogue group 1

template <int BLOCK_N, int STAGES>
struct ConvSmem {
  static constexpr int kBTileBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kPipeBytes = STAGES * kStageBytes;
  static constexpr int kStagingOffset = kPipeBytes;                    // one 16 KB store-staging tile per group
  static constexpr int kBarOffset = kStagingOffset + 2 * kATileBytes;   // full[S], empty[S], tmem_full[2], tmem_empty[2]
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * STAGES + 4) * 8;
  static constexpr int kStatsOffset = (kTmemPtrOffset + 4 + 15) / 16 * 16;  // [2 groups][4 warps][2][64] float
  static constexpr int kTotal = kStatsOffset + 2 * 4 * 2 * 64 * 4;
  static constexpr int kDynBytes = kTotal + 1024;  // slack for manual 1024-B alignment
  static_assert(kDynBytes <= 227 * 1024, "exceeds the 227 KB of shared memory a CTA may use");
};

// Persistent kernel: CTA c owns column tile n_tile = c % tiles_nn and the row tiles m = c / tiles_nn + i * groups.
// Three pipelines: smem ring (TMA -> MMA), two TMEM accumulators (MMA -> epilogue), one staging tile per epilogue
// group (epilogue -> TMA store). The producer runs ahead across tile boundaries; the two epilogue groups alternate
// tiles so that TMEM drain + bias/ReLU/statistics + store of tile i overlap the MMAs of tiles i+1, i+2.
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmOut, const ConvTcParams p) {
  using L = ConvSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kATileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);
  float* stats_smem = reinterpret_cast<float*>(smem + L::kStatsOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- static tile schedule -----------------------------------------------------------------
  const int n_tile = blockIdx.x % p.tiles_nn;
  const int mgroup = blockIdx.x / p.tiles_nn;
  const int num_mgroups = gridDim.x / p.tiles_nn;
  const int ncol0 = n_tile * BLOCK_N;  // first GEMM column of this CTA
  const int my_tiles = (p.tiles_m - mgroup + num_mgroups - 1) / num_mgroups;   // >= 1 by construction
  const int k_iters = p.num_taps * p.k_chunks;

  // ---- one-time setup -----------------------------------------------------------------------
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      tma_prefetch_desc(&tmOut);
    }
    tmem_alloc(tmem_ptr_smem, 2 * BLOCK_N);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&tmem_full_bar[g], 1);
      mbar_init(&tmem_empty_bar[g], 4);   // one arrival per epilogue warp of the group
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: runs ahead of the MMA warp across tile boundaries =====
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m_tile = mgroup + i * num_mgroups;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        for (int it = 0; it < k_iters; ++it) {
          const int tap = it / p.k_chunks;
          const int kc = it - tap * p.k_chunks;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          uint8_t* a_dst = smem_a + stage * kATileBytes;
          uint8_t* b_dst = smem_b + stage * L::kBTileBytes;
          if (p.a_mode == A_CONV3) {
            const int dh = tap / 3 - 1, dw = tap - (tap / 3) * 3 - 1;
            tma_load_4d(&tmA, &full_bar[stage], a_dst, kc * kBlockK, w0 + dw, h0 + dh, n0);
          } else if (p.a_mode == A_1X1) {
            tma_load_4d(&tmA, &full_bar[stage], a_dst, kc * kBlockK, w0, h0, n0);
          } else {  // A_CONVT_DGRAD: dY viewed as (C, b, j, a, i*N)
            const int a = tap >> 1, b = tap & 1;
            tma_load_5d(&tmA, &full_bar[stage], a_dst, kc * kBlockK, b, w0, a, n0 * p.H + h0);
          }
          // B rows: [tap][n_total] x Cin; the host encodes the box as (64, BLOCK_N)
          tma_load_2d(&tmB, &full_bar[stage], b_dst, kc * kBlockK, tap * p.n_total + ncol0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (single thread) =====
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int acc = i & 1;
        const uint32_t acc_phase = (i >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * BLOCK_N;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * kATileBytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * L::kBTileBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_acc, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue groups: TMEM -> regs -> (+bias, ReLU) -> bf16 -> swizzled smem -> TMA store; BN partial sums =====
    const int g = (warp - 2) >> 2;   // group 0: warps 2-5, group 1: warps 6-9; group g drains accumulator g
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;   // tile row == TMEM lane == linear thread id inside the group
    const bool do_relu = p.flags & B2S_FLAG_RELU;
    const bool do_stats = (p.flags & B2S_FLAG_STATS) && p.stats != nullptr;
    uint8_t* stage_buf = smem + L::kStagingOffset + g * kATileBytes;
    const int bar_id = 1 + g;
    const int wl = row % p.bw;
    const int hl = (row / p.bw) % p.bh;
    const int nl = row / (p.bw * p.bh);
    float st_acc[BLOCK_N / 64][4];   // per chunk: sum, sum of the two columns (2*lane, 2*lane+1), then squares
#pragma unroll
    for (int s = 0; s < BLOCK_N / 64; ++s) st_acc[s][0] = st_acc[s][1] = st_acc[s][2] = st_acc[s][3] = 0.f;

    for (int i = g; i < my_tiles; i += 2) {
      const int m_tile = mgroup + i * num_mgroups;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
      const bool valid = (w0 + wl < p.W) && (h0 + hl < p.H) && (n0 + nl < p.N);
      mbar_wait(&tmem_full_bar[g], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + g * BLOCK_N;

#pragma unroll
      for (int s = 0; s < BLOCK_N / 64; ++s) {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + s * 64;
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        if (s == BLOCK_N / 64 - 1) {
          // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[g]);
        }
        const int col_base = ncol0 + s * 64;           // GEMM column of v0[0]
        const int bias_base = col_base % p.cout_sub;
        uint32_t packed[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = __uint_as_float(j < 16 ? v0[2 * j] : v1[2 * j - 32]);
          float b = __uint_as_float(j < 16 ? v0[2 * j + 1] : v1[2 * j - 31]);
          if (p.bias != nullptr) {
            a += __ldg(p.bias + bias_base + 2 * j);
            b += __ldg(p.bias + bias_base + 2 * j + 1);
          }
          if (do_relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
          if (p.post_scale != nullptr) {
            a = fmaf(a, __ldg(p.post_scale + bias_base + 2 * j), __ldg(p.post_shift + bias_base + 2 * j));
            b = fmaf(b, __ldg(p.post_scale + bias_base + 2 * j + 1), __ldg(p.post_shift + bias_base + 2 * j + 1));
          }
          if (!valid) { a = 0.f; b = 0.f; }
          packed[j] = pack_bf16x2(a, b);
        }
        // the previous TMA store of this group must have finished READING the staging tile
        if (row == 0) tma_store_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        // 128-B row, 16-B chunk c stored at chunk (c ^ (row & 7)): the SWIZZLE_128B pattern of tmOut
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 val = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
          *reinterpret_cast<uint4*>(stage_buf + row * 128 + ((c ^ (row & 7)) << 4)) = val;
        }
        if (do_stats) {
          // column sums over this warp's own 32 rows, taken from the bf16-rounded values actually stored
          __syncwarp();
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const int chunk = lane >> 2, within = (lane & 3) * 4;  // lane <-> column pair (2*lane, 2*lane+1)
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const int rr = q * 32 + r;
            const uint32_t u =
                *reinterpret_cast<const uint32_t*>(stage_buf + rr * 128 + ((chunk ^ (rr & 7)) << 4) + within);
            const float x0 = bf16_lo(u), x1 = bf16_hi(u);
            s0 += x0; s1 += x1;
            q0 = fmaf(x0, x0, q0); q1 = fmaf(x1, x1, q1);
          }
          st_acc[s][0] += s0; st_acc[s][1] += s1; st_acc[s][2] += q0; st_acc[s][3] += q1;
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (row == 0) {
          if (p.out_mode == OUT_4D) {
            tma_store_4d(&tmOut, stage_buf, col_base, w0, h0, n0);
          } else {  // convT forward: column block belongs to one (a,b) sub-position
            const int ab = col_base / p.cout_sub;
            tma_store_5d(&tmOut, stage_buf, col_base - ab * p.cout_sub, ab & 1, w0, ab >> 1, n0 * p.H + h0);
          }
          tma_store_commit();
        }
      }
    }
    if (row == 0) tma_store_wait_read<0>();
    if (do_stats) {
      // one partial row per (CTA row-group, epilogue group): [2*mgroup + g][2][n_total]
      float* sb = stats_smem + g * (4 * 2 * 64);
      float* out_row = p.stats + static_cast<size_t>(2 * mgroup + g) * 2 * p.n_total;
#pragma unroll
      for (int s = 0; s < BLOCK_N / 64; ++s) {
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        sb[q * 128 + 2 * lane] = st_acc[s][0]; sb[q * 128 + 2 * lane + 1] = st_acc[s][1];
        sb[q * 128 + 64 + 2 * lane] = st_acc[s][2]; sb[q * 128 + 64 + 2 * lane + 1] = st_acc[s][3];
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const int which = row >> 6, col = row & 63;   // 128 threads: (sum | sumsq) x 64 columns
        float acc = 0.f;
#pragma unroll
        for (int wq = 0; wq < 4; ++wq) acc += sb[wq * 128 + which * 64 + col];
        out_row[which * p.n_total + ncol0 + s * 64 + col] = acc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[(tap,ci), co] = sum_pix A[pix, (tap,ci)] * B[pix, co], both operands MN-major in smem
// ------------------------------------------------------------------------------------------------
enum WgMode : int { WG_CONV3 = 0, WG_CONVT = 1, WG_1X1 = 2, WG_CONV3_S2 = 3 };

struct WgradParams {
  int pw, ph, pn;                 // pixel box of one K chunk, pw*ph*pn == 64
  int chunks_w, chunks_h, chunks_n;
  int H;                          // rows per image (5D merged coordinate)
  int num_taps, ci_blocks;        // row space = num_taps x (Cin/64) blocks of 64 rows
  int total_rb;                   // num_taps * ci_blocks
  int cin, cout;
  int tiles_nn;                   // cout / BLOCK_N
  int mt;                         // M tiles (pairs of row blocks) per CTA: 1 or 2
  int m_tiles;                    // ceil(total_rb / (2 * mt))
  int splits, chunks_total, chunks_per_split;
  int mode;
  int x_cstride;                  // WG_CONV3_S2: pixel stride (elements) of x, the offset of the odd-column parity
  float* ws;                      // [splits][num_taps*cin][cout] fp32
};

template <int BLOCK_N, int STAGES, int MT>
struct WgSmem {
  static constexpr int kABytes = MT * 2 * 64 * 128;       // MT M tiles x two (64 px x 64 ch) boxes
  static constexpr int kBBytes = (BLOCK_N / 64) * 64 * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPipeBytes = STAGES * kStageBytes;
  static constexpr int kBarOffset = kPipeBytes;
  static constexpr int kTmemPtrOffset = kBarOffset + (2 * STAGES + 1) * 8;
  static constexpr int kTotal = kTmemPtrOffset + 16;
  static constexpr int kDynBytes = kTotal + 1024;
  static constexpr int kTmemCols = MT * BLOCK_N;
  static_assert(kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns");
  static_assert(kDynBytes <= 227 * 1024, "exceeds the 227 KB of shared memory a CTA may use");
};

// MT = 2: two M tiles (four (tap, ci) row blocks) share every dz (B) stage: 131 FLOP per L2 byte at BLOCK_N = 256.
template <int BLOCK_N, int STAGES, int MT>
__global__ void __launch_bounds__(kNumThreads)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const WgradParams p) {
  using L = WgSmem<BLOCK_N, STAGES, MT>;
  constexpr int NRB = 2 * MT;   // row blocks (64 rows of the (tap, ci) space) per CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * L::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tile = blockIdx.x % p.tiles_nn;
  const int m_tile = (blockIdx.x / p.tiles_nn) % p.m_tiles;
  const int split = blockIdx.x / (p.tiles_nn * p.m_tiles);
  const int chunk_begin = split * p.chunks_per_split;
  const int chunk_end = min(chunk_begin + p.chunks_per_split, p.chunks_total);
  const int k_iters = max(chunk_end - chunk_begin, 0);

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    tmem_alloc(tmem_ptr_smem, L::kTmemCols);
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
  const uint32_t tmem_acc = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // row blocks (tap, ci block of 64) of this CTA; a padding block re-loads the last valid one
      int tap_j[NRB], cib_j[NRB];
#pragma unroll
      for (int j = 0; j < NRB; ++j) {
        int rb = min(m_tile * NRB + j, p.total_rb - 1);
        tap_j[j] = rb / p.ci_blocks;
        cib_j[j] = rb - tap_j[j] * p.ci_blocks;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < k_iters; ++it) {
        const int chunk = chunk_begin + it;
        const int cw = chunk % p.chunks_w;
        const int chh = (chunk / p.chunks_w) % p.chunks_h;
        const int cn = chunk / (p.chunks_w * p.chunks_h);
        const int w0 = cw * p.pw, h0 = chh * p.ph, n0 = cn * p.pn;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        uint8_t* a_dst = smem_a + stage * L::kABytes;
        uint8_t* b_dst = smem_b + stage * L::kBBytes;
#pragma unroll
        for (int j = 0; j < NRB; ++j) {
          int dh = 0, dw = 0;
          if (p.mode == WG_CONV3 || p.mode == WG_CONV3_S2) { dh = tap_j[j] / 3 - 1; dw = tap_j[j] % 3 - 1; }
          if (p.mode == WG_CONV3_S2) {   // stride-2 conv: one parity class of x per tap (make_act_map5_parity)
            const int ph = dh & 1, pw = dw & 1;
            tma_load_5d(&tmA, &full_bar[stage], a_dst + j * 8192, pw * p.x_cstride + cib_j[j] * 64, w0 + ((dw - pw) >> 1),
                        ph, h0 + ((dh - ph) >> 1), n0);
          } else
            tma_load_4d(&tmA, &full_bar[stage], a_dst + j * 8192, cib_j[j] * 64, w0 + dw, h0 + dh, n0);
        }
        if (p.mode != WG_CONVT) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_4d(&tmB, &full_bar[stage], b_dst + j * 8192, n_tile * BLOCK_N + j * 64, w0, h0, n0);
        } else {
          // convT: all row blocks of a CTA share one tap (host guarantees ci_blocks % NRB == 0)
          const int a = tap_j[0] >> 1, b = tap_j[0] & 1;
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_5d(&tmB, &full_bar[stage], b_dst + j * 8192, n_tile * BLOCK_N + j * 64, b, w0, a,
                        n0 * p.H + h0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp walks the ring (addresses stay in uniform registers); one elected lane issues
    if (k_iters > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 1, 1);
      const uint32_t tbase = __reduce_or_sync(0xffffffffu, tmem_acc);
      const bool leader = elect_one();
      const uint64_t proto = umma_smem_desc_sw128(0, 8192, 1024);   // 64-wide panels 8 KB apart, 8-pixel groups 1 KB
      const uint32_t desc_hi = static_cast<uint32_t>(proto >> 32);
      const uint32_t a_lo0 = static_cast<uint32_t>(proto) + (smem_u32(smem_a) >> 4);
      const uint32_t b_lo0 = static_cast<uint32_t>(proto) + (smem_u32(smem_b) >> 4);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < k_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + stage * (L::kABytes >> 4);
        const uint32_t b_lo = b_lo0 + stage * (L::kBBytes >> 4);
        if (leader) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 16 pixels (K rows of 128 B) per MMA
              umma_bf16_lohi(tbase + mt * BLOCK_N, a_lo + mt * (16384 >> 4) + k * (2048 >> 4), desc_hi,
                             b_lo + k * (2048 >> 4), desc_hi, idesc, k != 0 ? 1u : static_cast<uint32_t>(it != 0));
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
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (k_iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int rb = (m_tile * MT + mt) * 2 + (row >> 6);
      const bool valid = rb < p.total_rb;
      const size_t grow = static_cast<size_t>(rb) * 64 + (row & 63);  // row in (tap, ci) space
      float* dst = p.ws + (static_cast<size_t>(split) * p.num_taps * p.cin + grow) * p.cout + n_tile * BLOCK_N;
      if (k_iters > 0) {
#pragma unroll 1
        for (int s = 0; s < BLOCK_N / 32; ++s) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + mt * BLOCK_N + s * 32, v);
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
    tmem_dealloc(tmem_acc, L::kTmemCols);
  }
}

template <int BLOCK_N, int STAGES>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                       const ConvTcParams& p, cudaStream_t stream) {
  using L = ConvSmem<BLOCK_N, STAGES>;
  auto kfn = conv_tc_kernel<BLOCK_N, STAGES>;
  static std::atomic<unsigned long long> attr_devices{0};
  {
    cudaError_t e = allow_dynamic_smem(kfn, L::kDynBytes, attr_devices);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_tc_kernel)");
  }
  kfn<<<conv_grid(p.tiles_m, p.tiles_nn), kConvThreads, L::kDynBytes, stream>>>(tmA, tmB, tmOut, p);
  return check_launch("conv_tc_kernel");
}

// Routes a planned conv to conv_tc_kernel (one tile per stage) or to the tile-pair / halo kernels of conv2_tc.cu.
static int dispatch_conv(const ConvPlan& pl, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                         ConvTcParams& p, cudaStream_t stream) {
  p.bw = pl.bw; p.bh = pl.bh; p.bn = pl.bn;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.tiles_n = pl.tiles_n; p.tiles_m = pl.tiles_m;
  p.tiles_nn = pl.tiles_nn; p.items_m = pl.items_m;
  if (pl.kind != CONV_LEGACY) return launch_conv2(pl, tmA, tmB, tmOut, p, stream);
  count_launch();
  switch (pl.block_n) {
    case 64:  return launch_conv<64, 7>(tmA, tmB, tmOut, p, stream);    // 7 x 24 KB + 32 KB staging
    case 128: return launch_conv<128, 5>(tmA, tmB, tmOut, p, stream);   // 5 x 32 KB + 32 KB staging
    case 256: return launch_conv<256, 3>(tmA, tmB, tmOut, p, stream);   // 3 x 48 KB + 32 KB staging (all 512 TMEM cols)
    default:  return set_error(B2S_ERR_ARG, "unsupported tile_n");
  }
}

static int make_weight_map(CUtensorMap* m, const void* w_packed, int K, int rows, int block_n) {
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {64, (uint32_t)block_n};
  return make_tmap(m, w_packed, 2, dims, str, box);
}

}  // namespace b2s

using namespace b2s;

// Generic tap-GEMM conv. ksize 3 (pad 1) or 1. See include/b2s.h.
static int conv_fwd_impl(const void* x, int x_cstride, const void* w_packed, const float* bias, const float* post_scale,
                         const float* post_shift, void* y, int y_cstride, float* stats_partial, int N, int H, int W,
                         int Cin, int Cout, int ksize, int flags, int tile_n, void* stream_,
                         const void* red_r = nullptr, int red_cs = 0) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if ((post_scale == nullptr) != (post_shift == nullptr))
    return set_error(B2S_ERR_ARG, "b2s_conv_fwd: post_scale and post_shift must both be given");
  if (!x || !w_packed || !y) return set_error(B2S_ERR_ARG, "b2s_conv_fwd: null pointer");
  if (ksize != 3 && ksize != 1) return set_error(B2S_ERR_ARG, "b2s_conv_fwd: ksize must be 1 or 3");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_conv_fwd: Cin and Cout must be multiples of 64");
  if (N <= 0 || H <= 0 || W <= 0) return set_error(B2S_ERR_ARG, "b2s_conv_fwd: empty tensor");
  if ((flags & B2S_FLAG_STATS) && !stats_partial) return set_error(B2S_ERR_ARG, "b2s_conv_fwd: stats buffer missing");
  ConvTcParams p{};
  p.a_mode = ksize == 3 ? A_CONV3 : A_1X1; p.out_mode = OUT_4D;
  ConvPlan pl;
  if (conv_plan(N, H, W, Cout, Cout, p.a_mode, p.out_mode, tile_n, &pl))
    return set_error(B2S_ERR_ARG, "b2s_conv_fwd: tile_n must be 64/128/256, divide Cout and suit the forced variant");
  p.W = W; p.H = H; p.N = N;
  p.num_taps = ksize * ksize; p.k_chunks = Cin / 64;
  p.n_total = Cout; p.cout_sub = Cout;
  p.flags = flags; p.bias = bias; p.stats = stats_partial;
  p.post_scale = post_scale; p.post_shift = post_shift;
  if (flags & B2S_FLAG_BNRED) {
    if (!red_r || !stats_partial || red_cs % 8) return set_error(B2S_ERR_ARG, "b2s_conv_dgrad_bnred: r / partial missing");
    if (pl.kind == CONV_LEGACY) return 1;    // the one-tile kernel has no fused reduction: caller takes the two-pass path
    p.red_r = static_cast<const __nv_bfloat16*>(red_r); p.red_cs = red_cs;
  }

  CUtensorMap tmA, tmB, tmOut;
  int rc;
  if (pl.kind == CONV_HALO) {
    // one input row of 130 pixels (128 + halo) per load; out-of-image pixels are zero-filled by TMA
    if ((rc = make_act_map4(&tmA, x, Cin, W, H, N, x_cstride, 130, 1, 1))) return rc;
  } else {
    if ((rc = make_act_map4(&tmA, x, Cin, W, H, N, x_cstride, pl.bw, pl.bh, pl.bn))) return rc;
  }
  if ((rc = make_weight_map(&tmB, w_packed, Cin, p.num_taps * Cout, pl.block_n))) return rc;
  if ((rc = make_act_map4(&tmOut, y, Cout, W, H, N, y_cstride, pl.bw, pl.bh, pl.bn))) return rc;
  return dispatch_conv(pl, tmA, tmB, tmOut, p, stream);
}

extern "C" int b2s_conv_fwd(const void* x, int x_cstride, const void* w_packed, const float* bias, void* y,
                            int y_cstride, float* stats_partial, int N, int H, int W, int Cin, int Cout, int ksize,
                            int flags, int tile_n, void* stream) {
  return conv_fwd_impl(x, x_cstride, w_packed, bias, nullptr, nullptr, y, y_cstride, stats_partial, N, H, W, Cin, Cout,
                       ksize, flags, tile_n, stream);
}

// Inference: conv (+bias, ReLU) with the eval-mode BatchNorm affine applied in the epilogue (no separate BN pass).
extern "C" int b2s_conv_fwd_affine(const void* x, int x_cstride, const void* w_packed, const float* bias,
                                   const float* post_scale, const float* post_shift, void* y, int y_cstride, int N,
                                   int H, int W, int Cin, int Cout, int ksize, int flags, int tile_n, void* stream) {
  if (!post_scale || !post_shift) return set_error(B2S_ERR_ARG, "b2s_conv_fwd_affine: null pointer");
  return conv_fwd_impl(x, x_cstride, w_packed, bias, post_scale, post_shift, y, y_cstride, nullptr, N, H, W, Cin, Cout,
                       ksize, flags & ~B2S_FLAG_STATS, tile_n, stream);
}

// Input gradient of a conv (a conv over dz with the rotated weights) whose output dy is the gradient of a train-mode
// BatchNorm output: the epilogue also emits partial [b2s_conv_stats_rows][2][Cout] = {sum dy, sum dy * r} per channel
// (r = that BatchNorm's saved input), which replaces the separate reduce pass of the BatchNorm backward
// (b2s_bn_bwd_finalize_raw consumes it). Returns 1 (nothing launched) when the shape takes the one-tile kernel.
extern "C" int b2s_conv_dgrad_bnred(const void* dz, int dz_cstride, const void* w_dgrad_packed, void* dy, int dy_cstride,
                                    const void* r, int r_cstride, float* partial, int N, int H, int W, int Cin, int Cout,
                                    int ksize, int tile_n, void* stream) {
  return conv_fwd_impl(dz, dz_cstride, w_dgrad_packed, nullptr, nullptr, nullptr, dy, dy_cstride, partial, N, H, W, Cin,
                       Cout, ksize, B2S_FLAG_STATS | B2S_FLAG_BNRED, tile_n, stream, r, r_cstride);
}

// nn.Conv2d(Cin,Cout,3,stride=2,padding=1) forward (models/vnet.py:97): x [N,H,W,Cin] -> y [N,H/2,W/2,Cout].
extern "C" int b2s_conv3x3_s2_fwd(const void* x, int x_cstride, const void* w_packed, const float* bias, void* y,
                                  int y_cstride, int N, int H, int W, int Cin, int Cout, int tile_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w_packed || !y) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_fwd: null pointer");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_fwd: channels must be multiples of 64");
  if (N <= 0 || H < 2 || W < 2 || H % 2 || W % 2) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_fwd: H and W must be even");
  const int Ho = H / 2, Wo = W / 2;
  ConvTcParams p{};
  p.a_mode = A_CONV3_S2; p.out_mode = OUT_4D;
  ConvPlan pl;
  if (conv_plan(N, Ho, Wo, Cout, Cout, p.a_mode, p.out_mode, (tile_n & kTileNMask) | kVarPair, &pl))
    return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_fwd: tile_n must be 64/128/256 and divide Cout");
  p.W = Wo; p.H = Ho; p.N = N;
  p.num_taps = 9; p.k_chunks = Cin / 64;
  p.n_total = Cout; p.cout_sub = Cout;
  p.flags = 0; p.bias = bias; p.stats = nullptr;
  CUtensorMap tmA, tmB, tmOut;
  int rc;
  p.a_cstride = x_cstride;
  if ((rc = make_act_map5_parity(&tmA, x, Cin, W, H, N, x_cstride, pl.bw, pl.bh, pl.bn))) return rc;
  if ((rc = make_weight_map(&tmB, w_packed, Cin, 9 * Cout, pl.block_n))) return rc;
  if ((rc = make_act_map4(&tmOut, y, Cout, Wo, Ho, N, y_cstride, pl.bw, pl.bh, pl.bn))) return rc;
  return dispatch_conv(pl, tmA, tmB, tmOut, p, stream);
}

// Input gradient of nn.Conv2d(Cin,Cout,3,stride=2,padding=1): dz [N,H/2,W/2,Cout] -> dx [N,H,W,Cin], w_dgrad_packed
// [9][Cin][Cout] from b2s_pack_conv_weight. Four launches, one per output parity class (h%2, w%2) with 1/2/2/4 taps: no
// zero-insertion, no redundant FLOPs. Returns 1 (not an error) when the image is too small for the sub-lattice store;
// the caller then uses the zero-insertion path.
extern "C" int b2s_conv3x3_s2_dgrad(const void* dz, int dz_cstride, const void* w_dgrad_packed, void* dx, int dx_cstride,
                                    int N, int H, int W, int Cin, int Cout, int tile_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dz || !w_dgrad_packed || !dx) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_dgrad: null pointer");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_dgrad: channels must be multiples of 64");
  if (N <= 0 || H < 2 || W < 2 || H % 2 || W % 2) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_dgrad: H and W must be even");
  const int Ho = H / 2, Wo = W / 2;
  ConvTcParams p{};
  p.a_mode = A_TAPLIST; p.out_mode = OUT_SUB_5D;
  ConvPlan pl;
  if (conv_plan(N, Ho, Wo, Cin, Cin, p.a_mode, p.out_mode, (tile_n & kTileNMask) | kVarPair, &pl))
    return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_dgrad: tile_n must be 64/128/256 and divide Cin");
  if (pl.bn > 1 && pl.bh != Ho) return 1;   // tile would straddle images in the merged-row output view
  p.W = Wo; p.H = Ho; p.N = N;
  p.k_chunks = Cout / 64;
  p.n_total = Cin; p.cout_sub = Cin;
  p.flags = 0; p.bias = nullptr; p.stats = nullptr;
  CUtensorMap tmA, tmB, tmOut;
  int rc;
  if ((rc = make_act_map4(&tmA, dz, Cout, Wo, Ho, N, dz_cstride, pl.bw, pl.bh, pl.bn))) return rc;
  if ((rc = make_weight_map(&tmB, w_dgrad_packed, Cout, 9 * Cin, pl.block_n))) return rc;
  if ((rc = make_up_map5(&tmOut, dx, Cin, Wo, Ho, N, dx_cstride, pl.bw, pl.bh * pl.bn))) return rc;
  // dx[2i+ph, 2j+pw] = sum over (a, dh) in taps(ph), (b, dw) in taps(pw) of dz[i+dh, j+dw] * w[a, b];
  // taps(0) = {(1, 0)}, taps(1) = {(0, +1), (2, 0)}. The dgrad packing stores w[a, b] at tap row 8 - (3a + b).
  static const int ka[2][2] = {{1, -1}, {0, 2}}, kd[2][2] = {{0, 0}, {1, 0}}, kn[2] = {1, 2};
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      int nt = 0;
      for (int i = 0; i < kn[ph]; ++i)
        for (int j = 0; j < kn[pw]; ++j) {
          p.tap_dh[nt] = kd[ph][i]; p.tap_dw[nt] = kd[pw][j];
          p.tap_w[nt] = 8 - (3 * ka[ph][i] + ka[pw][j]);
          ++nt;
        }
      p.num_taps = nt; p.sub_a = ph; p.sub_b = pw;
      if ((rc = dispatch_conv(pl, tmA, tmB, tmOut, p, stream))) return rc;
    }
  return B2S_OK;
}

// Rows of the stats_partial buffer b2s_conv_fwd (ksize 3) writes for this shape: 2 per CTA row-group of the
// persistent grid of whichever kernel variant conv_plan picks.
extern "C" int b2s_conv_stats_rows(int N, int H, int W, int Cout, int tile_n) {
  ConvPlan pl;
  if (conv_plan(N, H, W, Cout, Cout, A_CONV3, OUT_4D, tile_n, &pl)) return -1;
  return pl.stats_rows;
}

// ConvTranspose2d(k=2,s=2) forward: x [N,Hi,Wi,Cin] -> y [N,2Hi,2Wi,Cout]; w_packed [(a*2+b)*Cout+co][Cin].
extern "C" int b2s_convt2x2_fwd(const void* x, int x_cstride, const void* w_packed, const float* bias, void* y,
                                int y_cstride, int N, int Hi, int Wi, int Cin, int Cout, int tile_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w_packed || !y) return set_error(B2S_ERR_ARG, "b2s_convt2x2_fwd: null pointer");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_convt2x2_fwd: channels must be multiples of 64");
  if (N <= 0 || Hi <= 0 || Wi <= 0) return set_error(B2S_ERR_ARG, "b2s_convt2x2_fwd: empty tensor");
  ConvTcParams p{};
  p.a_mode = A_1X1; p.out_mode = OUT_CONVT_5D;
  // no spatial halo in input space: rows of all images are merged into one dimension (Hm = Hi * N), which is also how
  // the 5-D view of the up-sampled output addresses them; works for any Hi, Wi
  const int Hm = Hi * N;
  ConvPlan pl;
  if (conv_plan(1, Hm, Wi, 4 * Cout, Cout, p.a_mode, p.out_mode, tile_n, &pl))
    return set_error(B2S_ERR_ARG, "b2s_convt2x2_fwd: tile_n must be 64/128/256 and divide Cout");
  p.W = Wi; p.H = Hm; p.N = 1;
  p.num_taps = 1; p.k_chunks = Cin / 64;
  p.n_total = 4 * Cout; p.cout_sub = Cout;
  p.flags = 0; p.bias = bias; p.stats = nullptr;
  CUtensorMap tmA, tmB, tmOut;
  int rc;
  if ((rc = make_act_map4(&tmA, x, Cin, Wi, Hm, 1, x_cstride, pl.bw, pl.bh, pl.bn))) return rc;
  if ((rc = make_weight_map(&tmB, w_packed, Cin, 4 * Cout, pl.block_n))) return rc;
  if ((rc = make_up_map5(&tmOut, y, Cout, Wi, Hm, 1, y_cstride, pl.bw, pl.bh * pl.bn))) return rc;
  return dispatch_conv(pl, tmA, tmB, tmOut, p, stream);
}

// ConvTranspose2d(k=2,s=2) input gradient: dy [N,2Hi,2Wi,Cout] -> dx [N,Hi,Wi,Cin]; w_packed [(a*2+b)*Cin+ci][Cout].
static int convt_dgrad_plan(int N, int Hi, int Wi, int Cin, int tile_n, ConvPlan* pl) {
  return conv_plan(1, Hi * N, Wi, Cin, Cin, A_CONVT_DGRAD, OUT_4D, tile_n, pl);
}

static int convt_dgrad_impl(const void* dy, int dy_cstride, const void* w_packed, void* dx, int dx_cstride, const void* red_r,
                            int red_cs, float* partial, int N, int Hi, int Wi, int Cin, int Cout, int tile_n,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dy || !w_packed || !dx) return set_error(B2S_ERR_ARG, "b2s_convt2x2_dgrad: null pointer");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_convt2x2_dgrad: channels must be multiples of 64");
  if (N <= 0 || Hi <= 0 || Wi <= 0) return set_error(B2S_ERR_ARG, "b2s_convt2x2_dgrad: empty tensor");
  ConvTcParams p{};
  p.a_mode = A_CONVT_DGRAD; p.out_mode = OUT_4D;
  const int Hm = Hi * N;   // image rows merged into one dimension (see b2s_convt2x2_fwd)
  ConvPlan pl;
  if (convt_dgrad_plan(N, Hi, Wi, Cin, tile_n, &pl))
    return set_error(B2S_ERR_ARG, "b2s_convt2x2_dgrad: tile_n must be 64/128/256 and divide Cin");
  p.W = Wi; p.H = Hm; p.N = 1;
  p.num_taps = 4; p.k_chunks = Cout / 64;
  p.n_total = Cin; p.cout_sub = Cin;
  p.flags = 0; p.bias = nullptr; p.stats = nullptr;
  if (red_r) {   // dx is the gradient of a train-mode BatchNorm output: fused {sum dx, sum dx * r} partials
    if (!partial || red_cs % 8) return set_error(B2S_ERR_ARG, "b2s_convt2x2_dgrad_bnred: partial missing / bad stride");
    if (pl.kind == CONV_LEGACY) return 1;
    p.flags = B2S_FLAG_STATS | B2S_FLAG_BNRED; p.stats = partial;
    p.red_r = static_cast<const __nv_bfloat16*>(red_r); p.red_cs = red_cs;
  }
  CUtensorMap tmA, tmB, tmOut;
  int rc;
  if ((rc = make_up_map5(&tmA, dy, Cout, Wi, Hm, 1, dy_cstride, pl.bw, pl.bh * pl.bn))) return rc;
  if ((rc = make_weight_map(&tmB, w_packed, Cout, 4 * Cin, pl.block_n))) return rc;
  if ((rc = make_act_map4(&tmOut, dx, Cin, Wi, Hm, 1, dx_cstride, pl.bw, pl.bh, pl.bn))) return rc;
  return dispatch_conv(pl, tmA, tmB, tmOut, p, stream);
}

extern "C" int b2s_convt2x2_dgrad(const void* dy, int dy_cstride, const void* w_packed, void* dx, int dx_cstride,
                                  int N, int Hi, int Wi, int Cin, int Cout, int tile_n, void* stream) {
  return convt_dgrad_impl(dy, dy_cstride, w_packed, dx, dx_cstride, nullptr, 0, nullptr, N, Hi, Wi, Cin, Cout, tile_n, stream);
}

// ... with the fused BatchNorm-backward reduction (see b2s_conv_dgrad_bnred); partial [b2s_convt2x2_dgrad_rows][2][Cin].
extern "C" int b2s_convt2x2_dgrad_bnred(const void* dy, int dy_cstride, const void* w_packed, void* dx, int dx_cstride,
                                        const void* r, int r_cstride, float* partial, int N, int Hi, int Wi, int Cin,
                                        int Cout, int tile_n, void* stream) {
  if (!r) return set_error(B2S_ERR_ARG, "b2s_convt2x2_dgrad_bnred: null pointer");
  return convt_dgrad_impl(dy, dy_cstride, w_packed, dx, dx_cstride, r, r_cstride, partial, N, Hi, Wi, Cin, Cout, tile_n, stream);
}

extern "C" int b2s_convt2x2_dgrad_rows(int N, int Hi, int Wi, int Cin, int tile_n) {
  ConvPlan pl;
  if (convt_dgrad_plan(N, Hi, Wi, Cin, tile_n, &pl)) return -1;
  return pl.stats_rows;
}

// ---- wgrad -----------------------------------------------------------------------------------
namespace b2s {
template <int BLOCK_N, int STAGES, int MT>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, int grid,
                        cudaStream_t stream) {
  using L = WgSmem<BLOCK_N, STAGES, MT>;
  auto kfn = wgrad_tc_kernel<BLOCK_N, STAGES, MT>;
  static std::atomic<unsigned long long> attr_devices{0};
  {
    cudaError_t e = allow_dynamic_smem(kfn, L::kDynBytes, attr_devices);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(wgrad_tc_kernel)");
  }
  kfn<<<grid, kNumThreads, L::kDynBytes, stream>>>(tmA, tmB, p);
  return check_launch("wgrad_tc_kernel");
}

static int dispatch_wgrad(int block_n, const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p,
                          cudaStream_t stream) {
  const int grid = p.m_tiles * p.tiles_nn * p.splits;
  count_launch();
  if (p.mt == 2) {
    switch (block_n) {
      case 128: return launch_wgrad<128, 4, 2>(tmA, tmB, p, grid, stream);   // 4 x 48 KB
      case 256: return launch_wgrad<256, 3, 2>(tmA, tmB, p, grid, stream);   // 3 x 64 KB, all 512 TMEM columns
    }
  } else {
    switch (block_n) {
      case 64:  return launch_wgrad<64, 6, 1>(tmA, tmB, p, grid, stream);    // 6 x 24 KB
      case 128: return launch_wgrad<128, 3, 1>(tmA, tmB, p, grid, stream);   // 3 x 32 KB
      case 256: return launch_wgrad<256, 4, 1>(tmA, tmB, p, grid, stream);   // 4 x 48 KB
    }
  }
  return set_error(B2S_ERR_ARG, "wgrad: unsupported tile_n");
}

static int wgrad_plan(int N, int H, int W, int Cin, int Cout, int num_taps, int tile_n, int splits_req,
                      WgradParams* p, int* block_n_out) {
  int block_n = tile_n ? tile_n : (Cout >= 256 ? 256 : Cout);
  if (block_n > Cout) block_n = Cout;
  if (Cout % block_n || (block_n != 64 && block_n != 128 && block_n != 256)) return -1;
  pick_box(W, H, N, 64, &p->pw, &p->ph, &p->pn);
  p->chunks_w = (W + p->pw - 1) / p->pw; p->chunks_h = (H + p->ph - 1) / p->ph; p->chunks_n = (N + p->pn - 1) / p->pn;
  p->H = H;
  p->num_taps = num_taps; p->ci_blocks = Cin / 64; p->total_rb = num_taps * p->ci_blocks;
  // two M tiles per dz stage when the row-block count allows it (transposed conv: every row block of a CTA must
  // belong to one tap, i.e. Cin % 256 == 0)
  p->mt = (block_n == 256 && p->total_rb >= 4 && (num_taps != 4 || p->ci_blocks % 4 == 0)) ? 2 : 1;   // 128: two CTAs/SM measured faster
  p->cin = Cin; p->cout = Cout; p->tiles_nn = Cout / block_n; p->m_tiles = (p->total_rb + 2 * p->mt - 1) / (2 * p->mt);
  p->chunks_total = p->chunks_w * p->chunks_h * p->chunks_n;
  int splits = splits_req;
  if (splits <= 0) {
    // keep >= 16 K-chunks per CTA (every split costs a K-sized fp32 partial) and
    // the grid must fit in ONE wave: CTAs resident per SM follow from the shared memory of the instantiation
    // (one tile per stage: block_n 64: 6 x 24 KB, 128: 3 x 32 KB, 256: 4 x 48 KB -> 1, 2, 1 CTAs per SM; tile pairs: 1)
    const int base = p->m_tiles * p->tiles_nn;
    const int resident = num_sms() * ((block_n == 128 && p->mt == 1) ? 2 : 1);
    splits = resident / base;
    const int max_by_work = p->chunks_total / 16 > 0 ? p->chunks_total / 16 : 1;
    if (splits > max_by_work) splits = max_by_work;
    if (splits < 1) splits = 1;
  }
  if (splits > p->chunks_total) splits = p->chunks_total;
  p->chunks_per_split = (p->chunks_total + splits - 1) / splits;
  p->splits = (p->chunks_total + p->chunks_per_split - 1) / p->chunks_per_split;
  *block_n_out = block_n;
  return 0;
}
}  // namespace b2s

// Workspace query: bytes needed and the number of K splits that b2s_conv_wgrad will use.
extern "C" long long b2s_conv_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int ksize_or_taps, int tile_n,
                                              int splits, int* splits_out) {
  WgradParams p{};
  int block_n;
  const int taps = ksize_or_taps == 3 ? 9 : ksize_or_taps;
  if (taps == 9 && !(tile_n & kVarLegacy) && wgrad_halo_eligible(N, H, W, Cin, Cout)) {
    const int s = wgrad_halo_splits(N, H, W, Cin, Cout, splits);
    if (splits_out) *splits_out = s;
    return static_cast<long long>(s) * taps * Cin * Cout * 4;
  }
  tile_n &= kTileNMask;
  if (taps == 4) { H *= N; N = 1; }   // the transposed-conv kernel merges the rows of all images (b2s_convt2x2_wgrad)
  if (Cin % 64 || Cout % 64 || wgrad_plan(N, H, W, Cin, Cout, taps, tile_n, splits, &p, &block_n)) {
    set_error(B2S_ERR_ARG, "b2s_conv_wgrad_workspace: unsupported shape");
    return -1;
  }
  if (splits_out) *splits_out = p.splits;
  return static_cast<long long>(p.splits) * taps * Cin * Cout * 4;
}

// 3x3 conv weight gradient partials: ws[split][tap*Cin+ci][co] = sum over the split's pixels.
extern "C" int b2s_conv3x3_wgrad(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H,
                                 int W, int Cin, int Cout, int tile_n, int splits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !dz || !ws) return set_error(B2S_ERR_ARG, "b2s_conv3x3_wgrad: null pointer");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_conv3x3_wgrad: channels must be multiples of 64");
  if (!(tile_n & kVarLegacy) && wgrad_halo_eligible(N, H, W, Cin, Cout))
    return launch_wgrad_halo(x, x_cstride, dz, dz_cstride, ws, N, H, W, Cin, Cout, splits, stream);
  tile_n &= kTileNMask;
  WgradParams p{};
  int block_n;
  if (wgrad_plan(N, H, W, Cin, Cout, 9, tile_n, splits, &p, &block_n))
    return set_error(B2S_ERR_ARG, "b2s_conv3x3_wgrad: bad tile_n");
  p.mode = WG_CONV3; p.ws = ws;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_act_map4(&tmA, x, Cin, W, H, N, x_cstride, p.pw, p.ph, p.pn))) return rc;
  if ((rc = make_act_map4(&tmB, dz, Cout, W, H, N, dz_cstride, p.pw, p.ph, p.pn))) return rc;
  return dispatch_wgrad(block_n, tmA, tmB, p, stream);
}

// Weight gradient of nn.Conv2d(Cin,Cout,3,stride=2,padding=1): x [N,H,W,Cin], dz [N,H/2,W/2,Cout] -> ws[split][tap*Cin+ci][co];
// workspace from b2s_conv_wgrad_workspace(N, H/2, W/2, Cin, Cout, 3, tile_n | 4096, ...) (the one-box-per-tap kernel).
extern "C" int b2s_conv3x3_s2_wgrad(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H,
                                    int W, int Cin, int Cout, int tile_n, int splits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !dz || !ws) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_wgrad: null pointer");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_wgrad: channels must be multiples of 64");
  if (N <= 0 || H < 2 || W < 2 || H % 2 || W % 2) return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_wgrad: H and W must be even");
  const int Ho = H / 2, Wo = W / 2;
  tile_n &= kTileNMask;
  WgradParams p{};
  int block_n;
  if (wgrad_plan(N, Ho, Wo, Cin, Cout, 9, tile_n, splits, &p, &block_n))
    return set_error(B2S_ERR_ARG, "b2s_conv3x3_s2_wgrad: bad tile_n");
  p.mode = WG_CONV3_S2; p.ws = ws; p.x_cstride = x_cstride;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_act_map5_parity(&tmA, x, Cin, W, H, N, x_cstride, p.pw, p.ph, p.pn))) return rc;
  if ((rc = make_act_map4(&tmB, dz, Cout, Wo, Ho, N, dz_cstride, p.pw, p.ph, p.pn))) return rc;
  return dispatch_wgrad(block_n, tmA, tmB, p, stream);
}

// 1x1 conv weight gradient partials (residual projections of models/vnet.py:46): ws[split][ci][co].
extern "C" int b2s_conv1x1_wgrad(const void* x, int x_cstride, const void* dz, int dz_cstride, float* ws, int N, int H,
                                 int W, int Cin, int Cout, int tile_n, int splits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !dz || !ws) return set_error(B2S_ERR_ARG, "b2s_conv1x1_wgrad: null pointer");
  if (Cin % 64 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_conv1x1_wgrad: channels must be multiples of 64");
  tile_n &= kTileNMask;
  WgradParams p{};
  int block_n;
  if (wgrad_plan(N, H, W, Cin, Cout, 1, tile_n, splits, &p, &block_n))
    return set_error(B2S_ERR_ARG, "b2s_conv1x1_wgrad: bad tile_n");
  p.mode = WG_1X1; p.ws = ws;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_act_map4(&tmA, x, Cin, W, H, N, x_cstride, p.pw, p.ph, p.pn))) return rc;
  if ((rc = make_act_map4(&tmB, dz, Cout, W, H, N, dz_cstride, p.pw, p.ph, p.pn))) return rc;
  return dispatch_wgrad(block_n, tmA, tmB, p, stream);
}

// ConvTranspose2d(k2,s2) weight gradient partials: ws[split][(a*2+b)*Cin+ci][co]; x [N,Hi,Wi,Cin], dy [N,2Hi,2Wi,Cout].
extern "C" int b2s_convt2x2_wgrad(const void* x, int x_cstride, const void* dy, int dy_cstride, float* ws, int N,
                                  int Hi, int Wi, int Cin, int Cout, int tile_n, int splits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !dy || !ws) return set_error(B2S_ERR_ARG, "b2s_convt2x2_wgrad: null pointer");
  if (Cin % 128 || Cout % 64) return set_error(B2S_ERR_ARG, "b2s_convt2x2_wgrad: need Cin % 128 == 0, Cout % 64 == 0");
  WgradParams p{};
  int block_n;
  const int Hm = Hi * N;   // image rows merged into one dimension (no spatial halo; see b2s_convt2x2_fwd)
  if (wgrad_plan(1, Hm, Wi, Cin, Cout, 4, tile_n, splits, &p, &block_n))
    return set_error(B2S_ERR_ARG, "b2s_convt2x2_wgrad: bad tile_n");
  p.mode = WG_CONVT; p.ws = ws;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_act_map4(&tmA, x, Cin, Wi, Hm, 1, x_cstride, p.pw, p.ph, p.pn))) return rc;
  if ((rc = make_up_map5(&tmB, dy, Cout, Wi, Hm, 1, dy_cstride, p.pw, p.ph * p.pn))) return rc;
  return dispatch_wgrad(block_n, tmA, tmB, p, stream);
}
