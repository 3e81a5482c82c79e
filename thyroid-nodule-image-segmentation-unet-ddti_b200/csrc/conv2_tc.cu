// Tile-pair and row-halo implicit-GEMM convolution kernels (tcgen05 + TMEM + TMA, sm_100a).
//
// Both kernels raise the FLOPs done per byte moved L2 -> shared memory, which is what bounds conv_tc_kernel
// (one 128-pixel tile per weight stage: 64 FLOP/B, ~45 B/clk/SM of L2 bandwidth = ~45 % of the tensor pipe):
//
//   PAIR  (HALO=false): every weight (B) stage is multiplied with TWO 128-pixel A tiles into two TMEM accumulators
//                       (an M=256 tile per CTA): 87 FLOP/B at BLOCK_N=128, 131 FLOP/B at BLOCK_N=256.
//   HALO  (HALO=true) : for images at least 128 pixels wide. A work item is two output rows x 128 pixels. The four
//                       input rows it needs (130 pixels each, zero-filled by TMA outside the image) are staged ONCE
//                       per 64-channel chunk; the nine filter taps are nine UMMA descriptors whose start address is
//                       shifted by whole 128-byte pixel rows inside that staging area. Activation traffic drops
//                       from 9x to 2x, weight traffic is shared by the two rows.
//
// Replaces the same reference call sites as conv_tc_kernel: nn.Conv2d(k=3,p=1) forward / input gradient and the
// 1x1 / transposed-conv GEMMs of models/model.py:36,39,49.
//
// Warp roles (352 threads): warp0 = TMEM alloc + halo producer, warp1 = MMA issuer, warp2 = ring producer,
// warps 3-6 / 7-10 = two epilogue groups (TMEM -> +bias, ReLU -> bf16 -> swizzled smem -> TMA store; BN partial sums).
#include "conv_common.cuh"

namespace b2s {

constexpr int kConv2Threads = 352;
constexpr int kHaloPix = 130;               // 128 output pixels + one halo pixel each side
constexpr int kHaloRowBytes = 17 * 1024;    // 130 x 128 B = 16640 B, padded so every row slot stays 1024-B aligned

template <int BLOCK_N, int MT, bool HALO, int STAGES, int STAGES_A>
struct Conv2Cfg {
  static constexpr int kSets = (2 * MT * BLOCK_N <= 512) ? 2 : 1;   // accumulator sets (double-buffered epilogue)
  static constexpr int kTmemCols = kSets * MT * BLOCK_N;
  static constexpr int kBTile = BLOCK_N * 128;
  static constexpr int kAInStage = HALO ? 0 : MT * kATileBytes;
  static constexpr int kStageBytes = kAInStage + kBTile;
  static constexpr int kHaloStage = (MT + 2) * kHaloRowBytes;
  static constexpr int kHaloOffset = STAGES * kStageBytes;
  static constexpr int kStagingOffset = kHaloOffset + (HALO ? STAGES_A * kHaloStage : 0);
  static constexpr int kBarOffset = kStagingOffset + 2 * kATileBytes;
  static constexpr int kNumBars = 2 * STAGES + 2 * STAGES_A + 2 * kSets;
  static constexpr int kTmemPtrOffset = kBarOffset + kNumBars * 8;
  static constexpr int kBiasOffset = kTmemPtrOffset + 16;            // BLOCK_N floats: the CTA's bias columns
  static constexpr int kTotal = kBiasOffset + BLOCK_N * 4;
  static constexpr int kDynBytes = kTotal + 1024;  // slack for manual 1024-B alignment
  static_assert(kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns");
  static_assert(kDynBytes <= 227 * 1024, "exceeds the 227 KB of shared memory a CTA may use");
  static_assert(MT == 1 || MT == 2, "one or two M tiles per work item");
};

// Epilogue conversion of one 64-column sub-tile row: accumulators (+bias) (ReLU) (eval-mode BatchNorm affine) -> 32
// packed bf16 pairs. The three options are compile-time: with them as run-time tests inside the unrolled loop the
// sub-tile cost 737 issued instructions per warp (ncu source page: 100 LDC, 96 predicated LDG, 142 IMAD, 64 FSEL ...),
// and one warp owns 32 rows, so a GEMM with a short K loop (the transposed-conv forward: 2 k-iterations per item) was
// bound by its epilogue warps' issue rate (3 600 clk per item with or without the stores, HBM- or L2-resident alike).
// Same arithmetic in the same order as before: bit-identical results.
template <bool BIAS, bool RELU, bool POST>
__device__ __forceinline__ void epi_convert(const uint32_t (&v0)[32], const uint32_t (&v1)[32], uint32_t (&packed)[32],
                                            const float* bias_s, const float* __restrict__ post_scale,
                                            const float* __restrict__ post_shift) {
#pragma unroll
  for (int j2 = 0; j2 < 16; ++j2) {
    float a0 = __uint_as_float(j2 < 8 ? v0[4 * j2] : v1[4 * j2 - 32]);
    float b0 = __uint_as_float(j2 < 8 ? v0[4 * j2 + 1] : v1[4 * j2 - 31]);
    float a1 = __uint_as_float(j2 < 8 ? v0[4 * j2 + 2] : v1[4 * j2 - 30]);
    float b1 = __uint_as_float(j2 < 8 ? v0[4 * j2 + 3] : v1[4 * j2 - 29]);
    if (BIAS) {
      const float4 bb = *reinterpret_cast<const float4*>(bias_s + 4 * j2);   // shared memory, broadcast
      f2_unpack(f2_add(f2_pack(a0, b0), f2_pack(bb.x, bb.y)), a0, b0);
      f2_unpack(f2_add(f2_pack(a1, b1), f2_pack(bb.z, bb.w)), a1, b1);
    }
    if (RELU) { a0 = fmaxf(a0, 0.f); b0 = fmaxf(b0, 0.f); a1 = fmaxf(a1, 0.f); b1 = fmaxf(b1, 0.f); }
    if (POST) {
      const float4 ps = __ldg(reinterpret_cast<const float4*>(post_scale) + j2);
      const float4 pt = __ldg(reinterpret_cast<const float4*>(post_shift) + j2);
      a0 = fmaf(a0, ps.x, pt.x); b0 = fmaf(b0, ps.y, pt.y); a1 = fmaf(a1, ps.z, pt.z); b1 = fmaf(b1, ps.w, pt.w);
    }
    packed[2 * j2] = pack_bf16x2(a0, b0);
    packed[2 * j2 + 1] = pack_bf16x2(a1, b1);
  }
}

// Row-shifted operands: the tensor core applies the SWIZZLE_128B XOR to the absolute shared-memory address bits, the
// same function TMA used when it wrote the 1024-B aligned row slots, so a descriptor may start at any 128-byte pixel
// row of a slot with base_offset = 0 (measured on B200: base_offset = (addr >> 7) & 7 gives wrong results).
template <int BLOCK_N, int MT, bool HALO, int STAGES, int STAGES_A, bool RED>
__global__ void __launch_bounds__(kConv2Threads, 1)
conv2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmOut, const ConvTcParams p) {
  using L = Conv2Cfg<BLOCK_N, MT, HALO, STAGES, STAGES_A>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* halo = smem + L::kHaloOffset;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* afull_bar = empty_bar + STAGES;
  uint64_t* aempty_bar = afull_bar + STAGES_A;
  uint64_t* tmem_full_bar = aempty_bar + STAGES_A;     // [kSets]
  uint64_t* tmem_empty_bar = tmem_full_bar + L::kSets;  // [kSets]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- static schedule: CTA c owns column tile c % tiles_nn and the work items mgroup + i * num_mgroups -------
  const int n_tile = blockIdx.x % p.tiles_nn;
  const int mgroup = blockIdx.x / p.tiles_nn;
  const int num_mgroups = gridDim.x / p.tiles_nn;
  const int ncol0 = n_tile * BLOCK_N;
  const int my_items = (p.items_m - mgroup + num_mgroups - 1) / num_mgroups;
  const int k_iters = p.num_taps * p.k_chunks;
  const int strips = p.tiles_w;          // HALO: 128-pixel strips per row
  const int hpairs = p.H >> 1;           // HALO: row pairs per image

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      tma_prefetch_desc(&tmOut);
    }
    tmem_alloc(tmem_ptr_smem, L::kTmemCols);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < STAGES_A; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int s = 0; s < L::kSets; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], MT == 2 ? 8 : 4);   // one arrival per epilogue warp that drains the set
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (HALO && lane == 0) {
      // ===== halo producer: (MT + 2) input rows x 130 pixels per (work item, 64-channel chunk) =====
      int sa = 0;
      uint32_t pa = 0;
      for (int i = 0; i < my_items; ++i) {
        const int item = mgroup + i * num_mgroups;
        const int ws = item % strips;
        const int hp = (item / strips) % hpairs;
        const int n = item / (strips * hpairs);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&aempty_bar[sa], pa ^ 1);
          mbar_arrive_expect_tx(&afull_bar[sa], (MT + 2) * kHaloPix * 128);
          uint8_t* dst = halo + sa * L::kHaloStage;
#pragma unroll
          for (int r = 0; r < MT + 2; ++r)
            tma_load_4d(&tmA, &afull_bar[sa], dst + r * kHaloRowBytes, kc * kBlockK, ws * 128 - 1, 2 * hp - 1 + r, n);
          if (++sa == STAGES_A) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ===== ring producer: weight tiles (HALO) or MT activation tiles + one weight tile per stage =====
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_items; ++i) {
        const int item = mgroup + i * num_mgroups;
        if (HALO) {
          for (int kc = 0; kc < p.k_chunks; ++kc)
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], L::kBTile);
              tma_load_2d(&tmB, &full_bar[stage], ring + stage * L::kStageBytes, kc * kBlockK, tap * p.n_total + ncol0);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        } else {
          int w0[MT], h0[MT], n0[MT];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int m_tile = item * MT + mt;
            w0[mt] = (m_tile % p.tiles_w) * p.bw;
            h0[mt] = ((m_tile / p.tiles_w) % p.tiles_h) * p.bh;
            n0[mt] = (m_tile / (p.tiles_w * p.tiles_h)) * p.bn;
          }
          for (int it = 0; it < k_iters; ++it) {
            const int tap = it / p.k_chunks;
            const int kc = it - tap * p.k_chunks;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            uint8_t* st = ring + stage * L::kStageBytes;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint8_t* a_dst = st + mt * kATileBytes;
              if (p.a_mode == A_CONV3) {
                const int dh = tap / 3 - 1, dw = tap - (tap / 3) * 3 - 1;
                tma_load_4d(&tmA, &full_bar[stage], a_dst, kc * kBlockK, w0[mt] + dw, h0[mt] + dh, n0[mt]);
              } else if (p.a_mode == A_CONV3_S2) {   // stride 2: one parity class of the input per tap (make_act_map5_parity)
                const int dh = tap / 3 - 1, dw = tap - (tap / 3) * 3 - 1;   // input pixel = 2 * output pixel + (dh, dw)
                const int ph = dh & 1, pw = dw & 1;
                tma_load_5d(&tmA, &full_bar[stage], a_dst, pw * p.a_cstride + kc * kBlockK, w0[mt] + ((dw - pw) >> 1), ph,
                            h0[mt] + ((dh - ph) >> 1), n0[mt]);
              } else if (p.a_mode == A_TAPLIST) {    // explicit tap offsets (stride-2 conv input gradient, one parity class)
                tma_load_4d(&tmA, &full_bar[stage], a_dst, kc * kBlockK, w0[mt] + p.tap_dw[tap], h0[mt] + p.tap_dh[tap],
                            n0[mt]);
              } else if (p.a_mode == A_1X1) {
                tma_load_4d(&tmA, &full_bar[stage], a_dst, kc * kBlockK, w0[mt], h0[mt], n0[mt]);
              } else {  // A_CONVT_DGRAD: dY viewed as (C, b, j, a, i*N)
                tma_load_5d(&tmA, &full_bar[stage], a_dst, kc * kBlockK, tap & 1, w0[mt], tap >> 1,
                            n0[mt] * p.H + h0[mt]);
              }
            }
            tma_load_2d(&tmB, &full_bar[stage], st + L::kAInStage, kc * kBlockK,
                        (p.a_mode == A_TAPLIST ? p.tap_w[tap] : tap) * p.n_total + ncol0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the pipeline so that every address stays in uniform registers; one
    // elected lane issues tcgen05.mma / tcgen05.commit. Descriptors are advanced by adding to their low word. =====
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
    const uint32_t tbase = __reduce_or_sync(0xffffffffu, tmem_base);   // warp-uniform copy
    const bool leader = elect_one();
    const uint32_t desc_hi = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, 1024) >> 32);
    const uint32_t desc_lo0 = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, 1024));   // LBO field, address 0
    const uint32_t ring_lo = desc_lo0 + (smem_u32(ring) >> 4);
    const uint32_t halo_lo = desc_lo0 + (smem_u32(halo) >> 4);
    int stage = 0, sa = 0;
    uint32_t phase = 0, pa = 0;
    for (int i = 0; i < my_items; ++i) {
      const int set = i % L::kSets;
      mbar_wait(&tmem_empty_bar[set], ((i / L::kSets) & 1) ^ 1);   // epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t acc0 = tbase + set * MT * BLOCK_N;
      if (HALO) {
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&afull_bar[sa], pa);
          const uint32_t a_lo = halo_lo + sa * (L::kHaloStage >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            constexpr int kRow16 = kHaloRowBytes >> 4;
            const int th = tap / 3, tw = tap - th * 3;   // input row offset (dh + 1), pixel offset (dw + 1)
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t b_lo = ring_lo + stage * (L::kStageBytes >> 4);
            if (leader) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_bf16_lohi(acc0 + mt * BLOCK_N, a_lo + (mt + th) * kRow16 + tw * 8 + k * 2, desc_hi, b_lo + k * 2,
                                 desc_hi, idesc, (tap | k) != 0 ? 1u : static_cast<uint32_t>(kc != 0));
              }
              umma_commit(&empty_bar[stage]);
              if (tap == 8) umma_commit(&aempty_bar[sa]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++sa == STAGES_A) { sa = 0; pa ^= 1; }
        }
      } else {
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = ring_lo + stage * (L::kStageBytes >> 4);
          const uint32_t b_lo = a_lo + (L::kAInStage >> 4);
          if (leader) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_bf16_lohi(acc0 + mt * BLOCK_N, a_lo + mt * (kATileBytes >> 4) + k * 2, desc_hi, b_lo + k * 2, desc_hi,
                               idesc, k != 0 ? 1u : static_cast<uint32_t>(it != 0));
            }
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (leader) umma_commit(&tmem_full_bar[set]);
      __syncwarp();
    }
  } else {
    // ===== epilogue groups =====
    const int g = (warp - 3) >> 2;   // group 0: warps 3-6, group 1: warps 7-10
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;   // tile row == TMEM lane
    const bool do_relu = p.flags & B2S_FLAG_RELU;
    const bool do_stats = (p.flags & B2S_FLAG_STATS) && p.stats != nullptr;
    // BatchNorm-backward reduction fused into an input-gradient launch: sum(dy) and sum(dy * r) per channel, where r
    // is the saved input of the BatchNorm that dy (this launch's output) is the gradient of
    // (RED is a template parameter: the forward / plain dgrad instantiations keep their register allocation)
    const bool do_red = RED && do_stats;
    const int sh_w = __ffs(p.bw) - 1, sh_h = __ffs(p.bh) - 1;   // pick_box: bw, bh, bn are powers of two
    uint8_t* stage_buf = smem + L::kStagingOffset + g * kATileBytes;
    const int bar_id = 1 + g;
    const int wl = row % p.bw;
    const int hl = (row / p.bw) % p.bh;
    const int nl = row / (p.bw * p.bh);
    float st_acc[BLOCK_N / 64][4];
#pragma unroll
    for (int s = 0; s < BLOCK_N / 64; ++s) st_acc[s][0] = st_acc[s][1] = st_acc[s][2] = st_acc[s][3] = 0.f;
    // this CTA's bias columns, once, in shared memory (a column's bias index is its position inside the (a,b)
    // sub-block for the transposed-conv forward); both epilogue groups fill and then meet on named barrier 3
    float* bias_smem = reinterpret_cast<float*>(smem + L::kBiasOffset);
    const int emode = (p.bias != nullptr ? 1 : 0) | (do_relu ? 2 : 0) | (p.post_scale != nullptr ? 4 : 0);
    if (p.bias != nullptr) {
      for (int c = threadIdx.x - 96; c < BLOCK_N; c += 256) bias_smem[c] = __ldg(p.bias + (ncol0 + c) % p.cout_sub);
      asm volatile("bar.sync 3, 256;" ::: "memory");
    }

    for (int i = (MT == 2 ? 0 : g); i < my_items; i += (MT == 2 ? 1 : 2)) {
      const int item = mgroup + i * num_mgroups;
      const int set = MT == 2 ? i % L::kSets : g;
      const uint32_t par = MT == 2 ? (i / L::kSets) & 1 : (i >> 1) & 1;
      int w0, h0, n0;
      bool valid;
      if (HALO) {
        w0 = (item % strips) * 128;
        h0 = 2 * ((item / strips) % hpairs) + g;
        n0 = item / (strips * hpairs);
        valid = true;
      } else {
        const int m_tile = MT == 2 ? item * 2 + g : item;
        w0 = (m_tile % p.tiles_w) * p.bw;
        h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.bh;
        n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.bn;
        valid = (w0 + wl < p.W) && (h0 + hl < p.H) && (n0 + nl < p.N);
      }
      if (RED && do_red) {
        // pull this tile's r rows (one pixel row of BLOCK_N channels per lane) into L2 while the MMAs of the item are
        // still running: the 32-bit loads of the column sums below then pay L2 instead of HBM latency
        const int rr = q * 32 + lane;
        const int rw_l = rr & (p.bw - 1), rh_l = (rr >> sh_w) & (p.bh - 1), rn_l = rr >> (sh_w + sh_h);
        if (HALO || ((w0 + rw_l < p.W) && (h0 + rh_l < p.H) && (n0 + rn_l < p.N))) {
          const long long pix = (static_cast<long long>(n0 + rn_l) * p.H + (h0 + rh_l)) * p.W + (w0 + rw_l);
          const __nv_bfloat16* line = p.red_r + pix * p.red_cs + ncol0;
#pragma unroll
          for (int s = 0; s < BLOCK_N / 64; ++s)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(line + s * 64));
        }
      }
      mbar_wait(&tmem_full_bar[set], par);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (set * MT + (MT == 2 ? g : 0)) * BLOCK_N;

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
          if (lane == 0) mbar_arrive(&tmem_empty_bar[set]);
        }
        const int col_base = ncol0 + s * 64;
        const int bias_base = col_base % p.cout_sub;
        uint32_t packed[32];
        {
          const float* bs = bias_smem + s * 64;
          const float* ps = p.post_scale + bias_base;
          const float* pt = p.post_shift + bias_base;
          switch (emode) {
            case 0: epi_convert<false, false, false>(v0, v1, packed, bs, ps, pt); break;
            case 1: epi_convert<true, false, false>(v0, v1, packed, bs, ps, pt); break;
            case 2: epi_convert<false, true, false>(v0, v1, packed, bs, ps, pt); break;
            case 3: epi_convert<true, true, false>(v0, v1, packed, bs, ps, pt); break;
            case 4: epi_convert<false, false, true>(v0, v1, packed, bs, ps, pt); break;
            case 5: epi_convert<true, false, true>(v0, v1, packed, bs, ps, pt); break;
            case 6: epi_convert<false, true, true>(v0, v1, packed, bs, ps, pt); break;
            default: epi_convert<true, true, true>(v0, v1, packed, bs, ps, pt); break;
          }
          if (!HALO && !valid) {      // rows of a ragged tile that lie outside the image store zeros
#pragma unroll
            for (int j = 0; j < 32; ++j) packed[j] = 0u;
          }
        }
        // r words of this warp's 32 rows for the lane's column pair (one coalesced 128-byte row per load), issued now
        // so that their L2 latency overlaps the staging write below
        uint32_t rw[RED ? 32 : 1];
        if (RED && do_red) {
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int rr = q * 32 + r;
            const int rw_l = rr & (p.bw - 1), rh_l = (rr >> sh_w) & (p.bh - 1), rn_l = rr >> (sh_w + sh_h);
            const bool rv = HALO || ((w0 + rw_l < p.W) && (h0 + rh_l < p.H) && (n0 + rn_l < p.N));
            const long long pix = (static_cast<long long>(n0 + rn_l) * p.H + (h0 + rh_l)) * p.W + (w0 + rw_l);
            rw[r] = rv ? __ldg(reinterpret_cast<const unsigned int*>(p.red_r + pix * p.red_cs + col_base) + lane) : 0u;
          }
        }
        // the previous TMA store of this group must have finished READING the staging tile
        if (row == 0) tma_store_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 val = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
          *reinterpret_cast<uint4*>(stage_buf + row * 128 + ((c ^ (row & 7)) << 4)) = val;
        }
        if (do_stats) {
          // column sums over this warp's own 32 rows, taken from the bf16-rounded values actually stored
          __syncwarp();
          // packed fp32 pairs (FADD2 / FFMA2: one issue slot for both columns, bit-identical per lane)
          f32x2 s01 = f2_pack(0.f, 0.f), q01 = f2_pack(0.f, 0.f);
          const int chunk = lane >> 2, within = (lane & 3) * 4;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int rr = q * 32 + r;
            const uint32_t u =
                *reinterpret_cast<const uint32_t*>(stage_buf + rr * 128 + ((chunk ^ (rr & 7)) << 4) + within);
            const f32x2 x01 = f2_pack(bf16_lo(u), bf16_hi(u));
            s01 = f2_add(s01, x01);
            q01 = f2_fma(x01, RED ? f2_pack(bf16_lo(rw[RED ? r : 0]), bf16_hi(rw[RED ? r : 0])) : x01, q01);
          }
          float s0, s1, q0, q1;
          f2_unpack(s01, s0, s1);
          f2_unpack(q01, q0, q1);
          st_acc[s][0] += s0; st_acc[s][1] += s1; st_acc[s][2] += q0; st_acc[s][3] += q1;
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (row == 0) {
          if (p.out_mode == OUT_4D) {
            tma_store_4d(&tmOut, stage_buf, col_base, w0, h0, n0);
          } else if (p.out_mode == OUT_SUB_5D) {   // one fixed sub-lattice of the 2x larger output
            tma_store_5d(&tmOut, stage_buf, col_base, p.sub_b, w0, p.sub_a, n0 * p.H + h0);
          } else {  // convT forward: the 64-column block belongs to one (a,b) sub-position
            const int ab = col_base / p.cout_sub;
            tma_store_5d(&tmOut, stage_buf, col_base - ab * p.cout_sub, ab & 1, w0, ab >> 1, n0 * p.H + h0);
          }
          tma_store_commit();
        }
      }
    }
    if (row == 0) tma_store_wait_read<0>();
    if (do_stats) {
      // one partial row per (CTA row-group, epilogue group): [2*mgroup + g][2][n_total]; scratch aliases the
      // group's staging tile (its last TMA store has been read out above)
      float* sb = reinterpret_cast<float*>(stage_buf);
      float* out_row = p.stats + static_cast<size_t>(2 * mgroup + g) * 2 * p.n_total;
#pragma unroll
      for (int s = 0; s < BLOCK_N / 64; ++s) {
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        sb[q * 128 + 2 * lane] = st_acc[s][0]; sb[q * 128 + 2 * lane + 1] = st_acc[s][1];
        sb[q * 128 + 64 + 2 * lane] = st_acc[s][2]; sb[q * 128 + 64 + 2 * lane + 1] = st_acc[s][3];
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        const int which = row >> 6, col = row & 63;
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
    tmem_dealloc(tmem_base, L::kTmemCols);
  }
}

template <int BLOCK_N, int MT, bool HALO, int STAGES, int STAGES_A, bool RED>
static int launch_conv2_r(int grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                          const ConvTcParams& p, cudaStream_t stream) {
  using L = Conv2Cfg<BLOCK_N, MT, HALO, STAGES, STAGES_A>;
  auto kfn = conv2_tc_kernel<BLOCK_N, MT, HALO, STAGES, STAGES_A, RED>;
  static std::atomic<unsigned long long> attr_devices{0};
  {
    cudaError_t e = allow_dynamic_smem(kfn, L::kDynBytes, attr_devices);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv2_tc_kernel)");
  }
  kfn<<<grid, kConv2Threads, L::kDynBytes, stream>>>(tmA, tmB, tmOut, p);
  return check_launch("conv2_tc_kernel");
}

// B2S_FLAG_BNRED selects the instantiation with the fused BatchNorm-backward reduction
template <int BLOCK_N, int MT, bool HALO, int STAGES, int STAGES_A>
static int launch_conv2_t(int grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                          const ConvTcParams& p, cudaStream_t stream) {
  if ((p.flags & B2S_FLAG_BNRED) && p.red_r != nullptr)
    return launch_conv2_r<BLOCK_N, MT, HALO, STAGES, STAGES_A, true>(grid, tmA, tmB, tmOut, p, stream);
  return launch_conv2_r<BLOCK_N, MT, HALO, STAGES, STAGES_A, false>(grid, tmA, tmB, tmOut, p, stream);
}

int launch_conv2(const ConvPlan& pl, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                 const ConvTcParams& p, cudaStream_t stream) {
  count_launch();
  if (pl.kind == CONV_HALO) {
    switch (pl.block_n) {
      case 64:  return launch_conv2_t<64, 2, true, 6, 2>(pl.grid, tmA, tmB, tmOut, p, stream);
      case 128: return launch_conv2_t<128, 2, true, 3, 2>(pl.grid, tmA, tmB, tmOut, p, stream);
    }
  } else if (pl.kind == CONV_PAIR) {
    switch (pl.block_n) {
      case 64:  return launch_conv2_t<64, 2, false, 4, 0>(pl.grid, tmA, tmB, tmOut, p, stream);
      case 128: return launch_conv2_t<128, 2, false, 4, 0>(pl.grid, tmA, tmB, tmOut, p, stream);
      case 256: return launch_conv2_t<256, 2, false, 3, 0>(pl.grid, tmA, tmB, tmOut, p, stream);
    }
  }
  return set_error(B2S_ERR_ARG, "launch_conv2: unsupported variant / tile_n");
}

}  // namespace b2s
