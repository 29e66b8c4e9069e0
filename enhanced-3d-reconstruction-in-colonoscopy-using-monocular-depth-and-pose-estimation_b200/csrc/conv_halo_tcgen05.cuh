// 3x3 / pad-1 convolution as an implicit GEMM with HALO REUSE, 2-CTA tcgen05 (cta_group::2).
//
// The plain implicit-GEMM path loads one shifted 128-pixel A box per (tap, channel block): every input
// element crosses L2->smem 9 times, and with N <= 256 every A box is private to one CTA, so the DPT
// convs were bound by the ~7-8 TB/s L2->SM fabric (ncu: 64-96 B/clk/SM demanded), not by the tensor pipe.
// Here each CTA loads, per 64-channel block, ONE halo patch of (16+2) x (8+2) pixels and issues the MMAs of
// all 9 taps from it: tap (dy,dx) is the same smem patch read through a UMMA descriptor whose start
// address is shifted by (dy*10 + dx)*128 bytes and whose 8-row groups are one 10-pixel line (1280 B)
// apart.  Neither the start nor the group stride is a multiple of the 1024 B swizzle atom: that is fine
// because both TMA (writing) and the MMA (reading) derive the 128B-swizzle phase from the absolute
// shared-memory address bits, with every stage base 1024 B-aligned (the first version padded lines to
// 16 pixels to keep all groups phase-aligned and moved 60 % more patch bytes for it; parity tests are
// identical for both).  A traffic drops 9 x 16 KB -> 22.5 KB per channel block.
// Weights (B) stream through their own ring, one [BN/2 x 64] tile per tap -- except for the N = 32 depth head, whose
// whole weight set (<= 18 tiles of 2 KB per CTA) is loaded ONCE per CTA and stays resident: with the ring that kernel
// moved 15.2 GB through the L2->SM crossbar per launch (10.1 GB of 16-wide halo patches + 5.0 GB of re-streamed weights)
// at the fabric's ~7.9 TB/s, i.e. it was crossbar-bound (profiles/ncu_conv_r02.txt).
//
// Tile = 16 rows x 8 columns of output pixels (M = 128 per CTA, 256 per CTA pair; N = 128 runs two such
// tiles per CTA against every weight tile); roles, TMEM double buffering, 8 epilogue warps and the
// fused epilogue are those of gemm2_tcgen05_kernel.
#pragma once

#include "gemm2_tcgen05.cuh"

#ifndef HALO_A_STAGES
#define HALO_A_STAGES 2  // halo patches in flight per CTA (A/B-measured: see DESIGN.md section 5)
#endif
#ifndef HALO_MT128
#define HALO_MT128 2     // N = 128: pixel tiles per CTA that share every weight tile (1 = the one-tile form, for A/B builds)
#endif

namespace dav2 {

template <int BN>
struct ConvHaloCfg {
  static constexpr int TW = 8, TH = 16;                 // output tile (pixels)
#ifdef HALO_PATCH_W
  static constexpr int HALO_W = HALO_PATCH_W, HALO_H = TH + 2;  // A/B builds only (16 = the round-1 line pitch)
#else
  static constexpr int HALO_W = TW + 2, HALO_H = TH + 2; // loaded patch: x0-1 .. x0+8, y0-1 .. y0+16
#endif
  static constexpr int A_TX_BYTES = HALO_H * HALO_W * 128; // 23040: [18][10] pixels x 64 channels x 2 B (what TMA delivers)
  static constexpr int A_BYTES = (A_TX_BYTES + 1023) / 1024 * 1024;  // stage stride: keeps every stage 1024 B-aligned
  static constexpr int A_STAGES = HALO_A_STAGES;
  // N = 128: 12.9 of the 17.4 GB a launch pulled through the L2->SM crossbar were weight tiles re-streamed for every
  // 128-pixel tile.  With 256 accumulator columns free in TMEM a CTA keeps TWO pixel tiles in flight and issues both
  // tiles' MMAs from each weight tile, halving that stream (N = 256 has no TMEM left for it and does not need it).
  static constexpr int MT = BN == 128 ? HALO_MT128 : 1;
  static constexpr int A_STAGE_BYTES = MT * A_BYTES;    // one stage = the halo patches of this CTA's MT tiles
  static constexpr int B_BYTES = (BN / 2) * 64 * 2;     // this CTA's half of one tap's weight tile
  static constexpr bool RESIDENT_B = BN == 32;          // depth head: all (tap, channel-block) weight tiles stay in smem
  static constexpr int B_STAGES = BN == 256 ? 5 : (RESIDENT_B ? 18 : 8);   // resident: 9 taps x <= 2 channel blocks
  static constexpr int EPI_WARPS = BN >= 128 ? 8 : 4;    // N = 32 (depth head): one warp per TMEM lane quadrant holds the whole row
  static constexpr int HN = BN / (EPI_WARPS / 4);        // accumulator columns drained per epilogue warp
  static constexpr int STAGING_BYTES = RESIDENT_B ? 0 : EPI_WARPS * 32 * ::dav2::STG_ROW_BYTES;  // the head epilogue stores straight from registers
  static constexpr int VEC_BYTES = EPI_WARPS * HN * 4;
  static constexpr int TMEM_COLS = 2 * MT * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int SMEM_BYTES = 1024 + A_STAGES * A_STAGE_BYTES + B_STAGES * B_BYTES + STAGING_BYTES + VEC_BYTES + BAR_BYTES;
  static_assert(TMEM_COLS <= 512 && SMEM_BYTES <= 227 * 1024, "conv_halo resources");
};

// Same as make_sw128_desc plus the 3-bit "matrix base offset" field (bits 49-51).
__device__ __forceinline__ uint64_t make_sw128_desc_bo(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                       uint32_t base_offset) {
  return make_sw128_desc(smem_addr, lbo_bytes, sbo_bytes) | ((uint64_t)(base_offset & 7u) << 49);
}

template <int BN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ConvHaloCfg<BN>::THREADS, 1)
conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p, const int bo_mode) {
  using Cfg = ConvHaloCfg<BN>;
  constexpr int AS = Cfg::A_STAGES, BS = Cfg::B_STAGES, MT = Cfg::MT;
  static_assert(MODE == GM_CONV_BF16 || (MODE == GM_CONV_HEAD && BN == 32), "conv modes only");
  extern __shared__ uint8_t smem_raw[];

  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t sA = base;
  const uint32_t sB = sA + AS * Cfg::A_STAGE_BYTES;
  const uint32_t staging = sB + BS * Cfg::B_BYTES;
  const uint32_t vecs = staging + Cfg::STAGING_BYTES;
  const uint32_t bars = vecs + Cfg::VEC_BYTES;
  constexpr int BB = Cfg::RESIDENT_B ? 1 : BS;  // weight barriers: one "all resident tiles landed", or a full/empty pair per ring stage
  constexpr int NBAR = 2 * AS + 2 * BB + 4;
  static_assert(NBAR * 8 + 8 <= Cfg::BAR_BYTES, "barrier area");
  const uint32_t tmem_slot = bars + 8 * NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + AS * Cfg::A_STAGE_BYTES + BS * Cfg::B_BYTES + Cfg::STAGING_BYTES + Cfg::VEC_BYTES + 8 * NBAR);
#define AFULL(s) (bars + 8u * (uint32_t)(s))
#define AEMPTY(s) (bars + 8u * (uint32_t)(AS + (s)))
#define BFULL(s) (bars + 8u * (uint32_t)(2 * AS + (s)))
#define BEMPTY(s) (bars + 8u * (uint32_t)(2 * AS + BB + (s)))
#define TFULL_BAR(a) (bars + 8u * (uint32_t)(2 * AS + 2 * BB + (a)))
#define TEMPTY_BAR(a) (bars + 8u * (uint32_t)(2 * AS + 2 * BB + 2 + (a)))

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < AS; ++s) { mbar_init(AFULL(s), 1); mbar_init(AEMPTY(s), 1); }
    for (int s = 0; s < BB; ++s) { mbar_init(BFULL(s), 1); mbar_init(BEMPTY(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(TFULL_BAR(a), 1); mbar_init(TEMPTY_BAR(a), 2 * Cfg::EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int pairs_m = (p.tiles_m + 2 * MT - 1) / (2 * MT);  // a CTA pair advances by 2 * MT pixel tiles
  const int num_pt = pairs_m * p.tiles_n;
  const int pt0 = (int)(blockIdx.x >> 1), pt_stride = (int)(gridDim.x >> 1);
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs; whole warp, elected lane issues) ======
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    if constexpr (Cfg::RESIDENT_B) {
      // tiles_n == 1: every (tap, channel block) tile of this CTA's half of the output channels, once
      if (pt0 < num_pt && elect_one()) {
        if (leader) mbar_expect_tx(BFULL(0), 2u * Cfg::B_BYTES * (uint32_t)p.num_kb);
        for (int t = 0; t < p.num_kb; ++t)
          tma_load_2d_2sm(sB + t * Cfg::B_BYTES, &tmB, mapa_shared(BFULL(0), 0), t * 64, (int)rank * (BN / 2));
      }
      __syncwarp();
    }
    for (int pt = pt0; pt < num_pt; pt += pt_stride) {
      const int tmp = pt / p.tiles_n, tn = pt - tmp * p.tiles_n;
      int tb[MT], ty0[MT], tx0[MT];
#pragma unroll
      for (int h = 0; h < MT; ++h) {
        const int tm = (2 * tmp + (int)rank) * MT + h;
        tb[h] = tm / per_img;  // >= batch for the tail tiles: TMA zero-fills, epilogue masks
        const int r = tm - tb[h] * per_img;
        const int ty = r / p.tiles_x;
        ty0[h] = ty * Cfg::TH;
        tx0[h] = (r - ty * p.tiles_x) * Cfg::TW;
      }
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(AEMPTY(sa), pha ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(AFULL(sa), 2 * MT * Cfg::A_TX_BYTES);
#pragma unroll
          for (int h = 0; h < MT; ++h)
            tma_load_4d_2sm(sA + sa * Cfg::A_STAGE_BYTES + h * Cfg::A_BYTES, &tmA, mapa_shared(AFULL(sa), 0), cb * 64,
                            tx0[h] - 1, ty0[h] - 1, tb[h]);
        }
        __syncwarp();
        if (++sa == AS) { sa = 0; pha ^= 1u; }
        for (int tap = 0; tap < (Cfg::RESIDENT_B ? 0 : 9); ++tap) {
          mbar_wait(BEMPTY(sb), phb ^ 1u);
          if (elect_one()) {
            if (leader) mbar_expect_tx(BFULL(sb), 2 * Cfg::B_BYTES);
            tma_load_2d_2sm(sB + sb * Cfg::B_BYTES, &tmB, mapa_shared(BFULL(sb), 0), (tap * p.cblocks + cb) * 64,
                            tn * BN + (int)rank * (BN / 2));
          }
          __syncwarp();
          if (++sb == BS) { sb = 0; phb ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ================================ MMA issuer (leader CTA; whole warp, elected lane issues) =======
    const uint32_t idesc = make_idesc_h(256, BN, 0, 0, p.fmt);
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    int as = 0;
    uint32_t aphase = 0;
    if constexpr (Cfg::RESIDENT_B) {
      if (pt0 < num_pt) mbar_wait(BFULL(0), 0);
      tc_fence_after();
    }
    for (int pt = pt0; pt < num_pt; pt += pt_stride) {
      mbar_wait(TEMPTY_BAR(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * MT * BN);
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(AFULL(sa), pha);
        tc_fence_after();
        const uint32_t a_base = sA + sa * Cfg::A_STAGE_BYTES;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          if constexpr (Cfg::RESIDENT_B) {
            sb = tap * p.cblocks + cb;  // the weight matrix's K order: (tap, channel block)
          } else {
            mbar_wait(BFULL(sb), phb);
            tc_fence_after();
          }
          // rows of one 8-pixel group are contiguous (8 x 128 B); groups (output rows) are one patch line apart
          const uint32_t a_tap = a_base + (uint32_t)(dy * Cfg::HALO_W + dx) * 128u;
          const uint64_t bdesc = make_sw128_desc(sB + sb * Cfg::B_BYTES, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int h = 0; h < MT; ++h) {  // every pixel tile of this CTA against the same weight tile
              const uint64_t adesc = make_sw128_desc_bo(a_tap + (uint32_t)(h * Cfg::A_BYTES), 16, Cfg::HALO_W * 128,
                                                        bo_mode ? (uint32_t)dx : 0u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_h16_2sm(d_tmem + (uint32_t)(h * BN), adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((cb | tap | k) != 0));
            }
            if constexpr (!Cfg::RESIDENT_B) umma_commit_2sm_mc(BEMPTY(sb), 3);
            if (tap == 8) umma_commit_2sm_mc(AEMPTY(sa), 3);
            if (tap == 8 && cb == p.cblocks - 1) umma_commit_2sm_mc(TFULL_BAR(as), 3);
          }
          __syncwarp();
          if constexpr (!Cfg::RESIDENT_B) {
            if (++sb == BS) { sb = 0; phb ^= 1u; }
          }
        }
        if (++sa == AS) { sa = 0; pha ^= 1u; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp >= 2) {
    // ================================ epilogue (both CTAs, own 128 pixels) ===================
    const int q = warp & 3;
    constexpr int HN = Cfg::HN;
    const int col0 = ((warp - 2) >> 2) * HN;
    const uint32_t stg = staging + (uint32_t)(warp - 2) * 32u * STG_ROW_BYTES;
    const uint32_t vec = vecs + (uint32_t)(warp - 2) * (HN * 4);
    int as = 0;
    uint32_t aphase = 0;
    for (int pt = pt0; pt < num_pt; pt += pt_stride) {
      const int tmp = pt / p.tiles_n, tn = pt - tmp * p.tiles_n;
      if constexpr (MODE != GM_CONV_HEAD) epi_fill_bias<HN, MODE>(p, vec, lane, tn * BN + col0);
      mbar_wait(TFULL_BAR(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < MT; ++h) {
        TileGeom g;
        g.tm = (2 * tmp + (int)rank) * MT + h;
        g.cb_img = g.tm / per_img;
        const int r = g.tm - g.cb_img * per_img;
        const int ty = r / p.tiles_x;
        g.y0 = ty * Cfg::TH;
        g.x0 = (r - ty * p.tiles_x) * Cfg::TW;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * MT + h) * BN + col0);
        if (g.tm < p.tiles_m) {
          if constexpr (MODE == GM_CONV_HEAD) epi_tile_head(p, t_row, lane, q, g);
          else epi_tile_dispatch<HN, MODE>(p, t_row, stg, vec, lane, q, g, tn * BN + col0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(TEMPTY_BAR(as));
        else mbar_arrive_cluster(mapa_shared(TEMPTY_BAR(as), 0));
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
#undef AFULL
#undef AEMPTY
#undef BFULL
#undef BEMPTY
#undef TFULL_BAR
#undef TEMPTY_BAR
}

// Host launcher (gemm.cu).  tmA: NHWC map with box {64, 16, 18, 1}; tmB box rows bn/2; p.tw = 8, p.th = 16.
int launch_conv_halo(int bn, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream);
bool conv_halo_eligible(int bn, int mode, int tiles_m);

}  // namespace dav2
