// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T      h16 operands (K-major, 128B-swizzled smem tiles fed by TMA),
//                                   fp32 accumulators in TMEM (double buffered), fused epilogues.
//
// Roles (192 threads, 1 CTA / SM, grid = min(tiles, #SM), static round-robin tile schedule):
//   warp 0 (1 lane)  TMA producer  : cp.async.bulk.tensor -> smem ring (STAGES deep)
//   warp 1 (1 lane)  MMA issuer    : tcgen05.mma 128 x BN x 16, commit -> frees smem slot / signals epilogue
//   warps 2..5       epilogue      : tcgen05.ld -> regs -> per-warp smem transpose -> coalesced global I/O
//
// A operand addressing modes:
//   linear : A is a row-major [M,K] matrix (tokens x features, or NHWC pixels x channels for 1x1 convs)
//   conv   : A is an NHWC activation; a 128-row tile is a (th x tw) spatial patch; K runs over
//            9 taps x channel blocks, each k-block being ONE shifted 4-D TMA box (zero fill outside
//            the image = the conv's zero padding).  Nothing is im2col-materialised.
#pragma once

#include "gemm_epilogue.cuh"

namespace dav2 {

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 4 * 32 * ::dav2::STG_ROW_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int VEC_BYTES = 4 * BN * 4;  // per epilogue warp: bias[BN]
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + VEC_BYTES + BAR_BYTES;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(192, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool IS_CONV = (MODE == GM_CONV_BF16 || MODE == GM_CONV_HEAD);
  extern __shared__ uint8_t smem_raw[];

  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t staging = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t vecs = staging + Cfg::STAGING_BYTES;
  const uint32_t bars = vecs + Cfg::VEC_BYTES;
  const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES + Cfg::VEC_BYTES + 8 * (2 * STAGES + 4));
#define FULL_BAR(s) (bars + 8u * (uint32_t)(s))
#define EMPTY_BAR(s) (bars + 8u * (uint32_t)(STAGES + (s)))
#define TFULL_BAR(a) (bars + 8u * (uint32_t)(2 * STAGES + (a)))
#define TEMPTY_BAR(a) (bars + 8u * (uint32_t)(2 * STAGES + 2 + (a)))

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(FULL_BAR(s), 1);
      mbar_init(EMPTY_BAR(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(TFULL_BAR(a), 1);
      mbar_init(TEMPTY_BAR(a), 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    // ================================ TMA producer (whole warp, elected lane issues) ==========
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      int b = 0, x0 = 0, y0 = 0;
      if (IS_CONV) {
        const int per_img = p.tiles_x * p.tiles_y;
        b = tm / per_img;
        const int r = tm - b * per_img;
        const int ty = r / p.tiles_x;
        y0 = ty * p.th;
        x0 = (r - ty * p.tiles_x) * p.tw;
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(EMPTY_BAR(stage), phase ^ 1u);
        const uint32_t a_dst = base + stage * Cfg::STAGE_BYTES;
        const uint32_t b_dst = a_dst + Cfg::A_BYTES;
        if (elect_one()) {
          mbar_expect_tx(FULL_BAR(stage), Cfg::STAGE_BYTES);
          if (IS_CONV) {
            const int tap = kb / p.cblocks;
            const int cb = kb - tap * p.cblocks;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            tma_load_4d(a_dst, &tmA, FULL_BAR(stage), cb * 64, x0 + dx, y0 + dy, b);
          } else {
            tma_load_2d(a_dst, &tmA, FULL_BAR(stage), kb * 64, tm * 128);
          }
          tma_load_2d(b_dst, &tmB, FULL_BAR(stage), kb * 64, tn * BN);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (whole warp, elected lane issues) ============
    const uint32_t idesc = make_idesc_h(128, BN, 0, 0, p.fmt);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(TEMPTY_BAR(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(FULL_BAR(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = base + stage * Cfg::STAGE_BYTES;
        const uint64_t adesc = make_sw128_desc(a_addr, 16, 1024);
        const uint64_t bdesc = make_sw128_desc(a_addr + Cfg::A_BYTES, 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_h16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
          umma_commit(EMPTY_BAR(stage));
          if (kb == p.num_kb - 1) umma_commit(TFULL_BAR(as));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp >= 2) {
    // ================================ epilogue ==============================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t stg = staging + (uint32_t)(warp - 2) * 32u * STG_ROW_BYTES;
    const uint32_t vec = vecs + (uint32_t)(warp - 2) * (BN * 4);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      int cb_img = 0, x0 = 0, y0 = 0;
      if (IS_CONV) {
        const int per_img = p.tiles_x * p.tiles_y;
        cb_img = tm / per_img;
        const int r = tm - cb_img * per_img;
        const int ty = r / p.tiles_x;
        y0 = ty * p.th;
        x0 = (r - ty * p.tiles_x) * p.tw;
      }
      if constexpr (MODE != GM_CONV_HEAD) epi_fill_bias<BN, MODE>(p, vec, lane, tn * BN);
      mbar_wait(TFULL_BAR(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);

      TileGeom g;
      g.tm = tm; g.cb_img = cb_img; g.x0 = x0; g.y0 = y0;
      if constexpr (MODE == GM_CONV_HEAD) {
        static_assert(MODE != GM_CONV_HEAD || BN == 32, "head epilogue expects N tile 32");
        epi_tile_head(p, t_row, lane, q, g);
      } else {
        epi_tile_dispatch<BN, MODE>(p, t_row, stg, vec, lane, q, g, tn * BN);
      }
      tc_fence_before();
      mbar_arrive(TEMPTY_BAR(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
#undef FULL_BAR
#undef EMPTY_BAR
#undef TFULL_BAR
#undef TEMPTY_BAR
}

// Host launcher (gemm.cu).  tmA/tmB must have been encoded with box rows 128 / BN.
int launch_gemm(int bn, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                cudaStream_t stream);
int pick_bn(int N);

}  // namespace dav2
