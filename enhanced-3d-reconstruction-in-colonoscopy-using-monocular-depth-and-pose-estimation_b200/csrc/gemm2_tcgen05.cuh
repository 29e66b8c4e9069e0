// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM / implicit-GEMM conv kernel.
//
// A CTA pair (cluster 2x1x1 = the two SMs of one TPC) computes a 256 x BN output tile with ONE
// tcgen05.mma.cta_group::2 stream issued by the leader CTA (UMMA M = 256):
//   * each CTA TMA-loads only ITS 128 rows of A and ITS BN/2 rows of B per k-block (32 KB instead of
//     48 KB per 128xBNx64 MACs): the per-SM L2->smem operand traffic and the smem fill bandwidth drop by
//     1/3, and the smem ring becomes 6 deep instead of 4 (the 1-CTA kernel is TMA-latency bound);
//   * both CTAs' loads complete_tx on the LEADER's full barrier; the leader's tcgen05.commit is
//     multicast to both CTAs' empty / tmem-full barriers; both CTAs' epilogue warps arrive (remotely for
//     the follower) on the leader's tmem-empty barrier;
//   * each CTA owns the accumulator rows of its A half in its own TMEM (double buffered) and runs its own
//     epilogue, so the epilogue code is per-CTA and identical in structure to the 1-CTA kernel's.
// Epilogue changes vs the 1-CTA kernel: LayerScale+residual uses cp.reduce.async.bulk (.add.f32) from
// smem -- the fp32 residual stream is never LOADED by the SM; GELU uses a 1.5e-7-accurate erf polynomial.
#pragma once

#include "gemm_tcgen05.cuh"

namespace dav2 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of THIS CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_h16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in every CTA of `mask` once all prior tcgen05 ops completed
__device__ __forceinline__ void umma_commit_2sm_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// x[global] += smem row (fp32), asynchronous, no load of x into the SM
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Sixteen epilogue warps (four per SM sub-partition, 64 accumulator columns each) for the 256-wide bias / GELU tile: with
// eight, the fc1 + GELU epilogue could not keep pace with the MMA stream (ncu: 70 % tensor pipe, 50 % issue with two
// warps per sub-partition; 81 % with four).  The LayerScale + residual mode keeps eight warps and the fifth smem stage:
// its epilogue is light and its K = 4096 main loop wants the deeper ring (96 % -> 87 % with four stages).
__host__ __device__ constexpr bool gemm2_wide_epi(int bn, int mode) { return bn == 256 && mode == GM_LINEAR_BF16; }

template <int BN, bool WIDE>
struct Gemm2Cfg {
  static constexpr int A_BYTES = 128 * 64 * 2;          // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * 64 * 2;     // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = WIDE ? 16 : 8;        // 2 (4) per SM sub-partition: each TMEM lane quadrant is drained by two (four) warps
  static constexpr int STAGES = WIDE ? 4 : (160 * 1024) / STAGE_BYTES;  // 5 (BN=256) / 6 (BN=128); 4 with the wide epilogue's staging
  static constexpr int HN = BN / (EPI_WARPS / 4);        // accumulator columns drained per epilogue warp
  static constexpr int STAGING_BYTES = EPI_WARPS * 32 * ::dav2::STG_ROW_BYTES;
  static constexpr int VEC_BYTES = EPI_WARPS * 2 * HN * 4;  // per epilogue warp: bias[HN] | gamma[HN] of its column slice
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + VEC_BYTES + BAR_BYTES;
};

template <int BN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * Gemm2Cfg<BN, gemm2_wide_epi(BN, MODE)>::EPI_WARPS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const GemmParams p) {
  using Cfg = Gemm2Cfg<BN, gemm2_wide_epi(BN, MODE)>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool IS_CONV = (MODE == GM_CONV_BF16);
  static_assert(MODE != GM_CONV_HEAD, "the N=32 head stays on the 1-CTA kernel");
  static_assert(STAGES * 2 + 4 <= 30, "barrier area");
  extern __shared__ uint8_t smem_raw[];

  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t staging = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t vecs = staging + Cfg::STAGING_BYTES;
  const uint32_t bars = vecs + Cfg::VEC_BYTES;
  const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES + Cfg::VEC_BYTES + 8 * (2 * STAGES + 4));
#define FULL_BAR(s) (bars + 8u * (uint32_t)(s))
#define EMPTY_BAR(s) (bars + 8u * (uint32_t)(STAGES + (s)))
#define TFULL_BAR(a) (bars + 8u * (uint32_t)(2 * STAGES + (a)))
#define TEMPTY_BAR(a) (bars + 8u * (uint32_t)(2 * STAGES + 2 + (a)))

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(FULL_BAR(s), 1);   // leader: its own arrive.expect_tx (bytes of BOTH CTAs)
      mbar_init(EMPTY_BAR(s), 1);  // the leader's multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(TFULL_BAR(a), 1);   // the leader's multicast commit
      mbar_init(TEMPTY_BAR(a), 2 * Cfg::EPI_WARPS);  // every epilogue warp of BOTH CTAs (only the leader's copy is waited on)
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int pairs_m = (p.tiles_m + 1) >> 1;
  const int num_pt = pairs_m * p.tiles_n;
  const int pt0 = (int)(blockIdx.x >> 1), pt_stride = (int)(gridDim.x >> 1);

  if (warp == 0) {
    // ================================ TMA producer (both CTAs; whole warp, elected lane issues) ======
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = pt0; pt < num_pt; pt += pt_stride) {
      const int tmp = pt / p.tiles_n, tn = pt - tmp * p.tiles_n;
      const int tm = 2 * tmp + (int)rank;
      int b = 0, x0 = 0, y0 = 0;
      if (IS_CONV) {
        const int per_img = p.tiles_x * p.tiles_y;
        b = tm / per_img;  // >= batch for the odd tail tile: TMA zero-fills, epilogue masks
        const int r = tm - b * per_img;
        const int ty = r / p.tiles_x;
        y0 = ty * p.th;
        x0 = (r - ty * p.tiles_x) * p.tw;
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(EMPTY_BAR(stage), phase ^ 1u);
        const uint32_t a_dst = base + stage * Cfg::STAGE_BYTES;
        const uint32_t b_dst = a_dst + Cfg::A_BYTES;
        const uint32_t full_leader = mapa_shared(FULL_BAR(stage), 0);
        if (elect_one()) {
          if (leader) mbar_expect_tx(FULL_BAR(stage), 2 * Cfg::STAGE_BYTES);
          if (IS_CONV) {
            const int tap = kb / p.cblocks;
            const int cb = kb - tap * p.cblocks;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            tma_load_4d_2sm(a_dst, &tmA, full_leader, cb * 64, x0 + dx, y0 + dy, b);
          } else {
            tma_load_2d_2sm(a_dst, &tmA, full_leader, kb * 64, tm * 128);
          }
          tma_load_2d_2sm(b_dst, &tmB, full_leader, kb * 64, tn * BN + (int)rank * (BN / 2));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ================================ MMA issuer (leader CTA; whole warp, elected lane issues) =======
    const uint32_t idesc = make_idesc_h(256, BN, 0, 0, p.fmt);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int pt = pt0; pt < num_pt; pt += pt_stride) {
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
            umma_h16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
          umma_commit_2sm_mc(EMPTY_BAR(stage), 3);
          if (kb == p.num_kb - 1) umma_commit_2sm_mc(TFULL_BAR(as), 3);
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
    // ================================ epilogue (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    constexpr int HN = Cfg::HN;                      // columns drained by this warp
    const int col0 = ((warp - 2) >> 2) * HN;           // warps 2..5 -> first half, 6..9 -> second half
    const uint32_t stg = staging + (uint32_t)(warp - 2) * 32u * STG_ROW_BYTES;
    const uint32_t vec = vecs + (uint32_t)(warp - 2) * (2 * HN * 4);
    int as = 0;
    uint32_t aphase = 0;
    for (int pt = pt0; pt < num_pt; pt += pt_stride) {
      const int tmp = pt / p.tiles_n, tn = pt - tmp * p.tiles_n;
      const int tm = 2 * tmp + (int)rank;
      int cb_img = 0, x0 = 0, y0 = 0;
      if (IS_CONV) {
        const int per_img = p.tiles_x * p.tiles_y;
        cb_img = tm / per_img;
        const int r = tm - cb_img * per_img;
        const int ty = r / p.tiles_x;
        y0 = ty * p.th;
        x0 = (r - ty * p.tiles_x) * p.tw;
      }
      const bool tile_valid = tm < p.tiles_m;
      if constexpr (MODE == GM_LINEAR_RESID) {
        // bias | gamma of this tile's BN columns -> per-warp smem (read back as warp-wide broadcasts)
        for (int j = lane; j < HN; j += 32) {
          const int n = tn * BN + col0 + j;
          const float bv = (n < p.N && p.bias) ? __ldg(p.bias + n) : 0.f;
          const float gv = (n < p.N) ? __ldg(p.gamma + n) : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + j * 4), "f"(bv) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + (HN + j) * 4), "f"(gv) : "memory");
        }
        __syncwarp();
      } else {
        epi_fill_bias<HN, MODE>(p, vec, lane, tn * BN + col0);
      }
      mbar_wait(TFULL_BAR(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + col0);

      TileGeom g;
      g.tm = tm; g.cb_img = cb_img; g.x0 = x0; g.y0 = y0;
      if constexpr (MODE == GM_LINEAR_RESID) {
        // thread = row: x[m, nc:nc+32] += gamma * (acc + bias) through an async bulk reduce-add (no load of x)
#pragma unroll 1
        for (int c = 0; c < HN / 32; ++c) {
          const int nc = tn * BN + col0 + c * 32;
          if (nc >= p.N || !tile_valid) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          bulk_wait_read0();  // my previous bulk op has finished reading my staging row
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b4, g4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                         : "r"(vec + (c * 32 + 4 * j) * 4) : "memory");
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g4.x), "=f"(g4.y), "=f"(g4.z), "=f"(g4.w)
                         : "r"(vec + (HN + c * 32 + 4 * j) * 4) : "memory");
            const float o0 = g4.x * (__uint_as_float(v[4 * j]) + b4.x), o1 = g4.y * (__uint_as_float(v[4 * j + 1]) + b4.y);
            const float o2 = g4.z * (__uint_as_float(v[4 * j + 2]) + b4.z), o3 = g4.w * (__uint_as_float(v[4 * j + 3]) + b4.w);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * STG_ROW_BYTES + j * 16), "f"(o0),
                         "f"(o1), "f"(o2), "f"(o3)
                         : "memory");
          }
          fence_proxy_async_smem();
          const int m = tm * 128 + q * 32 + lane;
          if (m < p.M) {
            const int ncols = min(32, p.N - nc);
            bulk_reduce_add_f32(reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + nc, stg + lane * STG_ROW_BYTES,
                                (uint32_t)ncols * 4u);
          }
          bulk_commit();
        }
      } else {
        if (tile_valid) epi_tile_dispatch<HN, MODE>(p, t_row, stg, vec, lane, q, g, tn * BN + col0);
      }
      // accumulator stage drained: one arrival per warp on the LEADER's tmem-empty barrier
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(TEMPTY_BAR(as));
        else mbar_arrive_cluster(mapa_shared(TEMPTY_BAR(as), 0));
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if constexpr (MODE == GM_LINEAR_RESID) bulk_wait_all0();
  }

  tc_fence_before();
  __syncwarp();        // re-converge the role-divergent warps 0 / 1 before the .aligned cluster barrier
  cluster_sync_all();  // nobody exits (smem / barriers / TMEM) while the partner may still signal it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
#undef FULL_BAR
#undef EMPTY_BAR
#undef TFULL_BAR
#undef TEMPTY_BAR
}

// Host launcher (gemm.cu): tmA box rows 128, tmB box rows bn/2.
int launch_gemm2(int bn, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                 cudaStream_t stream);
bool gemm2_eligible(int bn, int mode, int tiles_m);

}  // namespace dav2
