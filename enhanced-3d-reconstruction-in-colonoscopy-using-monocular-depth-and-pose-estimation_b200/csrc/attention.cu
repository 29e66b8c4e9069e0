// Fused flash-style attention for DINOv2 (d_head = 64, non-causal, arbitrary token count) on sm_100a.
//
// One CTA = one 128-query tile of one (image, head); 2 CTAs co-reside per SM so that one CTA's
// softmax overlaps the other's tensor-core work.  256 threads = two warpgroups:
//   WG0 warp 0      : TMA producer  - Q once, then K/V tiles through a 2-deep smem ring
//   WG0 warp 1      : MMA issuer    - S = Q K^T (tcgen05, 128x128x64 -> TMEM cols [0,128)),
//                                     O += P V (128x64x128 -> TMEM cols [128,192)); P is the A operand read from
//                                     TMEM cols [192,256), V the MN-major B operand straight from the [token, 3D]
//                                     qkv buffer
//   WG0 warps 2,3   : idle (they only exist so that setmaxnreg can hand WG0's registers to WG1)
//   WG1 warps 4..7  : softmax       - one query row per thread: tcgen05.ld S, online max / exp2 / sum in
//                                     fp32, P -> h16 pairs -> TMEM (tcgen05.st), O accumulates in TMEM (lazy rescale)
// Softmax arithmetic (the bound of this kernel: 16 MUFU/clk/SM = 1024 cycles per 128x128 tile against 512
// tensor-pipe cycles): packed fp32 (FFMA2/FADD2) for the scale-subtract and the row sums, and EMU of every 8
// element pairs take their 2^x from a Cody-Waite + degree-3 polynomial on the FMA pipe instead of MUFU.EX2
// (max rel. error 7.5e-5, below the 4.9e-4 rounding of P to fp16), which balances the XU pipe against issue slots.
// WG1 runs with 208 registers (setmaxnreg) so the 128-wide S row plus the exp pipeline stay in registers with ILP.
// Q was pre-scaled by d^-1/2 = 0.125 (folded exactly into the qkv weights), so S needs no scale.
// Reads qkv h16 [B*N, 3*D] (q | k | v, head h at columns h*64), writes out h16 [B*N, D].
#include <stdlib.h>

#include "gemm_epilogue.cuh"

namespace dav2 {

struct AttnParams {
  h16* out;
  int N;      // tokens per image
  int D;      // model width (= heads * 64)
  int nkv;    // ceil(N / 128)
  uint32_t v_lbo, v_sbo;  // MN-major descriptor strides for V (bytes)
  int fmt;                // FMT_F16 / FMT_BF16
};

static constexpr int ATT_TILE = 128 * 64 * 2;  // 16 KB: one [128 x 64] h16 tile
// shared memory: Q (16 KB) + two K and two V stages of KV x 64 h16 + barriers (P lives in TMEM)
static constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair, on the FMA/ALU pipes: n = round(x) by the 1.5*2^23 magic add, f = x - n in [-0.5, 0.5],
// p(f) ~ 2^f (degree-3 minimax, rel. err 7.5e-5), result = p with n added to its exponent field.
// SASS: 2 FMNMX + 3 FADD2 + 3 FFMA2 + 2 LEA per pair.
__device__ __forceinline__ float2 ex2_emu2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 r = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 n = __fadd2_rn(r, make_float2(-12582912.f, -12582912.f));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(make_float2(0.05517118f, 0.05517118f), f, make_float2(0.24260994f, 0.24260994f));
  p = __ffma2_rn(p, f, make_float2(0.69326096f, 0.69326096f));
  p = __ffma2_rn(p, f, make_float2(0.99992815f, 0.99992815f));
  float2 o;
  o.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23));
  o.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23));
  return o;
}

// Two round-2 experiments are kept as A/B builds (scripts/build_attn_variants.sh); both measured SLOWER than the default
// on B200 (vitl shape, 64 x 16 heads x 1370 tokens, same box: default 0.675 ms):
//   -DATTN_SPECULATIVE  no row max after the first tile: exponentiate against the first tile's max, the tile's row sum
//                       (<= 2^13) is the overflow guard, recentre + recompute on the rare violation.  Removes 32 FMNMX3 +
//                       vote per tile but keeps the S row live to the end of the tile (128 registers): 0.696 ms.
//   -DATTN_PREFETCH     tcgen05.ld of S(j+1) issued before the P hand-over of tile j when it is already committed: 0.733 ms.
// What they showed: the softmax warps are bound by their OWN instruction stream (ncu source page: 28 % fixed-latency
// dependency stalls, 18 % issuing, only 12 % lost arbitration; XU 52 %, issue 56 %), not by the max -> vote chain or the
// TMEM round trip.
#if !defined(ATTN_SPECULATIVE) && !defined(ATTN_PREFETCH)  // S registers die inside the exp loop, none prefetched
#define ATTN_REGS_WG0 40
#define ATTN_REGS_WG1 120
#else
#define ATTN_REGS_WG0 32
#define ATTN_REGS_WG1 128
#endif
#if ATTN_REGS_WG0 >= 40  // the unrolled issue loops keep ~10 descriptor registers live: they spill at 32
#define ATTN_MMA_UNROLL _Pragma("unroll")
#else
#define ATTN_MMA_UNROLL _Pragma("unroll 1")
#endif

template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
}

// max of a row of S held in registers: eight independent FMNMX3 chains (a single running max is a 64-deep dependent chain)
template <int KV>
__device__ __forceinline__ float row_max(const uint32_t (&sreg)[KV]) {
  float mxc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) mxc[c] = fmaxf(__uint_as_float(sreg[2 * c]), __uint_as_float(sreg[2 * c + 1]));
#pragma unroll
  for (int i = 16; i < KV; i += 16)
#pragma unroll
    for (int c = 0; c < 8; ++c) mxc[c] = fmax3(mxc[c], __uint_as_float(sreg[i + 2 * c]), __uint_as_float(sreg[i + 2 * c + 1]));
  return fmaxf(fmax3(mxc[0], mxc[1], mxc[2]), fmaxf(fmax3(mxc[3], mxc[4], mxc[5]), fmaxf(mxc[6], mxc[7])));
}

// In-kernel timeline (profiling builds only, -DATTN_TRACE): clock64 stamps of one softmax lane and of the MMA issuer for
// the CTAs of image 32 / head 8, 16 slots per (q-tile, kv-tile).
#ifdef ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;
#define ATTN_STAMP(slot)                                                                                   \
  do {                                                                                                     \
    if (g_attn_trace && blockIdx.z == 32 && blockIdx.y == 8 && threadIdx.x == 128)                         \
      g_attn_trace[((long long)blockIdx.x * p.nkv + j) * 16 + (slot)] = clock64();                         \
  } while (0)
#define ATTN_STAMP_M(slot)                                                                                 \
  do {                                                                                                     \
    if (g_attn_trace && blockIdx.z == 32 && blockIdx.y == 8 && threadIdx.x == 32)                          \
      g_attn_trace[((long long)blockIdx.x * p.nkv + j) * 16 + (slot)] = clock64();                         \
  } while (0)
#else
#define ATTN_STAMP(slot) do { } while (0)
#define ATTN_STAMP_M(slot) do { } while (0)
#endif

// KV = keys per tile.  128: two CTAs per SM (TMEM: S 128 + O 64 + P 64 = 256 columns each).  64: THREE CTAs per SM
// (S 64 + O 64 in one 128-column allocation, P in a second 32-column one: 480 of the 512 columns), i.e. three softmax
// warps per SM sub-partition instead of two to hide each other's TMEM / barrier / MUFU latencies -- tensor memory, not
// registers, is what limits the number of query tiles in flight.
template <bool FP16, int EMU, int KV>
__global__ void __launch_bounds__(256, KV == 64 ? 3 : 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  constexpr int KV_TILE = KV * 128;  // bytes of one [KV x 64] h16 tile
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();  // swizzle-128B tiles need 1024 B alignment
  const uint32_t sQ = base;
  const uint32_t sK = base + ATT_TILE;                 // 2 stages
  const uint32_t sV = base + ATT_TILE + 2 * KV_TILE;   // 2 stages
  const uint32_t bars = base + ATT_TILE + 4 * KV_TILE;
  // K and V tiles have separate full/empty barriers: a K stage is free as soon as its S MMA has completed (the start of
  // that tile's softmax), a V stage only after its PV MMA.  With one barrier pair per (K, V) stage the load of K(j+1)
  // could not start before PV(j-1) had finished, and every tile exposed a full TMA latency in front of S(j+1).
  const uint32_t BAR_Q = bars, BAR_K_FULL = bars + 8, BAR_K_EMPTY = bars + 24, BAR_V_FULL = bars + 40, BAR_V_EMPTY = bars + 56,
                 BAR_S_FULL = bars + 72, BAR_S_EMPTY = bars + 80, BAR_P_FULL = bars + 88, BAR_O_FULL = bars + 96;
  const uint32_t tmem_slot = bars + 104;  // two slots: main allocation, P allocation (KV = 64 only)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + ATT_TILE + 4 * KV_TILE + 104);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * p.N;            // first token row of this image in the [B*N, 3D] buffer
  const int q0 = qt * 128;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmKV);
    mbar_init(BAR_Q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR_K_FULL + 8 * s, 1);
      mbar_init(BAR_K_EMPTY + 8 * s, 1);
      mbar_init(BAR_V_FULL + 8 * s, 1);
      mbar_init(BAR_V_EMPTY + 8 * s, 1);
    }
    mbar_init(BAR_S_FULL, 1);
    mbar_init(BAR_S_EMPTY, 128);
    mbar_init(BAR_P_FULL, 128);
    mbar_init(BAR_O_FULL, 1);
    fence_mbar_init();
    // The first loads (Q and both K / V stages) go out BEFORE the CTA-wide sync below: their latency then overlaps the
    // TMEM allocation and the barrier round trip of the prologue (a CTA lives for only 11-22 KV tiles at 518^2).
    mbar_expect_tx(BAR_Q, ATT_TILE);
    tma_load_2d(sQ, &tmQKV, BAR_Q, h * 64, row0 + q0);
    for (int j = 0; j < 2 && j < p.nkv; ++j) {
      mbar_expect_tx(BAR_K_FULL + 8 * j, KV_TILE);
      tma_load_2d(sK + j * KV_TILE, &tmKV, BAR_K_FULL + 8 * j, p.D + h * 64, row0 + j * KV);
      mbar_expect_tx(BAR_V_FULL + 8 * j, KV_TILE);
      tma_load_2d(sV + j * KV_TILE, &tmKV, BAR_V_FULL + 8 * j, 2 * p.D + h * 64, row0 + j * KV);
    }
  }
  if (warp == 1) {
    if (KV == 128) {
      tmem_alloc(tmem_slot, 256);
    } else {  // 128 + 32 columns (allocations are powers of two; 160 would round up to 256 and cost the third CTA)
      tmem_alloc_hold(tmem_slot, 128);
      tmem_alloc(tmem_slot + 4, 32);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot_ptr[0];
  // S fp32 | O fp32 | P 16-bit pairs
  const uint32_t tS = tmem_base, tO = tmem_base + KV, tP = KV == 128 ? tmem_base + 192 : tmem_slot_ptr[1];

  if (warp < 4) {
  setmaxnreg_dec<KV == 64 ? ATTN_REGS_WG0 : 48>();
  if (warp == 0) {
    // ---------------- TMA producer (whole warp, elected lane issues); tiles 0 and 1 were issued in the prologue ----
    for (int j = 2; j < p.nkv; ++j) {
      const int s = j & 1;
      mbar_wait(BAR_K_EMPTY + 8 * s, ((uint32_t)(j >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(BAR_K_FULL + 8 * s, KV_TILE);
        tma_load_2d(sK + s * KV_TILE, &tmKV, BAR_K_FULL + 8 * s, p.D + h * 64, row0 + j * KV);
      }
      __syncwarp();
      mbar_wait(BAR_V_EMPTY + 8 * s, ((uint32_t)(j >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(BAR_V_FULL + 8 * s, KV_TILE);
        tma_load_2d(sV + s * KV_TILE, &tmKV, BAR_V_FULL + 8 * s, 2 * p.D + h * 64, row0 + j * KV);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp, elected lane issues) ----------------
    const uint32_t idesc_s = make_idesc_h(128, KV, 0, 0, p.fmt);   // S = Q K^T : both K-major
    const uint32_t idesc_o = make_idesc_h(128, 64, 0, 1, p.fmt);   // O = P V   : V is MN-major
    const uint64_t qdesc = make_sw128_desc(sQ, 16, 1024);
    mbar_wait(BAR_Q, 0);
    mbar_wait(BAR_K_FULL, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t kdesc = make_sw128_desc(sK, 16, 1024);
ATTN_MMA_UNROLL
      for (int k = 0; k < 4; ++k) umma_h16(tS, qdesc + 2u * k, kdesc + 2u * k, idesc_s, (uint32_t)(k != 0));
      umma_commit(BAR_S_FULL);
      umma_commit(BAR_K_EMPTY);
    }
    __syncwarp();
    for (int j = 0; j < p.nkv; ++j) {
      const int s = j & 1;
      if (j + 1 < p.nkv) {
        const int s1 = (j + 1) & 1;
        ATTN_STAMP_M(8);
        mbar_wait(BAR_K_FULL + 8 * s1, (uint32_t)((j + 1) >> 1) & 1u);
        ATTN_STAMP_M(9);
        mbar_wait(BAR_S_EMPTY, (uint32_t)j & 1u);  // softmax has pulled S(j) into registers
        tc_fence_after();
        if (elect_one()) {
          const uint64_t kdesc = make_sw128_desc(sK + s1 * KV_TILE, 16, 1024);
ATTN_MMA_UNROLL
          for (int k = 0; k < 4; ++k) umma_h16(tS, qdesc + 2u * k, kdesc + 2u * k, idesc_s, (uint32_t)(k != 0));
          umma_commit(BAR_S_FULL);
          umma_commit(BAR_K_EMPTY + 8 * s1);
        }
        __syncwarp();
      }
      mbar_wait(BAR_V_FULL + 8 * s, (uint32_t)(j >> 1) & 1u);
      ATTN_STAMP_M(10);
      mbar_wait(BAR_P_FULL, (uint32_t)j & 1u);  // P(j) in smem, any rescale of O finished
      ATTN_STAMP_M(11);
      tc_fence_after();
      if (elect_one()) {
ATTN_MMA_UNROLL
        for (int k = 0; k < KV / 16; ++k) {  // 16 keys per MMA = 8 TMEM columns of P
          const uint64_t vdesc = make_sw128_desc(sV + s * KV_TILE + k * 2048, p.v_lbo, p.v_sbo);
          umma_h16_ts(tO, tP + 8u * k, vdesc, idesc_o, (uint32_t)((j | k) != 0));  // O accumulates in TMEM across KV tiles
        }
        umma_commit(BAR_O_FULL);
        umma_commit(BAR_V_EMPTY + 8 * s);
        ATTN_STAMP_M(12);
      }
      __syncwarp();
    }
  }
  } else {
    setmaxnreg_inc<KV == 64 ? ATTN_REGS_WG1 : 208>();
    // ---------------- softmax / output (one query row per thread) ----------------
    // The whole S row (KV fp32) is pulled into registers with ONE exposed TMEM round trip, which frees the S buffer at
    // once (the issuer overlaps S(j+1) with this tile's softmax).  O accumulates in TMEM across KV tiles; it is rescaled
    // (TMEM load-scale-store) only when some row's running max grew by more than 2^8 since its reference max was taken
    // ("lazy rescaling": P <= 256 is harmless in fp16/bf16 with fp32 accumulation, and the final O / l is independent of
    // the reference).  Barrier "probes" are mbarrier.test_wait or absent: try_wait suspends the thread.
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row within the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float m_ref = 0.f, l = 0.f;
    uint32_t sreg[KV];
    bool prefetched = false;

    for (int j = 0; j < p.nkv; ++j) {
      const int nvalid = p.N - j * KV;  // keys of this tile inside the image (>= 1)
      ATTN_STAMP(0);
      if (!prefetched) {
        mbar_wait(BAR_S_FULL, (uint32_t)j & 1u);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < KV / 32; ++c) tmem_ld32(tS + lane_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sreg[c * 32]));
      }
      ATTN_STAMP(1);
      // PV(j-1) done <=> the P columns may be rewritten (and O rescaled on the slow path): waited for only where it is
      // needed, right before the P store (an early try_wait "probe" here blocked the warp until PV(j-1) had finished)
      uint32_t o_ready = (j == 0) ? 1u : 0u;
      tmem_ld_wait();
#ifndef ATTN_NO_REGFENCE
#pragma unroll
      for (int i = 0; i < KV; ++i) asm volatile("" : "+r"(sreg[i]));  // consumers stay below the wait
#endif
      ATTN_STAMP(2);
      tc_fence_before();
      mbar_arrive(BAR_S_EMPTY);  // S(j) is in registers: the issuer may overwrite the TMEM buffer with S(j+1)
      if (nvalid < KV) {         // warp-uniform: tail tile, keys beyond the image never contribute
#pragma unroll
        for (int i = 0; i < KV; ++i)
          if (i >= nvalid) sreg[i] = 0xff800000u;  // -inf
      }
      // O *= alpha (TMEM load-scale-store), l *= alpha, m_ref = m_new: needs PV(j-1) complete
      auto rescale_to = [&](float m_new) {
        const float alpha = fast_exp2((m_ref - m_new) * LOG2E);  // <= 1; exactly 1 for rows that did not move
        if (!o_ready) mbar_wait(BAR_O_FULL, (uint32_t)(j - 1) & 1u);  // PV(j-1) has finished accumulating into O
        o_ready = 1u;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t ov[16];
          tmem_ld16(tO + lane_addr + c * 16, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
          tmem_st16(tO + lane_addr + c * 16, ov);
        }
        tmem_st_wait();
        tc_fence_before();
        l *= alpha;
        m_ref = m_new;
      };
#ifndef ATTN_SPECULATIVE
      {
        const float mx = row_max<KV>(sreg);
        if (j == 0) {
          m_ref = mx;
        } else {
          const float m_new = fmaxf(m_ref, mx);
          const bool need = (m_new - m_ref) * LOG2E > 8.0f;  // lazy: P <= 2^8 is harmless
          if (__any_sync(0xffffffffu, need)) rescale_to(need ? m_new : m_ref);
        }
      }
#else
      if (j == 0) m_ref = row_max<KV>(sreg);
#endif
      ATTN_STAMP(3);
      uint32_t preg[KV / 2];
      float ts;
#pragma unroll 1
      for (int attempt = 0;; ++attempt) {
        const float mscaled = m_ref * LOG2E;
        const float2 l2e2 = make_float2(LOG2E, LOG2E), nm2 = make_float2(-mscaled, -mscaled);
        float2 rs[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < KV / 2; ++i) {  // pair i = keys 2i, 2i+1
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sreg[2 * i]), __uint_as_float(sreg[2 * i + 1])), l2e2, nm2);
          float2 e;
          if ((i & 7) < EMU) {
            e = ex2_emu2(x);
          } else {
            e.x = fast_exp2(x.x);
            e.y = fast_exp2(x.y);
          }
          rs[i & 3] = __fadd2_rn(rs[i & 3], e);
          preg[i] = FP16 ? pack2<FMT_F16>(e.x, e.y) : pack2<FMT_BF16>(e.x, e.y);
        }
        ts = ((rs[0].x + rs[0].y) + (rs[1].x + rs[1].y)) + ((rs[2].x + rs[2].y) + (rs[3].x + rs[3].y));
#ifndef ATTN_SPECULATIVE
        break;
#else
        // !(ts <= 2^13) also catches inf / NaN sums.  The first tile (true max) and a recomputed tile have ts <= KV.
        if (attempt != 0 || j == 0 || !__any_sync(0xffffffffu, !(ts <= 8192.0f))) break;
        // ---- slow path (rare): recentre on the true maximum seen so far, rescale O and l, recompute the tile ----
        rescale_to(fmaxf(m_ref, row_max<KV>(sreg)));
#endif
      }
      l += ts;
      ATTN_STAMP(5);
      // S(j+1) already committed?  Pull it into the (now dead) S registers before handing P over.
      prefetched = false;
#ifdef ATTN_PREFETCH
      if (j + 1 < p.nkv && __all_sync(0xffffffffu, mbar_test_wait(BAR_S_FULL, (uint32_t)(j + 1) & 1u))) {
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < KV / 32; ++c) tmem_ld32(tS + lane_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sreg[c * 32]));
        prefetched = true;
      }
#endif
      if (!o_ready) mbar_wait(BAR_O_FULL, (uint32_t)(j - 1) & 1u);  // PV(j-1) no longer reads the P columns
      tc_fence_after();
      ATTN_STAMP(6);
      // P (16-bit pairs: column i of row r = keys 2i, 2i+1) goes straight into TMEM as the A operand of the PV MMA: one
      // tcgen05.st burst instead of sixteen swizzled st.shared + a generic->async proxy fence (~370 cycles per tile in
      // the in-kernel timeline), and the MMA no longer reads 32 KB of P per tile through shared memory.
#pragma unroll
      for (int c = 0; c < KV / 64; ++c) tmem_st32(tP + lane_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&preg[c * 32]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(BAR_P_FULL);
      ATTN_STAMP(7);
    }
    // final: O / l
    mbar_wait(BAR_O_FULL, (uint32_t)(p.nkv - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l;
    const bool row_ok = q0 + r < p.N;
    h16* dst = p.out + (long long)(row0 + q0 + r) * p.D + h * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld32(tO + lane_addr + c * 32, ov);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          uint32_t* uw = &u.x;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float f0 = __uint_as_float(ov[8 * g + 2 * k]) * inv, f1 = __uint_as_float(ov[8 * g + 2 * k + 1]) * inv;
            uw[k] = FP16 ? pack2<FMT_F16>(f0, f1) : pack2<FMT_BF16>(f0, f1);
          }
          *reinterpret_cast<uint4*>(dst + c * 32 + 8 * g) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (KV == 128) {
      tmem_dealloc(tmem_base, 256);
    } else {
      tmem_dealloc(tmem_base, 128);
      tmem_dealloc(tmem_slot_ptr[1], 32);
    }
  }
}


int launch_attention(const h16* qkv, h16* out, int B, int N, int D, int fmt, cudaStream_t stream, uint32_t v_lbo,
                     uint32_t v_sbo) {
  DAV2_CHECK(D % 64 == 0 && N > 0 && B > 0, "attention: bad shape B=%d N=%d D=%d", B, N, D);
  // emu: pairs out of every 8 whose 2^x is emulated on the FMA pipe; kv: keys per tile (128 or 64).  Both were tuned on
  // B200.  The environment overrides exist in profiling builds only (-DDAV2_PROFILING_KNOBS).
  int emu = 2, kv = 64;
  constexpr int SMEM128 = ATT_TILE + 4 * 128 * 128 + 128, SMEM64 = ATT_TILE + 4 * 64 * 128 + 128;
#ifdef DAV2_PROFILING_KNOBS
  if (const char* e = getenv("DAV2_ATTN_EMU")) emu = atoi(e);
  DAV2_CHECK(emu == 0 || emu == 2 || emu == 3 || emu == 4, "DAV2_ATTN_EMU must be 0, 2, 3 or 4");
  if (const char* k = getenv("DAV2_ATTN_KV")) kv = atoi(k);
  DAV2_CHECK(kv == 64 || kv == 128, "DAV2_ATTN_KV must be 64 or 128");
#endif
  static char tag;
  if (!device_setup_done(&tag)) {  // registered only after every attribute call succeeded: a failure is retried
#define DAV2_ATTN_CFG(F, E)                                                                                                  \
  DAV2_CUDA_OK(cudaFuncSetAttribute(attention_kernel<F, E, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM128));     \
  DAV2_CUDA_OK(cudaFuncSetAttribute(attention_kernel<F, E, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM64))
    DAV2_ATTN_CFG(true, 0); DAV2_ATTN_CFG(true, 2); DAV2_ATTN_CFG(true, 3); DAV2_ATTN_CFG(true, 4);
    DAV2_ATTN_CFG(false, 0); DAV2_ATTN_CFG(false, 2); DAV2_ATTN_CFG(false, 3); DAV2_ATTN_CFG(false, 4);
#undef DAV2_ATTN_CFG
    device_setup_mark(&tag);
  }
  CUtensorMap tm, tmkv;
  if (int rc = make_tmap_2d(&tm, qkv, (uint64_t)B * N, (uint64_t)3 * D, (uint64_t)3 * D, 128)) return rc;
  if (int rc = make_tmap_2d(&tmkv, qkv, (uint64_t)B * N, (uint64_t)3 * D, (uint64_t)3 * D, (uint32_t)kv)) return rc;
  AttnParams p;
  p.out = out;
  p.N = N;
  p.D = D;
  p.nkv = (N + kv - 1) / kv;
  p.v_lbo = v_lbo;
  p.v_sbo = v_sbo;
  p.fmt = fmt;
  dim3 grid((N + 127) / 128, D / 64, B);
  ProfScope ps(PC_ATTN, 4.0 * B * (D / 64) * (double)N * N * 64.0, 2.0 * 4.0 * B * (double)N * D, stream);
#define DAV2_ATTN_GO(E)                                                                                  \
  do {                                                                                                   \
    if (kv == 64) {                                                                                      \
      if (fmt == FMT_F16) attention_kernel<true, E, 64><<<grid, 256, SMEM64, stream>>>(tm, tmkv, p);     \
      else attention_kernel<false, E, 64><<<grid, 256, SMEM64, stream>>>(tm, tmkv, p);                   \
    } else {                                                                                             \
      if (fmt == FMT_F16) attention_kernel<true, E, 128><<<grid, 256, SMEM128, stream>>>(tm, tmkv, p);   \
      else attention_kernel<false, E, 128><<<grid, 256, SMEM128, stream>>>(tm, tmkv, p);                 \
    }                                                                                                    \
  } while (0)
  switch (emu) {
    case 0: DAV2_ATTN_GO(0); break;
    case 3: DAV2_ATTN_GO(3); break;
    case 4: DAV2_ATTN_GO(4); break;
    default: DAV2_ATTN_GO(2); break;
  }
#undef DAV2_ATTN_GO
  DAV2_LAUNCH_OK();
  return 0;
}

#ifdef ATTN_TRACE
int set_attn_trace(long long* ptr) {
  DAV2_CUDA_OK(cudaMemcpyToSymbol(g_attn_trace, &ptr, sizeof(ptr)));
  return 0;
}
#endif

}  // namespace dav2

#ifdef ATTN_TRACE
extern "C" int dav2_debug_set_attn_trace(void* ptr) { return dav2::set_attn_trace((long long*)ptr); }
#endif
