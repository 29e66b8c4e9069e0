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

#include "common.cuh"

namespace dav2 {

enum GemmMode {
  GM_LINEAR_BF16 = 0,  // out h16 = act(acc+bias) [+add1][+add2]; optional second output relu(out)
  GM_LINEAR_RESID = 1, // x(fp32) += gamma * (acc + bias)               (LayerScale + residual)
  GM_PATCH = 2,        // x(fp32)[b, 1+p, :] = acc + bias + pos[1+p, :]   (patch embed + pos embed)
  GM_CONVT = 3,        // ConvTranspose2d(k = s): h16 scatter to (s*y+ky, s*x+kx), bias per out channel
  GM_CONV_BF16 = 4,    // 3x3 pad-1 conv, NHWC h16 out, same epilogue options as GM_LINEAR_BF16
  GM_CONV_HEAD = 5,    // 3x3 conv (N=32) + ReLU + 1x1 (32->1) + sigmoid * max_depth -> fp32 depth
};

struct GemmParams {
  int M, N, K;
  int num_kb;            // K / 64 (conv: 9 * cblocks)
  int tiles_m, tiles_n;
  // conv geometry (GM_CONV_*): image H x W, tile tw x th (tw*th == 128)
  int H, W, tw, th, tiles_x, tiles_y, cblocks;
  // epilogue
  void* out;
  long long ldo;
  h16* out_relu;
  const float* bias;
  const float* gamma;
  const h16* add1;
  const h16* add2;
  const float* pos;
  int P;                 // patches per image (GM_PATCH)
  int act;               // 0 none, 1 GELU(erf), 2 ReLU
  int fmt;               // FMT_F16 / FMT_BF16: operand + 16-bit output format
  int convt_s, convt_cout;
  const float* head_w;   // [32]
  float head_b, max_depth;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_ROW_BYTES = 36 * 4;  // 32 fp32 + 16 B pad: conflict-free both ways
  static constexpr int STAGING_BYTES = 4 * 32 * STG_ROW_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return gelu_erf(v);
  if (act == 2) return fmaxf(v, 0.0f);
  return v;
}

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
  const uint32_t bars = staging + Cfg::STAGING_BYTES;
  const uint32_t tmem_slot = bars + 8 * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES + 8 * (2 * STAGES + 4));
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

  if (threadIdx.x == 0) {
    // ================================ TMA producer =========================================
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
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (threadIdx.x == 32) {
    // ================================ MMA issuer ============================================
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
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_h16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
        umma_commit(EMPTY_BAR(stage));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(TFULL_BAR(as));
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp >= 2) {
    // ================================ epilogue ==============================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t stg = staging + (uint32_t)(warp - 2) * 32u * Cfg::STG_ROW_BYTES;
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
      mbar_wait(TFULL_BAR(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);

      if constexpr (MODE == GM_CONV_HEAD) {
        static_assert(MODE != GM_CONV_HEAD || BN == 32, "head epilogue expects N tile 32");
        uint32_t v[32];
        tmem_ld32(t_row, v);
        tmem_ld_wait();
        const int rt = q * 32 + lane;
        const int ly = rt / p.tw;
        const int y = y0 + ly, x = x0 + (rt - ly * p.tw);
        float acc = p.head_b;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          acc = fmaf(fmaxf(__uint_as_float(v[j]) + __ldg(p.bias + j), 0.0f), __ldg(p.head_w + j), acc);
        if (y < p.H && x < p.W) {
          const float s = 1.0f / (1.0f + __expf(-acc));
          reinterpret_cast<float*>(p.out)[((long long)cb_img * p.H + y) * p.W + x] = s * p.max_depth;
        }
      } else {
        for (int c = 0; c < BN / 32; ++c) {
          const int nc = tn * BN + c * 32;
          if (nc >= p.N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)(c * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * Cfg::STG_ROW_BYTES + j * 16),
                         "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                         : "memory");
          __syncwarp();
          const int n = nc + (lane & 7) * 4;
          const bool nvalid = n < p.N;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 gam4 = make_float4(1.f, 1.f, 1.f, 1.f);
          int co = n;  // bias / channel index
          int ky = 0, kx = 0;
          if constexpr (MODE == GM_CONVT) {
            const int kk = n / p.convt_cout;
            co = n - kk * p.convt_cout;
            ky = kk / p.convt_s;
            kx = kk - ky * p.convt_s;
          }
          if (nvalid) {
            if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + co));
            if constexpr (MODE == GM_LINEAR_RESID) gam4 = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            const int rt = q * 32 + r;
            float4 a;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                         : "r"(stg + r * Cfg::STG_ROW_BYTES + (lane & 7) * 16)
                         : "memory");
            a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
            bool valid = nvalid;
            long long off = 0;
            if constexpr (MODE == GM_CONV_BF16) {
              const int ly = rt / p.tw;
              const int y = y0 + ly, x = x0 + (rt - ly * p.tw);
              valid = valid && (y < p.H) && (x < p.W);
              off = (((long long)cb_img * p.H + y) * p.W + x) * p.ldo + n;
            } else {
              const int m = tm * 128 + rt;
              valid = valid && (m < p.M);
              if constexpr (MODE == GM_PATCH) {
                const int bi = m / p.P;
                const int pp = m - bi * p.P;
                off = ((long long)bi * (p.P + 1) + 1 + pp) * p.ldo + n;
                if (valid) {
                  const float4 ps = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(1 + pp) * p.ldo + n));
                  a.x += ps.x; a.y += ps.y; a.z += ps.z; a.w += ps.w;
                }
              } else if constexpr (MODE == GM_CONVT) {
                const int hw = p.H * p.W;
                const int bi = m / hw;
                const int rem = m - bi * hw;
                const int y = rem / p.W;
                const int x = rem - y * p.W;
                const int s = p.convt_s;
                off = ((((long long)bi * p.H * s + (y * s + ky)) * (p.W * s)) + (x * s + kx)) * p.convt_cout + co;
              } else {
                off = (long long)m * p.ldo + n;
              }
            }
            if (!valid) continue;
            if constexpr (MODE == GM_LINEAR_RESID) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
              float4 x = *dst;
              x.x = fmaf(gam4.x, a.x, x.x); x.y = fmaf(gam4.y, a.y, x.y);
              x.z = fmaf(gam4.z, a.z, x.z); x.w = fmaf(gam4.w, a.w, x.w);
              *dst = x;
            } else if constexpr (MODE == GM_PATCH) {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = a;
            } else {
              a.x = apply_act(a.x, p.act); a.y = apply_act(a.y, p.act);
              a.z = apply_act(a.z, p.act); a.w = apply_act(a.w, p.act);
              if (p.add1) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(p.add1 + off));
                const float2 f0 = unpack_h2(u.x, p.fmt), f1 = unpack_h2(u.y, p.fmt);
                a.x += f0.x; a.y += f0.y; a.z += f1.x; a.w += f1.y;
              }
              if (p.add2) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(p.add2 + off));
                const float2 f0 = unpack_h2(u.x, p.fmt), f1 = unpack_h2(u.y, p.fmt);
                a.x += f0.x; a.y += f0.y; a.z += f1.x; a.w += f1.y;
              }
              uint2 o;
              o.x = pack_h2(a.x, a.y, p.fmt);
              o.y = pack_h2(a.z, a.w, p.fmt);
              *reinterpret_cast<uint2*>(reinterpret_cast<h16*>(p.out) + off) = o;
              if (p.out_relu) {
                uint2 orl;
                orl.x = pack_h2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), p.fmt);
                orl.y = pack_h2(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f), p.fmt);
                *reinterpret_cast<uint2*>(p.out_relu + off) = orl;
              }
            }
          }
          __syncwarp();
        }
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
