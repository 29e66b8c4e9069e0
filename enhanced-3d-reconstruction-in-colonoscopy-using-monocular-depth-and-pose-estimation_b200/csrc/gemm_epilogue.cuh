// Shared epilogue of the tcgen05 GEMM / conv kernels (1-CTA and 2-CTA): one 128-row x BN-column fp32
// accumulator tile in TMEM -> fused op -> global memory.
//
// Everything that varies per LAUNCH but not per element (operand format, GELU yes/no) is a template
// parameter of the tile routine and is dispatched ONCE per tile; ReLU is a branch-free fmax against a
// launch-uniform floor (0 or -inf); optional skip adds / second ReLU output are uniform branches per
// ROW group, never per element.  (ncu on the first version showed the epilogue executing ~5 control
// instructions per useful one because `act` and `fmt` were tested per element.)
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
  float* out_logit;      // GM_CONV_HEAD, optional: pre-sigmoid logits [B,H,W] (parity instrumentation)
};

static constexpr int STG_ROW_BYTES = 36 * 4;  // 32 fp32 + 16 B pad: conflict-free row writes and column-group reads

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (FMT == FMT_BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
template <int FMT>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if constexpr (FMT == FMT_BF16) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
  } else {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
  }
}

struct TileGeom {
  int tm;              // 128-row tile index of this CTA
  int cb_img, x0, y0;  // conv modes: image index and tile origin
};

// Per-tile bias vector (BN floats) in a per-warp smem buffer; read back as warp-wide broadcasts in the
// row layout.  For the transposed-conv scatter the bias index is the output channel (n mod Cout).
template <int BN, int MODE>
__device__ __forceinline__ void epi_fill_bias(const GemmParams& p, uint32_t vec, int lane, int n0) {
#pragma unroll
  for (int j = lane; j < BN; j += 32) {
    const int n = n0 + j;
    float bv = 0.f;
    if (p.bias && n < p.N) {
      int co = n;
      if constexpr (MODE == GM_CONVT) co = n % p.convt_cout;
      bv = __ldg(p.bias + co);
    }
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + j * 4), "f"(bv) : "memory");
  }
  __syncwarp();
}

// One 32-column chunk held by the warp as v[32] (thread = accumulator row):
//   1. bias + activation in the ROW layout (32 independent elements per thread: full ILP for the GELU),
//   2. transpose through the per-warp staging buffer,
//   3. coalesced 8/16-byte global I/O per lane; address / validity / skip-add loads for all 8 row groups are
//      issued first (no branches in between), stores are predicated.
template <int MODE, int FMT, bool GELU>
__device__ __forceinline__ void epi_chunk_store(const GemmParams& p, uint32_t (&v)[32], uint32_t stg, uint32_t vec, int lane,
                                                int q, const TileGeom& g, int nc, int c, float relu_floor) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 b4;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                 : "r"(vec + (c * 32 + 4 * j) * 4) : "memory");
    float a0 = __uint_as_float(v[4 * j]) + b4.x, a1 = __uint_as_float(v[4 * j + 1]) + b4.y;
    float a2 = __uint_as_float(v[4 * j + 2]) + b4.z, a3 = __uint_as_float(v[4 * j + 3]) + b4.w;
    if constexpr (GELU) {
      const float2 g01 = gelu_erf2(make_float2(a0, a1)), g23 = gelu_erf2(make_float2(a2, a3));
      a0 = g01.x; a1 = g01.y; a2 = g23.x; a3 = g23.y;
    } else if constexpr (MODE != GM_LINEAR_RESID && MODE != GM_PATCH) {
      a0 = fmaxf(a0, relu_floor); a1 = fmaxf(a1, relu_floor); a2 = fmaxf(a2, relu_floor); a3 = fmaxf(a3, relu_floor);
    }
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * STG_ROW_BYTES + j * 16), "f"(a0), "f"(a1),
                 "f"(a2), "f"(a3)
                 : "memory");
  }
  __syncwarp();
  const int n = nc + (lane & 7) * 4;
  const bool nvalid = n < p.N;
  int co = n, ky = 0, kx = 0;
  if constexpr (MODE == GM_CONVT) {
    const int kk = n / p.convt_cout;
    co = n - kk * p.convt_cout;
    ky = kk / p.convt_s;
    kx = kk - ky * p.convt_s;
  }
  float4 gam4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if constexpr (MODE == GM_LINEAR_RESID)
    if (nvalid) gam4 = __ldg(reinterpret_cast<const float4*>(p.gamma + n));
  // skip adds / second (ReLU) output only exist on the conv path (DPT residual units)
  constexpr bool EXTRAS = (MODE == GM_CONV_BF16);
  const bool has_add1 = EXTRAS && p.add1 != nullptr, has_add2 = EXTRAS && p.add2 != nullptr;
  const bool has_relu2 = EXTRAS && p.out_relu != nullptr;

  long long off[8];
  bool valid[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rt = q * 32 + i * 4 + (lane >> 3);
    valid[i] = nvalid;
    if constexpr (MODE == GM_CONV_BF16) {
      const int ly = rt / p.tw;
      const int y = g.y0 + ly, x = g.x0 + (rt - ly * p.tw);
      valid[i] = valid[i] && (y < p.H) && (x < p.W);
      off[i] = (((long long)g.cb_img * p.H + y) * p.W + x) * p.ldo + n;
    } else {
      const int m = g.tm * 128 + rt;
      valid[i] = valid[i] && (m < p.M);
      if constexpr (MODE == GM_PATCH) {
        const int bi = m / p.P;
        off[i] = ((long long)bi * (p.P + 1) + 1 + (m - bi * p.P)) * p.ldo + n;
      } else if constexpr (MODE == GM_CONVT) {
        const int hw = p.H * p.W;
        const int bi = m / hw;
        const int rem = m - bi * hw;
        const int y = rem / p.W;
        const int x = rem - y * p.W;
        const int s = p.convt_s;
        off[i] = ((((long long)bi * p.H * s + (y * s + ky)) * (p.W * s)) + (x * s + kx)) * p.convt_cout + co;
      } else {
        off[i] = (long long)m * p.ldo + n;
      }
    }
    if (!valid[i]) off[i] = 0;  // keep speculative address arithmetic in range; accesses stay predicated
  }

  if constexpr (MODE == GM_LINEAR_RESID) {
    float4 xin[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (valid[i]) xin[i] = *reinterpret_cast<const float4*>(reinterpret_cast<float*>(p.out) + off[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 a;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                   : "r"(stg + (i * 4 + (lane >> 3)) * STG_ROW_BYTES + (lane & 7) * 16) : "memory");
      float4 x = xin[i];
      x.x = fmaf(gam4.x, a.x, x.x); x.y = fmaf(gam4.y, a.y, x.y);
      x.z = fmaf(gam4.z, a.z, x.z); x.w = fmaf(gam4.w, a.w, x.w);
      if (valid[i]) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off[i]) = x;
    }
  } else if constexpr (MODE == GM_PATCH) {
    float4 ps[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = g.tm * 128 + q * 32 + i * 4 + (lane >> 3);
      const int pp = m % p.P;
      if (valid[i]) ps[i] = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(1 + pp) * p.ldo + n));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 a;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                   : "r"(stg + (i * 4 + (lane >> 3)) * STG_ROW_BYTES + (lane & 7) * 16) : "memory");
      a.x += ps[i].x; a.y += ps[i].y; a.z += ps[i].z; a.w += ps[i].w;
      if (valid[i]) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off[i]) = a;
    }
  } else {
    uint2 s1[8], s2[8];
    if (has_add1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s1[i] = valid[i] ? __ldg(reinterpret_cast<const uint2*>(p.add1 + off[i])) : make_uint2(0u, 0u);
    }
    if (has_add2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s2[i] = valid[i] ? __ldg(reinterpret_cast<const uint2*>(p.add2 + off[i])) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 a;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                   : "r"(stg + (i * 4 + (lane >> 3)) * STG_ROW_BYTES + (lane & 7) * 16) : "memory");
      if (has_add1) {
        const float2 f0 = unpack2<FMT>(s1[i].x), f1 = unpack2<FMT>(s1[i].y);
        a.x += f0.x; a.y += f0.y; a.z += f1.x; a.w += f1.y;
      }
      if (has_add2) {
        const float2 f0 = unpack2<FMT>(s2[i].x), f1 = unpack2<FMT>(s2[i].y);
        a.x += f0.x; a.y += f0.y; a.z += f1.x; a.w += f1.y;
      }
      uint2 o;
      o.x = pack2<FMT>(a.x, a.y);
      o.y = pack2<FMT>(a.z, a.w);
      if (valid[i]) *reinterpret_cast<uint2*>(reinterpret_cast<h16*>(p.out) + off[i]) = o;
      if (has_relu2) {
        uint2 orl;
        orl.x = pack2<FMT>(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
        orl.y = pack2<FMT>(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
        if (valid[i]) *reinterpret_cast<uint2*>(p.out_relu + off[i]) = orl;
      }
    }
  }
  __syncwarp();
}

// All chunks of one accumulator tile for this warp (rows q*32 .. q*32+31).  `vec` must hold the tile's bias.
template <int BN, int MODE, int FMT, bool GELU>
__device__ __forceinline__ void epi_tile_store(const GemmParams& p, uint32_t t_row, uint32_t stg, uint32_t vec, int lane, int q,
                                               const TileGeom& g, int n0) {
  const float relu_floor = (p.act == 2) ? 0.0f : -INFINITY;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int nc = n0 + c * 32;
    if (nc >= p.N) break;  // warp-uniform
    uint32_t v[32];
    tmem_ld32(t_row + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    epi_chunk_store<MODE, FMT, GELU>(p, v, stg, vec, lane, q, g, nc, c, relu_floor);
  }
}

// Launch-uniform dispatch on (format, GELU) -- once per tile.
template <int BN, int MODE>
__device__ __forceinline__ void epi_tile_dispatch(const GemmParams& p, uint32_t t_row, uint32_t stg, uint32_t vec, int lane,
                                                  int q, const TileGeom& g, int n0) {
  if constexpr (MODE == GM_LINEAR_RESID || MODE == GM_PATCH) {
    epi_tile_store<BN, MODE, FMT_F16, false>(p, t_row, stg, vec, lane, q, g, n0);  // fp32 output: format unused
  } else if constexpr (MODE == GM_LINEAR_BF16) {
    if (p.fmt == FMT_BF16) {
      if (p.act == 1) epi_tile_store<BN, MODE, FMT_BF16, true>(p, t_row, stg, vec, lane, q, g, n0);
      else epi_tile_store<BN, MODE, FMT_BF16, false>(p, t_row, stg, vec, lane, q, g, n0);
    } else {
      if (p.act == 1) epi_tile_store<BN, MODE, FMT_F16, true>(p, t_row, stg, vec, lane, q, g, n0);
      else epi_tile_store<BN, MODE, FMT_F16, false>(p, t_row, stg, vec, lane, q, g, n0);
    }
  } else {
    if (p.fmt == FMT_BF16) epi_tile_store<BN, MODE, FMT_BF16, false>(p, t_row, stg, vec, lane, q, g, n0);
    else epi_tile_store<BN, MODE, FMT_F16, false>(p, t_row, stg, vec, lane, q, g, n0);
  }
}

// 3x3 conv (N = 32) + ReLU + 1x1 (32 -> 1) + sigmoid * max_depth: thread = pixel, whole row in registers.
__device__ __forceinline__ void epi_tile_head(const GemmParams& p, uint32_t t_row, int lane, int q, const TileGeom& g) {
  uint32_t v[32];
  tmem_ld32(t_row, v);
  tmem_ld_wait();
  const int rt = q * 32 + lane;
  const int ly = rt / p.tw;
  const int y = g.y0 + ly, x = g.x0 + (rt - ly * p.tw);
  float acc = p.head_b;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    acc = fmaf(fmaxf(__uint_as_float(v[j]) + __ldg(p.bias + j), 0.0f), __ldg(p.head_w + j), acc);
  if (y < p.H && x < p.W) {
    const float s = 1.0f / (1.0f + __expf(-acc));
    const long long idx = ((long long)g.cb_img * p.H + y) * p.W + x;
    reinterpret_cast<float*>(p.out)[idx] = s * p.max_depth;
    if (p.out_logit) p.out_logit[idx] = acc;  // launch-uniform
  }
}

}  // namespace dav2
