// HBM-bound helper kernels of the DepthAnythingV2 forward: LayerNorm (warp-shuffle), patch im2col,
// cls-token row, strided-conv im2col, bilinear (align_corners=True) NHWC resampler, depth resampler.
// All are coalesced + 16-byte vectorised; grids are sized in multiples of the SM count.
#include "elementwise.cuh"
#include "gemm_epilogue.cuh"

namespace dav2 {

// ----------------------------------------------------------------------------------------------
// LayerNorm over the last dim (D = 128*V4, V4 <= 8), fp32 in -> h16 out, one warp per row.
// rows_per_img / skip_cls: when skip_cls != 0 the first row of every image (cls token) is dropped and
// the remaining rows are written compactly ([B, N-1, D] == NHWC patch grid) - used for the DPT taps.
// ----------------------------------------------------------------------------------------------
template <int V4>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, h16* __restrict__ out,
                                                        long long rows, int rows_per_img, int skip_cls, float eps, int fmt) {
  constexpr int D = V4 * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp0; row < rows; row += nwarps) {
    long long orow = row;
    if (skip_cls) {
      const long long img = row / rows_per_img;
      const int t = (int)(row - img * rows_per_img);
      if (t == 0) continue;
      orow = img * (rows_per_img - 1) + (t - 1);
    }
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[V4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i] = xr[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + c * c) + (d * d + e * e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + eps);
    uint2* orp = reinterpret_cast<uint2*>(out + orow * D);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
      uint2 u;
      u.x = pack_h2((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y, fmt);
      u.y = pack_h2((v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w, fmt);
      orp[lane + 32 * i] = u;
    }
  }
}

int launch_layernorm(const float* x, const float* w, const float* b, h16* out, long long rows, int D,
                     int rows_per_img, int skip_cls, float eps, int fmt, cudaStream_t stream) {
  DAV2_CHECK(D % 128 == 0 && D / 128 <= 8, "layernorm: D=%d unsupported (need multiple of 128, <= 1024)", D);
  if (rows <= 0) return 0;
  const int wpb = 8;
  long long blocks = (rows + wpb - 1) / wpb;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ProfScope ps(PC_LAYERNORM, 0.0, (double)rows * D * 6.0, stream);
#define LN_CASE(V)                                                                                             \
  case V:                                                                                                      \
    layernorm_kernel<V><<<(int)blocks, wpb * 32, 0, stream>>>(x, w, b, out, rows, rows_per_img, skip_cls, eps, fmt); \
    break;
  switch (D / 128) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
  }
#undef LN_CASE
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Patch im2col for the 14x14 / stride-14 patch embedding: x fp32 NCHW [B,3,H,W] -> A h16 [B*ph*pw, KP]
// with k = c*196 + ky*14 + kx (the flattening of Conv2d weight [D,3,14,14]); columns 588..KP-1 are zero.
// One thread per (patch, c, ky) segment of 14 contiguous input floats.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_im2col_kernel(const float* __restrict__ x, h16* __restrict__ A, int B,
                                                           int H, int W, int ph, int pw, int KP, int fmt) {
  const long long nseg = (long long)B * ph * pw * 43;  // 42 data segments + 1 zero-pad segment
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nseg; i += (long long)gridDim.x * blockDim.x) {
    const long long patch = i / 43;
    const int seg = (int)(i - patch * 43);
    h16* dst = A + patch * KP;
    if (seg == 42) {
      for (int k = 588; k < KP; k += 2) *reinterpret_cast<uint32_t*>(dst + k) = 0u;
      continue;
    }
    const int c = seg / 14, ky = seg - c * 14;
    const int bimg = (int)(patch / (ph * pw));
    const int pr = (int)(patch - (long long)bimg * ph * pw);
    const int py = pr / pw, px = pr - py * pw;
    const float* src = x + (((long long)bimg * 3 + c) * H + (py * 14 + ky)) * W + px * 14;
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + seg * 14);
#pragma unroll
    for (int k = 0; k < 7; ++k) d32[k] = pack_h2(__ldg(src + 2 * k), __ldg(src + 2 * k + 1), fmt);
  }
}

int launch_patch_im2col(const float* x, h16* A, int B, int H, int W, int KP, int fmt, cudaStream_t stream) {
  DAV2_CHECK(H % 14 == 0 && W % 14 == 0 && KP >= 588 && KP % 64 == 0, "patch_im2col: bad shape H=%d W=%d KP=%d", H, W, KP);
  const int ph = H / 14, pw = W / 14;
  const long long nseg = (long long)B * ph * pw * 43;
  long long blocks = (nseg + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ProfScope ps(PC_IM2COL, 0.0, (double)B * 3 * H * W * 4.0 + (double)B * ph * pw * KP * 2.0, stream);
  patch_im2col_kernel<<<(int)blocks, 256, 0, stream>>>(x, A, B, H, W, ph, pw, KP, fmt);
  DAV2_LAUNCH_OK();
  return 0;
}

// cls row: x[b, 0, :] = cls + pos[0, :]
__global__ void cls_row_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                               int B, int ntok, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  x[(long long)b * ntok * D + d] = cls[d] + pos[d];
}

int launch_cls_row(float* x, const float* cls, const float* pos, int B, int ntok, int D, cudaStream_t stream) {
  ProfScope ps(PC_OTHER, 0.0, (double)B * D * 12.0, stream);
  cls_row_kernel<<<(B * D + 255) / 256, 256, 0, stream>>>(x, cls, pos, B, ntok, D);
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// im2col for Conv2d(k=3, stride=2, pad=1) on NHWC h16: [B,H,W,C] -> [B*Ho*Wo, 9*C], k = tap*C + c.
// One thread per 8-channel (16 B) vector.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_s2_kernel(const h16* __restrict__ in, h16* __restrict__ A, int B, int H,
                                                        int W, int C, int Ho, int Wo) {
  const int c8 = C / 8;
  const long long total = (long long)B * Ho * Wo * 9 * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    long long t = i / c8;
    const int tap = (int)(t % 9);
    t /= 9;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int yi = yo * 2 - 1 + tap / 3, xi = xo * 2 - 1 + tap % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yi >= 0 && yi < H && xi >= 0 && xi < W)
      v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)b * H + yi) * W + xi) * C) + cv);
    reinterpret_cast<uint4*>(A)[i] = v;
  }
}

int launch_im2col_s2(const h16* in, h16* A, int B, int H, int W, int C, cudaStream_t stream) {
  DAV2_CHECK(C % 8 == 0, "im2col_s2: C=%d must be a multiple of 8", C);
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)B * Ho * Wo * 9 * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ProfScope ps(PC_IM2COL, 0.0, (double)total * 32.0, stream);
  im2col_s2_kernel<<<(int)blocks, 256, 0, stream>>>(in, A, B, H, W, C, Ho, Wo);
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Bilinear resize, align_corners=True, NHWC h16 -> NHWC h16 (fp32 blend), 8 channels per thread.
// src coordinate = dst * (in-1)/(out-1), exactly like F.interpolate(..., align_corners=True).
// ----------------------------------------------------------------------------------------------
static constexpr int BILINEAR_ROWS = 8;

template <int FMT>
__global__ void __launch_bounds__(256) bilinear_nhwc_kernel(const h16* __restrict__ in, h16* __restrict__ out, int Hi, int Wi,
                                                            int Ho, int Wo, int C, float sy, float sx) {
  // grid = (ceil(Wo*C/8 / 256), ceil(Ho / BILINEAR_ROWS), B); one thread = 8 channels of one output column for
  // BILINEAR_ROWS consecutive output rows.
  // The first versions evaluated the four-tap weighted sum from scratch per output vector: 190 instructions per 16 output
  // bytes (32 half->float conversions, 38 FMAs, 64-bit address arithmetic for four loads) -- the kernel was ISSUE bound
  // at 2.9-3.5 TB/s.  Now the interpolation is separable: a source row is loaded once, interpolated horizontally into
  // eight fp32 registers, and kept while the output rows that use it go by (an up-sample advances the source row by 0 or
  // 1 per output row); an output vector then costs one vertical lerp (16 FMAs), four packs and one store, and the
  // addresses advance by constant strides.
  const int c8 = C >> 3;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Wo * c8) return;
  const int xo = t / c8, cv = t - xo * c8;
  const int b = blockIdx.z;
  const float fx = sx * xo;
  const int x0 = min((int)fx, Wi - 1);
  const int x1 = min(x0 + 1, Wi - 1);
  const float wx = fx - (float)x0;
  const int rstride = Wi * c8;  // uint4 units between source rows (one image: fits 32 bits, checked by the launcher)
  const uint4* col0 = reinterpret_cast<const uint4*>(in + ((long long)b * Hi * Wi + x0) * C) + cv;
  const uint4* col1 = reinterpret_cast<const uint4*>(in + ((long long)b * Hi * Wi + x1) * C) + cv;
  const int yo0 = blockIdx.y * BILINEAR_ROWS;
  uint4* optr = reinterpret_cast<uint4*>(out + (((long long)b * Ho + yo0) * Wo + xo) * C) + cv;
  const long long ostride = (long long)Wo * c8;

  float top[8], bot[8];
  int yt = -2, yb = -2;  // source rows held in top / bot
  auto load_row = [&](int y, float (&dst)[8]) {
    const int off = y * rstride;
    const uint4 pa = __ldg(col0 + off), pb = __ldg(col1 + off);
    const uint32_t* a = &pa.x;
    const uint32_t* bq = &pb.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack2<FMT>(a[k]), fb = unpack2<FMT>(bq[k]);
      dst[2 * k] = fmaf(wx, fb.x - fa.x, fa.x);
      dst[2 * k + 1] = fmaf(wx, fb.y - fa.y, fa.y);
    }
  };
#pragma unroll
  for (int rr = 0; rr < BILINEAR_ROWS; ++rr) {
    const int yo = yo0 + rr;
    if (yo >= Ho) break;
    const float fy = sy * yo;
    const int y0 = min((int)fy, Hi - 1);
    const int y1 = min(y0 + 1, Hi - 1);
    const float wy = fy - (float)y0;
    if (y0 != yt) {  // block-uniform
      if (y0 == yb) {
#pragma unroll
        for (int k = 0; k < 8; ++k) top[k] = bot[k];
      } else {
        load_row(y0, top);
      }
      yt = y0;
    }
    if (y1 != yb) {
      if (y1 == yt) {
#pragma unroll
        for (int k = 0; k < 8; ++k) bot[k] = top[k];
      } else {
        load_row(y1, bot);
      }
      yb = y1;
    }
    uint4 o;
    uint32_t* ow = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      ow[k] = pack2<FMT>(fmaf(wy, bot[2 * k] - top[2 * k], top[2 * k]), fmaf(wy, bot[2 * k + 1] - top[2 * k + 1], top[2 * k + 1]));
    __stcs(optr, o);
    optr += ostride;
  }
}

int launch_bilinear_nhwc(const h16* in, h16* out, int B, int Hi, int Wi, int Ho, int Wo, int C, int fmt, cudaStream_t stream) {
  DAV2_CHECK(C % 8 == 0, "bilinear: C=%d must be a multiple of 8", C);
  DAV2_CHECK(Ho <= 65535 && B <= 65535, "bilinear: Ho / B exceed the grid limits");
  DAV2_CHECK((long long)Hi * Wi * (C / 8) < (1ll << 31), "bilinear: one source image exceeds 2^31 16-byte vectors");
  const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  dim3 grid((unsigned)((Wo * (C / 8) + 255) / 256), (unsigned)((Ho + BILINEAR_ROWS - 1) / BILINEAR_ROWS), (unsigned)B);
  ProfScope ps(PC_RESAMPLE, 0.0, 2.0 * C * ((double)B * Hi * Wi + (double)B * Ho * Wo), stream);
  if (fmt == FMT_BF16) bilinear_nhwc_kernel<FMT_BF16><<<grid, 256, 0, stream>>>(in, out, Hi, Wi, Ho, Wo, C, sy, sx);
  else bilinear_nhwc_kernel<FMT_F16><<<grid, 256, 0, stream>>>(in, out, Hi, Wi, Ho, Wo, C, sy, sx);
  DAV2_LAUNCH_OK();
  return 0;
}

// fp32 single-channel bilinear (align_corners=True): depth [B,Hi,Wi] -> [B,Ho,Wo]  (infer_image resize-back)
__global__ void __launch_bounds__(256) bilinear_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                           int Hi, int Wi, int Ho, int Wo, float sy, float sx) {
  const long long total = (long long)B * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo);
    long long t = i / Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const float fy = sy * yo, fx = sx * xo;
    int y0 = min((int)fy, Hi - 1), x0 = min((int)fx, Wi - 1);
    const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
    const float wy = fy - (float)y0, wx = fx - (float)x0;
    const float* base = in + (long long)b * Hi * Wi;
    const float v00 = __ldg(base + (long long)y0 * Wi + x0), v01 = __ldg(base + (long long)y0 * Wi + x1);
    const float v10 = __ldg(base + (long long)y1 * Wi + x0), v11 = __ldg(base + (long long)y1 * Wi + x1);
    out[i] = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
  }
}

int launch_bilinear_f32(const float* in, float* out, int B, int Hi, int Wi, int Ho, int Wo, cudaStream_t stream) {
  const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)B * Ho * Wo;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 0;
  ProfScope ps(PC_RESAMPLE, 0.0, 4.0 * ((double)B * Hi * Wi + (double)B * Ho * Wo), stream);
  bilinear_f32_kernel<<<(int)blocks, 256, 0, stream>>>(in, out, B, Hi, Wi, Ho, Wo, sy, sx);
  DAV2_LAUNCH_OK();
  return 0;
}

}  // namespace dav2

namespace dav2 {

// ----------------------------------------------------------------------------------------------
// Upstream image2tensor on the GPU (external dpt.py / util/transform.py; reference call run.py:233-234):
// BGR uint8 [H,W,3] -> RGB/255 -> cv2.resize(INTER_CUBIC) to (nh, nw) -> (x - mean)/std -> fp32 CHW.
// OpenCV's bicubic: src = (dst + 0.5) * scale - 0.5, taps floor(src)-1 .. +2 with replicated borders,
// A = -0.75, fp32 coefficient tables, separable, fp64 pixels (the reference divides by 255.0 in fp64).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float x, float* c) {
  const float A = -0.75f;
  c[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  c[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
  c[3] = 1.f - c[0] - c[1] - c[2];
}

__global__ void __launch_bounds__(256) preprocess_bgr_kernel(const uint8_t* __restrict__ img, int H, int W, float* __restrict__ out,
                                                             int nh, int nw, double sy, double sx) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nh * nw) return;
  img += (long long)blockIdx.y * H * W * 3;       // frame blockIdx.y of the batch
  out += (long long)blockIdx.y * 3 * nh * nw;
  const int dy = t / nw, dx = t - dy * nw;
  float cy[4], cx[4];
  int iy[4], ix[4];
  if (nh == H) {  // cv2.resize with identical size is a copy
    cy[0] = 0.f; cy[1] = 1.f; cy[2] = 0.f; cy[3] = 0.f;
    for (int k = 0; k < 4; ++k) iy[k] = min(max(dy - 1 + k, 0), H - 1);
  } else {
    float fy = (float)((dy + 0.5) * sy - 0.5);
    const int s = (int)floorf(fy);
    fy -= (float)s;
    cubic_coeffs(fy, cy);
    for (int k = 0; k < 4; ++k) iy[k] = min(max(s - 1 + k, 0), H - 1);
  }
  if (nw == W) {
    cx[0] = 0.f; cx[1] = 1.f; cx[2] = 0.f; cx[3] = 0.f;
    for (int k = 0; k < 4; ++k) ix[k] = min(max(dx - 1 + k, 0), W - 1);
  } else {
    float fx = (float)((dx + 0.5) * sx - 0.5);
    const int s = (int)floorf(fx);
    fx -= (float)s;
    cubic_coeffs(fx, cx);
    for (int k = 0; k < 4; ++k) ix[k] = min(max(s - 1 + k, 0), W - 1);
  }
  double acc[3] = {0.0, 0.0, 0.0};
  for (int r = 0; r < 4; ++r) {
    double row[3] = {0.0, 0.0, 0.0};
    const uint8_t* line = img + (long long)iy[r] * W * 3;
    for (int k = 0; k < 4; ++k) {
      const uint8_t* px = line + ix[k] * 3;
      row[0] += (double)cx[k] * ((double)px[2] / 255.0);  // R
      row[1] += (double)cx[k] * ((double)px[1] / 255.0);  // G
      row[2] += (double)cx[k] * ((double)px[0] / 255.0);  // B
    }
    acc[0] += (double)cy[r] * row[0];
    acc[1] += (double)cy[r] * row[1];
    acc[2] += (double)cy[r] * row[2];
  }
  const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
  const long long plane = (long long)nh * nw;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[c * plane + t] = (float)((acc[c] - mean[c]) / stdv[c]);
}

int launch_preprocess_bgr(const uint8_t* img, int B, int H, int W, float* out, int nh, int nw, cudaStream_t stream) {
  DAV2_CHECK(img && out && B > 0 && H > 0 && W > 0 && nh > 0 && nw > 0, "preprocess: bad arguments");
  ProfScope ps(PC_RESAMPLE, 0.0, (double)B * ((double)H * W * 3.0 + (double)nh * nw * 12.0), stream);
  dim3 grid((unsigned)((nh * nw + 255) / 256), (unsigned)B);
  preprocess_bgr_kernel<<<grid, 256, 0, stream>>>(img, H, W, out, nh, nw, (double)H / nh, (double)W / nw);
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Dataset pre-processing (data_processing/simcol.py:104-135,161-168): ToTensor -> Resize((S, S), BICUBIC, antialias=True)
// [-> Normalize], i.e. torch's anti-aliased bicubic (aten _upsample_bicubic2d_aa): per output index i along an axis
//   scale = in / out, support = 2 * max(scale, 1), center = scale * (i + 0.5),
//   taps j in [xmin, xmin + xsize): xmin = max(int(center - support + 0.5), 0), xsize = min(int(center + support + 0.5), in) - xmin
//   weight_j = cubic_{a = -0.5}((j - center + 0.5) / max(scale, 1)), normalised to sum 1 (borders truncate, no clamping).
// One thread per output pixel, fp32 arithmetic like the CPU transform; the weights are recomputed in the tap loops (no
// per-thread arrays, any scale).  Inputs are divided by div_in with an IEEE division, like `image.astype(float32) / 255.0`.
// MODE 0: u8 [B,H,W,3] RGB / 255, ImageNet normalisation -> [B,3,Ho,Wo];
// MODE 1: u16 [B,H,W] / 65535 -> [B,1,Ho,Wo];  MODE 2: fp32 [B,H,W] / div_in -> [B,1,Ho,Wo].
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}
struct AaAxis {
  int lo, n;
  float center, inv, total;
};
__device__ __forceinline__ AaAxis aa_axis(int i, int in, float scale) {
  AaAxis ax;
  const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
  ax.inv = scale >= 1.0f ? 1.0f / scale : 1.0f;
  // __fmul_rn: the product must be ROUNDED to fp32 like ATen's `center`; left to the compiler, `tap - scale * (i + 0.5)` is
  // contracted into one FMA with an unrounded product -- half an ulp of 474 = 1.5e-5 in every weight argument
  ax.center = __fmul_rn(scale, (float)i + 0.5f);
  ax.lo = max((int)(ax.center - support + 0.5f), 0);
  ax.n = min((int)(ax.center + support + 0.5f), in) - ax.lo;
  float t = 0.f;
  for (int j = 0; j < ax.n; ++j) t += cubic_aa(((float)(j + ax.lo) - ax.center + 0.5f) * ax.inv);
  ax.total = t;
  return ax;
}

template <int MODE>
__global__ void __launch_bounds__(256) resample_aa_kernel(const void* __restrict__ in_, int H, int W, float* __restrict__ out,
                                                          int Ho, int Wo, float sy, float sx, float div_in) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Ho * Wo) return;
  const int oy = t / Wo, ox = t - oy * Wo;
  const AaAxis ay = aa_axis(oy, H, sy), ax = aa_axis(ox, W, sx);
  constexpr int C = MODE == 0 ? 3 : 1;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  const long long frame = (long long)blockIdx.y * H * W;
  for (int r = 0; r < ay.n; ++r) {
    const float wy = cubic_aa(((float)(r + ay.lo) - ay.center + 0.5f) * ay.inv) / ay.total;
    float row[C];
#pragma unroll
    for (int c = 0; c < C; ++c) row[c] = 0.f;
    const long long line = frame + (long long)(r + ay.lo) * W;
    for (int k = 0; k < ax.n; ++k) {
      const float wx = cubic_aa(((float)(k + ax.lo) - ax.center + 0.5f) * ax.inv) / ax.total;
      const long long px = line + k + ax.lo;
      if (MODE == 0) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(in_) + px * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] = fmaf(wx, (float)p[c] / div_in, row[c]);
      } else if (MODE == 1) {
        row[0] = fmaf(wx, (float)reinterpret_cast<const uint16_t*>(in_)[px] / div_in, row[0]);
      } else {
        row[0] = fmaf(wx, reinterpret_cast<const float*>(in_)[px] / div_in, row[0]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = fmaf(wy, row[c], acc[c]);
  }
  const long long plane = (long long)Ho * Wo;
  float* o = out + (long long)blockIdx.y * C * plane + t;
  if (MODE == 0) {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * plane] = (acc[c] - mean[c]) / stdv[c];
  } else {
    o[0] = acc[0];
  }
}

int launch_resample_aa(int mode, const void* in, int B, int H, int W, float* out, int Ho, int Wo, float div_in,
                       cudaStream_t stream) {
  DAV2_CHECK(in && out && B > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && div_in != 0.f, "resample_aa: bad arguments");
  DAV2_CHECK(mode >= 0 && mode <= 2, "resample_aa: mode must be 0 (u8 RGB image), 1 (u16 depth) or 2 (fp32 depth)");
  const int cin = mode == 0 ? 3 : (mode == 1 ? 2 : 4), cout = mode == 0 ? 12 : 4;
  ProfScope ps(PC_RESAMPLE, 0.0, (double)B * ((double)H * W * cin + (double)Ho * Wo * cout), stream);
  dim3 grid((unsigned)((Ho * Wo + 255) / 256), (unsigned)B);
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  if (mode == 0) resample_aa_kernel<0><<<grid, 256, 0, stream>>>(in, H, W, out, Ho, Wo, sy, sx, div_in);
  else if (mode == 1) resample_aa_kernel<1><<<grid, 256, 0, stream>>>(in, H, W, out, Ho, Wo, sy, sx, div_in);
  else resample_aa_kernel<2><<<grid, 256, 0, stream>>>(in, H, W, out, Ho, Wo, sy, sx, div_in);
  DAV2_LAUNCH_OK();
  return 0;
}

}  // namespace dav2
