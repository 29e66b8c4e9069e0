// fp32 validation engine (precision = 2): the whole DepthAnythingV2 forward in fp32 SIMT arithmetic.
//
// north_star asks for depth within 1e-4 of the reference "in fp32 mode"; 16-bit tensor-core operands cannot get
// there (2^-11 per rounding through ~70 layers).  This path is the parity instrument: same weights, same layer
// order and the same algebraic re-orderings as the tensor-core engine (engine.cu), but every contraction is an
// fp32 FMA chain -- one generic register-tiled SGEMM whose A operand is gathered on the fly (dense rows, 3x3
// conv taps of an NHWC map, 14x14 patches of the NCHW image), a two-pass softmax attention, and fp32 LayerNorm /
// bilinear kernels.  It is ~100x slower than the tcgen05 engine and is not the benchmarked path.
// Reference: the external dpt.DepthAnythingV2.forward driven at run.py:234 / lightning_model.py:301.
#include "engine.cuh"

#include <math.h>
#include <string.h>

namespace dav2 {

enum { GA_DENSE = 0, GA_CONV3 = 1, GA_PATCH = 2 };

struct SgemmParams {
  const float* A;     // dense: [M, lda]; conv: NHWC [B,H,W,Cin]; patch: NCHW image [B,3,H,W]
  const float* Wt;    // [N, K] row-major (K ordered like the gather)
  const float* bias;  // [N] (convT: [convt_cout]) or null
  float* C;           // [M, ldc] (or scattered, see convt_s / tok_P)
  int M, N, K;
  long long lda, ldc;
  int gather;
  int H, W, Cin, stride, Ho, Wo;  // conv / patch geometry (patch: Ho x Wo patch grid)
  int act;                        // 0 none, 1 GELU(erf), 2 ReLU
  const float* gamma;             // LayerScale [N] or null
  int accumulate;                 // C += result (residual stream)
  const float* add1;              // optional skip tensors [M, ldc]
  const float* add2;
  float* out_relu;                // optional second output relu(result)
  int convt_s, convt_cout;        // ConvTranspose2d(k = s, stride = s) pixel-shuffle scatter: n = (ky*s+kx)*cout + co
  int tok_P;                      // patch embed: row m = (b, p) goes to token row b*(P+1) + 1 + p, plus pos[(1+p)*N + n]
  const float* pos;
};

__device__ __forceinline__ float sgemm_load_a(const SgemmParams& p, int m, int k) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (p.gather == GA_DENSE) return p.A[(long long)m * p.lda + k];
  if (p.gather == GA_CONV3) {
    const int ox = m % p.Wo, oy = (m / p.Wo) % p.Ho, b = m / (p.Wo * p.Ho);
    const int tap = k / p.Cin, c = k - tap * p.Cin;
    const int iy = oy * p.stride + tap / 3 - 1, ix = ox * p.stride + tap % 3 - 1;
    if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) return 0.f;
    return p.A[(((long long)b * p.H + iy) * p.W + ix) * p.Cin + c];
  }
  // GA_PATCH: k = (c, ky, kx) of a 14x14 patch, m = (b, py, px)
  const int px = m % p.Wo, py = (m / p.Wo) % p.Ho, b = m / (p.Wo * p.Ho);
  const int c = k / 196, r = k - c * 196, ky = r / 14, kx = r - ky * 14;
  return p.A[(((long long)b * 3 + c) * p.H + py * 14 + ky) * p.W + px * 14 + kx];
}

// 64 x 64 output tile, 16-deep K steps, 256 threads x (4 x 4) accumulators
__global__ void __launch_bounds__(256) sgemm_f32_kernel(const SgemmParams p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int idx = threadIdx.x + t * 256;  // 0..1023
      const int r = idx >> 4, kk = idx & 15;
      As[kk][r] = sgemm_load_a(p, m0 + r, k0 + kk);
      const int n = n0 + r, k = k0 + kk;
      Bs[kk][r] = (n < p.N && k < p.K) ? p.Wt[(long long)n * p.K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[p.convt_s ? n % p.convt_cout : n];
      if (p.act == 1) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
      else if (p.act == 2) v = fmaxf(v, 0.f);
      long long o;
      if (p.convt_s) {
        const int s = p.convt_s, x = m % p.W, y = (m / p.W) % p.H, b = m / (p.W * p.H);
        const int kq = n / p.convt_cout, co = n - kq * p.convt_cout, ky = kq / s, kx = kq - ky * s;
        o = ((((long long)b * p.H * s + y * s + ky) * p.W * s) + x * s + kx) * p.convt_cout + co;
      } else if (p.tok_P) {
        const int b = m / p.tok_P, t = m - b * p.tok_P;
        v += p.pos[(long long)(1 + t) * p.N + n];
        o = ((long long)b * (p.tok_P + 1) + 1 + t) * p.ldc + n;
      } else {
        o = (long long)m * p.ldc + n;
      }
      if (p.gamma) v *= p.gamma[n];
      if (p.accumulate) v += p.C[o];
      if (p.add1) v += p.add1[o];
      if (p.add2) v += p.add2[o];
      p.C[o] = v;
      if (p.out_relu) p.out_relu[o] = fmaxf(v, 0.f);
    }
  }
}

static SgemmParams sg_blank() {
  SgemmParams p;
  memset(&p, 0, sizeof(p));
  return p;
}

static int sgemm(SgemmParams p, cudaStream_t stream) {
  DAV2_CHECK(p.A && p.Wt && p.C && p.M > 0 && p.N > 0 && p.K > 0, "fp32 sgemm: bad arguments");
  if (p.ldc == 0) p.ldc = p.N;
  dim3 grid((unsigned)((p.N + 63) / 64), (unsigned)((p.M + 63) / 64));
  ProfScope ps(PC_OTHER, 2.0 * p.M * (double)p.N * p.K, 0.0, stream);
  sgemm_f32_kernel<<<grid, 256, 0, stream>>>(p);
  DAV2_LAUNCH_OK();
  return 0;
}

static int sg_linear(const float* A, int M, int K, const h16* W, int N, const float* bias, float* C, int act, const float* gamma,
                     int accumulate, cudaStream_t stream) {
  SgemmParams p = sg_blank();
  p.A = A; p.M = M; p.K = K; p.lda = K; p.Wt = reinterpret_cast<const float*>(W); p.N = N; p.bias = bias; p.C = C;
  p.act = act; p.gamma = gamma; p.accumulate = accumulate;
  return sgemm(p, stream);
}

static int sg_conv3(const float* in, int B, int H, int W, int Cin, int stride, const h16* Wp, int Cout, const float* bias,
                    float* out, int act, const float* add1, const float* add2, float* out_relu, cudaStream_t stream) {
  SgemmParams p = sg_blank();
  p.gather = GA_CONV3;
  p.A = in; p.H = H; p.W = W; p.Cin = Cin; p.stride = stride;
  p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;  // k = 3, pad = 1
  p.M = B * p.Ho * p.Wo; p.K = 9 * Cin; p.N = Cout;
  p.Wt = reinterpret_cast<const float*>(Wp); p.bias = bias; p.C = out; p.act = act; p.add1 = add1; p.add2 = add2;
  p.out_relu = out_relu;
  return sgemm(p, stream);
}

// ----------------------------------------------------------------------------------------------
// attention: softmax(q k^T) v per (image, head, query row); q was pre-scaled by 1/8 in the packed qkv weights.
// One warp per query row; the row of scores lives in shared memory.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int N, int D) {
  extern __shared__ float sc[];  // [4][Npad]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= N) return;
  const int Npad = (N + 31) & ~31;
  float* s = sc + warp * Npad;
  const long long ld = 3ll * D;
  const float* q = qkv + ((long long)b * N + i) * ld + h * 64;
  const float* kbase = qkv + (long long)b * N * ld + D + h * 64;
  const float* vbase = qkv + (long long)b * N * ld + 2 * D + h * 64;
  float4 qv[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) qv[d] = reinterpret_cast<const float4*>(q)[d];
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    const float4* kr = reinterpret_cast<const float4*>(kbase + (long long)j * ld);
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float4 kv = kr[d];
      a = fmaf(qv[d].x, kv.x, a); a = fmaf(qv[d].y, kv.y, a); a = fmaf(qv[d].z, kv.z, a); a = fmaf(qv[d].w, kv.w, a);
    }
    s[j] = a;
    mx = fmaxf(mx, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < N; j += 32) {
    const float e = expf(s[j] - mx);
    s[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < N; ++j) {
    const float pj = s[j];
    const float* vr = vbase + (long long)j * ld;
    o0 = fmaf(pj, vr[lane], o0);
    o1 = fmaf(pj, vr[lane + 32], o1);
  }
  const float inv = 1.0f / sum;
  float* dst = out + ((long long)b * N + i) * D + h * 64;
  dst[lane] = o0 * inv;
  dst[lane + 32] = o1 * inv;
}

static int attention_f32(const float* qkv, float* out, int B, int N, int D, cudaStream_t stream) {
  const int Npad = (N + 31) & ~31;
  const size_t smem = (size_t)4 * Npad * sizeof(float);
  DAV2_CHECK(smem <= 200 * 1024, "fp32 attention: %d tokens need %zu bytes of shared memory", N, smem);
  static char tag;
  if (!device_setup_done(&tag)) {  // the largest size this path accepts: set once per device
    DAV2_CUDA_OK(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    device_setup_mark(&tag);
  }
  dim3 grid((unsigned)((N + 3) / 4), (unsigned)(D / 64), (unsigned)B);
  ProfScope ps(PC_OTHER, 4.0 * B * (D / 64) * (double)N * N * 64.0, 0.0, stream);
  attention_f32_kernel<<<grid, 128, smem, stream>>>(qkv, out, N, D);
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// LayerNorm (fp32 in, fp32 out; one warp per row; optional cls-drop compaction for the taps)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bvec, float* __restrict__ out, long long rows,
                                                            int D, int tokens, int drop_cls, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long orow = row;
  if (drop_cls) {
    const long long b = row / tokens, t = row - b * tokens;
    if (t == 0) return;
    orow = b * (tokens - 1) + t - 1;
  }
  const float* xr = x + row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += xr[d];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float v = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float t = xr[d] - mean;
    v = fmaf(t, t, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / (float)D + eps);
  float* orp = out + orow * D;
  for (int d = lane; d < D; d += 32) orp[d] = (xr[d] - mean) * rstd * w[d] + bvec[d];
}

static int layernorm_f32(const float* x, const float* w, const float* b, float* out, long long rows, int D, int tokens, int drop_cls,
                         cudaStream_t stream) {
  layernorm_f32_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(x, w, b, out, rows, D, tokens, drop_cls, 1e-6f);
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// bilinear resize, NHWC fp32, align_corners = True (F.interpolate in the DPT head)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bilinear_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Hi, int Wi,
                                                           int Ho, int Wo, int C) {
  const long long total = (long long)B * Ho * Wo * C;
  const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
    const float fy = oy * sy, fx = ox * sx;
    int y0 = (int)fy, x0 = (int)fx;
    y0 = min(y0, Hi - 1); x0 = min(x0, Wi - 1);
    const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
    const float wy = fy - (float)y0, wx = fx - (float)x0;
    const float* base = in + (long long)b * Hi * Wi * C + c;
    const float v00 = base[((long long)y0 * Wi + x0) * C], v01 = base[((long long)y0 * Wi + x1) * C];
    const float v10 = base[((long long)y1 * Wi + x0) * C], v11 = base[((long long)y1 * Wi + x1) * C];
    out[i] = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
  }
}

static int bilinear_f32(const float* in, float* out, int B, int Hi, int Wi, int Ho, int Wo, int C, cudaStream_t stream) {
  const long long total = (long long)B * Ho * Wo * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  bilinear_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, B, Hi, Wi, Ho, Wo, C);
  DAV2_LAUNCH_OK();
  return 0;
}

// output_conv2[2] (1x1, 32 -> 1) + sigmoid * max_depth on the ReLU'd 32-channel map
__global__ void __launch_bounds__(256) head_final_f32_kernel(const float* __restrict__ f32map, const float* __restrict__ w, float bias,
                                                             float max_depth, float* __restrict__ depth, long long npix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const float* r = f32map + i * 32;
    float a = bias;
#pragma unroll
    for (int c = 0; c < 32; ++c) a = fmaf(r[c], w[c], a);
    depth[i] = max_depth / (1.0f + expf(-a));
  }
}

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
#define RC(expr)                      \
  do {                                \
    if (int _rc = (expr)) return _rc; \
  } while (0)

int Model::forward_fp32(const float* x, int B, int H, int W, float* depth, cudaStream_t stream) {
  const int ph = H / 14, pw = W / 14, P = ph * pw, N = P + 1;
  const int M = B * N, MP = B * P;
  const float* posT = nullptr;
  if (ph == 37 && pw == 37) posT = pos;
  else {
    auto it = pos_tables.find(std::make_pair(ph, pw));
    DAV2_CHECK(it != pos_tables.end(), "forward: no position table for a %dx%d patch grid (call dav2_set_pos_embed)", ph, pw);
    posT = it->second;
  }
  float *X, *XN, *QKV, *ATT, *HID, *TAP[4];
  RC(buf("f32_x", (size_t)M * D * 4, (void**)&X));
  RC(buf("f32_xn", (size_t)M * D * 4, (void**)&XN));
  RC(buf("f32_qkv", (size_t)M * 3 * D * 4, (void**)&QKV));
  RC(buf("f32_attn", (size_t)M * D * 4, (void**)&ATT));
  RC(buf("f32_hid", (size_t)M * 4 * D * 4, (void**)&HID));
  for (int i = 0; i < 4; ++i) {
    char nm[24];
    snprintf(nm, sizeof(nm), "f32_tap%d", i);
    RC(buf(nm, (size_t)MP * D * 4, (void**)&TAP[i]));
  }
  // ---- patch embed + cls + pos ----
  {
    SgemmParams p = sg_blank();
    p.gather = GA_PATCH; p.A = x; p.H = H; p.W = W; p.Ho = ph; p.Wo = pw;
    p.M = MP; p.K = 588; p.N = D; p.Wt = reinterpret_cast<const float*>(patch_w); p.bias = patch_b;
    p.C = X; p.ldc = D; p.tok_P = P; p.pos = posT;
    RC(sgemm(p, stream));
  }
  RC(launch_cls_row(X, cls, posT, B, N, D, stream));
  // ---- transformer blocks ----
  int next_tap = 0;
  for (int l = 0; l < L; ++l) {
    const BlockW& w = blk[l];
    RC(layernorm_f32(X, w.n1w, w.n1b, XN, M, D, N, 0, stream));
    RC(sg_linear(XN, M, D, w.qkv_w, 3 * D, w.qkv_b, QKV, 0, nullptr, 0, stream));
    RC(attention_f32(QKV, ATT, B, N, D, stream));
    RC(sg_linear(ATT, M, D, w.proj_w, D, w.proj_b, X, 0, w.ls1, 1, stream));
    RC(layernorm_f32(X, w.n2w, w.n2b, XN, M, D, N, 0, stream));
    RC(sg_linear(XN, M, D, w.fc1_w, 4 * D, w.fc1_b, HID, 1, nullptr, 0, stream));
    RC(sg_linear(HID, M, 4 * D, w.fc2_w, D, w.fc2_b, X, 0, w.ls2, 1, stream));
    if (next_tap < 4 && l == cfg.tap_layers[next_tap]) {
      RC(layernorm_f32(X, norm_w, norm_b, TAP[next_tap], M, D, N, 1, stream));
      ++next_tap;
    }
  }
  DAV2_CHECK(next_tap == 4, "forward: tap layers must be increasing block indices < depth");
  // ---- DPT head ----
  const int* oc = cfg.out_channels;
  const int hh[4] = {4 * ph, 2 * ph, ph, (ph + 1) / 2};
  const int ww[4] = {4 * pw, 2 * pw, pw, (pw + 1) / 2};
  float* lvl[4];
  for (int i = 0; i < 4; ++i) {
    char nm[24];
    float* pr;
    snprintf(nm, sizeof(nm), "f32_proj%d", i);
    RC(buf(nm, (size_t)MP * oc[i] * 4, (void**)&pr));
    RC(sg_linear(TAP[i], MP, D, proj_w[i], oc[i], proj_b[i], pr, 0, nullptr, 0, stream));
    if (i == 2) {
      lvl[i] = pr;
      continue;
    }
    snprintf(nm, sizeof(nm), "f32_lvl%d", i);
    RC(buf(nm, (size_t)B * hh[i] * ww[i] * oc[i] * 4, (void**)&lvl[i]));
    if (i < 2) {
      const int s = i == 0 ? 4 : 2;
      SgemmParams p = sg_blank();
      p.A = pr; p.M = MP; p.K = oc[i]; p.lda = oc[i]; p.Wt = reinterpret_cast<const float*>(rs_w[i]); p.N = s * s * oc[i];
      p.bias = rs_b[i]; p.C = lvl[i]; p.convt_s = s; p.convt_cout = oc[i]; p.H = ph; p.W = pw;
      RC(sgemm(p, stream));
    } else {
      RC(sg_conv3(pr, B, ph, pw, oc[3], 2, rs_w[3], oc[3], rs_b[3], lvl[3], 0, nullptr, nullptr, nullptr, stream));
    }
  }
  float *rn[4], *rnr[4];
  for (int i = 0; i < 4; ++i) {
    char nm[24];
    snprintf(nm, sizeof(nm), "f32_rn%d", i);
    RC(buf(nm, (size_t)B * hh[i] * ww[i] * F * 4, (void**)&rn[i]));
    snprintf(nm, sizeof(nm), "f32_rn%d_relu", i);
    RC(buf(nm, (size_t)B * hh[i] * ww[i] * F * 4, (void**)&rnr[i]));
    RC(sg_conv3(lvl[i], B, hh[i], ww[i], oc[i], 1, rn_w[i], F, nullptr, rn[i], 0, nullptr, nullptr, rnr[i], stream));
  }
  // fusion blocks (refinenet4 -> refinenet1); out_conv runs before the resize exactly as in engine.cu (the two commute)
  const size_t big = (size_t)B * hh[0] * ww[0] * F * 4;
  float *T, *S, *SR, *Y, *OCb;
  RC(buf("f32_t", big, (void**)&T));
  RC(buf("f32_s", big, (void**)&S));
  RC(buf("f32_sr", big, (void**)&SR));
  RC(buf("f32_y", big, (void**)&Y));
  RC(buf("f32_oc", big, (void**)&OCb));
  float* up_prev = nullptr;
  for (int i = 3; i >= 0; --i) {
    const Fusion& f = ref[i];
    const int h = hh[i], w = ww[i];
    const float *in = rn[i], *in_relu = rnr[i];
    if (up_prev) {
      RC(sg_conv3(in_relu, B, h, w, F, 1, f.rcu_w[0][0], F, f.rcu_b[0][0], T, 2, nullptr, nullptr, nullptr, stream));
      RC(sg_conv3(T, B, h, w, F, 1, f.rcu_w[0][1], F, f.rcu_b[0][1], S, 0, in, up_prev, SR, stream));
      in = S;
      in_relu = SR;
    }
    RC(sg_conv3(in_relu, B, h, w, F, 1, f.rcu_w[1][0], F, f.rcu_b[1][0], T, 2, nullptr, nullptr, nullptr, stream));
    RC(sg_conv3(T, B, h, w, F, 1, f.rcu_w[1][1], F, f.rcu_b[1][1], Y, 0, in, nullptr, nullptr, stream));
    RC(sg_linear(Y, B * h * w, F, f.out_w, F, f.out_b, OCb, 0, nullptr, 0, stream));
    const int ho = i > 0 ? hh[i - 1] : 2 * hh[0], wo = i > 0 ? ww[i - 1] : 2 * ww[0];
    char nm[24];
    snprintf(nm, sizeof(nm), "f32_path%d", i + 1);
    float* up;
    RC(buf(nm, (size_t)B * ho * wo * F * 4, (void**)&up));
    RC(bilinear_f32(OCb, up, B, h, w, ho, wo, F, stream));
    up_prev = up;
  }
  const int h8 = 2 * hh[0], w8 = 2 * ww[0];
  float *O1, *O1U, *O2;
  RC(buf("f32_out1", (size_t)B * h8 * w8 * (F / 2) * 4, (void**)&O1));
  RC(buf("f32_out1_up", (size_t)B * H * W * (F / 2) * 4, (void**)&O1U));
  RC(buf("f32_out2", (size_t)B * H * W * 32 * 4, (void**)&O2));
  RC(sg_conv3(up_prev, B, h8, w8, F, 1, oc1_w, F / 2, oc1_b, O1, 0, nullptr, nullptr, nullptr, stream));
  RC(bilinear_f32(O1, O1U, B, h8, w8, H, W, F / 2, stream));
  RC(sg_conv3(O1U, B, H, W, F / 2, 1, oc2_w, 32, oc2_b, O2, 2, nullptr, nullptr, nullptr, stream));
  {
    const long long npix = (long long)B * H * W;
    long long blocks = (npix + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    head_final_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(O2, oc3_w, oc3_b, cfg.max_depth, depth, npix);
    DAV2_LAUNCH_OK();
  }
  return 0;
}

}  // namespace dav2
