// PoseEstimationNet forward (reference pose_estimation_model.py:35-105): ResNet-18 with an 8-channel stem
// on stacked frame pairs [rgb1, d1, rgb2, d2] (data_processing/pose_estimation.py:229-243), fc -> 256,
// MLP head 256 -> 128 -> 64 -> 7 = [t(3) | q xyzw(4)].  Eval semantics: BatchNorm uses running statistics
// (folded into the conv weights / biases by the host layer), Dropout is the identity.
//
// All convolutions run on the tcgen05 GEMM / implicit-GEMM kernels (NHWC 16-bit): the 7x7/2 stem and the
// stride-2 3x3 / 1x1 convs through small im2col / subsample gathers, the stride-1 3x3 convs without any
// im2col.  Helper kernels here: NCHW fp32 -> NHWC 16-bit pack, generic im2col, max-pool, add+ReLU,
// global average pool + fp32 MLP head.
#include "posenet.cuh"

#include <string.h>

namespace dav2 {

// ---------------------------------------------------------------------------------------------- kernels
template <int FMT>
__global__ void __launch_bounds__(256) pack_nchw8_kernel(const float* __restrict__ x, h16* __restrict__ out, int B, int HW) {
  // [B,8,H,W] fp32 -> [B,H,W,8] 16-bit: one pixel (16 B) per thread
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, px = i - b * HW;
    const float* src = x + b * 8 * HW + px;
    uint4 o;
    o.x = pack2<FMT>(__ldg(src), __ldg(src + HW));
    o.y = pack2<FMT>(__ldg(src + 2LL * HW), __ldg(src + 3LL * HW));
    o.z = pack2<FMT>(__ldg(src + 4LL * HW), __ldg(src + 5LL * HW));
    o.w = pack2<FMT>(__ldg(src + 6LL * HW), __ldg(src + 7LL * HW));
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// generic im2col on NHWC 16-bit, C multiple of 8: [B,H,W,C] -> [B*Ho*Wo, KP], k = (ky*ks+kx)*C + c, zero padded to KP
__global__ void __launch_bounds__(256) im2col_kernel(const h16* __restrict__ in, h16* __restrict__ A, int B, int H, int W, int C,
                                                     int ks, int stride, int pad, int Ho, int Wo, int KP) {
  const int c8 = C >> 3;
  const int kv = KP >> 3;  // 16-byte vectors per output row
  const long long total = (long long)B * Ho * Wo * kv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % kv);
    long long t = i / kv;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    uint4 val = make_uint4(0, 0, 0, 0);
    const int tap = v / c8, cv = v - tap * c8;
    if (tap < ks * ks) {
      const int ky = tap / ks, kx = tap - ky * ks;
      const int yi = yo * stride - pad + ky, xi = xo * stride - pad + kx;
      if (yi >= 0 && yi < H && xi >= 0 && xi < W)
        val = __ldg(reinterpret_cast<const uint4*>(in + (((long long)b * H + yi) * W + xi) * C) + cv);
    }
    reinterpret_cast<uint4*>(A)[i] = val;
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const h16* __restrict__ in, h16* __restrict__ out, int B, int H, int W,
                                                         int C, int Ho, int Wo) {
  const int c8 = C >> 3;
  const long long total = (long long)B * Ho * Wo * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % c8);
    long long t = i / c8;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int yi = yo * 2 - 1 + ky, xi = xo * 2 - 1 + kx;
        if (yi < 0 || yi >= H || xi < 0 || xi >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)b * H + yi) * W + xi) * C) + cv);
        const uint32_t* w = &v.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack2<FMT>(w[k]);
          m[2 * k] = fmaxf(m[2 * k], f.x);
          m[2 * k + 1] = fmaxf(m[2 * k + 1], f.y);
        }
      }
    uint4 o;
    o.x = pack2<FMT>(m[0], m[1]); o.y = pack2<FMT>(m[2], m[3]); o.z = pack2<FMT>(m[4], m[5]); o.w = pack2<FMT>(m[6], m[7]);
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// out = relu(a + b), 16-bit, 8 elements per thread
template <int FMT>
__global__ void __launch_bounds__(256) add_relu_kernel(const h16* __restrict__ a, const h16* __restrict__ b, h16* __restrict__ out,
                                                       long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 va = __ldg(reinterpret_cast<const uint4*>(a) + i), vb = __ldg(reinterpret_cast<const uint4*>(b) + i);
    const uint32_t* wa = &va.x;
    const uint32_t* wb = &vb.x;
    uint4 o;
    uint32_t* wo = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack2<FMT>(wa[k]), fb = unpack2<FMT>(wb[k]);
      wo[k] = pack2<FMT>(fmaxf(fa.x + fb.x, 0.f), fmaxf(fa.y + fb.y, 0.f));
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// global average pool over HW of [B,HW,512] (16-bit) + fc 512->256 + ReLU + 256->128 + ReLU + 128->64 + ReLU + 64->7,
// all fp32 on CUDA cores (172 k MAC per sample): one block of 256 threads per sample.
template <int FMT>
__global__ void __launch_bounds__(256) pool_mlp_kernel(const h16* __restrict__ feat, int HW, const float* __restrict__ fc_w,
                                                       const float* __restrict__ fc_b, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, const float* __restrict__ w3,
                                                       const float* __restrict__ b3, float* __restrict__ out) {
  __shared__ float s0[512], s1[256], s2[128], s3[64];
  const int b = blockIdx.x, t = threadIdx.x;
  const h16* f = feat + (long long)b * HW * 512;
  for (int c2 = t; c2 < 256; c2 += 256) {  // two channels per thread
    float a0 = 0.f, a1 = 0.f;
    for (int p = 0; p < HW; ++p) {
      const float2 v = unpack2<FMT>(__ldg(reinterpret_cast<const uint32_t*>(f + (long long)p * 512) + c2));
      a0 += v.x; a1 += v.y;
    }
    s0[2 * c2] = a0 / (float)HW;
    s0[2 * c2 + 1] = a1 / (float)HW;
  }
  __syncthreads();
  {  // fc 512 -> 256, then the head's leading ReLU
    float acc = fc_b[t];
    const float* w = fc_w + (long long)t * 512;
    for (int k = 0; k < 512; ++k) acc = fmaf(w[k], s0[k], acc);
    s1[t] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  if (t < 128) {
    float acc = b1[t];
    const float* w = w1 + t * 256;
    for (int k = 0; k < 256; ++k) acc = fmaf(w[k], s1[k], acc);
    s2[t] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  if (t < 64) {
    float acc = b2[t];
    const float* w = w2 + t * 128;
    for (int k = 0; k < 128; ++k) acc = fmaf(w[k], s2[k], acc);
    s3[t] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  if (t < 7) {
    float acc = b3[t];
    const float* w = w3 + t * 64;
    for (int k = 0; k < 64; ++k) acc = fmaf(w[k], s3[k], acc);
    out[b * 7 + t] = acc;
  }
}

static inline int grid_for(long long total, int per_block = 256, int waves = 16) {
  long long blocks = (total + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * waves;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

// ---------------------------------------------------------------------------------------------- model
PoseModel::PoseModel(int precision) : fmt(precision == 1 ? FMT_BF16 : FMT_F16) { cudaGetDevice(&device); }

PoseModel::~PoseModel() {
  for (void* p : owned) cudaFree(p);
  for (auto& kv : ws) cudaFree(kv.second.p);
}

static int up(PoseModel* m, const void* host, size_t bytes, void** out) {
  void* p = nullptr;
  DAV2_CUDA_OK(cudaMalloc(&p, bytes));
  DAV2_CUDA_OK(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
  m->owned.push_back(p);
  *out = p;
  return 0;
}

// Keys (BatchNorm already folded by the host layer):
//   "<conv>.weight" [Cout,Cin,k,k] + "<conv>.bias" [Cout] for conv in {conv1, layerL.B.conv1, layerL.B.conv2,
//   layerL.0.downsample}; "fc.weight/bias", "head.0|1|2.weight/bias" (fp32).
int PoseModel::set_weight(const char* key, const float* data, const int64_t* shape, int ndim) {
  DAV2_CHECK(key && data && shape, "pose set_weight: null argument");
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  const std::string K(key);
  if (K.rfind("fc.", 0) == 0 || K.rfind("head.", 0) == 0 || K.size() > 5 && K.compare(K.size() - 5, 5, ".bias") == 0) {
    void* p;
    if (int rc = up(this, data, (size_t)n * 4, &p)) return rc;
    f32[K] = (float*)p;
    return 0;
  }
  DAV2_CHECK(ndim == 4 && K.size() > 7 && K.compare(K.size() - 7, 7, ".weight") == 0, "pose set_weight: unexpected key %s", key);
  const int Cout = (int)shape[0], Cin = (int)shape[1], ks = (int)shape[2];
  ConvW w;
  w.Cout = Cout; w.Cin = Cin; w.ks = ks;
  std::vector<h16> v;
  if (ks == 3 && K.find("downsample") == std::string::npos) {
    // tap-major [Cout][tap*Cpad + c]; the stride-2 variant is fed by im2col with row pitch 9*Cin (Cin % 64 == 0 here)
    const int Cpad = (Cin + 63) / 64 * 64;
    w.K = 9 * Cpad;
    v.assign((size_t)Cout * w.K, f2h_host(0.f, fmt));
    for (int o = 0; o < Cout; ++o)
      for (int c = 0; c < Cin; ++c)
        for (int t = 0; t < 9; ++t) v[((size_t)o * 9 + t) * Cpad + c] = f2h_host(data[((size_t)o * Cin + c) * 9 + t], fmt);
  } else {
    // im2col order k = tap*Cin + c, zero padded to a multiple of 64 (7x7 stem: 392 -> 448; 1x1: Cin)
    const int kk = ks * ks;
    w.K = (kk * Cin + 63) / 64 * 64;
    v.assign((size_t)Cout * w.K, f2h_host(0.f, fmt));
    for (int o = 0; o < Cout; ++o)
      for (int c = 0; c < Cin; ++c)
        for (int t = 0; t < kk; ++t) v[(size_t)o * w.K + (size_t)t * Cin + c] = f2h_host(data[((size_t)o * Cin + c) * kk + t], fmt);
  }
  void* p;
  if (int rc = up(this, v.data(), v.size() * 2, &p)) return rc;
  w.w = (h16*)p;
  conv[K.substr(0, K.size() - 7)] = w;
  return 0;
}

int PoseModel::buf(const char* name, size_t bytes, void** out) {
  DevBuf& b = ws[name];
  if (b.cap < bytes) {
    if (b.p) DAV2_CUDA_OK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    const size_t cap = (bytes + 255) & ~(size_t)255;
    DAV2_CUDA_OK(cudaMalloc(&b.p, cap));
    b.cap = cap;
  }
  b.bytes = bytes;
  *out = b.p;
  return 0;
}

#define RC(expr)                      \
  do {                                \
    if (int _rc = (expr)) return _rc; \
  } while (0)

int PoseModel::get_conv(const std::string& name, ConvW* w, const float** bias) {
  auto it = conv.find(name);
  DAV2_CHECK(it != conv.end(), "pose forward: conv '%s' was never loaded", name.c_str());
  auto ib = f32.find(name + ".bias");
  DAV2_CHECK(ib != f32.end(), "pose forward: bias of '%s' was never loaded", name.c_str());
  *w = it->second;
  *bias = ib->second;
  return 0;
}

// conv via im2col + GEMM (any k / stride): in NHWC [B,H,W,Cin] -> out [B,Ho,Wo,Cout], act 0/2
int PoseModel::conv_im2col(const std::string& name, const h16* in, int B, int H, int W, int stride, int pad, int act, h16* out,
                           int* Ho_, int* Wo_, cudaStream_t stream) {
  ConvW w;
  const float* bias;
  RC(get_conv(name, &w, &bias));
  const int Ho = (H + 2 * pad - w.ks) / stride + 1, Wo = (W + 2 * pad - w.ks) / stride + 1;
  const int KP = w.ks == 3 ? 9 * w.Cin : w.K;  // the 3x3 tap-major packing has no per-tap padding when Cin % 64 == 0
  DAV2_CHECK(w.ks != 3 || w.Cin % 64 == 0, "pose conv %s: Cin must be a multiple of 64", name.c_str());
  h16* col;
  RC(buf("im2col", (size_t)B * Ho * Wo * KP * 2, (void**)&col));
  const long long total = (long long)B * Ho * Wo * (KP / 8);
  {
    ProfScope ps(PC_IM2COL, 0.0, (double)total * 32.0, stream);
    im2col_kernel<<<grid_for(total), 256, 0, stream>>>(in, col, B, H, W, w.Cin, w.ks, stride, pad, Ho, Wo, KP);
    DAV2_LAUNCH_OK();
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.fmt = fmt; p.out = out; p.ldo = w.Cout; p.bias = bias; p.act = act;
  RC(gemm_linear(GM_LINEAR_BF16, col, B * Ho * Wo, KP, KP, w.w, w.Cout, p, stream));
  *Ho_ = Ho;
  *Wo_ = Wo;
  return 0;
}

int PoseModel::add_relu(const h16* a, const h16* b, h16* out, long long n, cudaStream_t stream) {
  DAV2_CHECK(n % 8 == 0, "add_relu: size must be a multiple of 8");
  ProfScope ps(PC_OTHER, 0.0, (double)n * 6.0, stream);
  if (fmt == FMT_BF16) add_relu_kernel<FMT_BF16><<<grid_for(n / 8), 256, 0, stream>>>(a, b, out, n / 8);
  else add_relu_kernel<FMT_F16><<<grid_for(n / 8), 256, 0, stream>>>(a, b, out, n / 8);
  DAV2_LAUNCH_OK();
  return 0;
}

int PoseModel::forward(const float* x, int B, int H, int W, float* out7, cudaStream_t stream) {
  DAV2_CHECK(x && out7 && B > 0 && H >= 32 && W >= 32, "pose forward: bad arguments");
  {
    int cur = -1;
    DAV2_CUDA_OK(cudaGetDevice(&cur));
    DAV2_CHECK(cur == device, "pose forward: this handle lives on device %d but the call runs on device %d", device, cur);
  }
  for (const char* k : {"fc.weight", "fc.bias", "head.0.weight", "head.0.bias", "head.1.weight", "head.1.bias", "head.2.weight",
                        "head.2.bias"})
    DAV2_CHECK(f32.count(k), "pose forward: '%s' was never loaded", k);
  const size_t S2 = 2;
  // ---- stem: pack pair -> NHWC8, 7x7/2 conv (+folded BN, ReLU), 3x3/2 max-pool -----------------------
  h16 *x8, *c1, *cur, *nxt, *tmp, *idn;
  RC(buf("x8", (size_t)B * H * W * 8 * S2, (void**)&x8));
  {
    ProfScope ps(PC_IM2COL, 0.0, (double)B * H * W * 48.0, stream);
    if (fmt == FMT_BF16) pack_nchw8_kernel<FMT_BF16><<<grid_for((long long)B * H * W), 256, 0, stream>>>(x, x8, B, H * W);
    else pack_nchw8_kernel<FMT_F16><<<grid_for((long long)B * H * W), 256, 0, stream>>>(x, x8, B, H * W);
    DAV2_LAUNCH_OK();
  }
  const int H1 = (H + 6 - 7) / 2 + 1, W1 = (W + 6 - 7) / 2 + 1;
  RC(buf("stem", (size_t)B * H1 * W1 * 64 * S2, (void**)&c1));
  int ho, wo;
  RC(conv_im2col("conv1", x8, B, H, W, 2, 3, 2, c1, &ho, &wo, stream));
  int h = (H1 + 2 - 3) / 2 + 1, w = (W1 + 2 - 3) / 2 + 1;
  const size_t act_bytes = (size_t)B * h * w * 64 * S2;  // layer1 is the largest activation after the pool
  RC(buf("act_a", act_bytes, (void**)&cur));
  RC(buf("act_b", act_bytes, (void**)&nxt));
  RC(buf("act_c", act_bytes, (void**)&tmp));
  RC(buf("act_d", act_bytes, (void**)&idn));
  {
    const long long total = (long long)B * h * w * 8;
    ProfScope ps(PC_RESAMPLE, 0.0, (double)B * H1 * W1 * 128.0, stream);
    if (fmt == FMT_BF16) maxpool3s2_kernel<FMT_BF16><<<grid_for(total), 256, 0, stream>>>(c1, cur, B, H1, W1, 64, h, w);
    else maxpool3s2_kernel<FMT_F16><<<grid_for(total), 256, 0, stream>>>(c1, cur, B, H1, W1, 64, h, w);
    DAV2_LAUNCH_OK();
  }
  // ---- 4 stages x 2 BasicBlocks -----------------------------------------------------------------------
  int C = 64;
  for (int L = 1; L <= 4; ++L) {
    const int Cout = 64 << (L - 1);
    for (int blk = 0; blk < 2; ++blk) {
      char nm[64];
      const bool down = (L > 1 && blk == 0);
      ConvW w1, w2;
      const float *b1, *b2;
      snprintf(nm, sizeof(nm), "layer%d.%d.conv1", L, blk);
      const std::string n1(nm);
      snprintf(nm, sizeof(nm), "layer%d.%d.conv2", L, blk);
      const std::string n2(nm);
      RC(get_conv(n1, &w1, &b1));
      RC(get_conv(n2, &w2, &b2));
      int h2 = h, w2o = w;
      const h16* identity = cur;
      if (down) {
        // y = relu(bn1(conv1 3x3/2)); identity = bn(downsample 1x1/2)
        RC(conv_im2col(n1, cur, B, h, w, 2, 1, 2, tmp, &h2, &w2o, stream));
        snprintf(nm, sizeof(nm), "layer%d.0.downsample", L);
        int hd, wd;
        RC(conv_im2col(nm, cur, B, h, w, 2, 0, 0, idn, &hd, &wd, stream));
        identity = idn;
      } else {
        GemmParams p;
        memset(&p, 0, sizeof(p));
        p.fmt = fmt; p.out = tmp; p.bias = b1; p.act = 2;
        RC(conv3x3(GM_CONV_BF16, cur, B, h, w, C, w1.w, Cout, p, stream));
      }
      // out = relu(bn2(conv2 3x3/1)(y) + identity): the conv epilogue adds the identity and emits relu(out) as its
      // second output (the pre-ReLU sum lands in a scratch buffer)
      GemmParams p;
      memset(&p, 0, sizeof(p));
      p.fmt = fmt; p.out = idn == identity ? cur : idn; p.out_relu = nxt; p.bias = b2; p.add1 = identity;
      // (scratch for the pre-ReLU sum: whichever of cur / idn is NOT the identity and no longer needed)
      RC(conv3x3(GM_CONV_BF16, tmp, B, h2, w2o, Cout, w2.w, Cout, p, stream));
      std::swap(cur, nxt);
      h = h2; w = w2o; C = Cout;
    }
  }
  // ---- global average pool + fc + MLP head (fp32) -------------------------------------------------------
  {
    ProfScope ps(PC_OTHER, 2.0 * B * 172032.0, (double)B * h * w * 1024.0, stream);
    if (fmt == FMT_BF16)
      pool_mlp_kernel<FMT_BF16><<<B, 256, 0, stream>>>(cur, h * w, f32["fc.weight"], f32["fc.bias"], f32["head.0.weight"], f32["head.0.bias"],
                                                       f32["head.1.weight"], f32["head.1.bias"], f32["head.2.weight"], f32["head.2.bias"], out7);
    else
      pool_mlp_kernel<FMT_F16><<<B, 256, 0, stream>>>(cur, h * w, f32["fc.weight"], f32["fc.bias"], f32["head.0.weight"], f32["head.0.bias"],
                                                      f32["head.1.weight"], f32["head.1.bias"], f32["head.2.weight"], f32["head.2.bias"], out7);
    DAV2_LAUNCH_OK();
  }
  return 0;
}

}  // namespace dav2
