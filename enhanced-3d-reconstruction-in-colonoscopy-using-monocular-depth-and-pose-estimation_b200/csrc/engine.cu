// DepthAnythingV2 forward engine: packed weights + workspace + kernel orchestration.
// Mirrors the call structure of the external dpt.DepthAnythingV2.forward the reference drives
// (run.py:234, lightning_model.py:301): DINOv2 encoder (4 taps) -> DPT head -> sigmoid * max_depth.
#include "engine.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace dav2 {

static const int KP_PATCH = 640;  // 3*14*14 = 588 padded to a multiple of the 64-wide K block

// ----------------------------------------------------------------------------------------------
// small host utilities
// ----------------------------------------------------------------------------------------------
static int64_t numel(const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

static int dev_upload(const void* host, size_t bytes, void** out) {
  void* p = nullptr;
  DAV2_CUDA_OK(cudaMalloc(&p, bytes));
  DAV2_CUDA_OK(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}

static int upload_f32(const float* data, int64_t n, float** out) {
  return dev_upload(data, (size_t)n * 4, reinterpret_cast<void**>(out));
}

// Packed matrix weights are built in float in the kernel layout and then converted to the model's operand format:
// 16-bit (fp16 / bf16) for the tensor-core engine, or left in fp32 for the fp32 validation engine (FMT_F32, fp32_path.cu).
static std::vector<float> to_f(const float* d, int64_t n, float scale_first = 1.f, int64_t first = 0) {
  std::vector<float> v((size_t)n);
  for (int64_t i = 0; i < n; ++i) v[(size_t)i] = i < first ? d[i] * scale_first : d[i];
  return v;
}

// Conv2d weight [Cout, Cin, 3, 3] -> [Cout][tap*Cpad + c] (tap = ky*3+kx), zero padded to Cpad
static std::vector<float> pack_conv3x3(const float* w, int Cout, int Cin, int Cpad) {
  std::vector<float> v((size_t)Cout * 9 * Cpad, 0.f);
  for (int n = 0; n < Cout; ++n)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < 9; ++t) v[((size_t)n * 9 + t) * Cpad + c] = w[((size_t)n * Cin + c) * 9 + t];
  return v;
}

// ConvTranspose2d weight [Cin, Cout, s, s] -> [(ky*s+kx)*Cout + co][ci]
static std::vector<float> pack_convT(const float* w, int Cin, int Cout, int s) {
  std::vector<float> v((size_t)s * s * Cout * Cin);
  for (int ci = 0; ci < Cin; ++ci)
    for (int co = 0; co < Cout; ++co)
      for (int k = 0; k < s * s; ++k) v[((size_t)k * Cout + co) * Cin + ci] = w[((size_t)ci * Cout + co) * s * s + k];
  return v;
}

// upload in the operand format; the result is typed h16* for the 16-bit engine and really float* when fmt == FMT_F32
static int upload_w(int fmt, const std::vector<float>& v, h16** out) {
  if (fmt == FMT_F32) return dev_upload(v.data(), v.size() * 4, reinterpret_cast<void**>(out));
  std::vector<h16> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) h[i] = f2h_host(v[i], fmt);
  return dev_upload(h.data(), h.size() * 2, reinterpret_cast<void**>(out));
}

static bool shape_is(const int64_t* shape, int ndim, std::initializer_list<int64_t> want) {
  if ((int)want.size() != ndim) return false;
  int i = 0;
  for (int64_t w : want)
    if (shape[i++] != w) return false;
  return true;
}

#define WANT_SHAPE(...)                                                                 \
  DAV2_CHECK(shape_is(shape, ndim, {__VA_ARGS__}), "set_weight(%s): unexpected shape", key)

// ----------------------------------------------------------------------------------------------
// model lifecycle
// ----------------------------------------------------------------------------------------------
Model::Model(const dav2_config& c) : cfg(c) {
  D = c.embed_dim;
  L = c.depth;
  heads = c.num_heads;
  F = c.features;
  fmt = c.precision == 2 ? FMT_F32 : (c.precision == 1 ? FMT_BF16 : FMT_F16);
  cudaGetDevice(&device);
  blk.resize(L);
  memset(blk.data(), 0, sizeof(BlockW) * L);
  memset(proj_w, 0, sizeof(proj_w)); memset(proj_b, 0, sizeof(proj_b));
  memset(rs_w, 0, sizeof(rs_w)); memset(rs_b, 0, sizeof(rs_b));
  memset(rn_w, 0, sizeof(rn_w)); memset(ref, 0, sizeof(ref));
  char k[160];
  auto need = [&](const char* s) { required.insert(s); };
  need("pretrained.cls_token"); need("pretrained.pos_embed");
  need("pretrained.patch_embed.proj.weight"); need("pretrained.patch_embed.proj.bias");
  need("pretrained.norm.weight"); need("pretrained.norm.bias");
  static const char* blk_keys[] = {"norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
                                   "attn.proj.bias", "ls1.gamma", "norm2.weight", "norm2.bias", "mlp.fc1.weight",
                                   "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "ls2.gamma"};
  for (int i = 0; i < L; ++i)
    for (const char* s : blk_keys) {
      snprintf(k, sizeof(k), "pretrained.blocks.%d.%s", i, s);
      need(k);
    }
  for (int i = 0; i < 4; ++i) {
    snprintf(k, sizeof(k), "depth_head.projects.%d.weight", i); need(k);
    snprintf(k, sizeof(k), "depth_head.projects.%d.bias", i); need(k);
    if (i != 2) {
      snprintf(k, sizeof(k), "depth_head.resize_layers.%d.weight", i); need(k);
      snprintf(k, sizeof(k), "depth_head.resize_layers.%d.bias", i); need(k);
    }
    snprintf(k, sizeof(k), "depth_head.scratch.layer%d_rn.weight", i + 1); need(k);
    snprintf(k, sizeof(k), "depth_head.scratch.refinenet%d.out_conv.weight", i + 1); need(k);
    snprintf(k, sizeof(k), "depth_head.scratch.refinenet%d.out_conv.bias", i + 1); need(k);
    for (int u = 1; u <= 2; ++u) {
      if (i == 3 && u == 1) continue;  // refinenet4.resConfUnit1 is never executed
      for (int c = 1; c <= 2; ++c) {
        snprintf(k, sizeof(k), "depth_head.scratch.refinenet%d.resConfUnit%d.conv%d.weight", i + 1, u, c); need(k);
        snprintf(k, sizeof(k), "depth_head.scratch.refinenet%d.resConfUnit%d.conv%d.bias", i + 1, u, c); need(k);
      }
    }
  }
  need("depth_head.scratch.output_conv1.weight"); need("depth_head.scratch.output_conv1.bias");
  need("depth_head.scratch.output_conv2.0.weight"); need("depth_head.scratch.output_conv2.0.bias");
  need("depth_head.scratch.output_conv2.2.weight"); need("depth_head.scratch.output_conv2.2.bias");
}

Model::~Model() {
  for (void* p : owned) cudaFree(p);
  for (auto& kv : ws) cudaFree(kv.second.p);
}

int Model::set_weight(const char* key, const float* data, const int64_t* shape, int ndim) {
  DAV2_CHECK(key && data && shape, "set_weight: null argument");
  const int64_t n = numel(shape, ndim);
  const std::string K(key);
  int idx = 0, u = 0, c = 0;
  char tail[64];
  const int64_t Dl = D;
  // input channels of a packed 3x3 filter: padded to the 64-wide K block for the tensor-core kernels only
  auto cpad = [&](int c) { return fmt == FMT_F32 ? c : (c + 63) / 64 * 64; };

#define STORE_F32(dst)                                    \
  do {                                                    \
    float* _p = nullptr;                                  \
    if (int rc = upload_f32(data, n, &_p)) return rc;     \
    owned.push_back(_p);                                  \
    dst = _p;                                             \
  } while (0)
#define STORE_H16(dst, vec)                              \
  do {                                                    \
    h16* _p = nullptr;                                   \
    if (int rc = upload_w(fmt, vec, &_p)) return rc;    \
    owned.push_back(_p);                                  \
    dst = _p;                                             \
  } while (0)

  if (K == "pretrained.mask_token") {
    return 0;  // present in checkpoints, unused in eval
  } else if (K == "pretrained.cls_token") {
    DAV2_CHECK(n == Dl, "set_weight(%s): unexpected size", key);
    STORE_F32(cls);
  } else if (K == "pretrained.pos_embed") {
    WANT_SHAPE(1, 1370, Dl);
    STORE_F32(pos);
  } else if (K == "pretrained.patch_embed.proj.weight") {
    WANT_SHAPE(Dl, 3, 14, 14);
    const int kp = fmt == FMT_F32 ? 588 : KP_PATCH;  // the tensor-core GEMM wants K padded to the 64-wide block
    std::vector<float> v((size_t)D * kp, 0.f);
    for (int d = 0; d < D; ++d)
      for (int k = 0; k < 588; ++k) v[(size_t)d * kp + k] = data[(size_t)d * 588 + k];
    STORE_H16(patch_w, v);
  } else if (K == "pretrained.patch_embed.proj.bias") {
    WANT_SHAPE(Dl);
    STORE_F32(patch_b);
  } else if (K == "pretrained.norm.weight") {
    WANT_SHAPE(Dl);
    STORE_F32(norm_w);
  } else if (K == "pretrained.norm.bias") {
    WANT_SHAPE(Dl);
    STORE_F32(norm_b);
  } else if (sscanf(key, "pretrained.blocks.%d.%63s", &idx, tail) == 2) {
    DAV2_CHECK(idx >= 0 && idx < L, "set_weight(%s): block index out of range", key);
    BlockW& b = blk[idx];
    const std::string T(tail);
    if (T == "norm1.weight") { WANT_SHAPE(Dl); STORE_F32(b.n1w); }
    else if (T == "norm1.bias") { WANT_SHAPE(Dl); STORE_F32(b.n1b); }
    else if (T == "norm2.weight") { WANT_SHAPE(Dl); STORE_F32(b.n2w); }
    else if (T == "norm2.bias") { WANT_SHAPE(Dl); STORE_F32(b.n2b); }
    else if (T == "ls1.gamma") { WANT_SHAPE(Dl); STORE_F32(b.ls1); }
    else if (T == "ls2.gamma") { WANT_SHAPE(Dl); STORE_F32(b.ls2); }
    else if (T == "attn.qkv.weight") {
      WANT_SHAPE(3 * Dl, Dl);
      // q rows pre-scaled by d_head^-1/2 = 1/8 (exact in h16/fp32): upstream scales q before q@k^T
      STORE_H16(b.qkv_w, to_f(data, n, 0.125f, Dl * Dl));
    } else if (T == "attn.qkv.bias") {
      WANT_SHAPE(3 * Dl);
      std::vector<float> t(data, data + n);
      for (int i = 0; i < D; ++i) t[i] *= 0.125f;
      float* p = nullptr;
      if (int rc = upload_f32(t.data(), n, &p)) return rc;
      owned.push_back(p);
      b.qkv_b = p;
    }
    else if (T == "attn.proj.weight") { WANT_SHAPE(Dl, Dl); STORE_H16(b.proj_w, to_f(data, n)); }
    else if (T == "attn.proj.bias") { WANT_SHAPE(Dl); STORE_F32(b.proj_b); }
    else if (T == "mlp.fc1.weight") { WANT_SHAPE(4 * Dl, Dl); STORE_H16(b.fc1_w, to_f(data, n)); }
    else if (T == "mlp.fc1.bias") { WANT_SHAPE(4 * Dl); STORE_F32(b.fc1_b); }
    else if (T == "mlp.fc2.weight") { WANT_SHAPE(Dl, 4 * Dl); STORE_H16(b.fc2_w, to_f(data, n)); }
    else if (T == "mlp.fc2.bias") { WANT_SHAPE(Dl); STORE_F32(b.fc2_b); }
    else { set_last_error("set_weight: unknown key %s", key); return -4; }
  } else if (sscanf(key, "depth_head.projects.%d.%63s", &idx, tail) == 2) {
    DAV2_CHECK(idx >= 0 && idx < 4, "set_weight(%s): index", key);
    const int64_t oc = cfg.out_channels[idx];
    if (!strcmp(tail, "weight")) { WANT_SHAPE(oc, Dl, 1, 1); STORE_H16(proj_w[idx], to_f(data, n)); }
    else { WANT_SHAPE(oc); STORE_F32(proj_b[idx]); }
  } else if (sscanf(key, "depth_head.resize_layers.%d.%63s", &idx, tail) == 2) {
    DAV2_CHECK(idx == 0 || idx == 1 || idx == 3, "set_weight(%s): index", key);
    const int64_t oc = cfg.out_channels[idx];
    if (!strcmp(tail, "bias")) { WANT_SHAPE(oc); STORE_F32(rs_b[idx]); }
    else if (idx == 3) { WANT_SHAPE(oc, oc, 3, 3); STORE_H16(rs_w[3], pack_conv3x3(data, (int)oc, (int)oc, (int)oc)); }
    else {
      const int s = idx == 0 ? 4 : 2;
      WANT_SHAPE(oc, oc, s, s);
      STORE_H16(rs_w[idx], pack_convT(data, (int)oc, (int)oc, s));
    }
  } else if (sscanf(key, "depth_head.scratch.layer%d_rn.%63s", &idx, tail) == 2) {
    DAV2_CHECK(idx >= 1 && idx <= 4 && !strcmp(tail, "weight"), "set_weight(%s): bad key", key);
    const int64_t oc = cfg.out_channels[idx - 1];
    WANT_SHAPE(F, oc, 3, 3);
    STORE_H16(rn_w[idx - 1], pack_conv3x3(data, F, (int)oc, cpad((int)oc)));
  } else if (sscanf(key, "depth_head.scratch.refinenet%d.resConfUnit%d.conv%d.%63s", &idx, &u, &c, tail) == 4) {
    DAV2_CHECK(idx >= 1 && idx <= 4 && u >= 1 && u <= 2 && c >= 1 && c <= 2, "set_weight(%s): bad key", key);
    Fusion& f = ref[idx - 1];
    if (!strcmp(tail, "weight")) {
      WANT_SHAPE(F, F, 3, 3);
      STORE_H16(f.rcu_w[u - 1][c - 1], pack_conv3x3(data, F, F, cpad(F)));
    } else { WANT_SHAPE(F); STORE_F32(f.rcu_b[u - 1][c - 1]); }
  } else if (sscanf(key, "depth_head.scratch.refinenet%d.out_conv.%63s", &idx, tail) == 2) {
    DAV2_CHECK(idx >= 1 && idx <= 4, "set_weight(%s): bad key", key);
    if (!strcmp(tail, "weight")) { WANT_SHAPE(F, F, 1, 1); STORE_H16(ref[idx - 1].out_w, to_f(data, n)); }
    else { WANT_SHAPE(F); STORE_F32(ref[idx - 1].out_b); }
  } else if (K == "depth_head.scratch.output_conv1.weight") {
    WANT_SHAPE(F / 2, F, 3, 3);
    STORE_H16(oc1_w, pack_conv3x3(data, F / 2, F, cpad(F)));
  } else if (K == "depth_head.scratch.output_conv1.bias") {
    WANT_SHAPE(F / 2);
    STORE_F32(oc1_b);
  } else if (K == "depth_head.scratch.output_conv2.0.weight") {
    WANT_SHAPE(32, F / 2, 3, 3);
    STORE_H16(oc2_w, pack_conv3x3(data, 32, F / 2, cpad(F / 2)));
  } else if (K == "depth_head.scratch.output_conv2.0.bias") {
    WANT_SHAPE(32);
    STORE_F32(oc2_b);
  } else if (K == "depth_head.scratch.output_conv2.2.weight") {
    WANT_SHAPE(1, 32, 1, 1);
    STORE_F32(oc3_w);
  } else if (K == "depth_head.scratch.output_conv2.2.bias") {
    WANT_SHAPE(1);
    oc3_b = data[0];
  } else {
    set_last_error("set_weight: unknown key %s", key);
    return -4;
  }
  loaded.insert(K);
  return 0;
#undef STORE_F32
#undef STORE_H16
}

bool Model::weights_complete(std::string* missing) const {
  for (const auto& k : required)
    if (!loaded.count(k)) {
      if (missing) *missing = k;
      return false;
    }
  return true;
}

int Model::set_pos_embed(int ph, int pw, const float* table) {
  DAV2_CHECK(table && ph > 0 && pw > 0, "set_pos_embed: bad argument");
  float* p = nullptr;
  if (int rc = upload_f32(table, (int64_t)(1 + ph * pw) * D, &p)) return rc;
  owned.push_back(p);
  pos_tables[std::make_pair(ph, pw)] = p;
  return 0;
}

int Model::buf(const char* name, size_t bytes, void** out) {
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

// ----------------------------------------------------------------------------------------------
// GEMM / conv wrappers
// ----------------------------------------------------------------------------------------------
static GemmParams blank_params(int fmt) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.fmt = fmt;
  return p;
}

int gemm_linear(int mode, const h16* A, int M, int K, long long lda, const h16* Wt, int N, GemmParams p,
                cudaStream_t stream) {
  DAV2_CHECK(N % 4 == 0, "gemm: N=%d must be a multiple of 4", N);
  const int bn = pick_bn(N);
  p.M = M; p.N = N; p.K = K;
  p.num_kb = (K + 63) / 64;
  p.tiles_m = (M + 127) / 128;
  p.tiles_n = (N + bn - 1) / bn;
  const bool two_cta = gemm2_eligible(bn, mode, p.tiles_m);
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128)) return rc;
  if (int rc = make_tmap_2d(&tmB, Wt, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)(two_cta ? bn / 2 : bn))) return rc;
  ProfScope ps(PC_GEMM, 2.0 * M * (double)N * K, 2.0 * ((double)M * K + (double)N * K + (double)M * N), stream);
  if (two_cta) return launch_gemm2(bn, mode, tmA, tmB, p, stream);
  return launch_gemm(bn, mode, tmA, tmB, p, stream);
}

static void pick_conv_tile(int H, int W, int* tw, int* th) {
  const int cands[6][2] = {{16, 8}, {8, 16}, {32, 4}, {4, 32}, {64, 2}, {128, 1}};
  long long best = -1;
  for (int i = 0; i < 6; ++i) {
    const int w = cands[i][0], h = cands[i][1];
    const long long area = (long long)((W + w - 1) / w * w) * ((H + h - 1) / h * h);
    if (best < 0 || area < best) {
      best = area;
      *tw = w;
      *th = h;
    }
  }
}

int conv3x3(int mode, const h16* in, int B, int H, int W, int Cin, const h16* Wp, int Cout, GemmParams p,
            cudaStream_t stream) {
  DAV2_CHECK(Cout % 4 == 0 && Cin % 8 == 0, "conv3x3: Cin=%d Cout=%d unsupported", Cin, Cout);
  const int bn = mode == GM_CONV_HEAD ? 32 : pick_bn(Cout);
  int tw = 16, th = 8;
  pick_conv_tile(H, W, &tw, &th);
  const int cblocks = (Cin + 63) / 64;
  const int Kp = 9 * cblocks * 64;
  // halo-reuse kernel: fixed 8 x 16 output tile, one (10 x 18)-pixel patch per channel block
  const bool halo = conv_halo_eligible(bn, mode, B * ((W + 7) / 8) * ((H + 15) / 16));
  if (halo) { tw = 8; th = 16; }
  const int tiles_m_all = B * ((W + tw - 1) / tw) * ((H + th - 1) / th);
  const bool two_cta = halo || gemm2_eligible(bn, mode, tiles_m_all);
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_nhwc(&tmA, in, (uint64_t)B, (uint64_t)H, (uint64_t)W, (uint64_t)Cin, (uint32_t)(halo ? ConvHaloCfg<256>::HALO_W : tw),
                              (uint32_t)(halo ? 18 : th))) return rc;
  if (int rc = make_tmap_2d(&tmB, Wp, (uint64_t)Cout, (uint64_t)Kp, (uint64_t)Kp, (uint32_t)(two_cta ? bn / 2 : bn))) return rc;
  p.M = B * H * W; p.N = Cout; p.K = Kp;
  p.num_kb = 9 * cblocks;
  p.cblocks = cblocks;
  p.H = H; p.W = W; p.tw = tw; p.th = th;
  p.tiles_x = (W + tw - 1) / tw;
  p.tiles_y = (H + th - 1) / th;
  p.tiles_m = B * p.tiles_x * p.tiles_y;
  p.tiles_n = (Cout + bn - 1) / bn;
  if (p.ldo == 0) p.ldo = Cout;
  ProfScope ps2(PC_CONV, 2.0 * B * H * W * (double)Cout * 9.0 * Cin, 2.0 * ((double)B * H * W * (Cin + Cout) + 9.0 * Cin * Cout), stream);
  if (halo) return launch_conv_halo(bn, mode, tmA, tmB, p, stream);
  if (two_cta) return launch_gemm2(bn, mode, tmA, tmB, p, stream);
  return launch_gemm(bn, mode, tmA, tmB, p, stream);
}

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
#define RC(expr)                 \
  do {                           \
    if (int _rc = (expr)) return _rc; \
  } while (0)

int Model::forward(const float* x, int B, int H, int W, float* depth, cudaStream_t stream) {
  std::string missing;
  DAV2_CHECK(weights_complete(&missing), "forward: weight '%s' was never loaded", missing.c_str());
  DAV2_CHECK(x && depth && B > 0, "forward: null pointer or empty batch");
  {
    int cur = -1;
    DAV2_CUDA_OK(cudaGetDevice(&cur));
    cudaPointerAttributes pa;
    DAV2_CUDA_OK(cudaPointerGetAttributes(&pa, x));
    DAV2_CHECK(cur == device && (pa.type != cudaMemoryTypeDevice || pa.device == device),
               "forward: this handle lives on device %d but the call runs on device %d with x on device %d "
               "(create one handle per device)", device, cur, pa.device);
  }
  DAV2_CHECK(H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0, "forward: H=%d W=%d must be positive multiples of 14", H, W);
  if (fmt == FMT_F32) return forward_fp32(x, B, H, W, depth, stream);
  const int ph = H / 14, pw = W / 14, P = ph * pw, N = P + 1;
  const int M = B * N, MP = B * P;
  const float* posT = nullptr;
  if (ph == 37 && pw == 37) posT = pos;
  else {
    auto it = pos_tables.find(std::make_pair(ph, pw));
    DAV2_CHECK(it != pos_tables.end(), "forward: no position table for a %dx%d patch grid (call dav2_set_pos_embed)", ph, pw);
    posT = it->second;
  }
  const size_t S2 = sizeof(h16);

  h16 *patchA, *XN, *QKV, *ATT, *HID, *TAP[4];
  float* X;
  RC(buf("patch_A", (size_t)MP * KP_PATCH * S2, (void**)&patchA));
  RC(buf("x", (size_t)M * D * 4, (void**)&X));
  RC(buf("xn", (size_t)M * D * S2, (void**)&XN));
  RC(buf("qkv", (size_t)M * 3 * D * S2, (void**)&QKV));
  RC(buf("attn", (size_t)M * D * S2, (void**)&ATT));
  RC(buf("hid", (size_t)M * 4 * D * S2, (void**)&HID));
  for (int i = 0; i < 4; ++i) {
    char nm[16];
    snprintf(nm, sizeof(nm), "tap%d", i);
    RC(buf(nm, (size_t)MP * D * S2, (void**)&TAP[i]));
  }

  // ---- patch embed + cls + pos ------------------------------------------------------------
  RC(launch_patch_im2col(x, patchA, B, H, W, KP_PATCH, fmt, stream));
  {
    GemmParams p = blank_params(fmt);
    p.out = X; p.ldo = D; p.bias = patch_b; p.pos = posT; p.P = P;
    RC(gemm_linear(GM_PATCH, patchA, MP, KP_PATCH, KP_PATCH, patch_w, D, p, stream));
  }
  RC(launch_cls_row(X, cls, posT, B, N, D, stream));

  // ---- transformer blocks --------------------------------------------------------------------
  int next_tap = 0;
  for (int l = 0; l < L; ++l) {
    const BlockW& w = blk[l];
    RC(launch_layernorm(X, w.n1w, w.n1b, XN, M, D, N, 0, 1e-6f, fmt, stream));
    {
      GemmParams p = blank_params(fmt);
      p.out = QKV; p.ldo = 3 * D; p.bias = w.qkv_b;
      RC(gemm_linear(GM_LINEAR_BF16, XN, M, D, D, w.qkv_w, 3 * D, p, stream));
    }
    RC(launch_attention(QKV, ATT, B, N, D, fmt, stream));
    {
      GemmParams p = blank_params(fmt);
      p.out = X; p.ldo = D; p.bias = w.proj_b; p.gamma = w.ls1;
      RC(gemm_linear(GM_LINEAR_RESID, ATT, M, D, D, w.proj_w, D, p, stream));
    }
    RC(launch_layernorm(X, w.n2w, w.n2b, XN, M, D, N, 0, 1e-6f, fmt, stream));
    {
      GemmParams p = blank_params(fmt);
      p.out = HID; p.ldo = 4 * D; p.bias = w.fc1_b; p.act = 1;
      RC(gemm_linear(GM_LINEAR_BF16, XN, M, D, D, w.fc1_w, 4 * D, p, stream));
    }
    {
      GemmParams p = blank_params(fmt);
      p.out = X; p.ldo = D; p.bias = w.fc2_b; p.gamma = w.ls2;
      RC(gemm_linear(GM_LINEAR_RESID, HID, M, 4 * D, 4 * D, w.fc2_w, D, p, stream));
    }
    if (next_tap < 4 && l == cfg.tap_layers[next_tap]) {
      // final norm on the tap, cls dropped, written as the NHWC patch grid [B, ph, pw, D]
      RC(launch_layernorm(X, norm_w, norm_b, TAP[next_tap], M, D, N, 1, 1e-6f, fmt, stream));
      ++next_tap;
    }
  }
  DAV2_CHECK(next_tap == 4, "forward: tap layers must be increasing block indices < depth");

  // ---- DPT head ------------------------------------------------------------------------------
  const int* oc = cfg.out_channels;
  const int hh[4] = {4 * ph, 2 * ph, ph, (ph + 1) / 2};
  const int ww[4] = {4 * pw, 2 * pw, pw, (pw + 1) / 2};
  h16* lvl[4];
  // reassemble: 1x1 projection (+ resize)
  for (int i = 0; i < 4; ++i) {
    char nm[24];
    h16* pr;
    snprintf(nm, sizeof(nm), "proj%d", i);
    RC(buf(nm, (size_t)MP * oc[i] * S2, (void**)&pr));
    GemmParams p = blank_params(fmt);
    p.out = pr; p.ldo = oc[i]; p.bias = proj_b[i];
    RC(gemm_linear(GM_LINEAR_BF16, TAP[i], MP, D, D, proj_w[i], oc[i], p, stream));
    if (i == 2) {
      lvl[i] = pr;
    } else if (i < 2) {
      const int s = i == 0 ? 4 : 2;
      snprintf(nm, sizeof(nm), "lvl%d", i);
      RC(buf(nm, (size_t)B * hh[i] * ww[i] * oc[i] * S2, (void**)&lvl[i]));
      GemmParams q = blank_params(fmt);
      q.out = lvl[i]; q.bias = rs_b[i]; q.convt_s = s; q.convt_cout = oc[i]; q.H = ph; q.W = pw;
      RC(gemm_linear(GM_CONVT, pr, MP, oc[i], oc[i], rs_w[i], s * s * oc[i], q, stream));
    } else {
      h16* col;
      RC(buf("lvl3_im2col", (size_t)B * hh[3] * ww[3] * 9 * oc[3] * S2, (void**)&col));
      RC(buf("lvl3", (size_t)B * hh[3] * ww[3] * oc[3] * S2, (void**)&lvl[3]));
      RC(launch_im2col_s2(pr, col, B, ph, pw, oc[3], stream));
      GemmParams q = blank_params(fmt);
      q.out = lvl[3]; q.ldo = oc[3]; q.bias = rs_b[3];
      RC(gemm_linear(GM_LINEAR_BF16, col, B * hh[3] * ww[3], 9 * oc[3], 9 * oc[3], rs_w[3], oc[3], q, stream));
    }
  }
  // layer_rn 3x3 (no bias): keep x and relu(x)
  h16 *rn[4], *rnr[4];
  for (int i = 0; i < 4; ++i) {
    char nm[24];
    snprintf(nm, sizeof(nm), "rn%d", i);
    RC(buf(nm, (size_t)B * hh[i] * ww[i] * F * S2, (void**)&rn[i]));
    snprintf(nm, sizeof(nm), "rn%d_relu", i);
    RC(buf(nm, (size_t)B * hh[i] * ww[i] * F * S2, (void**)&rnr[i]));
    GemmParams p = blank_params(fmt);
    p.out = rn[i]; p.out_relu = rnr[i];
    RC(conv3x3(GM_CONV_BF16, lvl[i], B, hh[i], ww[i], oc[i], rn_w[i], F, p, stream));
  }
  // fusion blocks (refinenet4 -> refinenet1).  out_conv (1x1) commutes with the bilinear resize
  // (both linear, bilinear weights sum to 1), so it runs at the LOW resolution: 4x fewer FLOPs.
  const size_t big = (size_t)B * hh[0] * ww[0] * F * S2;
  h16 *T, *S, *SR, *Y, *OCb;
  RC(buf("scratch_t", big, (void**)&T));
  RC(buf("scratch_s", big, (void**)&S));
  RC(buf("scratch_sr", big, (void**)&SR));
  RC(buf("scratch_y", big, (void**)&Y));
  RC(buf("scratch_oc", big, (void**)&OCb));
  h16* up_prev = nullptr;
  for (int i = 3; i >= 0; --i) {
    const Fusion& f = ref[i];
    const int h = hh[i], w = ww[i];
    const h16 *in = rn[i], *in_relu = rnr[i];
    if (up_prev) {
      // S = resConfUnit1(rn) + up_prev ; SR = relu(S)
      GemmParams p = blank_params(fmt);
      p.out = T; p.bias = f.rcu_b[0][0]; p.act = 2;
      RC(conv3x3(GM_CONV_BF16, in_relu, B, h, w, F, f.rcu_w[0][0], F, p, stream));
      GemmParams q = blank_params(fmt);
      q.out = S; q.out_relu = SR; q.bias = f.rcu_b[0][1]; q.add1 = in; q.add2 = up_prev;
      RC(conv3x3(GM_CONV_BF16, T, B, h, w, F, f.rcu_w[0][1], F, q, stream));
      in = S;
      in_relu = SR;
    }
    {
      // Y = resConfUnit2(in)
      GemmParams p = blank_params(fmt);
      p.out = T; p.bias = f.rcu_b[1][0]; p.act = 2;
      RC(conv3x3(GM_CONV_BF16, in_relu, B, h, w, F, f.rcu_w[1][0], F, p, stream));
      GemmParams q = blank_params(fmt);
      q.out = Y; q.bias = f.rcu_b[1][1]; q.add1 = in;
      RC(conv3x3(GM_CONV_BF16, T, B, h, w, F, f.rcu_w[1][1], F, q, stream));
    }
    {
      GemmParams p = blank_params(fmt);
      p.out = OCb; p.ldo = F; p.bias = f.out_b;
      RC(gemm_linear(GM_LINEAR_BF16, Y, B * h * w, F, F, f.out_w, F, p, stream));
    }
    const int ho = i > 0 ? hh[i - 1] : 2 * hh[0], wo = i > 0 ? ww[i - 1] : 2 * ww[0];
    char nm[24];
    snprintf(nm, sizeof(nm), "path%d", i + 1);
    h16* up;
    RC(buf(nm, (size_t)B * ho * wo * F * S2, (void**)&up));
    RC(launch_bilinear_nhwc(OCb, up, B, h, w, ho, wo, F, fmt, stream));
    up_prev = up;
  }
  // head: output_conv1 -> bilinear to (H, W) -> 3x3 + ReLU + 1x1 + sigmoid * max_depth
  const int h8 = 2 * hh[0], w8 = 2 * ww[0];
  h16 *O1, *O1U;
  RC(buf("out1", (size_t)B * h8 * w8 * (F / 2) * S2, (void**)&O1));
  RC(buf("out1_up", (size_t)B * H * W * (F / 2) * S2, (void**)&O1U));
  {
    GemmParams p = blank_params(fmt);
    p.out = O1; p.bias = oc1_b;
    RC(conv3x3(GM_CONV_BF16, up_prev, B, h8, w8, F, oc1_w, F / 2, p, stream));
  }
  RC(launch_bilinear_nhwc(O1, O1U, B, h8, w8, H, W, F / 2, fmt, stream));
  {
    GemmParams p = blank_params(fmt);
    p.out = depth; p.bias = oc2_b; p.head_w = oc3_w; p.head_b = oc3_b; p.max_depth = cfg.max_depth;
    if (capture_logits) RC(buf("logits", (size_t)B * H * W * 4, (void**)&p.out_logit));
    RC(conv3x3(GM_CONV_HEAD, O1U, B, H, W, F / 2, oc2_w, 32, p, stream));
  }
  return 0;
}

int Model::debug_read(const char* name, void* dst, int64_t bytes, cudaStream_t stream) {
  auto it = ws.find(name);
  DAV2_CHECK(it != ws.end(), "debug_read: no buffer named '%s'", name);
  DAV2_CHECK(bytes >= 0 && (size_t)bytes <= it->second.bytes, "debug_read: %lld bytes requested, buffer '%s' holds %zu", (long long)bytes, name, it->second.bytes);
  DAV2_CUDA_OK(cudaMemcpyAsync(dst, it->second.p, (size_t)bytes, cudaMemcpyDeviceToDevice, stream));
  return 0;
}

int Model::debug_buffer(const char* name, void** ptr, int64_t* bytes) {
  auto it = ws.find(name);
  DAV2_CHECK(it != ws.end(), "debug_buffer: no buffer named '%s'", name);
  *ptr = it->second.p;
  *bytes = (int64_t)it->second.bytes;
  return 0;
}

}  // namespace dav2
