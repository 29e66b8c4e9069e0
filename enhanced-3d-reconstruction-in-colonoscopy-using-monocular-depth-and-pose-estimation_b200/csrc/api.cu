// extern "C" surface of libdav2_b200.so (declared in include/dav2_b200.h).
#include <string.h>
#include <new>

#include "posenet.cuh"

using namespace dav2;

struct dav2_pose {
  PoseModel impl;
  explicit dav2_pose(int precision) : impl(precision) {}
};

struct dav2_model {
  Model impl;
  explicit dav2_model(const dav2_config& c) : impl(c) {}
};

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int require_sm100() {
  int dev = 0;
  DAV2_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  DAV2_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DAV2_CHECK(major == 10, "dav2_b200 needs an sm_100 (B200) device; found compute capability major %d. There is no fallback path.", major);
  return 0;
}

extern "C" {

int dav2_create(dav2_model** out, const dav2_config* cfg) {
  DAV2_CHECK(out && cfg, "dav2_create: null argument");
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(cfg->embed_dim % 128 == 0 && cfg->embed_dim <= 1024 && cfg->embed_dim == cfg->num_heads * 64,
             "dav2_create: embed_dim=%d heads=%d unsupported (need D = 64*heads, multiple of 128, <= 1024)",
             cfg->embed_dim, cfg->num_heads);
  DAV2_CHECK(cfg->precision >= 0 && cfg->precision <= 2, "dav2_create: precision must be 0 (fp16), 1 (bf16) or 2 (fp32 validation engine)");
  DAV2_CHECK(cfg->features % 16 == 0 && cfg->depth > 0, "dav2_create: features=%d depth=%d unsupported", cfg->features, cfg->depth);
  for (int i = 0; i < 4; ++i)
    DAV2_CHECK(cfg->out_channels[i] % 8 == 0 && cfg->tap_layers[i] >= 0 && cfg->tap_layers[i] < cfg->depth,
               "dav2_create: out_channels / tap_layers entry %d invalid", i);
  dav2_model* m = new (std::nothrow) dav2_model(*cfg);
  DAV2_CHECK(m != nullptr, "dav2_create: out of host memory");
  *out = m;
  return 0;
}

void dav2_destroy(dav2_model* m) { delete m; }

int dav2_set_weight(dav2_model* m, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
  DAV2_CHECK(m, "null model");
  return m->impl.set_weight(key, data, shape, ndim);
}

int dav2_weights_complete(const dav2_model* m) {
  if (!m) return 0;
  std::string missing;
  const bool ok = m->impl.weights_complete(&missing);
  if (!ok) set_last_error("missing weight: %s", missing.c_str());
  return ok ? 1 : 0;
}

int dav2_set_pos_embed(dav2_model* m, int32_t ph, int32_t pw, const float* table) {
  DAV2_CHECK(m, "null model");
  return m->impl.set_pos_embed(ph, pw, table);
}

int dav2_forward(dav2_model* m, const float* x, int32_t B, int32_t H, int32_t W, float* depth, void* stream) {
  DAV2_CHECK(m, "null model");
  return m->impl.forward(x, B, H, W, depth, S(stream));
}

int dav2_set_capture_logits(dav2_model* m, int32_t on) {
  DAV2_CHECK(m, "null model");
  DAV2_CHECK(m->impl.fmt != FMT_F32 || !on, "dav2_set_capture_logits: not available in the fp32 validation engine");
  m->impl.capture_logits = on != 0;
  return 0;
}

int dav2_debug_buffer(dav2_model* m, const char* name, void** ptr, int64_t* bytes) {
  DAV2_CHECK(m && name && ptr && bytes, "dav2_debug_buffer: null argument");
  return m->impl.debug_buffer(name, ptr, bytes);
}

int dav2_debug_read(dav2_model* m, const char* name, void* dst, int64_t bytes, void* stream) {
  DAV2_CHECK(m && name && dst, "dav2_debug_read: null argument");
  return m->impl.debug_read(name, dst, bytes, S(stream));
}

int dav2_resize_depth(const float* in, int32_t B, int32_t Hi, int32_t Wi, float* out, int32_t Ho, int32_t Wo,
                      void* stream) {
  DAV2_CHECK(in && out, "dav2_resize_depth: null pointer");
  if (int rc = require_sm100()) return rc;
  return launch_bilinear_f32(in, out, B, Hi, Wi, Ho, Wo, S(stream));
}

int dav2_preprocess_bgr_u8(const uint8_t* img, int32_t H, int32_t W, float* out, int32_t nh, int32_t nw, void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_preprocess_bgr(img, 1, H, W, out, nh, nw, S(stream));
}

int dav2_preprocess_bgr_u8_batch(const uint8_t* img, int32_t B, int32_t H, int32_t W, float* out, int32_t nh, int32_t nw,
                                 void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_preprocess_bgr(img, B, H, W, out, nh, nw, S(stream));
}

int dav2_resize_aa(int32_t mode, const void* in, int32_t B, int32_t H, int32_t W, float* out, int32_t Ho, int32_t Wo,
                   float div_in, void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_resample_aa(mode, in, B, H, W, out, Ho, Wo, div_in, S(stream));
}

int dav2_backproject(const float* depth, int32_t B, int32_t H, int32_t W, const double* K4, int32_t k_per_frame,
                     const double* T12, float depth_scale, float depth_trunc, float* xyz, uint8_t* valid,
                     int32_t* counts, void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_backproject(depth, B, H, W, K4, k_per_frame, T12, depth_scale, depth_trunc, xyz, valid, counts, S(stream));
}

static int backproject_dsts(const char* who, int32_t H, int32_t W, float* const* xyz_dst, uint8_t* const* valid_dst,
                            int32_t* const* counts_dst, int32_t n_dst, int64_t frame_offset, float** x, uint8_t** v, int** c) {
  if (!xyz_dst || n_dst < 1 || n_dst > 8 || frame_offset < 0 || H <= 0 || W <= 0) {
    dav2::set_last_error("%s: bad destination list", who);
    return -2;
  }
  const long long HW = (long long)H * W;
  for (int p = 0; p < n_dst; ++p) {
    x[p] = xyz_dst[p] ? xyz_dst[p] + frame_offset * HW * 3 : nullptr;
    v[p] = (valid_dst && valid_dst[p]) ? valid_dst[p] + frame_offset * HW : nullptr;
    c[p] = (counts_dst && counts_dst[p]) ? counts_dst[p] + frame_offset : nullptr;
  }
  return 0;
}

int dav2_backproject_gather(const float* depth, int32_t B, int32_t H, int32_t W, const double* K4, int32_t k_per_frame,
                            const double* T12, float depth_scale, float depth_trunc, float* const* xyz_dst,
                            uint8_t* const* valid_dst, int32_t* const* counts_dst, int32_t n_dst, int64_t frame_offset,
                            void* stream) {
  if (int rc = require_sm100()) return rc;
  float* x[8];
  uint8_t* v[8];
  int* c[8];
  if (int rc = backproject_dsts("backproject_gather", H, W, xyz_dst, valid_dst, counts_dst, n_dst, frame_offset, x, v, c)) return rc;
  return launch_backproject_multi(depth, B, H, W, K4, k_per_frame, T12, depth_scale, depth_trunc, x, valid_dst ? v : nullptr,
                                  counts_dst ? c : nullptr, n_dst, S(stream));
}

int dav2_backproject_metrics(const float* depth, const float* gt, int32_t B, int32_t H, int32_t W, const double* K4,
                             int32_t k_per_frame, const double* T12, float depth_scale, float depth_trunc,
                             float* const* xyz_dst, uint8_t* const* valid_dst, int32_t* const* counts_dst, int32_t n_dst,
                             int64_t frame_offset, float lo, float hi, int32_t per_frame, double* partials, void* stream) {
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(gt && partials, "backproject_metrics: null gt / partials");
  float* x[8];
  uint8_t* v[8];
  int* c[8];
  if (int rc = backproject_dsts("backproject_metrics", H, W, xyz_dst, valid_dst, counts_dst, n_dst, frame_offset, x, v, c)) return rc;
  return launch_backproject_multi(depth, B, H, W, K4, k_per_frame, T12, depth_scale, depth_trunc, x, valid_dst ? v : nullptr,
                                  counts_dst ? c : nullptr, n_dst, S(stream), gt, lo, hi, per_frame, partials);
}

// ---- peer-mapped buffers (CUDA IPC) for the fused gather ----
int dav2_peer_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) { dav2::set_last_error("peer_alloc: bad arguments"); return -2; }
  DAV2_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));
  return 0;
}
int dav2_peer_free(void* ptr) {
  DAV2_CUDA_OK(cudaFree(ptr));
  return 0;
}
int dav2_peer_export(const void* ptr, uint8_t* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle64) { dav2::set_last_error("peer_export: null pointer"); return -2; }
  cudaIpcMemHandle_t h;
  DAV2_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, 64);
  return 0;
}
int dav2_peer_open(const uint8_t* handle64, void** ptr) {
  if (!ptr || !handle64) { dav2::set_last_error("peer_open: null pointer"); return -2; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DAV2_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int dav2_peer_close(void* ptr) {
  DAV2_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int dav2_voxel_downsample(const float* xyz, const float* rgb, const uint8_t* valid, int64_t n, double voxel_size, float* out_xyz,
                          float* out_rgb, int64_t* out_count, void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_voxel_downsample(xyz, rgb, valid, (long long)n, voxel_size, out_xyz, out_rgb, (long long*)out_count, S(stream));
}

int dav2_depth_metrics(const float* pred, const float* gt, int32_t B, int64_t HW, float lo, float hi,
                       int32_t variant, int32_t per_frame, double* partials, void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_depth_metrics(pred, gt, B, HW, lo, hi, variant, per_frame, partials, S(stream));
}

int dav2_transform_points(float* xyz, int64_t n, const double* T12, void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_transform_points(xyz, (long long)n, T12, S(stream));
}

int dav2_compose_poses(const float* rel, const float* init7, int32_t N, float* abs7, double* T12, void* stream) {
  if (int rc = require_sm100()) return rc;
  return launch_compose_poses(rel, init7, N, abs7, T12, S(stream));
}

static int check_fmt(int32_t fmt) {
  DAV2_CHECK(fmt == FMT_F16 || fmt == FMT_BF16, "fmt must be 0 (fp16) or 1 (bf16), got %d", fmt);
  return 0;
}

int dav2_linear_h16(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N, int32_t K,
                    int32_t act, int32_t fmt, void* stream) {
  if (int rc = check_fmt(fmt)) return rc;
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(A && W && C && K % 8 == 0, "dav2_linear_h16: null pointer or K %% 8 != 0");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = C; p.ldo = N; p.bias = bias; p.act = act; p.fmt = fmt;
  return gemm_linear(GM_LINEAR_BF16, (const h16*)A, M, K, K, (const h16*)W, N, p, S(stream));
}

int dav2_linear_resid(const void* A, const void* W, const float* bias, const float* gamma, float* x, int32_t M,
                      int32_t N, int32_t K, int32_t fmt, void* stream) {
  if (int rc = check_fmt(fmt)) return rc;
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(A && W && x && gamma && K % 8 == 0, "dav2_linear_resid: null pointer or K %% 8 != 0");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = x; p.ldo = N; p.bias = bias; p.gamma = gamma; p.fmt = fmt;
  return gemm_linear(GM_LINEAR_RESID, (const h16*)A, M, K, K, (const h16*)W, N, p, S(stream));
}

int dav2_conv3x3_h16(const void* in, const void* Wp, const float* bias, const void* add1, const void* add2,
                     void* out, void* out_relu, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                     int32_t act, int32_t fmt, void* stream) {
  if (int rc = check_fmt(fmt)) return rc;
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(in && Wp && out, "dav2_conv3x3_h16: null pointer");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = out; p.out_relu = (h16*)out_relu; p.bias = bias; p.add1 = (const h16*)add1; p.add2 = (const h16*)add2;
  p.act = act; p.fmt = fmt;
  return conv3x3(GM_CONV_BF16, (const h16*)in, B, H, W, Cin, (const h16*)Wp, Cout, p, S(stream));
}

int dav2_attention_h16(const void* qkv, void* out, int32_t B, int32_t N, int32_t D, int32_t fmt, void* stream) {
  if (int rc = check_fmt(fmt)) return rc;
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(qkv && out, "dav2_attention_h16: null pointer");
  uint32_t lbo = 1024, sbo = 1024;
#ifdef DAV2_PROFILING_KNOBS  // descriptor-stride experiments; never compiled into the shipped library
  if (const char* e = getenv("DAV2_ATT_VLBO")) lbo = (uint32_t)atoi(e);
  if (const char* e = getenv("DAV2_ATT_VSBO")) sbo = (uint32_t)atoi(e);
#endif
  return launch_attention((const h16*)qkv, (h16*)out, B, N, D, fmt, S(stream), lbo, sbo);
}

int dav2_layernorm(const float* x, const float* w, const float* b, void* out, int64_t rows, int32_t D, float eps,
                   int32_t fmt, void* stream) {
  if (int rc = check_fmt(fmt)) return rc;
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(x && w && b && out, "dav2_layernorm: null pointer");
  return launch_layernorm(x, w, b, (h16*)out, rows, D, 1, 0, eps, fmt, S(stream));
}

int dav2_bilinear_nhwc_h16(const void* in, void* out, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo,
                           int32_t C, int32_t fmt, void* stream) {
  if (int rc = check_fmt(fmt)) return rc;
  if (int rc = require_sm100()) return rc;
  DAV2_CHECK(in && out, "dav2_bilinear_nhwc_h16: null pointer");
  return launch_bilinear_nhwc((const h16*)in, (h16*)out, B, Hi, Wi, Ho, Wo, C, fmt, S(stream));
}

int dav2_pose_create(dav2_pose** out, int32_t precision) {
  DAV2_CHECK(out && (precision == 0 || precision == 1), "dav2_pose_create: bad argument");
  if (int rc = require_sm100()) return rc;
  dav2_pose* m = new (std::nothrow) dav2_pose(precision);
  DAV2_CHECK(m != nullptr, "dav2_pose_create: out of host memory");
  *out = m;
  return 0;
}
void dav2_pose_destroy(dav2_pose* m) { delete m; }
int dav2_pose_set_weight(dav2_pose* m, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
  DAV2_CHECK(m, "null pose model");
  return m->impl.set_weight(key, data, shape, ndim);
}
int dav2_pose_forward(dav2_pose* m, const float* x, int32_t B, int32_t H, int32_t W, float* pose7, void* stream) {
  DAV2_CHECK(m, "null pose model");
  return m->impl.forward(x, B, H, W, pose7, S(stream));
}

const char* dav2_last_error(void) { return get_last_error(); }
int64_t dav2_launch_count(void) { return launch_count(); }
const char* dav2_version(void) { return "dav2_b200 0.1 (sm_100a: tcgen05/TMEM/TMA)"; }

}  // extern "C"

extern "C" {
void dav2_profile_enable(int32_t on) { prof_enable(on); }
int dav2_profile_report(char* buf, int32_t cap) { return prof_report(buf, cap); }
}
