// Host-side helpers: last-error string, TMA tensor-map encoding, device properties.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <set>
#include <utility>
#include <stdarg.h>
#include <string.h>

namespace dav2 {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  DAV2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  DAV2_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  DAV2_CHECK((ld_elems * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%llu)",
             (unsigned long long)ld_elems);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DAV2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box_rows=%u) failed: %d",
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, (int)r);
  return 0;
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                   uint32_t tw, uint32_t th) {
  EncodeTiledFn fn = encode_fn();
  DAV2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  DAV2_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  DAV2_CHECK((C * 2) % 16 == 0, "NHWC channel count must be a multiple of 8 (C=%llu)", (unsigned long long)C);
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {64, tw, th, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DAV2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc B=%llu H=%llu W=%llu C=%llu box %ux%u) failed: %d",
             (unsigned long long)B, (unsigned long long)H, (unsigned long long)W, (unsigned long long)C, tw, th,
             (int)r);
  return 0;
}

// cudaFuncSetAttribute is per (function, device).  `device_setup_done(tag)` says whether the set-up keyed by `tag` (the
// address of a function-local static) has SUCCEEDED on the current device; the caller runs the set-up, returning early
// on any error, and only then calls `device_setup_mark(tag)` -- a failed set-up is therefore retried by the next call
// instead of being skipped for ever.  Works when several devices are driven from one process.
static std::mutex g_setup_mu;
static std::set<std::pair<int, const void*>> g_setup_seen;
bool device_setup_done(const void* tag) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_setup_mu);
  return g_setup_seen.count(std::make_pair(dev, tag)) != 0;
}
void device_setup_mark(const void* tag) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_setup_mu);
  g_setup_seen.insert(std::make_pair(dev, tag));
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace dav2

// ----------------------------------------------------------------------------------------------
// per-kernel-class event timing
// ----------------------------------------------------------------------------------------------
#include <vector>
namespace dav2 {
struct ProfRec { int cls; double flops, bytes; cudaEvent_t a, b; };
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static const char* kClassNames[PC_COUNT] = {"gemm_tcgen05", "conv_tcgen05", "attention_tcgen05", "layernorm", "resample",
                                            "im2col", "backproject", "depth_metrics", "other"};
void prof_enable(int on) { g_prof = on != 0; }
bool prof_enabled() { return g_prof; }
static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_begin(int cls, double flops, double bytes, cudaStream_t stream) {
  ProfRec r; r.cls = cls; r.flops = flops; r.bytes = bytes; r.a = get_event(); r.b = get_event();
  cudaEventRecord(r.a, stream);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t stream) { if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, stream); }
int prof_report(char* buf, int cap) {
  double ms[PC_COUNT] = {0}, fl[PC_COUNT] = {0}, by[PC_COUNT] = {0}; long long n[PC_COUNT] = {0};
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.b);
    float t = 0.f; cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.cls] += t; fl[r.cls] += r.flops; by[r.cls] += r.bytes; n[r.cls] += 1;
    g_pool.push_back(r.a); g_pool.push_back(r.b);
  }
  g_recs.clear();
  int off = snprintf(buf, cap, "{");
  bool first = true;
  for (int c = 0; c < PC_COUNT && off < cap; ++c) {
    if (!n[c]) continue;
    off += snprintf(buf + off, cap - off, "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
                    first ? "" : ", ", kClassNames[c], n[c], ms[c], fl[c], by[c]);
    first = false;
  }
  if (off < cap) off += snprintf(buf + off, cap - off, "}");
  return off < cap ? 0 : -1;
}
}  // namespace dav2
