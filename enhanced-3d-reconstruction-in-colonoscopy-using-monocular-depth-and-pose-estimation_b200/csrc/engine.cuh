// Model state for the DepthAnythingV2 forward engine (engine.cu) behind the C ABI (api.cu).
#pragma once

#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dav2_b200.h"
#include "elementwise.cuh"
#include "conv_halo_tcgen05.cuh"

namespace dav2 {

struct BlockW {
  float *n1w, *n1b, *n2w, *n2b, *qkv_b, *proj_b, *fc1_b, *fc2_b, *ls1, *ls2;
  h16 *qkv_w, *proj_w, *fc1_w, *fc2_w;
};
struct Fusion {
  h16* out_w;
  float* out_b;
  h16* rcu_w[2][2];   // [resConfUnit 1|2][conv 1|2], tap-major packed
  float* rcu_b[2][2];
};
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0, bytes = 0;
};

struct Model {
  dav2_config cfg;
  int D, L, heads, F;
  int device = 0;  // the CUDA device the handle (weights + workspace) lives on; forward() refuses any other
  int fmt;  // FMT_F16 (default; the reference's AMP precision), FMT_BF16, or FMT_F32 (fp32 validation engine)
  // encoder
  h16* patch_w = nullptr;
  float *patch_b = nullptr, *cls = nullptr, *pos = nullptr, *norm_w = nullptr, *norm_b = nullptr;
  std::vector<BlockW> blk;
  // DPT head
  h16* proj_w[4];
  float* proj_b[4];
  h16* rs_w[4];
  float* rs_b[4];
  h16* rn_w[4];
  Fusion ref[4];  // refinenet1..4
  h16 *oc1_w = nullptr, *oc2_w = nullptr;
  float *oc1_b = nullptr, *oc2_b = nullptr, *oc3_w = nullptr;
  float oc3_b = 0.f;

  bool capture_logits = false;  // dav2_set_capture_logits: keep the pre-sigmoid logits of the next forwards in buffer "logits"

  std::set<std::string> required, loaded;
  std::map<std::pair<int, int>, float*> pos_tables;
  std::vector<void*> owned;
  std::map<std::string, DevBuf> ws;

  explicit Model(const dav2_config& c);
  ~Model();
  int set_weight(const char* key, const float* data, const int64_t* shape, int ndim);
  bool weights_complete(std::string* missing) const;
  int set_pos_embed(int ph, int pw, const float* table);
  int buf(const char* name, size_t bytes, void** out);
  int forward(const float* x, int B, int H, int W, float* depth, cudaStream_t stream);
  int forward_fp32(const float* x, int B, int H, int W, float* depth, cudaStream_t stream);  // fp32_path.cu
  int debug_buffer(const char* name, void** ptr, int64_t* bytes);
  int debug_read(const char* name, void* dst, int64_t bytes, cudaStream_t stream);
};

int gemm_linear(int mode, const h16* A, int M, int K, long long lda, const h16* Wt, int N, GemmParams p,
                cudaStream_t stream);
int conv3x3(int mode, const h16* in, int B, int H, int W, int Cin, const h16* Wp, int Cout, GemmParams p,
            cudaStream_t stream);

}  // namespace dav2
