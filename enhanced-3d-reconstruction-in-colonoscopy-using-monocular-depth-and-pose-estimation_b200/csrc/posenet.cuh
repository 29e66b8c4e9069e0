// PoseEstimationNet engine (posenet.cu) behind dav2_pose_* (api.cu).
#pragma once

#include "engine.cuh"

namespace dav2 {

struct ConvW {
  h16* w = nullptr;
  int Cout = 0, Cin = 0, ks = 0, K = 0;  // K = packed row length
};

struct PoseModel {
  int fmt;
  int device = 0;  // the CUDA device the handle lives on
  std::map<std::string, ConvW> conv;
  std::map<std::string, float*> f32;
  std::vector<void*> owned;
  std::map<std::string, DevBuf> ws;

  explicit PoseModel(int precision);
  ~PoseModel();
  int set_weight(const char* key, const float* data, const int64_t* shape, int ndim);
  int forward(const float* x, int B, int H, int W, float* out7, cudaStream_t stream);

 private:
  int buf(const char* name, size_t bytes, void** out);
  int get_conv(const std::string& name, ConvW* w, const float** bias);
  int conv_im2col(const std::string& name, const h16* in, int B, int H, int W, int stride, int pad, int act, h16* out, int* Ho,
                  int* Wo, cudaStream_t stream);
  int add_relu(const h16* a, const h16* b, h16* out, long long n, cudaStream_t stream);  // standalone residual add (kept for non-fused callers)
};

}  // namespace dav2
