// Declarations of the helper-kernel launchers (elementwise.cu, attention.cu, geometry.cu).
#pragma once
#include "common.cuh"

namespace dav2 {

int launch_layernorm(const float* x, const float* w, const float* b, h16* out, long long rows, int D,
                     int rows_per_img, int skip_cls, float eps, int fmt, cudaStream_t stream);
int launch_patch_im2col(const float* x, h16* A, int B, int H, int W, int KP, int fmt, cudaStream_t stream);
int launch_cls_row(float* x, const float* cls, const float* pos, int B, int ntok, int D, cudaStream_t stream);
int launch_im2col_s2(const h16* in, h16* A, int B, int H, int W, int C, cudaStream_t stream);
int launch_bilinear_nhwc(const h16* in, h16* out, int B, int Hi, int Wi, int Ho, int Wo, int C, int fmt, cudaStream_t stream);
int launch_bilinear_f32(const float* in, float* out, int B, int Hi, int Wi, int Ho, int Wo, cudaStream_t stream);
int launch_cast_bf16(const float* in, h16* out, long long n, cudaStream_t stream);

int launch_preprocess_bgr(const uint8_t* img, int B, int H, int W, float* out, int nh, int nw, cudaStream_t stream);
int launch_resample_aa(int mode, const void* in, int B, int H, int W, float* out, int Ho, int Wo, float div_in,
                       cudaStream_t stream);

// voxel.cu
int launch_voxel_downsample(const float* xyz, const float* rgb, const uint8_t* valid, long long n, double voxel, float* out_xyz,
                            float* out_rgb, long long* out_count, cudaStream_t stream);

// attention.cu
int launch_attention(const h16* qkv, h16* out, int B, int N, int D, int fmt, cudaStream_t stream, uint32_t v_lbo = 1024,
                     uint32_t v_sbo = 1024);

// geometry.cu
int launch_backproject_multi(const float* depth, int B, int H, int W, const double* K4, int k_per_frame, const double* T12,
                             float depth_scale, float depth_trunc, float* const* xyz, uint8_t* const* valid,
                             int* const* counts, int n_dst, cudaStream_t stream, const float* gt = nullptr, float lo = 0.f,
                             float hi = 0.f, int per_frame = 0, double* partials = nullptr);
int launch_backproject(const float* depth, int B, int H, int W, const double* K4, int k_per_frame, const double* T12,
                       float depth_scale, float depth_trunc, float* xyz, uint8_t* valid, int* counts,
                       cudaStream_t stream);
int launch_depth_metrics(const float* pred, const float* gt, int B, long long HW, float lo, float hi, int variant,
                         int per_frame, double* partials, cudaStream_t stream);
int launch_transform_points(float* xyz, long long n, const double* T12, cudaStream_t stream);
int launch_compose_poses(const float* rel, const float* init7, int N, float* abs7, double* T12, cudaStream_t stream);

}  // namespace dav2
