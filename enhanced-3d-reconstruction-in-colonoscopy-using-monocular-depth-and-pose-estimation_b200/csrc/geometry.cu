// HBM-bound geometry / metric kernels of the depth -> point-cloud path.
//
//  backproject_kernel   fused pinhole back-projection + SE(3) world transform + validity mask
//                       (reference depth_to_pointcloud.py:218-239 via Open3D, formula
//                        depth_to_pointcloud_dav2.py:300-313).  16 B/px algorithmic traffic (+1 B mask).
//  depth_metrics_kernel one-pass masked reductions for BOTH metric definitions
//                       (eval/evaluation.py:16-60 and calculate_metrics.py:17-55) -> fp64 partial sums,
//                       finalised on the host AFTER any cross-GPU all-reduce.  8 B/px.
//  compose_poses_kernel the strictly sequential fp32 pose chain of eval/evaluation.py:279-382
//                       (+ quaternion -> [R|t] rows in fp64, depth_to_pointcloud.py:168-173).
#include "elementwise.cuh"

namespace dav2 {

// ----------------------------------------------------------------------------------------------
// back-projection: 8 pixels per thread (two float4 loads issued up front), 6x float4 + 2x uchar4 streaming
// stores.  fp64 math in registers; the fp64<->fp32 / int->fp64 conversions run on the XU pipe (16/clk/SM),
// which is what bounded the first version (ncu: XU 62 %, 34 % occupancy, 3.8 TB/s) -- so the pixel
// coordinates are converted once per float4 and advanced by fp64 adds, and (row, col) comes from a
// multiply-high instead of an integer division (another three XU ops).
// ----------------------------------------------------------------------------------------------
struct BpFrame {
  double fx_inv, fy_inv, cx, cy, inv_scale, trunc;
  double T[12];
  bool hasT;
};

__device__ __forceinline__ bool backproject_one(float d, double xf, double yf, const BpFrame& f, float& X, float& Y, float& Z) {
  double z = (double)d * f.inv_scale;
  const bool ok = (z > 0.0) && (z < f.trunc) && isfinite(d);  // NaN fails z > 0
  double x = xf * z;
  double y = yf * z;
  if (f.hasT) {
    const double xw = f.T[0] * x + f.T[1] * y + f.T[2] * z + f.T[3];
    const double yw = f.T[4] * x + f.T[5] * y + f.T[6] * z + f.T[7];
    const double zw = f.T[8] * x + f.T[9] * y + f.T[10] * z + f.T[11];
    x = xw; y = yw; z = zw;
  }
  X = ok ? (float)x : 0.f; Y = ok ? (float)y : 0.f; Z = ok ? (float)z : 0.f;
  return ok;
}

// 4 consecutive pixels starting at linear index p0 (row-major); returns the number of valid ones
__device__ __forceinline__ int backproject_quad(const float4 d4, unsigned p0, int W, unsigned wmagic, const BpFrame& f,
                                                float (&o)[12], uchar4& m) {
  // v = p0 / W: multiply-high by ceil(2^32 / W) is exact while p0 * W < 2^32 (checked by the launcher)
  const unsigned v = wmagic ? __umulhi(p0, wmagic) : p0 / (unsigned)W;
  int u = (int)(p0 - v * (unsigned)W);
  double ud = (double)u, vd = (double)v;
  const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
  uint8_t* mm = &m.x;
  int nvalid = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // x = (u-cx)/fx * z: the oracle divides; 1/fx in fp64 then one multiply differs by <= 1 ulp(fp64)
    const double xf = (ud - f.cx) * f.fx_inv, yf = (vd - f.cy) * f.fy_inv;
    const bool ok = backproject_one(dd[k], xf, yf, f, o[3 * k], o[3 * k + 1], o[3 * k + 2]);
    mm[k] = ok ? 1 : 0;
    nvalid += ok ? 1 : 0;
    const bool wrap = (++u == W);
    u = wrap ? 0 : u;
    ud = wrap ? 0.0 : ud + 1.0;
    vd = wrap ? vd + 1.0 : vd;
  }
  return nvalid;
}

template <bool VEC4>
__global__ void __launch_bounds__(256, 3) backproject_kernel(const float* __restrict__ depth, int H, int W, unsigned wmagic,
                                                             const double* __restrict__ K4, int k_per_frame,
                                                             const double* __restrict__ T12, double inv_scale,
                                                             double trunc, float* __restrict__ xyz,
                                                             uint8_t* __restrict__ valid, int* __restrict__ counts) {
  const int b = blockIdx.y;
  const long long HW = (long long)H * W;
  const double* K = K4 + (k_per_frame ? 4 * b : 0);
  BpFrame f;
  f.fx_inv = 1.0 / K[0]; f.fy_inv = 1.0 / K[1]; f.cx = K[2]; f.cy = K[3];
  f.inv_scale = inv_scale; f.trunc = trunc;
  f.hasT = T12 != nullptr;
#pragma unroll
  for (int i = 0; i < 12; ++i) f.T[i] = f.hasT ? T12[12 * b + i] : 0.0;
  const float* dfrm = depth + b * HW;
  float* ofrm = xyz + b * HW * 3;
  uint8_t* vfrm = valid ? valid + b * HW : nullptr;
  int nvalid = 0;
  if (VEC4) {
    const unsigned nvec = (unsigned)(HW >> 2);
    // a block owns 512 consecutive float4s; thread t takes t and t + 256 (both loads in flight before any math)
    for (unsigned base = blockIdx.x * 512u; base < nvec; base += gridDim.x * 512u) {
      const unsigned i0 = base + threadIdx.x, i1 = i0 + 256u;
      const bool h0 = i0 < nvec, h1 = i1 < nvec;
      float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
      if (h0) d0 = __ldcs(reinterpret_cast<const float4*>(dfrm) + i0);
      if (h1) d1 = __ldcs(reinterpret_cast<const float4*>(dfrm) + i1);
      float o[12];
      uchar4 m;
      if (h0) {
        nvalid += backproject_quad(d0, i0 * 4u, W, wmagic, f, o, m);
        float4* op = reinterpret_cast<float4*>(ofrm) + 3ll * i0;
        __stcs(op, make_float4(o[0], o[1], o[2], o[3]));
        __stcs(op + 1, make_float4(o[4], o[5], o[6], o[7]));
        __stcs(op + 2, make_float4(o[8], o[9], o[10], o[11]));
        if (vfrm) __stcs(reinterpret_cast<uchar4*>(vfrm) + i0, m);
      }
      if (h1) {
        nvalid += backproject_quad(d1, i1 * 4u, W, wmagic, f, o, m);
        float4* op = reinterpret_cast<float4*>(ofrm) + 3ll * i1;
        __stcs(op, make_float4(o[0], o[1], o[2], o[3]));
        __stcs(op + 1, make_float4(o[4], o[5], o[6], o[7]));
        __stcs(op + 2, make_float4(o[8], o[9], o[10], o[11]));
        if (vfrm) __stcs(reinterpret_cast<uchar4*>(vfrm) + i1, m);
      }
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
      const int v = (int)(i / W), u = (int)(i - (long long)v * W);
      float X, Y, Z;
      const bool ok = backproject_one(dfrm[i], ((double)u - f.cx) * f.fx_inv, ((double)v - f.cy) * f.fy_inv, f, X, Y, Z);
      ofrm[3 * i] = X; ofrm[3 * i + 1] = Y; ofrm[3 * i + 2] = Z;
      if (vfrm) vfrm[i] = ok ? 1 : 0;
      nvalid += ok ? 1 : 0;
    }
  }
  if (counts) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    __shared__ int wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = nvalid;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += wsum[w];
      if (s) atomicAdd(counts + b, s);
    }
  }
}

int launch_backproject(const float* depth, int B, int H, int W, const double* K4, int k_per_frame, const double* T12,
                       float depth_scale, float depth_trunc, float* xyz, uint8_t* valid, int* counts,
                       cudaStream_t stream) {
  DAV2_CHECK(depth && xyz && K4 && B > 0 && H > 0 && W > 0, "backproject: null pointer or empty shape");
  DAV2_CHECK(depth_scale > 0.f, "backproject: depth_scale must be > 0");
  if (counts) DAV2_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int) * B, stream));
  const long long HW = (long long)H * W;
  DAV2_CHECK(HW < (1ll << 31), "backproject: frame larger than 2^31 pixels");
  const double inv_scale = 1.0 / (double)depth_scale;
  const double trunc = (double)depth_trunc;  // +inf disables truncation
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(depth) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) && ((reinterpret_cast<uintptr_t>(valid) & 3) == 0);
  // multiply-high division by W is exact for p < 2^32 / W  (p < HW); otherwise the kernel divides
  const unsigned wmagic = (W > 1 && HW * (long long)W < (1ll << 32)) ? (unsigned)((1ull << 32) / (unsigned)W + 1ull) : 0u;
  long long bx = vec ? (HW / 4 + 511) / 512 : (HW + 255) / 256;
  // (bx * B) beyond ~16 waves of the 148 SMs x 3 resident CTAs buys nothing; the block-stride loop covers the rest
  const long long cap = ((long long)sm_count() * 3 * 16 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  ProfScope ps(PC_BACKPROJECT, 0.0, (double)B * HW * (valid ? 17.0 : 16.0), stream);
  dim3 grid((unsigned)bx, (unsigned)B);
  if (vec)
    backproject_kernel<true><<<grid, 256, 0, stream>>>(depth, H, W, wmagic, K4, k_per_frame, T12, inv_scale, trunc, xyz, valid, counts);
  else
    backproject_kernel<false><<<grid, 256, 0, stream>>>(depth, H, W, wmagic, K4, k_per_frame, T12, inv_scale, trunc, xyz, valid, counts);
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// depth metrics: partial sums {n, S|d|, S|d|/(gt+1e-6), Sd^2, Sgt, #(t<a), #(t<b), #(t<c)}, t = max(gt/pred, pred/gt)
//   variant 0 (compute_errors / test_step): valid = lo <= gt <= hi ; thresholds 1.1 (b, c unused = 1.1)
//   variant 1 (calculate_metrics):          valid = gt>0 & pred>0 & !isinf(gt) & !isinf(pred); 1.25, 1.25^2, 1.25^3
//   variants 2 / 3: variants 0 / 1 without a mask (every element counts)
// ----------------------------------------------------------------------------------------------
// Element-wise values are fp32 exactly as the reference computes them (torch / numpy fp32 ops, IEEE division);
// they are summed in fp32 over 8 pixels and then folded into fp64 accumulators.  (Converting every term to fp64
// put 4 F2F + 3 RCP per pixel on the XU pipe: ncu showed XU 94 % and 2.3 TB/s for the first version.)
struct MetricAccF {
  float s_abs, s_rel, s_sq, s_gt;
  unsigned int n, na, nb, nc;
};
struct MetricAcc {
  double s_abs, s_rel, s_sq, s_gt;
  unsigned int n, na, nb, nc;
};

template <int VARIANT>
__device__ __forceinline__ void metric_accum(MetricAccF& a, float p, float g, float lo, float hi) {
  bool ok;
  if (VARIANT == 0)
    ok = (g >= lo) && (g <= hi);
  else if (VARIANT == 1)
    ok = (g > 0.f) && (p > 0.f) && !isinf(g) && !isinf(p);
  else
    ok = true;  // variants 2 / 3: every element counts (plain compute_errors / calculate_metrics(mask_invalid=False))
  constexpr bool CM = (VARIANT == 1 || VARIANT == 3);  // calculate_metrics definitions
  const float d = p - g;
  const float ad = fabsf(d);
  const float t = fmaxf(g / p, p / g);
  a.s_abs += ok ? ad : 0.f;
  if (!CM) a.s_rel += ok ? ad / (g + 1e-6f) : 0.f;  // calculate_metrics divides the means instead
  a.s_sq += ok ? d * d : 0.f;
  a.s_gt += ok ? g : 0.f;
  a.n += ok ? 1 : 0;
  if (!CM) {
    a.na += (ok && t < 1.1f) ? 1 : 0;
    a.nb += (ok && isnan(p)) ? 1 : 0;  // eval/evaluation.py:33-36 NaN / Inf warnings
    a.nc += (ok && isinf(p)) ? 1 : 0;
  } else {
    a.na += (ok && t < 1.25f) ? 1 : 0;
    a.nb += (ok && t < 1.5625f) ? 1 : 0;
    a.nc += (ok && t < 1.953125f) ? 1 : 0;
  }
}

template <int VARIANT>
__device__ __forceinline__ void metric_quad(MetricAccF& a, const float4& p4, const float4& g4, float lo, float hi) {
  metric_accum<VARIANT>(a, p4.x, g4.x, lo, hi);
  metric_accum<VARIANT>(a, p4.y, g4.y, lo, hi);
  metric_accum<VARIANT>(a, p4.z, g4.z, lo, hi);
  metric_accum<VARIANT>(a, p4.w, g4.w, lo, hi);
}

__device__ __forceinline__ void metric_fold(MetricAcc& a, const MetricAccF& c) {
  a.s_abs += (double)c.s_abs; a.s_rel += (double)c.s_rel; a.s_sq += (double)c.s_sq; a.s_gt += (double)c.s_gt;
  a.n += c.n; a.na += c.na; a.nb += c.nb; a.nc += c.nc;
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 4) depth_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                               long long HW, float lo, float hi, int per_frame,
                                                               double* __restrict__ partials) {
  const int b = blockIdx.y;
  const float* pf = pred + b * HW;
  const float* gf = gt + b * HW;
  MetricAcc a = {0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u, 0u};
  const bool vec = (HW % 4 == 0) && (((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0);
  if (vec) {
    const long long nvec = HW >> 2;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // a block owns 512 consecutive float4 pairs per trip; four 16-byte loads are in flight per thread
    for (long long base = (long long)blockIdx.x * 512; base < nvec; base += (long long)gridDim.x * 512) {
      const long long i0 = base + threadIdx.x, i1 = i0 + 256;
      const bool h0 = i0 < nvec, h1 = i1 < nvec;
      float4 p0 = zero, g0 = zero, p1 = zero, g1 = zero;
      if (h0) { p0 = __ldcs(reinterpret_cast<const float4*>(pf) + i0); g0 = __ldcs(reinterpret_cast<const float4*>(gf) + i0); }
      if (h1) { p1 = __ldcs(reinterpret_cast<const float4*>(pf) + i1); g1 = __ldcs(reinterpret_cast<const float4*>(gf) + i1); }
      MetricAccF c = {0.f, 0.f, 0.f, 0.f, 0u, 0u, 0u, 0u};
      if (h0) metric_quad<VARIANT>(c, p0, g0, lo, hi);
      if (h1) metric_quad<VARIANT>(c, p1, g1, lo, hi);
      metric_fold(a, c);
    }
  } else {
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i0 < HW; i0 += (long long)gridDim.x * blockDim.x * 8) {
      MetricAccF c = {0.f, 0.f, 0.f, 0.f, 0u, 0u, 0u, 0u};
      for (long long i = i0; i < i0 + 8 && i < HW; ++i) metric_accum<VARIANT>(c, pf[i], gf[i], lo, hi);
      metric_fold(a, c);
    }
  }
  double v[8] = {(double)a.n, a.s_abs, a.s_rel, a.s_sq, a.s_gt, (double)a.na, (double)a.nb, (double)a.nc};
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __shared__ double sm[8][8];
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 8; ++k) sm[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w][threadIdx.x];
    atomicAdd(partials + (per_frame ? 8 * b : 0) + threadIdx.x, s);
  }
}

int launch_depth_metrics(const float* pred, const float* gt, int B, long long HW, float lo, float hi, int variant,
                         int per_frame, double* partials, cudaStream_t stream) {
  DAV2_CHECK(partials && B > 0 && HW >= 0, "depth_metrics: null pointer or bad shape");
  if (HW == 0) {  // empty selection: all-zero partials (mean of nothing -> NaN after finalisation)
    DAV2_CUDA_OK(cudaMemsetAsync(partials, 0, sizeof(double) * 8 * (per_frame ? B : 1), stream));
    return 0;
  }
  DAV2_CHECK(pred && gt, "depth_metrics: null pointer");
  DAV2_CHECK(variant >= 0 && variant <= 3,
             "depth_metrics: variant must be 0 (test_step mask), 1 (calculate_metrics), 2 (compute_errors, no mask) or 3 "
             "(calculate_metrics, no mask)");
  DAV2_CUDA_OK(cudaMemsetAsync(partials, 0, sizeof(double) * 8 * (per_frame ? B : 1), stream));
  long long bx = (HW / 4 + 511) / 512;
  const long long cap = ((long long)sm_count() * 4 * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  ProfScope ps(PC_METRICS, 0.0, (double)B * HW * 8.0, stream);
  dim3 grid((unsigned)bx, (unsigned)B);
  if (variant == 0)
    depth_metrics_kernel<0><<<grid, 256, 0, stream>>>(pred, gt, HW, lo, hi, per_frame, partials);
  else if (variant == 1)
    depth_metrics_kernel<1><<<grid, 256, 0, stream>>>(pred, gt, HW, lo, hi, per_frame, partials);
  else if (variant == 2)
    depth_metrics_kernel<2><<<grid, 256, 0, stream>>>(pred, gt, HW, lo, hi, per_frame, partials);
  else
    depth_metrics_kernel<3><<<grid, 256, 0, stream>>>(pred, gt, HW, lo, hi, per_frame, partials);
  DAV2_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// pose chain: q_{i+1} = q_i (x) r_i ; t_{i+1} = t_i + rot(q_i, tau_i); fp32, sequential, un-fused
// (intrinsics keep nvcc from contracting mul+add into FMA so the rounding sequence matches eager torch).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ void cross_(const float* a, const float* b, float* c) {
  c[0] = sub_(mul_(a[1], b[2]), mul_(a[2], b[1]));
  c[1] = sub_(mul_(a[2], b[0]), mul_(a[0], b[2]));
  c[2] = sub_(mul_(a[0], b[1]), mul_(a[1], b[0]));
}

__device__ void pose_to_T12(const float* p, double* T) {
  double x = p[3], y = p[4], z = p[5], w = p[6];
  const double n = sqrt(x * x + y * y + z * z + w * w);
  x /= n; y /= n; z /= n; w /= n;
  T[0] = 1 - 2 * (y * y + z * z); T[1] = 2 * (x * y - z * w);     T[2] = 2 * (x * z + y * w);      T[3] = p[0];
  T[4] = 2 * (x * y + z * w);     T[5] = 1 - 2 * (x * x + z * z); T[6] = 2 * (y * z - x * w);      T[7] = p[1];
  T[8] = 2 * (x * z - y * w);     T[9] = 2 * (y * z + x * w);     T[10] = 1 - 2 * (x * x + y * y); T[11] = p[2];
}

__global__ void compose_poses_kernel(const float* __restrict__ rel, const float* __restrict__ init7, int N,
                                     float* __restrict__ abs7) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float cur[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
  if (init7)
    for (int i = 0; i < 7; ++i) cur[i] = init7[i];
  for (int i = 0; i < 7; ++i) abs7[i] = cur[i];
  for (int s = 0; s < N; ++s) {
    const float* r = rel + 7 * s;
    float rq[4] = {r[3], r[4], r[5], r[6]};
    const float nrm = sqrtf(add_(add_(add_(mul_(rq[0], rq[0]), mul_(rq[1], rq[1])), mul_(rq[2], rq[2])), mul_(rq[3], rq[3])));
    if (nrm < 1e-8f) { rq[0] = rq[1] = rq[2] = 0.f; rq[3] = 1.f; }
    const float x1 = cur[3], y1 = cur[4], z1 = cur[5], w1 = cur[6];
    const float x2 = rq[0], y2 = rq[1], z2 = rq[2], w2 = rq[3];
    const float w = sub_(sub_(sub_(mul_(w1, w2), mul_(x1, x2)), mul_(y1, y2)), mul_(z1, z2));
    const float x = sub_(add_(add_(mul_(w1, x2), mul_(x1, w2)), mul_(y1, z2)), mul_(z1, y2));
    const float y = add_(add_(sub_(mul_(w1, y2), mul_(x1, z2)), mul_(y1, w2)), mul_(z1, x2));
    const float z = add_(sub_(add_(mul_(w1, z2), mul_(x1, y2)), mul_(y1, x2)), mul_(z1, w2));
    const float qv[3] = {x1, y1, z1};
    const float tv[3] = {r[0], r[1], r[2]};
    float uv[3], uuv[3];
    cross_(qv, tv, uv);
    cross_(qv, uv, uuv);
    for (int k = 0; k < 3; ++k)
      cur[k] = add_(cur[k], add_(tv[k], mul_(2.f, add_(mul_(uv[k], w1), uuv[k]))));
    cur[3] = x; cur[4] = y; cur[5] = z; cur[6] = w;
    for (int i = 0; i < 7; ++i) abs7[7 * (s + 1) + i] = cur[i];
  }
}

__global__ void poses_to_T12_kernel(const float* __restrict__ abs7, int n, double* __restrict__ T12) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double T[12];
  pose_to_T12(abs7 + 7 * i, T);
  for (int k = 0; k < 12; ++k) T12[12 * i + k] = T[k];
}

int launch_compose_poses(const float* rel, const float* init7, int N, float* abs7, double* T12, cudaStream_t stream) {
  DAV2_CHECK(abs7 && N >= 0 && (rel || N == 0), "compose_poses: null pointer");
  compose_poses_kernel<<<1, 32, 0, stream>>>(rel, init7, N, abs7);
  DAV2_LAUNCH_OK();
  if (T12) {
    poses_to_T12_kernel<<<(N + 1 + 127) / 128, 128, 0, stream>>>(abs7, N + 1, T12);
    DAV2_LAUNCH_OK();
  }
  return 0;
}

}  // namespace dav2
