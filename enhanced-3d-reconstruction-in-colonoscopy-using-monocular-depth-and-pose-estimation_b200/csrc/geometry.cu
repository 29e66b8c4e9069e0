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
#include <stdlib.h>

#include "elementwise.cuh"

namespace dav2 {

// ----------------------------------------------------------------------------------------------
// back-projection: 8 pixels per thread (two float4 loads issued up front), 6x float4 + 2x uchar4 streaming
// stores.  fp64 math in registers; the fp64<->fp32 / int->fp64 conversions run on the XU pipe (16/clk/SM),
// which is what bounded the first version (ncu: XU 62 %, 34 % occupancy, 3.8 TB/s) -- so the pixel
// coordinates are converted once per float4 and advanced by fp64 adds, and (row, col) comes from a
// multiply-high instead of an integer division (another three XU ops).
// ----------------------------------------------------------------------------------------------
// Destinations of one launch: the local output, or -- for the fused back-projection + cloud all-gather -- the
// gather buffer of EVERY rank (peer-mapped over NVLink; the store is the collective, SURVEY.md 8e phase 2).
// All pointers are already offset to this rank's first frame.
static constexpr int BP_MAX_DST = 8;
struct BpDst {
  float* xyz[BP_MAX_DST];
  uint8_t* valid[BP_MAX_DST];
  int* counts[BP_MAX_DST];
  int n;
};

// Per-frame constants.  X_w = R (z [xf, yf, 1]) + t = z * ray(u, v) + t with the world-frame ray
// ray(u, v) = u*dx + v*dy + r0,  dx = R[:,0]/fx,  dy = R[:,1]/fy,  r0 = R[:,2] - cx*dx - cy*dy,
// so a pixel costs 3 DFMA for its ray (ray_k = ray_0 + k*dx inside a quad) and 3 DFMA for the point instead of the
// 20+ fp64 operations of the literal formula (the first versions were fp64-issue bound: 90 instructions / pixel,
// 55 % issue at 33 % occupancy, 3.0 TB/s).  Differences to the oracle's operation order are ~1e-16 relative, far below
// the single rounding to fp32 at the end.  Without a pose R = I, t = 0.
// Validity (z = d / depth_scale in fp64; 0 < z < depth_trunc; finite) is decided on the fp32 depth itself: z is monotone
// in d, so one lane finds the smallest float d_thr whose z reaches depth_trunc and every pixel tests 0 < d < d_thr
// (d_thr = +inf without truncation, which still rejects inf / NaN) -- two FSETP instead of an F2F, two DSETP and a class
// test per pixel.
struct BpFrame {
  double dx[3], dy[3], r0[3], wrap[3], t[3];  // wrap = dy - W*dx: step from (u, v) to (u - W, v + 1)
  double inv_scale;
  float d_thr;
};

__device__ __forceinline__ void bp_frame_setup(BpFrame& fs, int c, const double* K, const double* T12, int b, int W,
                                               double inv_scale, float d_thr) {
  // x = (u-cx)/fx * z: the oracle divides; 1/fx in fp64 then multiplies differs by <= 1 ulp(fp64)
  const double fx_inv = 1.0 / K[0], fy_inv = 1.0 / K[1], cx = K[2], cy = K[3];
  const bool hasT = T12 != nullptr;
  const double r0c = hasT ? T12[12 * b + 4 * c] : (c == 0 ? 1.0 : 0.0), r1c = hasT ? T12[12 * b + 4 * c + 1] : (c == 1 ? 1.0 : 0.0),
               r2c = hasT ? T12[12 * b + 4 * c + 2] : (c == 2 ? 1.0 : 0.0);
  fs.dx[c] = r0c * fx_inv;
  fs.dy[c] = r1c * fy_inv;
  fs.r0[c] = r2c - cx * (r0c * fx_inv) - cy * (r1c * fy_inv);
  fs.wrap[c] = r1c * fy_inv - (double)W * (r0c * fx_inv);
  fs.t[c] = hasT ? T12[12 * b + 4 * c + 3] : 0.0;
  if (c == 0) {
    fs.inv_scale = inv_scale;
    fs.d_thr = d_thr;
  }
}

// smallest float d whose z = (double)d * inv_scale reaches depth_trunc (host side: a function of two launch scalars)
static float depth_threshold(double inv_scale, double trunc) {
  if (!(trunc < (double)INFINITY)) return INFINITY;
  float thr = (float)(trunc / inv_scale);
  for (int i = 0; i < 4 && (double)thr * inv_scale < trunc; ++i) thr = nextafterf(thr, INFINITY);
  for (int i = 0; i < 4 && (double)nextafterf(thr, -INFINITY) * inv_scale >= trunc; ++i) thr = nextafterf(thr, -INFINITY);
  return thr;
}

// One pixel: returns an all-ones / all-zero validity mask; the outputs are AND-ed with it (a `cond ? v : 0` select made
// the compiler sink the three DFMA + F2F of a pixel under a branch: BSSY / BRA / BSYNC + zero-initialisation per pixel
// although nearly every pixel is valid).
__device__ __forceinline__ unsigned backproject_one(float d, const double (&ray)[3], const BpFrame& f, float& X, float& Y, float& Z) {
  const unsigned mk = ((d > 0.f) && (d < f.d_thr)) ? 0xffffffffu : 0u;  // NaN fails both tests; +-inf fails one of them
  const double z = (double)d * f.inv_scale;
  X = __uint_as_float(__float_as_uint((float)fma(z, ray[0], f.t[0])) & mk);
  Y = __uint_as_float(__float_as_uint((float)fma(z, ray[1], f.t[1])) & mk);
  Z = __uint_as_float(__float_as_uint((float)fma(z, ray[2], f.t[2])) & mk);
  return mk;
}

// 4 consecutive pixels starting at linear index p0 (row-major); m = validity bytes (uchar4 of 0 / 1), returns their count.
// CROSS = false: the caller knows that no quad of the warp straddles two image rows (7 of 8 warps at W = 518); otherwise
// pixels k >= W - u are on row v + 1 (at most one row change: W >= 4).  (An out-of-line CROSS path made the compiler keep
// o[] in local memory for every quad: 58 -> 93 us.)
template <bool CROSS>
__device__ __forceinline__ int backproject_quad(const float4 d4, unsigned p0, int W, unsigned wmagic, const BpFrame& f,
                                                float (&o)[12], unsigned& m) {
  // v = p0 / W: multiply-high by ceil(2^32 / W) is exact while p0 * W < 2^32 (the launcher sends larger frames to the
  // scalar kernel)
  const unsigned v = __umulhi(p0, wmagic);
  const int u = (int)(p0 - v * (unsigned)W);
  const int nrow = W - u;
  const double ud = (double)u, vd = (double)(int)v;
  double ra[3], rb[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ra[c] = fma(ud, f.dx[c], fma(vd, f.dy[c], f.r0[c]));
    if (CROSS) rb[c] = ra[c] + f.wrap[c];
  }
  const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
  unsigned mk[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double ray[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double base = (!CROSS || k < nrow) ? ra[c] : rb[c];
      ray[c] = k == 0 ? base : fma((double)k, f.dx[c], base);  // (fma(0, dx, base) is not folded: dx could be inf / NaN)
    }
    mk[k] = backproject_one(dd[k], ray, f, o[3 * k], o[3 * k + 1], o[3 * k + 2]);
  }
  m = (mk[0] & 0x1u) | (mk[1] & 0x100u) | (mk[2] & 0x10000u) | (mk[3] & 0x1000000u);
  return __popc(m);
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
__device__ __forceinline__ bool metric_valid(float p, float g, float lo, float hi) {
  if (VARIANT == 0) return (g >= lo) && (g <= hi);
  if (VARIANT == 1) return (g > 0.f) && (p > 0.f) && !isinf(g) && !isinf(p);
  return true;  // variants 2 / 3: every element counts (plain compute_errors / calculate_metrics(mask_invalid=False))
}

// EXACT = false: pred and gt of every counted pixel are finite and > 0, so t = max(g/p, p/g) < thr is evaluated as
// (g < thr*p) && (p < thr*g) (equal up to the rounding of one quotient) and pred is neither NaN nor inf.
// EXACT = true: the reference's literal formula (two IEEE divisions) for zero / negative / non-finite values.
// (Two divisions + NaN bookkeeping on every pixel made the first version issue-bound: 80 % issue, 2.3 TB/s.)
template <int VARIANT, bool EXACT>
__device__ __forceinline__ void metric_accum(MetricAccF& a, float p, float g, float lo, float hi) {
  constexpr bool CM = (VARIANT == 1 || VARIANT == 3);  // calculate_metrics definitions
  constexpr float A = CM ? 1.25f : 1.1f, B2 = 1.5625f, C3 = 1.953125f;
  const bool ok = metric_valid<VARIANT>(p, g, lo, hi);
  if (!ok) return;  // short body: compiled to predicated instructions, no divergence penalty beyond the predicate
  const float d = p - g;
  const float ad = fabsf(d);
  a.s_abs += ad;
  if (!CM) {  // calculate_metrics divides the means instead
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(g + 1e-6f));  // <= 1 ulp; g + 1e-6 is a normal number here
    a.s_rel = EXACT ? a.s_rel + ad / (g + 1e-6f) : fmaf(ad, r, a.s_rel);
  }
  a.s_sq = fmaf(d, d, a.s_sq);
  a.s_gt += g;
  a.n += 1;
  bool ta, tb = false, tc = false;
  if (!EXACT) {
    ta = (g < A * p) && (p < A * g);
    if (CM) {
      tb = (g < B2 * p) && (p < B2 * g);
      tc = (g < C3 * p) && (p < C3 * g);
    }
  } else {
    const float q0 = g / p, q1 = p / g;
    const float t = fmaxf(q0, q1);
    const bool tn = isnan(q0) || isnan(q1);  // torch.max / np.maximum propagate NaN, and NaN < thr is false
    ta = !tn && t < A;
    tb = !tn && t < B2;
    tc = !tn && t < C3;
  }
  a.na += ta ? 1 : 0;
  if (CM) {
    a.nb += tb ? 1 : 0;
    a.nc += tc ? 1 : 0;
  } else if (EXACT) {
    a.nb += isnan(p) ? 1 : 0;  // eval/evaluation.py:33-36 NaN / Inf warnings
    a.nc += isinf(p) ? 1 : 0;
  }
}

// Does any counted pixel of the quad need the literal formula?  Branch-free: bits(x) - 1 (unsigned) is below
// 0x7f7fffff exactly for finite x > 0, so one running unsigned max over the counted operands decides.
template <int VARIANT>
__device__ __forceinline__ unsigned metric_exact_key(const float4& p4, const float4& g4, float lo, float hi) {
  if (VARIANT == 1) return 0u;  // the mask itself guarantees finite positive operands
  const float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
  unsigned key = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned kk = max(__float_as_uint(pp[k]) - 1u, __float_as_uint(gg[k]) - 1u);
    key = max(key, (VARIANT == 0 && !metric_valid<0>(pp[k], gg[k], lo, hi)) ? 0u : kk);
  }
  return key;
}

template <int VARIANT, bool EXACT>
__device__ __forceinline__ void metric_quad(MetricAccF& a, const float4& p4, const float4& g4, float lo, float hi) {
  metric_accum<VARIANT, EXACT>(a, p4.x, g4.x, lo, hi);
  metric_accum<VARIANT, EXACT>(a, p4.y, g4.y, lo, hi);
  metric_accum<VARIANT, EXACT>(a, p4.z, g4.z, lo, hi);
  metric_accum<VARIANT, EXACT>(a, p4.w, g4.w, lo, hi);
}

__device__ __forceinline__ void metric_fold(MetricAcc& a, const MetricAccF& c) {
  a.s_abs += (double)c.s_abs; a.s_rel += (double)c.s_rel; a.s_sq += (double)c.s_sq; a.s_gt += (double)c.s_gt;
  a.n += c.n; a.na += c.na; a.nb += c.nb; a.nc += c.nc;
}

int launch_depth_metrics(const float* pred, const float* gt, int B, long long HW, float lo, float hi, int variant,
                         int per_frame, double* partials, cudaStream_t stream);

// ----------------------------------------------------------------------------------------------
// back-projection kernels
// ----------------------------------------------------------------------------------------------
// block-wide sum of per-thread metric accumulators -> 8 fp64 atomics.
// fp64 inputs (the stand-alone metric kernel: a thread has already folded several trips): butterfly in fp64.
__device__ __forceinline__ void metric_block_reduce(const MetricAcc& a, double* __restrict__ dst) {
  // warp reduce: one REDUX per integer count, butterfly shuffles for the four fp64 sums
  double v[8] = {(double)__reduce_add_sync(0xffffffffu, a.n), a.s_abs, a.s_rel, a.s_sq, a.s_gt,
                 (double)__reduce_add_sync(0xffffffffu, a.na), (double)__reduce_add_sync(0xffffffffu, a.nb),
                 (double)__reduce_add_sync(0xffffffffu, a.nc)};
#pragma unroll
  for (int k = 1; k < 5; ++k)
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
    atomicAdd(dst + threadIdx.x, s);
  }
}
// fp32 inputs (the fused kernel: 4 * QPT pixels per thread): the warp's 512 pixels are summed with fp32 shuffles (20 SHFL
// + 20 FADD per thread instead of 40 + 20 DADD -- the fp64 butterfly was a fifth of the fused kernel's instructions), the
// eight warp sums and everything after them stay fp64.  512 terms in fp32: <= 2e-6 relative on a warp sum, random in sign.
__device__ __forceinline__ void metric_block_reduce(const MetricAccF& a, double* __restrict__ dst) {
  float v[4] = {a.s_abs, a.s_rel, a.s_sq, a.s_gt};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  const unsigned n = __reduce_add_sync(0xffffffffu, a.n), na = __reduce_add_sync(0xffffffffu, a.na),
                 nb = __reduce_add_sync(0xffffffffu, a.nb), nc = __reduce_add_sync(0xffffffffu, a.nc);
  __shared__ double sm[8][8];
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sm[warp][0] = (double)n; sm[warp][1] = (double)v[0]; sm[warp][2] = (double)v[1]; sm[warp][3] = (double)v[2];
    sm[warp][4] = (double)v[3]; sm[warp][5] = (double)na; sm[warp][6] = (double)nb; sm[warp][7] = (double)nc;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w][threadIdx.x];
    atomicAdd(dst + threadIdx.x, s);
  }
}

// Vector path: a block owns QPT * 256 consecutive float4s of ONE frame; thread t takes quads t, t + 256, ...: all loads
// (depth, and gt when the metric sums ride along) are in flight before any math, and the per-block costs (frame constants,
// count / metric reductions, atomics) are spread over 4 * QPT pixels per thread.  (A persistent variant -- one resident
// wave, each block walking ~20 tiles of its frame with the next tile's loads issued before the current tile's math -- was
// built and measured in round 2: 62 / 74 us instead of 53 / 72, and the same 97 us inside the step: the kernel is bound by
// the core clock it gets, not by exposed load latency.)
// Each thread's 4 points are 48 contiguous bytes; storing them directly makes every warp-wide store touch 12 lines with
// 16 of every 48 bytes (ncu: 1.94x the ideal number of L2 store sectors), so the warp stages its 1536 bytes in shared
// memory (48-byte thread stride = conflict-free for 128-bit accesses) and writes three fully coalesced 512-byte rows.
// MULTI = false: one destination (dst.*[0]); a destination loop over a runtime count kept eight sets of predicated 64-bit
// address arithmetic alive in the single-destination kernel as well.
// METRICS = true (round 2): the test_step metric partial sums (variant 0: lo <= gt <= hi, eval/evaluation.py:16-60) are
// accumulated from the SAME depth registers -- the separate metric kernel re-read the 68.7 MB depth map the head conv had
// just written (8 B/px of its own traffic); fused, one pass moves 21 B/px instead of 17 + 8.
template <bool MULTI, bool METRICS, int BP_QPT>
__global__ void __launch_bounds__(256, BP_QPT == 2 ? 4 : 3) backproject_vec_kernel(const float* __restrict__ depth, const float* __restrict__ gt,
                                                                 int H, int W, unsigned wmagic, const double* __restrict__ K4,
                                                                 int k_per_frame, const double* __restrict__ T12,
                                                                 double inv_scale, float d_thr, const BpDst dst, float lo,
                                                                 float hi, int per_frame, double* __restrict__ partials) {
  const int b = blockIdx.y;
  const long long HW = (long long)H * W;
  // the frame constants (two fp64 divisions, ~150 instructions) are computed by three lanes and broadcast through
  // shared memory; per thread that prologue cost a third of the 8-pixel body
  __shared__ BpFrame fs;
  __shared__ float4 stage[8][96];
  if (threadIdx.x < 3) bp_frame_setup(fs, threadIdx.x, K4 + (k_per_frame ? 4 * b : 0), T12, b, W, inv_scale, d_thr);
  const float4* dfrm = reinterpret_cast<const float4*>(depth + b * HW);
  const float4* gfrm = METRICS ? reinterpret_cast<const float4*>(gt + b * HW) : nullptr;
  const long long ooff = b * HW * 3, voff = b * HW;
  const bool has_valid = dst.valid[0] != nullptr;
  const int ndst = MULTI ? dst.n : 1;
  const unsigned nvec = (unsigned)(HW >> 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned i0 = blockIdx.x * (256u * BP_QPT) + threadIdx.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dq[BP_QPT], gq[BP_QPT];
#pragma unroll
  for (int q = 0; q < BP_QPT; ++q) {
    const unsigned iv = i0 + 256u * q;
    dq[q] = iv < nvec ? __ldcs(dfrm + iv) : zero;
  }
  if (METRICS) {
#pragma unroll
    for (int q = 0; q < BP_QPT; ++q) {
      const unsigned iv = i0 + 256u * q;
      gq[q] = iv < nvec ? __ldcs(gfrm + iv) : zero;
    }
  }
  __syncthreads();  // frame constants; the loads above are already in flight
  int nvalid = 0;
#pragma unroll
  for (int q = 0; q < BP_QPT; ++q) {
    const unsigned iv = i0 + 256u * q;
    const unsigned wbase = iv - lane;  // first float4 index of this warp's 32
    if (wbase >= nvec) continue;       // warp-uniform
    float o[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    unsigned m = 0u;
    // does any quad of this warp straddle two rows?  first pixel of the warp's span: wbase*4, 128 pixels long
    const unsigned pw = wbase * 4u;
    const bool cross = (pw - __umulhi(pw, wmagic) * (unsigned)W) + 128u > (unsigned)W;  // warp-uniform
    if (iv < nvec) {
      if (cross) nvalid += backproject_quad<true>(dq[q], iv * 4u, W, wmagic, fs, o, m);
      else nvalid += backproject_quad<false>(dq[q], iv * 4u, W, wmagic, fs, o, m);
    }
    stage[warp][3 * lane] = make_float4(o[0], o[1], o[2], o[3]);
    stage[warp][3 * lane + 1] = make_float4(o[4], o[5], o[6], o[7]);
    stage[warp][3 * lane + 2] = make_float4(o[8], o[9], o[10], o[11]);
    __syncwarp();
    const float4 r0 = stage[warp][lane], r1 = stage[warp][lane + 32], r2 = stage[warp][lane + 64];
    __syncwarp();
    const unsigned nout = 3u * min(32u, nvec - wbase);  // float4s of this warp inside the frame
    for (int p = 0; p < ndst; ++p) {
      float4* op = reinterpret_cast<float4*>(dst.xyz[p] + ooff) + 3ll * wbase;
      if ((unsigned)lane < nout) __stcs(op + lane, r0);
      if ((unsigned)lane + 32u < nout) __stcs(op + lane + 32, r1);
      if ((unsigned)lane + 64u < nout) __stcs(op + lane + 64, r2);
      if (has_valid && iv < nvec) __stcs(reinterpret_cast<unsigned*>(dst.valid[p] + voff) + iv, m);
    }
  }
  if (dst.counts[0]) {
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    __shared__ int wsum[8];
    if (lane == 0) wsum[warp] = nvalid;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < 8; ++w) s += wsum[w];
      if (s)
        for (int p = 0; p < ndst; ++p) atomicAdd(dst.counts[p] + b, s);  // peer destinations: NVLink atomics
    }
  }
  if (METRICS) {
    MetricAccF c = {0.f, 0.f, 0.f, 0.f, 0u, 0u, 0u, 0u};
    // absent quads (beyond the frame) are skipped explicitly; does any counted pixel need the literal formula?
    // With 0 < lo and hi < inf (launch-uniform; the test_step mask) every COUNTED gt is finite and positive by the mask
    // itself, so only the predictions decide -- and testing uncounted ones too is harmless (1.5 instead of 6 instructions
    // per pixel; absent quads are zero-filled: key 0xffffffff, so they are skipped here as well).
    const bool gt_safe = lo > 0.f && hi < INFINITY;
    unsigned key = 0u;
#pragma unroll
    for (int q = 0; q < BP_QPT; ++q) {
      if (i0 + 256u * q >= nvec) continue;
      if (gt_safe)
        key = max(key, max(max(__float_as_uint(dq[q].x) - 1u, __float_as_uint(dq[q].y) - 1u),
                           max(__float_as_uint(dq[q].z) - 1u, __float_as_uint(dq[q].w) - 1u)));
      else
        key = max(key, metric_exact_key<0>(dq[q], gq[q], lo, hi));
    }
    if (key < 0x7f7fffffu) {
#pragma unroll
      for (int q = 0; q < BP_QPT; ++q)
        if (i0 + 256u * q < nvec) metric_quad<0, false>(c, dq[q], gq[q], lo, hi);
    } else {
#pragma unroll
      for (int q = 0; q < BP_QPT; ++q)
        if (i0 + 256u * q < nvec) metric_quad<0, true>(c, dq[q], gq[q], lo, hi);
    }
    metric_block_reduce(c, partials + (per_frame ? 8 * b : 0));
  }
}

// Scalar path (frames whose pixel count is not a multiple of 4, narrower than 4 pixels, or unaligned buffers)
__global__ void __launch_bounds__(256) backproject_scalar_kernel(const float* __restrict__ depth, int H, int W,
                                                                 const double* __restrict__ K4, int k_per_frame,
                                                                 const double* __restrict__ T12, double inv_scale, float d_thr,
                                                                 const BpDst dst) {
  const int b = blockIdx.y;
  const long long HW = (long long)H * W;
  __shared__ BpFrame fs;
  if (threadIdx.x < 3) bp_frame_setup(fs, threadIdx.x, K4 + (k_per_frame ? 4 * b : 0), T12, b, W, inv_scale, d_thr);
  __syncthreads();
  const BpFrame f = fs;
  const float* dfrm = depth + b * HW;
  const long long ooff = b * HW * 3, voff = b * HW;
  const bool has_valid = dst.valid[0] != nullptr;
  int nvalid = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / W), u = (int)(i - (long long)v * W);
    float X, Y, Z;
    double ray[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ray[c] = fma((double)u, f.dx[c], fma((double)v, f.dy[c], f.r0[c]));
    const bool ok = backproject_one(dfrm[i], ray, f, X, Y, Z) != 0u;
    for (int p = 0; p < dst.n; ++p) {
      float* ofrm = dst.xyz[p] + ooff;
      ofrm[3 * i] = X; ofrm[3 * i + 1] = Y; ofrm[3 * i + 2] = Z;
      if (has_valid) dst.valid[p][voff + i] = ok ? 1 : 0;
    }
    nvalid += ok ? 1 : 0;
  }
  if (dst.counts[0]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    __shared__ int wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = nvalid;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += wsum[w];
      if (s)
        for (int p = 0; p < dst.n; ++p) atomicAdd(dst.counts[p] + b, s);
    }
  }
}

// xyz / valid / counts: n_dst destination pointers each (valid, counts: all NULL or all set), already offset to the
// first frame this launch writes.  gt / partials non-NULL: the variant-0 metric partial sums of (depth, gt) are produced
// by the same pass (partials is zeroed here; fp64 [B,8] if per_frame else [8]).
int launch_backproject_multi(const float* depth, int B, int H, int W, const double* K4, int k_per_frame, const double* T12,
                             float depth_scale, float depth_trunc, float* const* xyz, uint8_t* const* valid,
                             int* const* counts, int n_dst, cudaStream_t stream, const float* gt, float lo, float hi,
                             int per_frame, double* partials) {
  DAV2_CHECK(depth && xyz && K4 && B > 0 && H > 0 && W > 0, "backproject: null pointer or empty shape");
  DAV2_CHECK(n_dst >= 1 && n_dst <= BP_MAX_DST, "backproject: 1..%d destinations", BP_MAX_DST);
  DAV2_CHECK(depth_scale > 0.f, "backproject: depth_scale must be > 0");
  DAV2_CHECK((gt == nullptr) == (partials == nullptr), "backproject: gt and partials go together");
  const long long HW = (long long)H * W;
  DAV2_CHECK(HW < (1ll << 31), "backproject: frame larger than 2^31 pixels");
  BpDst dst;
  // the 4-pixel path assumes at most ONE row change inside a quad (backproject_quad): frames narrower than 4 pixels take
  // the scalar path
  // multiply-high division by W (vector path) is exact for p < 2^32 / W, p < HW
  const unsigned wmagic = (W > 1 && HW * (long long)W < (1ll << 32)) ? (unsigned)((1ull << 32) / (unsigned)W + 1ull) : 0u;
  bool vec = (HW % 4 == 0) && (W >= 4) && wmagic != 0u && ((reinterpret_cast<uintptr_t>(depth) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(gt) & 15) == 0);
  for (int p = 0; p < BP_MAX_DST; ++p) {
    const bool on = p < n_dst;
    dst.xyz[p] = on ? xyz[p] : nullptr;
    dst.valid[p] = (on && valid) ? valid[p] : nullptr;
    dst.counts[p] = (on && counts) ? counts[p] : nullptr;
    if (on) {
      DAV2_CHECK(dst.xyz[p] && (!valid || dst.valid[p]) && (!counts || dst.counts[p]), "backproject: null destination %d", p);
      vec = vec && ((reinterpret_cast<uintptr_t>(dst.xyz[p]) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst.valid[p]) & 3) == 0);
      // zero this launch's slice of every destination's count vector (peer pointers are mapped: stream-ordered memset)
      if (dst.counts[p]) DAV2_CUDA_OK(cudaMemsetAsync(dst.counts[p], 0, sizeof(int) * B, stream));
    }
  }
  dst.n = n_dst;
  const double inv_scale = 1.0 / (double)depth_scale;
  const float d_thr = depth_threshold(inv_scale, (double)depth_trunc);  // +inf disables truncation
  if (gt && !vec) {
    // unaligned / odd-sized frames: the two passes run separately (same results)
    if (int rc = launch_depth_metrics(depth, gt, B, HW, lo, hi, 0, per_frame, partials, stream)) return rc;
    gt = nullptr;
    partials = nullptr;
  }
  if (gt) DAV2_CUDA_OK(cudaMemsetAsync(partials, 0, sizeof(double) * 8 * (per_frame ? B : 1), stream));
  // one trip per block: a block-stride loop with a fractional number of passes leaves most SMs idle during the last one
  int qpt = 4;  // quads per thread: 72 us (4) vs 74 us (2) for the fused pass at 64 x 518^2
#ifdef DAV2_PROFILING_KNOBS
  if (const char* e = getenv("DAV2_BP_QPT")) qpt = atoi(e) == 2 ? 2 : 4;
#endif
  long long bx = vec ? (HW / 4 + 256 * qpt - 1) / (256 * qpt) : (HW + 255) / 256;
  if (bx > 1048576) bx = 1048576;
  if (bx < 1) bx = 1;
  ProfScope ps(PC_BACKPROJECT, 0.0, (double)B * HW * ((gt ? 8.0 : 4.0) + n_dst * (valid ? 13.0 : 12.0)), stream);
  dim3 grid((unsigned)bx, (unsigned)B);
#define DAV2_BP_GO(MULTI, MET)                                                                                          \
  do {                                                                                                                 \
    if (qpt == 4)                                                                                                      \
      backproject_vec_kernel<MULTI, MET, 4><<<grid, 256, 0, stream>>>(depth, gt, H, W, wmagic, K4, k_per_frame, T12,    \
                                                                     inv_scale, d_thr, dst, lo, hi, per_frame, partials); \
    else                                                                                                               \
      backproject_vec_kernel<MULTI, MET, 2><<<grid, 256, 0, stream>>>(depth, gt, H, W, wmagic, K4, k_per_frame, T12,    \
                                                                     inv_scale, d_thr, dst, lo, hi, per_frame, partials); \
  } while (0)
  if (vec && n_dst == 1 && gt) DAV2_BP_GO(false, true);
  else if (vec && n_dst == 1) DAV2_BP_GO(false, false);
  else if (vec && gt) DAV2_BP_GO(true, true);
  else if (vec) DAV2_BP_GO(true, false);
  else backproject_scalar_kernel<<<grid, 256, 0, stream>>>(depth, H, W, K4, k_per_frame, T12, inv_scale, d_thr, dst);
#undef DAV2_BP_GO
  DAV2_LAUNCH_OK();
  return 0;
}

int launch_backproject(const float* depth, int B, int H, int W, const double* K4, int k_per_frame, const double* T12,
                       float depth_scale, float depth_trunc, float* xyz, uint8_t* valid, int* counts,
                       cudaStream_t stream) {
  return launch_backproject_multi(depth, B, H, W, K4, k_per_frame, T12, depth_scale, depth_trunc, &xyz, valid ? &valid : nullptr,
                                  counts ? &counts : nullptr, 1, stream, nullptr, 0.f, 0.f, 0, nullptr);
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
    // a block owns 1024 consecutive float4 pairs per pass and walks them in two trips of 512 (four 16-byte loads in
    // flight per thread)
    for (long long trip = (long long)blockIdx.x * 2; trip * 512 < nvec; trip = (trip & 1) ? trip + 2 * (long long)gridDim.x - 1 : trip + 1) {
      const long long i0 = trip * 512 + threadIdx.x, i1 = i0 + 256;
      const bool h0 = i0 < nvec, h1 = i1 < nvec;
      float4 p0 = zero, g0 = zero, p1 = zero, g1 = zero;
      if (h0) { p0 = __ldcs(reinterpret_cast<const float4*>(pf) + i0); g0 = __ldcs(reinterpret_cast<const float4*>(gf) + i0); }
      if (h1) { p1 = __ldcs(reinterpret_cast<const float4*>(pf) + i1); g1 = __ldcs(reinterpret_cast<const float4*>(gf) + i1); }
      MetricAccF c = {0.f, 0.f, 0.f, 0.f, 0u, 0u, 0u, 0u};
      // zero-filled (absent) quads are invalid under variants 0 / 1; variants 2 / 3 skip them explicitly
      const bool exact = max(h0 ? metric_exact_key<VARIANT>(p0, g0, lo, hi) : 0u, h1 ? metric_exact_key<VARIANT>(p1, g1, lo, hi) : 0u) >= 0x7f7fffffu;
      if (!exact) {
        if (h0) metric_quad<VARIANT, false>(c, p0, g0, lo, hi);
        if (h1) metric_quad<VARIANT, false>(c, p1, g1, lo, hi);
      } else {
        if (h0) metric_quad<VARIANT, true>(c, p0, g0, lo, hi);
        if (h1) metric_quad<VARIANT, true>(c, p1, g1, lo, hi);
      }
      metric_fold(a, c);
    }
  } else {
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i0 < HW; i0 += (long long)gridDim.x * blockDim.x * 8) {
      MetricAccF c = {0.f, 0.f, 0.f, 0.f, 0u, 0u, 0u, 0u};
      for (long long i = i0; i < i0 + 8 && i < HW; ++i) metric_accum<VARIANT, true>(c, pf[i], gf[i], lo, hi);
      metric_fold(a, c);
    }
  }
  metric_block_reduce(a, partials + (per_frame ? 8 * b : 0));
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
  // every block makes (up to) two trips of 512 float4 pairs: the warp/block reduction and the 8 atomics are a fixed cost
  // per block, and a whole number of trips avoids a mostly idle last pass
  long long bx = (HW / 4 + 1023) / 1024;
  const long long cap = 1048576;
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
// rigid transform of an existing cloud: p <- R p + t in fp64, one rounding to fp32 (o3d PointCloud.transform,
// depth_to_pointcloud.py:236-239, for clouds that are already on the device)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transform_points_kernel(float* __restrict__ xyz, long long n, const double* __restrict__ T12) {
  __shared__ double T[12];
  if (threadIdx.x < 12) T[threadIdx.x] = T12[threadIdx.x];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    xyz[3 * i] = (float)fma(T[0], x, fma(T[1], y, fma(T[2], z, T[3])));
    xyz[3 * i + 1] = (float)fma(T[4], x, fma(T[5], y, fma(T[6], z, T[7])));
    xyz[3 * i + 2] = (float)fma(T[8], x, fma(T[9], y, fma(T[10], z, T[11])));
  }
}

int launch_transform_points(float* xyz, long long n, const double* T12, cudaStream_t stream) {
  DAV2_CHECK(xyz && T12 && n >= 0, "transform_points: null pointer");
  if (n == 0) return 0;
  long long bx = (n + 255) / 256;
  if (bx > 148 * 16) bx = 148 * 16;
  ProfScope ps(PC_OTHER, 0.0, (double)n * 24.0, stream);
  transform_points_kernel<<<(unsigned)bx, 256, 0, stream>>>(xyz, n, T12);
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
