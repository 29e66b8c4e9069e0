// Voxel-grid down-sampling of a fused point cloud (reference depth_to_pointcloud.py:357-359 ->
// Open3D PointCloud::VoxelDownSample): voxel index = floor((p - (min_bound - voxel/2)) / voxel) per axis,
// output = per-voxel mean of the points (and colours), accumulated in double like Open3D.
// GPU formulation: min/max-bound reduction -> 63-bit voxel keys -> radix sort (CUB) of (key, point index) ->
// reduce-by-key of fp64 sums (values gathered on the fly through the sorted index) -> means.
// Output order is ascending (ix, iy, iz); Open3D's is hash-map order, so results compare as sets.
#include <cub/cub.cuh>

#include <vector>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "elementwise.cuh"

namespace dav2 {

struct VoxAcc {
  double x, y, z, r, g, b;
  long long n;
};
struct VoxAdd {
  __host__ __device__ VoxAcc operator()(const VoxAcc& a, const VoxAcc& b) const {
    VoxAcc o;
    o.x = a.x + b.x; o.y = a.y + b.y; o.z = a.z + b.z; o.r = a.r + b.r; o.g = a.g + b.g; o.b = a.b + b.b; o.n = a.n + b.n;
    return o;
  }
};
struct Bounds {
  float lo[3], hi[3];
};
struct BoundsMerge {
  __host__ __device__ Bounds operator()(const Bounds& a, const Bounds& b) const {
    Bounds o;
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.lo[k] = fminf(a.lo[k], b.lo[k]); o.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return o;
  }
};
// point i -> its (degenerate) bounding box; masked-out / non-finite points give the empty box
struct PointBounds {
  const float* xyz;
  const uint8_t* valid;
  __host__ __device__ Bounds operator()(long long i) const {
    Bounds o;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const bool ok = (!valid || valid[i]) && isfinite(x) && isfinite(y) && isfinite(z);
    o.lo[0] = ok ? x : INFINITY; o.lo[1] = ok ? y : INFINITY; o.lo[2] = ok ? z : INFINITY;
    o.hi[0] = ok ? x : -INFINITY; o.hi[1] = ok ? y : -INFINITY; o.hi[2] = ok ? z : -INFINITY;
    return o;
  }
};
// sorted position -> accumulator of the point it refers to
struct GatherAcc {
  const float* xyz;
  const float* rgb;
  __host__ __device__ VoxAcc operator()(unsigned int j) const {
    VoxAcc a;
    const long long o = 3ll * j;
    a.x = xyz[o]; a.y = xyz[o + 1]; a.z = xyz[o + 2];
    a.r = rgb ? rgb[o] : 0.0; a.g = rgb ? rgb[o + 1] : 0.0; a.b = rgb ? rgb[o + 2] : 0.0;
    a.n = 1;
    return a;
  }
};

static constexpr long long kAxisMax = (1ll << 21) - 1;  // 21 bits per axis in the 63-bit key

__global__ void __launch_bounds__(256) vox_keys(const float* __restrict__ xyz, const uint8_t* __restrict__ valid, long long n,
                                                const Bounds* __restrict__ bounds, double voxel,
                                                unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx,
                                                int* __restrict__ overflow) {
  const double bx = (double)bounds->lo[0] - 0.5 * voxel, by = (double)bounds->lo[1] - 0.5 * voxel,
               bz = (double)bounds->lo[2] - 0.5 * voxel;
  if (blockIdx.x == 0 && threadIdx.x == 0 && bounds->lo[0] <= bounds->hi[0]) {
    // Open3D raises "voxel_size is too small" when the grid does not fit an int; ours must fit 21 bits per axis
    const double ex = fmax(fmax((double)bounds->hi[0] - bx, (double)bounds->hi[1] - by), (double)bounds->hi[2] - bz);
    if (floor(ex / voxel) > (double)kAxisMax) *overflow = 1;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const bool ok = (!valid || valid[i]) && isfinite(x) && isfinite(y) && isfinite(z);
    unsigned long long k = ~0ull;  // dropped points sort to the end
    if (ok) {
      const long long ix = (long long)floor(((double)x - bx) / voxel), iy = (long long)floor(((double)y - by) / voxel),
                      iz = (long long)floor(((double)z - bz) / voxel);
      k = ((unsigned long long)(ix & kAxisMax) << 42) | ((unsigned long long)(iy & kAxisMax) << 21) | (unsigned long long)(iz & kAxisMax);
    }
    keys[i] = k;
    idx[i] = (unsigned int)i;
  }
}

__global__ void __launch_bounds__(256) vox_finalize(const unsigned long long* __restrict__ ukeys, const VoxAcc* __restrict__ sums,
                                                    const long long* __restrict__ nruns, const int* __restrict__ overflow,
                                                    float* __restrict__ oxyz, float* __restrict__ orgb,
                                                    long long* __restrict__ out_count) {
  const long long runs = *nruns;
  // the last run is the dropped-point bucket (key ~0) if any point was dropped
  const long long nv = (runs > 0 && ukeys[runs - 1] == ~0ull) ? runs - 1 : runs;
  if (*overflow) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = -1;
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const VoxAcc a = sums[i];
    const double dn = (double)a.n;
    oxyz[3 * i] = (float)(a.x / dn); oxyz[3 * i + 1] = (float)(a.y / dn); oxyz[3 * i + 2] = (float)(a.z / dn);
    if (orgb) { orgb[3 * i] = (float)(a.r / dn); orgb[3 * i + 1] = (float)(a.g / dn); orgb[3 * i + 2] = (float)(a.b / dn); }
  }
}

static inline int vgrid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

int launch_voxel_downsample(const float* xyz, const float* rgb, const uint8_t* valid, long long n, double voxel, float* out_xyz,
                            float* out_rgb, long long* out_count, cudaStream_t stream) {
  DAV2_CHECK(out_count && n >= 0 && voxel > 0.0 && (n == 0 || (xyz && out_xyz)), "voxel_downsample: bad arguments");
  DAV2_CHECK(n < (1ll << 31), "voxel_downsample: at most 2^31-1 points per call");
  DAV2_CHECK(!out_rgb || rgb, "voxel_downsample: out_rgb without rgb");
  if (n == 0) {
    DAV2_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(long long), stream));
    return 0;
  }
  using CountIt = thrust::counting_iterator<long long>;
  using BoundsIt = thrust::transform_iterator<PointBounds, CountIt>;
  using GatherIt = thrust::transform_iterator<GatherAcc, const unsigned int*>;
  Bounds empty;
  for (int k = 0; k < 3; ++k) { empty.lo[k] = INFINITY; empty.hi[k] = -INFINITY; }

  // stream-ordered scratch, released on every exit path
  struct Scratch {
    cudaStream_t s;
    std::vector<void*> p;
    ~Scratch() {
      for (void* q : p) cudaFreeAsync(q, s);
    }
    int get(void** out, size_t bytes) {
      DAV2_CUDA_OK(cudaMallocAsync(out, bytes, s));
      p.push_back(*out);
      return 0;
    }
  } scratch{stream, {}};
#define VOX_ALLOC(ptr, bytes)                                                    \
  do {                                                                           \
    if (int rc = scratch.get(reinterpret_cast<void**>(&(ptr)), (bytes))) return rc; \
  } while (0)
  Bounds* bounds = nullptr;
  unsigned long long *keys = nullptr, *keys_s = nullptr;  // keys is reused for the unique keys after the sort
  unsigned int *idx = nullptr, *idx_s = nullptr;
  VoxAcc* sums = nullptr;
  long long* nruns = nullptr;
  int* overflow = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, need = 0;
  VOX_ALLOC(bounds, sizeof(Bounds));
  VOX_ALLOC(keys, n * 8);
  VOX_ALLOC(keys_s, n * 8);
  VOX_ALLOC(idx, n * 4);
  VOX_ALLOC(idx_s, n * 4);
  VOX_ALLOC(sums, n * sizeof(VoxAcc));
  VOX_ALLOC(nruns, sizeof(long long));
  VOX_ALLOC(overflow, sizeof(int));
  DAV2_CUDA_OK(cudaMemsetAsync(overflow, 0, sizeof(int), stream));
  BoundsIt bit(CountIt(0), PointBounds{xyz, valid});
  GatherIt git(idx_s, GatherAcc{xyz, rgb});
  // temp storage: max over the three CUB calls
  DAV2_CUDA_OK(cub::DeviceReduce::Reduce(nullptr, need, bit, bounds, (int)n, BoundsMerge(), empty, stream));
  tmp_bytes = need;
  DAV2_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_s, idx, idx_s, (int)n, 0, 64, stream));
  tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
  DAV2_CUDA_OK(cub::DeviceReduce::ReduceByKey(nullptr, need, keys_s, keys, git, sums, nruns, VoxAdd(), (int)n, stream));
  tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
  VOX_ALLOC(tmp, tmp_bytes);
#undef VOX_ALLOC

  // bytes: bounds 12n, keys 12n+12n, sort ~4 passes x 24n, gather 12n(+12n) + keys 8n, outputs
  ProfScope ps(PC_OTHER, 0.0, (double)n * 160.0, stream);
  size_t tb = tmp_bytes;
  DAV2_CUDA_OK(cub::DeviceReduce::Reduce(tmp, tb, bit, bounds, (int)n, BoundsMerge(), empty, stream));
  vox_keys<<<vgrid(n), 256, 0, stream>>>(xyz, valid, n, bounds, voxel, keys, idx, overflow);
  DAV2_LAUNCH_OK();
  tb = tmp_bytes;
  DAV2_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_s, idx, idx_s, (int)n, 0, 64, stream));
  tb = tmp_bytes;
  DAV2_CUDA_OK(cub::DeviceReduce::ReduceByKey(tmp, tb, keys_s, keys, git, sums, nruns, VoxAdd(), (int)n, stream));
  vox_finalize<<<vgrid(n), 256, 0, stream>>>(keys, sums, nruns, overflow, out_xyz, out_rgb, out_count);
  DAV2_LAUNCH_OK();
  return 0;
}

}  // namespace dav2
