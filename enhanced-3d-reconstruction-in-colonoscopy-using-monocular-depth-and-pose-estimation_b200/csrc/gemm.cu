// Host launcher + template instantiations for the tcgen05 GEMM / implicit-GEMM conv kernel.
#include "gemm_tcgen05.cuh"

namespace dav2 {

int pick_bn(int N) {
  // smallest padded width wins; ties go to the wider tile (fewer A re-reads)
  const int cands[4] = {256, 128, 64, 32};
  int best = 256, best_pad = 1 << 30;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int pad = (N + bn - 1) / bn * bn;
    if (pad < best_pad) {
      best_pad = pad;
      best = bn;
    }
  }
  return best;
}

template <int BN, int MODE>
static int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;  // per instantiation; attribute is per-function, per-device (single device per process)
  if (!configured) {
    DAV2_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  if (tiles <= 0) return 0;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tcgen05_kernel<BN, MODE><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  DAV2_LAUNCH_OK();
  return 0;
}

#define DISPATCH_BN(MODE)                                                \
  switch (bn) {                                                          \
    case 256: return launch_t<256, MODE>(tmA, tmB, p, stream);           \
    case 128: return launch_t<128, MODE>(tmA, tmB, p, stream);           \
    case 64: return launch_t<64, MODE>(tmA, tmB, p, stream);             \
    case 32: return launch_t<32, MODE>(tmA, tmB, p, stream);             \
    default: break;                                                      \
  }

int launch_gemm(int bn, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                cudaStream_t stream) {
  switch (mode) {
    case GM_LINEAR_BF16: DISPATCH_BN(GM_LINEAR_BF16); break;
    case GM_LINEAR_RESID: DISPATCH_BN(GM_LINEAR_RESID); break;
    case GM_PATCH: DISPATCH_BN(GM_PATCH); break;
    case GM_CONVT: DISPATCH_BN(GM_CONVT); break;
    case GM_CONV_BF16: DISPATCH_BN(GM_CONV_BF16); break;
    case GM_CONV_HEAD:
      if (bn == 32) return launch_t<32, GM_CONV_HEAD>(tmA, tmB, p, stream);
      break;
    default: break;
  }
  set_last_error("launch_gemm: unsupported (bn=%d, mode=%d)", bn, mode);
  return -3;
}

}  // namespace dav2
