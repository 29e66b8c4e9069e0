// Host launcher + template instantiations for the tcgen05 GEMM / implicit-GEMM conv kernel.
#include "conv_halo_tcgen05.cuh"

#include <stdlib.h>

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
  static char tag;  // per instantiation; the attribute is per function AND per device
  if (!device_setup_done(&tag)) {
    DAV2_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    device_setup_mark(&tag);
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

template <int BN, int MODE>
static int launch2_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, gemm2_wide_epi(BN, MODE)>;
  static char tag;
  if (!device_setup_done(&tag)) {
    DAV2_CUDA_OK(cudaFuncSetAttribute(gemm2_tcgen05_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    device_setup_mark(&tag);
  }
  const int pairs = ((p.tiles_m + 1) / 2) * p.tiles_n;
  if (pairs <= 0) return 0;
  const int max_pairs = sm_count() / 2;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  gemm2_tcgen05_kernel<BN, MODE><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  DAV2_LAUNCH_OK();
  return 0;
}

#define DISPATCH2_BN(MODE)                                               \
  switch (bn) {                                                          \
    case 256: return launch2_t<256, MODE>(tmA, tmB, p, stream);          \
    case 128: return launch2_t<128, MODE>(tmA, tmB, p, stream);          \
    default: break;                                                      \
  }

// 2-CTA (cta_group::2) kernel: bn in {128, 256}; tmB must have been encoded with box rows bn/2.
int launch_gemm2(int bn, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                 cudaStream_t stream) {
  switch (mode) {
    case GM_LINEAR_BF16: DISPATCH2_BN(GM_LINEAR_BF16); break;
    case GM_LINEAR_RESID: DISPATCH2_BN(GM_LINEAR_RESID); break;
    case GM_PATCH: DISPATCH2_BN(GM_PATCH); break;
    case GM_CONVT: DISPATCH2_BN(GM_CONVT); break;
    case GM_CONV_BF16: DISPATCH2_BN(GM_CONV_BF16); break;
    default: break;
  }
  set_last_error("launch_gemm2: unsupported (bn=%d, mode=%d)", bn, mode);
  return -3;
}

#ifdef DAV2_PROFILING_KNOBS
static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#endif

template <int BN, int MODE>
static int launch_halo_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = ConvHaloCfg<BN>;
  static char tag;
  int bo_mode = 0;
#ifdef DAV2_PROFILING_KNOBS
  bo_mode = env_flag("DAV2_HALO_BO", 0);
#endif
  if (!device_setup_done(&tag)) {
    DAV2_CUDA_OK(cudaFuncSetAttribute(conv_halo_tcgen05_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    device_setup_mark(&tag);
  }
  const int pairs = ((p.tiles_m + 2 * Cfg::MT - 1) / (2 * Cfg::MT)) * p.tiles_n;
  if (pairs <= 0) return 0;
  if (Cfg::RESIDENT_B && (p.num_kb > Cfg::B_STAGES || p.tiles_n != 1)) {
    set_last_error("conv_halo<%d>: %d weight tiles do not fit the resident set (%d)", BN, p.num_kb, Cfg::B_STAGES);
    return -3;
  }
  // the N=32 head variant needs ~110 KB smem and 64 TMEM columns: two CTA pairs fit per TPC
  const int max_pairs = (sm_count() / 2) * (Cfg::SMEM_BYTES <= 112 * 1024 ? 2 : 1);
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  conv_halo_tcgen05_kernel<BN, MODE><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p, bo_mode);
  DAV2_LAUNCH_OK();
  return 0;
}

int launch_conv_halo(int bn, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  if (mode == GM_CONV_HEAD && bn == 32) return launch_halo_t<32, GM_CONV_HEAD>(tmA, tmB, p, stream);
  if (mode == GM_CONV_BF16 && bn == 256) return launch_halo_t<256, GM_CONV_BF16>(tmA, tmB, p, stream);
  if (mode == GM_CONV_BF16 && bn == 128) return launch_halo_t<128, GM_CONV_BF16>(tmA, tmB, p, stream);
  set_last_error("launch_conv_halo: unsupported bn=%d mode=%d", bn, mode);
  return -3;
}

// Kernel selection is a pure function of the shape in the shipped library; profiling builds (-DDAV2_PROFILING_KNOBS) can
// force the per-tap / 1-CTA kernels with DAV2_CONV_HALO=0 / DAV2_GEMM2=0 for A/B measurements.
static bool knob_on(const char* name) {
#ifdef DAV2_PROFILING_KNOBS
  return env_flag(name, 1) == 1;
#else
  (void)name;
  return true;
#endif
}

bool conv_halo_eligible(int bn, int mode, int tiles_m) {
  if (!knob_on("DAV2_CONV_HALO") || tiles_m < 2) return false;
  if (mode == GM_CONV_HEAD) return bn == 32 && knob_on("DAV2_GEMM2");
  return mode == GM_CONV_BF16 && (bn == 128 || bn == 256) && gemm2_eligible(bn, mode, tiles_m);
}

bool gemm2_eligible(int bn, int mode, int tiles_m) {
  return knob_on("DAV2_GEMM2") && mode != GM_CONV_HEAD && (bn == 128 || bn == 256) && tiles_m >= 2;
}

}  // namespace dav2
