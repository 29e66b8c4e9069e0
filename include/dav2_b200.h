/* dav2_b200 -- C ABI of the B200-native depth + point-cloud hot path.
 *
 * Drop-in boundary.  The reference (prototypeanugrah/Enhanced-3D-Reconstruction-in-Colonoscopy-...)
 * is pure Python and has no FFI layer of its own; the hot path sits behind Python call signatures
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it replaces
 * (paths relative to the reference root).  The Python mirror of those interfaces lives in the
 * package (dpt.py, depth_to_pointcloud.py, evaluation.py, calculate_metrics.py) and binds this
 * library with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer documented as "device" is a CUDA device pointer owned by the
 *     CALLER (e.g. a torch tensor's data_ptr()); the library never frees caller memory;
 *   - the library owns its packed weights and workspace (allocated lazily, released by dav2_destroy);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *     no call synchronises except dav2_create / dav2_set_weight (host-side packing + H2D copy);
 *   - return 0 on success, <0 on error; dav2_last_error() returns a thread-local message;
 *   - one handle per (device, host thread); a handle is not thread-safe, and its workspace is shared by all of its
 *     calls: issue the forwards of ONE handle on ONE stream at a time (two streams need two handles, or an event
 *     between them); the handle must be used on the device it was created on (checked);
 *   - there is NO CPU fallback: without an sm_100 device every compute call fails.
 */
#ifndef DAV2_B200_H
#define DAV2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dav2_model dav2_model;

/* DepthAnythingV2(encoder, features, out_channels, use_bn=False, use_clstoken=False, max_depth)
 * -- external dpt.DepthAnythingV2.__init__, called at run.py:120-125, lightning_model.py:116-121,
 *    depth_to_pointcloud_dav2.py:159-164. */
typedef struct dav2_config {
  int32_t embed_dim;      /* 384 / 768 / 1024 (vits / vitb / vitl)            */
  int32_t depth;          /* 12 / 12 / 24 transformer blocks                   */
  int32_t num_heads;      /* 6 / 12 / 16 (head dim is always 64)               */
  int32_t features;       /* DPT width: 64 / 128 / 256                         */
  int32_t out_channels[4];/* reassemble widths, run.py:97-118                  */
  int32_t tap_layers[4];  /* 0-based block indices tapped for the DPT head     */
  float max_depth;        /* sigmoid scale, run.py:76 / configs/model/large.yaml:3 */
  int32_t precision;      /* tensor-core operand format: 0 = fp16 (the reference's AMP '16-mixed',
                             configs/trainer/default.yaml:4), 1 = bf16; accumulation is always fp32.
                             2 = fp32 validation engine: every contraction in fp32 SIMT arithmetic (the
                             "fp32 mode" of the 1e-4 parity gate; ~100x slower, not the benchmarked path) */
} dav2_config;

int dav2_create(dav2_model** out, const dav2_config* cfg);
void dav2_destroy(dav2_model* m);

/* load_state_dict (run.py:128-147, lightning_model.py:130-140): one call per upstream state-dict
 * key ("pretrained.blocks.3.attn.qkv.weight", ...).  `data` is HOST fp32, C-contiguous, `shape`
 * its dims.  The library converts / re-lays-out into its kernel formats (fp16/bf16 K-major GEMM operands,
 * tap-major 3x3 filters, pixel-shuffle-major transposed-conv filters, q pre-scaled by 1/8).
 * Unknown keys return a negative code (the Python layer implements strict=False on top). */
int dav2_set_weight(dav2_model* m, const char* key, const float* data, const int64_t* shape, int32_t ndim);
/* 1 once every tensor the forward pass needs has been supplied. */
int dav2_weights_complete(const dav2_model* m);
/* Position-embedding table for a ph x pw patch grid other than the checkpoint's 37x37: HOST fp32
 * [1 + ph*pw, embed_dim] (bicubic interpolation of the checkpoint table is input independent, so the
 * host layer computes it once per resolution exactly like upstream interpolate_pos_encoding). */
int dav2_set_pos_embed(dav2_model* m, int32_t ph, int32_t pw, const float* table);

/* DepthAnythingV2.forward(x[B,3,H,W]) -> depth[B,H,W]   (lightning_model.py:301, :358; external dpt.py)
 * x: device fp32 NCHW, already ImageNet-normalised; H, W multiples of 14.  depth: device fp32. */
int dav2_forward(dav2_model* m, const float* x, int32_t B, int32_t H, int32_t W, float* depth, void* stream);

/* Parity instrumentation (compare PRE-sigmoid logits, not only the saturating depth): when on, every
 * following dav2_forward also keeps the head's pre-sigmoid logits as the fp32 [B,H,W] debug buffer "logits"
 * (depth = max_depth * sigmoid(logit); external dpt.py DPTHead output_conv2).  Tensor-core engine only. */
int dav2_set_capture_logits(dav2_model* m, int32_t on);

/* Parity / debugging: look up an internal activation buffer of the LAST forward by name
 * ("tap0".."tap3" h16 [B*ph*pw, D]; "x" fp32 residual stream; "path1" ...).  Returns device ptr + bytes. */
int dav2_debug_buffer(dav2_model* m, const char* name, void** ptr, int64_t* bytes);
/* Copy the first `bytes` of that buffer into caller-owned device memory `dst` (async on `stream`). */
int dav2_debug_read(dav2_model* m, const char* name, void* dst, int64_t bytes, void* stream);

/* infer_image's final F.interpolate(depth[:,None], (h,w), mode="bilinear", align_corners=True)
 * (external dpt.py infer_image; used at run.py:234).  Device fp32 in/out. */
int dav2_resize_depth(const float* in, int32_t B, int32_t Hi, int32_t Wi, float* out, int32_t Ho, int32_t Wo,
                      void* stream);

/* Upstream image2tensor on the GPU (external dpt.py image2tensor, called through infer_image at run.py:233-234):
 * img device u8 [H,W,3] BGR (cv2.imread layout) -> RGB/255 -> cv2.INTER_CUBIC resize to (nh, nw) -> ImageNet
 * normalisation -> out device fp32 [3,nh,nw].  (nh, nw) = the lower-bound, multiple-of-14 size computed by the host. */
int dav2_preprocess_bgr_u8(const uint8_t* img, int32_t H, int32_t W, float* out, int32_t nh, int32_t nw, void* stream);

/* The same for a batch of equally sized frames (the run.py frame loop in batches): img device u8 [B,H,W,3] BGR ->
 * out device fp32 [B,3,nh,nw]. */
int dav2_preprocess_bgr_u8_batch(const uint8_t* img, int32_t B, int32_t H, int32_t W, float* out, int32_t nh, int32_t nw,
                                 void* stream);

/* Dataset pre-processing of the SimCol loader (data_processing/simcol.py:104-135 transform_input / transform_output,
 * :161-168 __getitem__): ToTensor -> transforms.Resize((Ho, Wo), BICUBIC, antialias=True) [-> Normalize], i.e. torch's
 * anti-aliased bicubic (a = -0.5, taps renormalised at the borders, kernel widened by the scale when down-sampling).
 * Inputs are divided by div_in (IEEE division, like `image.astype(np.float32) / 255.0`) before the resampling.
 *   mode 0: in device u8  [B,H,W,3] RGB (PIL order) / div_in (255) -> resize -> ImageNet normalisation -> out fp32 [B,3,Ho,Wo]
 *   mode 1: in device u16 [B,H,W] / div_in (65535) -> resize                                        -> out fp32 [B,1,Ho,Wo]
 *   mode 2: in device fp32 [B,H,W] / div_in -> resize                                               -> out fp32 [B,1,Ho,Wo] */
int dav2_resize_aa(int32_t mode, const void* in, int32_t B, int32_t H, int32_t W, float* out, int32_t Ho, int32_t Wo,
                   float div_in, void* stream);

/* Fused back-projection + SE(3) world transform + validity mask.
 * Replaces depth_to_pointcloud.py:218-239 (Open3D RGBD -> PointCloud.create_from_rgbd_image -> transform)
 * and the explicit formula at depth_to_pointcloud_dav2.py:300-313.
 *   depth  device fp32 [B,H,W]
 *   K4     device fp64 [B,4] (k_per_frame=1) or [4] (k_per_frame=0): fx, fy, cx, cy
 *   T12    device fp64 [B,12] row-major [R|t] (depth_to_pointcloud.py:170-173) or NULL (camera frame)
 *   z = depth / depth_scale; pixel valid iff 0 < z < depth_trunc and finite (Open3D defaults 1000 / 3.0;
 *   pass depth_scale=1, depth_trunc=INFINITY for metric depth straight from the network)
 *   xyz    device fp32 [B,H*W,3] dense, row-major pixels (invalid -> 0,0,0)
 *   valid  device u8   [B,H*W] or NULL;   counts device i32 [B] or NULL (# valid points per frame) */
int dav2_backproject(const float* depth, int32_t B, int32_t H, int32_t W, const double* K4, int32_t k_per_frame,
                     const double* T12, float depth_scale, float depth_trunc, float* xyz, uint8_t* valid,
                     int32_t* counts, void* stream);

/* Fused back-projection + point-cloud all-gather ("the kernel's output write is the collective").
 * Same arithmetic as dav2_backproject, but every result is stored to n_dst (<= 8) destination buffers -- the gather
 * buffers of all ranks, peer-mapped over NVLink (dav2_peer_*) -- at frame index frame_offset + b, so no separate
 * all-gather runs afterwards.  xyz_dst / valid_dst / counts_dst are HOST arrays of n_dst device pointers to the BASE of
 * each gathered buffer (xyz [world*B,H*W,3] fp32, valid [world*B,H*W] u8 or NULL array, counts [world*B] i32 or NULL
 * array).  The caller orders readers after all writers with any later collective on the same streams (the metric
 * all-reduce of the step does it) and double-buffers across steps. */
int dav2_backproject_gather(const float* depth, int32_t B, int32_t H, int32_t W, const double* K4, int32_t k_per_frame,
                            const double* T12, float depth_scale, float depth_trunc, float* const* xyz_dst,
                            uint8_t* const* valid_dst, int32_t* const* counts_dst, int32_t n_dst, int64_t frame_offset,
                            void* stream);

/* Back-projection AND the test_step metric partial sums in ONE pass over the depth map (round 2): the same arithmetic and
 * destinations as dav2_backproject_gather (n_dst = 1 with one-element arrays for a purely local cloud), plus
 * dav2_depth_metrics variant 0 (mask lo <= gt <= hi, lightning_model.py:304-313 + eval/evaluation.py:16-60) on
 * (pred = depth, gt) -- gt device fp32 [B,H*W] -- accumulated from the registers that already hold the depth, so the
 * 4 B/px re-read of the stand-alone metric kernel disappears (21 B/px for the fused pass instead of 17 + 8).
 * partials device fp64 [B,8] (per_frame = 1) or [8], zeroed by the call; same slots as dav2_depth_metrics. */
int dav2_backproject_metrics(const float* depth, const float* gt, int32_t B, int32_t H, int32_t W, const double* K4,
                             int32_t k_per_frame, const double* T12, float depth_scale, float depth_trunc,
                             float* const* xyz_dst, uint8_t* const* valid_dst, int32_t* const* counts_dst, int32_t n_dst,
                             int64_t frame_offset, float lo, float hi, int32_t per_frame, double* partials, void* stream);

/* Peer-mapped device buffers for dav2_backproject_gather: one process per GPU allocates its gather buffer
 * (dav2_peer_alloc: cudaMalloc, so the allocation is IPC-exportable), exports a 64-byte handle (cudaIpcGetMemHandle),
 * exchanges the handles through torch.distributed, and maps every other rank's buffer (dav2_peer_open:
 * cudaIpcOpenMemHandle with lazy peer access).  Replaces the NCCL all_gather the reference path would otherwise need
 * after depth_to_pointcloud.py:332-354 when the frames are sharded over GPUs. */
int dav2_peer_alloc(void** ptr, int64_t bytes);
int dav2_peer_free(void* ptr);
int dav2_peer_export(const void* ptr, uint8_t* handle64);
int dav2_peer_open(const uint8_t* handle64, void** ptr);
int dav2_peer_close(void* ptr);

/* Voxel-grid down-sample of a fused cloud: depth_to_pointcloud.py:357-359 (`combined.voxel_down_sample(voxel_size=0.01)`,
 * i.e. Open3D PointCloud::VoxelDownSample): voxel index = floor((p - (min_bound - voxel/2)) / voxel) per axis, one output
 * point per occupied voxel = the mean (accumulated in fp64) of its points, colours averaged the same way.
 *   xyz    device fp32 [n,3];  rgb device fp32 [n,3] or NULL;  valid device u8 [n] or NULL (0 / non-finite -> dropped)
 *   out_xyz device fp32 [n,3] (worst case), out_rgb device fp32 [n,3] or NULL; rows [0, *out_count) are written,
 *   ordered by ascending (ix, iy, iz) (Open3D emits hash-map order: compare as sets)
 *   out_count device i64 [1]: number of voxels, or -1 when the grid exceeds 2^21 cells on an axis
 *   (Open3D's "voxel_size is too small" error).  n < 2^31. */
int dav2_voxel_downsample(const float* xyz, const float* rgb, const uint8_t* valid, int64_t n, double voxel_size,
                          float* out_xyz, float* out_rgb, int64_t* out_count, void* stream);

/* Depth-metric partial sums (finalise on the host AFTER any cross-GPU sum).
 *   variant 0: evaluation.compute_errors on the batch-wide mask lo <= gt <= hi
 *              (eval/evaluation.py:16-60, lightning_model.py:304-313)
 *   variant 1: calculate_metrics.calculate_metrics mask gt>0 & pred>0 & finite (calculate_metrics.py:17-55)
 *   variant 2: evaluation.compute_errors on inputs the caller already masked (every element counts)
 *   variant 3: calculate_metrics(mask_invalid=False) (every element counts)
 *   partials   device fp64 [B,8] (per_frame=1) or [8]:
 *              {n, sum|d|, sum|d|/(gt+1e-6), sum d^2, sum gt, #(t<a), #(t<b), #(t<c)}, t = max(gt/pred, pred/gt),
 *              (a,b,c) = (1.25, 1.25^2, 1.25^3) for variants 1 and 3; for variants 0 and 2 slot 5 is #(t<1.1) and
 *              slots 6, 7 count NaN / Inf predictions (the warnings of eval/evaluation.py:33-36). */
int dav2_depth_metrics(const float* pred, const float* gt, int32_t B, int64_t HW, float lo, float hi,
                       int32_t variant, int32_t per_frame, double* partials, void* stream);

/* o3d PointCloud.transform(T) (depth_to_pointcloud.py:236-239) for a cloud already on the device: xyz device fp32 [n,3],
 * in place, p <- R p + t evaluated in fp64 with one rounding; T12 device fp64 [12] = rows of [R|t]. */
int dav2_transform_points(float* xyz, int64_t n, const double* T12, void* stream);

/* evaluation.compose_poses (eval/evaluation.py:279-382): rel device fp32 [N,7] (t | q xyzw),
 * init7 device fp32 [7] or NULL (identity) -> abs7 device fp32 [N+1,7]; optional T12 device fp64
 * [N+1,12] = rows of [R|t] with R = Rotation.from_quat(q).as_matrix() (depth_to_pointcloud.py:168-173). */
int dav2_compose_poses(const float* rel, const float* init7, int32_t N, float* abs7, double* T12, void* stream);

/* PoseEstimationNet(in_channels=8).forward (pose_estimation_model.py:35-105): ResNet-18 with an 8-channel stem on
 * stacked frame pairs [rgb1, d1, rgb2, d2] (data_processing/pose_estimation.py:229-243) -> fc 256 -> MLP -> 7
 * = [t(3) | q xyzw(4)], eval semantics (BatchNorm running statistics, Dropout = identity).
 * dav2_pose_set_weight takes HOST fp32 tensors with BatchNorm ALREADY FOLDED by the host layer:
 *   "<conv>.weight" [Cout,Cin,k,k] and "<conv>.bias" [Cout] for conv1, layer{1..4}.{0,1}.conv{1,2},
 *   layer{2..4}.0.downsample; "fc.weight" [256,512], "fc.bias"; "head.{0,1,2}.weight/bias" (the three Linear layers).
 * dav2_pose_forward: x device fp32 [B,8,H,W] -> pose7 device fp32 [B,7]. */
typedef struct dav2_pose dav2_pose;
int dav2_pose_create(dav2_pose** out, int32_t precision);
void dav2_pose_destroy(dav2_pose* m);
int dav2_pose_set_weight(dav2_pose* m, const char* key, const float* data, const int64_t* shape, int32_t ndim);
int dav2_pose_forward(dav2_pose* m, const float* x, int32_t B, int32_t H, int32_t W, float* pose7, void* stream);

/* Operator-level entry points (unit parity tests + reuse).  "h16" operands are 16-bit device tensors whose
 * numeric format is given by `fmt` (0 = fp16, 1 = bf16); accumulation is fp32.
 *   C[M,N] = act(A[M,K] * W[N,K]^T + bias)    A, W, C h16 row-major; bias fp32 or NULL; act 0/1(GELU)/2(ReLU) */
int dav2_linear_h16(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N, int32_t K,
                    int32_t act, int32_t fmt, void* stream);
/*   x[M,N](fp32) += gamma[N] * (A[M,K] * W[N,K]^T + bias[N])   (LayerScale + residual epilogue) */
int dav2_linear_resid(const void* A, const void* W, const float* bias, const float* gamma, float* x, int32_t M,
                      int32_t N, int32_t K, int32_t fmt, void* stream);
/*   3x3 / pad 1 / stride 1 conv, NHWC h16: in [B,H,W,Cin], Wp [Cout, 9*Cpad] (tap-major, Cpad = ceil64(Cin)),
 *   out [B,H,W,Cout] = act(conv + bias) + add1 + add2; out_relu (optional) = relu(out). */
int dav2_conv3x3_h16(const void* in, const void* Wp, const float* bias, const void* add1, const void* add2,
                     void* out, void* out_relu, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                     int32_t act, int32_t fmt, void* stream);
/*   softmax(q k^T) v per (image, head), d_head 64, q pre-scaled: qkv h16 [B*N, 3*D] -> out h16 [B*N, D] */
int dav2_attention_h16(const void* qkv, void* out, int32_t B, int32_t N, int32_t D, int32_t fmt, void* stream);
/*   LayerNorm(eps) fp32 [rows, D] -> h16 */
int dav2_layernorm(const float* x, const float* w, const float* b, void* out, int64_t rows, int32_t D, float eps,
                   int32_t fmt, void* stream);
/*   bilinear align_corners=True, NHWC h16 */
int dav2_bilinear_nhwc_h16(const void* in, void* out, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo,
                           int32_t C, int32_t fmt, void* stream);

/* Per-kernel-class timing with CUDA events recorded on the launching stream (bench.py's live roofline
 * measurement).  dav2_profile_report synchronises on the pending events, then writes a JSON object
 * {"gemm_tcgen05": {"launches":..,"ms":..,"flops":..,"bytes":..}, ...} (algorithmic flops / bytes). */
void dav2_profile_enable(int32_t on);
int dav2_profile_report(char* buf, int32_t cap);

const char* dav2_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
int64_t dav2_launch_count(void);
const char* dav2_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DAV2_B200_H */
