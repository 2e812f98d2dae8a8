/*
 * ray3d_b200.h -- C ABI of the B200-native Ray3D lifting forward pass.
 *
 * One shared library (libray3d_b200.so, hand-written sm_100a CUDA) replaces, for eval-mode
 * inference, the PyTorch module graph the reference dispatches for its 2D->3D lifting path.
 * Every entry point names the reference interface it stands in for (paths are relative to the
 * reference checkout, YxZhxn/Ray3D @ aff4b9f).  No torch types cross this boundary: plain
 * pointers, sizes and a cudaStream_t passed as void*.
 *
 * Conventions
 *   - every function returns 0 on success or a negative r3d_status; r3d_last_error() gives the
 *     message for the calling thread.  Nothing throws across the ABI.
 *   - "_dev" pointers are device pointers on the plan's device; "_host" pointers are host memory.
 *   - forward calls are asynchronous on `stream` (the caller's current stream) and do no host
 *     allocation; the *_host variants copy in/out themselves and synchronise before returning.
 *   - a plan's activation workspace is single-buffered per lane: forwards that share a lane are ordered one after
 *     the other ON THE DEVICE (an event recorded after each, waited for by the next), whatever streams they were
 *     issued on -- mixing r3d_forward_* on several streams with r3d_submit_* is safe, it just does not overlap.
 *   - outputs are always freshly written caller-owned buffers (the reference's callers mutate
 *     the returned tensors in place, lib/train_val/trainer.py:215,340,353).
 */
#ifndef RAY3D_B200_H_
#define RAY3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define R3D_API __attribute__((visibility("default")))
#else
#define R3D_API
#endif

#define R3D_ABI_VERSION 2
#define R3D_MAX_WIDTHS 8

typedef enum {
  R3D_OK = 0,
  R3D_ERR_BAD_ARG = -1,        /* null pointer, bad shape; mirrors the asserts at lib/model/rie.py:285-287 */
  R3D_ERR_UNSUPPORTED = -2,    /* config outside what the reference itself supports */
  R3D_ERR_MISSING_WEIGHT = -3, /* finalize() before every state_dict tensor was supplied */
  R3D_ERR_CUDA = -4,           /* a CUDA runtime/driver call failed */
  R3D_ERR_STATE = -5,          /* call order violated (e.g. forward before upload) */
  R3D_ERR_NO_DEVICE = -6       /* no usable sm_100 device: the product path never falls back to CPU */
} r3d_status;

/* Arithmetic used for the conv/linear contractions. */
typedef enum {
  R3D_PREC_FP32 = 0,   /* FP32 FFMA everywhere (bit-for-bit fp32 operands, fp32 accumulate) */
  R3D_PREC_BF16X3 = 1, /* tcgen05 tensor cores, operands split hi+lo bf16, 3 products, fp32 accumulate
                          (~1e-5 normwise vs fp64; meets the 1e-4 fp32 parity bar) */
  R3D_PREC_BF16 = 2    /* tcgen05, single bf16 product (BASELINE config 3; ~3e-3 normwise) */
} r3d_precision;

/* Which of the reference's two nn.Modules a plan computes. */
#define R3D_NET_POS 1 /* RIEModel            lib/model/rie.py:172-434 */
#define R3D_NET_TRJ 2 /* RIETrajectoryModel  lib/model/rie.py:437-559 */

/* Mirrors the constructor arguments lib/model/__init__.py:11-46 derives from cfg_*.model_config
 * (NUM_KPTS, INPUT_DIM, ARCHITECTURE, CHANNELS, LATENT_FEATURES_DIM, STAGE, EXTRINSIC_DIM,
 * EMBEDD_DIM; CAUSAL/DENSE/DISABLE_OPTIMIZATIONS must be False -- the only combination that runs in
 * the reference). */
typedef struct {
  int32_t num_joints;               /* 17, 15 or 14 */
  int32_t in_features;              /* 3 (ray encoding) or 2 (cfg_rie_*) */
  int32_t n_widths;                 /* len(ARCHITECTURE) */
  int32_t widths[R3D_MAX_WIDTHS];   /* odd filter widths; receptive field = product */
  int32_t channels;                 /* CHANNELS (multiple of 64) */
  int32_t latent;                   /* LATENT_FEATURES_DIM (multiple of 64) */
  int32_t stage;                    /* 1, or !=1 for FuseBlocks */
  int32_t extrinsic_dim;            /* 0 disables the camera embedding */
  int32_t embed_dim;
  int32_t nets;                     /* R3D_NET_POS | R3D_NET_TRJ */
  int32_t precision;                /* r3d_precision */
} r3d_config;

typedef struct r3d_plan r3d_plan;

R3D_API int r3d_abi_version(void);
R3D_API const char* r3d_last_error(void);

/* --- plan lifecycle: replaces Model(model_config, ...) + load_state_dict ------------------------
 * r3d_plan_create    <- RIEModel.__init__ / RIETrajectoryModel.__init__ (rie.py:178-253, 443-494).
 *                       Host only, no CUDA call.
 * r3d_plan_set_tensor<- one state_dict entry, by the reference's own key (e.g.
 *                       "LocalLayer_Torso.expand_conv.weight"; a leading "module." from
 *                       nn.DataParallel checkpoints is accepted, trainer.py:232-240).  `net` is
 *                       R3D_NET_POS or R3D_NET_TRJ.  Shape-checked: unlike lib/utils/utils.py:208-218
 *                       (load_weight silently drops mismatches) an unknown key or wrong shape is an error.
 * r3d_plan_finalize  <- folds eval-mode BatchNorm (running stats, eps 1e-5) into the preceding
 *                       conv/linear in float64, repacks to the kernels' K-major layout and splits to
 *                       bf16 hi/lo for the tensor-core precisions.  Host only.
 * r3d_plan_upload    <- .cuda() (lib/model/__init__.py:51-53): copies packed weights to `device`.
 */
R3D_API int r3d_plan_create(const r3d_config* cfg, r3d_plan** out);
R3D_API int r3d_plan_set_tensor(r3d_plan* plan, int net, const char* name, const float* data_host,
                        const int64_t* shape, int ndim);
R3D_API int r3d_plan_finalize(r3d_plan* plan);
R3D_API int r3d_plan_upload(r3d_plan* plan, int device);
R3D_API void r3d_plan_destroy(r3d_plan* plan);

/* Introspection used by the CPU test-suite (no GPU needed): packed, BN-folded fp32 weights of one
 * layer.  `layer` is the reference module path, e.g. "LocalLayer_LArm.layers_conv.0" or
 * "GlobalInfo.fc_1".  Returns rows/cols of the K-major matrix [n_pad][k_pad]; copies
 * min(cap, n_pad*k_pad) weights and min(cap_b, n_pad) biases. */
R3D_API int r3d_plan_packed_layer(const r3d_plan* plan, int net, const char* layer, int32_t* n_pad, int32_t* k_pad,
                          float* w_out, int64_t cap, float* b_out, int64_t cap_b);
/* JSON description of the launch graph (activation buffers, grouped GEMM ops and their bindings, input-stage
 * gather tables, output slots).  Writes at most cap bytes (NUL terminated) and the required size to *needed.
 * Used by the CPU tests to replay the wiring with numpy; host only. */
R3D_API int r3d_plan_describe(const r3d_plan* plan, char* out, int64_t cap, int64_t* needed);
R3D_API int64_t r3d_plan_weight_bytes(const r3d_plan* plan);     /* device bytes of packed weights */
R3D_API int64_t r3d_plan_workspace_bytes(const r3d_plan* plan);  /* device bytes of activations at current capacity */
R3D_API int r3d_plan_receptive_field(const r3d_plan* plan);       /* rie.py:278-282 */
R3D_API int r3d_plan_kernel_launches(const r3d_plan* plan);
/* Forward calls on device pointers with batch <= 64 (R3D_GRAPH_MAX_BATCH) replay a CUDA graph of the launch sequence,
 * captured once per (batch, input kind, output set): one cudaGraphLaunch instead of ~20 launches on two streams.
 * Returns how many forwards of this plan went through a graph (instrumentation). */
R3D_API int64_t r3d_plan_graph_launches(const r3d_plan* plan);       /* kernels one forward enqueues */

/* Per-launch device timing (benchmark instrumentation): when enabled every forward records CUDA events on the
 * launch stream around each kernel.  r3d_plan_launch_times synchronises on the recorded events and returns the
 * mean duration (ms) of each launch position over the recorded forwards (ring of the last 64).
 * Position 0 is the input stage, the last is the output stage, the others the grouped GEMMs in graph order. */
R3D_API int r3d_plan_set_profiling(r3d_plan* plan, int enable);
R3D_API int r3d_plan_launch_times(r3d_plan* plan, float* ms_out, int32_t cap, int32_t* n_launches, int32_t* n_runs);
R3D_API const char* r3d_plan_launch_name(const r3d_plan* plan, int32_t index);

/* --- what a forward reads ----------------------------------------------------------------------------
 * One descriptor covers every input form the reference's eval loop produces:
 *   src_kind  R3D_SRC_RAYS  windows of encoded input (B, T, J, Cin) float32 -- what nn.Module.forward receives
 *                           (rie.py:284-304; T == receptive field)
 *             R3D_SRC_UV    windows of pixel keypoints (B, T, J, 2) float32; CameraInfoPacket.get_cam_ray_given_uv
 *                           (lib/camera/camera.py:423-441, 460-471) runs inside the input stage, in float64, rounded to
 *                           float32 exactly like trainer.py:298.  in_features must be 3.
 *   window_stride           floats between consecutive windows: T*J*C for materialised windows; J*C for the
 *                           frames of ONE edge-padded video (F + RF - 1, J, C), whose F sliding windows are then
 *                           indexed in place -- replaces Trainer.eval_data_prepare (trainer.py:47-58, 323-337)
 *   cam_kind  R3D_CAM_PARAM (B, extrinsic_dim) float32 = the reference's `param` [height_m, pitch_rad]
 *                           (trainer.py:297,324); only with R3D_SRC_RAYS; may be NULL when the embedding is off
 *             R3D_CAM_F32   (B, 6) float32 [fx, fy, cx, cy, pitch_rad, height_m]; sin/cos evaluated on the device
 *             R3D_CAM_F64   (B, R3D_CAM64_STRIDE) float64 [fx, fy, ppx, ppy, cos(pitch), sin(pitch), pitch, height,
 *                           k1, k2, p1, p2, k3, undistort (0/1), K02, K12]: the reference's own float64 calibration
 *                           (camera.py:438-439 divides by float64 K entries; Rc2n from libm cos/sin, :333-338), pp =
 *                           CameraInfoPacket.pp_cam (camera.py:253-259: the undistorted principal point of a distorted
 *                           lens), and the 5-coefficient lens model of cv2.undistortPoints around K's own principal
 *                           point K02, K12 (camera.py:412-421, 435-436)
 *   cam_stride              elements between camera rows; 0 = one row shared by every window (a video)
 *   flags     R3D_IN_UNDISTORT  some row has its undistort flag set (selects the kernel variant with the lens model)
 *             R3D_IN_FLIP_TTA   flip test-time augmentation (trainer.py:299-302, 338-353; needs r3d_plan_set_flip)
 * Overlapping R3D_SRC_UV windows with a shared camera row are encoded ONCE per frame (not once per window). */
#define R3D_SRC_RAYS 0
#define R3D_SRC_UV 1
#define R3D_CAM_PARAM 0
#define R3D_CAM_F32 1
#define R3D_CAM_F64 2
#define R3D_CAM64_STRIDE 16
#define R3D_IN_UNDISTORT 1
#define R3D_IN_FLIP_TTA 2

typedef struct {
  const float* src;
  int64_t window_stride;
  int32_t src_kind;
  int32_t cam_kind;
  const void* cam;
  int64_t cam_stride;
  int32_t flags;
  int32_t reserved;
} r3d_input;

/* Generic forms; every named entry point below is one of these with a fixed descriptor.
 * r3d_forward       device pointers, asynchronous on `stream`.
 * r3d_submit/r3d_join  device pointers, on one of the plan's two lanes (see r3d_submit_rays).
 * r3d_forward_host  host pointers; H2D, kernels, D2H inside the call (see r3d_forward_rays_host).
 * r3d_submit_host / r3d_wait  the streaming form of it (see r3d_submit_rays_host). */
R3D_API int r3d_forward(r3d_plan* plan, const r3d_input* in_dev, float* pos_dev, float* trj_dev, float* sum_dev,
                        int32_t n_windows, void* stream);
R3D_API int r3d_submit(r3d_plan* plan, const r3d_input* in_dev, float* pos_dev, float* trj_dev, float* sum_dev,
                       int32_t n_windows, void* stream, uint64_t* ticket);
R3D_API int r3d_forward_host(r3d_plan* plan, const r3d_input* in_host, float* pos_host, float* trj_host,
                             float* sum_host, int32_t n_windows);
R3D_API int r3d_submit_host(r3d_plan* plan, const r3d_input* in_host, float* pos_host, float* trj_host, float* sum_host,
                            int32_t n_windows, uint64_t* ticket);

/* Result-neutral tuning: "graph_max_batch" (CUDA-graph replay for batches <= n, default 64, 0 off), "lanes" (1 or 2,
 * default 2), "side_stream" (0/1, default 1), "host_chunk" (windows per staged chunk of a host-buffer call, 0 = auto),
 * and, before r3d_plan_finalize only: "side_chain" (0 off / 1 when worth it (default) / 2 always: the GlobalInfo chain,
 * rie.py:362, as ONE persistent kernel on a few CTA pairs of the side stream, its work units ordered by per-row-group
 * completion counters, instead of six under-filled launches racing the tree's kernels for SMs), "side_clusters" (CTA
 * pairs it may hold; 0 = from its flop share), "tail_fusion" (0/1, default 0: the same for the one-row layers of the main
 * chain -- top tree level, shrink, FuseBlocks, Integration; bit-identical, measured slower on B200, kept for study),
 * "tail_width" (128/256, its unit width) and "tile_policy" (GEMM tile width of the narrow launches: 1 = latency, the
 * shortest launch for a forward that has the GPU to itself; 2 = throughput, the widest tile, least SM time per flop
 * when two batches are in flight; 0 = auto (default): throughput for plans of >= 256 windows capacity). */
R3D_API int r3d_plan_set_option(r3d_plan* plan, const char* name, int32_t value);

/* --- forward: replaces nn.Module.forward(x, param) ------------------------------------------------
 * x_dev     (B, T, J, Cin) float32 contiguous, T == receptive field (rie.py:284-304; the reference
 *           only works for T == RF).
 * param_dev (B, extrinsic_dim) float32 = [height_m, pitch_rad] (trainer.py:297,324); ignored (may be
 *           NULL) when the camera embedding is off.
 * pos_dev   (B, 1, J, 3) float32 or NULL   <- RIEModel.forward            rie.py:284-434
 * trj_dev   (B, 1, 1, 3) float32 or NULL   <- RIETrajectoryModel.forward  rie.py:518-559
 * sum_dev   (B, 1, J, 3) float32 or NULL   <- predicted_3d_pos += predicted_3d_trj, trainer.py:353
 */
R3D_API int r3d_forward_rays(r3d_plan* plan, const float* x_dev, const float* param_dev, float* pos_dev,
                     float* trj_dev, float* sum_dev, int32_t batch, void* stream);

/* Same, but starting from pixel keypoints: fuses CameraInfoPacket.get_cam_ray_given_uv
 * (lib/camera/camera.py:423-441, 460-471; undistort=False) into the input stage.
 * uv_dev  (B, T, J, 2) float32 pixels;  cam_dev (B, 6) float32 = [fx, fy, cx, cy, pitch_rad, height_m].
 * The encode runs in float64 and rounds to float32 exactly like trainer.py:298. in_features must be 3. */
R3D_API int r3d_forward_uv(r3d_plan* plan, const float* uv_dev, const float* cam_dev, float* pos_dev,
                   float* trj_dev, float* sum_dev, int32_t batch, void* stream);
/* Same with the reference's float64 calibration rows (R3D_CAM_F64, see r3d_input) and, when `undistort` != 0, the lens
 * undistortion of CameraInfoPacket.encode_uv_with_intrinsic (camera.py:435-436) inside the input stage. */
R3D_API int r3d_forward_uv_cam64(r3d_plan* plan, const float* uv_dev, const double* cam64_dev, int32_t undistort,
                                 float* pos_dev, float* trj_dev, float* sum_dev, int32_t batch, void* stream);

/* Host-buffer variants (the end-to-end call: H2D of inputs, forward, D2H of results, chunked and
 * double-buffered on internal streams; returns after the results are in host memory).
 * Stand in for trainer.py:329-356 (.cuda() ... forward ... .cpu()). */
R3D_API int r3d_forward_rays_host(r3d_plan* plan, const float* x_host, const float* param_host, float* pos_host,
                          float* trj_host, float* sum_host, int32_t batch);
R3D_API int r3d_forward_uv_host(r3d_plan* plan, const float* uv_host, const float* cam_host, float* pos_host,
                        float* trj_host, float* sum_host, int32_t batch);

/* Asynchronous form of the two calls above for streaming use (a loader thread feeding batches): the H2D copy, the
 * launches and the D2H copy are only enqueued -- on the plan's own copy/compute streams with two device staging slots,
 * so the copy of one submission overlaps the kernels of the previous one -- and *ticket names the submission.  Host
 * buffers must be page-locked and stay untouched until r3d_wait(plan, ticket) returns; results are then in pos/trj/sum.
 * Consecutive submissions alternate between the plan's two lanes (each with its own workspace and streams, sharing
 * the weights), so besides the copy/compute overlap the under-filled tail launches of one batch run beside the large
 * launches of the next; they may therefore complete out of order.  At most 8 may be outstanding (the 9th submit
 * blocks on the oldest).
 * Stands in for the reference's DataLoader(pin_memory) + .cuda() + forward + .cpu() loop body, trainer.py:318-364. */
R3D_API int r3d_submit_rays_host(r3d_plan* plan, const float* x_host, const float* param_host, float* pos_host, float* trj_host,
                                 float* sum_host, int32_t batch, uint64_t* ticket);
R3D_API int r3d_submit_uv_host(r3d_plan* plan, const float* uv_host, const float* cam_host, float* pos_host, float* trj_host,
                               float* sum_host, int32_t batch, uint64_t* ticket);
R3D_API int r3d_wait(r3d_plan* plan, uint64_t ticket);

/* Asynchronous DEVICE-buffer form of r3d_forward_rays / r3d_forward_uv: the launch sequence is ordered after the work
 * already enqueued on `stream` and runs on one of the plan's two lanes (alternating per submission); `*ticket` names
 * it.  r3d_join makes `stream` wait for it on the device (no host blocking), r3d_wait blocks the host.  All buffers
 * must stay valid and untouched until then.  Keeping two submissions in flight is what a loop over independent
 * batches of windows (trainer.py:318-364) should do: throughput is then bounded by the large launches only. */
R3D_API int r3d_submit_rays(r3d_plan* plan, const float* x_dev, const float* param_dev, float* pos_dev, float* trj_dev,
                            float* sum_dev, int32_t batch, void* stream, uint64_t* ticket);
R3D_API int r3d_submit_uv(r3d_plan* plan, const float* uv_dev, const float* cam_dev, float* pos_dev, float* trj_dev,
                          float* sum_dev, int32_t batch, void* stream, uint64_t* ticket);
R3D_API int r3d_join(r3d_plan* plan, uint64_t ticket, void* stream);

/* Sliding-window evaluation of one video without materialising windows:
 * replaces Trainer.eval_data_prepare + np.tile(cam_param) + forward (trainer.py:47-58, 323-337).
 * seq_dev (F + RF - 1, J, Cin) float32 (already edge-padded like generators.py:209-234);
 * param_dev (extrinsic_dim) float32 shared by all windows; outputs have F rows. */
R3D_API int r3d_forward_video(r3d_plan* plan, const float* seq_dev, const float* param_dev, float* pos_dev,
                      float* trj_dev, float* sum_dev, int32_t frames_out, void* stream);
/* The whole inner step of Trainer.evaluate_core for one video, from pixels (trainer.py:297-353): uv_seq
 * (F + RF - 1, J, 2) float32 pixel keypoints of the edge-padded video, ONE R3D_CAM_F64 row; every frame is ray-encoded
 * once, windows are indexed in place.  flags: R3D_IN_UNDISTORT | R3D_IN_FLIP_TTA.  The _host forms take host pointers
 * (136 bytes per frame cross PCIe instead of RF x 136 per window) and return / complete with results in host memory. */
R3D_API int r3d_forward_video_uv(r3d_plan* plan, const float* uv_seq_dev, const double* cam64_row_dev, int32_t flags,
                                 float* pos_dev, float* trj_dev, float* sum_dev, int32_t frames_out, void* stream);
R3D_API int r3d_forward_video_uv_host(r3d_plan* plan, const float* uv_seq_host, const double* cam64_row_host, int32_t flags,
                                      float* pos_host, float* trj_host, float* sum_host, int32_t frames_out);
R3D_API int r3d_submit_video_uv_host(r3d_plan* plan, const float* uv_seq_host, const double* cam64_row_host, int32_t flags,
                                     float* pos_host, float* trj_host, float* sum_host, int32_t frames_out, uint64_t* ticket);

/* --- test-time flip augmentation: Trainer.evaluate_core with flip_test=True (trainer.py:299-302, 338-353) ---------
 * r3d_plan_set_flip: in_perm[j] = source joint of input joint j in the mirrored copy (the kps_left/kps_right swap of
 *   trainer.py:302), out_perm[s] = output slot whose mirrored prediction lands in slot s (trainer.py:341-342); both
 *   have num_joints entries.
 * r3d_forward_*_tta: every window is lifted twice inside ONE launch sequence (direct, and mirrored: x negated + L/R
 *   swapped in the input stage), the mirrored prediction is un-mirrored and averaged with the direct one in the output
 *   stage (torch.mean of the two), then pos + trj.  Same argument meaning as r3d_forward_rays / r3d_forward_video. */
R3D_API int r3d_plan_set_flip(r3d_plan* plan, const int32_t* in_perm, const int32_t* out_perm);
R3D_API int r3d_forward_rays_tta(r3d_plan* plan, const float* x_dev, const float* param_dev, float* pos_dev, float* trj_dev,
                                 float* sum_dev, int32_t batch, void* stream);
R3D_API int r3d_forward_video_tta(r3d_plan* plan, const float* seq_dev, const float* param_dev, float* pos_dev, float* trj_dev,
                                  float* sum_dev, int32_t frames_out, void* stream);

/* --- standalone camera encode: CameraInfoPacket.get_cam_ray_given_uv in float64 ------------------
 * uv_dev (n_points, 2) float64 -> ray_dev (n_points, 3) float64; one camera (fx, fy, ppx, ppy) and
 * cos/sin of the pitch computed by the caller with libm (math.cos/math.sin, camera.py:333-338) so the
 * result is bit-identical to the reference's numpy arithmetic. */
R3D_API int r3d_ray_encode_f64(const double* uv_dev, double* ray_dev, int64_t n_points, double fx, double fy,
                       double ppx, double ppy, double cos_pitch, double sin_pitch, void* stream);

/* CameraInfoPacket.undistort_point (lib/camera/camera.py:412-421): cv2.undistortPoints(points, K, dist_coeff, P=K) with the
 * 5-coefficient radial/tangential model (k1, k2, p1, p2, k3; dist5_host is a HOST array).  uv_dev/out_dev (n_points, 2)
 * float64 pixels; may alias.  Five fixed-point iterations like OpenCV's default for this overload; bit-identical to
 * opencv-python 4.13 (the reference pins 4.4.0.42, requirements.txt:40). */
R3D_API int r3d_undistort_points_f64(const double* uv_dev, double* out_dev, int64_t n_points, double fx, double fy, double cx,
                                     double cy, const double* dist5_host, void* stream);

/* normalize_screen_coordinates (lib/camera/camera.py:11-18) in float64 for the cfg_rie_* (in_features == 2)
 * input encoding: out = xy / w * 2 - [1, h / w].  xy_dev/out_dev (n_points, 2) float64; may alias. */
R3D_API int r3d_normalize_screen_f64(const double* xy_dev, double* out_dev, int64_t n_points, double w, double h,
                                     void* stream);

/* --- evaluation tail on the device: cam.normalized2world (lib/camera/camera.py:401-410) followed by the error sums of
 * mpjpe / root mpjpe / n_mpjpe / mean_velocity_error (lib/loss/loss.py:12-18, 72-81, 95-104) as evaluate_core applies
 * them (trainer.py:355-395), in float64 like the reference.  pred_dev/target_dev (frames, joints, 3) float32;
 * rn2w_tn2w_dev: 12 doubles = Rn2w row-major then Tn2w, or NULL to stay in the normalised frame; sums_dev: 5 doubles
 * written with [sum ||p-t||, sum over frames of the root error, sum ||s*p-t||, sum of velocity errors, sum of the
 * errors after per-frame Procrustes alignment (p_mpjpe, loss.py:30-69: 3x3 SVD by Jacobi rotations)]; the caller
 * divides by frames*joints, frames, frames*joints, (frames-1)*joints and frames*joints. */
R3D_API int r3d_eval_metrics(const float* pred_dev, const float* target_dev, int32_t frames, int32_t joints,
                             const double* rn2w_tn2w_dev, double* sums_dev, void* stream);

/* --- self tests (used by tests/ and smoke(); run on the device, compare the tensor-core GEMM with
 * the FP32 FFMA GEMM on seeded data).  Returns max |tc - ffma| / max|ffma| in *rel_err. */
R3D_API int r3d_selftest_gemm(int32_t m, int32_t n, int32_t k, int32_t nprob, int32_t precision, int32_t device,
                      double* rel_err, double* ms_tc, double* ms_ffma);

#ifdef __cplusplus
}
#endif
#endif /* RAY3D_B200_H_ */
