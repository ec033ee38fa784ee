/*
 * difusion_b200 -- C-ABI of the B200-native DI-Fusion per-frame map hot path.
 *
 * This header is the drop-in boundary: every entry point replaces one operator of the reference's
 * torch-extension layer (system/ext/__init__.py:13-42 and third-party torch_scatter) or one fused
 * stretch of system/map.py / system/tracker.py.  The reference interface each one replaces is cited.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch types.  All pointers are DEVICE pointers unless the name
 *     starts with h_ (host).  Arrays are dense row-major.  `stream` is a cudaStream_t passed as void*.
 *   - return 0 on success, a negative DFB_E_* code otherwise; dfb_last_error() gives the message
 *     (thread-local).  No exceptions cross the ABI.  Kernel launch errors ARE checked (the reference
 *     never checks them, imgproc/common.cuh:7-9).
 *   - no hidden allocation: temporaries live in a caller-provided workspace (`ws`, `ws_bytes`), sized by the
 *     companion *_ws_bytes() function.  Calls are asynchronous on `stream`; data-dependent output sizes
 *     are written to device counters the caller reads.
 *   - built for sm_100a only; there is no CPU fallback.
 */
#ifndef DIFUSION_B200_H_
#define DIFUSION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_OK 0
#define DFB_E_INVALID (-1)   /* bad argument */
#define DFB_E_WORKSPACE (-2) /* workspace too small */
#define DFB_E_CUDA (-3)      /* CUDA runtime / launch error */
#define DFB_E_CAPACITY (-4)  /* a fixed-capacity structure overflowed */

#define DFB_LATENT_DIM 29
#define DFB_DIV_IEEE 0  /* x / vs   : torch CPU semantics (the parity oracle, BASELINE config 1) */
#define DFB_DIV_RECIP 1 /* x * (1/vs): torch CUDA semantics for division by a Python scalar    */

int dfb_version(void);
const char* dfb_last_error(void);
/* SM count / name of the current device (so hosts can size grids); returns 0 or DFB_E_CUDA. */
int dfb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Network weights.  Host code folds weight-norm / batch-norm and packs the blobs
 * (nerf-fusion_b200/weights.py documents the layout); sizes are validated here.
 * Replaces: network/utility.py:22-58 load_model + the per-forward weight-norm recompute
 * (network/utility.py:211-220) and BatchNorm eval (utils/pt_util.py:193-206).
 * ---------------------------------------------------------------------------------------------- */
size_t dfb_decoder_blob_floats(void);
size_t dfb_encoder_blob_floats(void);
/* Decoder engine used by dfb_decoder_forward / dfb_get_sdf / dfb_sdf_hg / dfb_decode_cubes:
 * 1 = tcgen05 (FP16 operands, FP32 accumulation in TMEM; default), 0 = FP32 CUDA cores (exact-precision path).
 * Also settable with the environment variable DFB_DECODER_ENGINE before the first call. */
int dfb_set_decoder_engine(int engine);
int dfb_get_decoder_engine(void);
/* Frame pipelining (no reference counterpart; the reference tracks frames strictly one after the other, tracker.py:75-134):
 * SMs the fused Gauss-Newton evaluation kernel leaves free (its grid = SM count - n) so that the NEXT frame's front end,
 * queued on another stream, runs beside the pose solve instead of time-slicing with it.  0 = use every SM (default). */
int dfb_set_gn_reserved_sms(int n);
/* Same switch for the encoder (dfb_integrate_commit, dfb_encoder_forward); env DFB_ENCODER_ENGINE.  Default 1: the
 * tcgen05 encoder splits weights and the inputs of layers 0..2 into FP16 hi + lo parts, which keeps latents within
 * 1.8e-4 relative of the FP32 reference on the golden keyframe (tolerance 1e-3); 0 = FP32 CUDA cores (6e-7). */
int dfb_set_encoder_engine(int engine);
int dfb_get_encoder_engine(void);

/* ------------------------------------------------------------------------------------------------
 * Stage 1 -- image / point-cloud preprocessing
 * ---------------------------------------------------------------------------------------------- */
/* system.ext.unproject_depth (imgproc.cpp:3, imgproc.cu:5-44).  depth (H,W) f32, NaN = invalid ->
 * pc (H,W,3) f32.  Invalid pixels get NaN in ALL three channels (the reference leaves 1,2 uninitialised). */
/* Frame ingest: replaces dataset/production/icl_nuim.py:110-114 (cv2 decode -> float32 -> .cuda() -> / 5000, / 255.) and the
 * clipping of main.py:56-57.  depth_raw uint16[H,W] and color_raw uint8[H,W,3] are DEVICE copies of the decoded PNGs;
 * depth_out[i] = depth_raw[i] / depth_scale (NaN outside [cut_min, cut_max] when cut_min < cut_max),
 * rgb_out = color_raw / 255 with an optional BGR -> RGB swap.  div_mode: DFB_DIV_RECIP multiplies by the fp32 reciprocal
 * like torch CUDA's tensor / scalar (what the reference executes), DFB_DIV_IEEE divides (numpy / torch CPU).
 * Either pair (depth_raw, depth_out) or (color_raw, rgb_out) may be NULL. */
int dfb_ingest_frame(const uint16_t* depth_raw, const uint8_t* color_raw, int H, int W, float depth_scale, int div_mode,
                     float cut_min, float cut_max, int bgr, float* depth_out, float* rgb_out, void* stream);

/* Isometry @ points and rotation @ normals (utils/motion_util.py:323-328, main.py:83-84): out[i] = R xyz[i] (+ t), fp32,
 * h_R row-major 3x3 on the host, h_t may be NULL (rotation only). */
int dfb_transform_points(const float* xyz, int n, const float* h_R, const float* h_t, float* out, void* stream);

int dfb_unproject_depth(const float* depth, int H, int W, float fx, float fy, float cx, float cy, float* pc,
                        void* stream);

/* system.ext.remove_radius_outlier (pcproc.cpp:3-7, pcproc.cu:160-187): mask[i] = (the nb_points-th smallest
 * squared distance from point i, self included) < radius^2.  pc (n,4) f32 with w = 0. */
size_t dfb_pcproc_ws_bytes(int n);
int dfb_remove_radius_outlier(const float* pc4, int n, int nb_points, float radius, uint8_t* mask, void* ws,
                              size_t ws_bytes, void* stream);
/* system.ext.estimate_normals (pcproc.cpp:9-14, pcproc.cu:189-210): PCA normal over the <= max_nn-1 nearest
 * neighbours within radius (>= 5 required, else NaN), oriented towards cam.  max_nn <= 32. */
int dfb_estimate_normals(const float* pc4, int n, int max_nn, float radius, const float* h_cam_xyz,
                         float* normals, void* ws, size_t ws_bytes, void* stream);

/* torch_scatter.scatter_mean(src, index, dim=0) (tracker.py:22-23).  src (n,d) f32, index (n,) i64 in
 * [0, n_out).  out (n_out,d) f32; empty groups give 0.  Deterministic: rows are summed in ascending row
 * order.  */
size_t dfb_scatter_mean_ws_bytes(int n, int n_out);
int dfb_scatter_mean(const float* src, const int64_t* index, int n, int d, int n_out, float* out, void* ws,
                     size_t ws_bytes, void* stream);

/* tracker.point_box_filter (tracker.py:14-24) fused: bounding box, cell keys, unique (ascending key),
 * per-cell mean of points and normals.  out_points/out_normals hold up to n rows; *d_n_out (device int32)
 * receives the number of cells.  div_mode selects how (p - min)/voxel_size is evaluated (DFB_DIV_*). */
size_t dfb_box_filter_ws_bytes(int n);
int dfb_point_box_filter(const float* points, const float* normals, int n, float voxel_size, int div_mode,
                         float* out_points, float* out_normals, int32_t* d_n_out, void* ws, size_t ws_bytes,
                         void* stream);

/* The geometry half of SDFTracker.track_camera (tracker.py:89-120) as ONE asynchronous call with no host
 * synchronisation: nearest x0.5 subsample of the (H,W) depth (NaN = invalid), unproject with halved intrinsics, drop
 * invalid pixels, radius-outlier filter, PCA normals, drop NaN normals, 2 cm box filter.  Rows keep their order (like the
 * reference's boolean-mask indexing), intermediate counts stay on the device.  When outlier_radius <= normal_radius / 2
 * (the reference's 0.05 / 0.1 m) the cloud stays uncompacted until the box filter and ONE uniform grid (cells = the outlier
 * radius) serves both the radius count and the kNN search; otherwise each stage compacts and builds its own grid.  Same rows,
 * bit for bit, as the op-by-op chain either way.  out_points/out_normals hold up to (H/2)*(W/2) rows; *d_n_out (device int32)
 * receives the number of rows (-1: box-filter key range overflow). */
size_t dfb_preprocess_ws_bytes(int H, int W);
int dfb_preprocess_frame(const float* depth, int H, int W, float fx, float fy, float cx, float cy, int nb_points,
                         float outlier_radius, int max_nn, float normal_radius, const float* h_cam_xyz, float box_voxel,
                         int div_mode, float* out_points, float* out_normals, int32_t* d_n_out, void* ws, size_t ws_bytes,
                         void* stream);

/* system.ext.groupby_sum (indexing.cpp:3-4, indexing.cu:59-71,89-109): sum (C,L) f32 and count (C,) i32 of
 * values (n,L) grouped by indices (n,) i64.  Outputs are zeroed here.  Like the reference kernel, `count` is
 * incremented once per element, i.e. it holds L x (number of rows in the group). */
int dfb_groupby_sum(const float* values, const int64_t* indices, int n, int L, int C, float* sum, int32_t* count,
                    void* stream);

/* system.ext.gradient_xy (imgproc.cpp:21, photometric.cu:3-22,79-93). (H,W) -> (H,W,2), NaN border. */
int dfb_gradient_xy(const float* intensity, int H, int W, float* grad, void* stream);
/* Image half of the tracker front end in two launches: depth clipping (main.py:56-57; cut = 0 disables it), intensity =
 * torch.mean(rgb, -1) (tracker.py:84), the 3-level pyramid of tracker.py:42-57 (F.interpolate bilinear align_corners=True for
 * intensity, nearest for depth) and gradient_xy of every level.  Outputs: I0,D0 [H,W], I1,D1 [H/2,W/2], I2,D2 [H/4,W/4],
 * G0..G2 [h,w,2].  Bit-identical to the torch CUDA kernels the reference calls (test_frame_images_equals_torch). */
int dfb_frame_images(const float* rgb, const float* depth, int H, int W, float cut_near, float cut_far, int cut, float* I0, float* D0,
                     float* I1, float* D1, float* I2, float* D2, float* G0, float* G1, float* G2, void* stream);
/* system.ext.rgb_odometry (imgproc.cpp:14-20, photometric.cu:24-77,95-138).  f (H,W) and, if J != NULL,
 * J (H,W,6); NaN where invalid.  h_intr = fx,fy,cx,cy; h_krkinv row-major 3x3; h_kt 3. */
int dfb_rgb_odometry(const float* prev_I, const float* prev_D, const float* cur_I, const float* cur_D,
                     const float* cur_dIdxy, int H, int W, const float* h_intr, const float* h_krkinv,
                     const float* h_kt, float min_grad_scale, float max_depth_delta, float* f, float* J,
                     void* stream);
/* Fused tracker.compute_rgb_Hg (tracker.py:136-177): the same per-pixel residual/Jacobian reduced in-kernel to
 * out[0..35] = sum w J J^T (with the reference's sign flip J := -J), out[36..41] = sum w f J, out[42] = sum w f^2,
 * out[43] = number of valid pixels (doubles, unnormalised).  `out44` must hold 80 doubles: [0..43] is the result,
 * the rest is scratch; all of it is zeroed here.  robust: 0 none, 1 huber, 2 tukey. */
int dfb_rgb_hg(const float* prev_I, const float* prev_D, const float* cur_I, const float* cur_D,
               const float* cur_dIdxy, int H, int W, const float* h_intr, const float* h_krkinv, const float* h_kt,
               float min_grad_scale, float max_depth_delta, int robust, float robust_k, int compute_J, double* out44,
               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stages 2+3 -- integrate_keyframe (system/map.py:341-453, do_optimize = False)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t nx, ny, nz;          /* grid size (map.py:178)                                   */
  float bound_min[3];          /* map.py:182                                               */
  float voxel_size;            /* map.py:177                                               */
  int32_t div_mode;            /* DFB_DIV_*                                                */
  int32_t prune_min_vox_obs;   /* fusion-lr-kt.yaml:32, strict '>'  (map.py:376)           */
  float ignore_count_th;       /* :33, strict '>' (map.py:572,632)                         */
  float encoder_count_th;      /* :34, strict '<' (map.py:410)                             */
} dfb_map_params;

/* Persistent per-map scratch (all-zero between calls; the kernels restore the zeros they disturb):
 *   grid_count  int32[G]            per-cell point counter
 *   grid_bits   uint32[ceil(G/32)]  allocation bitmap
 *   acc         int64[cap*29], acc_n int32[cap], touched int32[cap]   per-slot encoder accumulators: 64-bit FIXED-POINT sums in
 *               units of 2^-34 (integer adds commute: the scatter-add of map.py:446-449 gives the same bits for any order of the
 *               atomics and any sharding of the samples, unlike the float atomics of indexing.cu:59-71)                       */
size_t dfb_integrate_ws_bytes(int n, int64_t n_cells);

/* Phase A (map.py:367-388 up to the allocation count): voxel ids, count-prune (unq_mask, the function's return
 * value map.py:520), mark unseen home voxels and their 6 clamped face neighbours, count them.
 * *d_n_new (device int32) receives how many slots phase B will allocate, so the host can grow its buffers
 * exactly like map.py:263-285 before calling phase B. */
int dfb_integrate_plan(const dfb_map_params* h_params, const float* xyz, const float* normal, int n,
                       const int64_t* indexer, int32_t* grid_count, uint32_t* grid_bits, uint8_t* unq_mask,
                       int32_t* d_n_new, void* ws, size_t ws_bytes, void* stream);
/* Phase B (map.py:388-453): assign slots n_occupied.. in ascending voxel-id order, gather (point, offset)
 * samples of candidate voxels, run the encoder MLP, scatter-add per voxel, running-mean update, mark
 * updated slots.  `ws` must be the workspace phase A filled, unq_mask phase A's mask, n_new the value read from
 * *d_n_new.  capacity = rows of the per-slot arrays (DFB_E_CAPACITY if n_occupied + n_new exceeds it).
 * d_stats (device int32[4]) receives {n_samples, n_voxels_updated, n_allocated, 0}. */
int dfb_integrate_commit(const dfb_map_params* h_params, const float* normal, const uint8_t* unq_mask, int n,
                         int64_t* indexer, float* latent_vecs, int64_t* latent_vecs_pos, float* voxel_obs_count,
                         uint8_t* updated_flag, int64_t n_occupied, int64_t capacity, int32_t n_new, uint32_t* grid_bits,
                         int64_t* acc, int32_t* acc_n, int32_t* touched, const float* encoder_blob, int32_t* d_stats,
                         void* ws, size_t ws_bytes, void* stream);

/* network encoder forward on explicit inputs (map.py:446-447 -> di_encoder.py:26-30): x (m,6) -> (m,29). */
int dfb_encoder_forward(const float* x, int m, const float* encoder_blob, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 4 -- decoder queries
 * ---------------------------------------------------------------------------------------------- */
/* net_util.forward_model on explicit rows (utility.py:61 -> di_decoder.py:55-86): x (n,32) = [latent, xyz]
 * -> sdf (n,), std (n,). */
int dfb_decoder_forward(const float* x, int n, const float* decoder_blob, float* sdf, float* std, void* stream);

/* DenseIndexedMap.get_sdf (map.py:560-580), un-compacted: for every xyz row writes valid[i] (u8) and, where
 * valid, sdf[i], std[i].  If g_sdf/g_std/grad_xyz are non-NULL it also back-propagates
 * grad_xyz[i] = d(g_sdf[i]*sdf_i + g_std[i]*std_i)/d xyz_i  (what autograd does on the reference graph). */
int dfb_get_sdf(const dfb_map_params* h_params, const float* xyz, int n, const int64_t* indexer,
                const float* latent_vecs, const float* voxel_obs_count, const float* decoder_blob, float* sdf,
                float* std, uint8_t* valid, const float* g_sdf, const float* g_std, float* grad_xyz, void* stream);

/* SDFTracker.compute_sdf_Hg (tracker.py:179-223) fused: transform, map lookup, decoder forward + reverse pass,
 * Jacobian, robust weights, reduction.  obs_xyz (n,3) camera space.  h_pose = 33 floats:
 * R_total(9), t_total(3), R_delta(9), t_delta(3), R_last(9) (row-major, fp32 casts of the f64 poses).
 * out44 (device doubles, 80 allocated: 44 results + scratch, zeroed here): [0..35] sum w J J^T, [36..41] sum w r J, [42] sum w r^2, [43] M' (valid
 * count) -- unnormalised; the host divides by M' like tracker.py:215-223.  robust: 0 none, 1 huber, 2 tukey. */
int dfb_sdf_hg(const dfb_map_params* h_params, const float* obs_xyz, int n, const float* h_pose,
               const int64_t* indexer, const float* latent_vecs, const float* voxel_obs_count,
               const float* decoder_blob, int robust, float robust_k, int compute_J, double* out44, void* stream);

/* SDFTracker.gauss_newton (tracker.py:225-288) as ONE host call: for every group of the iteration config, up to n_iter
 * Gauss-Newton steps plus one evaluation-only pass, each evaluating the fused SDF term and/or the fused photometric
 * term.  The loop state lives on the device: a one-warp step (the tail of the evaluation's last block) scales and adds the
 * terms, solves the 6x6 system and updates the pose in float64, rolls a step back when its energy rises (which ends the
 * group, tracker.py:269-271) and publishes the pose of the next evaluation, so nothing is read back between
 * iterations; the host keeps one evaluation of look-ahead queued and only polls a 16-byte record per step (the 96-byte pose follows when a group ends).
 * obs_xyz holds n rows; if d_n (device int32, may be NULL) is given, only the first min(n, max(*d_n, 0)) rows are used, so
 * the caller need not read the row count of dfb_preprocess_frame back before queueing the solve.
 * Poses are 12 doubles: R row-major (9) then t (3).  h_delta_pose is in/out (initial guess -> result; untouched on error).
 * d_scratch: device, DFB_GN_SCRATCH_DOUBLES doubles.  h_pinned: PINNED (device-visible) host memory,
 * DFB_GN_PINNED_DOUBLES doubles.
 * h_stats[8] = {last value of the iteration counter (tracker.py:281), #sdf evaluations, #rgb evaluations, status, ...}.
 * If h_stats[4] == 0x54494d45 on entry, the SDF-term launches are timed with CUDA events on `stream` and
 * h_stats[4..6] return {microseconds, valid queries with reverse pass, valid queries forward-only}.
 * h_stats[7] = kernels launched by the call. */
#define DFB_GN_SCRATCH_DOUBLES 160
#define DFB_GN_PINNED_DOUBLES 64
typedef struct {
  int32_t n_groups;              /* <= 8 */
  int32_t n_iter[8];             /* iter_config[i]["n"]                                  */
  int32_t use_sdf[8];            /* group has an ['sdf'] term                             */
  int32_t rgb_level[8];          /* pyramid level of its ['rgb', level] term, -1 = none   */
  int32_t sdf_robust;  float sdf_robust_k;                       /* 0 none, 1 huber, 2 tukey */
  int32_t rgb_robust;  float rgb_robust_k;
  float rgb_weight, rgb_min_grad_scale, rgb_max_depth_delta;     /* fusion-lr-kt.yaml:51-56  */
} dfb_gn_config;
typedef struct {
  const float* prev_I; const float* prev_D; const float* cur_I; const float* cur_D; const float* cur_G;
  int32_t H, W;
} dfb_rgb_level;
int dfb_gauss_newton(const dfb_map_params* h_params, const dfb_gn_config* h_cfg, const float* obs_xyz, int n, const int32_t* d_n,
                     const int64_t* indexer, const float* latent_vecs, const float* voxel_obs_count,
                     const float* decoder_blob, const dfb_rgb_level* h_levels, const double* h_intr,
                     const double* h_last_pose, double* h_delta_pose, double* d_scratch, double* h_pinned,
                     int32_t* h_stats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 5 -- meshing
 * ---------------------------------------------------------------------------------------------- */
/* do_meshing sampling (map.py:637-688, fast=True): per voxel slot occ[b], decode an r^3 lattice, trilinear
 * x2 (align_corners), re-decode the |sdf| < refine_band samples, negate.  cube_sdf/std (B,2r,2r,2r). */
size_t dfb_decode_cubes_ws_bytes(int B, int r);
int dfb_decode_cubes(const float* latent_vecs, const int64_t* occ, int B, int r, float refine_band,
                     const float* decoder_blob, float* cube_sdf, float* cube_std, void* ws, size_t ws_bytes,
                     void* stream);

/* system.ext.marching_cubes_interp (mc.cpp:3-12, mc_interp_kernel.cu:202-382).  indexer (nx,ny,nz) i64,
 * valid_blocks (U,) i64, vec_batch_mapping (n_map,) i32, cubes (B,2r,2r,2r).  Writes up to max_tri triangles
 * (tri (T,3,3) f32 voxel units, flat_id (T,) i64, tri_std (T,3) f32) and the TOTAL count found to *d_n_tri
 * (device int32; may exceed max_tri, in which case the output is truncated like mc_interp_kernel.cu:375-379). */
int dfb_marching_cubes(const int64_t* indexer, int nx, int ny, int nz, const int64_t* valid_blocks, int U,
                       const int32_t* vec_batch_mapping, int n_map, const float* cube_sdf, const float* cube_std,
                       int B, int r, float max_std, int max_tri, float* tri, int64_t* flat_id, float* tri_std,
                       int32_t* d_n_tri, void* stream);

/* One Adam iteration of the latent optimiser, OptimizeProcess.do_optimize (map.py:81-113; disabled by the target config,
 * do_optimize = False): decoder forward + reverse pass w.r.t. the 29 latent inputs of every sample, Gaussian negative
 * log-likelihood of the clamped targets (map.py:88-96) averaged over n_samples, gradient scatter-added per latent row, then
 * torch.optim.Adam's update (lr, betas 0.9 / 0.999, eps 1e-8; `step` is 1-based) with the code regulariser's gradient
 * reg_scale * x / ||x|| (reg_scale = code_reg_lambda * chunks / n_samples, 0 = off).  latents (n_rows,29) is updated in place;
 * grad / adam_m / adam_v are (n_rows,29) buffers, zero before the first step (grad is left zero).  FP32 CUDA-core engine. */
int dfb_latent_adam_step(float* latents, int n_rows, const int64_t* inv, const float* rel_xyz, const float* gt_sdf, int n_samples,
                         const float* decoder_blob, float* grad, float* adam_m, float* adam_v, int step, float lr, float reg_scale,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sharded map (BASELINE config 5; new functionality, the reference has no multi-GPU map).
 * integrate_keyframe of system/map.py:341-453 over the union of all ranks' points, voxel ids partitioned in 8^3 bricks dealt
 * round-robin to the ranks.  Records travel by PEER STORES from the producing kernel into the owner's receive buffer
 * (`peer_*[d]` = base of this rank's segment inside rank d's inbox; for d == rank it is the local inbox).  The host side
 * (nerf-fusion_b200/sharded.py) owns the memory, maps the peers (dfb_peer_* below, CUDA IPC) and runs a stream-ordered
 * barrier between the five phases; see csrc/sharded.cu for the protocol.  All pointers are device pointers.
 * ---------------------------------------------------------------------------------------------- */
#define DFB_SHARD_MAX_WORLD 8
typedef struct {
  int32_t nx, ny, nz;
  float bound_min[3];
  float voxel_size;
  int32_t div_mode, prune_min_vox_obs;
  float encoder_count_th;
  int32_t world, rank;
  /* owned state */
  int32_t* indexer_local;      /* dfb_shard_local_cells() ints, -1 = unallocated, else slot */
  float* latent_vecs;          /* (capacity, 29), zero for unused slots */
  int32_t* latent_vecs_pos;    /* (capacity,) linear voxel id of a slot */
  float* voxel_obs_count;      /* (capacity,) */
  int32_t capacity;
  uint32_t* cand_bits;         /* ceil(nx ny nz / 32) words: candidate voxels (allocated, obs_count < encoder_count_th), all ranks' */
  int32_t* grid_count;         /* dfb_shard_local_cells() ints, zero between keyframes */
  int64_t* acc;                /* (capacity, 29) fixed-point sums (2^-34 units), zero between keyframes */
  int32_t* acc_n;              /* (capacity,) zero between keyframes */
  int32_t* touched;            /* (capacity,) scratch */
  int32_t* counters;           /* dfb_shard_counter_ints() ints, zero-initialised once; [3*8+3] = n_occupied */
  int32_t* delta_list;         /* (delta_cap,) */
  int32_t* next_delta;         /* (delta_cap,) */
  int32_t* n_next_delta;       /* 1 int, zero-initialised */
  int32_t delta_cap;
  /* receive buffers of THIS rank: `world` segments each, segment s written by rank s */
  void* pts_inbox;    int32_t pts_cap; int32_t* pts_count;   /* 32-byte point records; counts: `world` ints */
  int32_t* ids_inbox; int32_t ids_cap; int32_t* ids_count;   /* 4-byte voxel ids (allocation requests) */
  void* smp_inbox;    int32_t smp_cap; int32_t* smp_count;   /* 32-byte samples {id -> slot, rel[3], n[3]} */
  int32_t* dlt_inbox; int32_t dlt_cap; int32_t* dlt_count;   /* 4-byte candidate-set deltas (id << 1 | removed) */
  /* where THIS rank writes: its segment inside rank d's buffers, and rank d's count slot for this rank */
  void* peer_pts[DFB_SHARD_MAX_WORLD]; int32_t* peer_pts_count[DFB_SHARD_MAX_WORLD];
  void* peer_ids[DFB_SHARD_MAX_WORLD]; int32_t* peer_ids_count[DFB_SHARD_MAX_WORLD];
  void* peer_smp[DFB_SHARD_MAX_WORLD]; int32_t* peer_smp_count[DFB_SHARD_MAX_WORLD];
  void* peer_dlt[DFB_SHARD_MAX_WORLD]; int32_t* peer_dlt_count[DFB_SHARD_MAX_WORLD];
  /* barrier flags: `flags` = this rank's `world` epoch slots (slot s written by rank s), peer_flags[d] = this rank's slot at rank d */
  int32_t* flags; int32_t* peer_flags[DFB_SHARD_MAX_WORLD];
  int32_t fuse_publish;        /* 1: the per-pair record counts are published by dfb_shard_barrier (one launch less per phase) */
} dfb_shard;

int dfb_shard_counter_ints(void);
int64_t dfb_shard_local_cells(int nx, int ny, int nz, int world);
/* phase 1: this rank's share of the keyframe (n <= pts_cap points) -> owners of the home voxels.            [barrier] */
int dfb_shard_phase1(const dfb_shard* S, const float* xyz, const float* normal, int n, void* stream);
/* phase 2: count, prune (map.py:373-379), allocate home voxels + face neighbours (:382-388), request remote ones. [barrier] */
int dfb_shard_phase2(const dfb_shard* S, void* stream);
/* phase 3: allocate requested ids, broadcast this rank's candidate-set deltas.                              [barrier] */
int dfb_shard_phase3(const dfb_shard* S, void* stream);
/* phase 4: apply deltas, build the (point, offset) samples (map.py:390-436), push them to their voxels' owners. [barrier] */
int dfb_shard_phase4(const dfb_shard* S, void* stream);
/* phase 5: id -> slot, encoder on the receive buffer, running mean (map.py:446-452).  d_stats: 8 + 2*8 device ints
 * {points received, samples received, voxels allocated, error bits, n_occupied, voxels updated, -, -, points sent to rank d..,
 * samples sent to rank d..}. */
int dfb_shard_phase5(const dfb_shard* S, const float* encoder_blob, int32_t* d_stats, void* stream);

/* Barrier between two phases, as a kernel: every rank stores `epoch` (monotonically increasing, > 0) into its slot of every
 * peer's flag array (peer stores, after a system-scope fence that orders the phase's record stores before it) and waits until all
 * of its own slots have reached `epoch`.  With S->fuse_publish it first publishes the record counts of the phase that just ran
 * (`after_phase` = 1, 2 or 4; 3 publishes nothing).  Each rank runs this on its OWN GPU, so the waits cannot starve one another; the wait is
 * bounded (~2 s) and traps instead of hanging.  An alternative to a 4-byte NCCL all-reduce (a few microseconds instead of ~20). */
int dfb_shard_barrier(const dfb_shard* S, int epoch, int after_phase, void* stream);

/* Peer-mappable device memory (cudaMalloc + CUDA IPC): alloc returns the pointer and a 64-byte handle another process on the
 * same node opens with dfb_peer_open (peer access over NVLink is enabled lazily by the driver). */
int dfb_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int dfb_peer_open(const unsigned char* handle64, void** ptr);
int dfb_peer_close(void* ptr);
int dfb_peer_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* DIFUSION_B200_H_ */
