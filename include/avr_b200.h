/*
 * avr_b200 — C ABI of the B200 (sm_100a) ray-sampling + volume-compositing path.
 *
 * This is the drop-in boundary.  The reference (yankeesong/adaptive-volume-rendering)
 * has no native layer at all: its hot path is ~30 stock ATen calls in
 * renderers.py:4-119.  Each entry point below replaces the op sequence of one
 * reference function (cited per function); the Python host side that mirrors the
 * reference's own API (`sample_coarse`, `sample_fine`, `sample_depth`,
 * `volume_integral`, `VolumeRenderer`, `AdaptiveVolumeRenderer`) binds these with
 * ctypes and nothing else (see INTEGRATION.md for the stub a maintainer adds).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the function name ends in `_host`;
 *     all floating-point data is fp32, contiguous, row-major;
 *   - buffers are owned by the caller and must stay alive until the work queued on
 *     `stream` has completed; the library never allocates device memory in the
 *     device-pointer entry points, never synchronises, holds no mutable global
 *     state (safe to call from autograd's backward thread);
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - return value: AVR_OK (0) or a negative AVR_ERR_* code; nothing throws;
 *   - "rgbs" is the radiance field's output exactly as the reference slices it:
 *     4 floats per sample in the order (r, g, b, sigma)  [models.py:856-862,
 *     renderers.py:177-178];
 *   - dense layout:  R rays x K samples;  packed layout: `offsets[R+1]` (int64,
 *     offsets[0] == 0, non-decreasing) gives each ray's slice of an S-sample stream;
 *   - `bound_stride`: 1 if `near`/`far` hold one value per ray, 0 if they hold a
 *     single value shared by all rays (the reference passes stride-0 expands of a
 *     1-element tensor, renderers.py:169).
 */
#ifndef AVR_B200_H_
#define AVR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVR_B200_ABI_VERSION 2

typedef void* avr_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define AVR_API __attribute__((visibility("default")))
#else
#define AVR_API
#endif

enum avr_status {
  AVR_OK = 0,
  AVR_ERR_BAD_ARG = -1,    /* null / misaligned pointer, negative size, K out of range */
  AVR_ERR_LAUNCH = -2,     /* cudaPeekAtLastError() after the launch was not cudaSuccess */
  AVR_ERR_NO_DEVICE = -3,  /* current device is not compute capability 10.x */
  AVR_ERR_UNSUPPORTED = -4,/* shape outside what the kernels implement (e.g. K > AVR_MAX_SORT) */
  AVR_ERR_RUNTIME = -5     /* a CUDA runtime call (malloc/memcpy/event) failed; host entry points only */
};

/* Largest per-ray sample count the merge-sort kernel takes (coarse + fine + depth). */
#define AVR_MAX_SORT 1024

AVR_API int avr_abi_version(void);
AVR_API const char* avr_status_string(int status);
/* Last CUDA error string recorded by the calling thread's most recent failing call. */
AVR_API const char* avr_last_cuda_error(void);
/* AVR_OK if the current device can run the kernels (sm_100 family). */
AVR_API int avr_device_check(void);

/* Which composite kernel family a call will use (for tests/bench): 0 = generic
 * thread-per-ray, 1 = TMA-staged blocked scan ("span" kernels). */
AVR_API int avr_composite_plan(int64_t R, int K, const void* rgbs, const void* z);
/* Details of that choice (each out pointer may be NULL): samples per lane, whole rays per
 * warp tile, and how many leading rays the span kernel covers (the remaining R - main_rays
 * tail rays take the generic kernel in a second launch).  Returns 1 / 0 like the above. */
AVR_API int avr_composite_plan_info(int64_t R, int K, const void* rgbs, const void* z,
                                    int* samples_per_lane, int* rays_per_tile, int64_t* main_rays);
/* Force the generic kernels (1) or restore automatic choice (0); process-wide, for tests. */
AVR_API void avr_set_force_generic(int on);
/* Tuning / A-B switches (process-wide).  Each is read from the environment variable of the same
 * name ONCE, at first use (no getenv on any launch path); this call overrides it (unset != 0:
 * back to the built-in default).  Names: AVR_SPAN_L, AVR_SPAN_STAGES, AVR_SPAN_WARPS,
 * AVR_COARSE_PACKED (1 = one warp per ray), AVR_PACKED_SPAN, AVR_IMPORTANCE_GRP,
 * AVR_PACKED_CLASSES, AVR_GRP_G, AVR_IMPORTANCE_BINS, AVR_FIELD_*.  AVR_ERR_BAD_ARG for an
 * unknown name. */
AVR_API int avr_set_option(const char* name, int value, int unset);
/* Which kernel family served the composite calls so far (process-wide counters; tests and
 * bench.py assert that training shapes stay on the span kernels).  Fills out[0..n) and returns
 * AVR_DISPATCH_COUNT. */
enum avr_dispatch {
  AVR_DISPATCH_FWD_SPAN = 0, AVR_DISPATCH_FWD_WRAY, AVR_DISPATCH_FWD_GENERIC, AVR_DISPATCH_FWD_SPAN_PACKED,
  AVR_DISPATCH_BWD_SPAN, AVR_DISPATCH_BWD_WRAY, AVR_DISPATCH_BWD_GENERIC, AVR_DISPATCH_BWD_SPAN_PACKED,
  AVR_DISPATCH_IMPORTANCE_BINS, AVR_DISPATCH_IMPORTANCE_GRP, AVR_DISPATCH_IMPORTANCE_REG, AVR_DISPATCH_IMPORTANCE_SMEM,
  AVR_DISPATCH_COUNT
};
AVR_API int avr_dispatch_counters(int64_t* out, int n);
AVR_API void avr_dispatch_reset(void);

/* ---------------------------------------------------------------- samplers -- */

/* Replaces sample_coarse, renderers.py:12-14:
 *   z[r,j] = near_r + (far_r - near_r) * (j / K)  +  (u[r,j] * (far_r - near_r)) / K
 * evaluated with the reference's operation order (separate mul / add / div, no FMA)
 * so the result is bit-identical to torch-CPU.  u: [R,K] uniforms in [0,1). */
AVR_API int avr_coarse_sample_fwd(const float* near, const float* far, int bound_stride,
                          const float* u, int64_t R, int K, float* z, avr_stream_t stream);

/* Gradient of the above w.r.t. per-ray near / far (AdaptiveVolumeRenderer only,
 * renderers.py:490-494):  d_near[r] = sum_j g_z[r,j] * (1 - (j + u[r,j]) / K),
 * d_far[r] = sum_j g_z[r,j] * ((j + u[r,j]) / K). */
AVR_API int avr_coarse_sample_bwd(const float* g_z, const float* u, int64_t R, int K,
                          float* d_near, float* d_far, avr_stream_t stream);

/* Replaces sample_fine (renderers.py:27-54), sample_depth + clamp (:56-66, :255) and
 * cat + sort (:257-258) in one pass per ray.
 *   weights [R,Kc]   coarse compositing weights (detached by construction)
 *   z_coarse [R,Kc]  coarse depths to merge in (may be NULL when z_sorted is NULL)
 *   u, u2 [R,n_imp]  inverse-CDF draws and in-bin jitter
 *   normals [R,n_depth] N(0,1) draws (may be NULL iff n_depth == 0); the reference's
 *                    sample_depth returns normals * depth_std WITHOUT the depth added;
 *                    reproduced, then clamped to [near_r, far_r]
 * outputs (each may be NULL):
 *   z_fine  [R,n_imp]            unsorted importance samples (what sample_fine returns)
 *   z_sorted[R,Kc+n_imp+n_depth] ascending merge of coarse, fine and depth samples
 *   cdf     [R,Kc+1]             the CDF the search ran on (cdf[r,0] == 0)
 *   idx     [R,n_imp] int32      bin index in [0, Kc] — bit-exact with
 *                                clamp_min(searchsorted(cdf, u, right=True) - 1, 0) */
AVR_API int avr_importance_sample(const float* weights, const float* z_coarse,
                          const float* u, const float* u2, const float* normals,
                          const float* near, const float* far, int bound_stride,
                          int64_t R, int Kc, int n_imp, int n_depth, float depth_std,
                          float* z_fine, float* z_sorted, float* cdf, int32_t* idx,
                          avr_stream_t stream);

/* Per-ray ascending sort (torch.sort, renderers.py:494).  perm (int32 [R,K], may be
 * NULL) receives the source index of each output element (stable), for routing
 * gradients back (d_in[r, perm[r,k]] = d_out[r,k]). */
AVR_API int avr_sort_rays(const float* z_in, int64_t R, int K, float* z_out, int32_t* perm,
                  avr_stream_t stream);

/* Gradient of the sort: d_in[r, perm[r,k]] = g_out[r,k] (every element of d_in is written: perm is
 * a permutation of each row). */
AVR_API int avr_sort_rays_bwd(const float* g_out, const int32_t* perm, int64_t R, int K, float* d_in,
                              avr_stream_t stream);

/* ------------------------------------------------------------- compositing -- */

/* Replaces volume_integral, renderers.py:69-119 (fused: deltas, alpha, exclusive
 * cumprod of (1-alpha)+1e-10, weights, weighted sums, white background).
 *   rgbs [R,K,4], z [R,K]  ->  w [R,K] (may be NULL), rgb [R,3], depth [R]
 * `infinity` is the z paired with the last sample in the depth sum (reference
 * default 1.8, never overridden by its callers). */
AVR_API int avr_composite_fwd(const float* rgbs, const float* z, int64_t R, int K,
                      int white_back, float infinity,
                      float* w, float* rgb, float* depth, avr_stream_t stream);

/* The same with utils.depth_from_world folded in (renderers.py:274-275, 508-509): `depth_affine`
 * [R,2] holds, per ray, (A, B) with  camera depth of (ros + rds * dist) = A * dist + B  (written by
 * avr_world_rays / avr_rays_coarse_sample_points_fwd); `depth` then receives that camera depth instead
 * of the composited distance, and in the backward call `g_depth` is the gradient w.r.t. it.
 * depth_affine == NULL: exactly avr_composite_fwd / avr_composite_bwd. */
AVR_API int avr_composite_fwd_camera(const float* rgbs, const float* z, const float* depth_affine,
                                     int64_t R, int K, int white_back, float infinity,
                                     float* w, float* rgb, float* depth, avr_stream_t stream);
AVR_API int avr_composite_bwd_camera(const float* rgbs, const float* z, const float* depth_affine,
                                     const float* g_rgb, const float* g_depth, const float* g_w,
                                     int64_t R, int K, int white_back, float infinity,
                                     float* d_rgbs, float* d_z, avr_stream_t stream);

/* Forward compositing with the output all-gather fused into the kernel's epilogue (multi-GPU):
 * besides rgb/depth, every ray's (r,g,b,depth) is stored as one float4 into row `row0 + ray` of
 * each of the `n_peers` buffers in `peer_gathered` (HOST array of DEVICE pointers, each a
 * [world*R, 4] fp32 buffer; peers' buffers are peer-mapped over NVLink, e.g. the `buffer_ptrs`
 * of a torch symmetric-memory allocation; include the local buffer to fill the local copy).
 * The caller orders the ranks afterwards (a barrier) before reading the gathered buffers.
 * Returns AVR_ERR_UNSUPPORTED when the batch is not eligible for the span kernels (the
 * caller then composites normally and all-gathers with NCCL). */
AVR_API int avr_composite_fwd_gather(const float* rgbs, const float* z, int64_t R, int K,
                                     int white_back, float infinity,
                                     float* w, float* rgb, float* depth,
                                     void* const* peer_gathered, int n_peers, int64_t row0,
                                     avr_stream_t stream);

/* The fused all-gather WITHOUT a cross-rank barrier.  Same stores as avr_composite_fwd_gather
 * (multicast != 0: `peer_gathered[0]` is an NVSwitch multicast address, n_peers == 1), and when the
 * last CTA of the launch has pushed its rows it writes `value` into word `self_rank` of each of the
 * `n_flag_peers` flag arrays in `peer_flags` (HOST array of DEVICE pointers to uint32 arrays in peer-
 * mapped memory, one array per rank, the local one included), release / system scope.  A consumer
 * runs avr_gather_wait on ITS array before reading its gathered buffer.  `done_counter`: one zeroed
 * device word owned by this rank (used to find the last CTA; left at zero).  `value` is a step
 * number: later launches must pass larger values (serial-number arithmetic, wraps allowed).
 * Double-buffer the gathered rows by step parity: a rank may run one step ahead of its peers. */
AVR_API int avr_composite_fwd_gather_signal(const float* rgbs, const float* z, int64_t R, int K,
                                            int white_back, float infinity,
                                            float* w, float* rgb, float* depth,
                                            void* const* peer_gathered, int n_peers, int multicast, int64_t row0,
                                            uint32_t* const* peer_flags, int n_flag_peers, int self_rank,
                                            uint32_t value, uint32_t* done_counter, avr_stream_t stream);
/* Stream-ordered wait (one tiny kernel) until flags[0..n_sources) have all reached `value`
 * (ld.acquire.sys).  The spin is bounded (~4 s): on expiry *status (device word, may be NULL) is set
 * to 1 and the kernel returns, so a dead peer cannot hang this GPU. */
AVR_API int avr_gather_wait(const uint32_t* flags, int n_sources, uint32_t value, uint32_t* status,
                            avr_stream_t stream);

/* Same, through an NVSwitch MULTICAST mapping of the gathered buffers (e.g. the multicast_ptr of a
 * torch symmetric-memory allocation on an NVLS-capable box): every 32 finished rays leave as one
 * 512-byte multimem.st that the switch replicates into every rank's buffer, this rank's included —
 * 16 bytes per ray cross the GPU's links instead of 16 bytes per ray and peer. */
AVR_API int avr_composite_fwd_gather_multicast(const float* rgbs, const float* z, int64_t R, int K,
                                               int white_back, float infinity,
                                               float* w, float* rgb, float* depth,
                                               void* multicast_gathered, int64_t row0, avr_stream_t stream);

/* All-gather by the copy engines (multi-GPU): copy rows [row0, row0 + rows) of this rank's
 * gathered buffer `peer_gathered[self]` into the same rows of every other rank's buffer
 * (one cudaMemcpyAsync per peer over NVLink: no SM is involved, so the transfer overlaps a
 * persistent kernel — the backward pass — running on another stream).  `peer_gathered` is a
 * HOST array of `n_peers` DEVICE pointers to [world*R, 4] fp32 buffers, peer-mapped (e.g. the
 * buffer_ptrs of a torch symmetric-memory allocation).  Typical use: avr_composite_fwd_gather
 * with only the local buffer as target, then this call on a side stream, then a cross-rank
 * barrier. */
AVR_API int avr_gather_push_rows(void* const* peer_gathered, int n_peers, int self, int64_t row0, int64_t rows,
                                 avr_stream_t stream);

/* Gradient of the above; transmittance is recomputed, nothing is saved by forward.
 *   g_rgb [R,3] (may be NULL = zeros), g_depth [R] (may be NULL), g_w [R,K] (may be NULL)
 *   d_rgbs [R,K,4] (required), d_z [R,K] (may be NULL: VolumeRenderer's z carries no grad) */
AVR_API int avr_composite_bwd(const float* rgbs, const float* z,
                      const float* g_rgb, const float* g_depth, const float* g_w,
                      int64_t R, int K, int white_back, float infinity,
                      float* d_rgbs, float* d_z, avr_stream_t stream);

/* Packed (ragged) variants: ray r owns samples [offsets[r], offsets[r+1]) of an
 * S-sample stream.  rgbs [S,4], z [S], w [S], g_w [S], d_rgbs [S,4], d_z [S].
 * A ray with zero samples composites to background (rgb = white_back, depth = 0). */
AVR_API int avr_composite_fwd_packed(const float* rgbs, const float* z, const int64_t* offsets,
                             int64_t R, int64_t S, int white_back, float infinity,
                             float* w, float* rgb, float* depth, avr_stream_t stream);
AVR_API int avr_composite_bwd_packed(const float* rgbs, const float* z, const int64_t* offsets,
                             const float* g_rgb, const float* g_depth, const float* g_w,
                             int64_t R, int64_t S, int white_back, float infinity,
                             float* d_rgbs, float* d_z, avr_stream_t stream);

/* Packed samplers.  counts are implied by offsets; `u` is laid out like z. */
AVR_API int avr_coarse_sample_fwd_packed(const float* near, const float* far, int bound_stride,
                                 const float* u, const int64_t* offsets, int64_t R, int64_t S,
                                 float* z, avr_stream_t stream);
/* Importance sampling on packed rays: ray r draws n_r = fine_offsets[r+1]-fine_offsets[r]
 * samples from its Kc_r = offsets[r+1]-offsets[r] coarse weights (Kc_r >= 1 wherever
 * n_r > 0), and the merge has Kc_r + n_r entries at offsets[r] + fine_offsets[r].
 * max_coarse / max_fine: upper bounds on Kc_r / n_r over all rays (they size the
 * per-warp shared-memory tables and pick the kernel classes; max_coarse + max_fine <=
 * AVR_MAX_SORT).  PRECONDITION: they really are upper bounds — a ray with more samples than
 * stated is processed at the stated shape (its tail ignored) without an error.
 * cdf (may be NULL): the table each ray's search ran on, Kc_r + 1 entries per ray starting at
 * offsets[r] + r (S + R floats in all); idx (may be NULL): int32 bin index per new sample, laid
 * out like u — bit-exact with clamp_min(searchsorted(cdf_r, u_r, right=True) - 1, 0). */
AVR_API int avr_importance_sample_packed(const float* weights, const float* z_coarse,
                                 const float* u, const float* u2,
                                 const float* near, const float* far, int bound_stride,
                                 const int64_t* offsets, const int64_t* fine_offsets,
                                 int64_t R, int max_coarse, int max_fine,
                                 float* z_fine, float* z_sorted, float* cdf, int32_t* idx,
                                 avr_stream_t stream);

/* ------------------------------------------ ray setup / sample points / depth -- */

/* Sample points handed to the radiance-field callback (renderers.py:171-175, 260-265, 496-500):
 *   pts[r,k,:] = ros[r,:] + rds[r,:] * z[r,k]      (separate mul and add, like the reference)
 *   viewdirs[r,k,:] = rds[r,:]                      (the expand().reshape() copy; may be NULL)
 * ros, rds [R,3]; z [R,K]; pts, viewdirs [R,K,3] — i.e. the callback's (SB, R*K, 3) buffers. */
AVR_API int avr_ray_points_fwd(const float* ros, const float* rds, const float* z, int64_t R, int K,
                               float* pts, float* viewdirs, avr_stream_t stream);
/* Gradient of the above w.r.t. z (AdaptiveVolumeRenderer, where the depths carry grad):
 *   d_z[r,k] = g_pts[r,k,:] . rds[r,:] */
AVR_API int avr_ray_points_bwd(const float* rds, const float* g_pts, int64_t R, int K, float* d_z,
                               avr_stream_t stream);
/* sample_coarse fused with the point generation (renderers.py:169-175 in one pass over the
 * uniforms): writes z [R,K] (bit-identical to avr_coarse_sample_fwd), pts and viewdirs. */
AVR_API int avr_coarse_sample_points_fwd(const float* near, const float* far, int bound_stride, const float* u,
                                         const float* ros, const float* rds, int64_t R, int K,
                                         float* z, float* pts, float* viewdirs, avr_stream_t stream);

/* Packed (ragged) rays: ray r owns samples [offsets[r], offsets[r+1]) of the S-sample streams
 * z [S], pts / viewdirs / g_pts [S,3], d_z [S]. */
AVR_API int avr_ray_points_fwd_packed(const float* ros, const float* rds, const float* z, const int64_t* offsets,
                                      int64_t R, int64_t S, float* pts, float* viewdirs, avr_stream_t stream);
AVR_API int avr_ray_points_bwd_packed(const float* rds, const float* g_pts, const int64_t* offsets, int64_t R,
                                      int64_t S, float* d_z, avr_stream_t stream);

/* utils.get_world_rays (utils.py:309-336): ray origins and unit directions in world coordinates.
 *   x_pix [R,2]; intrinsics [n_cams,3,3], ray r uses camera r / rays_per_cam (the 3x3 inverse of
 *   utils.py:263 is formed inside the kernel);
 *   cam2world [R,4,4] (one pose per ray, as the reference's callers expand it, train.py:83);
 *   ros, rds [R,3];
 *   depth_affine [R,2] (may be NULL): per ray (A, B) with depth_from_world(ros + rds * t) = A t + B,
 *   for avr_composite_fwd_camera. */
AVR_API int avr_world_rays(const float* x_pix, const float* intrinsics, const float* cam2world, int64_t R,
                           int64_t rays_per_cam, float* ros, float* rds, float* depth_affine,
                           avr_stream_t stream);
/* VolumeRenderer's first launch, renderers.py:166-175 in one pass over the uniforms: ray setup
 * (avr_world_rays) + sample_coarse + sample points + view directions.  Outputs ros, rds [R,3],
 * depth_affine [R,2] (may be NULL), z [R,K], pts / viewdirs [R,K,3]; each bit-identical to the
 * separate calls. */
AVR_API int avr_rays_coarse_sample_points_fwd(const float* x_pix, const float* intrinsics, const float* cam2world,
                                              int64_t rays_per_cam, const float* near, const float* far,
                                              int bound_stride, const float* u, int64_t R, int K,
                                              float* ros, float* rds, float* depth_affine,
                                              float* z, float* pts, float* viewdirs, avr_stream_t stream);
/* utils.depth_from_world (utils.py:358-361) of the composited point ros + rds * dist
 * (renderers.py:274-275, 508-509); dist == NULL: `ros` holds the world points themselves and
 * rds is ignored (renderers.py:486).  depth [R] = -(cam2world^-1 [p,1])_z.
 * grad_row [R,3] (may be NULL) receives d depth / d p, the per-ray constant the backward pass
 * needs (depth is affine in the point; the loss's depth penalty, utils.py:374-376, sends
 * gradient through it). */
AVR_API int avr_depth_from_world(const float* ros, const float* rds, const float* dist, const float* cam2world,
                                 int64_t R, float* depth, float* grad_row, avr_stream_t stream);

/* ------------------------------------------ radiance-field front end -- */

/* what NewPixelNeRFNet.forward computes between receiving the
 * renderer's sample points and calling its MLP (models.py:754-826) — world->view transform,
 * positional encoding (models.py:62-71), view direction into the view frame, projection and
 * the bilinear fetch of pixel-aligned encoder features (ConvEncoder.index, models.py:256-279:
 * F.grid_sample, align_corners=True, border padding), concatenated into the MLP's input
 *     out[v*B + b] = [ features (C) | xyz code (3 + 6*num_freqs) | view direction (3) ]
 * for source view v = obj*NS + s and point b of object obj.  The encoder and the MLP stay the
 * reference's torch modules; this replaces the ~25 ATen calls (and ~5x the traffic) between them.
 *
 * The descriptor carries the module state the reference keeps on `self` (poses, focal, c,
 * image_shape / latent_scaling folded into `scale_*`, PositionalEncoding._freqs / ._phases).
 * The feature map is channels-last, (NV, H, W, C) — torch.channels_last strides of the
 * reference's (NV, C, H, W) `encoder.latent`; C % 4 == 0 and C + code width must be even.
 * Given identical inputs every value except the sines is bit-identical to torch-CPU. */
#define AVR_FIELD_MAX_SIN 32
typedef struct avr_field_inputs {
  const float* xyz;       /* (SB, B, 3) world points                                          */
  const float* viewdirs;  /* (SB, B, 3), may be NULL when use_viewdirs == 0                    */
  const float* poses;     /* (NV, 3, 4) world -> view, NV = SB*NS           models.py:705-707 */
  const float* focal;     /* (1 or SB, 2), y already negated                 models.py:721-722 */
  const float* c;         /* (1 or SB, 2) principal point                    models.py:724-733 */
  const float* latent;    /* (NV, H, W, C) channels-last feature map                           */
  float* out;             /* (NV*B, C + code) — (NV*B, C) when features_only                   */
  int64_t B;              /* points per object                                                 */
  int64_t NV;             /* source views in total                                             */
  int NS;                 /* views per object                                                  */
  int focal_per_obj;      /* 0: focal[0] for every view; 1: focal[v / NS]    models.py:805-807 */
  int c_per_obj;          /* same for c                                      models.py:808-810 */
  float scale_x, scale_y; /* latent_scaling / image_shape                    models.py:268-270 */
  int C, H, W;
  int n_sin;              /* 2 * num_freqs rows of the encoding, <= AVR_FIELD_MAX_SIN          */
  float freqs[AVR_FIELD_MAX_SIN];  /* PositionalEncoding._freqs   (f1 f1 f2 f2 ...)  :53-56    */
  float phases[AVR_FIELD_MAX_SIN]; /* PositionalEncoding._phases  (0 pi/2 0 pi/2 ...) :58-60   */
  int include_input;      /* raw coordinates precede the encoding            models.py:69-70   */
  int normalize_z;        /* 1: encode R p, 0: encode R p + t                models.py:766-769 */
  int use_viewdirs;       /* append R d                                      models.py:783-794 */
  int features_only;      /* return_features=True: the features alone        models.py:828-829 */
  /* backward only */
  const float* g_out;     /* (NV*B, row) gradient w.r.t. `out`                                 */
  float* d_latent;        /* (NV, H, W, C) or NULL; overwritten                                */
  float* d_xyz;           /* (SB, B, 3) or NULL; overwritten (summed over the NS views)        */
  float* d_viewdirs;      /* (SB, B, 3) or NULL; overwritten                                   */
} avr_field_inputs;

/* Forward: fills desc->out.  The descriptor is read during the call (host memory). */
AVR_API int avr_field_inputs_fwd(const avr_field_inputs* desc, avr_stream_t stream);
/* Backward of the above: the gradients whose pointers are non-NULL (feature map via vector
 * atomics, summed in registers while consecutive samples stay in one texel cell; points and
 * view directions through the projection, the encoding and the rigid transform). */
AVR_API int avr_field_inputs_bwd(const avr_field_inputs* desc, avr_stream_t stream);

/* ------------------------------------------ the adaptive renderer's LSTM ray march -- */

/* AdaptiveVolumeRenderer's march loop (renderers.py:411-435; the same loop
 * is Raymarcher.forward, :313-351) in ONE launch:
 *     world_0 = ros + rds * init_dist
 *     repeat `steps` times:  v = phi(world_t, return_features=True)      pixel-aligned encoder features,
 *                                                                        models.py:757-761, 803-829
 *                            (h, c) = LSTMCell(v, (h, c))               torch.nn.LSTMCell(C -> 16), gates i,f,g,o
 *                            world_{t+1} = world_t + rds * Linear(h)     out_layer, 16 -> 1
 * The feature fetch is described by an avr_field_inputs descriptor with features_only = 1 and NS == 1
 * (one source view per object — the only case the reference's reshape at :423/:430 admits); its xyz /
 * viewdirs / out pointers are ignored (the points live in registers), B is ignored.  C in {128, 256, 512}.
 * All pointers are device pointers.  fp32 FMAs throughout (no tensor cores: the recurrence would amplify
 * TF32 rounding beyond the 1e-5 bar). */
typedef struct avr_lstm_march {
  const float* ros;        /* (R, 3) ray origins                                        utils.py:315-336 */
  const float* rds;        /* (R, 3) unit ray directions                                                  */
  const float* init_dist;  /* (R)    the N(0.8, 0.05) draw of renderers.py:413 (made by the caller)       */
  int64_t R;
  int64_t rays_per_obj;    /* ray r belongs to object (= source view) r / rays_per_obj                    */
  int steps;               /* raymarch_steps                                                              */
  const float* w_ih;       /* (64, C)  lstm.weight_ih                                   renderers.py:371   */
  const float* w_hh;       /* (64, 16) lstm.weight_hh                                                      */
  const float* b_ih;       /* (64)     lstm.bias_ih                                                        */
  const float* b_hh;       /* (64)     lstm.bias_hh                                                        */
  const float* w_out;      /* (16)     out_layer.weight                                 renderers.py:377   */
  const float* b_out;      /* (1)      out_layer.bias                                                      */
  float* world;            /* (steps + 1, R, 3): world[t] = the point step t starts from; world[steps] is
                              the march's result (forward writes, backward reads)                          */
  /* saved by forward for backward; all four NULL = inference (nothing but `world` is written)            */
  float* feats;            /* (steps, R, C)  the fetched features v_t (the weight-gradient GEMM reads them) */
  float* gates;            /* (steps, R, 64) activated gates i, f, g, o                                     */
  float* cells;            /* (steps, R, 16) c_t                                                            */
  float* hidden;           /* (steps, R, 16) h_t                                                            */
  /* backward only */
  const float* g_world;    /* (R, 3)  dL / d world[steps]                                                   */
  float* d_gates;          /* (steps, R, 64) dL / d (gate pre-activations)                                  */
  float* d_dist;           /* (steps, R)     dL / d (signed distance of step t)                             */
} avr_lstm_march;

/* Forward march.  AVR_ERR_UNSUPPORTED for NS != 1 or a channel count other than 128 / 256 / 512. */
AVR_API int avr_lstm_march_fwd(const avr_field_inputs* field, const avr_lstm_march* march, avr_stream_t stream);
/* Backward through time, including the gradient clamp hook on the hidden state (renderers.py:427-428,
 * clamp to [-10, 10]).  Writes d_gates and d_dist, adds the feature-map gradient into field->d_latent
 * (NOT zeroed here; may be NULL) with vector atomics, and follows the gradient through the sample
 * positions (d features / d point) from step to step.  The parameter gradients are plain GEMMs over
 * the rows this writes (N = steps * R):  d w_ih = d_gates^T feats,  d w_hh = d_gates^T [0; hidden[:-1]],
 * d b_ih = d b_hh = sum_N d_gates,  d w_out = d_dist^T hidden,  d b_out = sum_N d_dist  — left to the
 * caller's BLAS (the Python host side uses torch.matmul). */
AVR_API int avr_lstm_march_bwd(const avr_field_inputs* field, const avr_lstm_march* march, avr_stream_t stream);

/* ------------------------------------------------ host-buffer (end to end) -- */

/* One forward+backward compositing pass over HOST buffers (pinned for full speed):
 * chunks the rays, overlaps H2D copies, the two kernels and D2H copies on the workspace's
 * streams, and returns when every output is on the host.  This is the call `bench.py`
 * times for the end-to-end figure.  Outputs rgb [R,3], depth [R], d_rgbs [R,K,4] and, when `w`
 * is not NULL, the weights [R,K] (the coarse pass needs them, renderers.py:180/252; the fine
 * pass discards them, :270).
 *
 * The workspace owns the device staging buffers and streams (3 slots of `chunk_rays` rays
 * x K samples); it is created once and reused across calls, is not thread-safe, and must
 * be destroyed by the caller.  K of a call must equal the workspace's K.
 * `chunk_rays` <= 0 picks a default (~48 MiB of rgbs per chunk). */
typedef struct avr_host_workspace avr_host_workspace;
AVR_API int avr_host_workspace_create(int K, int64_t chunk_rays, avr_host_workspace** out);
AVR_API int avr_host_workspace_destroy(avr_host_workspace* ws);
AVR_API int avr_composite_fwd_bwd_host(avr_host_workspace* ws,
                                       const float* rgbs, const float* z,
                                       const float* g_rgb, const float* g_depth,
                                       int64_t R, int K, int white_back, float infinity,
                                       float* rgb, float* depth, float* w, float* d_rgbs);

#ifdef __cplusplus
}
#endif
#endif /* AVR_B200_H_ */
