// Internal launcher declarations (host side).  Each returns an avr_status.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "avr_b200.h"

namespace avr {

// runtime.cu — switches (environment read once; avr_set_option overrides), SM count, counters
enum Opt {
  OPT_SPAN_L, OPT_SPAN_STAGES, OPT_SPAN_WARPS, OPT_COARSE_PACKED_RAY,
  OPT_PACKED_SPAN, OPT_IMPORTANCE_GRP, OPT_PACKED_CLASSES, OPT_GRP_G,
  OPT_FIELD_NOCACHE, OPT_FIELD_BWD_SPLIT, OPT_FIELD_SHARE_POINT, OPT_FIELD_BWD_PREFETCH,
  OPT_FIELD_STAGE, OPT_IMPORTANCE_BINS, OPT_FIELD_BWD_RING, OPT_HOST_SLOTS, OPT_HOST_CHUNK_MIB,
  OPT_FIELD_BWD_ASYNC,
  OPT_COUNT
};
int option(Opt o, int dflt);
int num_sms();                  // SMs of the current device (cached)
void count_dispatch(int which); // AVR_DISPATCH_* of avr_b200.h

// composite_generic.cu — any shape, dense (offsets == nullptr) or packed
// depth_affine (nullable, [R,2]): `depth` / `g_depth` are the camera depth A*dist + B (avr_common.cuh)
int launch_composite_fwd_generic(const float* rgbs, const float* z, const int64_t* offsets, int64_t R,
                                 int K, int white_back, float infinity, float* w, float* rgb,
                                 float* depth, cudaStream_t stream, const float* depth_affine = nullptr);
int launch_composite_bwd_generic(const float* rgbs, const float* z, const int64_t* offsets,
                                 const float* g_rgb, const float* g_depth, const float* g_w, int64_t R,
                                 int K, int white_back, float infinity, float* d_rgbs, float* d_z,
                                 cudaStream_t stream, const float* depth_affine = nullptr);

// composite_wray.cu — warp per ray, coalesced; any shape, dense (offsets == nullptr) or packed
int launch_composite_fwd_wray(const float* rgbs, const float* z, const int64_t* offsets, int64_t R, int K,
                              int white_back, float infinity, float* w, float* rgb, float* depth,
                              cudaStream_t stream, const float* depth_affine = nullptr);
int launch_composite_bwd_wray(const float* rgbs, const float* z, const int64_t* offsets, const float* g_rgb,
                              const float* g_depth, const float* g_w, int64_t R, int K, int white_back,
                              float infinity, float* d_rgbs, float* d_z, cudaStream_t stream,
                              const float* depth_affine = nullptr);

// composite_span.cu — dense, TMA-staged blocked scan.  `span_plan` says whether a
// shape is eligible and how many leading rays the span kernel covers (the caller
// runs the generic kernel on the remaining tail rays).
struct SpanPlan {
  int L;              // samples per lane (odd)
  int rays_per_tile;  // whole rays per warp tile
  int64_t main_rays;  // rays covered by full tiles
};
bool span_plan(int64_t R, int K, const void* rgbs, const void* z, SpanPlan* plan);
// completion signal of a fused-gather launch (see SpanArgs::signal_flags)
struct GatherSignal {
  uint32_t* const* flags;   // per peer: that peer's flag array (device pointers, peer-mapped)
  int n;                    // peers to signal
  int slot;                 // word written in each array (this rank's index)
  uint32_t value;
  unsigned* done_counter;   // local device word, zero between launches
};
int launch_composite_fwd_span(const SpanPlan& plan, const float* rgbs, const float* z, int K,
                              int white_back, float infinity, float* w, float* rgb, float* depth,
                              cudaStream_t stream, void* const* peers = nullptr, int n_peers = 0,
                              int64_t peer_row0 = 0, bool multicast = false,
                              const GatherSignal* signal = nullptr, const float* depth_affine = nullptr);
// runtime.cu: wait until flags[0..n) >= value (acquire, system scope); bounded spin
int launch_gather_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* status, cudaStream_t stream);
int launch_composite_bwd_span(const SpanPlan& plan, const float* rgbs, const float* z, const float* g_rgb,
                              const float* g_depth, int K, int white_back, float infinity, float* d_rgbs,
                              float* d_z /* nullable; needs K > plan.L */, cudaStream_t stream,
                              const float* depth_affine = nullptr);

// composite_span_packed.cu — packed layout, TMA-staged whole-ray tiles packed on the fly
bool span_packed_eligible(const void* rgbs, const void* z, const void* w_or_null, const void* d_rgbs_or_null);
int launch_composite_fwd_span_packed(const float* rgbs, const float* z, const int64_t* offsets, int64_t R,
                                     int64_t S, int white_back, float infinity, float* w, float* rgb,
                                     float* depth, cudaStream_t stream);
int launch_composite_bwd_span_packed(const float* rgbs, const float* z, const int64_t* offsets,
                                     const float* g_rgb, const float* g_depth, int64_t R, int64_t S,
                                     int white_back, float infinity, float* d_rgbs, cudaStream_t stream);

// samplers.cu
int launch_coarse_fwd(const float* near, const float* far, int bound_stride, const float* u,
                      const int64_t* offsets, int64_t R, int K, int64_t S, float* z, cudaStream_t stream);
int launch_coarse_bwd(const float* g_z, const float* u, int64_t R, int K, float* d_near, float* d_far,
                      cudaStream_t stream);
int launch_importance(const float* weights, const float* z_coarse, const float* u, const float* u2,
                      const float* normals, const float* near, const float* far, int bound_stride,
                      const int64_t* offsets, const int64_t* fine_offsets, int64_t R, int Kc, int n_imp,
                      int n_depth, float depth_std, float* z_fine, float* z_sorted, float* cdf,
                      int32_t* idx, bool allow_fast, cudaStream_t stream);
// importance_reg.cu — dense layout, register-resident sort/merge; AVR_ERR_UNSUPPORTED if the
// shape does not fit (more than 256 new samples or more than 512 merged samples per ray)
int launch_importance_reg(const float* weights, const float* z_coarse, const float* u, const float* u2,
                          const float* normals, const float* near, const float* far, int bound_stride,
                          const int64_t* offsets, const int64_t* fine_offsets, int64_t R, int Kc, int n_imp,
                          int n_depth, float depth_std, float* z_fine, float* z_sorted, float* cdf, int32_t* idx,
                          cudaStream_t stream);
// geometry.cu — ray setup, sample points for the radiance-field callback, depth re-projection
int launch_ray_points(const float* ros, const float* rds, const float* z_or_u, const float* near, const float* far,
                      int bound_stride, bool from_u, int64_t R, int K, float* z_out, float* pts, float* viewdirs,
                      cudaStream_t stream);
int launch_ray_points_bwd(const float* rds, const float* g_pts, int64_t R, int K, float* d_z, cudaStream_t stream);
int launch_ray_points_packed(const float* ros, const float* rds, const float* z, const int64_t* offsets, int64_t R,
                             float* pts, float* viewdirs, const float* g_pts, float* d_z, cudaStream_t stream);
int launch_world_rays(const float* x_pix, const float* kinv, const float* c2w, int64_t R, int64_t rays_per_cam,
                      float* ros, float* rds, float* depth_affine, cudaStream_t stream);
// ray setup + stratified depths + sample points + view directions in one launch
int launch_rays_coarse_points(const float* x_pix, const float* intr, const float* c2w, int64_t rays_per_cam,
                              const float* near, const float* far, int bound_stride, const float* u, int64_t R, int K,
                              float* ros, float* rds, float* depth_affine, float* z, float* pts, float* viewdirs,
                              cudaStream_t stream);
int launch_depth_from_world(const float* ros, const float* rds, const float* dist, const float* c2w, int64_t R,
                            float* depth, float* grad_row, cudaStream_t stream);
// field_inputs.cu — radiance-field front end (models.py:754-826), SURVEY 8(f) row 3
using FieldInputsArgs = ::avr_field_inputs;
int launch_field_inputs_fwd(const FieldInputsArgs& a, cudaStream_t stream);
int launch_field_inputs_bwd(const FieldInputsArgs& a, int64_t SB, cudaStream_t stream);
// lstm_march.cu — the adaptive renderer's LSTM ray march (renderers.py:411-435), SURVEY 8(f) row 4
int launch_lstm_march(const FieldInputsArgs& f, const ::avr_lstm_march& m, bool backward, cudaStream_t stream);
int launch_sort_rays_bwd(const float* g_out, const int32_t* perm, int64_t R, int K, float* d_in, cudaStream_t stream);
int launch_sort_rays(const float* z_in, int64_t R, int K, float* z_out, int32_t* perm, cudaStream_t stream);

}  // namespace avr
