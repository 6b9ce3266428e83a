// extern "C" entry points of libavr_b200.so: argument checks and kernel selection.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "avr_common.cuh"
#include "kernels.h"

namespace avr {

static thread_local char g_last_error[256] = "";
static std::atomic<int> g_force_generic{0};

void set_last_cuda_error(cudaError_t e) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
}

int check_launch() {
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return AVR_OK;
  set_last_cuda_error(e);
  (void)cudaGetLastError();  // clear the (non-sticky) launch error so the next call starts clean
  return AVR_ERR_LAUNCH;
}

static inline cudaStream_t as_stream(avr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace avr

using namespace avr;

extern "C" {

int avr_abi_version(void) { return AVR_B200_ABI_VERSION; }

const char* avr_status_string(int status) {
  switch (status) {
    case AVR_OK: return "ok";
    case AVR_ERR_BAD_ARG: return "bad argument";
    case AVR_ERR_LAUNCH: return "kernel launch failed";
    case AVR_ERR_NO_DEVICE: return "no sm_100 device";
    case AVR_ERR_UNSUPPORTED: return "unsupported shape";
    case AVR_ERR_RUNTIME: return "cuda runtime call failed";
    default: return "unknown status";
  }
}

const char* avr_last_cuda_error(void) { return g_last_error; }

int avr_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    (void)cudaGetLastError();
    return AVR_ERR_NO_DEVICE;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    if (e != cudaSuccess) set_last_cuda_error(e);
    return AVR_ERR_NO_DEVICE;
  }
  return AVR_OK;
}

void avr_set_force_generic(int on) { g_force_generic.store(on ? 1 : 0); }

int avr_composite_plan(int64_t R, int K, const void* rgbs, const void* z) {
  SpanPlan p;
  if (g_force_generic.load()) return 0;
  return span_plan(R, K, rgbs, z, &p) ? 1 : 0;
}

int avr_composite_plan_info(int64_t R, int K, const void* rgbs, const void* z, int* samples_per_lane,
                            int* rays_per_tile, int64_t* main_rays) {
  SpanPlan p{};
  const bool ok = !g_force_generic.load() && span_plan(R, K, rgbs, z, &p);
  if (samples_per_lane) *samples_per_lane = ok ? p.L : 0;
  if (rays_per_tile) *rays_per_tile = ok ? p.rays_per_tile : 0;
  if (main_rays) *main_rays = ok ? p.main_rays : 0;
  return ok ? 1 : 0;
}

/* ---------------------------------------------------------------- samplers -- */

int avr_coarse_sample_fwd(const float* near, const float* far, int bound_stride, const float* u, int64_t R,
                          int K, float* z, avr_stream_t stream) {
  if (R < 0 || K < 1 || (bound_stride != 0 && bound_stride != 1)) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!near || !far || !u || !z) return AVR_ERR_BAD_ARG;
  return launch_coarse_fwd(near, far, bound_stride, u, nullptr, R, K, R * K, z, as_stream(stream));
}

int avr_coarse_sample_fwd_packed(const float* near, const float* far, int bound_stride, const float* u,
                                 const int64_t* offsets, int64_t R, int64_t S, float* z,
                                 avr_stream_t stream) {
  if (R < 0 || S < 0 || (bound_stride != 0 && bound_stride != 1)) return AVR_ERR_BAD_ARG;
  if (R == 0 || S == 0) return AVR_OK;
  if (!near || !far || !u || !z || !offsets) return AVR_ERR_BAD_ARG;
  return launch_coarse_fwd(near, far, bound_stride, u, offsets, R, 0, S, z, as_stream(stream));
}

int avr_coarse_sample_bwd(const float* g_z, const float* u, int64_t R, int K, float* d_near, float* d_far,
                          avr_stream_t stream) {
  if (R < 0 || K < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!g_z || !u || !d_near || !d_far) return AVR_ERR_BAD_ARG;
  return launch_coarse_bwd(g_z, u, R, K, d_near, d_far, as_stream(stream));
}

int avr_importance_sample(const float* weights, const float* z_coarse, const float* u, const float* u2,
                          const float* normals, const float* near, const float* far, int bound_stride,
                          int64_t R, int Kc, int n_imp, int n_depth, float depth_std, float* z_fine,
                          float* z_sorted, float* cdf, int32_t* idx, avr_stream_t stream) {
  if (R < 0 || Kc < 1 || n_imp < 0 || n_depth < 0 || (bound_stride != 0 && bound_stride != 1))
    return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!weights || !near || !far) return AVR_ERR_BAD_ARG;
  if (n_imp > 0 && (!u || !u2)) return AVR_ERR_BAD_ARG;
  if (z_sorted && !z_coarse) return AVR_ERR_BAD_ARG;
  if (z_sorted && n_depth > 0 && !normals) return AVR_ERR_BAD_ARG;
  if (!z_sorted) n_depth = 0;
  return launch_importance(weights, z_coarse, u, u2, normals, near, far, bound_stride, nullptr, nullptr, R, Kc,
                           n_imp, n_depth, depth_std, z_fine, z_sorted, cdf, idx, !g_force_generic.load(), as_stream(stream));
}

int avr_importance_sample_packed(const float* weights, const float* z_coarse, const float* u, const float* u2,
                                 const float* near, const float* far, int bound_stride, const int64_t* offsets,
                                 const int64_t* fine_offsets, int64_t R, int max_coarse, int max_fine,
                                 float* z_fine, float* z_sorted, float* cdf, int32_t* idx, avr_stream_t stream) {
  if (R < 0 || max_coarse < 1 || max_fine < 0 || (bound_stride != 0 && bound_stride != 1))
    return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!weights || !near || !far || !offsets || !fine_offsets || !u || !u2) return AVR_ERR_BAD_ARG;
  if (z_sorted && !z_coarse) return AVR_ERR_BAD_ARG;
  return launch_importance(weights, z_coarse, u, u2, nullptr, near, far, bound_stride, offsets, fine_offsets, R,
                           max_coarse, max_fine, 0, 0.f, z_fine, z_sorted, cdf, idx, !g_force_generic.load(), as_stream(stream));
}

int avr_sort_rays(const float* z_in, int64_t R, int K, float* z_out, int32_t* perm, avr_stream_t stream) {
  if (R < 0 || K < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!z_in || !z_out) return AVR_ERR_BAD_ARG;
  return launch_sort_rays(z_in, R, K, z_out, perm, as_stream(stream));
}

int avr_sort_rays_bwd(const float* g_out, const int32_t* perm, int64_t R, int K, float* d_in, avr_stream_t stream) {
  if (R < 0 || K < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!g_out || !perm || !d_in) return AVR_ERR_BAD_ARG;
  return launch_sort_rays_bwd(g_out, perm, R, K, d_in, as_stream(stream));
}

/* ------------------------------------------------------------- compositing -- */

int avr_composite_fwd(const float* rgbs, const float* z, int64_t R, int K, int white_back, float infinity,
                      float* w, float* rgb, float* depth, avr_stream_t stream) {
  return avr_composite_fwd_camera(rgbs, z, nullptr, R, K, white_back, infinity, w, rgb, depth, stream);
}

int avr_composite_fwd_camera(const float* rgbs, const float* z, const float* depth_affine, int64_t R, int K,
                             int white_back, float infinity, float* w, float* rgb, float* depth, avr_stream_t stream) {
  if (R < 0 || K < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!rgbs || !z || !rgb || !depth || !aligned16(rgbs)) return AVR_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  SpanPlan plan;
  int64_t done = 0;
  if (!g_force_generic.load() && span_plan(R, K, rgbs, z, &plan) && (w == nullptr || aligned16(w))) {
    int rc = launch_composite_fwd_span(plan, rgbs, z, K, white_back, infinity, w, rgb, depth, st, nullptr, 0, 0, false,
                                       nullptr, depth_affine);
    if (rc != AVR_OK) return rc;
    count_dispatch(AVR_DISPATCH_FWD_SPAN);
    done = plan.main_rays;
  }
  if (done < R) {
    count_dispatch(g_force_generic.load() ? AVR_DISPATCH_FWD_GENERIC : AVR_DISPATCH_FWD_WRAY);
    auto rest = g_force_generic.load() ? launch_composite_fwd_generic : launch_composite_fwd_wray;
    return rest(rgbs + done * K * 4, z + done * K, nullptr, R - done, K, white_back, infinity,
                w ? w + done * K : nullptr, rgb + done * 3, depth + done, st,
                depth_affine ? depth_affine + done * 2 : nullptr);
  }
  return AVR_OK;
}

int avr_composite_fwd_gather(const float* rgbs, const float* z, int64_t R, int K, int white_back, float infinity,
                             float* w, float* rgb, float* depth, void* const* peer_gathered, int n_peers,
                             int64_t row0, avr_stream_t stream) {
  if (R < 0 || K < 1 || n_peers < 1 || row0 < 0 || !peer_gathered) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!rgbs || !z || !rgb || !depth || !aligned16(rgbs)) return AVR_ERR_BAD_ARG;
  for (int p = 0; p < n_peers; ++p)
    if (!peer_gathered[p] || !aligned16(peer_gathered[p])) return AVR_ERR_BAD_ARG;
  SpanPlan plan;
  // the fused epilogue lives in the span kernel: the whole batch must be tileable
  if (g_force_generic.load() || !span_plan(R, K, rgbs, z, &plan) || plan.main_rays != R || (w && !aligned16(w)))
    return AVR_ERR_UNSUPPORTED;
  return launch_composite_fwd_span(plan, rgbs, z, K, white_back, infinity, w, rgb, depth, as_stream(stream),
                                   peer_gathered, n_peers, row0);
}

int avr_composite_fwd_gather_signal(const float* rgbs, const float* z, int64_t R, int K, int white_back,
                                    float infinity, float* w, float* rgb, float* depth, void* const* peer_gathered,
                                    int n_peers, int multicast, int64_t row0, uint32_t* const* peer_flags,
                                    int n_flag_peers, int self_rank, uint32_t value, uint32_t* done_counter,
                                    avr_stream_t stream) {
  if (R < 1 || K < 1 || n_peers < 1 || row0 < 0 || !peer_gathered || !peer_flags || !done_counter) return AVR_ERR_BAD_ARG;
  if (n_flag_peers < 1 || n_flag_peers > 16 || self_rank < 0 || self_rank >= 32) return AVR_ERR_BAD_ARG;
  if (multicast && n_peers != 1) return AVR_ERR_BAD_ARG;
  if (!rgbs || !z || !rgb || !depth || !aligned16(rgbs)) return AVR_ERR_BAD_ARG;
  for (int p = 0; p < n_peers; ++p)
    if (!peer_gathered[p] || !aligned16(peer_gathered[p])) return AVR_ERR_BAD_ARG;
  for (int p = 0; p < n_flag_peers; ++p)
    if (!peer_flags[p]) return AVR_ERR_BAD_ARG;
  SpanPlan plan;
  if (g_force_generic.load() || !span_plan(R, K, rgbs, z, &plan) || plan.main_rays != R || (w && !aligned16(w)))
    return AVR_ERR_UNSUPPORTED;
  GatherSignal sig{peer_flags, n_flag_peers, self_rank, value, done_counter};
  count_dispatch(AVR_DISPATCH_FWD_SPAN);
  return launch_composite_fwd_span(plan, rgbs, z, K, white_back, infinity, w, rgb, depth, as_stream(stream),
                                   peer_gathered, n_peers, row0, multicast != 0, &sig);
}

int avr_composite_fwd_gather_multicast(const float* rgbs, const float* z, int64_t R, int K, int white_back,
                                       float infinity, float* w, float* rgb, float* depth, void* multicast_gathered,
                                       int64_t row0, avr_stream_t stream) {
  if (R < 0 || K < 1 || row0 < 0 || !multicast_gathered || !aligned16(multicast_gathered)) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!rgbs || !z || !rgb || !depth || !aligned16(rgbs)) return AVR_ERR_BAD_ARG;
  SpanPlan plan;
  if (g_force_generic.load() || !span_plan(R, K, rgbs, z, &plan) || plan.main_rays != R || (w && !aligned16(w)))
    return AVR_ERR_UNSUPPORTED;
  void* one[1] = {multicast_gathered};
  return launch_composite_fwd_span(plan, rgbs, z, K, white_back, infinity, w, rgb, depth, as_stream(stream), one, 1,
                                   row0, true);
}

int avr_gather_push_rows(void* const* peer_gathered, int n_peers, int self, int64_t row0, int64_t rows,
                         avr_stream_t stream) {
  if (!peer_gathered || n_peers < 1 || self < 0 || self >= n_peers || row0 < 0 || rows < 0) return AVR_ERR_BAD_ARG;
  if (rows == 0) return AVR_OK;
  for (int p = 0; p < n_peers; ++p)
    if (!peer_gathered[p]) return AVR_ERR_BAD_ARG;
  const size_t off = (size_t)row0 * 16, bytes = (size_t)rows * 16;
  const char* src = static_cast<const char*>(peer_gathered[self]) + off;
  for (int d = 1; d < n_peers; ++d) {
    const int p = (self + d) % n_peers;  // staggered: at any moment every rank targets a different peer
    cudaError_t e = cudaMemcpyAsync(static_cast<char*>(peer_gathered[p]) + off, src, bytes, cudaMemcpyDefault,
                                    as_stream(stream));
    if (e != cudaSuccess) {
      set_last_cuda_error(e);
      (void)cudaGetLastError();
      return AVR_ERR_RUNTIME;
    }
  }
  return AVR_OK;
}

int avr_composite_bwd(const float* rgbs, const float* z, const float* g_rgb, const float* g_depth,
                      const float* g_w, int64_t R, int K, int white_back, float infinity, float* d_rgbs,
                      float* d_z, avr_stream_t stream) {
  return avr_composite_bwd_camera(rgbs, z, nullptr, g_rgb, g_depth, g_w, R, K, white_back, infinity, d_rgbs, d_z, stream);
}

int avr_composite_bwd_camera(const float* rgbs, const float* z, const float* depth_affine, const float* g_rgb,
                             const float* g_depth, const float* g_w, int64_t R, int K, int white_back, float infinity,
                             float* d_rgbs, float* d_z, avr_stream_t stream) {
  if (R < 0 || K < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!rgbs || !z || !d_rgbs || !aligned16(rgbs) || !aligned16(d_rgbs)) return AVR_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  SpanPlan plan;
  int64_t done = 0;
  // the span kernel covers the training configurations: no grad into the weights; grad into z
  // (adaptive renderer) when a lane's run holds at most one ray end (K > L) and d_z is 16-byte
  // aligned.  g_w requests and the remaining shapes take the warp-per-ray kernel.
  if (!g_force_generic.load() && !g_w && span_plan(R, K, rgbs, z, &plan) &&
      (!d_z || (K > plan.L && aligned16(d_z)))) {
    int rc = launch_composite_bwd_span(plan, rgbs, z, g_rgb, g_depth, K, white_back, infinity, d_rgbs, d_z, st, depth_affine);
    if (rc != AVR_OK) return rc;
    count_dispatch(AVR_DISPATCH_BWD_SPAN);
    done = plan.main_rays;
  }
  if (done < R) {
    count_dispatch(g_force_generic.load() ? AVR_DISPATCH_BWD_GENERIC : AVR_DISPATCH_BWD_WRAY);
    auto rest = g_force_generic.load() ? launch_composite_bwd_generic : launch_composite_bwd_wray;
    return rest(rgbs + done * K * 4, z + done * K, nullptr, g_rgb ? g_rgb + done * 3 : nullptr,
                g_depth ? g_depth + done : nullptr, g_w ? g_w + done * K : nullptr, R - done, K, white_back,
                infinity, d_rgbs + done * K * 4, d_z ? d_z + done * K : nullptr, st,
                depth_affine ? depth_affine + done * 2 : nullptr);
  }
  return AVR_OK;
}

int avr_composite_fwd_packed(const float* rgbs, const float* z, const int64_t* offsets, int64_t R, int64_t S,
                             int white_back, float infinity, float* w, float* rgb, float* depth,
                             avr_stream_t stream) {
  if (R < 0 || S < 0) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!offsets || !rgb || !depth) return AVR_ERR_BAD_ARG;
  if (S > 0 && (!rgbs || !z || !aligned16(rgbs))) return AVR_ERR_BAD_ARG;
  if (!g_force_generic.load() && S > 0 && span_packed_eligible(rgbs, z, w, nullptr)) {
    count_dispatch(AVR_DISPATCH_FWD_SPAN_PACKED);
    return launch_composite_fwd_span_packed(rgbs, z, offsets, R, S, white_back, infinity, w, rgb, depth,
                                            as_stream(stream));
  }
  count_dispatch(g_force_generic.load() ? AVR_DISPATCH_FWD_GENERIC : AVR_DISPATCH_FWD_WRAY);
  auto fn = g_force_generic.load() ? launch_composite_fwd_generic : launch_composite_fwd_wray;
  return fn(rgbs, z, offsets, R, 0, white_back, infinity, w, rgb, depth, as_stream(stream), nullptr);
}

int avr_composite_bwd_packed(const float* rgbs, const float* z, const int64_t* offsets, const float* g_rgb,
                             const float* g_depth, const float* g_w, int64_t R, int64_t S, int white_back,
                             float infinity, float* d_rgbs, float* d_z, avr_stream_t stream) {
  if (R < 0 || S < 0) return AVR_ERR_BAD_ARG;
  if (R == 0 || S == 0) return AVR_OK;
  if (!offsets || !rgbs || !z || !d_rgbs || !aligned16(rgbs) || !aligned16(d_rgbs)) return AVR_ERR_BAD_ARG;
  if (!g_force_generic.load() && !g_w && !d_z && span_packed_eligible(rgbs, z, nullptr, d_rgbs)) {
    count_dispatch(AVR_DISPATCH_BWD_SPAN_PACKED);
    return launch_composite_bwd_span_packed(rgbs, z, offsets, g_rgb, g_depth, R, S, white_back, infinity, d_rgbs,
                                            as_stream(stream));
  }
  count_dispatch(g_force_generic.load() ? AVR_DISPATCH_BWD_GENERIC : AVR_DISPATCH_BWD_WRAY);
  auto fn = g_force_generic.load() ? launch_composite_bwd_generic : launch_composite_bwd_wray;
  return fn(rgbs, z, offsets, g_rgb, g_depth, g_w, R, 0, white_back, infinity, d_rgbs, d_z, as_stream(stream), nullptr);
}

/* ------------------------------------------ ray setup / sample points / depth -- */

int avr_ray_points_fwd(const float* ros, const float* rds, const float* z, int64_t R, int K, float* pts,
                       float* viewdirs, avr_stream_t stream) {
  if (R < 0 || K < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!ros || !rds || !z || !pts) return AVR_ERR_BAD_ARG;
  return launch_ray_points(ros, rds, z, nullptr, nullptr, 0, false, R, K, nullptr, pts, viewdirs, as_stream(stream));
}

int avr_ray_points_bwd(const float* rds, const float* g_pts, int64_t R, int K, float* d_z, avr_stream_t stream) {
  if (R < 0 || K < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!rds || !g_pts || !d_z) return AVR_ERR_BAD_ARG;
  return launch_ray_points_bwd(rds, g_pts, R, K, d_z, as_stream(stream));
}

int avr_coarse_sample_points_fwd(const float* near, const float* far, int bound_stride, const float* u,
                                 const float* ros, const float* rds, int64_t R, int K, float* z, float* pts,
                                 float* viewdirs, avr_stream_t stream) {
  if (R < 0 || K < 1 || (bound_stride != 0 && bound_stride != 1)) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!near || !far || !u || !ros || !rds || !z || !pts) return AVR_ERR_BAD_ARG;
  return launch_ray_points(ros, rds, u, near, far, bound_stride, true, R, K, z, pts, viewdirs, as_stream(stream));
}

int avr_ray_points_fwd_packed(const float* ros, const float* rds, const float* z, const int64_t* offsets, int64_t R,
                              int64_t S, float* pts, float* viewdirs, avr_stream_t stream) {
  if (R < 0 || S < 0) return AVR_ERR_BAD_ARG;
  if (R == 0 || S == 0) return AVR_OK;
  if (!ros || !rds || !z || !offsets || !pts) return AVR_ERR_BAD_ARG;
  return launch_ray_points_packed(ros, rds, z, offsets, R, pts, viewdirs, nullptr, nullptr, as_stream(stream));
}

int avr_ray_points_bwd_packed(const float* rds, const float* g_pts, const int64_t* offsets, int64_t R, int64_t S,
                              float* d_z, avr_stream_t stream) {
  if (R < 0 || S < 0) return AVR_ERR_BAD_ARG;
  if (R == 0 || S == 0) return AVR_OK;
  if (!rds || !g_pts || !offsets || !d_z) return AVR_ERR_BAD_ARG;
  return launch_ray_points_packed(rds, rds, nullptr, offsets, R, nullptr, nullptr, g_pts, d_z, as_stream(stream));
}

int avr_world_rays(const float* x_pix, const float* kinv, const float* cam2world, int64_t R, int64_t rays_per_cam,
                   float* ros, float* rds, float* depth_affine, avr_stream_t stream) {
  if (R < 0 || rays_per_cam < 1) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!x_pix || !kinv || !cam2world || !ros || !rds || !aligned16(cam2world)) return AVR_ERR_BAD_ARG;
  return launch_world_rays(x_pix, kinv, cam2world, R, rays_per_cam, ros, rds, depth_affine, as_stream(stream));
}

int avr_rays_coarse_sample_points_fwd(const float* x_pix, const float* intrinsics, const float* cam2world,
                                      int64_t rays_per_cam, const float* near, const float* far, int bound_stride,
                                      const float* u, int64_t R, int K, float* ros, float* rds, float* depth_affine,
                                      float* z, float* pts, float* viewdirs, avr_stream_t stream) {
  if (R < 0 || K < 1 || rays_per_cam < 1 || (bound_stride != 0 && bound_stride != 1)) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!x_pix || !intrinsics || !cam2world || !near || !far || !u || !ros || !rds || !z || !pts || !aligned16(cam2world))
    return AVR_ERR_BAD_ARG;
  return launch_rays_coarse_points(x_pix, intrinsics, cam2world, rays_per_cam, near, far, bound_stride, u, R, K, ros, rds,
                                   depth_affine, z, pts, viewdirs, as_stream(stream));
}

int avr_depth_from_world(const float* ros, const float* rds, const float* dist, const float* cam2world, int64_t R,
                         float* depth, float* grad_row, avr_stream_t stream) {
  if (R < 0) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!ros || !cam2world || !depth || (dist && !rds) || !aligned16(cam2world)) return AVR_ERR_BAD_ARG;
  return launch_depth_from_world(ros, rds, dist, cam2world, R, depth, grad_row, as_stream(stream));
}

/* ------------------------------------------ radiance-field front end -- */

static int check_field_desc(const avr_field_inputs* d, bool backward) {
  if (!d || d->B < 0 || d->NV < 0 || d->NS < 1 || d->NV % d->NS != 0) return AVR_ERR_BAD_ARG;
  if (d->C < 4 || d->C % 4 != 0 || d->H < 1 || d->W < 1) return AVR_ERR_BAD_ARG;
  if (d->n_sin < 0 || d->n_sin > AVR_FIELD_MAX_SIN) return AVR_ERR_BAD_ARG;
  if (d->NV * d->B == 0) return AVR_OK;
  const int width = d->features_only ? 0 : (d->include_input ? 3 : 0) + 3 * d->n_sin + (d->use_viewdirs ? 3 : 0);
  if ((d->C + width) % 2 != 0) return AVR_ERR_UNSUPPORTED;  // rows are written in 8-byte pieces
  if (!d->xyz || !d->poses || !d->focal || !d->c || !d->latent) return AVR_ERR_BAD_ARG;
  if (d->use_viewdirs && !d->features_only && !d->viewdirs) return AVR_ERR_BAD_ARG;
  if (!aligned16(d->latent)) return AVR_ERR_BAD_ARG;
  if (backward) {
    if (!d->g_out || (reinterpret_cast<uintptr_t>(d->g_out) & 7u)) return AVR_ERR_BAD_ARG;
    if (d->d_latent && !aligned16(d->d_latent)) return AVR_ERR_BAD_ARG;
  } else {
    if (!d->out || (reinterpret_cast<uintptr_t>(d->out) & 7u)) return AVR_ERR_BAD_ARG;
  }
  return 1;  // work to do
}

int avr_field_inputs_fwd(const avr_field_inputs* desc, avr_stream_t stream) {
  const int rc = check_field_desc(desc, false);
  if (rc <= 0) return rc;
  return launch_field_inputs_fwd(*desc, as_stream(stream));
}

int avr_field_inputs_bwd(const avr_field_inputs* desc, avr_stream_t stream) {
  const int rc = check_field_desc(desc, true);
  if (rc < 0) return rc;
  return launch_field_inputs_bwd(*desc, desc->NV / desc->NS, as_stream(stream));
}

/* ------------------------------------------ the adaptive renderer's LSTM ray march -- */

static int check_march(const avr_field_inputs* f, const avr_lstm_march* m, bool backward) {
  if (!f || !m || m->R < 0 || m->steps < 0 || m->rays_per_obj < 1) return AVR_ERR_BAD_ARG;
  if (m->R == 0 || m->steps == 0) return AVR_OK;
  if (f->NS != 1 || !f->features_only) return AVR_ERR_UNSUPPORTED;
  if (f->C != 128 && f->C != 256 && f->C != 512) return AVR_ERR_UNSUPPORTED;
  if (f->H < 1 || f->W < 1 || !f->poses || !f->focal || !f->c || !f->latent || !aligned16(f->latent)) return AVR_ERR_BAD_ARG;
  if ((m->R + m->rays_per_obj - 1) / m->rays_per_obj > f->NV) return AVR_ERR_BAD_ARG;
  if (!m->ros || !m->rds || !m->w_ih || !m->w_hh || !m->b_ih || !m->b_hh || !m->w_out || !m->b_out || !m->world)
    return AVR_ERR_BAD_ARG;
  if (!aligned16(m->w_ih)) return AVR_ERR_BAD_ARG;
  const int saved = (m->feats ? 1 : 0) + (m->gates ? 1 : 0) + (m->cells ? 1 : 0) + (m->hidden ? 1 : 0);
  if (backward) {
    if (saved != 4 || !m->g_world || !m->d_gates || !m->d_dist) return AVR_ERR_BAD_ARG;
    if (f->d_latent && !aligned16(f->d_latent)) return AVR_ERR_BAD_ARG;
  } else {
    if (!m->init_dist || (saved != 0 && saved != 4)) return AVR_ERR_BAD_ARG;
    if (m->feats && !aligned16(m->feats)) return AVR_ERR_BAD_ARG;
  }
  return 1;
}

int avr_lstm_march_fwd(const avr_field_inputs* field, const avr_lstm_march* march, avr_stream_t stream) {
  const int rc = check_march(field, march, false);
  if (rc <= 0) return rc;
  return launch_lstm_march(*field, *march, false, as_stream(stream));
}

int avr_lstm_march_bwd(const avr_field_inputs* field, const avr_lstm_march* march, avr_stream_t stream) {
  const int rc = check_march(field, march, true);
  if (rc <= 0) return rc;
  return launch_lstm_march(*field, *march, true, as_stream(stream));
}

/* ------------------------------------------------ host-buffer (end to end) -- */

}  // extern "C"

// Rays are independent, so the pass is chunked and software-pipelined over three slots:
// while chunk c computes, chunk c+1 uploads and chunk c-1 downloads (PCIe is full duplex).
// Stream order inside a slot serialises the reuse of its buffers.
struct avr_host_workspace {
  static constexpr int kMaxSlots = 8;
  int n_slots = 3;   // AVR_HOST_SLOTS (2..8): chunks in flight
  struct Slot {
    float *rgbs = nullptr, *z = nullptr, *g_rgb = nullptr, *g_depth = nullptr;
    float *rgb = nullptr, *depth = nullptr, *d_rgbs = nullptr, *w = nullptr;
    cudaStream_t stream = nullptr;
  } slots[kMaxSlots];
  int K = 0;
  int64_t chunk_rays = 0;
  int device = 0;
};

static void free_workspace(avr_host_workspace* ws) {
  for (int i = 0; i < ws->n_slots; ++i) {
    auto& sl = ws->slots[i];
    if (sl.stream) {
      cudaStreamSynchronize(sl.stream);
      cudaStreamDestroy(sl.stream);
    }
    cudaFree(sl.rgbs);
    cudaFree(sl.z);
    cudaFree(sl.g_rgb);
    cudaFree(sl.g_depth);
    cudaFree(sl.rgb);
    cudaFree(sl.depth);
    cudaFree(sl.d_rgbs);
    cudaFree(sl.w);
  }
  delete ws;
}

#define AVR_RT(call)            \
  do {                          \
    cudaError_t e_ = (call);    \
    if (e_ != cudaSuccess) {    \
      set_last_cuda_error(e_);  \
      (void)cudaGetLastError(); \
      return AVR_ERR_RUNTIME;   \
    }                           \
  } while (0)

static int alloc_workspace(avr_host_workspace* ws) {
  const size_t nk = (size_t)ws->chunk_rays * ws->K;
  const size_t nr = (size_t)ws->chunk_rays;
  for (int i = 0; i < ws->n_slots; ++i) {
    auto& sl = ws->slots[i];
    AVR_RT(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    AVR_RT(cudaMalloc(&sl.rgbs, nk * 16));
    AVR_RT(cudaMalloc(&sl.z, nk * 4));
    AVR_RT(cudaMalloc(&sl.g_rgb, nr * 12));
    AVR_RT(cudaMalloc(&sl.g_depth, nr * 4));
    AVR_RT(cudaMalloc(&sl.rgb, nr * 12));
    AVR_RT(cudaMalloc(&sl.depth, nr * 4));
    AVR_RT(cudaMalloc(&sl.d_rgbs, nk * 16));
    AVR_RT(cudaMalloc(&sl.w, nk * 4));
  }
  return AVR_OK;
}

static int enqueue_host_pass(avr_host_workspace* ws, const float* rgbs, const float* z, const float* g_rgb,
                             const float* g_depth, int64_t R, int white_back, float infinity, float* rgb,
                             float* depth, float* w, float* d_rgbs) {
  const int K = ws->K;
  const int64_t chunk = ws->chunk_rays;
  const int64_t n_chunks = (R + chunk - 1) / chunk;
  int rc = AVR_OK;
  for (int64_t c = 0; c < n_chunks; ++c) {
    auto& sl = ws->slots[c % ws->n_slots];
    const int64_t r0 = c * chunk;
    const int64_t rn = (R - r0 < chunk) ? R - r0 : chunk;
    const size_t n = (size_t)rn * K;
    AVR_RT(cudaMemcpyAsync(sl.rgbs, rgbs + r0 * K * 4, n * 16, cudaMemcpyHostToDevice, sl.stream));
    AVR_RT(cudaMemcpyAsync(sl.z, z + r0 * K, n * 4, cudaMemcpyHostToDevice, sl.stream));
    if (g_rgb) AVR_RT(cudaMemcpyAsync(sl.g_rgb, g_rgb + r0 * 3, (size_t)rn * 12, cudaMemcpyHostToDevice, sl.stream));
    if (g_depth) AVR_RT(cudaMemcpyAsync(sl.g_depth, g_depth + r0, (size_t)rn * 4, cudaMemcpyHostToDevice, sl.stream));
    rc = avr_composite_fwd(sl.rgbs, sl.z, rn, K, white_back, infinity, w ? sl.w : nullptr, sl.rgb, sl.depth, sl.stream);
    if (rc != AVR_OK) return rc;
    rc = avr_composite_bwd(sl.rgbs, sl.z, g_rgb ? sl.g_rgb : nullptr, g_depth ? sl.g_depth : nullptr, nullptr, rn, K,
                           white_back, infinity, sl.d_rgbs, nullptr, sl.stream);
    if (rc != AVR_OK) return rc;
    AVR_RT(cudaMemcpyAsync(rgb + r0 * 3, sl.rgb, (size_t)rn * 12, cudaMemcpyDeviceToHost, sl.stream));
    AVR_RT(cudaMemcpyAsync(depth + r0, sl.depth, (size_t)rn * 4, cudaMemcpyDeviceToHost, sl.stream));
    if (w) AVR_RT(cudaMemcpyAsync(w + r0 * K, sl.w, n * 4, cudaMemcpyDeviceToHost, sl.stream));
    AVR_RT(cudaMemcpyAsync(d_rgbs + r0 * K * 4, sl.d_rgbs, n * 16, cudaMemcpyDeviceToHost, sl.stream));
  }
  return AVR_OK;
}

// Whatever was enqueued is drained before returning — also on the error path, so no copy into the
// caller's host buffers is still in flight once the call has reported its status.
static int run_host_pass(avr_host_workspace* ws, const float* rgbs, const float* z, const float* g_rgb,
                         const float* g_depth, int64_t R, int white_back, float infinity, float* rgb,
                         float* depth, float* w, float* d_rgbs) {
  int prev = 0;
  AVR_RT(cudaGetDevice(&prev));
  if (prev != ws->device) AVR_RT(cudaSetDevice(ws->device));  // streams and buffers live on the workspace's device
  int rc = enqueue_host_pass(ws, rgbs, z, g_rgb, g_depth, R, white_back, infinity, rgb, depth, w, d_rgbs);
  for (int i = 0; i < ws->n_slots; ++i) {
    cudaError_t e = cudaStreamSynchronize(ws->slots[i].stream);
    if (e != cudaSuccess && rc == AVR_OK) {
      set_last_cuda_error(e);
      (void)cudaGetLastError();
      rc = AVR_ERR_RUNTIME;
    }
  }
  if (prev != ws->device) (void)cudaSetDevice(prev);
  return rc;
}

extern "C" {

int avr_host_workspace_create(int K, int64_t chunk_rays, avr_host_workspace** out) {
  if (!out || K < 1) return AVR_ERR_BAD_ARG;
  *out = nullptr;
  int rc = avr_device_check();
  if (rc != AVR_OK) return rc;
  if (chunk_rays <= 0) {
    const int64_t mib = option(OPT_HOST_CHUNK_MIB, 48);
    int64_t target = (int64_t)((mib < 1 ? 1 : mib) << 20) / ((int64_t)K * 16);
    chunk_rays = target < 96 ? 96 : target;
  }
  auto* ws = new avr_host_workspace();
  ws->K = K;
  ws->chunk_rays = chunk_rays;
  int ns = option(OPT_HOST_SLOTS, 3);
  ws->n_slots = ns < 2 ? 2 : (ns > avr_host_workspace::kMaxSlots ? avr_host_workspace::kMaxSlots : ns);
  cudaGetDevice(&ws->device);
  rc = alloc_workspace(ws);
  if (rc != AVR_OK) {
    free_workspace(ws);
    return rc;
  }
  *out = ws;
  return AVR_OK;
}

int avr_host_workspace_destroy(avr_host_workspace* ws) {
  if (!ws) return AVR_OK;
  free_workspace(ws);
  return AVR_OK;
}

int avr_composite_fwd_bwd_host(avr_host_workspace* ws, const float* rgbs, const float* z, const float* g_rgb,
                               const float* g_depth, int64_t R, int K, int white_back, float infinity,
                               float* rgb, float* depth, float* w, float* d_rgbs) {
  if (!ws || R < 0 || K < 1 || K != ws->K) return AVR_ERR_BAD_ARG;
  if (R == 0) return AVR_OK;
  if (!rgbs || !z || !rgb || !depth || !d_rgbs) return AVR_ERR_BAD_ARG;
  return run_host_pass(ws, rgbs, z, g_rgb, g_depth, R, white_back, infinity, rgb, depth, w, d_rgbs);
}

}  // extern "C"
