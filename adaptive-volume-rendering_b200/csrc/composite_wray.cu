// Warp-per-ray compositing kernels: the path for the packed (ragged) layout and for dense
// shapes the span kernels do not take (unaligned views, K not tileable, gradients into
// weights / depths for the adaptive renderer).
//
// One warp walks one ray; lanes <-> consecutive samples, 32 at a time, so every global
// access is a coalesced 512-byte (rgbs) / 128-byte (z, w) segment and short rays of the
// packed layout do not strand a thread on a long serial loop.  The transmittance is an
// exclusive product scan by warp shuffles with a running carry between chunks
// (renderers.py:90-93); the per-ray sums are reduced once per ray.  Backward keeps e_k and
// T_k of up to 256 samples in registers from the forward sweep and walks the chunks back
// to front with a reverse affine scan (Q_{k-1} = g_k alpha_k + t_k Q_k, see
// composite_generic.cu); longer rays park T_k in the sigma slot of d_rgbs instead.
#include "avr_common.cuh"
#include "kernels.h"
#include "wray_device.cuh"

namespace avr {

constexpr int kWrayWarps = 8;    // warps per CTA

struct WraySpan {
  int64_t begin;
  int64_t count;
};
__device__ __forceinline__ WraySpan wray_span(const int64_t* __restrict__ offsets, int64_t r, int K) {
  WraySpan s;
  if (offsets) {
    s.begin = offsets[r];
    s.count = offsets[r + 1] - s.begin;
  } else {
    s.begin = r * (int64_t)K;
    s.count = K;
  }
  return s;
}

__global__ void __launch_bounds__(kWrayWarps * 32)
composite_fwd_wray_kernel(const float4* __restrict__ rgbs, const float* __restrict__ z,
                          const int64_t* __restrict__ offsets, int64_t R, int K, int white_back,
                          float infinity, float* __restrict__ w_out, float* __restrict__ rgb_out,
                          float* __restrict__ depth_out, const float* __restrict__ depth_affine) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * kWrayWarps;
  for (int64_t r = blockIdx.x * (int64_t)kWrayWarps + (threadIdx.x >> 5); r < R; r += warps) {
    const WraySpan s = wray_span(offsets, r, K);
    wray_fwd_ray(rgbs, z, s.begin, s.count, r, white_back, infinity, w_out, rgb_out, depth_out, lane, depth_affine);
  }
}

__global__ void __launch_bounds__(kWrayWarps * 32)
composite_bwd_wray_kernel(const float4* __restrict__ rgbs, const float* __restrict__ z,
                          const int64_t* __restrict__ offsets, const float* __restrict__ g_rgb,
                          const float* __restrict__ g_depth, const float* __restrict__ g_w, int64_t R, int K,
                          int white_back, float infinity, float4* __restrict__ d_rgbs,
                          float* __restrict__ d_z, const float* __restrict__ depth_affine) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * kWrayWarps;
  for (int64_t r = blockIdx.x * (int64_t)kWrayWarps + (threadIdx.x >> 5); r < R; r += warps) {
    const WraySpan s = wray_span(offsets, r, K);
    if (s.count == 0) continue;
    if (s.count <= 32 * kWrayMaxChunks) {
      wray_bwd_ray<true>(rgbs, z, s.begin, s.count, r, g_rgb, g_depth, g_w, white_back, infinity, d_rgbs, d_z, lane, depth_affine);
    } else {
      wray_bwd_ray<false>(rgbs, z, s.begin, s.count, r, g_rgb, g_depth, g_w, white_back, infinity, d_rgbs, d_z, lane, depth_affine);
    }
  }
}

static unsigned wray_grid(int64_t R) {
  int64_t b = (R + kWrayWarps - 1) / kWrayWarps;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

int launch_composite_fwd_wray(const float* rgbs, const float* z, const int64_t* offsets, int64_t R, int K,
                              int white_back, float infinity, float* w, float* rgb, float* depth,
                              cudaStream_t stream, const float* depth_affine) {
  if (R == 0) return AVR_OK;
  composite_fwd_wray_kernel<<<wray_grid(R), kWrayWarps * 32, 0, stream>>>(
      reinterpret_cast<const float4*>(rgbs), z, offsets, R, K, white_back, infinity, w, rgb, depth, depth_affine);
  return check_launch();
}

int launch_composite_bwd_wray(const float* rgbs, const float* z, const int64_t* offsets, const float* g_rgb,
                              const float* g_depth, const float* g_w, int64_t R, int K, int white_back,
                              float infinity, float* d_rgbs, float* d_z, cudaStream_t stream,
                              const float* depth_affine) {
  if (R == 0) return AVR_OK;
  composite_bwd_wray_kernel<<<wray_grid(R), kWrayWarps * 32, 0, stream>>>(
      reinterpret_cast<const float4*>(rgbs), z, offsets, g_rgb, g_depth, g_w, R, K, white_back, infinity,
      reinterpret_cast<float4*>(d_rgbs), d_z, depth_affine);
  return check_launch();
}

}  // namespace avr
