// Generic compositing kernels: one thread walks one ray straight from global memory.
//
// They take ANY shape — any K, unaligned views, the packed layout with arbitrary
// per-ray counts (including 0) — and are the path for shapes the TMA-staged span
// kernels (composite_span.cu) do not cover.  Arithmetic follows volume_integral,
// renderers.py:69-119, sample by sample; see Appendix A of SURVEY.md.
#include "avr_common.cuh"
#include "kernels.h"

namespace avr {

// Ray r of a dense [R,K] tensor or of a packed stream.
struct RaySpan {
  int64_t begin;
  int64_t count;
};
__device__ __forceinline__ RaySpan ray_span(const int64_t* __restrict__ offsets, int64_t r, int K) {
  RaySpan s;
  if (offsets) {
    s.begin = offsets[r];
    s.count = offsets[r + 1] - s.begin;
  } else {
    s.begin = r * (int64_t)K;
    s.count = K;
  }
  return s;
}

__global__ void __launch_bounds__(128)
composite_fwd_ray_kernel(const float4* __restrict__ rgbs, const float* __restrict__ z,
                         const int64_t* __restrict__ offsets, int64_t R, int K, int white_back,
                         float infinity, float* __restrict__ w_out, float* __restrict__ rgb_out,
                         float* __restrict__ depth_out, const float* __restrict__ depth_affine) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R) return;
  RaySpan s = ray_span(offsets, r, K);
  float T = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, acc = 0.f;
  float zk = s.count > 0 ? z[s.begin] : 0.f;
  for (int64_t k = 0; k < s.count; ++k) {
    const bool last = (k + 1 == s.count);
    float4 c = rgbs[s.begin + k];
    float zn = last ? infinity : z[s.begin + k + 1];
    float delta = last ? kLastDelta : zn - zk;
    Opacity o = opacity(c.w, delta);
    float w = o.alpha * T;
    ar += w * c.x;
    ag += w * c.y;
    ab += w * c.z;
    ad += w * zn;
    acc += w;
    if (w_out) w_out[s.begin + k] = w;
    T *= o.t;
    zk = zn;
  }
  if (white_back) {
    float bg = 1.0f - acc;
    ar += bg;
    ag += bg;
    ab += bg;
  }
  rgb_out[r * 3 + 0] = ar;
  rgb_out[r * 3 + 1] = ag;
  rgb_out[r * 3 + 2] = ab;
  depth_out[r] = cam_depth(depth_affine, r, ad);
}

// Backward, two sweeps per ray.  Sweep 1 (front to back) recomputes the transmittance
// T_k and parks it in the sigma slot of d_rgbs (which this thread owns until sweep 2
// overwrites it).  Sweep 2 (back to front) carries
//     Q_k = sum_{i>k} g_i * alpha_i * prod_{k<j<i} t_j          (Q_{k-1} = g_k*alpha_k + t_k*Q_k)
// so that dL/dalpha_k = T_k * (g_k - Q_k): the same value autograd's cumprod backward
// forms as g_k*T_k - (sum_{i>k} g_i w_i)/t_k, without the division.
__global__ void __launch_bounds__(128)
composite_bwd_ray_kernel(const float4* __restrict__ rgbs, const float* __restrict__ z,
                         const int64_t* __restrict__ offsets, const float* __restrict__ g_rgb,
                         const float* __restrict__ g_depth, const float* __restrict__ g_w, int64_t R,
                         int K, int white_back, float infinity, float4* __restrict__ d_rgbs,
                         float* __restrict__ d_z, const float* __restrict__ depth_affine) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R) return;
  RaySpan s = ray_span(offsets, r, K);
  if (s.count == 0) return;
  float* park = reinterpret_cast<float*>(d_rgbs);
  {
    float T = 1.0f;
    float zk = z[s.begin];
    for (int64_t k = 0; k < s.count; ++k) {
      const bool last = (k + 1 == s.count);
      float sigma = rgbs[s.begin + k].w;
      float zn = last ? infinity : z[s.begin + k + 1];
      float delta = last ? kLastDelta : zn - zk;
      park[(s.begin + k) * 4 + 3] = T;
      T *= opacity(sigma, delta).t;
      zk = zn;
    }
  }
  const float gr = g_rgb ? g_rgb[r * 3 + 0] : 0.f;
  const float gg = g_rgb ? g_rgb[r * 3 + 1] : 0.f;
  const float gb = g_rgb ? g_rgb[r * 3 + 2] : 0.f;
  const float gd = g_depth ? cam_depth_grad(depth_affine, r, g_depth[r]) : 0.f;
  const float gbg = white_back ? (gr + gg + gb) : 0.f;
  float Q = 0.f;
  float ddelta_next = 0.f;  // dL/ddelta_{k+1}
  float zn = infinity;
  for (int64_t k = s.count - 1; k >= 0; --k) {
    const bool last = (k + 1 == s.count);
    float4 c = rgbs[s.begin + k];
    float zk = z[s.begin + k];
    float delta = last ? kLastDelta : zn - zk;
    Opacity o = opacity(c.w, delta);
    float T = park[(s.begin + k) * 4 + 3];
    float g = gr * c.x + gg * c.y + gb * c.z + gd * zn - gbg;
    if (g_w) g += g_w[s.begin + k];
    float dalpha = T * (g - Q);
    Q = g * o.alpha + o.t * Q;
    float dsd = dalpha * o.e;
    float w = o.alpha * T;
    d_rgbs[s.begin + k] = make_float4(w * gr, w * gg, w * gb, dsd * delta);
    if (d_z) {
      float ddelta = last ? 0.f : dsd * c.w;
      if (!last) d_z[s.begin + k + 1] = (ddelta + gd * w) - ddelta_next;
      ddelta_next = ddelta;
    }
    zn = zk;
  }
  if (d_z) d_z[s.begin] = -ddelta_next;
}

int launch_composite_fwd_generic(const float* rgbs, const float* z, const int64_t* offsets, int64_t R,
                                 int K, int white_back, float infinity, float* w, float* rgb,
                                 float* depth, cudaStream_t stream, const float* depth_affine) {
  if (R == 0) return AVR_OK;
  const int threads = 128;
  const int64_t blocks = (R + threads - 1) / threads;
  composite_fwd_ray_kernel<<<(unsigned)blocks, threads, 0, stream>>>(
      reinterpret_cast<const float4*>(rgbs), z, offsets, R, K, white_back, infinity, w, rgb, depth, depth_affine);
  return check_launch();
}

int launch_composite_bwd_generic(const float* rgbs, const float* z, const int64_t* offsets,
                                 const float* g_rgb, const float* g_depth, const float* g_w, int64_t R,
                                 int K, int white_back, float infinity, float* d_rgbs, float* d_z,
                                 cudaStream_t stream, const float* depth_affine) {
  if (R == 0) return AVR_OK;
  const int threads = 128;
  const int64_t blocks = (R + threads - 1) / threads;
  composite_bwd_ray_kernel<<<(unsigned)blocks, threads, 0, stream>>>(
      reinterpret_cast<const float4*>(rgbs), z, offsets, g_rgb, g_depth, g_w, R, K, white_back, infinity,
      reinterpret_cast<float4*>(d_rgbs), d_z, depth_affine);
  return check_launch();
}

}  // namespace avr
