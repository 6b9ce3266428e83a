// Warp-per-ray compositing routines (device side): one warp walks one ray, lanes <->
// consecutive samples, 32 at a time.  Used by the warp-per-ray kernels (composite_wray.cu)
// and, for the rays a packed span tile cannot take, by composite_span_packed.cu.
#pragma once

#include "avr_common.cuh"

namespace avr {

constexpr int kWrayMaxChunks = 8;  // chunks of 32 samples whose e/T stay in registers (256 samples)

// inclusive product scan over the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float p = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v *= p;
  }
  return v;
}

struct ChunkIn {
  float4 c;    // r,g,b,sigma (sigma forced to 0 on lanes past the ray's end: neutral element)
  float zk;    // depth of this sample
  float zn;    // depth paired with this sample in the depth sum (next sample, or `infinity`)
  float delta; // interval length (1e10 for the ray's last sample)
  bool valid, last;
};

__device__ __forceinline__ ChunkIn load_chunk(const float4* __restrict__ rgbs, const float* __restrict__ z,
                                              int64_t begin, int64_t count, int64_t k, float infinity) {
  ChunkIn in;
  in.valid = k < count;
  in.last = (k == count - 1);
  in.c = in.valid ? rgbs[begin + k] : make_float4(0.f, 0.f, 0.f, 0.f);
  in.zk = in.valid ? z[begin + k] : 0.f;
  const float z_after = (k + 1 < count) ? z[begin + k + 1] : 0.f;
  in.zn = in.last ? infinity : z_after;
  in.delta = in.last ? kLastDelta : in.zn - in.zk;
  if (!in.valid) in.delta = 0.f;
  return in;
}

// One chunk of the back-to-front sweep.  Q_carry enters from the right (samples after this
// chunk) and leaves to the left.
struct BwdRay {
  float gr, gg, gb, gd, gbg;
};

__device__ __forceinline__ void bwd_chunk(const ChunkIn& in, float e, float T, const BwdRay& g, float g_w,
                                          float& Q_carry, int lane, float4& d_out, float& ddelta, float& a_next) {
  const float alpha = 1.0f - e;
  const float t = (1.0f - alpha) + kTransEps;
  float gs = g.gr * in.c.x + g.gg * in.c.y + g.gb * in.c.z + g.gd * in.zn - g.gbg + g_w;
  if (!in.valid) gs = 0.f;
  // reverse inclusive scan of the affine maps Q_left = A + B * Q_right
  float A = gs * alpha, B = t;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float Ap = __shfl_down_sync(0xffffffffu, A, d);
    const float Bp = __shfl_down_sync(0xffffffffu, B, d);
    if (lane + d < 32) {
      A = A + B * Ap;
      B = B * Bp;
    }
  }
  // Q for this sample = everything to its right: the inclusive result of lane+1 applied to Q_carry
  float An = __shfl_down_sync(0xffffffffu, A, 1);
  float Bn = __shfl_down_sync(0xffffffffu, B, 1);
  if (lane == 31) {
    An = 0.f;
    Bn = 1.0f;
  }
  const float Q = An + Bn * Q_carry;
  const float dalpha = T * (gs - Q);
  const float dsd = dalpha * e;
  const float w = alpha * T;
  d_out = make_float4(w * g.gr, w * g.gg, w * g.gb, dsd * in.delta);
  ddelta = (in.valid && !in.last) ? dsd * in.c.w : 0.f;          // dL/d delta_k
  a_next = (in.valid && !in.last) ? ddelta + g.gd * w : 0.f;     // contribution to d z_{k+1}
  const float A0 = __shfl_sync(0xffffffffu, A, 0), B0 = __shfl_sync(0xffffffffu, B, 0);
  Q_carry = A0 + B0 * Q_carry;
}


// Forward for one ray.  All 32 lanes of the warp call this together.
__device__ __forceinline__ void wray_fwd_ray(const float4* __restrict__ rgbs, const float* __restrict__ z,
                                             int64_t begin, int64_t count, int64_t r, int white_back,
                                             float infinity, float* __restrict__ w_out,
                                             float* __restrict__ rgb_out, float* __restrict__ depth_out, int lane,
                                             const float* __restrict__ depth_affine = nullptr) {
  float carry = 1.0f;
  float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, acc = 0.f;
  for (int64_t c0 = 0; c0 < count; c0 += 32) {
    const int64_t k = c0 + lane;
    const ChunkIn in = load_chunk(rgbs, z, begin, count, k, infinity);
    const Opacity o = opacity(in.c.w, in.delta);   // invalid lanes: alpha = 0, t = 1
    const float incl = warp_scan_mul(o.t, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float w = o.alpha * (carry * excl);
    if (w_out && in.valid) w_out[begin + k] = w;
    ar += w * in.c.x;
    ag += w * in.c.y;
    ab += w * in.c.z;
    ad += w * in.zn;
    acc += w;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, d);
    ag += __shfl_xor_sync(0xffffffffu, ag, d);
    ab += __shfl_xor_sync(0xffffffffu, ab, d);
    ad += __shfl_xor_sync(0xffffffffu, ad, d);
    acc += __shfl_xor_sync(0xffffffffu, acc, d);
  }
  if (lane == 0) {
    const float bg = white_back ? 1.0f - acc : 0.f;
    rgb_out[r * 3 + 0] = ar + bg;
    rgb_out[r * 3 + 1] = ag + bg;
    rgb_out[r * 3 + 2] = ab + bg;
    depth_out[r] = cam_depth(depth_affine, r, ad);
  }
}

// Backward for one ray (count > 0).  kRegs: e_k / T_k of up to 256 samples stay in registers
// between the two sweeps; otherwise T_k is parked in the sigma slot of d_rgbs and e_k recomputed.
template <bool kRegs>
__device__ __forceinline__ void wray_bwd_ray(const float4* __restrict__ rgbs, const float* __restrict__ z,
                                             int64_t begin, int64_t count, int64_t r,
                                             const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
                                             const float* __restrict__ g_w, int white_back, float infinity,
                                             float4* __restrict__ d_rgbs, float* __restrict__ d_z, int lane,
                                             const float* __restrict__ depth_affine = nullptr) {
  float* park = reinterpret_cast<float*>(d_rgbs);
  BwdRay g;
  g.gr = g_rgb ? g_rgb[r * 3 + 0] : 0.f;
  g.gg = g_rgb ? g_rgb[r * 3 + 1] : 0.f;
  g.gb = g_rgb ? g_rgb[r * 3 + 2] : 0.f;
  g.gd = g_depth ? cam_depth_grad(depth_affine, r, g_depth[r]) : 0.f;
  g.gbg = white_back ? (g.gr + g.gg + g.gb) : 0.f;
  const int n_chunks = (int)((count + 31) >> 5);

  // sweep 1, front to back: transmittance before every sample
  float e_c[kRegs ? kWrayMaxChunks : 1], T_c[kRegs ? kWrayMaxChunks : 1];
  if (kRegs) {
    float carry = 1.0f;
#pragma unroll
    for (int i = 0; i < kWrayMaxChunks; ++i) {
      e_c[kRegs ? i : 0] = 1.0f;
      T_c[kRegs ? i : 0] = 0.f;
      if (i < n_chunks) {
        const ChunkIn in = load_chunk(rgbs, z, begin, count, (int64_t)i * 32 + lane, infinity);
        const Opacity o = opacity(in.c.w, in.delta);
        const float incl = warp_scan_mul(o.t, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        e_c[kRegs ? i : 0] = o.e;
        T_c[kRegs ? i : 0] = carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
      }
    }
  } else {
    float carry = 1.0f;
    for (int i = 0; i < n_chunks; ++i) {
      const int64_t k = (int64_t)i * 32 + lane;
      const ChunkIn in = load_chunk(rgbs, z, begin, count, k, infinity);
      const Opacity o = opacity(in.c.w, in.delta);
      const float incl = warp_scan_mul(o.t, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      if (in.valid) park[(begin + k) * 4 + 3] = carry * excl;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
  }

  // sweep 2, back to front
  float Q_carry = 0.f;
  float pend_ddelta0 = 0.f;  // dL/d delta of the first sample of the chunk to the right
  bool have_pend = false;
  auto do_chunk = [&](int i, float e, float T, const ChunkIn& in) {
    const int64_t k = (int64_t)i * 32 + lane;
    const float gw = (g_w && in.valid) ? g_w[begin + k] : 0.f;
    float4 d_out;
    float ddelta, a_next;
    bwd_chunk(in, e, T, g, gw, Q_carry, lane, d_out, ddelta, a_next);
    if (in.valid) d_rgbs[begin + k] = d_out;
    if (d_z) {
      // d z_k = (contribution of sample k-1 through its interval and the depth sum) - dL/d delta_k
      float from_prev = __shfl_up_sync(0xffffffffu, a_next, 1);
      const float a_last = __shfl_sync(0xffffffffu, a_next, 31);
      const float ddelta_first = __shfl_sync(0xffffffffu, ddelta, 0);
      if (have_pend && lane == 0) {
        // first sample of the chunk to the right: its left neighbour is this chunk's lane 31
        d_z[begin + (int64_t)(i + 1) * 32] = a_last - pend_ddelta0;
      }
      if (lane > 0 && in.valid) d_z[begin + k] = from_prev - ddelta;
      pend_ddelta0 = ddelta_first;
      have_pend = true;
    }
  };
  if (kRegs) {
#pragma unroll
    for (int i = kWrayMaxChunks - 1; i >= 0; --i) {
      if (i < n_chunks) {
        do_chunk(i, e_c[kRegs ? i : 0], T_c[kRegs ? i : 0],
                 load_chunk(rgbs, z, begin, count, (int64_t)i * 32 + lane, infinity));
      }
    }
  } else {
    for (int i = n_chunks - 1; i >= 0; --i) {
      const int64_t k = (int64_t)i * 32 + lane;
      const ChunkIn in = load_chunk(rgbs, z, begin, count, k, infinity);
      const float T = in.valid ? park[(begin + k) * 4 + 3] : 0.f;
      const float e = opacity(in.c.w, in.delta).e;  // recomputed from the inputs
      __syncwarp();
      do_chunk(i, e, T, in);
    }
  }
  if (d_z && lane == 0) d_z[begin] = -pend_ddelta0;
}

}  // namespace avr
