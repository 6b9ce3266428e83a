// The seams either side of the sampling/compositing path:
//
//   * sample points for the radiance-field callback: pts = o + d*z and the K-fold view-direction
//     broadcast the reference materialises with expand().reshape()  (renderers.py:171-175,
//     260-265, 496-500), optionally fused with the stratified coarse sampler so z, pts and
//     viewdirs leave in one pass over the uniforms;
//   * ray setup, utils.get_world_rays (utils.py:246-267, 309-336): unproject the pixel through
//     K^-1, flip to the camera's -z convention, normalise, rotate into the world;
//   * depth re-projection, utils.depth_from_world (utils.py:358-361): -z of the composited
//     point in camera coordinates.  The reference inverts the (SB,R,4,4) pose tensor per ray;
//     here each thread forms only the third row of its pose's inverse (cofactors).
//
// All of it is elementwise, HBM-bound streaming: 16-byte accesses, one thread per 4 samples.
#include "avr_common.cuh"
#include "kernels.h"

namespace avr {

// grid cap of the grid-stride kernels: `mult` CTAs per SM
static inline int grid_cap(int mult) { return num_sms() * mult; }


struct Ray3 {
  float ox, oy, oz, dx, dy, dz;
};
__device__ __forceinline__ Ray3 load_ray3(const float* __restrict__ ros, const float* __restrict__ rds, int64_t r) {
  Ray3 q;
  q.ox = ros[r * 3 + 0]; q.oy = ros[r * 3 + 1]; q.oz = ros[r * 3 + 2];
  q.dx = rds[r * 3 + 0]; q.dy = rds[r * 3 + 1]; q.dz = rds[r * 3 + 2];
  return q;
}

// pts = ros + rds * z with the reference's two roundings (mul, then add: renderers.py:171)
__device__ __forceinline__ float3 point_on_ray(const Ray3& q, float z) {
  return make_float3(__fadd_rn(q.ox, __fmul_rn(q.dx, z)), __fadd_rn(q.oy, __fmul_rn(q.dy, z)),
                     __fadd_rn(q.oz, __fmul_rn(q.dz, z)));
}

// z = near + (far-near)*(j/K) + (u*(far-near))/K       renderers.py:12-14
__device__ __forceinline__ float coarse_z(float near, float span, int j, float kf, float inv_k, bool pow2, float u) {
  const float base = __fadd_rn(near, __fmul_rn(span, __fdiv_rn((float)j, kf)));
  const float jit = __fmul_rn(u, span);
  return __fadd_rn(base, pow2 ? __fmul_rn(jit, inv_k) : __fdiv_rn(jit, kf));
}

// utils.get_world_rays for ray r: origin = pose[:3,3]; dir = R_pose * normalize(flip(K^-1 [x,y,1])).
// intr: [n_cams,3,3] intrinsics, one per object (ray r uses camera r / rays_per_cam); the 3x3 inverse
// (utils.py:263) is formed from the cofactors — 9 cached loads and ~30 flops.
__device__ __forceinline__ Ray3 ray_from_pixel(const float* __restrict__ x_pix, const float* __restrict__ intr,
                                               const float* __restrict__ c2w, int64_t r, int64_t rays_per_cam) {
  const float* km = intr + (r / rays_per_cam) * 9;
  const float a00 = km[0], a01 = km[1], a02 = km[2], a10 = km[3], a11 = km[4], a12 = km[5], a20 = km[6],
              a21 = km[7], a22 = km[8];
  const float c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
  const float idet = 1.0f / (a00 * c00 + a01 * c01 + a02 * c02);
  const float ki[9] = {c00 * idet, (a02 * a21 - a01 * a22) * idet, (a01 * a12 - a02 * a11) * idet,
                       c01 * idet, (a00 * a22 - a02 * a20) * idet, (a02 * a10 - a00 * a12) * idet,
                       c02 * idet, (a01 * a20 - a00 * a21) * idet, (a00 * a11 - a01 * a10) * idet};
  const float x = x_pix[r * 2 + 0], y = x_pix[r * 2 + 1];
  // einsum('ij,kj->ki', K^-1, [x,y,1])  (utils.py:263)
  float cx = ki[0] * x + ki[1] * y + ki[2];
  float cy = ki[3] * x + ki[4] * y + ki[5];
  float cz = ki[6] * x + ki[7] * y + ki[8];
  // unproject negates x, then scales by z = -1 (utils.py:264-266, :312)
  cx = -cx;
  cx *= -1.0f; cy *= -1.0f; cz *= -1.0f;
  const float nrm = sqrtf(cx * cx + cy * cy + cz * cz);   // torch.norm (utils.py:313)
  cx = cx / nrm; cy = cy / nrm; cz = cz / nrm;
  const float4* m = reinterpret_cast<const float4*>(c2w + r * 16);
  const float4 r0 = m[0], r1 = m[1], r2 = m[2];
  Ray3 q;
  q.ox = r0.w; q.oy = r1.w; q.oz = r2.w;                                      // pose[:3, -1]
  q.dx = r0.x * cx + r0.y * cy + r0.z * cz;                                   // pose @ [dir, 0]
  q.dy = r1.x * cx + r1.y * cy + r1.z * cz;
  q.dz = r2.x * cx + r2.y * cy + r2.z * cz;
  return q;
}

// Row 2 of pose^-1 as (c02, c12, c22, c32) / det: camera z of a world point p is row . [p, 1]
// (inverse(M)[2][j] = cofactor(M)[j][2] / det(M); the cofactors delete column 2 of M).
struct PoseRow2 {
  float x, y, z, w;
};
__device__ __forceinline__ PoseRow2 pose_inverse_row2(const float* __restrict__ c2w, int64_t r) {
  const float4* mp = reinterpret_cast<const float4*>(c2w + r * 16);
  const float4 a = mp[0], b = mp[1], c = mp[2], d = mp[3];
  auto det3 = [](float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
    return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0);
  };
  const float c02 = det3(b.x, b.y, b.w, c.x, c.y, c.w, d.x, d.y, d.w);    // delete row 0, col 2, sign +
  const float c12 = -det3(a.x, a.y, a.w, c.x, c.y, c.w, d.x, d.y, d.w);   // row 1, sign -
  const float c22 = det3(a.x, a.y, a.w, b.x, b.y, b.w, d.x, d.y, d.w);    // row 2, sign +
  const float c32 = -det3(a.x, a.y, a.w, b.x, b.y, b.w, c.x, c.y, c.w);   // row 3, sign -
  const float det = a.z * c02 + b.z * c12 + c.z * c22 + d.z * c32;         // expansion along column 2
  return PoseRow2{c02 / det, c12 / det, c22 / det, c32 / det};
}

// depth_from_world(o + d * dist) = A * dist + B  (utils.py:358-361 at renderers.py:274-275, 508-509):
// what the compositing kernels need to return the camera depth themselves (avr_common.cuh cam_depth)
__device__ __forceinline__ float2 depth_affine_of(const Ray3& q, const PoseRow2& w) {
  return make_float2(-(w.x * q.dx + w.y * q.dy + w.z * q.dz), -(w.x * q.ox + w.y * q.oy + w.z * q.oz + w.w));
}

// One thread per 4 consecutive samples (K % 4 == 0 keeps them in one ray): 16 B of z (or u) in,
// 48 B of points and 48 B of view directions out.  A thread's 48 bytes are contiguous but the
// warp's stores would interleave at a 48-byte stride (half-filled sectors), so each warp
// transposes its 1536 bytes through shared memory and writes three fully coalesced 512-byte rows
// (the 48-byte stride is conflict-free for 16-byte shared-memory accesses).
// kFromU: `zu` holds the uniforms; z is computed here (and stored) — the fused coarse sampler.
// Ray setup folded into the first launch of a render (renderers.py:166-175): where the rays come from and
// where their origins / directions / camera-depth coefficients go (the fine pass and the adaptive tail read
// them).  Used by rays_coarse_points_kernel (K % 4 == 0) and by the scalar kernel's kSetup variant.
struct RaySetupArgs {
  const float* x_pix;
  const float* intr;
  const float* c2w;
  int64_t rays_per_cam;
  float* ros_out;
  float* rds_out;
  float* affine_out;   // nullable
};

template <bool kFromU>
__global__ void __launch_bounds__(256)
ray_points_vec4_kernel(const float* __restrict__ ros, const float* __restrict__ rds, const float* __restrict__ zu,
                       const float* __restrict__ near, const float* __restrict__ far, int bound_stride,
                       int64_t n_vec, int K, float* __restrict__ z_out, float* __restrict__ pts,
                       float* __restrict__ viewdirs) {
  __shared__ float4 s_stage[8][96];
  const int lane = threadIdx.x & 31;
  float4* sw = s_stage[threadIdx.x >> 5];
  const float kf = (float)K;
  const bool pow2 = (K & (K - 1)) == 0;
  const float inv_k = 1.0f / kf;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // warp-uniform loop: v0 is the warp's first vector, every lane stays until the warp is done
  for (int64_t v0 = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x - lane); v0 < n_vec; v0 += stride) {
    const int64_t v = v0 + lane;
    const bool valid = v < n_vec;
    const int64_t i = (valid ? v : n_vec - 1) * 4;
    const int64_t r = i / K;
    const int j = (int)(i - r * K);
    const Ray3 q = load_ray3(ros, rds, r);
    float4 z4 = ldg_stream(reinterpret_cast<const float4*>(zu + i));
    if (kFromU) {
      const int64_t b = bound_stride ? r : 0;
      const float n0 = near[b];
      const float span = __fsub_rn(far[b], n0);
      z4.x = coarse_z(n0, span, j + 0, kf, inv_k, pow2, z4.x);
      z4.y = coarse_z(n0, span, j + 1, kf, inv_k, pow2, z4.y);
      z4.z = coarse_z(n0, span, j + 2, kf, inv_k, pow2, z4.z);
      z4.w = coarse_z(n0, span, j + 3, kf, inv_k, pow2, z4.w);
      if (valid) stg_stream(reinterpret_cast<float4*>(z_out + i), z4);
    }
    const float3 p0 = point_on_ray(q, z4.x), p1 = point_on_ray(q, z4.y), p2 = point_on_ray(q, z4.z),
                 p3 = point_on_ray(q, z4.w);
    const int64_t n_left = n_vec - v0;
    const int n_f4 = (int)(n_left < 32 ? n_left : 32) * 3;  // float4s this warp writes per output
    sw[lane * 3 + 0] = make_float4(p0.x, p0.y, p0.z, p1.x);
    sw[lane * 3 + 1] = make_float4(p1.y, p1.z, p2.x, p2.y);
    sw[lane * 3 + 2] = make_float4(p2.z, p3.x, p3.y, p3.z);
    __syncwarp();
    float4* po = reinterpret_cast<float4*>(pts + v0 * 12);
#pragma unroll
    for (int t = 0; t < 3; ++t)
      if (t * 32 + lane < n_f4) stg_stream(po + t * 32 + lane, sw[t * 32 + lane]);
    __syncwarp();
    if (viewdirs) {
      sw[lane * 3 + 0] = make_float4(q.dx, q.dy, q.dz, q.dx);
      sw[lane * 3 + 1] = make_float4(q.dy, q.dz, q.dx, q.dy);
      sw[lane * 3 + 2] = make_float4(q.dz, q.dx, q.dy, q.dz);
      __syncwarp();
      float4* vo = reinterpret_cast<float4*>(viewdirs + v0 * 12);
#pragma unroll
      for (int t = 0; t < 3; ++t)
        if (t * 32 + lane < n_f4) stg_stream(vo + t * 32 + lane, sw[t * 32 + lane]);
      __syncwarp();
    }
  }
}

// VolumeRenderer's first launch (renderers.py:166-175) for K % 4 == 0: a warp owns a BLOCK of 32 consecutive rays.
// Lane l forms ray (block*32 + l) once — pixel, intrinsics, pose in; origin, direction, camera-depth coefficients
// out, all coalesced — and the warp then walks the block's 32*K/4 sample vectors, every lane fetching the ray
// its four samples belong to from the lane that formed it (six shuffles) instead of redoing ~150 instructions of
// setup per 128 samples (the per-thread version ran at 0.70 of the roofline; this one is bound by its stores).
__global__ void __launch_bounds__(256)
rays_coarse_points_kernel(const float* __restrict__ u, const float* __restrict__ near, const float* __restrict__ far,
                          int bound_stride, int64_t R, int K, float* __restrict__ z_out, float* __restrict__ pts,
                          float* __restrict__ viewdirs, const RaySetupArgs rs) {
  __shared__ float4 s_stage[8][96];
  const int lane = threadIdx.x & 31;
  float4* sw = s_stage[threadIdx.x >> 5];
  const float kf = (float)K;
  const bool pow2 = (K & (K - 1)) == 0;
  const float inv_k = 1.0f / kf;
  const int vec_per_ray = K >> 2;
  const int64_t n_blocks = (R + 31) >> 5;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t rb = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); rb < n_blocks; rb += warps) {
    const int64_t r0 = rb << 5;
    const int n_rays = (int)(R - r0 < 32 ? R - r0 : 32);
    const int64_t mine = r0 + (lane < n_rays ? lane : n_rays - 1);
    const Ray3 q_mine = ray_from_pixel(rs.x_pix, rs.intr, rs.c2w, mine, rs.rays_per_cam);
    const int64_t bm = bound_stride ? mine : 0;
    const float near_mine = near[bm], span_mine = __fsub_rn(far[bm], near_mine);
    if (lane < n_rays) {
      rs.ros_out[mine * 3 + 0] = q_mine.ox; rs.ros_out[mine * 3 + 1] = q_mine.oy; rs.ros_out[mine * 3 + 2] = q_mine.oz;
      rs.rds_out[mine * 3 + 0] = q_mine.dx; rs.rds_out[mine * 3 + 1] = q_mine.dy; rs.rds_out[mine * 3 + 2] = q_mine.dz;
      if (rs.affine_out) {
        const float2 ab = depth_affine_of(q_mine, pose_inverse_row2(rs.c2w, mine));
        rs.affine_out[mine * 2 + 0] = ab.x;
        rs.affine_out[mine * 2 + 1] = ab.y;
      }
    }
    const int n_vec = n_rays * vec_per_ray;               // sample vectors of this block
    const int64_t vbase = r0 * vec_per_ray;
    for (int v0 = 0; v0 < n_vec; v0 += 32) {
      const int v = v0 + lane;
      const bool valid = v < n_vec;
      const int vc = valid ? v : n_vec - 1;
      const int lr = vc / vec_per_ray;                     // ray of the block: the lane that formed it
      const int j = (vc - lr * vec_per_ray) << 2;
      Ray3 q;
      q.ox = __shfl_sync(0xffffffffu, q_mine.ox, lr); q.oy = __shfl_sync(0xffffffffu, q_mine.oy, lr);
      q.oz = __shfl_sync(0xffffffffu, q_mine.oz, lr); q.dx = __shfl_sync(0xffffffffu, q_mine.dx, lr);
      q.dy = __shfl_sync(0xffffffffu, q_mine.dy, lr); q.dz = __shfl_sync(0xffffffffu, q_mine.dz, lr);
      const float n0 = __shfl_sync(0xffffffffu, near_mine, lr), span = __shfl_sync(0xffffffffu, span_mine, lr);
      const int64_t i = (vbase + vc) << 2;
      float4 z4 = ldg_stream(reinterpret_cast<const float4*>(u + i));
      z4.x = coarse_z(n0, span, j + 0, kf, inv_k, pow2, z4.x);
      z4.y = coarse_z(n0, span, j + 1, kf, inv_k, pow2, z4.y);
      z4.z = coarse_z(n0, span, j + 2, kf, inv_k, pow2, z4.z);
      z4.w = coarse_z(n0, span, j + 3, kf, inv_k, pow2, z4.w);
      if (valid) stg_stream(reinterpret_cast<float4*>(z_out + i), z4);
      const float3 p0 = point_on_ray(q, z4.x), p1 = point_on_ray(q, z4.y), p2 = point_on_ray(q, z4.z),
                   p3 = point_on_ray(q, z4.w);
      const int n_left = n_vec - v0;
      const int n_f4 = (n_left < 32 ? n_left : 32) * 3;    // float4s this warp writes per output
      sw[lane * 3 + 0] = make_float4(p0.x, p0.y, p0.z, p1.x);
      sw[lane * 3 + 1] = make_float4(p1.y, p1.z, p2.x, p2.y);
      sw[lane * 3 + 2] = make_float4(p2.z, p3.x, p3.y, p3.z);
      __syncwarp();
      float4* po = reinterpret_cast<float4*>(pts + (vbase + v0) * 12);
#pragma unroll
      for (int t = 0; t < 3; ++t)
        if (t * 32 + lane < n_f4) stg_stream(po + t * 32 + lane, sw[t * 32 + lane]);
      __syncwarp();
      if (viewdirs) {
        sw[lane * 3 + 0] = make_float4(q.dx, q.dy, q.dz, q.dx);
        sw[lane * 3 + 1] = make_float4(q.dy, q.dz, q.dx, q.dy);
        sw[lane * 3 + 2] = make_float4(q.dz, q.dx, q.dy, q.dz);
        __syncwarp();
        float4* vo = reinterpret_cast<float4*>(viewdirs + (vbase + v0) * 12);
#pragma unroll
        for (int t = 0; t < 3; ++t)
          if (t * 32 + lane < n_f4) stg_stream(vo + t * 32 + lane, sw[t * 32 + lane]);
        __syncwarp();
      }
    }
  }
}

// any K / alignment: one thread per sample
template <bool kFromU, bool kSetup>
__global__ void __launch_bounds__(256)
ray_points_scalar_kernel(const float* __restrict__ ros, const float* __restrict__ rds, const float* __restrict__ zu,
                         const float* __restrict__ near, const float* __restrict__ far, int bound_stride,
                         int64_t total, int K, float* __restrict__ z_out, float* __restrict__ pts,
                         float* __restrict__ viewdirs, const RaySetupArgs rs) {
  const float kf = (float)K;
  const bool pow2 = (K & (K - 1)) == 0;
  const float inv_k = 1.0f / kf;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / K;
    Ray3 q;
    if (kSetup) {
      q = ray_from_pixel(rs.x_pix, rs.intr, rs.c2w, r, rs.rays_per_cam);
      if (i == r * K) {
        rs.ros_out[r * 3 + 0] = q.ox; rs.ros_out[r * 3 + 1] = q.oy; rs.ros_out[r * 3 + 2] = q.oz;
        rs.rds_out[r * 3 + 0] = q.dx; rs.rds_out[r * 3 + 1] = q.dy; rs.rds_out[r * 3 + 2] = q.dz;
        if (rs.affine_out) {
          const float2 ab = depth_affine_of(q, pose_inverse_row2(rs.c2w, r));
          rs.affine_out[r * 2 + 0] = ab.x;
          rs.affine_out[r * 2 + 1] = ab.y;
        }
      }
    } else {
      q = load_ray3(ros, rds, r);
    }
    float z = zu[i];
    if (kFromU) {
      const int64_t b = bound_stride ? r : 0;
      const float n0 = near[b];
      z = coarse_z(n0, __fsub_rn(far[b], n0), (int)(i - r * K), kf, inv_k, pow2, z);
      z_out[i] = z;
    }
    const float3 p = point_on_ray(q, z);
    pts[i * 3 + 0] = p.x; pts[i * 3 + 1] = p.y; pts[i * 3 + 2] = p.z;
    if (viewdirs) {
      viewdirs[i * 3 + 0] = q.dx; viewdirs[i * 3 + 1] = q.dy; viewdirs[i * 3 + 2] = q.dz;
    }
  }
}

// Packed (ragged) rays: ray r owns samples [offsets[r], offsets[r+1]).  One warp per ray, lanes over
// its samples; a warp instruction writes 32 consecutive 12-byte points = 384 contiguous bytes.
// g_pts == nullptr: forward (pts, viewdirs); otherwise backward (d_z = g_pts . rds).
__global__ void __launch_bounds__(128)
ray_points_packed_kernel(const float* __restrict__ ros, const float* __restrict__ rds, const float* __restrict__ z,
                         const int64_t* __restrict__ offsets, int64_t R, float* __restrict__ pts,
                         float* __restrict__ viewdirs, const float* __restrict__ g_pts, float* __restrict__ d_z) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < R; r += warps) {
    const int64_t begin = offsets[r], end = offsets[r + 1];
    const Ray3 q = load_ray3(ros, rds, r);
    for (int64_t i = begin + lane; i < end; i += 32) {
      if (g_pts) {
        d_z[i] = g_pts[i * 3 + 0] * q.dx + g_pts[i * 3 + 1] * q.dy + g_pts[i * 3 + 2] * q.dz;
      } else {
        const float3 p = point_on_ray(q, z[i]);
        pts[i * 3 + 0] = p.x; pts[i * 3 + 1] = p.y; pts[i * 3 + 2] = p.z;
        if (viewdirs) {
          viewdirs[i * 3 + 0] = q.dx; viewdirs[i * 3 + 1] = q.dy; viewdirs[i * 3 + 2] = q.dz;
        }
      }
    }
  }
}

// d_z[r,k] = g_pts[r,k,:] . rds[r,:]   (AdaptiveVolumeRenderer: the sample depths carry grad)
__global__ void __launch_bounds__(256)
ray_points_bwd_kernel(const float* __restrict__ rds, const float* __restrict__ g_pts, int64_t total, int K,
                      float* __restrict__ d_z) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / K;
    const float gx = g_pts[i * 3 + 0], gy = g_pts[i * 3 + 1], gz = g_pts[i * 3 + 2];
    d_z[i] = gx * rds[r * 3 + 0] + gy * rds[r * 3 + 1] + gz * rds[r * 3 + 2];
  }
}

// utils.get_world_rays as a kernel of its own (the adaptive renderer marches before it samples),
// optionally with the camera-depth coefficients of every ray.
__global__ void __launch_bounds__(256)
world_rays_kernel(const float* __restrict__ x_pix, const float* __restrict__ intr, const float* __restrict__ c2w,
                  int64_t R, int64_t rays_per_cam, float* __restrict__ ros, float* __restrict__ rds,
                  float* __restrict__ affine) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += stride) {
    const Ray3 q = ray_from_pixel(x_pix, intr, c2w, r, rays_per_cam);
    ros[r * 3 + 0] = q.ox; ros[r * 3 + 1] = q.oy; ros[r * 3 + 2] = q.oz;
    rds[r * 3 + 0] = q.dx; rds[r * 3 + 1] = q.dy; rds[r * 3 + 2] = q.dz;
    if (affine) {
      const float2 ab = depth_affine_of(q, pose_inverse_row2(c2w, r));
      affine[r * 2 + 0] = ab.x;
      affine[r * 2 + 1] = ab.y;
    }
  }
}

// utils.depth_from_world on the composited point p = o + d*dist (renderers.py:274-275, 508-509),
// or on given points (dist == nullptr: `ros` holds the points; renderers.py:486):
// depth = -(pose^-1 [p,1])_z.  Only row 2 of the inverse is needed: cofactors of the 4x4.
__global__ void __launch_bounds__(256)
depth_from_world_kernel(const float* __restrict__ ros, const float* __restrict__ rds, const float* __restrict__ dist,
                        const float* __restrict__ c2w, int64_t R, float* __restrict__ depth,
                        float* __restrict__ grad_row) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += stride) {
    float px = ros[r * 3 + 0], py = ros[r * 3 + 1], pz = ros[r * 3 + 2];
    if (dist) {
      const float t = dist[r];
      px = __fadd_rn(px, __fmul_rn(rds[r * 3 + 0], t));
      py = __fadd_rn(py, __fmul_rn(rds[r * 3 + 1], t));
      pz = __fadd_rn(pz, __fmul_rn(rds[r * 3 + 2], t));
    }
    const float4* mp = reinterpret_cast<const float4*>(c2w + r * 16);
    const float4 a = mp[0], b = mp[1], c = mp[2], d = mp[3];
    // inverse(M)[2][j] = cofactor(M)[j][2] / det(M); the cofactors C_j2 delete column 2 of M
    // 3x3 minors over columns (0,1,3):
    auto det3 = [](float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
      return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0);
    };
    const float c02 = det3(b.x, b.y, b.w, c.x, c.y, c.w, d.x, d.y, d.w);    // delete row 0, col 2, sign +
    const float c12 = -det3(a.x, a.y, a.w, c.x, c.y, c.w, d.x, d.y, d.w);   // row 1, sign -
    const float c22 = det3(a.x, a.y, a.w, b.x, b.y, b.w, d.x, d.y, d.w);    // row 2, sign +
    const float c32 = -det3(a.x, a.y, a.w, b.x, b.y, b.w, c.x, c.y, c.w);   // row 3, sign -
    // det(M) by expansion along column 2
    const float det = a.z * c02 + b.z * c12 + c.z * c22 + d.z * c32;
    const float zc = (c02 * px + c12 * py + c22 * pz + c32) / det;
    depth[r] = -zc;
    if (grad_row) {  // d depth / d p: what the backward pass multiplies by (depth is affine in p)
      grad_row[r * 3 + 0] = -(c02 / det);
      grad_row[r * 3 + 1] = -(c12 / det);
      grad_row[r * 3 + 2] = -(c22 / det);
    }
  }
}

static int grid_1d(int64_t work, int max_blocks) {
  int64_t b = (work + 255) / 256;
  if (b < 1) b = 1;
  return (int)(b < max_blocks ? b : max_blocks);
}

int launch_ray_points(const float* ros, const float* rds, const float* z_or_u, const float* near, const float* far,
                      int bound_stride, bool from_u, int64_t R, int K, float* z_out, float* pts, float* viewdirs,
                      cudaStream_t stream) {
  const int64_t total = R * (int64_t)K;
  if (total == 0) return AVR_OK;
  const bool vec = (K % 4 == 0) && aligned16(z_or_u) && aligned16(pts) && aligned16(viewdirs) && aligned16(z_out);
  const int max_blocks = grid_cap(16);
  const RaySetupArgs none{};
  if (vec) {
    const int64_t n_vec = total / 4;
    // NO grid cap here: with one vector per thread the CTAs in flight cover a compact, moving window of the output
    // streams; under a grid-stride loop every CTA's next vector lies grid*256 vectors further on and the stores of an
    // SM scatter over the whole array (K = 96: 0.473 ms at 16 CTAs per SM, 0.455 / 0.444 / 0.434 at 64 / 128 / 256,
    // 0.423 = 1.04 of the copy roofline without a cap; the backward kernel and the dense coarse sampler, which set up
    // per-CTA state, are the other way round)
    const int vec_blocks = grid_1d(n_vec, 1 << 24);
    if (from_u) {
      ray_points_vec4_kernel<true><<<vec_blocks, 256, 0, stream>>>(
          ros, rds, z_or_u, near, far, bound_stride, n_vec, K, z_out, pts, viewdirs);
    } else {
      ray_points_vec4_kernel<false><<<vec_blocks, 256, 0, stream>>>(
          ros, rds, z_or_u, near, far, bound_stride, n_vec, K, z_out, pts, viewdirs);
    }
  } else if (from_u) {
    ray_points_scalar_kernel<true, false><<<grid_1d(total, max_blocks), 256, 0, stream>>>(
        ros, rds, z_or_u, near, far, bound_stride, total, K, z_out, pts, viewdirs, none);
  } else {
    ray_points_scalar_kernel<false, false><<<grid_1d(total, max_blocks), 256, 0, stream>>>(
        ros, rds, z_or_u, near, far, bound_stride, total, K, z_out, pts, viewdirs, none);
  }
  return check_launch();
}

int launch_rays_coarse_points(const float* x_pix, const float* intr, const float* c2w, int64_t rays_per_cam,
                              const float* near, const float* far, int bound_stride, const float* u, int64_t R, int K,
                              float* ros, float* rds, float* depth_affine, float* z, float* pts, float* viewdirs,
                              cudaStream_t stream) {
  const int64_t total = R * (int64_t)K;
  if (total == 0) return AVR_OK;
  const bool vec = (K % 4 == 0) && aligned16(u) && aligned16(pts) && aligned16(viewdirs) && aligned16(z);
  const int max_blocks = grid_cap(16);
  const RaySetupArgs rs{x_pix, intr, c2w, rays_per_cam, ros, rds, depth_affine};
  // Large batches: the ray setup as its own small launch and the sample-parallel kernel after it (same per-ray code,
  // same bits).  One thread per four samples with no grid cap streams its outputs at 1.04 of the copy roofline; the
  // one-launch kernel below, a warp per block of 32 rays, stays at 0.89 (2^20 rays x 64: 0.390 ms against
  // 0.024 + 0.324).  The single launch is what a training batch or a frame wants.
  if (vec && total >= ((int64_t)1 << 23)) {
    const int rc = launch_world_rays(x_pix, intr, c2w, R, rays_per_cam, ros, rds, depth_affine, stream);
    if (rc != AVR_OK) return rc;
    return launch_ray_points(ros, rds, u, near, far, bound_stride, true, R, K, z, pts, viewdirs, stream);
  }
  if (vec) {
    const int64_t n_blocks = (R + 31) / 32;   // one warp per block of 32 rays
    int64_t blocks = (n_blocks + 7) / 8;
    if (blocks > max_blocks) blocks = max_blocks;
    rays_coarse_points_kernel<<<(unsigned)blocks, 256, 0, stream>>>(u, near, far, bound_stride, R, K, z, pts, viewdirs, rs);
  } else {
    ray_points_scalar_kernel<true, true><<<grid_1d(total, max_blocks), 256, 0, stream>>>(
        nullptr, nullptr, u, near, far, bound_stride, total, K, z, pts, viewdirs, rs);
  }
  return check_launch();
}

int launch_ray_points_bwd(const float* rds, const float* g_pts, int64_t R, int K, float* d_z, cudaStream_t stream) {
  const int64_t total = R * (int64_t)K;
  if (total == 0) return AVR_OK;
  ray_points_bwd_kernel<<<grid_1d(total, grid_cap(16)), 256, 0, stream>>>(rds, g_pts, total, K, d_z);
  return check_launch();
}

int launch_ray_points_packed(const float* ros, const float* rds, const float* z, const int64_t* offsets, int64_t R,
                             float* pts, float* viewdirs, const float* g_pts, float* d_z, cudaStream_t stream) {
  if (R == 0) return AVR_OK;
  int64_t blocks = (R + 3) / 4;
  if (blocks > grid_cap(16)) blocks = grid_cap(16);
  ray_points_packed_kernel<<<(unsigned)blocks, 128, 0, stream>>>(ros, rds, z, offsets, R, pts, viewdirs, g_pts, d_z);
  return check_launch();
}

int launch_world_rays(const float* x_pix, const float* kinv, const float* c2w, int64_t R, int64_t rays_per_cam,
                      float* ros, float* rds, float* depth_affine, cudaStream_t stream) {
  if (R == 0) return AVR_OK;
  world_rays_kernel<<<grid_1d(R, grid_cap(16)), 256, 0, stream>>>(x_pix, kinv, c2w, R, rays_per_cam, ros, rds,
                                                                     depth_affine);
  return check_launch();
}

int launch_depth_from_world(const float* ros, const float* rds, const float* dist, const float* c2w, int64_t R,
                            float* depth, float* grad_row, cudaStream_t stream) {
  if (R == 0) return AVR_OK;
  depth_from_world_kernel<<<grid_1d(R, grid_cap(16)), 256, 0, stream>>>(ros, rds, dist, c2w, R, depth, grad_row);
  return check_launch();
}

}  // namespace avr
