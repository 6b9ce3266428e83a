// Depth samplers: stratified coarse sampling (+ its gradient), inverse-CDF importance
// sampling with the per-ray merge sort, and a stand-alone per-ray sort.
//
// Reference: sample_coarse renderers.py:4-24, sample_fine :27-54, sample_depth :56-66,
// clamp/cat/sort :255-258, sort :494.  The samplers reproduce the reference's operation
// order with explicitly rounded intrinsics (no FMA contraction) so that, given the same
// draws (and, for the importance sampler, the same CDF), depths are bit-identical to
// torch-CPU.
#include <math_constants.h>

#include <cstdlib>

#include "avr_common.cuh"
#include "coarse_packed_core.h"
#include "kernels.h"

namespace avr {

// grid cap of the grid-stride kernels: `mult` CTAs per SM
static inline int grid_cap(int mult) { return num_sms() * mult; }


constexpr int kSamplerWarps = 4;

// z = near + (far-near)*(j/K) + (u*(far-near))/K       renderers.py:12-14
__device__ __forceinline__ float coarse_depth(float near, float far, int j, int K, float u) {
  const float span = __fsub_rn(far, near);
  const float bin = __fdiv_rn((float)j, (float)K);
  const float z = __fadd_rn(near, __fmul_rn(span, bin));
  return __fadd_rn(z, __fdiv_rn(__fmul_rn(u, span), (float)K));
}

// Dense layout, one thread per 4 consecutive samples (16-byte loads/stores; K % 4 == 0 keeps
// the four in one ray).  The per-bin term (far-near)*(j/K) only needs j/K, which is
// tabulated once per block in shared memory; dividing by a power-of-two K is done as an
// exact multiply (bit-identical).
template <bool kVec4>
__global__ void __launch_bounds__(256)
coarse_fwd_dense_kernel(const float* __restrict__ near, const float* __restrict__ far, int bound_stride,
                        const float* __restrict__ u, int64_t total, int K, float* __restrict__ z) {
  extern __shared__ float s_bins[];  // j / K
  for (int j = threadIdx.x; j < K; j += blockDim.x) s_bins[j] = __fdiv_rn((float)j, (float)K);
  __syncthreads();
  const float kf = (float)K;
  const bool pow2 = (K & (K - 1)) == 0;
  const float inv_k = 1.0f / kf;
  constexpr int V = kVec4 ? 4 : 1;
  const int64_t n_vec = total / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float n0 = near[0], f0 = far[0];
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_vec; v += stride) {
    const int64_t i = v * V;
    const int64_t r = i / K;
    const int j = (int)(i - r * K);
    if (bound_stride) {
      n0 = near[r];
      f0 = far[r];
    }
    const float span = __fsub_rn(f0, n0);
    float uu[V], out[V];
    if (kVec4) {
      const float4 q = ldg_stream(reinterpret_cast<const float4*>(u + i));
      uu[0] = q.x; uu[V > 1 ? 1 : 0] = q.y; uu[V > 2 ? 2 : 0] = q.z; uu[V > 3 ? 3 : 0] = q.w;
    } else {
      uu[0] = u[i];
    }
#pragma unroll
    for (int q = 0; q < V; ++q) {
      const float base = __fadd_rn(n0, __fmul_rn(span, s_bins[j + q]));
      const float jit = __fmul_rn(uu[q], span);
      out[q] = __fadd_rn(base, pow2 ? __fmul_rn(jit, inv_k) : __fdiv_rn(jit, kf));
    }
    if (kVec4) {
      stg_stream(reinterpret_cast<float4*>(z + i), make_float4(out[0], out[V > 1 ? 1 : 0], out[V > 2 ? 2 : 0], out[V > 3 ? 3 : 0]));
    } else {
      z[i] = out[0];
    }
  }
}

// packed, flat: a warp owns 32 consecutive rays = one contiguous slice of the packed streams and
// walks it four samples per lane at a time (coarse_packed_core.h has the per-lane code and the
// reasoning; the same code runs on the host in the CPU test suite)
#ifndef AVR_COARSE_FLAT_BLOCKS
#define AVR_COARSE_FLAT_BLOCKS 8  // resident CTAs per SM the register budget is cut for (8 x 4 warps: 64 registers)
#endif
__global__ void __launch_bounds__(kSamplerWarps * 32, AVR_COARSE_FLAT_BLOCKS)
coarse_fwd_packed_flat_kernel(const float* __restrict__ near, const float* __restrict__ far, int bound_stride,
                              const float* __restrict__ u, const int64_t* __restrict__ offsets, int64_t R,
                              float* __restrict__ z, int vec_ok) {
  __shared__ __align__(16) CoarseSegment s_seg[kSamplerWarps];
  const int lane = threadIdx.x & 31;
  CoarseSegment* seg = &s_seg[threadIdx.x >> 5];
  const int64_t n_seg = (R + kSegRays - 1) / kSegRays;
  const int64_t warps = (int64_t)gridDim.x * kSamplerWarps;
  for (int64_t sidx = blockIdx.x * (int64_t)kSamplerWarps + (threadIdx.x >> 5); sidx < n_seg; sidx += warps) {
    const int64_t r0 = sidx * kSegRays;
    const bool ok = coarse_segment_build(lane, r0, R, offsets, near, far, bound_stride, seg);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    __syncwarp();  // table writes before the reads
    if (all_ok) {
      coarse_segment_run(lane, seg, offsets[r0], u, z, vec_ok != 0);
    } else {
      coarse_segment_slow(lane, r0, R, offsets, near, far, bound_stride, u, z);
    }
    __syncwarp();  // the next segment's build overwrites the tables
  }
}

// packed: one warp per ray, each ray stratified over its own count (AVR_COARSE_PACKED=ray)
__global__ void __launch_bounds__(kSamplerWarps * 32)
coarse_fwd_packed_kernel(const float* __restrict__ near, const float* __restrict__ far, int bound_stride,
                         const float* __restrict__ u, const int64_t* __restrict__ offsets, int64_t R,
                         float* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * kSamplerWarps;
  for (int64_t r = blockIdx.x * (int64_t)kSamplerWarps + (threadIdx.x >> 5); r < R; r += warps) {
    const int64_t begin = offsets[r];
    const int cnt = (int)(offsets[r + 1] - begin);
    const int64_t b = bound_stride ? r : 0;
    const float nr = near[b], fr = far[b];
    for (int j = lane; j < cnt; j += 32) z[begin + j] = coarse_depth(nr, fr, j, cnt, u[begin + j]);
  }
}

// d_near = sum_j g_j (1 - f_j), d_far = sum_j g_j f_j with f_j = j/K + u_j/K
__global__ void __launch_bounds__(kSamplerWarps * 32)
coarse_bwd_kernel(const float* __restrict__ g_z, const float* __restrict__ u, int64_t R, int K,
                  float* __restrict__ d_near, float* __restrict__ d_far) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * kSamplerWarps;
  for (int64_t r = blockIdx.x * (int64_t)kSamplerWarps + (threadIdx.x >> 5); r < R; r += warps) {
    float dn = 0.f, df = 0.f;
    for (int j = lane; j < K; j += 32) {
      const float g = g_z[r * K + j];
      const float f = (float)j / (float)K + u[r * K + j] / (float)K;
      df += g * f;
      dn += g * (1.0f - f);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      dn += __shfl_xor_sync(0xffffffffu, dn, d);
      df += __shfl_xor_sync(0xffffffffu, df, d);
    }
    if (lane == 0) {
      d_near[r] = dn;
      d_far[r] = df;
    }
  }
}

// Short rays (the adaptive renderer's K = 20): one THREAD per ray, 16-byte loads of its own row
// (a warp's 32 rows are one contiguous block, so every sector is used), no shuffles, no idle lanes.
template <int V>
__global__ void __launch_bounds__(256)
coarse_bwd_thread_kernel(const float* __restrict__ g_z, const float* __restrict__ u, int64_t R, int K,
                         float* __restrict__ d_near, float* __restrict__ d_far) {
  const float inv_k = 1.0f / (float)K;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += stride) {
    const float* gr = g_z + r * K;
    const float* ur = u + r * K;
    float dn = 0.f, df = 0.f;
    for (int j = 0; j < K; j += V) {
      float gv[V], uv[V];
      if (V == 4) {
        const float4 a = *reinterpret_cast<const float4*>(gr + j), b = *reinterpret_cast<const float4*>(ur + j);
        gv[0] = a.x; gv[V > 1 ? 1 : 0] = a.y; gv[V > 2 ? 2 : 0] = a.z; gv[V > 3 ? 3 : 0] = a.w;
        uv[0] = b.x; uv[V > 1 ? 1 : 0] = b.y; uv[V > 2 ? 2 : 0] = b.z; uv[V > 3 ? 3 : 0] = b.w;
      } else {
        gv[0] = gr[j];
        uv[0] = ur[j];
      }
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float f = (float)(j + q) * inv_k + uv[q] * inv_k;
        df += gv[q] * f;
        dn += gv[q] * (1.0f - f);
      }
    }
    d_near[r] = dn;
    d_far[r] = df;
  }
}

// ---- warp-synchronous bitonic sort of P (power of two) keys in shared memory --------
// With `idx` non-null the (key, index) pairs are ordered lexicographically, which makes
// the result the stable sort of the keys.
__device__ __forceinline__ void warp_bitonic_sort(float* key, int* idx, int P, int lane) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = lane; t < (P >> 1); t += 32) {
        const int lo = 2 * t - (t & (stride - 1));  // index with the `stride` bit clear
        const int hi = lo + stride;
        const bool ascending = ((lo & size) == 0);
        const float a = key[lo], b = key[hi];
        bool swap;
        if (idx) {
          const int ia = idx[lo], ib = idx[hi];
          const bool a_gt_b = (a > b) || (a == b && ia > ib);
          swap = (a_gt_b == ascending);
          if (swap) {
            idx[lo] = ib;
            idx[hi] = ia;
          }
        } else {
          swap = ((a > b) == ascending) && (a != b);
        }
        if (swap) {
          key[lo] = b;
          key[hi] = a;
        }
      }
    }
  }
  __syncwarp();
}

__host__ __device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- importance sampling + merge ----------------------------------------------------
struct ImportanceArgs {
  const float* weights;
  const float* z_coarse;
  const float* u;
  const float* u2;
  const float* normals;
  const float* near;
  const float* far;
  int bound_stride;
  const int64_t* offsets;       // packed coarse layout or null
  const int64_t* fine_offsets;  // packed fine layout or null
  int64_t R;
  int Kc, n_imp, n_depth;
  float depth_std;
  float* z_fine;
  float* z_sorted;
  float* cdf;
  int32_t* idx;
  int cdf_cap;   // floats reserved per warp for the cdf
  int sort_cap;  // floats reserved per warp for the sort buffer (power of two)
};

__global__ void __launch_bounds__(kSamplerWarps * 32)
importance_kernel(const ImportanceArgs a) {
  extern __shared__ __align__(16) float fsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* cdf = fsmem + warp * (a.cdf_cap + a.sort_cap);
  float* sortbuf = cdf + a.cdf_cap;
  const int64_t warps = (int64_t)gridDim.x * kSamplerWarps;

  for (int64_t r = blockIdx.x * (int64_t)kSamplerWarps + warp; r < a.R; r += warps) {
    int64_t cbase, fbase;
    int kc, n;
    if (a.offsets) {
      cbase = a.offsets[r];
      kc = (int)(a.offsets[r + 1] - cbase);
      fbase = a.fine_offsets[r];
      n = (int)(a.fine_offsets[r + 1] - fbase);
    } else {
      cbase = r * (int64_t)a.Kc;
      kc = a.Kc;
      fbase = r * (int64_t)a.n_imp;
      n = a.n_imp;
    }
    const int nd = a.n_depth;
    const int64_t bi = a.bound_stride ? r : 0;
    const float near = a.near[bi], far = a.far[bi];
    const float span = __fsub_rn(far, near);

    // (a) S = sum_j (w_j + 1e-5)                                           renderers.py:36-37
    float part = 0.f;
    for (int j = lane; j < kc; j += 32) part += __fadd_rn(a.weights[cbase + j], kPdfEps);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    const float S = part;

    // (b) cdf_0 = 0, cdf_{j+1} = cdf_j + pdf_j: 32-wide shuffle scan with a running carry,
    // then a running max so the table is non-decreasing by construction (a parallel
    // prefix sum is not monotone in floating point; the search below requires it).
    __syncwarp();
    if (lane == 0) cdf[0] = 0.f;
    float carry = 0.f, hi = 0.f;
    for (int j0 = 0; j0 < kc; j0 += 32) {
      const int j = j0 + lane;
      float v = (j < kc) ? __fdiv_rn(__fadd_rn(a.weights[cbase + j], kPdfEps), S) : 0.f;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float p = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += p;
      }
      v += carry;
      float m = fmaxf(v, hi);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float p = __shfl_up_sync(0xffffffffu, m, d);
        if (lane >= d) m = fmaxf(m, p);
      }
      if (j < kc) cdf[j + 1] = m;
      carry = __shfl_sync(0xffffffffu, v, 31);
      hi = __shfl_sync(0xffffffffu, m, 31);
    }
    __syncwarp();
    if (a.cdf) {
      float* out = a.cdf + (cbase + r);  // kc+1 entries per ray: dense r*(Kc+1), packed offsets[r]+r
      for (int j = lane; j <= kc; j += 32) out[j] = cdf[j];
    }

    // (c) inverse-CDF search + in-bin jitter                              renderers.py:41-46
    const int total = kc + n + nd;
    const int P = next_pow2(total);
    const bool do_sort = (a.z_sorted != nullptr);
    for (int i = lane; i < n; i += 32) {
      const float uu = a.u[fbase + i];
      // number of entries <= u among cdf[0..kc]  (searchsorted right=True)
      int lo = 0, hi_i = kc + 1;
      while (lo < hi_i) {
        const int mid = (lo + hi_i) >> 1;
        if (cdf[mid] <= uu) lo = mid + 1; else hi_i = mid;
      }
      int bin = lo - 1;
      if (bin < 0) bin = 0;
      const float t = __fdiv_rn(__fadd_rn((float)bin, a.u2[fbase + i]), (float)kc);
      const float zf = __fadd_rn(near, __fmul_rn(span, t));
      if (a.idx) a.idx[fbase + i] = bin;
      if (a.z_fine) a.z_fine[fbase + i] = zf;
      if (do_sort) sortbuf[kc + i] = zf;
    }
    if (do_sort) {
      for (int j = lane; j < kc; j += 32) sortbuf[j] = a.z_coarse[cbase + j];
      // sample_depth + clamp: normals*std (depth NOT added), clamped to [near, far]   :62-66, :255
      for (int i = lane; i < nd; i += 32) {
        const float v = __fmul_rn(a.normals[r * (int64_t)nd + i], a.depth_std);
        sortbuf[kc + n + i] = fminf(fmaxf(v, near), far);
      }
      for (int i = total + lane; i < P; i += 32) sortbuf[i] = CUDART_INF_F;
      warp_bitonic_sort(sortbuf, nullptr, P, lane);
      float* out = a.z_sorted + (cbase + fbase + r * (int64_t)nd);
      for (int i = lane; i < total; i += 32) out[i] = sortbuf[i];
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kSamplerWarps * 32)
sort_rays_kernel(const float* __restrict__ z_in, int64_t R, int K, int P, float* __restrict__ z_out,
                 int32_t* __restrict__ perm) {
  extern __shared__ __align__(16) float fsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* key = fsmem + warp * 2 * P;
  int* idx = reinterpret_cast<int*>(key + P);
  const int64_t warps = (int64_t)gridDim.x * kSamplerWarps;
  for (int64_t r = blockIdx.x * (int64_t)kSamplerWarps + warp; r < R; r += warps) {
    for (int i = lane; i < P; i += 32) {
      key[i] = (i < K) ? z_in[r * K + i] : CUDART_INF_F;
      idx[i] = i;
    }
    warp_bitonic_sort(key, idx, P, lane);
    for (int i = lane; i < K; i += 32) {
      z_out[r * K + i] = key[i];
      if (perm) perm[r * K + i] = idx[i];
    }
    __syncwarp();
  }
}

// Short rays that are (almost always) ALREADY ascending — the adaptive renderer sorts stratified
// depths (renderers.py:492-494), so the sort is the identity and only routes gradients.  A warp
// owns 32 consecutive rays = ONE contiguous stream of 32*K floats and walks it 16 bytes per lane and
// step, lanes on consecutive pieces (a piece lies inside one ray: K % 4 == 0): every load and store
// instruction covers 512 contiguous bytes.  (One thread per ray with 16-byte accesses at a stride of
// 4K bytes cost K/4 * 20 LSU wavefronts per instruction slot instead of 4: 0.087 ms for 2^20 x 20.)
// A piece checks its own four keys and its last key against the next piece's first unless it ends
// its ray; an ascending ray is copied and gets the identity permutation, the rare unsorted ray is
// then sorted by the whole warp with the shared-memory network above (stable, like
// torch.sort(stable=True)) and overwrites its row.
__global__ void __launch_bounds__(kSamplerWarps * 32)
sort_rays_presorted_kernel(const float* __restrict__ z_in, int64_t R, int K, int P, float* __restrict__ z_out,
                           int32_t* __restrict__ perm) {
  extern __shared__ __align__(16) float fsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* key = fsmem + warp * 2 * P;
  int* idx = reinterpret_cast<int*>(key + P);
  const int64_t warps = (int64_t)gridDim.x * kSamplerWarps;
  const int ppr = K >> 2;  // pieces per ray
  for (int64_t r0 = (blockIdx.x * (int64_t)kSamplerWarps + warp) * 32; r0 < R; r0 += warps * 32) {
    const int nrays = (int)(R - r0 < 32 ? R - r0 : 32);
    const int npieces = nrays * ppr;
    const float4* in4 = reinterpret_cast<const float4*>(z_in + r0 * K);
    float4* out4 = reinterpret_cast<float4*>(z_out + r0 * K);
    int4* perm4 = perm ? reinterpret_cast<int4*>(perm + r0 * K) : nullptr;
    unsigned bad_rays = 0u;
    // ray and in-ray position of this lane's piece, advanced by 32 pieces per step without a division per step
    int ray = lane / ppr, pos = lane - ray * ppr;
    const int step_rays = 32 / ppr, step_pos = 32 - step_rays * ppr;
    float4 q = lane < npieces ? in4[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int f = lane; f - lane < npieces; f += 32) {
      const bool have = f < npieces;
      const float4 nq = (f + 32 < npieces) ? in4[f + 32] : make_float4(0.f, 0.f, 0.f, 0.f);
      // the next piece's first key: lane + 1 of this step, or lane 0 of the next step
      float nx = __shfl_down_sync(0xffffffffu, q.x, 1);
      const float nx_wrap = __shfl_sync(0xffffffffu, nq.x, 0);
      if (lane == 31) nx = nx_wrap;
      // strictly "not descending": ties keep their order under a stable sort, so they are fine
      bool ok = (q.x <= q.y) && (q.y <= q.z) && (q.z <= q.w);
      if (pos + 1 < ppr) ok = ok && (q.w <= nx);
      if (have) {
        out4[f] = q;                       // overwritten below if the ray turns out unsorted
        if (perm4) perm4[f] = make_int4(4 * pos, 4 * pos + 1, 4 * pos + 2, 4 * pos + 3);
      }
      bad_rays |= __reduce_or_sync(0xffffffffu, (have && !ok) ? (1u << ray) : 0u);
      q = nq;
      ray += step_rays;
      pos += step_pos;
      if (pos >= ppr) {
        pos -= ppr;
        ++ray;
      }
    }
    unsigned todo = bad_rays;
    while (todo) {                          // NaNs land here too (comparisons fail) and take the network
      const int b = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t rr = r0 + b;
      __syncwarp();
      for (int i = lane; i < P; i += 32) {
        key[i] = (i < K) ? z_in[rr * K + i] : CUDART_INF_F;
        idx[i] = i;
      }
      warp_bitonic_sort(key, idx, P, lane);
      for (int i = lane; i < K; i += 32) {
        z_out[rr * K + i] = key[i];
        if (perm) perm[rr * K + i] = idx[i];
      }
    }
    __syncwarp();
  }
}

// gradient routing of the sort: d_in[r, perm[r,k]] = g_out[r,k]   (torch.sort's backward, renderers.py:494)
__global__ void __launch_bounds__(256)
sort_rays_bwd_kernel(const float* __restrict__ g_out, const int32_t* __restrict__ perm, int64_t total, int K,
                     float* __restrict__ d_in) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / K;
    d_in[r * K + perm[i]] = g_out[i];
  }
}

// ---- launchers ------------------------------------------------------------------------
static int grid_for(int64_t work_items, int per_block, int max_blocks) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (int)(b < max_blocks ? b : max_blocks);
}

int launch_coarse_fwd(const float* near, const float* far, int bound_stride, const float* u,
                      const int64_t* offsets, int64_t R, int K, int64_t S, float* z, cudaStream_t stream) {
  if (R == 0) return AVR_OK;
  if (offsets) {
    // AVR_COARSE_PACKED=ray selects the one-warp-per-ray kernel (kept for A/B measurements)
    if (option(OPT_COARSE_PACKED_RAY, 0)) {
      coarse_fwd_packed_kernel<<<grid_for(R, kSamplerWarps, grid_cap(16)), kSamplerWarps * 32, 0, stream>>>(
          near, far, bound_stride, u, offsets, R, z);
    } else {
      const int64_t n_seg = (R + kSegRays - 1) / kSegRays;
      // one segment per warp, no grid cap: the CTAs in flight then cover a compact window of the packed streams
      // (0.247 -> 0.237 ms on 2^20 rays of 8..256 samples)
      coarse_fwd_packed_flat_kernel<<<grid_for(n_seg, kSamplerWarps, 1 << 24), kSamplerWarps * 32, 0, stream>>>(
          near, far, bound_stride, u, offsets, R, z, (aligned16(u) && aligned16(z)) ? 1 : 0);
    }
  } else {
    const int64_t total = R * (int64_t)K;
    if (total == 0) return AVR_OK;
    const bool vec4 = (K % 4 == 0) && aligned16(u) && aligned16(z);
    const int smem = K * (int)sizeof(float);
    if (smem > 48 * 1024) return AVR_ERR_UNSUPPORTED;
    const int blocks = grid_for(total / (vec4 ? 4 : 1), 256, grid_cap(8));
    if (vec4) {
      coarse_fwd_dense_kernel<true><<<blocks, 256, smem, stream>>>(near, far, bound_stride, u, total, K, z);
    } else {
      coarse_fwd_dense_kernel<false><<<blocks, 256, smem, stream>>>(near, far, bound_stride, u, total, K, z);
    }
  }
  (void)S;
  return check_launch();
}

int launch_coarse_bwd(const float* g_z, const float* u, int64_t R, int K, float* d_near, float* d_far,
                      cudaStream_t stream) {
  if (R == 0) return AVR_OK;
  if (K <= 64) {  // short rays: one thread per ray
    if (K % 4 == 0 && aligned16(g_z) && aligned16(u)) {
      coarse_bwd_thread_kernel<4><<<grid_for(R, 256, grid_cap(8)), 256, 0, stream>>>(g_z, u, R, K, d_near, d_far);
    } else {
      coarse_bwd_thread_kernel<1><<<grid_for(R, 256, grid_cap(8)), 256, 0, stream>>>(g_z, u, R, K, d_near, d_far);
    }
    return check_launch();
  }
  coarse_bwd_kernel<<<grid_for(R, kSamplerWarps, grid_cap(16)), kSamplerWarps * 32, 0, stream>>>(g_z, u, R, K,
                                                                                                d_near, d_far);
  return check_launch();
}

int launch_importance(const float* weights, const float* z_coarse, const float* u, const float* u2,
                      const float* normals, const float* near, const float* far, int bound_stride,
                      const int64_t* offsets, const int64_t* fine_offsets, int64_t R, int Kc, int n_imp,
                      int n_depth, float depth_std, float* z_fine, float* z_sorted, float* cdf,
                      int32_t* idx, bool allow_fast, cudaStream_t stream) {
  if (R == 0) return AVR_OK;
  if (allow_fast && n_imp > 0) {
    const int rc = launch_importance_reg(weights, z_coarse, u, u2, normals, near, far, bound_stride, offsets,
                                         fine_offsets, R, Kc, n_imp, n_depth, depth_std, z_fine, z_sorted, cdf, idx,
                                         stream);
    if (rc != AVR_ERR_UNSUPPORTED) return rc;
  }
  // dense: Kc/n_imp are exact; packed: they are the caller's per-ray maxima
  const int total_cap = Kc + n_imp + n_depth;
  if (total_cap > AVR_MAX_SORT) return AVR_ERR_UNSUPPORTED;
  ImportanceArgs a;
  a.weights = weights;
  a.z_coarse = z_coarse;
  a.u = u;
  a.u2 = u2;
  a.normals = normals;
  a.near = near;
  a.far = far;
  a.bound_stride = bound_stride;
  a.offsets = offsets;
  a.fine_offsets = fine_offsets;
  a.R = R;
  a.Kc = Kc;
  a.n_imp = n_imp;
  a.n_depth = n_depth;
  a.depth_std = depth_std;
  a.z_fine = z_fine;
  a.z_sorted = z_sorted;
  a.cdf = cdf;
  a.idx = idx;
  a.cdf_cap = ((Kc + 1) + 3) & ~3;
  a.sort_cap = z_sorted ? next_pow2(total_cap) : 0;
  const int smem = kSamplerWarps * (a.cdf_cap + a.sort_cap) * (int)sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(importance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_last_cuda_error(e);
      return AVR_ERR_LAUNCH;
    }
  }
  importance_kernel<<<grid_for(R, kSamplerWarps, grid_cap(8)), kSamplerWarps * 32, smem, stream>>>(a);
  return check_launch();
}

int launch_sort_rays_bwd(const float* g_out, const int32_t* perm, int64_t R, int K, float* d_in, cudaStream_t stream) {
  const int64_t total = R * (int64_t)K;
  if (total == 0) return AVR_OK;
  sort_rays_bwd_kernel<<<grid_for(total, 256, grid_cap(16)), 256, 0, stream>>>(g_out, perm, total, K, d_in);
  return check_launch();
}

int launch_sort_rays(const float* z_in, int64_t R, int K, float* z_out, int32_t* perm, cudaStream_t stream) {
  if (R == 0 || K == 0) return AVR_OK;
  if (K > AVR_MAX_SORT) return AVR_ERR_UNSUPPORTED;
  const int P = next_pow2(K);
  const int smem = kSamplerWarps * 2 * P * (int)sizeof(float);
  if (K % 4 == 0 && K <= 128 && aligned16(z_in) && aligned16(z_out) && aligned16(perm)) {
    // a short streaming kernel: as many resident warps as the registers allow (48 -> 40 per SM), each with one
    // 512-byte load in flight
    sort_rays_presorted_kernel<<<grid_for(R, kSamplerWarps * 32, 1 << 24), kSamplerWarps * 32, smem, stream>>>(
        z_in, R, K, P, z_out, perm);
    return check_launch();
  }
  sort_rays_kernel<<<grid_for(R, kSamplerWarps, grid_cap(8)), kSamplerWarps * 32, smem, stream>>>(z_in, R, K, P,
                                                                                                z_out, perm);
  return check_launch();
}

}  // namespace avr
