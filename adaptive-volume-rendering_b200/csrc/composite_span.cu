// "Span" compositing kernels for the dense [R,K] layout — the fast path.
//
// Shape of the work: 20 B read + 4..16 B written per sample, ~30 flops and one expf,
// a product scan (transmittance) and, backward, a reverse affine scan along each ray.
// Nothing here is a contraction, so the design is about HBM: every byte moves through
// shared memory by 1-D bulk async copies (TMA, cp.async.bulk + mbarrier), fully
// coalesced and with no registers tied up while in flight, and the SM only does math.
//
//   * A *tile* is a group of whole rays: n_r rays = n_s = n_r*K consecutive samples,
//     contiguous in both the rgbs and z streams.  One WARP owns a tile; warps are
//     autonomous (private stages, private mbarriers, no __syncthreads) and walk the
//     tile list with a grid-wide stride (persistent grid, 148 x occupancy CTAs).
//   * Each warp keeps a ring of NS stages; lane 0 issues the bulk loads NS-2 tiles
//     ahead, all lanes wait on the stage's mbarrier.  Results are written back into
//     the same stage (w over z, d_rgbs over rgbs) and leave by bulk store.
//   * Inside a tile each LANE owns L consecutive samples (a "run"; L is odd so the
//     float4 and scalar shared-memory accesses of the 32 lanes hit distinct banks)
//     and walks them sequentially in registers — ~30 instructions per sample and no
//     shuffles.  Runs are stitched together with ONE segmented warp scan per tile
//     over the per-lane aggregates (transmittance product + partial sums), which is
//     the blocked form of the per-ray exclusive cumprod (renderers.py:90-93).
//
// Backward recomputes the transmittance (walk 1, also yields the reverse-scan
// aggregates) and then walks each run back to front (walk 2) carrying
//     Q_k = sum_{i>k} g_i alpha_i prod_{k<j<i} t_j ,   dL/dalpha_k = T_k (g_k - Q_k),
// see composite_generic.cu for the derivation against torch's cumprod backward.
#include "avr_common.cuh"
#include "kernels.h"

namespace avr {

constexpr int kSpanWarps = 4;  // warps per CTA

template <int L>
struct SpanCfg {
  static constexpr int kStages = (L <= 9) ? 4 : 3;
  static constexpr int kTileSamples = 32 * L;
  static constexpr int kRgbsBytes = kTileSamples * 16;
  static constexpr int kZBytes = kTileSamples * 4;
  static constexpr int kStageBytes = kRgbsBytes + kZBytes;  // multiple of 128
  static constexpr int kWarpBytes = kStages * kStageBytes;
  static constexpr int kSmemBytes = kSpanWarps * kWarpBytes + kSpanWarps * kStages * 8;
};

struct SpanArgs {
  const float* rgbs;
  const float* z;
  float* w;               // fwd (nullable)
  float* rgb;             // fwd
  float* depth;           // fwd
  const float* g_rgb;     // bwd (nullable)
  const float* g_depth;   // bwd (nullable)
  float* d_rgbs;          // bwd
  int64_t n_tiles;
  int K;
  int rays_per_tile;
  int white_back;
  float infinity;
};

// Static description of a lane's run (identical for every tile of a dense launch).
struct Run {
  int s0;         // first sample (tile-relative)
  int nvalid;     // samples in the run (0..L)
  int k0;         // position of the first sample inside its ray
  int ray0;       // tile-relative ray of the first sample
  int carry_len;  // leading samples that belong to a ray started in an earlier lane
};

template <int L>
__device__ __forceinline__ Run make_run(int lane, int K, int n_s) {
  Run r;
  r.s0 = lane * L;
  int rem = n_s - r.s0;
  r.nvalid = rem < 0 ? 0 : (rem > L ? L : rem);
  r.k0 = r.s0 % K;
  r.ray0 = r.s0 / K;
  int to_head = (r.k0 == 0) ? 0 : K - r.k0;
  r.carry_len = to_head < r.nvalid ? to_head : r.nvalid;
  return r;
}

struct Sums {
  float r, g, b, d, a;
};
__device__ __forceinline__ Sums zero_sums() { return Sums{0.f, 0.f, 0.f, 0.f, 0.f}; }

// ---- warp scans over per-lane aggregates ------------------------------------------
// Forward, segmented: element = (flag, T, sums); combine(A earlier, B later) =
// (A.T*B.T, A.s + A.T*B.s) unless B.flag.  Returns the EXCLUSIVE result (carry into the lane).
__device__ __forceinline__ void scan_fwd_exclusive(int lane, bool flag, float& T, Sums& s) {
  unsigned f = flag ? 1u : 0u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float Tp = __shfl_up_sync(0xffffffffu, T, d);
    float pr = __shfl_up_sync(0xffffffffu, s.r, d);
    float pg = __shfl_up_sync(0xffffffffu, s.g, d);
    float pb = __shfl_up_sync(0xffffffffu, s.b, d);
    float pd = __shfl_up_sync(0xffffffffu, s.d, d);
    float pa = __shfl_up_sync(0xffffffffu, s.a, d);
    unsigned fp = __shfl_up_sync(0xffffffffu, f, d);
    if (lane >= d && !f) {
      s.r = pr + Tp * s.r;
      s.g = pg + Tp * s.g;
      s.b = pb + Tp * s.b;
      s.d = pd + Tp * s.d;
      s.a = pa + Tp * s.a;
      T = Tp * T;
      f = fp;
    }
  }
  T = __shfl_up_sync(0xffffffffu, T, 1);
  s.r = __shfl_up_sync(0xffffffffu, s.r, 1);
  s.g = __shfl_up_sync(0xffffffffu, s.g, 1);
  s.b = __shfl_up_sync(0xffffffffu, s.b, 1);
  s.d = __shfl_up_sync(0xffffffffu, s.d, 1);
  s.a = __shfl_up_sync(0xffffffffu, s.a, 1);
  if (lane == 0) {
    T = 1.0f;
    s = zero_sums();
  }
}

__device__ __forceinline__ float scan_fwd_exclusive_T(int lane, bool flag, float T) {
  unsigned f = flag ? 1u : 0u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float Tp = __shfl_up_sync(0xffffffffu, T, d);
    unsigned fp = __shfl_up_sync(0xffffffffu, f, d);
    if (lane >= d && !f) {
      T = Tp * T;
      f = fp;
    }
  }
  T = __shfl_up_sync(0xffffffffu, T, 1);
  return lane == 0 ? 1.0f : T;
}

// Reverse: element = affine map Q_left = A + B*Q_right (B == 0 where a ray ends inside
// the run, which is what stops the carry).  Returns Q entering the lane from the right.
__device__ __forceinline__ float scan_rev_exclusive(int lane, float A, float B) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float Ap = __shfl_down_sync(0xffffffffu, A, d);
    float Bp = __shfl_down_sync(0xffffffffu, B, d);
    if (lane + d < 32) {
      A = A + B * Ap;
      B = B * Bp;
    }
  }
  float q = __shfl_down_sync(0xffffffffu, A, 1);
  return lane == 31 ? 0.f : q;
}

// ---- the per-warp tile pipeline ---------------------------------------------------
template <int L>
struct WarpPipe {
  using Cfg = SpanCfg<L>;
  unsigned char* base;  // this warp's stages
  uint64_t* bars;       // this warp's mbarriers
  int lane;

  __device__ __forceinline__ float4* rgbs_stage(int st) const {
    return reinterpret_cast<float4*>(base + st * Cfg::kStageBytes);
  }
  __device__ __forceinline__ float* z_stage(int st) const {
    return reinterpret_cast<float*>(base + st * Cfg::kStageBytes + Cfg::kRgbsBytes);
  }
  __device__ __forceinline__ void init(unsigned char* smem, int warp, int lane_) {
    lane = lane_;
    base = smem + warp * Cfg::kWarpBytes;
    bars = reinterpret_cast<uint64_t*>(smem + kSpanWarps * Cfg::kWarpBytes) + warp * Cfg::kStages;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < Cfg::kStages; ++s) mbar_init(&bars[s], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  // lane 0 only
  __device__ __forceinline__ void load(int st, const float* rgbs, const float* z, int64_t tile, int n_s) {
    const uint32_t rb = (uint32_t)n_s * 16u, zb = (uint32_t)n_s * 4u;
    mbar_expect_tx(&bars[st], rb + zb);
    bulk_g2s(rgbs_stage(st), rgbs + tile * (int64_t)n_s * 4, rb, &bars[st]);
    bulk_g2s(z_stage(st), z + tile * (int64_t)n_s, zb, &bars[st]);
  }
};

template <int L, bool kWriteW>
__global__ void __launch_bounds__(kSpanWarps * 32)
composite_fwd_span_kernel(const SpanArgs a) {
  using Cfg = SpanCfg<L>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpPipe<L> pipe;
  pipe.init(smem, warp, lane);

  const int K = a.K;
  const int n_s = a.rays_per_tile * K;
  const Run run = make_run<L>(lane, K, n_s);

  const int64_t first = (int64_t)blockIdx.x * kSpanWarps + warp;
  const int64_t stride = (int64_t)gridDim.x * kSpanWarps;
  const int64_t n_my = first < a.n_tiles ? (a.n_tiles - first + stride - 1) / stride : 0;

  if (lane == 0) {
    for (int p = 0; p < NS - 2 && p < n_my; ++p) pipe.load(p, a.rgbs, a.z, first + p * stride, n_s);
  }

  for (int64_t i = 0; i < n_my; ++i) {
    const int st = (int)(i % NS);
    const int64_t tile = first + i * stride;
    {
      const int64_t pf = i + NS - 2;
      if (lane == 0 && pf < n_my) {
        if (kWriteW) bulk_wait_read<1>();  // the store that last read stage pf%NS (tile i-2) is done
        pipe.load((int)(pf % NS), a.rgbs, a.z, first + pf * stride, n_s);
      }
    }
    mbar_wait(&pipe.bars[st], (uint32_t)((i / NS) & 1));

    const float4* rg = pipe.rgbs_stage(st);
    float* zs = pipe.z_stage(st);
    const int64_t ray_base = tile * a.rays_per_tile;

    // ---- walk the run front to back with a local transmittance starting at 1
    float wl[L];
    float Tl = 1.0f;
    Sums s = zero_sums();
    Sums first_seg = zero_sums();  // continuing ray's part, if that ray ends inside this run
    int first_seg_ray = -1;
    bool seen_head = false, closed = false;
    int k = run.k0, ray = run.ray0;
    float zk = run.nvalid > 0 ? zs[run.s0] : 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      if (j < run.nvalid) {
        const int si = run.s0 + j;
        if (k == 0) seen_head = true;
        const bool last = (k == K - 1);
        const float4 c = rg[si];
        const float z_after = (si + 1 < n_s) ? zs[si + 1] : 0.f;
        const float zn = last ? a.infinity : z_after;
        const float delta = last ? kLastDelta : zn - zk;
        const Opacity o = opacity(c.w, delta);
        const float w = o.alpha * Tl;
        wl[j] = w;
        s.r += w * c.x;
        s.g += w * c.y;
        s.b += w * c.z;
        s.d += w * zn;
        s.a += w;
        Tl *= o.t;
        zk = z_after;
        if (last) {
          if (seen_head) {  // ray lies entirely inside this run: finished here
            const float bg = a.white_back ? 1.0f - s.a : 0.f;
            float* o3 = a.rgb + (ray_base + ray) * 3;
            o3[0] = s.r + bg;
            o3[1] = s.g + bg;
            o3[2] = s.b + bg;
            a.depth[ray_base + ray] = s.d;
          } else {
            first_seg = s;
            first_seg_ray = ray;
          }
          closed = true;
          Tl = 1.0f;
          s = zero_sums();
          k = 0;
          ++ray;
        } else {
          ++k;
        }
      } else {
        wl[j] = 0.f;
      }
    }

    // ---- stitch the runs: carry = aggregate of the open ray over the lanes before this one
    float T_in = Tl;
    Sums s_in = s;
    scan_fwd_exclusive(lane, seen_head || closed || run.nvalid == 0, T_in, s_in);

    if (first_seg_ray >= 0) {
      Sums t;
      t.r = s_in.r + T_in * first_seg.r;
      t.g = s_in.g + T_in * first_seg.g;
      t.b = s_in.b + T_in * first_seg.b;
      t.d = s_in.d + T_in * first_seg.d;
      t.a = s_in.a + T_in * first_seg.a;
      const float bg = a.white_back ? 1.0f - t.a : 0.f;
      float* o3 = a.rgb + (ray_base + first_seg_ray) * 3;
      o3[0] = t.r + bg;
      o3[1] = t.g + bg;
      o3[2] = t.b + bg;
      a.depth[ray_base + first_seg_ray] = t.d;
    }

    if (kWriteW) {
      __syncwarp();  // every lane has finished reading z from this stage
#pragma unroll
      for (int j = 0; j < L; ++j) {
        if (j < run.nvalid) zs[run.s0 + j] = (j < run.carry_len) ? wl[j] * T_in : wl[j];
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(a.w + tile * (int64_t)n_s, zs, (uint32_t)n_s * 4u);
        bulk_commit();
      }
    } else {
      __syncwarp();  // stage may be refilled two iterations from now
    }
  }
  if (kWriteW && lane == 0) bulk_wait_all<0>();
}

template <int L>
__global__ void __launch_bounds__(kSpanWarps * 32)
composite_bwd_span_kernel(const SpanArgs a) {
  using Cfg = SpanCfg<L>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpPipe<L> pipe;
  pipe.init(smem, warp, lane);

  const int K = a.K;
  const int n_s = a.rays_per_tile * K;
  const Run run = make_run<L>(lane, K, n_s);
  const int last_idx = run.s0 + run.nvalid - 1;  // tile-relative, valid only if nvalid > 0
  const int k_end = run.nvalid > 0 ? last_idx % K : 0;
  const int ray_end = run.nvalid > 0 ? last_idx / K : 0;

  const int64_t first = (int64_t)blockIdx.x * kSpanWarps + warp;
  const int64_t stride = (int64_t)gridDim.x * kSpanWarps;
  const int64_t n_my = first < a.n_tiles ? (a.n_tiles - first + stride - 1) / stride : 0;

  if (lane == 0) {
    for (int p = 0; p < NS - 2 && p < n_my; ++p) pipe.load(p, a.rgbs, a.z, first + p * stride, n_s);
  }

  for (int64_t i = 0; i < n_my; ++i) {
    const int st = (int)(i % NS);
    const int64_t tile = first + i * stride;
    {
      const int64_t pf = i + NS - 2;
      if (lane == 0 && pf < n_my) {
        bulk_wait_read<1>();
        pipe.load((int)(pf % NS), a.rgbs, a.z, first + pf * stride, n_s);
      }
    }
    const int64_t ray_base = tile * a.rays_per_tile;

    // upstream gradients of the first and last ray this run touches (issued before the
    // wait so their latency hides behind the tile load / walk 1)
    float gA[4] = {0.f, 0.f, 0.f, 0.f}, gB[4] = {0.f, 0.f, 0.f, 0.f};
    if (run.nvalid > 0) {
      const int64_t ra = ray_base + run.ray0, rb = ray_base + ray_end;
      if (a.g_rgb) {
        gA[0] = a.g_rgb[ra * 3 + 0];
        gA[1] = a.g_rgb[ra * 3 + 1];
        gA[2] = a.g_rgb[ra * 3 + 2];
        gB[0] = a.g_rgb[rb * 3 + 0];
        gB[1] = a.g_rgb[rb * 3 + 1];
        gB[2] = a.g_rgb[rb * 3 + 2];
      }
      if (a.g_depth) {
        gA[3] = a.g_depth[ra];
        gB[3] = a.g_depth[rb];
      }
    }

    mbar_wait(&pipe.bars[st], (uint32_t)((i / NS) & 1));
    float4* rg = pipe.rgbs_stage(st);
    const float* zs = pipe.z_stage(st);

    // ---- walk 1, front to back: cache e_j and the local transmittance before sample j;
    // accumulate the first segment's weighted sums (the part of the run that belongs to
    // the ray entering from the left) — they give the reverse-scan aggregate A.
    float ej[L], Tj[L];
    float Tl = 1.0f;
    Sums s = zero_sums();
    bool seen_head = false, closed = false;
    float A = 0.f, B = 1.0f;
    {
      int k = run.k0;
      float zk = run.nvalid > 0 ? zs[run.s0] : 0.f;
#pragma unroll
      for (int j = 0; j < L; ++j) {
        if (j < run.nvalid) {
          const int si = run.s0 + j;
          if (k == 0) seen_head = true;
          const bool last = (k == K - 1);
          const float4 c = rg[si];
          const float z_after = (si + 1 < n_s) ? zs[si + 1] : 0.f;
          const float zn = last ? a.infinity : z_after;
          const float delta = last ? kLastDelta : zn - zk;
          const Opacity o = opacity(c.w, delta);
          ej[j] = o.e;
          Tj[j] = Tl;
          if (!closed) {
            const float w = o.alpha * Tl;
            s.r += w * c.x;
            s.g += w * c.y;
            s.b += w * c.z;
            s.d += w * zn;
            s.a += w;
          }
          Tl *= o.t;
          zk = z_after;
          if (last) {
            if (!closed) {
              B = 0.f;  // a ray ends inside the run: nothing from the right reaches the left
              closed = true;
            }
            Tl = 1.0f;
            k = 0;
          } else {
            ++k;
          }
        } else {
          ej[j] = 1.0f;
          Tj[j] = 1.0f;
        }
      }
    }
    if (run.nvalid == 0) {
      B = 0.f;
    } else {
      // A = sum over the first segment of g_j * (alpha_j * Tlocal_j), g_j linear in the ray's grads
      const float gbg = a.white_back ? (gA[0] + gA[1] + gA[2]) : 0.f;
      A = gA[0] * s.r + gA[1] * s.g + gA[2] * s.b + gA[3] * s.d - gbg * s.a;
      if (!closed) B = Tl;  // whole run is one segment: Tl is its transmittance product
    }
    const float T_in = scan_fwd_exclusive_T(lane, seen_head || closed || run.nvalid == 0, Tl);
    const float Q_in = scan_rev_exclusive(lane, A, B);

    // ---- walk 2, back to front: final gradients, written over the rgbs stage in place
    {
      int k = k_end, ray = ray_end;
      float gr = gB[0], gg = gB[1], gb = gB[2], gd = gB[3];
      float gbg = a.white_back ? (gr + gg + gb) : 0.f;
      float Q = Q_in;
      float zn = 0.f;
      if (run.nvalid > 0 && last_idx + 1 < n_s) zn = zs[last_idx + 1];
#pragma unroll
      for (int j = L - 1; j >= 0; --j) {
        if (j < run.nvalid) {
          const int si = run.s0 + j;
          const bool last = (k == K - 1);
          if (last) {
            Q = 0.f;
            zn = a.infinity;
          }
          const float4 c = rg[si];
          const float zk = zs[si];
          const float delta = last ? kLastDelta : zn - zk;
          const float e = ej[j];
          const float alpha = 1.0f - e;
          const float t = (1.0f - alpha) + kTransEps;
          const float T = (j < run.carry_len) ? Tj[j] * T_in : Tj[j];
          const float g = gr * c.x + gg * c.y + gb * c.z + gd * zn - gbg;
          const float dalpha = T * (g - Q);
          Q = g * alpha + t * Q;
          const float dsd = dalpha * e;
          const float w = alpha * T;
          rg[si] = make_float4(w * gr, w * gg, w * gb, dsd * delta);
          zn = zk;
          if (k == 0) {
            k = K - 1;
            --ray;
            if (j > 0) {  // the next (earlier) sample belongs to the previous ray
              if (ray == run.ray0) {
                gr = gA[0];
                gg = gA[1];
                gb = gA[2];
                gd = gA[3];
              } else {
                const int64_t rr = ray_base + ray;
                gr = a.g_rgb ? a.g_rgb[rr * 3 + 0] : 0.f;
                gg = a.g_rgb ? a.g_rgb[rr * 3 + 1] : 0.f;
                gb = a.g_rgb ? a.g_rgb[rr * 3 + 2] : 0.f;
                gd = a.g_depth ? a.g_depth[rr] : 0.f;
              }
              gbg = a.white_back ? (gr + gg + gb) : 0.f;
            }
          } else {
            --k;
          }
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(a.d_rgbs + tile * (int64_t)n_s * 4, rg, (uint32_t)n_s * 16u);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_all<0>();
}

// ---- host side ----------------------------------------------------------------------
static const int kLs[] = {5, 7, 9, 11, 13};

bool span_plan(int64_t R, int K, const void* rgbs, const void* z, SpanPlan* plan) {
  if (K < 1 || R < 1) return false;
  if (!aligned16(rgbs) || !aligned16(z)) return false;
  int best_L = 0, best_nr = 0;
  double best_util = 0.0;
  for (int L : kLs) {
    int cap = 32 * L;
    int nr = cap / K;
    // tile byte counts must be multiples of 16 for the bulk copies: (nr*K) % 4 == 0
    while (nr > 0 && ((int64_t)nr * K) % 4 != 0) --nr;
    if (nr <= 0) continue;
    double util = (double)nr * K / cap;
    // prefer fuller tiles; on ties prefer the middle of the range (L = 9)
    if (util > best_util + 1e-9) {
      best_util = util;
      best_L = L;
      best_nr = nr;
    }
  }
  if (best_L == 0 || best_util < 0.5) return false;
  int64_t tiles = R / best_nr;
  if (tiles < 1) return false;
  plan->L = best_L;
  plan->rays_per_tile = best_nr;
  plan->main_rays = tiles * best_nr;
  return true;
}

template <typename KernelT>
static int span_grid(KernelT kernel, int smem_bytes, int64_t n_tiles, int* grid) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return AVR_ERR_LAUNCH;
  }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSpanWarps * 32, smem_bytes);
  if (e != cudaSuccess || occ < 1) {
    set_last_cuda_error(e);
    return AVR_ERR_LAUNCH;
  }
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t want = (n_tiles + kSpanWarps - 1) / kSpanWarps;
  int64_t cap = (int64_t)sms * occ;
  *grid = (int)(want < cap ? want : cap);
  return AVR_OK;
}

template <int L>
static int fwd_span_L(const SpanArgs& a, bool write_w, cudaStream_t stream) {
  using Cfg = SpanCfg<L>;
  int grid = 0, rc;
  if (write_w) {
    if ((rc = span_grid(composite_fwd_span_kernel<L, true>, Cfg::kSmemBytes, a.n_tiles, &grid))) return rc;
    composite_fwd_span_kernel<L, true><<<grid, kSpanWarps * 32, Cfg::kSmemBytes, stream>>>(a);
  } else {
    if ((rc = span_grid(composite_fwd_span_kernel<L, false>, Cfg::kSmemBytes, a.n_tiles, &grid))) return rc;
    composite_fwd_span_kernel<L, false><<<grid, kSpanWarps * 32, Cfg::kSmemBytes, stream>>>(a);
  }
  return check_launch();
}

template <int L>
static int bwd_span_L(const SpanArgs& a, cudaStream_t stream) {
  using Cfg = SpanCfg<L>;
  int grid = 0, rc;
  if ((rc = span_grid(composite_bwd_span_kernel<L>, Cfg::kSmemBytes, a.n_tiles, &grid))) return rc;
  composite_bwd_span_kernel<L><<<grid, kSpanWarps * 32, Cfg::kSmemBytes, stream>>>(a);
  return check_launch();
}

#define AVR_DISPATCH_L(L_, CALL)          \
  switch (L_) {                           \
    case 5: return CALL(5);               \
    case 7: return CALL(7);               \
    case 9: return CALL(9);               \
    case 11: return CALL(11);             \
    case 13: return CALL(13);             \
    default: return AVR_ERR_UNSUPPORTED;  \
  }

int launch_composite_fwd_span(const SpanPlan& plan, const float* rgbs, const float* z, int K,
                              int white_back, float infinity, float* w, float* rgb, float* depth,
                              cudaStream_t stream) {
  SpanArgs a{};
  a.rgbs = rgbs;
  a.z = z;
  a.w = w;
  a.rgb = rgb;
  a.depth = depth;
  a.n_tiles = plan.main_rays / plan.rays_per_tile;
  a.K = K;
  a.rays_per_tile = plan.rays_per_tile;
  a.white_back = white_back;
  a.infinity = infinity;
  const bool write_w = (w != nullptr);
#define CALL(LL) fwd_span_L<LL>(a, write_w, stream)
  AVR_DISPATCH_L(plan.L, CALL)
#undef CALL
}

int launch_composite_bwd_span(const SpanPlan& plan, const float* rgbs, const float* z, const float* g_rgb,
                              const float* g_depth, int K, int white_back, float infinity, float* d_rgbs,
                              cudaStream_t stream) {
  SpanArgs a{};
  a.rgbs = rgbs;
  a.z = z;
  a.g_rgb = g_rgb;
  a.g_depth = g_depth;
  a.d_rgbs = d_rgbs;
  a.n_tiles = plan.main_rays / plan.rays_per_tile;
  a.K = K;
  a.rays_per_tile = plan.rays_per_tile;
  a.white_back = white_back;
  a.infinity = infinity;
#define CALL(LL) bwd_span_L<LL>(a, stream)
  AVR_DISPATCH_L(plan.L, CALL)
#undef CALL
}

}  // namespace avr
