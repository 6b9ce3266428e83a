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
//     and walks them sequentially in registers — no shuffles per sample.  Runs are
//     stitched together with ONE segmented warp scan per tile over the per-lane
//     aggregates (transmittance product + partial sums): the blocked form of the
//     per-ray exclusive cumprod (renderers.py:90-93).
//   * With K > L (every shipped configuration) a run holds at most one ray boundary,
//     at a position that is fixed per lane for the whole launch.  The "simple" tile
//     bodies below exploit that: straight-line code, two accumulator sets selected by
//     a per-lane predicate, no per-sample branches.  K <= L takes the general bodies.
//
// Backward recomputes the transmittance (walk 1, also yields the reverse-scan
// aggregates) and then walks each run back to front (walk 2) carrying
//     Q_k = sum_{i>k} g_i alpha_i prod_{k<j<i} t_j ,   dL/dalpha_k = T_k (g_k - Q_k),
// see composite_generic.cu for the derivation against torch's cumprod backward.
#include <cstdlib>

#include "avr_common.cuh"
#include "kernels.h"
#include "span_bodies.cuh"

namespace avr {

// ---- forward, any K (several boundaries per run possible) ----------------------------
template <int L, bool kWriteW>
__device__ __forceinline__ void fwd_tile_general(const SpanArgs& a, const Run& run, const float4* rg, float* zs,
                                                 int64_t ray_base, int lane, float4* gather_ring = nullptr) {
  const int K = a.K;
  float wl[L];
  float Tl = 1.0f;
  Sums s = zero_sums();
  Sums first_seg = zero_sums();  // continuing ray's part, if that ray ends inside this run
  int first_seg_ray = -1;
  bool seen_head = false, closed = false;
  int k = run.k0, ray = run.ray0;
  float zk = zs[0];
#pragma unroll
  for (int j = 0; j < L; ++j) {
    if (j < run.nvalid) {
      if (k == 0) seen_head = true;
      const bool last = (k == K - 1);
      const float4 c = rg[j];
      const float z_after = zs[j + 1];
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
          store_ray(a, ray_base + ray, s, gather_ring);
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
    store_ray(a, ray_base + first_seg_ray, t, gather_ring);
  }
  if (kWriteW) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < L; ++j) {
      if (j < run.nvalid) zs[j] = (j < run.carry_len) ? wl[j] * T_in : wl[j];
    }
  }
}

// ---- backward, any K ----------------------------------------------------------------
template <int L>
__device__ __forceinline__ void bwd_tile_general(const SpanArgs& a, const Run& run, float4* rg, const float* zs,
                                                 const RayGrad& gA, const RayGrad& gB, int64_t ray_base,
                                                 int lane) {
  const int K = a.K;
  const int last_idx = run.nvalid - 1;  // run-relative
  const int k_end = run.nvalid > 0 ? (run.s0 + last_idx) % K : 0;
  const int ray_end = run.nvalid > 0 ? (run.s0 + last_idx) / K : 0;
  float ej[L], Tj[L];
  float Tl = 1.0f;
  Sums s = zero_sums();
  bool seen_head = false, closed = false;
  float A = 0.f, Bm = 1.0f;
  {
    int k = run.k0;
    float zk = zs[0];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      if (j < run.nvalid) {
        if (k == 0) seen_head = true;
        const bool last = (k == K - 1);
        const float4 c = rg[j];
        const float z_after = zs[j + 1];
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
            Bm = 0.f;  // a ray ends inside the run: nothing from the right reaches the left
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
    Bm = 0.f;
  } else {
    A = gA.r * s.r + gA.g * s.g + gA.b * s.b + gA.d * s.d - gA.bg * s.a;
    if (!closed) Bm = Tl;  // whole run is one segment: Tl is its transmittance product
  }
  const float T_in = scan_fwd_exclusive_T(lane, seen_head || closed || run.nvalid == 0, Tl);
  const float Q_in = scan_rev_exclusive(lane, A, Bm);
  {
    int k = k_end, ray = ray_end;
    RayGrad g = gB;
    float Q = Q_in;
    float zn = run.nvalid > 0 ? zs[last_idx + 1] : 0.f;
#pragma unroll
    for (int j = L - 1; j >= 0; --j) {
      if (j < run.nvalid) {
        const bool last = (k == K - 1);
        if (last) {
          Q = 0.f;
          zn = a.infinity;
        }
        const float4 c = rg[j];
        const float zk = zs[j];
        const float delta = last ? kLastDelta : zn - zk;
        const float e = ej[j];
        const float alpha = 1.0f - e;
        const float t = (1.0f - alpha) + kTransEps;
        const float T = (j < run.carry_len) ? Tj[j] * T_in : Tj[j];
        const float gs = g.r * c.x + g.g * c.y + g.b * c.z + g.d * zn - g.bg;
        const float dalpha = T * (gs - Q);
        Q = gs * alpha + t * Q;
        const float dsd = dalpha * e;
        const float w = alpha * T;
        rg[j] = make_float4(w * g.r, w * g.g, w * g.b, dsd * delta);
        zn = zk;
        if (k == 0) {
          k = K - 1;
          --ray;
          if (j > 0) g = (ray == run.ray0) ? gA : load_ray_grad(a, ray_base + ray);
        } else {
          --k;
        }
      }
    }
  }
}

// =====================================================================================
// Kernels: per-warp tile pipeline around the bodies.
// =====================================================================================
// byte offset of the gather rings in a forward CTA's dynamic shared memory (16-byte aligned)
__host__ __device__ inline int gather_ring_offset(int warps, int stages, int stage_bytes) {
  return (warps * stages * stage_bytes + warps * stages * 8 + 15) & ~15;
}

template <int L, int NS>
struct WarpPipe {
  using Cfg = SpanCfg<L>;
  unsigned char* base;  // this warp's stages
  uint64_t* bars;       // this warp's mbarriers

  __device__ __forceinline__ float4* rgbs_stage(int st) const {
    return reinterpret_cast<float4*>(base + st * Cfg::kStageBytes);
  }
  __device__ __forceinline__ float* z_stage(int st) const {
    return reinterpret_cast<float*>(base + st * Cfg::kStageBytes + Cfg::kRgbsBytes);
  }
  __device__ __forceinline__ void init(unsigned char* smem, int warps, int warp, int lane) {
    base = smem + warp * (NS * Cfg::kStageBytes);
    bars = reinterpret_cast<uint64_t*>(smem + warps * (NS * Cfg::kStageBytes)) + warp * NS;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
  // lane 0 only
  __device__ __forceinline__ void load(int st, const SpanArgs& a, int64_t tile, int n_s_full, int n_s) {
    const uint32_t rb = (uint32_t)n_s * 16u, zb = (uint32_t)n_s * 4u;
    mbar_expect_tx(&bars[st], rb + zb);
    bulk_g2s(rgbs_stage(st), a.rgbs + tile * (int64_t)n_s_full * 4, rb, &bars[st]);
    bulk_g2s(z_stage(st), a.z + tile * (int64_t)n_s_full, zb, &bars[st]);
  }
};

template <int L, int NS, bool kSimple, bool kWriteW>
__global__ void __launch_bounds__(256)
composite_fwd_span_kernel(const SpanArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpPipe<L, NS> pipe;
  pipe.init(smem, warps, warp, lane);

  const int K = a.K;
  const int n_s = a.rays_per_tile * K;
  const int n_s_tail = a.tail_rays * K;
  const Run run_full = make_run<L>(lane, K, n_s);

  // tiles are dealt round-robin over all warps (neighbouring warps stream neighbouring
  // addresses); with the fused all-gather each warp takes a contiguous block of tiles instead,
  // so that the rays it finishes are consecutive and leave for the peers in 512-byte stores
  const bool gather = a.n_peers > 0;
  // gather mode: this warp's ring of finished rays (r,g,b,depth), behind the stages and their barriers
  float4* gring = gather ? reinterpret_cast<float4*>(smem + gather_ring_offset(warps, NS, WarpPipe<L, NS>::Cfg::kStageBytes)) +
                               warp * kGatherRing
                         : nullptr;
  const int64_t gw = (int64_t)blockIdx.x * warps + warp, n_warps = (int64_t)gridDim.x * warps;
  const int64_t per_warp = (a.n_tiles + n_warps - 1) / n_warps;
  const int64_t first = gather ? gw * per_warp : gw;
  const int64_t stride = gather ? 1 : n_warps;
  const int64_t n_my = gather ? (first < a.n_tiles ? (a.n_tiles - first < per_warp ? a.n_tiles - first : per_warp) : 0)
                              : (first < a.n_tiles ? (a.n_tiles - first + stride - 1) / stride : 0);
  int64_t flushed = first * a.rays_per_tile;  // gather: rays below this are already at the peers
  auto samples_of = [&](int64_t tile) { return tile == a.n_tiles - 1 ? n_s_tail : n_s; };

  if (lane == 0) {
    for (int p = 0; p < NS - 2 && p < n_my; ++p) {
      const int64_t t = first + p * stride;
      pipe.load(p, a, t, n_s, samples_of(t));
    }
  }
  for (int64_t i = 0; i < n_my; ++i) {
    const int st = (int)(i % NS);
    const int64_t tile = first + i * stride;
    {
      const int64_t pf = i + NS - 2;
      if (lane == 0 && pf < n_my) {
        if (kWriteW) bulk_wait_read<1>();  // the store that last read stage pf%NS (tile i-2) is done
        const int64_t t = first + pf * stride;
        pipe.load((int)(pf % NS), a, t, n_s, samples_of(t));
      }
    }
    const int n_cur = samples_of(tile);
    Run run = run_full;
    if (n_cur != n_s) run = make_run<L>(lane, K, n_cur);  // partial last tile (warp-uniform branch)
    mbar_wait(&pipe.bars[st], (uint32_t)((i / NS) & 1));

    const float4* rg = pipe.rgbs_stage(st) + run.s0;
    float* zs = pipe.z_stage(st) + run.s0;
    const int64_t ray_base = tile * a.rays_per_tile;
    if (kSimple) {
      fwd_tile_simple<L, kWriteW>(a, run, rg, zs, ray_base, lane, gring);
    } else {
      fwd_tile_general<L, kWriteW>(a, run, rg, zs, ray_base, lane, gring);
    }
    if (kWriteW) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(a.w + tile * (int64_t)n_s, pipe.z_stage(st), (uint32_t)n_cur * 4u);
        bulk_commit();
      }
    } else {
      __syncwarp();  // stage may be refilled two iterations from now
    }
    if (gather) {
      const int64_t done = ray_base + n_cur / K;
      const int64_t upto = (i + 1 == n_my) ? done : flushed + ((done - flushed) & ~(int64_t)31);  // whole 512-byte rows
      if (upto > flushed) {
        flush_rays_to_peers(a, flushed, upto, lane, gring);
        flushed = upto;
      }
    }
  }
  if (kWriteW && lane == 0) bulk_wait_all<0>();
  if (gather && a.n_signal > 0) signal_gather_done(a);
}

template <int L, int NS, bool kSimple, bool kDz, bool kCam = true>
__global__ void __launch_bounds__(256)
composite_bwd_span_kernel(const SpanArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpPipe<L, NS> pipe;
  pipe.init(smem, warps, warp, lane);

  const int K = a.K;
  const int n_s = a.rays_per_tile * K;
  const int n_s_tail = a.tail_rays * K;
  const Run run_full = make_run<L>(lane, K, n_s);

  const int64_t first = (int64_t)blockIdx.x * warps + warp;
  const int64_t stride = (int64_t)gridDim.x * warps;
  const int64_t n_my = first < a.n_tiles ? (a.n_tiles - first + stride - 1) / stride : 0;
  auto samples_of = [&](int64_t tile) { return tile == a.n_tiles - 1 ? n_s_tail : n_s; };

  if (lane == 0) {
    for (int p = 0; p < NS - 2 && p < n_my; ++p) {
      const int64_t t = first + p * stride;
      pipe.load(p, a, t, n_s, samples_of(t));
    }
  }
  for (int64_t i = 0; i < n_my; ++i) {
    const int st = (int)(i % NS);
    const int64_t tile = first + i * stride;
    {
      const int64_t pf = i + NS - 2;
      if (lane == 0 && pf < n_my) {
        bulk_wait_read<1>();
        const int64_t t = first + pf * stride;
        pipe.load((int)(pf % NS), a, t, n_s, samples_of(t));
      }
    }
    const int n_cur = samples_of(tile);
    Run run = run_full;
    if (n_cur != n_s) run = make_run<L>(lane, K, n_cur);
    const int64_t ray_base = tile * a.rays_per_tile;

    // upstream gradients of the first and last ray this run touches (issued before the
    // wait so their latency hides behind the tile load)
    RayGrad gA{0.f, 0.f, 0.f, 0.f, 0.f}, gB{0.f, 0.f, 0.f, 0.f, 0.f};
    if (run.nvalid > 0) {
      gA = load_ray_grad<kCam>(a, ray_base + run.ray0);
      const int ray_end = (run.s0 + run.nvalid - 1) / K;
      gB = (ray_end != run.ray0) ? load_ray_grad<kCam>(a, ray_base + ray_end) : gA;
    }
    mbar_wait(&pipe.bars[st], (uint32_t)((i / NS) & 1));
    float4* rg = pipe.rgbs_stage(st) + run.s0;
    float* zs = pipe.z_stage(st) + run.s0;
    if (kSimple) {
      bwd_tile_simple<L, kDz>(a, run, rg, zs, gA, gB, lane);
    } else {
      bwd_tile_general<L>(a, run, rg, zs, gA, gB, ray_base, lane);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(a.d_rgbs + tile * (int64_t)n_s * 4, pipe.rgbs_stage(st), (uint32_t)n_cur * 16u);
      if (kDz) bulk_s2g(a.d_z + tile * (int64_t)n_s, pipe.z_stage(st), (uint32_t)n_cur * 4u);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_all<0>();
}

// ---- host side ----------------------------------------------------------------------
static const int kLs[] = {5, 7, 9, 11, 13};


// Defaults from the B200 sweep in profiles/r01_span_sweep.md: a 3-deep ring (one tile in
// flight behind the one being computed) with as many resident warps as shared memory
// allows (12/SM at L <= 9, 8/SM above) beat deeper rings with fewer warps.
// Tuning knobs (experiments only): AVR_SPAN_L forces the samples-per-lane, AVR_SPAN_STAGES
// (3|4) the ring depth, AVR_SPAN_WARPS (1..8) the warps per CTA.
static int stages_for(int L) {
  (void)L;
  int s = option(OPT_SPAN_STAGES, 0);
  return (s == 3 || s == 4) ? s : 3;
}
static int warps_per_cta(int L) {
  int w = option(OPT_SPAN_WARPS, L >= 11 ? 2 : 4);
  return w < 1 ? 1 : (w > 8 ? 8 : w);
}

bool span_plan(int64_t R, int K, const void* rgbs, const void* z, SpanPlan* plan) {
  if (K < 1 || R < 1) return false;
  if (!aligned16(rgbs) || !aligned16(z)) return false;
  const int forced = option(OPT_SPAN_L, 0);
  int best_L = 0, best_nr = 0;
  double best_util = 0.0;
  for (int L : kLs) {
    if (forced && L != forced) continue;
    if (K <= L && L != 5) continue;  // the general (multi-boundary) bodies are built for L = 5 only
    int cap = 32 * L;
    int nr = cap / K;
    // tile byte counts must be multiples of 16 for the bulk copies: (nr*K) % 4 == 0
    while (nr > 0 && ((int64_t)nr * K) % 4 != 0) --nr;
    if (nr <= 0) continue;
    double util = (double)nr * K / cap;
    if (util > best_util + 1e-9) {
      best_util = util;
      best_L = L;
      best_nr = nr;
    }
  }
  // L = 5 fills best for short rays (K = 20, 32, 40 ...) but its 160-sample tiles are mostly per-tile
  // overhead: a larger run that fills within 5 % is faster (K = 20: fwd 0.083 -> 0.069 ms, 81 -> 98 %
  // of the roofline with L = 9)
  if (best_L == 5 && !forced && K > 9) {
    for (int L : {9, 7}) {
      int nr = (32 * L) / K;
      while (nr > 0 && ((int64_t)nr * K) % 4 != 0) --nr;
      if (nr > 0 && (double)nr * K / (32 * L) >= best_util - 0.05) {
        best_util = (double)nr * K / (32 * L);
        best_L = L;
        best_nr = nr;
        break;
      }
    }
  }
  if (best_L == 0 || best_util < 0.5) return false;
  int64_t tiles = R / best_nr;
  if (tiles < 1) return false;
  plan->L = best_L;
  plan->rays_per_tile = best_nr;
  plan->main_rays = tiles * best_nr;
  // a partial last tile joins the span launch when its byte counts stay 16-byte multiples
  const int64_t rem = R - plan->main_rays;
  if (rem > 0 && (rem * K) % 4 == 0) plan->main_rays = R;
  return true;
}

template <typename KernelT>
static int span_launch(KernelT kernel, int L, int stage_bytes, int stages, const SpanArgs& a, cudaStream_t stream) {
  const int warps = warps_per_cta(L);
  const int smem_bytes = a.n_peers > 0 ? gather_ring_offset(warps, stages, stage_bytes) + warps * kGatherRing * 16
                                       : warps * stages * stage_bytes + warps * stages * 8;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    (void)cudaGetLastError();
    return AVR_ERR_LAUNCH;
  }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, warps * 32, smem_bytes);
  if (e != cudaSuccess || occ < 1) {
    set_last_cuda_error(e);
    (void)cudaGetLastError();
    return AVR_ERR_LAUNCH;
  }
  const int sms = num_sms();
  const int64_t want = (a.n_tiles + warps - 1) / warps;
  const int64_t cap = (int64_t)sms * occ;
  const int grid = (int)(want < cap ? want : cap);
  kernel<<<grid, warps * 32, smem_bytes, stream>>>(a);
  return check_launch();
}

template <int L, int NS>
static int fwd_span_LN(const SpanArgs& a, bool simple, bool write_w, cudaStream_t stream) {
  constexpr int sb = SpanCfg<L>::kStageBytes;
  if (simple) {
    return write_w ? span_launch(composite_fwd_span_kernel<L, NS, true, true>, L, sb, NS, a, stream)
                   : span_launch(composite_fwd_span_kernel<L, NS, true, false>, L, sb, NS, a, stream);
  }
  if constexpr (L == 5) {
    return write_w ? span_launch(composite_fwd_span_kernel<L, NS, false, true>, L, sb, NS, a, stream)
                   : span_launch(composite_fwd_span_kernel<L, NS, false, false>, L, sb, NS, a, stream);
  }
  return AVR_ERR_UNSUPPORTED;
}

template <int L, int NS>
static int bwd_span_LN(const SpanArgs& a, bool simple, cudaStream_t stream) {
  constexpr int sb = SpanCfg<L>::kStageBytes;
  if (simple) {
    if (!a.d_z && !a.depth_affine) return span_launch(composite_bwd_span_kernel<L, NS, true, false, false>, L, sb, NS, a, stream);
    return a.d_z ? span_launch(composite_bwd_span_kernel<L, NS, true, true>, L, sb, NS, a, stream)
                 : span_launch(composite_bwd_span_kernel<L, NS, true, false>, L, sb, NS, a, stream);
  }
  if (a.d_z) return AVR_ERR_UNSUPPORTED;  // callers check span_supports_dz() first
  if constexpr (L == 5) return span_launch(composite_bwd_span_kernel<L, NS, false, false>, L, sb, NS, a, stream);
  return AVR_ERR_UNSUPPORTED;
}

#define AVR_DISPATCH_LN(L_, NS_, CALL)                             \
  switch (L_) {                                                    \
    case 5: return NS_ == 3 ? CALL(5, 3) : CALL(5, 4);             \
    case 7: return NS_ == 3 ? CALL(7, 3) : CALL(7, 4);             \
    case 9: return NS_ == 3 ? CALL(9, 3) : CALL(9, 4);             \
    case 11: return NS_ == 3 ? CALL(11, 3) : CALL(11, 4);          \
    case 13: return NS_ == 3 ? CALL(13, 3) : CALL(13, 4);          \
    default: return AVR_ERR_UNSUPPORTED;                           \
  }

static SpanArgs make_args(const SpanPlan& plan, int K, int white_back, float infinity) {
  SpanArgs a{};
  a.n_tiles = (plan.main_rays + plan.rays_per_tile - 1) / plan.rays_per_tile;
  a.K = K;
  a.rays_per_tile = plan.rays_per_tile;
  const int rem = (int)(plan.main_rays % plan.rays_per_tile);
  a.tail_rays = rem ? rem : plan.rays_per_tile;
  a.white_back = white_back;
  a.infinity = infinity;
  return a;
}

int launch_composite_fwd_span(const SpanPlan& plan, const float* rgbs, const float* z, int K,
                              int white_back, float infinity, float* w, float* rgb, float* depth,
                              cudaStream_t stream, void* const* peers, int n_peers, int64_t peer_row0,
                              bool multicast, const GatherSignal* signal, const float* depth_affine) {
  SpanArgs a = make_args(plan, K, white_back, infinity);
  a.depth_affine = depth_affine;
  if (n_peers > kMaxPeers) return AVR_ERR_UNSUPPORTED;
  if (signal && signal->n > 0) {
    if (signal->n > kMaxPeers || n_peers < 1) return AVR_ERR_UNSUPPORTED;
    a.n_signal = signal->n;
    a.signal_slot = signal->slot;
    a.signal_value = signal->value;
    a.done_counter = signal->done_counter;
    for (int p = 0; p < signal->n; ++p) a.signal_flags[p] = signal->flags[p];
  }
  a.n_peers = n_peers;
  a.peers_multicast = multicast ? 1 : 0;
  a.peer_row0 = peer_row0;
  for (int p = 0; p < n_peers; ++p) a.peers[p] = reinterpret_cast<float4*>(peers[p]);
  a.rgbs = rgbs;
  a.z = z;
  a.w = w;
  a.rgb = rgb;
  a.depth = depth;
  const bool write_w = (w != nullptr);
  const bool simple = K > plan.L;
  const int ns = stages_for(plan.L);
#define CALL(LL, NN) fwd_span_LN<LL, NN>(a, simple, write_w, stream)
  AVR_DISPATCH_LN(plan.L, ns, CALL)
#undef CALL
}

int launch_composite_bwd_span(const SpanPlan& plan, const float* rgbs, const float* z, const float* g_rgb,
                              const float* g_depth, int K, int white_back, float infinity, float* d_rgbs,
                              float* d_z, cudaStream_t stream, const float* depth_affine) {
  SpanArgs a = make_args(plan, K, white_back, infinity);
  a.depth_affine = depth_affine;
  a.d_z = d_z;
  a.rgbs = rgbs;
  a.z = z;
  a.g_rgb = g_rgb;
  a.g_depth = g_depth;
  a.d_rgbs = d_rgbs;
  const bool simple = K > plan.L;
  const int ns = stages_for(plan.L);
#define CALL(LL, NN) bwd_span_LN<LL, NN>(a, simple, stream)
  AVR_DISPATCH_LN(plan.L, ns, CALL)
#undef CALL
}

}  // namespace avr
