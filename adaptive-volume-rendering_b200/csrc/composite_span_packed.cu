// Span compositing kernels for the PACKED (ragged) layout: per-ray sample counts given by
// offsets[R+1] (BASELINE.json config 4, "adaptive" rays of 8..256 samples).
//
// Same machinery as the dense span kernels (composite_span.cu): whole-ray tiles staged in
// shared memory by 1-D bulk async copies, per-warp private ring + mbarriers, each lane walks
// L consecutive samples, one segmented warp scan per tile — but the tiling is dynamic:
//
//   * each warp owns a contiguous range of rays holding ~1/n-th of the SAMPLES (ranges are
//     cut by a warp-wide 32-ary search of `offsets`, so short and long rays balance out);
//   * walking its range, the warp packs consecutive rays greedily into tiles of at most
//     C = 32*L samples (at most 32 rays): the next 32 ray ends are read with one coalesced
//     load one iteration ahead, a ballot finds how many fit;
//   * a tile only takes a ray whose last sample falls into another lane's run than the last
//     sample of the ray before it (any ray longer than L; in the backward kernel also a shorter
//     one unless it is unlucky, and then it opens the next tile), so a lane's run holds at most one ray boundary and the
//     branch-free "simple" tile bodies apply unchanged; the per-lane run description (ray,
//     boundary position) is found per tile by a 6-step search of the tile's ray-end table in
//     shared memory;
//   * rays a tile cannot take — empty ones, and in the backward kernel count > C — are composited
//     by the same warp with the warp-per-ray routine straight from global memory;
//   * the forward kernels do not need whole rays in a tile: see "streaming forward" below;
//   * z and w slices start at arbitrary sample offsets: the stage keeps the global 16-byte
//     phase (bulk copies cover the aligned interior, up to 3 elements at either end go
//     through ordinary loads/stores);
//   * the ring has two slots and results (w, d_rgbs) leave by ordinary coalesced 16-byte stores
//     inside the iteration instead of bulk stores: these kernels are bound by per-warp
//     instruction latency, so 12 resident warps/SM (2 x 8.6 KB each) beat 8 with a third slot.
#include <climits>
#include <cstdlib>
#include <type_traits>

#include "avr_common.cuh"
#include "kernels.h"
#include "span_bodies.cuh"
#include "wray_device.cuh"

namespace avr {

// tuning knobs (compile-time; -DAVR_PK_L=.. -DAVR_PK_WARPS=.. -DAVR_PK_STAGES=.. for experiments)
#ifndef AVR_PK_L
#define AVR_PK_L 13
#endif
#ifndef AVR_PK_WARPS
#define AVR_PK_WARPS 1
#endif
#ifndef AVR_PK_STAGES
#define AVR_PK_STAGES 2
#endif
#ifndef AVR_PK_L2
#define AVR_PK_L2 9  // run length of the forward kernels' second body (small tiles)
#endif
constexpr int kPkL = AVR_PK_L;   // samples per lane (odd: conflict-free shared-memory runs)
constexpr int kPkC = 32 * kPkL;  // samples per tile
constexpr int kPkStages = AVR_PK_STAGES;
// Results leave the stage by bulk store when the ring has a third slot to cover the store's read of
// shared memory; with a 2-slot ring they leave by ordinary coalesced stores inside the iteration, which
// frees the slot at once.  The packed kernels are bound by per-warp instruction latency (a tile's walk is
// one dependent chain; 0.25 IPC per warp), so what pays is resident warps: 2 x 8.6 KB per warp = 13
// warps/SM with one-warp CTAs instead of 8 (profiles/r01_span_sweep.md).
constexpr bool kOutByTma = kPkStages >= 3;
constexpr int kPkWarps = AVR_PK_WARPS;
constexpr int kPkRgbsBytes = kPkC * 16;
constexpr int kPkZFloats = kPkC + 8;  // + phase shift (<= 3) + one z past the tile + slack
constexpr int kPkStageBytes = kPkRgbsBytes + kPkZFloats * 4 + 32 * 4;  // rgbs | z | ray ends
constexpr int kPkSmemBytes = kPkWarps * kPkStages * kPkStageBytes + kPkWarps * kPkStages * 8;

struct PackedArgs {
  SpanArgs sp;  // rgbs, z, outputs, gradients, white_back, infinity (tile fields unused)
  const int64_t* offsets;
  int64_t R, S;
};

enum { kItemNone = 0, kItemTile = 1, kItemRay = 2 };
struct Item {
  int64_t r0;  // first ray
  int64_t sb;  // first sample
  int64_t n_s; // samples
  int nr;      // rays
  int kind;
};

// smallest r in [0, R] with offsets[r] >= target (offsets[R] = S >= target); all lanes cooperate
__device__ __forceinline__ int64_t warp_lower_bound(const int64_t* __restrict__ offsets, int64_t R,
                                                    int64_t target, int lane) {
  if (target <= 0) return 0;
  int64_t lo = 0, hi = R;  // offsets[lo] < target <= offsets[hi]
  while (hi - lo > 1) {
    const int64_t chunk = (hi - lo + 31) >> 5;
    int64_t q = lo + (int64_t)(lane + 1) * chunk;
    if (q > hi) q = hi;
    const bool ge = offsets[q] >= target;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    const int f = __ffs(m) - 1;  // m != 0: the last probe is hi
    const int64_t q_f = __shfl_sync(0xffffffffu, q, f);
    const int64_t q_prev = __shfl_sync(0xffffffffu, q, f > 0 ? f - 1 : 0);
    hi = q_f;
    if (f > 0) lo = q_prev;
  }
  return hi;
}

struct Window {
  int64_t ostart;  // offsets[r]
  int64_t oend;    // offsets[r + 1 + lane] (INT64_MAX past the warp's range)
};
__device__ __forceinline__ Window load_window(const int64_t* __restrict__ offsets, int64_t r, int64_t rb, int lane) {
  Window w;
  w.ostart = (r < rb) ? offsets[r] : 0;
  w.oend = (r + lane < rb) ? offsets[r + 1 + lane] : INT64_MAX;
  return w;
}

// The tile bodies allow ONE ray end per lane run.  Rays longer than a run guarantee that; a shorter ray still fits
// when its last sample and the last sample of the ray before it fall into different runs (always true for the
// first ray of a tile).  `rel`: the lane's ray end relative to the tile start, >= 1.  Checked for both run lengths
// the kernels use.  (Short rays used to leave the tiles altogether: each cut a tile short and went through the
// warp-per-ray routine — 4.5 % of the kernels' time on counts 8..256, where 2.4 % of the rays have <= 13 samples.)
__device__ __forceinline__ bool ends_share_a_run(int rel, int lane) {
  const int run_a = (rel - 1) / kPkL, run_b = (rel - 1) / AVR_PK_L2;
  int prev_a = __shfl_up_sync(0xffffffffu, run_a, 1), prev_b = __shfl_up_sync(0xffffffffu, run_b, 1);
  if (lane == 0) prev_a = prev_b = -1;
  return run_a == prev_a || run_b == prev_b;
}

// Decide what the next work item starting at ray r is.  `rel_end` returns the lane's ray end
// relative to the item start (INT_MAX for lanes past it); meaningful for tiles.
__device__ __forceinline__ Item next_item(const Window& w, int64_t r, int64_t rb, int lane, int& rel_end) {
  Item it;
  it.r0 = r;
  it.sb = w.ostart;
  it.kind = kItemNone;
  it.nr = 0;
  it.n_s = 0;
  rel_end = INT_MAX;
  if (r >= rb) return it;
  const bool valid = (r + lane < rb);
  const int64_t rel64 = valid ? (w.oend - w.ostart) : (int64_t)INT_MAX;
  const int rel = rel64 > (int64_t)INT_MAX ? INT_MAX : (int)rel64;
  int prev = __shfl_up_sync(0xffffffffu, rel, 1);
  if (lane == 0) prev = 0;
  const int cnt = rel - prev;
  const bool shares = ends_share_a_run(rel, lane);  // (shuffles: every lane, before any short-circuit)
  const bool ok = valid && rel <= kPkC && cnt >= 1 && !shares;
  const unsigned mask = __ballot_sync(0xffffffffu, ok);
  const int nr = (mask == 0xffffffffu) ? 32 : (__ffs(~mask) - 1);  // leading rays that fit
  if (nr == 0) {  // first ray is too short or too long for a tile: composite it on its own
    it.kind = kItemRay;
    it.nr = 1;
    it.n_s = __shfl_sync(0xffffffffu, rel64, 0);
    return it;
  }
  it.kind = kItemTile;
  it.nr = nr;
  it.n_s = __shfl_sync(0xffffffffu, rel, nr - 1);
  rel_end = (lane < nr) ? rel : INT_MAX;
  return it;
}

struct PkPipe {
  unsigned char* base;
  uint64_t* bars;
  __device__ __forceinline__ float4* rgbs_stage(int st) const {
    return reinterpret_cast<float4*>(base + st * kPkStageBytes);
  }
  __device__ __forceinline__ float* z_stage(int st) const {
    return reinterpret_cast<float*>(base + st * kPkStageBytes + kPkRgbsBytes);
  }
  __device__ __forceinline__ int* ends_stage(int st) const {
    return reinterpret_cast<int*>(base + st * kPkStageBytes + kPkRgbsBytes + kPkZFloats * 4);
  }
  __device__ __forceinline__ void init(unsigned char* smem, int warp, int lane) {
    base = smem + warp * (kPkStages * kPkStageBytes);
    bars = reinterpret_cast<uint64_t*>(smem + kPkWarps * (kPkStages * kPkStageBytes)) + warp * kPkStages;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < kPkStages; ++s) mbar_init(&bars[s], 1);
      fence_mbar_init();
    }
    __syncwarp();
  }
};

// z elements [zb, zb + bulk) go by bulk copy (zb = sb rounded down to 4; everything inside the
// buffer's last full 16 bytes); the rest of the tile's z (<= 3 elements at the very end of the
// buffer) is patched by plain loads after the wait
__device__ __forceinline__ int z_bulk_elems(int64_t sb, int n_s, int64_t S) {
  const int shift = (int)(sb & 3);
  const int64_t zb = sb - shift;
  int64_t end = zb + ((shift + n_s + 3) & ~3);
  const int64_t cap = S & ~(int64_t)3;
  if (end > cap) end = cap;
  return end > zb ? (int)(end - zb) : 0;
}

// all lanes: publish the tile's ray ends, lane 0: issue the bulk loads (n_s samples of rgbs, n_z >= n_s of z: a
// tile that leaves its last ray open also needs the depth after its last sample)
__device__ __forceinline__ void issue_tile_n(const PkPipe& pipe, int st, const PackedArgs& a, int64_t sb, int n_s,
                                             int n_z, int rel_end, int lane) {
  pipe.ends_stage(st)[lane] = rel_end;
  __syncwarp();
  if (lane == 0) {
    if (!kOutByTma) fence_proxy_async_smem();  // generic reads/writes of this slot precede the async refill
    const int shift = (int)(sb & 3);
    const int zel = z_bulk_elems(sb, n_z, a.S);
    const uint32_t rb = (uint32_t)n_s * 16u, zbytes = (uint32_t)zel * 4u;
    mbar_expect_tx(&pipe.bars[st], rb + zbytes);
    bulk_g2s(pipe.rgbs_stage(st), a.sp.rgbs + sb * 4, rb, &pipe.bars[st]);
    if (zel > 0) bulk_g2s(pipe.z_stage(st), a.sp.z + (sb - shift), zbytes, &pipe.bars[st]);
  }
}
__device__ __forceinline__ void issue_tile(const PkPipe& pipe, int st, const PackedArgs& a, const Item& it,
                                           int rel_end, int lane) {
  issue_tile_n(pipe, st, a, it.sb, (int)it.n_s, (int)it.n_s, rel_end, lane);
}

// after the mbarrier wait: fetch the (<= 4) staged elements of z the bulk copy could not cover
__device__ __forceinline__ void patch_z_tail_n(const PkPipe& pipe, int st, const PackedArgs& a, int64_t sb, int n_z, int lane) {
  const int shift = (int)(sb & 3);
  const int64_t zb = sb - shift;
  const int64_t covered = zb + z_bulk_elems(sb, n_z, a.S);
  const int64_t gi = covered + lane;
  if (gi < sb + n_z && gi < a.S) pipe.z_stage(st)[gi - zb] = a.sp.z[gi];  // a few lanes (warp-uniform loop-free)
  __syncwarp();
}
__device__ __forceinline__ void patch_z_tail(const PkPipe& pipe, int st, const PackedArgs& a, const Item& it, int lane) {
  patch_z_tail_n(pipe, st, a, it.sb, (int)it.n_s, lane);
}

// The lane's run inside a ragged tile: which ray its first sample belongs to and where (if
// anywhere) that ray ends inside the run.  ends[i] = end of the tile's i-th ray, INT_MAX padded.
// first_k0: position of the tile's first sample inside its ray (> 0 only in the streaming forward kernel, whose
// tiles may begin inside a ray)
template <int L>
__device__ __forceinline__ Run ragged_run(const int* ends, int n_s, int lane, int first_k0 = 0) {
  Run run;
  run.s0 = lane * L;
  const int rem = n_s - run.s0;
  run.nvalid = rem < 0 ? 0 : (rem > L ? L : rem);
  int cnt = 0;  // rays ending at or before s0
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    if (ends[cnt + step - 1] <= run.s0) cnt += step;
  }
  if (ends[cnt] <= run.s0) ++cnt;   // cnt in [0, 32]
  if (cnt > 31) cnt = 31;           // idle lanes only
  const int start = cnt > 0 ? ends[cnt - 1] : -first_k0;
  const int end = ends[cnt];
  run.ray0 = cnt;
  run.k0 = run.s0 - start;
  const int to_head = (run.k0 == 0) ? 0 : (end == INT_MAX ? L : end - run.s0);
  run.carry_len = to_head < run.nvalid ? to_head : run.nvalid;
  const int e = (end == INT_MAX) ? INT_MAX : end - 1 - run.s0;
  run.end_pos = e < run.nvalid ? e : -1;
  return run;
}

// Samples per lane for ONE tile.  The greedy packing of whole rays leaves a tile 77 % full on config 4's
// distribution (322 of 416 samples), and a tile body costs its L samples per lane whatever the fill.  The forward
// kernels therefore walk a tile of at most 288 samples with 9 samples per lane (odd: conflict-free; every tiled
// ray has more than kPkL samples, so a run still holds at most one ray boundary): 0.616 -> 0.586 ms on 2^20 rays of
// 8..256 samples.  More bodies lose to their code size — {7, 9, 11, 13}: forward 0.89 ms, backward 0.88 -> 1.80 ms;
// the backward kernel, whose body is twice as long, is slower even with two (0.876 -> 0.924 ms) and keeps one.
template <bool kTwoBodies, typename F>
__device__ __forceinline__ void with_run_length(int n_s, F&& body) {
  if (kTwoBodies && kPkL == 13 && AVR_PK_L2 < 13) {
    if (n_s <= 32 * AVR_PK_L2) {
      body(std::integral_constant<int, AVR_PK_L2>());
    } else {
      body(std::integral_constant<int, 13>());
    }
  } else {
    body(std::integral_constant<int, kPkL>());
  }
}

template <bool kWriteW>
__global__ void __launch_bounds__(kPkWarps * 32)
composite_fwd_span_packed_kernel(const PackedArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PkPipe pipe;
  pipe.init(smem, warp, lane);

  const int64_t gw = (int64_t)blockIdx.x * kPkWarps + warp;
  const int64_t n_warps = (int64_t)gridDim.x * kPkWarps;
  const int64_t ra = warp_lower_bound(a.offsets, a.R, (a.S / n_warps) * gw + min(gw, a.S % n_warps), lane);
  const int64_t rb = (gw + 1 == n_warps)
                         ? a.R
                         : warp_lower_bound(a.offsets, a.R, (a.S / n_warps) * (gw + 1) + min(gw + 1, a.S % n_warps), lane);
  const float4* rgbs4 = reinterpret_cast<const float4*>(a.sp.rgbs);

  int64_t tiles = 0;  // tiles issued so far (tile t lives in stage t % NS, parity (t / NS) & 1)
  int rel_end;
  Window win = load_window(a.offsets, ra, rb, lane);
  Item cur = next_item(win, ra, rb, lane, rel_end);
  int64_t cur_tile = -1;
  if (cur.kind == kItemTile) {
    cur_tile = tiles++;
    issue_tile(pipe, (int)(cur_tile % kPkStages), a, cur, rel_end, lane);
  }
  int64_t r_next = ra + cur.nr;
  win = load_window(a.offsets, r_next, rb, lane);

  while (cur.kind != kItemNone) {
    // ---- prepare the next item (its window was loaded one iteration ago) and start its loads
    Item nxt = next_item(win, r_next, rb, lane, rel_end);
    int64_t nxt_tile = -1;
    if (nxt.kind == kItemTile) {
      nxt_tile = tiles++;
      if (kOutByTma && kWriteW && lane == 0) bulk_wait_read<1>();  // the store that last read this stage is done
      issue_tile(pipe, (int)(nxt_tile % kPkStages), a, nxt, rel_end, lane);
    }
    r_next += nxt.nr;
    if (nxt.kind != kItemNone) win = load_window(a.offsets, r_next, rb, lane);

    // ---- process the current item
    if (cur.kind == kItemTile) {
      const int st = (int)(cur_tile % kPkStages);
      const int n_s = (int)cur.n_s;
      const int shift = (int)(cur.sb & 3);
      mbar_wait(&pipe.bars[st], (uint32_t)((cur_tile / kPkStages) & 1));
      patch_z_tail(pipe, st, a, cur, lane);
      with_run_length<true>(n_s, [&](auto len) {
        constexpr int L = decltype(len)::value;
        const Run run = ragged_run<L>(pipe.ends_stage(st), n_s, lane);
        const float4* rg = pipe.rgbs_stage(st) + run.s0;
        float* zs = pipe.z_stage(st) + shift + run.s0;
        fwd_tile_simple<L, kWriteW>(a.sp, run, rg, zs, cur.r0, lane);
      });
      if (kWriteW) {
        fence_proxy_async_smem();
        __syncwarp();
        // aligned interior by bulk store, up to 3 samples at either end by ordinary stores
        int a0 = (4 - shift) & 3;
        if (a0 > n_s) a0 = n_s;
        const int len = (n_s - a0) & ~3;
        const float* zt = pipe.z_stage(st) + shift;  // tile sample s at zt[s]
        float* wg = a.sp.w + cur.sb;
        if (kOutByTma) {
          if (lane == 0 && len > 0) bulk_s2g(wg + a0, zt + a0, (uint32_t)len * 4u);
          if (lane == 0) bulk_commit();
        } else {  // coalesced 16-byte stores: the stage is free again when this iteration ends
          const float4* src4 = reinterpret_cast<const float4*>(zt + a0);
          float4* dst4 = reinterpret_cast<float4*>(wg + a0);
          for (int v = lane; v < (len >> 2); v += 32) dst4[v] = src4[v];
        }
        if (lane < a0) wg[lane] = zt[lane];
        const int t0 = a0 + len;
        if (t0 + lane < n_s) wg[t0 + lane] = zt[t0 + lane];
        if (!kOutByTma) __syncwarp();
      } else {
        __syncwarp();
      }
    } else {
      wray_fwd_ray(rgbs4, a.sp.z, cur.sb, cur.n_s, cur.r0, a.sp.white_back, a.sp.infinity, a.sp.w, a.sp.rgb,
                   a.sp.depth, lane);
    }
    cur = nxt;
    cur_tile = nxt_tile;
  }
  if (kOutByTma && kWriteW && lane == 0) bulk_wait_all<0>();
}

__global__ void __launch_bounds__(kPkWarps * 32)
composite_bwd_span_packed_kernel(const PackedArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PkPipe pipe;
  pipe.init(smem, warp, lane);

  const int64_t gw = (int64_t)blockIdx.x * kPkWarps + warp;
  const int64_t n_warps = (int64_t)gridDim.x * kPkWarps;
  const int64_t ra = warp_lower_bound(a.offsets, a.R, (a.S / n_warps) * gw + min(gw, a.S % n_warps), lane);
  const int64_t rb = (gw + 1 == n_warps)
                         ? a.R
                         : warp_lower_bound(a.offsets, a.R, (a.S / n_warps) * (gw + 1) + min(gw + 1, a.S % n_warps), lane);
  const float4* rgbs4 = reinterpret_cast<const float4*>(a.sp.rgbs);
  float4* d4 = reinterpret_cast<float4*>(a.sp.d_rgbs);

  int64_t tiles = 0;
  int rel_end;
  Window win = load_window(a.offsets, ra, rb, lane);
  Item cur = next_item(win, ra, rb, lane, rel_end);
  int64_t cur_tile = -1;
  if (cur.kind == kItemTile) {
    cur_tile = tiles++;
    issue_tile(pipe, (int)(cur_tile % kPkStages), a, cur, rel_end, lane);
  }
  int64_t r_next = ra + cur.nr;
  win = load_window(a.offsets, r_next, rb, lane);

  while (cur.kind != kItemNone) {
    Item nxt = next_item(win, r_next, rb, lane, rel_end);
    int64_t nxt_tile = -1;
    if (nxt.kind == kItemTile) {
      nxt_tile = tiles++;
      if (kOutByTma && lane == 0) bulk_wait_read<1>();
      issue_tile(pipe, (int)(nxt_tile % kPkStages), a, nxt, rel_end, lane);
    }
    r_next += nxt.nr;
    if (nxt.kind != kItemNone) win = load_window(a.offsets, r_next, rb, lane);

    if (cur.kind == kItemTile) {
      const int st = (int)(cur_tile % kPkStages);
      const int n_s = (int)cur.n_s;
      const int shift = (int)(cur.sb & 3);
      mbar_wait(&pipe.bars[st], (uint32_t)((cur_tile / kPkStages) & 1));
      patch_z_tail(pipe, st, a, cur, lane);
      with_run_length<false>(n_s, [&](auto len) {
        constexpr int L = decltype(len)::value;
        const Run run = ragged_run<L>(pipe.ends_stage(st), n_s, lane);
        RayGrad gA{0.f, 0.f, 0.f, 0.f, 0.f}, gB{0.f, 0.f, 0.f, 0.f, 0.f};
        if (run.nvalid > 0) {
          gA = load_ray_grad<false>(a.sp, cur.r0 + run.ray0);  // the packed entry points have no camera-depth map
          gB = (run.end_pos >= 0 && run.end_pos + 1 < run.nvalid) ? load_ray_grad<false>(a.sp, cur.r0 + run.ray0 + 1) : gA;
        }
        float4* rg = pipe.rgbs_stage(st) + run.s0;
        float* zs = pipe.z_stage(st) + shift + run.s0;
        bwd_tile_simple<L, false>(a.sp, run, rg, zs, gA, gB, lane);
      });
      if (kOutByTma) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          bulk_s2g(a.sp.d_rgbs + cur.sb * 4, pipe.rgbs_stage(st), (uint32_t)n_s * 16u);
          bulk_commit();
        }
      } else {
        __syncwarp();
        const float4* src4 = pipe.rgbs_stage(st) + lane;
        float4* dst4 = d4 + cur.sb + lane;
#pragma unroll
        for (int it = 0; it < kPkL; ++it)  // 512-byte coalesced rows at constant offsets from two pointers
          if (it * 32 + lane < n_s) dst4[it * 32] = src4[it * 32];
        __syncwarp();
      }
    } else if (cur.n_s > 0) {
      wray_bwd_ray<false>(rgbs4, a.sp.z, cur.sb, cur.n_s, cur.r0, a.sp.g_rgb, a.sp.g_depth, nullptr,
                          a.sp.white_back, a.sp.infinity, d4, nullptr, lane);
    }
    cur = nxt;
    cur_tile = nxt_tile;
  }
  if (kOutByTma && lane == 0) bulk_wait_all<0>();
}

// ---- streaming forward -----------------------------------------------------------------
// The kernels above pack WHOLE rays into a tile, which leaves a 416-sample tile 77 % full on config 4's counts and
// cuts it short at every ray of <= kPkL samples.  The forward pass does not need whole rays: the warp walks its
// sample range front to back, so a ray may continue from one tile into the next with its transmittance and partial
// sums carried in registers.  A tile is then exactly kPkC consecutive samples unless it has to stop at a ray end
// (before a short ray, after 32 ray ends, at the end of the warp's range) — only tiles that stop at a ray end are
// partial, so every run of an open tile is full and the tile bodies' "no valid sample after the last ray end" rule
// still holds.  (The backward pass needs the transmittance from the ray's start AND the reverse scan from its end;
// it keeps whole-ray tiles.)
#ifndef AVR_PK_STREAM
#define AVR_PK_STREAM 1
#endif

struct StreamItem {
  int64_t r0;     // ray of the item's first sample
  int64_t sb;     // first sample
  int64_t n_ray;  // kItemRay: samples of the short ray
  int n_s;        // tile: samples
  int first_k0;   // tile: position of its first sample inside ray r0 (> 0: the ray continues from the tile before)
  bool open;      // tile: its last sample does not end a ray
  int kind;
};

// What starts at sample k_in of ray r (k_in > 0: inside a ray a tile left open).  Updates (r, k_in) to the state
// after the item; `rel_end` returns the lane's ray end relative to the tile start (INT_MAX unless the ray ends
// inside the tile).
__device__ __forceinline__ StreamItem next_stream_item(const Window& w, int64_t& r, int64_t& k_in, int64_t rb, int lane,
                                                       int& rel_end) {
  StreamItem it;
  it.r0 = r;
  it.sb = w.ostart + k_in;
  it.kind = kItemNone;
  it.n_ray = 0;
  it.n_s = 0;
  it.first_k0 = 0;
  it.open = false;
  rel_end = INT_MAX;
  if (r >= rb) return it;
  const bool valid = (r + lane < rb);
  const int64_t rel64 = valid ? (w.oend - it.sb) : (int64_t)INT_MAX;
  const int rel = rel64 > (int64_t)INT_MAX ? INT_MAX : (int)rel64;
  int64_t prev64 = __shfl_up_sync(0xffffffffu, rel64, 1);
  if (lane == 0) prev64 = -k_in;
  const int64_t cnt = rel64 - prev64;                       // the ray's length
  // cannot be (the next) part of a tile.  (The backward kernel's finer rule, ends_share_a_run, buys nothing here:
  // with tiles that stay full across rays the forward kernels run at the HBM roofline either way, 0.536 vs 0.541 ms.)
  const bool bad = !valid || cnt <= kPkL;
  const unsigned badmask = __ballot_sync(0xffffffffu, bad);
  const int first_bad = badmask ? (__ffs(badmask) - 1) : 32;
  if (first_bad == 0) {  // ray r is short (k_in == 0: a ray a tile left open is long)
    it.kind = kItemRay;
    it.n_ray = __shfl_sync(0xffffffffu, rel64, 0);
    r += 1;
    k_in = 0;
    return it;
  }
  const int limit = __shfl_sync(0xffffffffu, rel, first_bad - 1);  // end of the last ray a tile may take
  const int n_s = limit < kPkC ? limit : kPkC;
  const bool closed = lane < first_bad && rel <= n_s;
  const int nr_closed = __popc(__ballot_sync(0xffffffffu, closed));
  const int last_end = __shfl_sync(0xffffffffu, rel, nr_closed > 0 ? nr_closed - 1 : 0);
  it.kind = kItemTile;
  it.n_s = n_s;
  it.first_k0 = k_in > (int64_t)INT_MAX ? INT_MAX : (int)k_in;
  it.open = nr_closed == 0 || last_end < n_s;
  rel_end = closed ? rel : INT_MAX;
  k_in = it.open ? (nr_closed > 0 ? (int64_t)(n_s - last_end) : k_in + n_s) : 0;
  r += nr_closed;
  return it;
}

// fwd_tile_simple (span_bodies.cuh) with a ray carried in from the tile before (cT, cS: its transmittance and
// partial sums up to this tile; identity when the tile starts at a ray head) and out to the next one.
template <int L, bool kWriteW>
__device__ __forceinline__ void fwd_tile_stream(const SpanArgs& a, const Run& run, const float4* rg, float* zs,
                                                int64_t ray_base, int lane, bool carried_in, bool open, int n_s,
                                                float& cT, Sums& cS) {
  const int p = run.end_pos;
  float wl[L];
  float Tl = 1.0f;
  Sums A = zero_sums(), B = zero_sums();
  float zk = zs[0];
#pragma unroll
  for (int j = 0; j < L; ++j) {
    const bool last = (j == p);
    const bool in_a = (j <= p);
    const float4 c = rg[j];
    const float z_after = zs[j + 1];
    const float zn = last ? a.infinity : z_after;
    const float delta = last ? kLastDelta : zn - zk;
    const Opacity o = opacity(c.w, delta);
    const float w = o.alpha * Tl;
    wl[j] = w;
    if (in_a) {
      A.r += w * c.x;
      A.g += w * c.y;
      A.b += w * c.z;
      A.d += w * zn;
      A.a += w;
    } else {
      B.r += w * c.x;
      B.g += w * c.y;
      B.b += w * c.z;
      B.d += w * zn;
      B.a += w;
    }
    Tl = last ? 1.0f : Tl * o.t;
    zk = z_after;
  }
  const bool head0 = (run.k0 == 0);
  float T_in = Tl;
  Sums s_in = B;
  scan_fwd_exclusive(lane, head0 || p >= 0 || run.nvalid == 0, T_in, s_in);
  if (head0) {
    T_in = 1.0f;
    s_in = zero_sums();
  } else if (carried_in && run.ray0 == 0) {  // still inside the ray the tile before left open
    s_in.r = cS.r + cT * s_in.r;
    s_in.g = cS.g + cT * s_in.g;
    s_in.b = cS.b + cT * s_in.b;
    s_in.d = cS.d + cT * s_in.d;
    s_in.a = cS.a + cT * s_in.a;
    T_in = cT * T_in;
  }
  if (p >= 0) {
    Sums t;
    t.r = s_in.r + T_in * A.r;
    t.g = s_in.g + T_in * A.g;
    t.b = s_in.b + T_in * A.b;
    t.d = s_in.d + T_in * A.d;
    t.a = s_in.a + T_in * A.a;
    store_ray(a, ray_base + run.ray0, t);
  }
  // what the open ray has gathered up to the tile's end: the last lane's run after its ray end, if it has one,
  // else everything that entered the lane composed with the whole run
  if (open) {
    const int ll = (n_s - 1) / L;
    const bool fresh = p >= 0;
    const float oT = fresh ? Tl : T_in * Tl;
    Sums o;
    o.r = fresh ? B.r : s_in.r + T_in * B.r;
    o.g = fresh ? B.g : s_in.g + T_in * B.g;
    o.b = fresh ? B.b : s_in.b + T_in * B.b;
    o.d = fresh ? B.d : s_in.d + T_in * B.d;
    o.a = fresh ? B.a : s_in.a + T_in * B.a;
    cT = __shfl_sync(0xffffffffu, oT, ll);
    cS.r = __shfl_sync(0xffffffffu, o.r, ll);
    cS.g = __shfl_sync(0xffffffffu, o.g, ll);
    cS.b = __shfl_sync(0xffffffffu, o.b, ll);
    cS.d = __shfl_sync(0xffffffffu, o.d, ll);
    cS.a = __shfl_sync(0xffffffffu, o.a, ll);
  } else {
    cT = 1.0f;
    cS = zero_sums();
  }
  if (kWriteW) {
    __syncwarp();  // every lane has finished reading z from this stage
#pragma unroll
    for (int j = 0; j < L; ++j) zs[j] = (p < 0 || j <= p) ? wl[j] * T_in : wl[j];
  }
}

template <bool kWriteW>
__global__ void __launch_bounds__(kPkWarps * 32)
composite_fwd_stream_packed_kernel(const PackedArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PkPipe pipe;
  pipe.init(smem, warp, lane);

  const int64_t gw = (int64_t)blockIdx.x * kPkWarps + warp;
  const int64_t n_warps = (int64_t)gridDim.x * kPkWarps;
  const int64_t ra = warp_lower_bound(a.offsets, a.R, (a.S / n_warps) * gw + min(gw, a.S % n_warps), lane);
  const int64_t rb = (gw + 1 == n_warps)
                         ? a.R
                         : warp_lower_bound(a.offsets, a.R, (a.S / n_warps) * (gw + 1) + min(gw + 1, a.S % n_warps), lane);
  const float4* rgbs4 = reinterpret_cast<const float4*>(a.sp.rgbs);

  int64_t tiles = 0;  // tiles issued so far (tile t lives in stage t % NS, parity (t / NS) & 1)
  int rel_end;
  int64_t r = ra, k_in = 0;  // the walk's position: sample k_in of ray r
  Window win = load_window(a.offsets, r, rb, lane);
  StreamItem cur = next_stream_item(win, r, k_in, rb, lane, rel_end);
  int64_t cur_tile = -1;
  if (cur.kind == kItemTile) {
    cur_tile = tiles++;
    issue_tile_n(pipe, (int)(cur_tile % kPkStages), a, cur.sb, cur.n_s, cur.n_s + 1, rel_end, lane);
  }
  win = load_window(a.offsets, r, rb, lane);
  float cT = 1.0f;  // the ray the previous tile left open
  Sums cS = zero_sums();

  while (cur.kind != kItemNone) {
    // ---- prepare the next item (its window was loaded one iteration ago) and start its loads
    StreamItem nxt = next_stream_item(win, r, k_in, rb, lane, rel_end);
    int64_t nxt_tile = -1;
    if (nxt.kind == kItemTile) {
      nxt_tile = tiles++;
      if (kOutByTma && kWriteW && lane == 0) bulk_wait_read<1>();  // the store that last read this stage is done
      issue_tile_n(pipe, (int)(nxt_tile % kPkStages), a, nxt.sb, nxt.n_s, nxt.n_s + 1, rel_end, lane);
    }
    if (nxt.kind != kItemNone) win = load_window(a.offsets, r, rb, lane);

    // ---- process the current item
    if (cur.kind == kItemTile) {
      const int st = (int)(cur_tile % kPkStages);
      const int n_s = cur.n_s;
      const int shift = (int)(cur.sb & 3);
      mbar_wait(&pipe.bars[st], (uint32_t)((cur_tile / kPkStages) & 1));
      patch_z_tail_n(pipe, st, a, cur.sb, n_s + 1, lane);
      with_run_length<true>(n_s, [&](auto len) {
        constexpr int L = decltype(len)::value;
        const Run run = ragged_run<L>(pipe.ends_stage(st), n_s, lane, cur.first_k0);
        const float4* rg = pipe.rgbs_stage(st) + run.s0;
        float* zs = pipe.z_stage(st) + shift + run.s0;
        fwd_tile_stream<L, kWriteW>(a.sp, run, rg, zs, cur.r0, lane, cur.first_k0 > 0, cur.open, n_s, cT, cS);
      });
      if (kWriteW) {
        fence_proxy_async_smem();
        __syncwarp();
        // aligned interior by bulk store, up to 3 samples at either end by ordinary stores
        int a0 = (4 - shift) & 3;
        if (a0 > n_s) a0 = n_s;
        const int len = (n_s - a0) & ~3;
        const float* zt = pipe.z_stage(st) + shift;  // tile sample s at zt[s]
        float* wg = a.sp.w + cur.sb;
        if (kOutByTma) {
          if (lane == 0 && len > 0) bulk_s2g(wg + a0, zt + a0, (uint32_t)len * 4u);
          if (lane == 0) bulk_commit();
        } else {  // coalesced 16-byte stores: the stage is free again when this iteration ends
          const float4* src4 = reinterpret_cast<const float4*>(zt + a0);
          float4* dst4 = reinterpret_cast<float4*>(wg + a0);
          for (int v = lane; v < (len >> 2); v += 32) dst4[v] = src4[v];
        }
        if (lane < a0) wg[lane] = zt[lane];
        const int t0 = a0 + len;
        if (t0 + lane < n_s) wg[t0 + lane] = zt[t0 + lane];
        if (!kOutByTma) __syncwarp();
      } else {
        __syncwarp();
      }
    } else {
      wray_fwd_ray(rgbs4, a.sp.z, cur.sb, cur.n_ray, cur.r0, a.sp.white_back, a.sp.infinity, a.sp.w, a.sp.rgb,
                   a.sp.depth, lane);
    }
    cur = nxt;
    cur_tile = nxt_tile;
  }
  if (kOutByTma && kWriteW && lane == 0) bulk_wait_all<0>();
}

// ---- host side -----------------------------------------------------------------------
bool span_packed_eligible(const void* rgbs, const void* z, const void* w_or_null, const void* d_rgbs_or_null) {
  if (!option(OPT_PACKED_SPAN, 1)) return false;
  return aligned16(rgbs) && aligned16(z) && aligned16(w_or_null) && aligned16(d_rgbs_or_null);
}

template <typename KernelT>
static int packed_launch(KernelT kernel, const PackedArgs& a, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPkSmemBytes);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    (void)cudaGetLastError();
    return AVR_ERR_LAUNCH;
  }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kPkWarps * 32, kPkSmemBytes);
  if (e != cudaSuccess || occ < 1) {
    set_last_cuda_error(e);
    (void)cudaGetLastError();
    return AVR_ERR_LAUNCH;
  }
  const int sms = num_sms();
  // enough samples per warp to amortise the range search; never more warps than tiles
  int64_t want = (a.S / kPkC + kPkWarps) / kPkWarps;
  if (want < 1) want = 1;
  const int64_t cap = (int64_t)sms * occ;
  const int grid = (int)(want < cap ? want : cap);
  kernel<<<grid, kPkWarps * 32, kPkSmemBytes, stream>>>(a);
  return check_launch();
}

int launch_composite_fwd_span_packed(const float* rgbs, const float* z, const int64_t* offsets, int64_t R,
                                     int64_t S, int white_back, float infinity, float* w, float* rgb,
                                     float* depth, cudaStream_t stream) {
  if (R == 0) return AVR_OK;
  PackedArgs a{};
  a.sp.rgbs = rgbs;
  a.sp.z = z;
  a.sp.w = w;
  a.sp.rgb = rgb;
  a.sp.depth = depth;
  a.sp.white_back = white_back;
  a.sp.infinity = infinity;
  a.offsets = offsets;
  a.R = R;
  a.S = S;
#if AVR_PK_STREAM
  return w ? packed_launch(composite_fwd_stream_packed_kernel<true>, a, stream)
           : packed_launch(composite_fwd_stream_packed_kernel<false>, a, stream);
#else
  return w ? packed_launch(composite_fwd_span_packed_kernel<true>, a, stream)
           : packed_launch(composite_fwd_span_packed_kernel<false>, a, stream);
#endif
}

int launch_composite_bwd_span_packed(const float* rgbs, const float* z, const int64_t* offsets,
                                     const float* g_rgb, const float* g_depth, int64_t R, int64_t S,
                                     int white_back, float infinity, float* d_rgbs, cudaStream_t stream) {
  if (R == 0 || S == 0) return AVR_OK;
  PackedArgs a{};
  a.sp.rgbs = rgbs;
  a.sp.z = z;
  a.sp.g_rgb = g_rgb;
  a.sp.g_depth = g_depth;
  a.sp.d_rgbs = d_rgbs;
  a.sp.white_back = white_back;
  a.sp.infinity = infinity;
  a.offsets = offsets;
  a.R = R;
  a.S = S;
  return packed_launch(composite_bwd_span_packed_kernel, a, stream);
}

}  // namespace avr
