// Tile bodies and warp scans shared by the dense (composite_span.cu) and packed
// (composite_span_packed.cu) span kernels.  See composite_span.cu for the design.
#pragma once

#include "avr_common.cuh"

namespace avr {

constexpr int kMaxPeers = 16;

template <int L>
struct SpanCfg {
  static constexpr int kTileSamples = 32 * L;
  static constexpr int kRgbsBytes = kTileSamples * 16;
  static constexpr int kZBytes = kTileSamples * 4;
  // +16: the lane that owns the last sample reads one z past the tile (value unused)
  static constexpr int kStageBytes = kRgbsBytes + kZBytes + 16;
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
  float* d_z;             // bwd (nullable): gradient w.r.t. the depths (adaptive renderer)
  const float* depth_affine;  // nullable [R,2]: `depth` is the camera depth A*dist + B (see cam_depth)
  int64_t n_tiles;
  int K;
  int rays_per_tile;
  int tail_rays;          // rays in the last tile (== rays_per_tile when it is full)
  int white_back;
  float infinity;
  // fused all-gather (forward only): every finished ray also lands, packed as (r,g,b,depth),
  // in row peer_row0 + ray of each peer's gathered [world*R,4] buffer (peer memory mapped over
  // NVLink).  A warp owns a CONTIGUOUS range of rays in this mode and flushes them 32 at a
  // time: one coalesced 512-byte store per peer instead of per-ray 16-byte stores (which are
  // NVLink-packet-rate bound: 1.23 vs 0.93 ms per step measured on 2 GPUs).
  float4* peers[kMaxPeers];
  int n_peers;
  int peers_multicast;    // peers[0] is an NVSwitch multicast address: one multimem.st reaches every rank
  int64_t peer_row0;
  // completion signal of the fused all-gather (nullable): when the LAST CTA of the launch has
  // pushed its rows, it writes `signal_value` into word `signal_slot` of every peer's flag array
  // (release, system scope).  A consumer waits for all sources' words (avr_gather_wait) instead
  // of a cross-rank barrier, so no rank ever stalls on a peer that is merely behind.
  uint32_t* signal_flags[kMaxPeers];
  int n_signal;
  int signal_slot;
  uint32_t signal_value;
  unsigned* done_counter;  // local device word, zero between launches
};

// A lane's run inside a tile (identical for every full tile of a launch).
struct Run {
  int s0;         // first sample (tile-relative)
  int nvalid;     // samples of the run that exist (0..L)
  int k0;         // position of the first sample inside its ray
  int ray0;       // tile-relative ray of the first sample
  int carry_len;  // leading samples that belong to a ray started in an earlier lane
  int end_pos;    // simple bodies: run index of the sample that ends a ray, or -1
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
  int e = K - 1 - r.k0;
  r.end_pos = e < r.nvalid ? e : -1;
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

// `gather_ring` (fused all-gather only, else null): the warp's 64-entry shared-memory ring of finished rays,
// indexed by ray & 63 — what flush_rays_to_peers forwards, so that the peer stores need no trip through L2.
constexpr int kGatherRing = 64;
// `aff`: the ray's camera-depth pair if the caller fetched it ahead of the walk (fwd_tile_simple), else null and
// it is read here.
__device__ __forceinline__ void store_ray(const SpanArgs& a, int64_t ray, const Sums& t, float4* gather_ring = nullptr,
                                          const float2* aff = nullptr) {
  const float bg = a.white_back ? 1.0f - t.a : 0.f;
  float* o3 = a.rgb + ray * 3;
  const float r = t.r + bg, g = t.g + bg, b = t.b + bg;
  const float d = (aff && a.depth_affine) ? fmaf(aff->x, t.d, aff->y) : cam_depth(a.depth_affine, ray, t.d);
  o3[0] = r;
  o3[1] = g;
  o3[2] = b;
  a.depth[ray] = d;
  if (gather_ring) gather_ring[ray & (kGatherRing - 1)] = make_float4(r, g, b, d);
}

// rays [lo, hi) were finished by THIS warp (hi - lo <= 64, the ring's size): forward them to every peer's
// gathered buffer, lanes <-> consecutive rays, from the warp's shared-memory ring (the first version re-read
// them from global memory with ld.cg: a store -> L2 -> load round trip of ~1.5 us every ~11 tiles, 7 % of
// the forward kernel).  __syncwarp orders the ring writes of the other lanes before the reads.
__device__ __forceinline__ void flush_rays_to_peers(const SpanArgs& a, int64_t lo, int64_t hi, int lane,
                                                    const float4* gather_ring) {
  __syncwarp();
  for (int64_t ray = lo + lane; ray < hi; ray += 32) {
    const float4 v = gather_ring[ray & (kGatherRing - 1)];
    if (a.peers_multicast) {
      // the switch replicates the store into every rank's buffer (this rank's included): 16 B per ray
      // leave the GPU instead of 16 B per ray and peer
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a.peers[0] + a.peer_row0 + ray),
                   "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    } else {
      for (int p = 0; p < a.n_peers; ++p) a.peers[p][a.peer_row0 + ray] = v;
    }
  }
}

// End of a gather launch: every thread has issued its peer stores.  Count the CTA in; the last one
// publishes the launch's signal value to every peer.  Ordering: each thread fences its own (weak)
// peer stores at system scope, the CTA barrier carries them to thread 0, whose fence + atomic
// releases them; the last CTA's thread 0 has observed every CTA's atomic, fences, then stores the
// flags with release semantics — a consumer's ld.acquire.sys of the flag sees every row.
__device__ __forceinline__ void signal_gather_done(const SpanArgs& a) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(a.done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *a.done_counter = 0u;  // the next launch on this stream starts from zero
      __threadfence_system();
      for (int p = 0; p < a.n_signal; ++p)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.signal_flags[p] + a.signal_slot), "r"(a.signal_value)
                     : "memory");
    }
  }
}

struct RayGrad {
  float r, g, b, d, bg;  // g_rgb, g_depth, and the white-background term sum(g_rgb)
};
// kCam == false: the launch has no camera-depth map (a.depth_affine == nullptr) and the code for it is left
// out — with it, the K = 96 backward kernel went from 115 to 128 registers (one spill) and from 0.544 to
// 0.572 ms, the K = 192 one from 2.26 to 2.61 ms.
template <bool kCam = true>
__device__ __forceinline__ RayGrad load_ray_grad(const SpanArgs& a, int64_t ray) {
  RayGrad g{0.f, 0.f, 0.f, 0.f, 0.f};
  if (a.g_rgb) {
    g.r = a.g_rgb[ray * 3 + 0];
    g.g = a.g_rgb[ray * 3 + 1];
    g.b = a.g_rgb[ray * 3 + 2];
  }
  if (a.g_depth) g.d = kCam ? cam_depth_grad(a.depth_affine, ray, a.g_depth[ray]) : a.g_depth[ray];
  g.bg = a.white_back ? (g.r + g.g + g.b) : 0.f;
  return g;
}

// =====================================================================================
// Tile bodies.  rg / zs point at the lane's OWN first sample inside the stage.
// =====================================================================================

// ---- forward, K > L: at most one ray end per run, at run index p -----------------------
// Samples [0..p] (segment A) close the ray that contains the run's first sample; samples
// after p (segment B) open the next ray.  Without an end the whole run is segment B.
// Samples past the tile's end (idle lanes / partial tail tile) come after a ray end, so
// whatever stale shared memory they read only reaches aggregates nobody consumes.
template <int L, bool kWriteW>
__device__ __forceinline__ void fwd_tile_simple(const SpanArgs& a, const Run& run, const float4* rg, float* zs,
                                                int64_t ray_base, int lane, float4* gather_ring = nullptr) {
  const int p = run.end_pos;
  // the camera-depth pair of the ray this run closes, requested before the walk: read where the ray is stored, the
  // load's latency stalled the warp at every ray end (K = 96 forward with camera depth: 0.373 ms against 0.296 without)
  float2 aff = make_float2(1.0f, 0.0f);
  const bool aff_ahead = a.depth_affine && (reinterpret_cast<uintptr_t>(a.depth_affine) & 7u) == 0;  // 8-byte loads
  if (aff_ahead && p >= 0) aff = __ldg(reinterpret_cast<const float2*>(a.depth_affine) + (ray_base + run.ray0));
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
  }
  if (p >= 0) {
    Sums t;
    t.r = s_in.r + T_in * A.r;
    t.g = s_in.g + T_in * A.g;
    t.b = s_in.b + T_in * A.b;
    t.d = s_in.d + T_in * A.d;
    t.a = s_in.a + T_in * A.a;
    store_ray(a, ray_base + run.ray0, t, gather_ring, aff_ahead ? &aff : nullptr);
  }
  if (kWriteW) {
    __syncwarp();  // every lane has finished reading z from this stage
#pragma unroll
    for (int j = 0; j < L; ++j) zs[j] = (p < 0 || j <= p) ? wl[j] * T_in : wl[j];
  }
}

// ---- backward, K > L -------------------------------------------------------------------
// kDz: also form d_z (the adaptive renderer's depths carry grad, renderers.py:490-509):
//   d_z[k] = [k > 0] (dL/ddelta_{k-1} + g_depth * w_{k-1}) - [k < K-1] dL/ddelta_k,
//   dL/ddelta_k = dL/d(sigma*delta)_k * sigma_k.
// The values are written over the tile's z (once every lane has read it) and leave by bulk store.
template <int L, bool kDz>
__device__ __forceinline__ void bwd_tile_simple(const SpanArgs& a, const Run& run, float4* rg, float* zs,
                                                const RayGrad& gA, const RayGrad& gB, int lane) {
  const int p = run.end_pos;
  // walk 1, front to back: cache e_j and the local transmittance before sample j; sum the
  // first segment's weighted terms (they give the reverse-scan aggregate A)
  float ej[L], Tj[L];
  float Tl = 1.0f;
  Sums s = zero_sums();
  {
    float zk = zs[0];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const bool last = (j == p);
      const bool first_seg = (p < 0 || j <= p);
      const float4 c = rg[j];
      const float z_after = zs[j + 1];
      const float zn = last ? a.infinity : z_after;
      const float delta = last ? kLastDelta : zn - zk;
      const Opacity o = opacity(c.w, delta);
      ej[j] = o.e;
      Tj[j] = Tl;
      if (first_seg) {
        const float w = o.alpha * Tl;
        s.r += w * c.x;
        s.g += w * c.y;
        s.b += w * c.z;
        s.d += w * zn;
        s.a += w;
      }
      Tl = last ? 1.0f : Tl * o.t;
      zk = z_after;
    }
  }
  const bool head0 = (run.k0 == 0);
  const bool idle = (run.nvalid == 0);
  // reverse-scan element of this run: Q_left = A + B * Q_right
  float A = gA.r * s.r + gA.g * s.g + gA.b * s.b + gA.d * s.d - gA.bg * s.a;
  float Bm = Tl;
  if (p >= 0 || idle) Bm = 0.f;
  if (idle) A = 0.f;
  float T_in = scan_fwd_exclusive_T(lane, head0 || p >= 0 || idle, Tl);
  if (head0) T_in = 1.0f;
  float Q = scan_rev_exclusive(lane, A, Bm);

  // walk 2, back to front: final gradients, written over the rgbs stage in place
  float dzl[kDz ? L : 1];   // d_z of the run, completed below
  float dd_next = 0.f;      // dL/ddelta of sample j+1 (kDz)
  float a_last = 0.f;       // what the run's last sample sends to the NEXT lane's first sample (kDz)
  float zn = zs[L];
#pragma unroll
  for (int j = L - 1; j >= 0; --j) {
    const bool last = (j == p);
    const bool first_seg = (p < 0 || j <= p);
    const float4 c = rg[j];
    const float zk = zs[j];
    if (last) {
      Q = 0.f;
      zn = a.infinity;
    }
    const float delta = last ? kLastDelta : zn - zk;
    const float e = ej[j];
    const float alpha = 1.0f - e;
    const float t = (1.0f - alpha) + kTransEps;
    const float T = first_seg ? Tj[j] * T_in : Tj[j];
    const float gr = first_seg ? gA.r : gB.r;
    const float gg = first_seg ? gA.g : gB.g;
    const float gb = first_seg ? gA.b : gB.b;
    const float gd = first_seg ? gA.d : gB.d;
    const float gbg = first_seg ? gA.bg : gB.bg;
    const float g = gr * c.x + gg * c.y + gb * c.z + gd * zn - gbg;
    const float dalpha = T * (g - Q);
    Q = g * alpha + t * Q;
    const float dsd = dalpha * e;
    const float w = alpha * T;
    rg[j] = make_float4(w * gr, w * gg, w * gb, dsd * delta);
    if (kDz) {
      const float dd = last ? 0.f : dsd * c.w;           // dL/ddelta_j
      const float to_next = last ? 0.f : dd + gd * w;     // lands on z_{j+1}
      if (j == L - 1) a_last = to_next;
      else dzl[kDz ? j + 1 : 0] = to_next - dd_next;
      dd_next = dd;
    }
    zn = zk;
  }
  if (kDz) {
    float from_prev = __shfl_up_sync(0xffffffffu, a_last, 1);   // the previous lane's last sample (0 if it ended a ray)
    if (lane == 0) from_prev = 0.f;                              // tiles start at a ray head
    dzl[0] = from_prev - dd_next;
    __syncwarp();  // every lane has finished reading z from this stage
#pragma unroll
    for (int j = 0; j < L; ++j) zs[j] = dzl[kDz ? j : 0];
  }
}


}  // namespace avr
