// Importance sampler + merge by BUCKET RANKING (sample_fine + sample_depth + clamp + cat + sort,
// renderers.py:27-66, 255-258): the successor of the sorting-network kernels for the hot shapes.
//
// The sorting networks (importance_grp.cu / importance_reg.cu) are issue-bound: ~300 of their
// ~670 warp instructions per ray are compare-exchanges, sign flips and shuffles, whatever the
// data.  But the data has structure: all Kc + n (+ nd) depths of a ray live in [near, far(1+1/Kc)],
// the coarse ones stratified, the new ones drawn from a piecewise-uniform density.  So the merge is
// done as a bucket sort BY VALUE, which is exact for any input:
//
//   1. every element (coarse depth, new sample, clamped "depth" sample) computes a LEVEL
//      q = round((v - near) * scale), clamped — the same monotone fp32 function of its VALUE for
//      every element, so elements in lower levels are <= elements in higher levels, bit for bit;
//      it counts itself into a table of per-level counters in shared memory (three 8-bit counters
//      per word; one atomicAdd, whose return value is the element's arrival order in its level);
//   2. the group turns the counters into exclusive prefix sums IN PLACE (byte 3 of a word = the
//      prefix at the word's end, so a level's start and end are in ONE word):
//      x * 0x01010100 + r * 0x01010101 does a word's four prefixes in two IMADs;
//   3. an element alone in its level stores its value at out[start]; elements that share a level
//      are parked at tmp[start + arrival] and rank themselves among the level's members by value
//      (ties by arrival), a loop of `count` steps — 1.3 on average, levels being ~0.5 occupied;
//   4. the row is read back in order and stored with full 16-byte coalesced writes.
//
// Cost per element is ~constant (~50 thread instructions against ~110 for the networks) and does
// not depend on a power-of-two box, which is what the packed (ragged) layout needs.  No assumption
// is made about the coarse depths being sorted.  Everything before step 1 (cdf scan with running
// max, breadth-first cdf tree, branch-free descent, the reference's rounding) is the code of
// importance_grp.cu, so cdf / idx / z_fine are bit-identical to it.
#include <math_constants.h>

#include "avr_common.cuh"
#include "importance_args.cuh"
#include "kernels.h"
#include "sort_net.cuh"   // ilog2_c

namespace avr {

__host__ __device__ constexpr int bins_pow2ceil(int v) { return v <= 1 ? 1 : 2 * bins_pow2ceil((v + 1) / 2); }

template <int G, int KC, int NI, int ND, int WPB>
struct BinsCfg {
  static constexpr int RPW = 32 / G;                 // rays per warp
  static constexpr int NIL = (NI + G - 1) / G, NDL = (ND + G - 1) / G, CW = KC / G;
  static constexpr int T = KC + NI + ND;
  static constexpr int FB = T <= 255 ? 8 : 16;       // bits per counter field
  static constexpr int LPW = FB == 8 ? 3 : 1;        // levels per word (the last field is the word's end prefix)
  static constexpr unsigned FMASK = FB == 8 ? 0xffu : 0xffffu;
  static constexpr int W = KC * WPB;                 // counter words per ray
  static constexpr int WL = W / G;                   // words per lane in the prefix pass
  static constexpr int Q = W * LPW;                  // levels
  static constexpr int DEPTH = ilog2_c(KC);
  static constexpr int TREE_STRIDE = KC + 8;
  static constexpr int TS = ((T + 3) & ~3) + 4;      // out / tmp row stride (floats)
  static constexpr int SMEM_PER_RAY = (TREE_STRIDE + W + 2 * TS) * 4;
  static constexpr int WARPS = (8 * RPW * SMEM_PER_RAY <= 48 * 1024) ? 8 : ((4 * RPW * SMEM_PER_RAY <= 48 * 1024) ? 4 : 2);
  static_assert(G == 8 || G == 16 || G == 32, "group width");
  static_assert((KC & (KC - 1)) == 0 && KC % G == 0, "KC: power of two, split evenly over the group");
  static_assert(WL % 4 == 0, "each lane owns whole uint4s of the counter table");
  static_assert(T < 1024 && W < 4096, "packed per-element state");
  static_assert(WARPS * RPW * SMEM_PER_RAY <= 48 * 1024, "static shared memory");
};

// The class of a packed ray for the ragged kernels: the smallest (KC, NI) box that holds its kc
// coarse and n new samples, or -1 (empty ray, or larger than the largest box).
__host__ __device__ inline int bins_ragged_class(int kc, int n) {
  if (kc < 1 || kc > 256 || n > 128) return -1;
  if (kc <= 32 && n <= 16) return 0;
  if (kc <= 64 && n <= 32) return 1;
  if (kc <= 128 && n <= 64) return 2;
  return 3;
}

template <int G, int KC, int NI, int ND, bool kRagged, int kClass, int WPB>
__global__ void __launch_bounds__(BinsCfg<G, KC, NI, ND, WPB>::WARPS * 32, 32 / BinsCfg<G, KC, NI, ND, WPB>::WARPS)
importance_bins_kernel(const ImportanceRegArgs a) {
  using C = BinsCfg<G, KC, NI, ND, WPB>;
  constexpr int RPW = C::RPW, NIL = C::NIL, NDL = C::NDL, CW = C::CW, DEPTH = C::DEPTH, kWarps = C::WARPS;
  constexpr int FB = C::FB, LPW = C::LPW, W = C::W, WL = C::WL, Q = C::Q, TS = C::TS;
  constexpr unsigned FMASK = C::FMASK;
  constexpr bool kNearLevel = ND > 0;     // the clamped "depth" samples all equal `near`: give that value its own level
  static_assert(!kRagged || ND == 0, "the packed entry point has no depth samples");
  __shared__ __align__(16) float s_tree[kWarps][RPW][C::TREE_STRIDE];
  __shared__ __align__(16) unsigned s_cnt[kWarps][RPW][W];
  __shared__ __align__(16) float s_out[kWarps][RPW][TS];
  __shared__ __align__(16) float s_tmp[kWarps][RPW][TS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane & (G - 1), sub = lane / G;
  float* tree = s_tree[warp][sub];
  unsigned* cnt = s_cnt[warp][sub];
  float* out = s_out[warp][sub];
  float* tmp = s_tmp[warp][sub];
  const bool do_sort = (a.z_sorted != nullptr);
  const float inv_kc = 1.0f / (float)KC;  // KC is a power of two: x * inv_kc == x / KC bit for bit

  // tree slots of this lane's cdf entries q = g*CW + i + 1 (breadth-first order, root = 1);
  // the last entry q == KC is not part of the tree: it is dumped into the unused slot 0
  int hidx[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    const int q = g * CW + i + 1;
    const int tz = __ffs(q) - 1;
    hidx[i] = (q < KC) ? (1 << (DEPTH - 1 - tz)) + (q >> (tz + 1)) : 0;
  }
  {
    uint4* c4 = reinterpret_cast<uint4*>(cnt) + g * (WL / 4);
#pragma unroll
    for (int i = 0; i < WL / 4; ++i) c4[i] = make_uint4(0u, 0u, 0u, 0u);
  }

  auto process = [&](const int64_t r, const bool live) {
    const int64_t bi = a.bound_stride ? r : 0;
    const float near = a.near[bi], far = a.far[bi];
    const float span = __fsub_rn(far, near);
    int kc = KC, n = NI;
    int64_t cbase = r * KC, fbase = r * NI;
    if (kRagged) {
      cbase = a.offsets[r];
      kc = (int)(a.offsets[r + 1] - cbase);
      fbase = a.fine_offsets[r];
      n = (int)(a.fine_offsets[r + 1] - fbase);
    }

    // ---- loads ------------------------------------------------------------------------------
    float w[CW], uu[NIL > 0 ? NIL : 1], jj[NIL > 0 ? NIL : 1], nn[NDL > 0 ? NDL : 1], zc[CW];
    if (kRagged) {
      const float* wrow = a.weights + cbase;
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        const int j = g * CW + i;
        w[i] = wrow[j < kc ? j : kc - 1];
      }
#pragma unroll
      for (int q = 0; q < NIL; ++q) {
        const int e = q * G + g;                      // striped: coalesced, and a short ray keeps every lane busy
        const int ec = n > 0 ? (e < n ? e : n - 1) : 0;
        uu[q] = n > 0 ? a.u[fbase + ec] : 0.f;
        jj[q] = n > 0 ? a.u2[fbase + ec] : 0.f;
      }
      if (do_sort) {
        const float* zrow = a.z_coarse + cbase;
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const int j = i * G + g;
          zc[i] = zrow[j < kc ? j : kc - 1];
        }
      }
    } else {
      if (a.vecw) {
#pragma unroll
        for (int i = 0; i < CW; i += 4) {
          const float4 p = *reinterpret_cast<const float4*>(a.weights + r * KC + g * CW + i);
          w[i] = p.x; w[(i + 1) % CW] = p.y; w[(i + 2) % CW] = p.z; w[(i + 3) % CW] = p.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) w[i] = a.weights[r * KC + g * CW + i];
      }
#pragma unroll
      for (int q = 0; q < NIL; ++q) {       // striped (coalesced 4-byte loads; the sample order is free)
        uu[q] = a.u[r * NI + q * G + g];
        jj[q] = a.u2[r * NI + q * G + g];
      }
      if (do_sort) {
#pragma unroll
        for (int q = 0; q < NDL; ++q) nn[q] = a.normals[r * ND + q * G + g];
#pragma unroll
        for (int i = 0; i < CW; ++i) zc[i] = a.z_coarse[r * KC + i * G + g];
      }
    }

    // level of a value: the SAME monotone function for every element of the ray
    const float kcf = (float)kc;
    // values reach near + span*(kc+1)/kc (bin index kc, renderers.py:42-43); any positive scale is
    // correct, this one spreads [near, that] over the levels
    const float qtop = (float)(Q - 1 - (kNearLevel ? 2 : 0));
    const float scale = span > 0.f ? __fdividef(qtop * kcf, span * (kcf + 1.0f)) : 0.f;
    auto count_in = [&](const float v) -> unsigned {
      float t = __fmul_rn(__fsub_rn(v, near), scale);
      t = fminf(fmaxf(t, 0.f), qtop);
      int q = __float_as_int(__fadd_rn(t, 8388608.0f)) & 0x7fffff;      // round to nearest: monotone
      if (kNearLevel) q += (v > near) ? 2 : ((v == near) ? 1 : 0);
      unsigned word, k;
      if (LPW == 3) {
        word = __umulhi((unsigned)q, 0x55555556u);
        k = (unsigned)q - 3u * word;
      } else {
        word = (unsigned)q;
        k = 0u;
      }
      const unsigned sh = k * FB;
      const unsigned old = atomicAdd(&cnt[word], 1u << sh);
      return word | (k << 12) | (((old >> sh) & FMASK) << 16);
    };

    // coarse depths can be counted at once (their loads are the only dependency)
    unsigned stc[CW];
    __syncwarp();  // the previous ray's readers of tree / out / tmp and the zeroing of cnt are done
    if (do_sort) {
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        stc[i] = 0u;
        if (!kRagged || i * G + g < kc) stc[i] = count_in(zc[i]);
      }
    }

    // ---- 1. cdf (renderers.py:36-39): blocked scan + running max -----------------------------
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      w[i] = (!kRagged || g * CW + i < kc) ? __fadd_rn(w[i], kPdfEps) : 0.f;
      part += w[i];
    }
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    const float S = part;
    const float rS = kRagged ? __fdiv_rn(1.0f, S) : 0.f;
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      run += kRagged ? __fmul_rn(w[i], rS) : __fdiv_rn(w[i], S);
      w[i] = run;
    }
    float incl = run;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const float p = __shfl_up_sync(0xffffffffu, incl, d);
      if (g >= d) incl += p;
    }
    float off = __shfl_up_sync(0xffffffffu, incl, 1);
    if (g == 0) off = 0.f;
    float mx = off + run;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const float p = __shfl_up_sync(0xffffffffu, mx, d);
      if (g >= d) mx = fmaxf(mx, p);
    }
    float floor_prev = __shfl_up_sync(0xffffffffu, mx, 1);
    if (g == 0) floor_prev = 0.f;
    // packed layout: ray r's kc + 1 cdf entries start at offsets[r] + r
    float* cdf_out = (a.cdf && live) ? a.cdf + (kRagged ? cbase + r : r * (KC + 1)) : nullptr;
    float last = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      float val = fmaxf(off + w[i], floor_prev);
      if (kRagged && g * CW + i >= kc) val = CUDART_INF_F;
      tree[hidx[i]] = val;
      if (cdf_out && (!kRagged || g * CW + i < kc)) cdf_out[g * CW + i + 1] = val;
      last = val;
    }
    if (cdf_out && g == 0) cdf_out[0] = 0.f;
    last = __shfl_sync(0xffffffffu, last, lane | (G - 1));  // cdf[KC]
    __syncwarp();

    // ---- 2. this lane's new samples: search, place, count ---------------------------------------
    float v[NIL + NDL > 0 ? NIL + NDL : 1];
    unsigned stn[NIL + NDL > 0 ? NIL + NDL : 1];
    if (NIL > 0) {
      unsigned node[NIL > 0 ? NIL : 1];
#pragma unroll
      for (int q = 0; q < NIL; ++q) node[q] = 1;
#pragma unroll
      for (int step = 0; step < DEPTH; ++step) {
#pragma unroll
        for (int q = 0; q < NIL; ++q) {
          const float d = __fsub_rn(uu[q], tree[node[q]]);
          node[q] = 2 * node[q] + 1 - (__float_as_uint(d) >> 31);
        }
      }
#pragma unroll
      for (int q = 0; q < NIL; ++q) {
        const float dl = __fsub_rn(uu[q], last);
        const int bin = (int)(node[q] - KC) + 1 - (int)(__float_as_uint(dl) >> 31);
        const float num = __fadd_rn((float)bin, jj[q]);                    // renderers.py:45
        const float t = kRagged ? __fdiv_rn(num, kcf) : __fmul_rn(num, inv_kc);
        v[q] = __fadd_rn(near, __fmul_rn(span, t));                          // :46
        const int e = q * G + g;
        const bool has = !kRagged || e < n;
        if (live && has) {
          if (a.idx) a.idx[fbase + e] = bin;
          if (a.z_fine) a.z_fine[fbase + e] = v[q];
        }
        stn[q] = 0u;
        if (do_sort && has) stn[q] = count_in(v[q]);
      }
    }
    if (!do_sort) return;
#pragma unroll
    for (int q = 0; q < NDL; ++q) {
      // sample_depth's randn*std (the depth is NOT added), clamped (renderers.py:62-66, :255)
      v[NIL + q] = fminf(fmaxf(__fmul_rn(nn[q], a.depth_std), near), far);
      stn[NIL + q] = count_in(v[NIL + q]);
    }
    __syncwarp();

    // ---- 3. counters -> exclusive prefixes, in place ---------------------------------------------
    {
      uint4* c4 = reinterpret_cast<uint4*>(cnt) + g * (WL / 4);
      unsigned wv[WL];
#pragma unroll
      for (int i = 0; i < WL / 4; ++i) {
        const uint4 p = c4[i];
        wv[4 * i] = p.x; wv[4 * i + 1] = p.y; wv[4 * i + 2] = p.z; wv[4 * i + 3] = p.w;
      }
      unsigned s = 0;
#pragma unroll
      for (int i = 0; i < WL; ++i) s = (LPW == 3) ? __dp4a(wv[i], 0x01010101u, s) : s + wv[i];
      unsigned run_s = s;
#pragma unroll
      for (int d = 1; d < G; d <<= 1) {
        const unsigned p = __shfl_up_sync(0xffffffffu, run_s, d);
        if (g >= d) run_s += p;
      }
      unsigned rr = run_s - s;
#pragma unroll
      for (int i = 0; i < WL; ++i) {
        unsigned x;
        if (LPW == 3) {
          x = wv[i] * 0x01010100u + rr * 0x01010101u;
          rr = x >> 24;
        } else {
          x = rr | ((rr + wv[i]) << 16);
          rr += wv[i];
        }
        wv[i] = x;
      }
#pragma unroll
      for (int i = 0; i < WL / 4; ++i) c4[i] = make_uint4(wv[4 * i], wv[4 * i + 1], wv[4 * i + 2], wv[4 * i + 3]);
    }
    __syncwarp();

    // ---- 4. placement: alone in the level -> final slot; shared level -> parked by arrival ---------
    auto place = [&](const float val, const unsigned st) -> unsigned {
      const unsigned x = cnt[st & 0xfffu];
      const unsigned sh = ((st >> 12) & 3u) * FB;
      const unsigned p = (x >> sh) & FMASK;
      const unsigned c = ((x >> (sh + FB)) & FMASK) - p;
      const unsigned o = st >> 16;
      if (c == 1u) out[p] = val;
      else tmp[p + o] = val;
      return p | (c << 10) | (o << 20);
    };
#pragma unroll
    for (int i = 0; i < CW; ++i)
      if (!kRagged || i * G + g < kc) stc[i] = place(zc[i], stc[i]);
#pragma unroll
    for (int q = 0; q < NIL + NDL; ++q)
      if (!kRagged || q >= NIL || q * G + g < n) stn[q] = place(v[q], stn[q]);
    __syncwarp();
    {  // the table is free again: zero it for the next ray while the ranks are being fixed
      uint4* c4 = reinterpret_cast<uint4*>(cnt) + g * (WL / 4);
#pragma unroll
      for (int i = 0; i < WL / 4; ++i) c4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // ---- 5. members of a shared level rank themselves by value (ties by arrival) ---------------------
    auto settle = [&](const float val, const unsigned st) {
      const unsigned p = st & 0x3ffu, c = (st >> 10) & 0x3ffu, o = st >> 20;
      if (c > 1u && !(kNearLevel && val == near)) {
        unsigned rank = 0;
        for (unsigned i = 0; i < c; ++i) {
          const float f = tmp[p + i];
          rank += (f < val || (f == val && i < o)) ? 1u : 0u;
        }
        out[p + rank] = val;
      } else if (c > 1u) {
        out[p + o] = val;          // the level of `near` itself: all members are equal
      }
    };
#pragma unroll
    for (int i = 0; i < CW; ++i)
      if (!kRagged || i * G + g < kc) settle(zc[i], stc[i]);
#pragma unroll
    for (int q = 0; q < NIL + NDL; ++q)
      if (!kRagged || q >= NIL || q * G + g < n) settle(v[q], stn[q]);
    __syncwarp();

    // ---- 6. the row, in order ------------------------------------------------------------------------
    if (live) {
      if (!kRagged && a.vecz) {
        float4* orow = reinterpret_cast<float4*>(a.z_sorted + r * C::T);
        const float4* o4 = reinterpret_cast<const float4*>(out);
#pragma unroll
        for (int i = 0; i < (C::T / 4 + G - 1) / G; ++i) {
          const int c4i = i * G + g;
          if (c4i < C::T / 4) orow[c4i] = o4[c4i];
        }
      } else {
        float* orow = a.z_sorted + (kRagged ? cbase + fbase : r * C::T);
        const int total = kRagged ? kc + n : C::T;
#pragma unroll
        for (int i = 0; i < (C::T + G - 1) / G; ++i) {
          const int q = i * G + g;
          if (q < total) orow[q] = out[q];
        }
      }
    }
  };

  if (kRagged) {
    const int64_t step = (int64_t)gridDim.x * kWarps * 32;
    for (int64_t base = (blockIdx.x * (int64_t)kWarps + warp) * 32; base < a.R; base += step) {
      const int64_t rr = base + lane;
      bool mine = false;
      if (rr < a.R) {
        const int kc = (int)(a.offsets[rr + 1] - a.offsets[rr]);
        const int n = (int)(a.fine_offsets[rr + 1] - a.fine_offsets[rr]);
        mine = bins_ragged_class(kc, n) == kClass;
      }
      unsigned todo = __ballot_sync(0xffffffffu, mine);
      while (todo) {
        const unsigned pick = __fns(todo, 0, sub + 1);
        const bool live = pick != 0xffffffffu;
        const int64_t r = base + (live ? (int)pick : __ffs(todo) - 1);
#pragma unroll
        for (int k = 0; k < RPW; ++k) todo &= todo - 1;
        process(r, live);
      }
    }
  } else {
    const int64_t n_wg = (a.R + RPW - 1) / RPW;
    for (int64_t wg = blockIdx.x * (int64_t)kWarps + warp; wg < n_wg; wg += (int64_t)gridDim.x * kWarps) {
      const int64_t r_raw = wg * RPW + sub;
      process(r_raw < a.R ? r_raw : a.R - 1, r_raw < a.R);
    }
  }
}

template <int G, int KC, int NI, int ND, int WPB>
static int launch_bins(const ImportanceRegArgs& a, cudaStream_t stream) {
  using C = BinsCfg<G, KC, NI, ND, WPB>;
  const int64_t n_wg = (a.R + C::RPW - 1) / C::RPW;
  int64_t blocks = (n_wg + C::WARPS - 1) / C::WARPS;
  const int64_t cap = (int64_t)num_sms() * (32 / C::WARPS) * 4;
  if (blocks > cap) blocks = cap;
  importance_bins_kernel<G, KC, NI, ND, false, 0, WPB><<<(unsigned)blocks, C::WARPS * 32, 0, stream>>>(a);
  return check_launch();
}

template <int G, int KC, int NI, int kClass, int WPB>
static int launch_bins_ragged(const ImportanceRegArgs& a, cudaStream_t stream) {
  using C = BinsCfg<G, KC, NI, 0, WPB>;
  int64_t blocks = (a.R + C::WARPS * 32 - 1) / (C::WARPS * 32);
  const int64_t cap = (int64_t)num_sms() * (32 / C::WARPS) * 2;
  if (blocks > cap) blocks = cap;
  importance_bins_kernel<G, KC, NI, 0, true, kClass, WPB><<<(unsigned)blocks, C::WARPS * 32, 0, stream>>>(a);
  return check_launch();
}

// Packed layout: one launch per ray class that the caller's maxima allow; rays outside every class
// (kc > 256 or n > 128) are left to the caller (importance_reg.cu), which is told by `*covers_all`.
int launch_importance_bins_ragged(const ImportanceRegArgs& a, int max_coarse, int max_fine, bool* covers_all,
                                  cudaStream_t stream) {
  *covers_all = (max_coarse <= 256 && max_fine <= 128);
  int rc = launch_bins_ragged<8, 32, 16, 0, 2>(a, stream);
  if (rc == AVR_OK && (max_coarse > 32 || max_fine > 16)) rc = launch_bins_ragged<16, 64, 32, 1, 2>(a, stream);
  if (rc == AVR_OK && (max_coarse > 64 || max_fine > 32)) rc = launch_bins_ragged<16, 128, 64, 2, 2>(a, stream);
  if (rc == AVR_OK && (max_coarse > 128 || max_fine > 64)) rc = launch_bins_ragged<32, 256, 128, 3, 2>(a, stream);
  return rc;
}

// Dense shapes compiled with the shape as constants (the same list as importance_grp.cu).
int launch_importance_bins(const ImportanceRegArgs& a, cudaStream_t stream) {
  if (a.offsets) return AVR_ERR_UNSUPPORTED;
#define AVR_BINS_CASE(G_, KC_, NI_, ND_, WPB_) \
  if (a.Kc == KC_ && a.n_imp == NI_ && a.n_depth == ND_) return launch_bins<G_, KC_, NI_, ND_, WPB_>(a, stream);
  AVR_BINS_CASE(16, 64, 128, 0, 2)
  AVR_BINS_CASE(16, 64, 16, 16, 2)
  AVR_BINS_CASE(16, 64, 16, 0, 2)
  AVR_BINS_CASE(16, 64, 64, 0, 2)
  AVR_BINS_CASE(8, 32, 8, 8, 2)
  AVR_BINS_CASE(8, 32, 8, 0, 2)
#undef AVR_BINS_CASE
  return AVR_ERR_UNSUPPORTED;
}

}  // namespace avr
