// Importance sampler + merge for the hot dense shapes and the packed ray classes, G = 8 / 16 / 32 lanes per ray
// (four / two / one ray per warp).
//
// Same algorithm as importance_reg.cu (cdf scan -> inverse-CDF search -> jitter -> register
// bitonic sort of the new samples -> one bitonic merge with the ascending coarse depths;
// renderers.py:27-66, 255-258), re-mapped so that more of it is lane-local:
//   * a lane owns KC/G coarse bins and (NI+ND)/G new samples, so at G = 8 22 of the 28 compare-exchange
//     stages of a 128-key sort (and 5 of the 8 merge stages) are register-to-register; only
//     the stages over the lane bits inside a group need a shuffle;
//   * the per-ray fixed work (sum, scan, running max, bounds, addressing) is issued once per
//     32/G rays;
//   * the cdf is searched as a perfect binary tree over cdf[1..KC-1] (log2 KC probes, stored
//     breadth-first so a probe step's nodes are contiguous) plus one compare against the last
//     entry, which stays in a register; with four or more draws per lane the first two levels
//     come from registers.
// What limits the kernel is the SM's LSU data pipe — shuffles, shared-memory and global accesses share
// it, one 128-byte wavefront per cycle — and, once that is relieved, issue slots and the ALU pipe
// together (profiles/r02_lsu_wavefronts.md).  Hence: dense rows move as 16-byte pieces DEALT over a
// ray's lanes (G*16 contiguous bytes per ray and instruction), the merged keys live in a layout
// where a lane owns groups of four consecutive keys (sort_net.cuh: merge_net<..., kQuad>), and every
// dense shape runs at 8 lanes per ray.  Shapes are template constants; anything else takes
// importance_reg.cu.
#include <math_constants.h>

#include <cstdlib>

#include "avr_common.cuh"
#include "importance_args.cuh"
#include "kernels.h"
#include "sort_net.cuh"

namespace avr {



__host__ __device__ constexpr int pow2ceil_c(int v) { return v <= 1 ? 1 : 2 * pow2ceil_c((v + 1) / 2); }

template <int G, int KC, int NI, int ND>
struct GrpCfg {
  static constexpr int RPW = 32 / G;                 // rays per warp
  static constexpr int NIL = NI / G, NDL = ND / G;   // importance / depth samples per lane
  static constexpr int EPF = pow2ceil_c(NIL + NDL);  // new samples per lane incl. +inf padding
  static constexpr int M = G * EPF;
  static constexpr int CW = KC / G;                  // coarse bins per lane
  static constexpr int EPT = pow2ceil_c(CW + EPF);   // merged samples per lane incl. padding
  static constexpr int EPC = EPT - EPF;
  static constexpr int P = G * EPT;
  static constexpr int TOTAL = KC + NI + ND;
  static constexpr int DEPTH = ilog2_c(KC);          // perfect tree over cdf[1..KC-1]
  static constexpr int TREE_STRIDE = KC + 8;         // +8 floats: the rays of a warp start in different banks
  static constexpr int PAD = EPF >= 8 ? 4 : 0;       // floats after each lane's chunk of the exchange buffer
  static constexpr int BUF_FLOATS = (P > M + G * PAD ? P : M + G * PAD);
  static constexpr int BUF_STRIDE = BUF_FLOATS + 8;
  static constexpr unsigned kInf0 = ((EPC >= 32 ? 0xffffffffu : ((1u << EPC) - 1u)) & ~((1u << CW) - 1u));
  // quad layout of the merged keys (dense rays): register i of lane g holds key (i >> 2)*4G + 4g + (i & 3).
  // A group of four registers is +inf on every lane iff its 4G keys lie inside the padding [KC, P - M).
  static constexpr int QG = 4 * G;
  static constexpr unsigned quad_inf_mask() {
    unsigned m = 0;
    for (int hi = 0; hi * 4 < EPT; ++hi)
      if (hi * QG >= KC && (hi + 1) * QG <= P - M) m |= 0xfu << (4 * hi);
    return m;
  }
  static constexpr unsigned kInfQ = quad_inf_mask();
  // warps per CTA: as many as the static shared-memory limit (48 KiB) allows, at most 8
  static constexpr int SMEM_PER_WARP = RPW * (TREE_STRIDE + BUF_STRIDE) * 4;
  static constexpr int WARPS = (8 * SMEM_PER_WARP <= 48 * 1024) ? 8 : 4;
  // register budget: 64/thread (32 resident warps) where the per-lane arrays are small; otherwise 80 (24 warps) —
  // the dense shapes fit that without a spill, and the 64 -> 128 kernel runs 0.628 -> 0.605 ms against the former
  // 128 / 16 warps (95 registers / 20 warps: 0.606, 72 / 28 warps with 16 B spilled: 0.625, 64 with 44 B: 0.638)
  // ... and 48 (40 warps, 12 B spilled) for the largest ragged class, one ray per warp: 1.239 -> 1.212 ms on config 4's
  // distribution (24 warps at 78 registers: 1.283, 48 warps at 40 registers: 1.281)
  // (the second largest class, two rays per warp: 24 / 32 / 40 warps within 1 %)
  // conf/default.conf's shape (64 -> 16 + 16): 48 registers / 40 warps without a spill, 0.239 -> 0.233 ms with the merge
  // (24 warps: 0.268, 48 warps at 40 registers with 36 B spilled: 0.248)
  static constexpr int MIN_BLOCKS = (G == 32 && KC == 256)                 ? (40 / WARPS)
                                    : (G == 8 && KC == 64 && NI == 16)     ? (40 / WARPS)
                                    : (EPT <= 16 && (G >= 16 || EPF <= 4)) ? (32 / WARPS)
                                                                           : (24 / WARPS);
  static_assert(G == 8 || G == 16 || G == 32, "group width");
  static_assert((KC & (KC - 1)) == 0 && KC % G == 0 && NI % G == 0 && ND % G == 0, "shape must split evenly over the group");
  static_assert(EPT <= 32 && BUF_STRIDE % 4 == 0 && TREE_STRIDE % 4 == 0, "layout");
  static constexpr bool kQuadOK = (EPT % 4 == 0 && KC % QG == 0 && (P - M) % 4 == 0 && TOTAL % 4 == 0 && (EPF % 4 == 0 || PAD == 0));
};

// shared-memory bitonic sort of P keys by the G lanes of a group (all groups of the warp in
// lock step); used only when a ray's coarse depths arrive unsorted
__device__ __noinline__ void sort_smem_grp(float* key, int P, int g, int G) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = g; t < (P >> 1); t += G) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool ascending = ((lo & size) == 0);
        const float a = key[lo], b = key[hi];
        if ((a > b) == ascending && a != b) {
          key[lo] = b;
          key[hi] = a;
        }
      }
    }
  }
  __syncwarp();
}

template <int N>
__device__ __forceinline__ void load_consecutive(float (&dst)[N], const float* src, bool vec_ok) {
  if (N % 4 == 0 && vec_ok) {
#pragma unroll
    for (int q = 0; q < N; q += 4) {
      const float4 p = *reinterpret_cast<const float4*>(src + q);
      dst[q] = p.x; dst[(q + 1) % N] = p.y; dst[(q + 2) % N] = p.z; dst[(q + 3) % N] = p.w;
    }
  } else if (N % 2 == 0 && vec_ok) {
#pragma unroll
    for (int q = 0; q < N; q += 2) {
      const float2 p = *reinterpret_cast<const float2*>(src + q);
      dst[q] = p.x; dst[(q + 1) % N] = p.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) dst[q] = src[q];
  }
}

template <int N>
__device__ __forceinline__ void store_consecutive(float* dst, const float (&src)[N], bool vec_ok) {
  if (N % 4 == 0 && vec_ok) {
#pragma unroll
    for (int q = 0; q < N; q += 4)
      *reinterpret_cast<float4*>(dst + q) = make_float4(src[q], src[(q + 1) % N], src[(q + 2) % N], src[(q + 3) % N]);
  } else if (N % 2 == 0 && vec_ok) {
#pragma unroll
    for (int q = 0; q < N; q += 2) *reinterpret_cast<float2*>(dst + q) = make_float2(src[q], src[(q + 1) % N]);
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) dst[q] = src[q];
  }
}

// A row of N*G floats split over the G lanes of a group as 16-byte pieces dealt round-robin: lane g takes
// pieces g, g + G, ... — every instruction touches G*16 CONTIGUOUS bytes per ray (one LSU wavefront per 128
// bytes; a lane-blocked split of 32 bytes per lane costs two).  Which new sample a lane draws is free: they
// are sorted afterwards; the exports below use the same split.
template <int N>
__device__ __forceinline__ void load_dealt(float (&dst)[N], const float* row, int g, int G, bool vec_ok) {
  if (N % 4 == 0 && N > 4 && vec_ok) {
#pragma unroll
    for (int q = 0; q < N; q += 4) {
      const float4 p = *reinterpret_cast<const float4*>(row + 4 * g + q * G);
      dst[q] = p.x; dst[(q + 1) % N] = p.y; dst[(q + 2) % N] = p.z; dst[(q + 3) % N] = p.w;
    }
  } else {
    load_consecutive<N>(dst, row + g * N, vec_ok);
  }
}
template <int N>
__device__ __forceinline__ void store_dealt(float* row, const float (&src)[N], int g, int G, bool vec_ok) {
  if (N % 4 == 0 && N > 4 && vec_ok) {
#pragma unroll
    for (int q = 0; q < N; q += 4)
      *reinterpret_cast<float4*>(row + 4 * g + q * G) = make_float4(src[q], src[(q + 1) % N], src[(q + 2) % N], src[(q + 3) % N]);
  } else {
    store_consecutive<N>(row + g * N, src, vec_ok);
  }
}

// The class of a packed ray for the ragged kernels below: the smallest (KC, NI) box that holds
// its kc coarse and n new samples, or -1 (empty ray, or larger than the largest box: those
// rays take importance_reg.cu's per-class kernels).
__host__ __device__ inline int grp_ragged_class(int kc, int n) {
  if (kc < 1 || kc > 256 || n > 128) return -1;
  if (kc <= 32 && n <= 16) return 0;
  if (kc <= 64 && n <= 32) return 1;
  if (kc <= 128 && n <= 64) return 2;
  return 3;
}

// kRagged == false: dense rays of exactly (KC, NI, ND) samples.
// kRagged == true (packed layout): rays of class kClass, i.e. at most KC coarse and NI new
// samples, padded to the box with masked lanes / +inf keys; the launch handles only its class
// (warp-level filter over 32 consecutive rays), so a short ray never pays for the longest.
template <int G, int KC, int NI, int ND, bool kRagged, int kClass>
__global__ void __launch_bounds__(GrpCfg<G, KC, NI, ND>::WARPS * 32, GrpCfg<G, KC, NI, ND>::MIN_BLOCKS)
importance_grp_kernel(const ImportanceRegArgs a) {
  using C = GrpCfg<G, KC, NI, ND>;
  constexpr int RPW = C::RPW, NIL = C::NIL, NDL = C::NDL, EPF = C::EPF, CW = C::CW, EPT = C::EPT, EPC = C::EPC;
  constexpr int DEPTH = C::DEPTH, PAD = C::PAD, kGrpWarps = C::WARPS;
  constexpr bool kQuad = !kRagged && C::kQuadOK;  // layout of the merged keys (sort_net.cuh: merge_net)
  static_assert(!kRagged || ND == 0, "the packed entry point has no depth samples");
  __shared__ __align__(16) float s_tree[kGrpWarps][RPW][C::TREE_STRIDE];
  __shared__ __align__(16) float s_buf[kGrpWarps][RPW][C::BUF_STRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane & (G - 1), sub = lane / G;
  float* tree = s_tree[warp][sub];
  float* buf = s_buf[warp][sub];
  const bool do_sort = (a.z_sorted != nullptr);
  const LaneSigns sg = lane_signs(lane);
  const float inv_kc = 1.0f / (float)KC;  // KC is a power of two: x * inv_kc == x / KC bit for bit

  // tree slots of this lane's cdf entries q = g*CW + i + 1 (breadth-first order, root = 1):
  // slot(q) = 2^(DEPTH-1-tz) + (q >> (tz+1)) with tz = trailing zeros of q.  For i < CW-1 the
  // trailing zeros of q are those of i + 1 (g*CW has more), a compile-time constant, so the slot is
  // linear in g and costs one IMAD where it is used; only the lane's LAST entry (q = (g+1)*CW) needs
  // a run-time count, kept in one register.  q == KC is not part of the tree: it goes to the unused slot 0.
  const int q_last = (g + 1) * CW;
  const int tz_last = __ffs(q_last) - 1;
  const int h_last = (q_last < KC) ? (1 << (DEPTH - 1 - tz_last)) + (q_last >> (tz_last + 1)) : 0;
  auto tree_slot = [&](const int i) -> int {
    if (i == CW - 1) return h_last;
    const int tz = __ffs(i + 1) - 1;     // i is a constant after unrolling
    return (1 << (DEPTH - 1 - tz)) + (g * (CW >> (tz + 1))) + ((i + 1) >> (tz + 1));
  };

  // packed layout: lengths of the coarse and the new-sample streams (the last offsets)
  const int64_t total_c = kRagged ? a.offsets[a.R] : 0, total_f = kRagged ? a.fine_offsets[a.R] : 0;

  // one ray per group: `live` == false marks a group without work of its own (it recomputes a
  // valid ray alongside the others and stores nothing)
  auto process = [&](const int64_t r, const bool live, const int64_t cbase_in, const int kc_in, const int64_t fbase_in,
                     const int n_in) {
    const int64_t bi = a.bound_stride ? r : 0;
    const float near = a.near[bi], far = a.far[bi];
    const float span = __fsub_rn(far, near);
    // ragged: this ray's counts and its slices of the packed streams
    int kc = KC, n = NI;
    int64_t cbase = r * KC, fbase = r * NI;
    if (kRagged) {  // handed over by the lane that classified the ray: no second trip to the offsets
      cbase = cbase_in;
      kc = kc_in;
      fbase = fbase_in;
      n = n_in;
    }

    // ---- loads: everything this lane needs of its ray --------------------------------------
    float w[CW], uu[NIL > 0 ? NIL : 1], jj[NIL > 0 ? NIL : 1], nn[NDL > 0 ? NDL : 1];
    float x[EPT];
    if (kRagged && __all_sync(0xffffffffu, cbase + KC <= total_c && fbase + NI <= total_f)) {
      // the whole class box lies inside the packed streams (every ray but the last few of a launch): plain loads
      // at constant offsets from one pointer per stream.  What they fetch past the ray's own counts belongs to
      // the following rays and is masked where it is used (selects: a NaN there goes nowhere).
      const float* wrow = a.weights + cbase + g * CW;
#pragma unroll
      for (int i = 0; i < CW; ++i) w[i] = wrow[i];
      // draws striped over the group (lane g takes samples g, g + G, ...: G consecutive floats per instruction;
      // which lane draws which sample is free, they are sorted afterwards)
      const float* urow = a.u + fbase + g;
      const float* jrow = a.u2 + fbase + g;
#pragma unroll
      for (int q = 0; q < NIL; ++q) {
        uu[q] = urow[q * G];
        jj[q] = jrow[q * G];
      }
      if (do_sort) {
        const float* zrow = a.z_coarse + cbase + g;
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const float zv = zrow[i * G];
          x[i] = i * G + g < kc ? zv : CUDART_INF_F;
        }
      }
    } else if (kRagged) {
      // clamped unconditional loads; entries past the ray's counts are masked where they are used
      const float* wrow = a.weights + cbase;
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        const int j = g * CW + i;
        w[i] = wrow[j < kc ? j : kc - 1];
      }
#pragma unroll
      for (int q = 0; q < NIL; ++q) {
        const int e = q * G + g;
        const int ec = n > 0 ? (e < n ? e : n - 1) : 0;
        uu[q] = n > 0 ? a.u[fbase + ec] : 0.f;
        jj[q] = n > 0 ? a.u2[fbase + ec] : 0.f;
      }
      if (do_sort) {
        const float* zrow = a.z_coarse + cbase;
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const int j = i * G + g;
          const float zv = zrow[j < kc ? j : kc - 1];
          x[i] = j < kc ? zv : CUDART_INF_F;
        }
      }
    } else {
      load_consecutive<CW>(w, a.weights + r * KC + g * CW, a.vecw != 0);
      if (NIL > 0) {
        load_dealt<(NIL > 0 ? NIL : 1)>(uu, a.u + r * NI, g, G, a.vec4 != 0);
        load_dealt<(NIL > 0 ? NIL : 1)>(jj, a.u2 + r * NI, g, G, a.vec4 != 0);
      }
      if (do_sort) {
        if (NDL > 0) load_consecutive<(NDL > 0 ? NDL : 1)>(nn, a.normals + r * ND + g * NDL, a.vecz != 0);
        const float* zrow = a.z_coarse + r * KC;  // quads: merged position q = (i >> 2)*4G + 4g + (i & 3)
        if (!kQuad) {
#pragma unroll
          for (int i = 0; i < CW; ++i) x[i] = zrow[i * G + g];  // striped: merged position q = i*G + g
        } else if (a.vecz) {
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            const float4 p = *reinterpret_cast<const float4*>(zrow + (i >> 2) * C::QG + 4 * g);
            x[i] = p.x; x[i + 1] = p.y; x[i + 2] = p.z; x[i + 3] = p.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < CW; ++i) x[i] = zrow[(i >> 2) * C::QG + 4 * g + (i & 3)];
        }
      }
    }

    // ---- 1. cdf (renderers.py:36-39): blocked scan + running max (the search needs a
    //         non-decreasing table; a parallel prefix sum is not monotone in floating point)
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      w[i] = (!kRagged || g * CW + i < kc) ? __fadd_rn(w[i], kPdfEps) : 0.f;
      part += w[i];
    }
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    const float S = part;
    // pdf = w / S (renderers.py:37) as w * (1/S) with a correctly rounded reciprocal: <= 1 ulp per
    // entry, far inside the cdf's tolerance, and the search stays bit-exact with respect to the cdf
    // the kernel exports (a true division is 16 instructions per entry on this ISA)
    const float rS = __fdiv_rn(1.0f, S);
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      run += __fmul_rn(w[i], rS);
      w[i] = run;  // local inclusive prefix
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
    __syncwarp();  // the previous iteration's readers of tree/buf are done
    float* cdf_out = (!kRagged && a.cdf && live) ? a.cdf + r * (KC + 1) : nullptr;
    float last = 0.f;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      float val = fmaxf(off + w[i], floor_prev);
      if (kRagged && g * CW + i >= kc) val = CUDART_INF_F;  // past the ray's last bin
      tree[tree_slot(i)] = val;
      if (cdf_out) cdf_out[g * CW + i + 1] = val;
      last = val;
    }
    if (cdf_out && g == 0) cdf_out[0] = 0.f;
    last = __shfl_sync(0xffffffffu, last, lane | (G - 1));  // cdf[KC]
    __syncwarp();

    // ---- 2. this lane's new samples -----------------------------------------------------------
    float v[EPF];
    if (NIL > 0) {
      int32_t* irow = (!kRagged && a.idx && live) ? a.idx + r * NI : nullptr;           // dense: the row, dealt as u was
      float* frow = (a.z_fine && live) ? a.z_fine + fbase + (kRagged ? g : 0) : nullptr;
      const float kcf = (float)kc;
      // ragged rays divide by their own count (renderers.py:45).  1 <= kc <= 256, so y = RN(1/kc) and Markstein's
      // sequence q0 = RN(a*y); r = fma(-kc, q0, a); q = fma(r, y, q0) give the correctly rounded a/kc for every
      // numerator whose intermediates stay normal (coarse_packed_core.h has the same shortcut and its CPU walk);
      // anything else — no draw of torch.rand gets there — divides the IEEE way.
      const float rkc = kRagged ? __frcp_rn(kcf) : 0.f;
      // node <- 2*node + (tree[node] <= u): after DEPTH probes node - KC counts the entries
      // cdf[1..KC-1] <= u; adding (cdf[KC] <= u) gives clamp_min(searchsorted(cdf, u, right=True) - 1, 0).
      // (c <= u) is the complement of the sign bit of u - c (exact in sign: nothing is flushed).
      // Held as shared-memory byte addresses a = base + 4 * node, so a probe step is LDS, one 3-input add
      // (a <- 2a - base) and a predicated +4: four instructions.
      unsigned node[NIL > 0 ? NIL : 1];
      int bins[NIL > 0 ? NIL : 1];
      const unsigned tbase = smem_u32(tree);
      const unsigned neg_base = 0u - tbase;
      // With four or more draws per lane the first two probe levels come from registers: the root and its two
      // children are read once per ray (three broadcast reads) instead of twice per draw — the kernel is bound by
      // LSU wavefronts, and a select plus a compare is cheaper than a shared-memory round trip.
      constexpr int kRegLevels = (NIL >= 4 && DEPTH >= 3) ? 2 : 0;
      if (kRegLevels == 2) {
        const float n1 = tree[1], n2 = tree[2], n3 = tree[3];
#pragma unroll
        for (int q = 0; q < NIL; ++q) {
          const bool right = n1 <= uu[q];
          const float c = right ? n3 : n2;
          unsigned nd = right ? tbase + 24u : tbase + 16u;   // node 4 + 2*right ...
          if (c <= uu[q]) nd += 4u;                           // ... + (child <= u)
          node[q] = nd;
        }
      } else {
#pragma unroll
        for (int q = 0; q < NIL; ++q) node[q] = tbase + 4u;
      }
#pragma unroll
      for (int step = kRegLevels; step < DEPTH; ++step) {
#pragma unroll
        for (int q = 0; q < NIL; ++q) {
          float cv;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cv) : "r"(node[q]));
          node[q] = node[q] + node[q] + neg_base;
          if (cv <= uu[q]) node[q] += 4u;
        }
      }
#pragma unroll
      for (int q = 0; q < NIL; ++q) node[q] = (node[q] - tbase) >> 2;
#pragma unroll
      for (int q = 0; q < NIL; ++q) {
        const float dl = __fsub_rn(uu[q], last);
        const int bin = (int)(node[q] - KC) + 1 - (int)(__float_as_uint(dl) >> 31);
        const float num = __fadd_rn((float)bin, jj[q]);                    // renderers.py:45
        float t;
        if (kRagged) {
          const float an = fabsf(num);
          if ((an >= 0x1p-90f && an < 0x1p100f) || an == 0.f) {
            const float q0 = __fmul_rn(num, rkc);
            t = __fmaf_rn(__fmaf_rn(-kcf, q0, num), rkc, q0);
          } else {
            t = __fdiv_rn(num, kcf);
          }
        } else {
          t = __fmul_rn(num, inv_kc);
        }
        const float val = __fadd_rn(near, __fmul_rn(span, t));              // :46
        bins[q] = bin;
        v[q] = val;
      }
      if (kRagged) {
#pragma unroll
        for (int q = 0; q < NIL; ++q) {
          const bool has = q * G + g < n;
          if (frow && has) frow[q * G] = v[q];
          if (!has) v[q] = CUDART_INF_F;  // padding of the class box
        }
      } else {
        if (irow) store_dealt<(NIL > 0 ? NIL : 1)>(reinterpret_cast<float*>(irow), reinterpret_cast<const float(&)[NIL > 0 ? NIL : 1]>(bins), g, G, a.vec4 != 0);
        if (frow) store_dealt<(NIL > 0 ? NIL : 1)>(frow, reinterpret_cast<const float(&)[NIL > 0 ? NIL : 1]>(v), g, G, a.vec4 != 0);
      }
    }
    if (!do_sort) return;
#pragma unroll
    for (int q = 0; q < NDL; ++q) {
      // sample_depth's randn*std (the depth is NOT added), clamped (renderers.py:62-66, :255)
      v[NIL + q] = fminf(fmaxf(__fmul_rn(nn[q], a.depth_std), near), far);
    }
#pragma unroll
    for (int q = NIL + NDL; q < EPF; ++q) v[q] = CUDART_INF_F;

    // ---- 3. sort the group's G*EPF new samples in registers -------------------------------------
    sort_blocked<EPF, G>(v, sg);

    // ---- 4. [coarse ascending | +inf | new descending], striped over the group --------------------
    {
      float* chunk = buf + (G - 1 - g) * (EPF + PAD);  // descending: this lane's keys go to the mirrored chunk
      if (EPF % 4 == 0) {
#pragma unroll
        for (int o = 0; o < EPF; o += 4)
          *reinterpret_cast<float4*>(chunk + o) =
              make_float4(v[EPF - 1 - o], v[(EPF - 2 - o + EPF) % EPF], v[(EPF - 3 - o + EPF) % EPF], v[(EPF - 4 - o + EPF) % EPF]);
      } else {
#pragma unroll
        for (int o = 0; o < EPF; ++o) chunk[o] = v[EPF - 1 - o];
      }
    }
    __syncwarp();
    // key index of register i: quads (dense rays) or striped (packed rays, whose rows have no alignment)
    auto key_index = [&](const int i) -> int { return kQuad ? (i >> 2) * C::QG + 4 * g + (i & 3) : i * G + g; };
    if (kQuad) {
      // groups of four registers: wholly coarse (loaded above), wholly padding, or — from key P - M on — the
      // new samples in descending order, read back 16 bytes at a time
#pragma unroll
      for (int i = CW; i < EPT; i += 4) {
        const int q0 = (i >> 2) * C::QG + 4 * g;
        float4 p = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
        if (!((C::kInfQ >> i) & 1u) && q0 >= C::P - C::M) {
          const int idx = q0 - (C::P - C::M);
          p = *reinterpret_cast<const float4*>(buf + idx + (idx / EPF) * PAD);
        }
        x[i] = p.x; x[i + 1] = p.y; x[i + 2] = p.z; x[i + 3] = p.w;
      }
    } else {
#pragma unroll
      for (int i = CW; i < EPT; ++i) {
        if (i < EPC) {
          x[i] = CUDART_INF_F;
        } else {
          const int idx = (i - EPC) * G + g;          // position in the descending new-sample sequence
          x[i] = buf[idx + (idx / EPF) * PAD];
        }
      }
    }
    // coarse depths must be ascending for the merge (they are when they come from sample_coarse)
    bool unsorted = false;
    if (kQuad) {
#pragma unroll
      for (int i = 0; i < CW; i += 4) {
        // inside the lane's four, then its last against the next lane's first (the group's last lane: against
        // lane 0's first of the next group of registers)
        unsorted = unsorted || x[i] > x[i + 1] || x[i + 1] > x[i + 2] || x[i + 2] > x[i + 3];
        float nx = __shfl_down_sync(0xffffffffu, x[i], 1);
        if (i + 4 < CW) {
          const float first_next = __shfl_sync(0xffffffffu, x[(i + 4 < CW) ? i + 4 : i], lane & ~(G - 1));
          if (g == G - 1) nx = first_next;
        } else if (g == G - 1) {
          nx = CUDART_INF_F;
        }
        unsorted = unsorted || x[i + 3] > nx;
      }
    } else {
      // key i*G + g + 1 sits in register i of the next lane — for the group's last lane in register i + 1 of its
      // first: that lane offers x[i + 1] instead, so one rotating shuffle per register serves both
      const int next_lane = (lane & ~(G - 1)) | ((g + 1) & (G - 1));
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        const float offer = (g == 0 && i + 1 < CW) ? x[(i + 1 < CW) ? i + 1 : i] : x[i];
        float nx = __shfl_sync(0xffffffffu, offer, next_lane);
        if (i + 1 == CW && g == G - 1) nx = CUDART_INF_F;
        if ((!kRagged || i * G + g + 1 < kc) && x[i] > nx) unsorted = true;
      }
    }
    if (__any_sync(0xffffffffu, unsorted)) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < EPT; ++i) buf[key_index(i)] = x[i];
      sort_smem_grp(buf, C::P, g, G);
#pragma unroll
      for (int i = 0; i < EPT; ++i) x[i] = buf[key_index(i)];
    } else {
      merge_net<EPT, (kQuad ? C::kInfQ : C::kInf0), G, kQuad>(x, sg);
    }

    // ---- 5. store the first TOTAL keys: G*16 contiguous bytes per ray and instruction (quads), 4G (striped)
    if (live) {
      float* out = a.z_sorted + (kRagged ? cbase + fbase : r * C::TOTAL);
      const int total = kRagged ? kc + n : C::TOTAL;
      if (kQuad && a.vecz) {
#pragma unroll
        for (int i = 0; i < EPT; i += 4) {
          const int q0 = (i >> 2) * C::QG + 4 * g;
          if (q0 < total) *reinterpret_cast<float4*>(out + q0) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          const int q = key_index(i);
          if (!kQuad && i * G >= C::TOTAL) continue;  // striped: registers wholly past the largest row
          if (q < total) out[q] = x[i];
        }
      }
    }
  };

  if (kRagged) {
    // 32 consecutive rays per step: each lane classifies one; the groups take RPW matches at a time
    const int64_t step = (int64_t)gridDim.x * kGrpWarps * 32;
    for (int64_t base = (blockIdx.x * (int64_t)kGrpWarps + warp) * 32; base < a.R; base += step) {
      const int64_t rr = base + lane;
      bool mine = false;
      int64_t c_lane = 0, f_lane = 0;
      int kc_lane = 0, n_lane = 0;
      if (rr < a.R) {
        c_lane = a.offsets[rr];
        f_lane = a.fine_offsets[rr];
        kc_lane = (int)(a.offsets[rr + 1] - c_lane);
        n_lane = (int)(a.fine_offsets[rr + 1] - f_lane);
        mine = grp_ragged_class(kc_lane, n_lane) == kClass;
      }
      unsigned todo = __ballot_sync(0xffffffffu, mine);
      while (todo) {
        unsigned mine_on = todo;                              // the (sub+1)-th match, if there is one
#pragma unroll
        for (int k = 1; k < RPW; ++k)
          if (k <= sub) mine_on &= mine_on - 1;
        const bool live = mine_on != 0u;
        const int src = __ffs(live ? mine_on : todo) - 1;
#pragma unroll
        for (int k = 0; k < RPW; ++k) todo &= todo - 1;       // (x & (x-1)) of 0 stays 0
        process(base + src, live, __shfl_sync(0xffffffffu, c_lane, src), __shfl_sync(0xffffffffu, kc_lane, src),
                __shfl_sync(0xffffffffu, f_lane, src), __shfl_sync(0xffffffffu, n_lane, src));
      }
    }
  } else {
    const int64_t n_wg = (a.R + RPW - 1) / RPW;
    const int64_t wg_stride = (int64_t)gridDim.x * kGrpWarps;
    for (int64_t wg = blockIdx.x * (int64_t)kGrpWarps + warp; wg < n_wg; wg += wg_stride) {
      const int64_t r_raw = wg * RPW + sub;
      process(r_raw < a.R ? r_raw : a.R - 1, r_raw < a.R, 0, 0, 0, 0);   // a dead group (last warp only) recomputes ray R-1
    }
  }
}

template <int G, int KC, int NI, int ND>
static int launch_grp(const ImportanceRegArgs& a, cudaStream_t stream) {
  using C = GrpCfg<G, KC, NI, ND>;
  constexpr int RPW = 32 / G;
  const int64_t n_wg = (a.R + RPW - 1) / RPW;
  int64_t blocks = (n_wg + C::WARPS - 1) / C::WARPS;
  const int64_t cap = (int64_t)num_sms() * C::MIN_BLOCKS * 4;  // a few waves for balance
  if (blocks > cap) blocks = cap;
  importance_grp_kernel<G, KC, NI, ND, false, 0><<<(unsigned)blocks, C::WARPS * 32, 0, stream>>>(a);
  return check_launch();
}

template <int G, int KC, int NI, int kClass>
static int launch_grp_ragged(const ImportanceRegArgs& a, cudaStream_t stream) {
  using C = GrpCfg<G, KC, NI, 0>;
  int64_t blocks = (a.R + C::WARPS * 32 - 1) / (C::WARPS * 32);
  const int64_t cap = (int64_t)num_sms() * C::MIN_BLOCKS * 2;
  if (blocks > cap) blocks = cap;
  importance_grp_kernel<G, KC, NI, 0, true, kClass><<<(unsigned)blocks, C::WARPS * 32, 0, stream>>>(a);
  return check_launch();
}

// Packed layout: one launch per ray class that the caller's maxima allow.  Returns AVR_OK after
// launching; rays outside every class (kc > 256 or n > 128) are left to the caller
// (importance_reg.cu), which is told by `*covers_all`.
int launch_importance_grp_ragged(const ImportanceRegArgs& a, int max_coarse, int max_fine, bool* covers_all,
                                 cudaStream_t stream) {
  *covers_all = (max_coarse <= 256 && max_fine <= 128);
  int rc = launch_grp_ragged<8, 32, 16, 0>(a, stream);
  if (rc == AVR_OK && (max_coarse > 32 || max_fine > 16)) rc = launch_grp_ragged<8, 64, 32, 1>(a, stream);
  if (rc == AVR_OK && (max_coarse > 64 || max_fine > 32)) rc = launch_grp_ragged<16, 128, 64, 2>(a, stream);
  // the largest box runs one ray per warp: 64 registers / 32 resident warps beat 16 lanes per ray at
  // 128 registers / 16 warps (1.44 vs 1.56 ms on BASELINE.json config 4's distribution)
  if (rc == AVR_OK && (max_coarse > 128 || max_fine > 64)) rc = launch_grp_ragged<32, 256, 128, 3>(a, stream);
  return rc;
}

// Hot dense shapes, compiled with the shape as constants: BASELINE.json config 3 (64 -> 128),
// conf/default.conf's renderer (64 coarse, 16 importance + 16 depth, renderers.py:252-258) and
// VolumeRenderer.from_conf's defaults (32 coarse, 8 + 8, renderers.py:279-289).
int launch_importance_grp(const ImportanceRegArgs& a, cudaStream_t stream) {
  if (a.offsets) return AVR_ERR_UNSUPPORTED;
  // lanes per ray: 8 for every shape.  The kernel is bound by LSU wavefronts (shuffles, shared memory and
  // global accesses share one data pipe per SM), and 8 lanes per ray need 84 shuffles per two rays of the
  // 128-sample shape where 16 lanes need 128; that outweighs 127 registers / 16 resident warps against 64 / 32
  // (0.653 vs 0.695 ms for 2^20 rays on B200).  AVR_GRP_G=8|16 overrides (experiments).
  const int force_g = option(OPT_GRP_G, 0);
  if (force_g == 32 && a.Kc == 64 && a.n_imp == 128 && a.n_depth == 0) return launch_grp<32, 64, 128, 0>(a, stream);
#define AVR_GRP_CASE(G_, KC_, NI_, ND_)                                        \
  if (a.Kc == KC_ && a.n_imp == NI_ && a.n_depth == ND_) {                     \
    if (force_g == 8) return launch_grp<8, KC_, NI_, ND_>(a, stream);          \
    if (force_g == 16) return launch_grp<16, KC_, NI_, ND_>(a, stream);        \
    return launch_grp<G_, KC_, NI_, ND_>(a, stream);                           \
  }
  AVR_GRP_CASE(8, 64, 128, 0)
  AVR_GRP_CASE(8, 64, 16, 16)
  AVR_GRP_CASE(8, 64, 16, 0)
  AVR_GRP_CASE(8, 64, 64, 0)
  // VolumeRenderer.from_conf's defaults (renderers.py:279-289): 8 + 8 new samples only split over 8 lanes
  if (a.Kc == 32 && a.n_imp == 8 && a.n_depth == 8) return launch_grp<8, 32, 8, 8>(a, stream);
  if (a.Kc == 32 && a.n_imp == 8 && a.n_depth == 0) return launch_grp<8, 32, 8, 0>(a, stream);
#undef AVR_GRP_CASE
  return AVR_ERR_UNSUPPORTED;
}

}  // namespace avr
