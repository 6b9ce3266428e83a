// Register-resident importance sampler + merge (the fast path of avr_importance_sample;
// csrc/samplers.cu keeps the general shared-memory kernel).
//
// One warp per ray, no block-level synchronisation:
//   1. pdf/cdf: sum and inclusive scan of (w + 1e-5)/S by warp shuffles (renderers.py:36-39);
//      the table goes to shared memory as an implicit binary search tree padded with +inf to
//      a power of two, so the search is a fixed number of branch-free steps.
//   2. each lane owns EPF *consecutive* new samples (importance samples, then the clamped
//      "depth" samples, then +inf padding): inverse-CDF search (renderers.py:41-43), jitter
//      and placement (:45-46) with the reference's rounding.
//   3. the 32*EPF new samples are sorted IN REGISTERS by a bitonic network (lane-local
//      compare-exchanges for small strides, one shuffle + one min per element for the rest).
//   4. coarse depths are already ascending (stratified sampling), so
//      [coarse ascending | +inf | new samples descending] is a bitonic sequence of
//      P = 32*EPT elements: ONE bitonic merge (log2 P steps) in a lane-striped register
//      layout finishes the sort (renderers.py:257-258), and the striped layout stores
//      straight to global memory fully coalesced.  If a ray's coarse depths are not
//      ascending (arbitrary caller input), that ray takes a full shared-memory sort instead.
//
// The kernel is bound by the SM's ALU pipe (min/max, compares, integer logic issue at half
// the FP32 rate on sm_100), so the hot shapes are compiled with the shape as template
// constants (no per-ray bounds logic), and work is moved to the FMA pipe where possible:
//   * cross-lane compare-exchanges run in a "signed domain": a lane that must keep the
//     maximum holds its keys negated, so BOTH partners execute the same min(w, -other)
//     (one FMNMX instead of a min/max pair); the sign pattern changes between stages by
//     one multiply by +-1 (FMA pipe, exact);
//   * the search step's compare is the sign bit of u - cdf[node] (an FADD);
//   * merge comparators whose inputs are known +inf padding are dropped at compile time.
#include <math_constants.h>

#include <cstdlib>

#include "avr_common.cuh"
#include "kernels.h"
#include "importance_args.cuh"
#include "sort_net.cuh"

namespace avr {

constexpr int kRegWarps = 8;  // warps per CTA


// smem bitonic sort (any content), P power of two; used for the rare unsorted-coarse ray
__device__ __noinline__ void sort_smem(float* key, int P, int lane) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = lane; t < (P >> 1); t += 32) {
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

// One ray's inputs, held in registers so the NEXT ray's loads are in flight while the
// current ray is being processed (the kernel is otherwise bound by global-load latency:
// weights -> sum -> cdf -> search is one dependent chain per ray).
template <int EPF, int EPC>
struct RayIn {
  float w[EPC];    // coarse weights, blocked: j = lane*ceil(Kc/32) + i
  float zc[EPC];   // coarse depths, striped
  float a[EPF];    // lane's consecutive new samples: u (e < n) or N(0,1) draw (n <= e < n+nd)
  float b[EPF];    // in-bin jitter u2 (e < n)
  float near, far;
  int kc, n;            // this ray's coarse / importance sample counts
  int64_t cbase, fbase; // first coarse / importance sample of the ray in the streams
};

// KC > 0: the dense shape (KC coarse, NI importance, ND depth samples) is a compile-time constant
// with KC % 32 == 0; KC == 0: shapes are read at run time (packed layout, unusual dense shapes).
template <int EPF, int EPC, int KC, int NI, int ND>
__device__ __forceinline__ void load_ray(RayIn<EPF, EPC>& in, const ImportanceRegArgs& a, int64_t r, int lane,
                                         bool do_sort) {
  constexpr bool kStatic = KC > 0;
  const int nd = kStatic ? ND : a.n_depth;
  if (kStatic) {
    in.kc = KC;
    in.n = NI;
    in.cbase = r * (int64_t)KC;
    in.fbase = r * (int64_t)NI;
  } else if (a.offsets) {
    in.cbase = a.offsets[r];
    in.kc = (int)(a.offsets[r + 1] - in.cbase);
    in.fbase = a.fine_offsets[r];
    in.n = (int)(a.fine_offsets[r + 1] - in.fbase);
  } else {
    in.kc = a.Kc;
    in.n = a.n_imp;
    in.cbase = r * (int64_t)a.Kc;
    in.fbase = r * (int64_t)a.n_imp;
  }
  const int kc = in.kc, n = in.n;
  const int64_t bi = a.bound_stride ? r : 0;
  in.near = a.near[bi];
  in.far = a.far[bi];
  const float* wrow = a.weights + in.cbase;
  if (kStatic) {
    constexpr int CW = KC / 32;
    if (CW == 2 && a.vecw) {
      const float2 p = *reinterpret_cast<const float2*>(wrow + 2 * lane);
      in.w[0] = p.x;
      in.w[CW > 1 ? 1 : 0] = p.y;
    } else if (CW == 4 && a.vecw) {
      const float4 p = *reinterpret_cast<const float4*>(wrow + 4 * lane);
      in.w[0] = p.x; in.w[CW > 1 ? 1 : 0] = p.y; in.w[CW > 2 ? 2 : 0] = p.z; in.w[CW > 3 ? 3 : 0] = p.w;
    } else {
#pragma unroll
      for (int i = 0; i < CW; ++i) in.w[i] = wrow[lane * CW + i];
    }
    if (do_sort) {
      const float* zrow = a.z_coarse + in.cbase;
#pragma unroll
      for (int i = 0; i < CW; ++i) in.zc[i] = zrow[i * 32 + lane];
    }
  } else if (kc > 0) {
    // run-time shapes use a FIXED layout per kernel variant (EPC bins per lane, whatever the
    // ray's count): clamped unconditional loads + selects instead of guarded loads, and every
    // index that depends only on the lane is hoisted out of the ray loop
    const int last = kc - 1;
#pragma unroll
    for (int i = 0; i < EPC; ++i) {
      const int j = lane * EPC + i;
      in.w[i] = wrow[j < last ? j : last];  // entries with j >= kc are masked where they are used
    }
    if (do_sort) {
      const float* zrow = a.z_coarse + in.cbase;
#pragma unroll
      for (int i = 0; i < EPC; ++i) {
        const int j = i * 32 + lane;
        const float zv = zrow[j < last ? j : last];
        in.zc[i] = (j < kc) ? zv : CUDART_INF_F;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < EPC; ++i) {
      in.w[i] = 0.f;
      in.zc[i] = CUDART_INF_F;
    }
  }
  const int e0 = lane * EPF;
  const float* urow = a.u + in.fbase;
  const float* u2row = a.u2 + in.fbase;
  if (kStatic) {
    if (EPF >= 4 && a.vec4 && NI == 32 * EPF) {  // 16-byte loads of the lane's consecutive draws
#pragma unroll
      for (int q = 0; q < EPF; q += 4) {
        const float4 p4 = *reinterpret_cast<const float4*>(urow + e0 + q);
        const float4 j4 = *reinterpret_cast<const float4*>(u2row + e0 + q);
        in.a[q] = p4.x; in.a[(q + 1) % EPF] = p4.y; in.a[(q + 2) % EPF] = p4.z; in.a[(q + 3) % EPF] = p4.w;
        in.b[q] = j4.x; in.b[(q + 1) % EPF] = j4.y; in.b[(q + 2) % EPF] = j4.z; in.b[(q + 3) % EPF] = j4.w;
      }
    } else {
      const float* nrow = a.normals ? a.normals + r * (int64_t)nd : nullptr;
#pragma unroll
      for (int q = 0; q < EPF; ++q) {
        const int e = e0 + q;
        float va = 0.f, vb = 0.f;
        if (e < n) {
          va = urow[e];
          vb = u2row[e];
        } else if (e < n + nd && do_sort) {
          va = nrow[e - n];
        }
        in.a[q] = va;
        in.b[q] = vb;
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < EPF; ++q) {
      in.a[q] = 0.f;
      in.b[q] = 0.f;
    }
    if (n > 0) {
#pragma unroll
      for (int q = 0; q < EPF; ++q) {
        const int e = e0 + q;
        const int ec = e < n - 1 ? e : n - 1;
        in.a[q] = urow[ec];
        in.b[q] = u2row[ec];
      }
    }
    if (nd > 0 && do_sort) {  // dense only: the packed entry point has no depth samples
      const float* nrow = a.normals + r * (int64_t)nd;
#pragma unroll
      for (int q = 0; q < EPF; ++q) {
        const int e = e0 + q - n;
        const float nv = nrow[e < 0 ? 0 : (e < nd - 1 ? e : nd - 1)];
        if (e >= 0) in.a[q] = nv;  // only read where n <= e0 + q < n + nd
      }
    }
  }
}

// Everything that happens to one ray once its inputs are in registers (one warp, all lanes).
template <int EPF, int EPT, int KC, int NI, int ND>
__device__ __forceinline__ void process_ray(const ImportanceRegArgs& a, const RayIn<EPF, EPT - EPF>& cur, int64_t r,
                                            float* cdf, float* buf, int lane, const LaneSigns& sg, bool& first_ray) {
  constexpr bool kStatic = KC > 0;
  constexpr int P = 32 * EPT;    // merged length incl. padding; also the padded cdf table length
  constexpr int EPC = EPT - EPF; // coarse samples per lane (striped), Kc <= 32*EPC
  constexpr int CW = kStatic ? KC / 32 : EPC;  // registers that can hold real coarse samples
  // merge registers known to be +inf padding: coarse registers beyond KC (static shapes only)
  constexpr unsigned kInf0 = kStatic ? (((1u << EPC) - 1u) & ~((1u << CW) - 1u)) : 0u;
  const int nd = kStatic ? ND : a.n_depth;
  const bool do_sort = (a.z_sorted != nullptr);
  const int e0 = lane * EPF;
  const float near = cur.near, far = cur.far;
  const float span = __fsub_rn(far, near);
  const int kc = kStatic ? KC : cur.kc, n = kStatic ? NI : cur.n;
  const int total = kc + n + nd;
  const float kcf = (float)kc;
  const bool kc_pow2 = (kc & (kc - 1)) == 0;
  const float inv_kc = 1.0f / kcf;  // exact when kc is a power of two: x*inv_kc == x/kc bit for bit

  // ---- 1. cdf table -----------------------------------------------------------------
  // blocked scan: each lane owns c consecutive bins; local running sums, one exclusive
  // warp scan of the lane totals, then a running max (a parallel prefix sum is not
  // monotone in floating point; the search needs a non-decreasing table)
  constexpr int c = CW;  // bins per lane: KC/32 (static shapes) or the variant's EPC (run-time shapes)
  auto valid = [&](int i) { return kStatic || (lane * c + i < kc); };
  float wp[CW];
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    wp[i] = valid(i) ? __fadd_rn(cur.w[i], kPdfEps) : 0.f;
    part += wp[i];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  const float S = part;
  // static shapes divide like the reference (renderers.py:37); run-time shapes (a lane may own
  // up to 12 bins) multiply by 1/S: <= 1 ulp per pdf entry, far inside the cdf's tolerance, and
  // the search stays bit-exact with respect to the cdf the kernel exports
  const float rS = kStatic ? 0.f : __fdiv_rn(1.0f, S);
  float ps[CW];
  float run = 0.f;
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    run += kStatic ? __fdiv_rn(wp[i], S) : __fmul_rn(wp[i], rS);
    ps[i] = run;
  }
  float incl = run;  // inclusive scan of the lane totals
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float p = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += p;
  }
  float off = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) off = 0.f;
  float mx = off + run;  // this lane's last (largest) entry
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float p = __shfl_up_sync(0xffffffffu, mx, d);
    if (lane >= d) mx = fmaxf(mx, p);
  }
  float floor_prev = __shfl_up_sync(0xffffffffu, mx, 1);
  if (lane == 0) floor_prev = 0.f;
  // The table is stored as an implicit binary search tree in breadth-first order (heap
  // index 1 = root): the nodes one search step can touch are contiguous, so the 32 lanes'
  // probes fall into distinct banks (a sorted table probed at power-of-two strides is a
  // worst case for bank conflicts).  Sorted position q in [1, Pc) holds cdf[q], +inf
  // beyond Kc; cdf[0] = 0 is implicit.  The depth is a constant of the variant.
  constexpr int m = kStatic ? ilog2_c(KC) + 1 : ilog2_c(P);
  constexpr int Pc = 1 << m;
  auto heap_index = [](int q) {
    const int tz = __ffs(q) - 1;
    return (1 << (m - 1 - tz)) + (q >> (tz + 1));
  };
  __syncwarp();  // previous ray's readers of cdf/buf are done
  float* cdf_out = a.cdf ? a.cdf + (cur.cbase + r) : nullptr;
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    const int j = lane * c + i;
    const float val = fmaxf(off + ps[i], floor_prev);
    cdf[heap_index(j + 1)] = valid(i) ? val : CUDART_INF_F;
    if (cdf_out && valid(i)) cdf_out[j + 1] = val;
  }
  if (cdf_out && lane == 0) cdf_out[0] = 0.f;
  if (first_ray) {  // slots no lane owns: +inf once
    for (int q = 32 * c + 1 + lane; q < Pc; q += 32) cdf[heap_index(q)] = CUDART_INF_F;
    first_ray = false;
  }
  __syncwarp();

  // ---- 2. the lane's EPF consecutive new samples --------------------------------------
  float v[EPF];
  int32_t* irow = a.idx ? a.idx + cur.fbase : nullptr;
  float* frow = a.z_fine ? a.z_fine + cur.fbase : nullptr;
  // descend the tree: node <- 2*node + (tree[node] <= u); after m steps node - Pc is the
  // number of entries cdf[q >= 1] <= u, i.e. clamp_min(searchsorted(cdf, u, right=True) - 1, 0).
  // (tree[node] <= u) is the complement of the sign bit of u - tree[node]: the subtraction
  // is exact in sign (no flush to zero), +inf padding gives -inf.
  unsigned node[EPF];
#pragma unroll
  for (int q = 0; q < EPF; ++q) node[q] = 1;
#pragma unroll
  for (int step = 0; step < m; ++step) {
#pragma unroll
    for (int q = 0; q < EPF; ++q) {
      const float d = __fsub_rn(cur.a[q], cdf[node[q]]);
      node[q] = 2 * node[q] + 1 - (__float_as_uint(d) >> 31);
    }
  }
#pragma unroll
  for (int q = 0; q < EPF; ++q) {
    const int e = e0 + q;
    const int bin = (int)node[q] - Pc;
    const float num = __fadd_rn((float)bin, cur.b[q]);
    const float t = kc_pow2 ? __fmul_rn(num, inv_kc) : __fdiv_rn(num, kcf);
    float val = __fadd_rn(near, __fmul_rn(span, t));
    if (e < n) {
      if (irow) irow[e] = bin;
      if (frow) frow[e] = val;
    } else if (e < n + nd) {
      // sample_depth's randn*std (depth NOT added), clamped (renderers.py:62-66, :255)
      val = fminf(fmaxf(__fmul_rn(cur.a[q], a.depth_std), near), far);
    } else {
      val = CUDART_INF_F;
    }
    v[q] = val;
  }
  if (do_sort) {
    // ---- 3. sort the new samples in registers -----------------------------------------
    sort_blocked<EPF>(v, sg);

    // ---- 4. lay out [coarse | +inf | new descending] striped over the lanes -------------
#pragma unroll
    for (int q = 0; q < EPF; ++q) buf[P - 1 - (e0 + q)] = v[q];
    __syncwarp();
    float x[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if (i < CW) x[i] = cur.zc[i];                 // runtime shapes: +inf beyond kc
      else if (i < EPC) x[i] = CUDART_INF_F;        // static shapes: pure padding
      else x[i] = buf[i * 32 + lane];               // positions >= P - M
    }
    // coarse depths must be ascending for the merge; check the (q, q+1) pairs inside [0, kc)
    bool unsorted = false;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      float nx = __shfl_down_sync(0xffffffffu, x[i], 1);
      const float first_next = __shfl_sync(0xffffffffu, x[(i + 1 < CW) ? i + 1 : i], 0);
      if (lane == 31) nx = (i + 1 < CW) ? first_next : CUDART_INF_F;
      if ((kStatic || i * 32 + lane + 1 < kc) && x[i] > nx) unsorted = true;
    }
    if (__any_sync(0xffffffffu, unsorted)) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < EPT; ++i) buf[i * 32 + lane] = x[i];
      sort_smem(buf, P, lane);
#pragma unroll
      for (int i = 0; i < EPT; ++i) x[i] = buf[i * 32 + lane];
    } else {
      merge_striped<EPT, kInf0>(x, sg);
    }

    // ---- 5. coalesced store of the first `total` keys ------------------------------------
    float* out = a.z_sorted + (cur.cbase + cur.fbase + r * (int64_t)nd);
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const int q = i * 32 + lane;
      if (q < total) out[q] = x[i];
    }
  }
}

// The (EPF, EPT) variant a ray of kc coarse and m new samples needs: the smallest register
// footprint that holds it (same rule as the host-side choice for dense shapes).
__host__ __device__ inline void ray_class(int kc, int m, int& epf, int& ept) {
  epf = 1;
  while (32 * epf < m) epf <<= 1;
  ept = 2 * epf;
  while (32 * ept < kc + 32 * epf || 32 * ept < kc + 2) ept <<= 1;
}

// kFilter == false: every ray (dense layout, or packed rays all treated at the maximum shape).
// kFilter == true (packed layout): the launch handles only the rays whose class is exactly
// (EPF, EPT); the host launches one such kernel per class, so a short ray never pays for the
// longest one and each class gets its own register budget.
template <int EPF, int EPT, int KC, int NI, int ND, bool kFilter>
__global__ void __launch_bounds__(kRegWarps * 32, (EPT >= 16 ? 2 : 3))
importance_reg_kernel(const ImportanceRegArgs a) {
  constexpr int P = 32 * EPT;
  constexpr int EPC = EPT - EPF;
  static_assert(EPT > EPF, "room for the coarse samples");
  static_assert(KC == 0 || (KC % 32 == 0 && KC <= 32 * EPC && NI + ND <= 32 * EPF), "static shape must fit");
  __shared__ float s_cdf[kRegWarps][P];
  __shared__ float s_buf[kRegWarps][P];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* cdf = s_cdf[warp];
  float* buf = s_buf[warp];
  const bool do_sort = (a.z_sorted != nullptr);
  const int64_t warps = (int64_t)gridDim.x * kRegWarps;
  const LaneSigns sg = lane_signs(lane);
  bool first_ray = true;
  if (kFilter) {
    // 32 consecutive rays per step: each lane classifies one, the warp walks the matches
    // (prefetching the next match's inputs was tried: the extra registers cost more than the
    // exposed load latency, 3.2 vs 2.95 ms on BASELINE.json config 4's distribution)
    for (int64_t base = (blockIdx.x * (int64_t)kRegWarps + warp) * 32; base < a.R; base += warps * 32) {
      const int64_t rr = base + lane;
      bool mine = false;
      if (rr < a.R) {
        const int kc = (int)(a.offsets[rr + 1] - a.offsets[rr]);
        const int n = (int)(a.fine_offsets[rr + 1] - a.fine_offsets[rr]);
        int epf, ept;
        ray_class(kc, n + a.n_depth, epf, ept);
        mine = (epf == EPF && ept == EPT) && kc > 0 && !(a.skip_grp_classes && kc <= 256 && n <= 128);
      }
      unsigned todo = __ballot_sync(0xffffffffu, mine);
      while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        RayIn<EPF, EPC> cur;
        load_ray<EPF, EPC, KC, NI, ND>(cur, a, base + b, lane, do_sort);
        process_ray<EPF, EPT, KC, NI, ND>(a, cur, base + b, cdf, buf, lane, sg, first_ray);
      }
    }
    return;
  }
  int64_t r = blockIdx.x * (int64_t)kRegWarps + warp;
  RayIn<EPF, EPC> cur;
  if (r < a.R) load_ray<EPF, EPC, KC, NI, ND>(cur, a, r, lane, do_sort);
  for (; r < a.R; r += warps) {
    RayIn<EPF, EPC> nxt;
    const bool more = (r + warps < a.R);
    if (more) load_ray<EPF, EPC, KC, NI, ND>(nxt, a, r + warps, lane, do_sort);
    process_ray<EPF, EPT, KC, NI, ND>(a, cur, r, cdf, buf, lane, sg, first_ray);
    if (more) cur = nxt;
  }
}

template <int EPF, int EPT, int KC = 0, int NI = 0, int ND = 0>
static int launch_reg(const ImportanceRegArgs& a, cudaStream_t stream) {
  int64_t blocks = (a.R + kRegWarps - 1) / kRegWarps;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  importance_reg_kernel<EPF, EPT, KC, NI, ND, false><<<(unsigned)blocks, kRegWarps * 32, 0, stream>>>(a);
  return check_launch();
}

// packed layout: only the rays of class (EPF, EPT)
template <int EPF, int EPT>
static int launch_reg_class(const ImportanceRegArgs& a, cudaStream_t stream) {
  int64_t blocks = (a.R + kRegWarps * 32 - 1) / (kRegWarps * 32);
  const int64_t cap = (int64_t)num_sms() * (EPT >= 16 ? 2 : 3);
  if (blocks > cap) blocks = cap;
  importance_reg_kernel<EPF, EPT, 0, 0, 0, true><<<(unsigned)blocks, kRegWarps * 32, 0, stream>>>(a);
  return check_launch();
}

// Returns AVR_ERR_UNSUPPORTED when the shape does not fit a register variant (caller falls
// back to the shared-memory kernel).
int launch_importance_reg(const float* weights, const float* z_coarse, const float* u, const float* u2,
                          const float* normals, const float* near, const float* far, int bound_stride,
                          const int64_t* offsets, const int64_t* fine_offsets, int64_t R, int Kc, int n_imp,
                          int n_depth, float depth_std, float* z_fine, float* z_sorted, float* cdf, int32_t* idx,
                          cudaStream_t stream) {
  // packed layout: Kc / n_imp are the caller's per-ray maxima and size the register variant
  const int nd_eff = z_sorted ? n_depth : 0;
  const int m = n_imp + nd_eff;
  int epf = 1;
  while (32 * epf < m) epf <<= 1;
  int ept = 2 * epf;
  while (32 * ept < Kc + 32 * epf || 32 * ept < Kc + 2) ept <<= 1;
  if (epf > 8 || ept > 16 || m < 1) return AVR_ERR_UNSUPPORTED;
  ImportanceRegArgs a;
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
  a.n_depth = nd_eff;
  a.depth_std = depth_std;
  a.vec4 = !offsets && ((n_imp & 3) == 0) && aligned16(u) && aligned16(u2);
  a.vecw = !offsets && ((Kc & 3) == 0) && aligned16(weights);
  a.vecz = !offsets && ((Kc & 3) == 0) && ((nd_eff & 3) == 0) && ((n_imp & 3) == 0) && aligned16(z_coarse) &&
           aligned16(normals) && aligned16(z_sorted);
  a.skip_grp_classes = 0;
  a.z_fine = z_fine;
  a.z_sorted = z_sorted;
  a.cdf = cdf;
  a.idx = idx;
  // hot dense shapes: eight lanes per ray, shape as template constants (importance_grp.cu);
  // AVR_IMPORTANCE_GRP=0 keeps them on the warp-per-ray kernel (A/B experiments)
  if (!offsets) {
    // AVR_IMPORTANCE_BINS=1: bucket ranking (importance_bins.cu) instead of the sorting networks — exact and
    // shape-agnostic, but measured 2x slower on B200 (profiles/r02_importance_bins.md), so it is opt-in
    if (z_sorted && option(OPT_IMPORTANCE_BINS, 0)) {
      const int rc = launch_importance_bins(a, stream);
      if (rc == AVR_OK) count_dispatch(AVR_DISPATCH_IMPORTANCE_BINS);
      if (rc != AVR_ERR_UNSUPPORTED) return rc;
    }
    if (option(OPT_IMPORTANCE_GRP, 1)) {
      const int rc = launch_importance_grp(a, stream);
      if (rc == AVR_OK) count_dispatch(AVR_DISPATCH_IMPORTANCE_GRP);
      if (rc != AVR_ERR_UNSUPPORTED) return rc;
    }
#define AVR_REG_STATIC(F, T, KC_, NI_, ND_) \
  if (Kc == KC_ && n_imp == NI_ && nd_eff == ND_) return launch_reg<F, T, KC_, NI_, ND_>(a, stream);
    AVR_REG_STATIC(4, 8, 64, 128, 0)
    AVR_REG_STATIC(1, 4, 64, 16, 16)
#undef AVR_REG_STATIC
  }
  // packed layout with enough rays to amortise the launches: one launch per ray class up to
  // the caller's maxima (AVR_PACKED_CLASSES=0: one launch, every ray at the maximum shape)
  if (offsets && R >= 4096) {
    if (option(OPT_PACKED_CLASSES, 1)) {
      // rays of at most 256 coarse / 128 new samples: class kernels with 8..32 lanes per ray — the sorting
      // networks (importance_grp.cu) or, with AVR_IMPORTANCE_BINS=1, bucket ranking (importance_bins.cu);
      // a request for the cdf / the indices goes to the kernels that export them for packed rays
      const bool bins = z_sorted && a.n_depth == 0 && option(OPT_IMPORTANCE_BINS, 0);
      const bool grp = !bins && a.n_depth == 0 && !cdf && !idx && option(OPT_IMPORTANCE_GRP, 1);
      if (bins || grp) {
        bool covers_all = false;
        const int rc = bins ? launch_importance_bins_ragged(a, Kc, n_imp, &covers_all, stream)
                            : launch_importance_grp_ragged(a, Kc, n_imp, &covers_all, stream);
        if (rc == AVR_OK) count_dispatch(bins ? AVR_DISPATCH_IMPORTANCE_BINS : AVR_DISPATCH_IMPORTANCE_GRP);
        if (rc != AVR_OK || covers_all) return rc;
        a.skip_grp_classes = 1;
      }
#define AVR_REG_CLASS(F, T)                              \
  if (F <= epf && T <= ept) {                            \
    const int rc = launch_reg_class<F, T>(a, stream);    \
    if (rc != AVR_OK) return rc;                         \
  }
      AVR_REG_CLASS(1, 2)
      AVR_REG_CLASS(1, 4)
      AVR_REG_CLASS(1, 8)
      AVR_REG_CLASS(1, 16)
      AVR_REG_CLASS(2, 4)
      AVR_REG_CLASS(2, 8)
      AVR_REG_CLASS(2, 16)
      AVR_REG_CLASS(4, 8)
      AVR_REG_CLASS(4, 16)
      AVR_REG_CLASS(8, 16)
#undef AVR_REG_CLASS
      return AVR_OK;
    }
  }
#define AVR_REG_CASE(F, T) \
  if (epf == F && ept == T) return launch_reg<F, T>(a, stream);
  AVR_REG_CASE(1, 2)
  AVR_REG_CASE(1, 4)
  AVR_REG_CASE(1, 8)
  AVR_REG_CASE(1, 16)
  AVR_REG_CASE(2, 4)
  AVR_REG_CASE(2, 8)
  AVR_REG_CASE(2, 16)
  AVR_REG_CASE(4, 8)
  AVR_REG_CASE(4, 16)
  AVR_REG_CASE(8, 16)
#undef AVR_REG_CASE
  return AVR_ERR_UNSUPPORTED;
}

}  // namespace avr
