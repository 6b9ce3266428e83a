// Register-resident sorting networks shared by the importance samplers
// (importance_reg.cu: one warp per ray; importance_grp.cu: G lanes per ray).
//
// Cross-lane compare-exchanges run in a "signed domain": a lane that must keep the maximum
// holds its keys negated, so BOTH partners execute the same min(w, -other) (one FMNMX instead
// of a min/max pair on the half-rate ALU pipe); the sign pattern changes between stages by one
// multiply by +-1 (FMA pipe, exact).
#pragma once

#include <math_constants.h>

namespace avr {

__host__ __device__ constexpr int ilog2_c(int v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }

__device__ __forceinline__ void cmp_swap_asc(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}

// sg[b] = -1 if lane bit b is set, +1 otherwise.  A cross-lane compare-exchange over lane
// bit b in the signed domain w = sg[b] * v is  w <- min(w, -w_partner)  on both partners.
struct LaneSigns {
  float s[5];
};
__device__ __forceinline__ LaneSigns lane_signs(int lane) {
  LaneSigns g;
#pragma unroll
  for (int b = 0; b < 5; ++b) g.s[b] = ((lane >> b) & 1) ? -1.0f : 1.0f;
  return g;
}

// Ascending sort of a lane's own EPF keys by the smallest known comparator networks (5 / 19 / 60 compare-exchanges
// for 4 / 8 / 16 keys: Knuth, TAOCP 3, section 5.3.4; the lane-local phases of a bitonic sort need 6 / 24 / 80).
// Each list was checked over all 2^EPF zero-one inputs (tools/check_sort_nets.py).
template <int EPF>
__device__ __forceinline__ bool presort_lane(float (&v)[EPF]) {
#define AVR_CE(a, b) cmp_swap_asc(v[a], v[b]);
  if constexpr (EPF == 4) {
    AVR_CE(0, 2) AVR_CE(1, 3) AVR_CE(0, 1) AVR_CE(2, 3) AVR_CE(1, 2)
    return true;
  } else if constexpr (EPF == 8) {
    AVR_CE(0, 2) AVR_CE(1, 3) AVR_CE(4, 6) AVR_CE(5, 7) AVR_CE(0, 4) AVR_CE(1, 5) AVR_CE(2, 6) AVR_CE(3, 7)
    AVR_CE(0, 1) AVR_CE(2, 3) AVR_CE(4, 5) AVR_CE(6, 7) AVR_CE(2, 4) AVR_CE(3, 5) AVR_CE(1, 4) AVR_CE(3, 6)
    AVR_CE(1, 2) AVR_CE(3, 4) AVR_CE(5, 6)
    return true;
  } else if constexpr (EPF == 16) {
    AVR_CE(0, 13) AVR_CE(1, 12) AVR_CE(2, 15) AVR_CE(3, 14) AVR_CE(4, 8) AVR_CE(5, 6) AVR_CE(7, 11) AVR_CE(9, 10)
    AVR_CE(0, 5) AVR_CE(1, 7) AVR_CE(2, 9) AVR_CE(3, 4) AVR_CE(6, 13) AVR_CE(8, 14) AVR_CE(10, 15) AVR_CE(11, 12)
    AVR_CE(0, 1) AVR_CE(2, 3) AVR_CE(4, 5) AVR_CE(6, 8) AVR_CE(7, 9) AVR_CE(10, 11) AVR_CE(12, 13) AVR_CE(14, 15)
    AVR_CE(0, 2) AVR_CE(1, 3) AVR_CE(4, 10) AVR_CE(5, 11) AVR_CE(6, 7) AVR_CE(8, 9) AVR_CE(12, 14) AVR_CE(13, 15)
    AVR_CE(1, 2) AVR_CE(3, 12) AVR_CE(4, 6) AVR_CE(5, 7) AVR_CE(8, 10) AVR_CE(9, 11) AVR_CE(13, 14)
    AVR_CE(1, 4) AVR_CE(2, 6) AVR_CE(5, 8) AVR_CE(7, 10) AVR_CE(9, 13) AVR_CE(11, 14)
    AVR_CE(2, 4) AVR_CE(3, 6) AVR_CE(9, 12) AVR_CE(11, 13)
    AVR_CE(3, 5) AVR_CE(6, 8) AVR_CE(7, 9) AVR_CE(10, 12)
    AVR_CE(3, 4) AVR_CE(5, 6) AVR_CE(7, 8) AVR_CE(9, 10) AVR_CE(11, 12)
    AVR_CE(6, 7) AVR_CE(8, 9)
    return true;
  }
#undef AVR_CE
  return false;
}

// Full bitonic sort of G*EPF keys, blocked layout: key index e = (lane % G)*EPF + r.
// All comparators point the same way (the lower index keeps the minimum): each merge phase
// opens with a "mirror" step (e <-> e ^ (size-1)) instead of alternating directions, so the
// lane-local compare-exchanges need no direction selects.
// G = lanes that share one key set (32: the whole warp; 8: four independent sorts per warp).
template <int EPF, int G = 32>
__device__ __forceinline__ void sort_blocked(float (&v)[EPF], const LaneSigns& sg) {
  constexpr int M = G * EPF;
  int cur = -1;  // lane bit whose sign pattern the keys carry; -1 = true values (compile-time after unrolling)
  constexpr bool presorted = (EPF == 4 || EPF == 8 || EPF == 16);  // the phases up to size EPF, by a smaller network
  presort_lane<EPF>(v);
#pragma unroll
  for (int size = 2; size <= M; size <<= 1) {
    if (presorted && size <= EPF) continue;
    // mirror step
    if (size <= EPF) {
#pragma unroll
      for (int r = 0; r < EPF; ++r) {
        const int pr = r ^ (size - 1);
        if (r < pr) cmp_swap_asc(v[r], v[pr]);
      }
    } else {
      const int lmask = size / EPF - 1;          // partner lane = lane ^ lmask
      const int kb = ilog2_c(size / EPF / 2);    // the lane bit that says "upper index of the pair"
      const float f = (cur < 0) ? sg.s[kb] : sg.s[cur] * sg.s[kb];
      float o[EPF];
#pragma unroll
      for (int r = 0; r < EPF; ++r) v[r] *= f;
#pragma unroll
      for (int r = 0; r < EPF; ++r) o[r] = __shfl_xor_sync(0xffffffffu, v[EPF - 1 - r], lmask);
#pragma unroll
      for (int r = 0; r < EPF; ++r) v[r] = fminf(v[r], -o[r]);
      cur = kb;
    }
    // half-cleaner steps
#pragma unroll
    for (int stride = size >> 2; stride > 0; stride >>= 1) {
      if (stride < EPF) {
        if (cur >= 0) {
#pragma unroll
          for (int r = 0; r < EPF; ++r) v[r] *= sg.s[cur];
          cur = -1;
        }
#pragma unroll
        for (int r = 0; r < EPF; ++r) {
          if ((r & stride) == 0) cmp_swap_asc(v[r], v[r | stride]);
        }
      } else {
        const int lb = ilog2_c(stride / EPF);
        const float f = (cur < 0) ? sg.s[lb] : sg.s[cur] * sg.s[lb];
#pragma unroll
        for (int r = 0; r < EPF; ++r) {
          const float w = v[r] * f;
          const float o = __shfl_xor_sync(0xffffffffu, w, 1 << lb);
          v[r] = fminf(w, -o);
        }
        cur = lb;
      }
    }
  }
  if (cur >= 0) {
#pragma unroll
    for (int r = 0; r < EPF; ++r) v[r] *= sg.s[cur];
  }
}

// Ascending bitonic merge of G*EPT keys.  Two register layouts (g = lane % G):
//   striped (kQuad == false): key index q = i*G + g                       — a load/store of register i covers
//                             G consecutive floats per ray;
//   quads   (kQuad == true):  key index q = (i >> 2)*4G + 4g + (i & 3)    — a lane owns groups of FOUR
//                             consecutive keys, so rows move with 16-byte accesses (G*16 contiguous bytes per
//                             ray and instruction: the fewest LSU wavefronts a row can cost).
// Index bits from the top: the high register bits (lane-local compare-exchanges), the lane bits (shuffles,
// signed domain), then — quads only — the two low register bits.
// kInf0: bit i set = register i is known to hold +inf on every lane (padding); comparators
// with such an input reduce to nothing or to a register move, resolved at compile time.
template <int EPT, unsigned kInf0, int G, bool kQuad>
__device__ __forceinline__ void merge_net(float (&x)[EPT], const LaneSigns& sg) {
  constexpr int kTop = ilog2_c(G) - 1;  // highest lane bit inside a group
  constexpr int kLow = kQuad ? 4 : 1;   // register strides below this one come after the lane bits
  static_assert(!kQuad || EPT % 4 == 0, "quads need whole groups of four registers");
  unsigned infm = kInf0;
  auto local_stage = [&](const int istride) {
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if ((i & istride) == 0) {
        const int j = i | istride;
        if ((infm >> j) & 1u) {
          // upper input is +inf: already ordered
        } else if ((infm >> i) & 1u) {
          x[i] = x[j];
          x[j] = CUDART_INF_F;
          infm = (infm & ~(1u << i)) | (1u << j);
        } else {
          cmp_swap_asc(x[i], x[j]);
        }
      }
    }
  };
#pragma unroll
  for (int istride = EPT / 2; istride >= kLow; istride >>= 1) local_stage(istride);
#pragma unroll
  for (int b = kTop; b >= 0; --b) {
    const float f = (b == kTop) ? sg.s[kTop] : sg.s[b + 1] * sg.s[b];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if (!((infm >> i) & 1u)) {
        const float w = x[i] * f;
        const float o = __shfl_xor_sync(0xffffffffu, w, 1 << b);
        x[i] = fminf(w, -o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    if (!((infm >> i) & 1u)) x[i] *= sg.s[0];
  }
#pragma unroll
  for (int istride = kLow / 2; istride >= 1; istride >>= 1) local_stage(istride);
}

template <int EPT, unsigned kInf0, int G = 32>
__device__ __forceinline__ void merge_striped(float (&x)[EPT], const LaneSigns& sg) {
  merge_net<EPT, kInf0, G, false>(x, sg);
}

}  // namespace avr
