// Packed (ragged) stratified coarse sampler: the per-lane work of
// coarse_fwd_packed_flat_kernel (samplers.cu), written so that the SAME code also compiles as
// plain host C++.  tests/test_host_kernel_cores.py builds it with g++ and walks a whole launch
// lane by lane against the reference formula (renderers.py:12-14) — the kernel's index
// arithmetic and its division shortcut are checked in the CPU suite, the GPU tests then only
// have to confirm the launch itself.
//
// Layout of the work.  A warp owns a SEGMENT = 32 consecutive rays, i.e. one contiguous slice
// [offsets[r0], offsets[r0+32]) of the packed u / z streams.  The 33 offsets (relative to the
// segment start) and the rays' parameters go to shared memory once; after that the segment is
// processed as a flat stream, four consecutive samples (one 16-byte access) per lane and step.
// Rays of any length, including empty ones, cost nothing beyond their samples: no lane idles on a
// short ray and there is one dependent global-load chain per 32 rays instead of one per ray.
//
// Finding the ray of a group.  The 32 groups a warp handles in one step ("a row") are 128
// consecutive samples, so they touch few rays: a vote over the lanes' own offsets gives the ray of
// the row's first sample, and a lane adds how many of the next two ray starts lie at or before its
// group (rows that hold more than two ray starts take the 5-probe search of the 33 offsets).
//
// Inside a group.  A group lies in one ray (A) or crosses into the next one (B); as long as B
// reaches the group's end the four samples are computed without a branch, each selecting its
// ray's constants.  Everything else (rays shorter than the rest of a group, empty rays in between,
// partial groups at the segment's ends, unaligned streams, numerators outside the division
// shortcut's range) takes the general walk, sample by sample.
//
// Division.  z = near + span*(j/K) + (u*span)/K has two IEEE divisions per sample
// (renderers.py:12, :14).  K is constant along a ray, so y = RN(1/K) is formed once per ray and
// every a/K is  q0 = RN(a*y); r = fma(-K, q0, a); q = fma(r, y, q0)  — Markstein's sequence,
// which returns the correctly rounded quotient whenever y is the correctly rounded reciprocal,
// K's significand is not all ones and nothing leaves the normal range on the way.  Rays and
// numerators outside those conditions take the IEEE division.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define AVR_HD __device__ __forceinline__
#else
#include <math.h>
#define AVR_HD inline
#endif

namespace avr {

#if defined(__CUDACC__)
AVR_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
AVR_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
AVR_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
AVR_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
AVR_HD float f_rcp(float a) { return __frcp_rn(a); }
AVR_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
// host build: -ffp-contract=off keeps these as single IEEE operations
AVR_HD float f_mul(float a, float b) { return a * b; }
AVR_HD float f_add(float a, float b) { return a + b; }
AVR_HD float f_sub(float a, float b) { return a - b; }
AVR_HD float f_div(float a, float b) { return a / b; }
AVR_HD float f_rcp(float a) { return 1.0f / a; }
AVR_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

constexpr int kSegRays = 32;                       // rays per segment (= lanes per warp)
constexpr int kMarksteinMaxCount = (1 << 24) - 2;  // above: float(K) may have an all-ones significand

// Per-ray constants of a segment (one 16-byte shared-memory load per ray change).
struct CoarseRay {
  float near, span, kf, rcp;  // rcp == 0 marks a ray that divides the IEEE way
};

struct CoarseSegment {
  int rel[kSegRays + 4];  // offsets relative to the segment start; rel[32] = segment length; rel[33..35] = INT_MAX
  CoarseRay ray[kSegRays + 1];  // ray[32]: a stand-in (rcp == 0) so that "the ray after k" always exists
};

// a / K given y = RN(1/K): Markstein's correction, the correctly rounded quotient under the
// conditions checked by the caller (see the header comment)
AVR_HD float div_markstein(float a, float kf, float y) {
  const float q0 = f_mul(a, y);
  const float r = f_fma(-kf, q0, a);
  return f_fma(r, y, q0);
}

// renderers.py:12-14 for sample j of a ray.  The numerators are j (an integer below 2^24) and
// u*span; the latter takes the shortcut only for magnitudes in [2^-90, 2^100) or zero, where no
// intermediate of the sequence can leave the normal range (inf / nan / denormals divide the IEEE way).
AVR_HD float coarse_depth_ray(const CoarseRay& p, int j, float u) {
  const float jf = (float)j;
  const float jit = f_mul(u, p.span);
  const float aj = jit < 0.f ? -jit : jit;
  const bool fast = p.rcp != 0.f && ((aj >= 0x1p-90f && aj < 0x1p100f) || aj == 0.f);
  float bin, t;
  if (fast) {
    bin = div_markstein(jf, p.kf, p.rcp);
    t = div_markstein(jit, p.kf, p.rcp);
  } else {
    bin = f_div(jf, p.kf);
    t = f_div(jit, p.kf);
  }
  return f_add(f_add(p.near, f_mul(p.span, bin)), t);
}

// Phase 1, lane `lane` of the warp that owns rays [r0, r0+32): fill this lane's table entries.
// Returns false when the segment cannot use 32-bit relative offsets (its caller then takes the
// per-ray loop for the whole segment; warp-uniform after a vote).
AVR_HD bool coarse_segment_build(int lane, int64_t r0, int64_t R, const int64_t* offsets, const float* near,
                                 const float* far, int bound_stride, CoarseSegment* seg) {
  const int64_t seg_begin = offsets[r0];
  const int64_t r = r0 + lane < R ? r0 + lane : R;  // rays past the end are empty
  const int64_t begin = offsets[r];
  const int64_t end = offsets[r < R ? r + 1 : R];
  const int64_t rel = begin - seg_begin;
  const int64_t cnt = end - begin;
  const bool ok = rel >= 0 && cnt >= 0 && (end - seg_begin) < (int64_t)0x7ffffff0;  // head + len + 3 stays an int
  seg->rel[lane] = (int)rel;
  if (lane == kSegRays - 1) seg->rel[kSegRays] = (int)(end - seg_begin);
  if (lane == 0) {
    seg->rel[kSegRays + 1] = seg->rel[kSegRays + 2] = seg->rel[kSegRays + 3] = 0x7fffffff;
    CoarseRay none;
    none.near = none.span = none.kf = none.rcp = 0.f;
    seg->ray[kSegRays] = none;
  }
  CoarseRay p;
  const int64_t b = bound_stride ? (r < R ? r : R - 1) : 0;
  p.near = near[b];
  p.span = f_sub(far[b], p.near);
  p.kf = (float)cnt;
  p.rcp = (cnt > 0 && cnt <= kMarksteinMaxCount) ? f_rcp(p.kf) : 0.f;
  seg->ray[lane] = p;
  return ok;
}

// largest k in [0, 31] with rel[k] <= i  (0 <= i < rel[32]); with empty rays (equal entries)
// this is the ray that actually holds sample i
AVR_HD int coarse_segment_find(const CoarseSegment* seg, int i) {
  int k = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int step = kSegRays / 2; step > 0; step >>= 1)
    if (seg->rel[k + step] <= i) k += step;
  return k;
}

// Phase 2, lane `lane`: samples of the segment, four consecutive ones (one aligned 16-byte
// access of the packed streams) per group.  A group is described by the relative index i0 of
// its first slot; slots outside [0, seg_len) belong to the neighbouring segments and are
// neither read nor written.
struct CoarseGroup {
  int i0, lo, hi;  // slots [lo, hi) of [i0, i0+4) are this segment's
  bool full;       // all four, and the streams are 16-byte aligned: one vector access
  float u[4];
};

// The draws of a whole, vector-loadable group: requested as soon as its index is known (a warp has
// AVR_COARSE_ROWS such loads per lane in flight; with AVR_COARSE_AHEAD those of the NEXT step, issued before
// the current step is computed).  Anything else is left to coarse_group_load.
struct CoarseDraws {
  float u[4];
};

AVR_HD CoarseDraws coarse_group_request(int g, int n_groups, int head, int seg_len, const float* ua, bool vec_ok) {
  CoarseDraws d;
  d.u[0] = d.u[1] = d.u[2] = d.u[3] = 0.f;
  const int i0 = 4 * g - head;
  if (vec_ok && g < n_groups && i0 >= 0 && i0 + 4 <= seg_len) {
#if defined(__CUDACC__)
    const float4 q = __ldcs(reinterpret_cast<const float4*>(ua + 4 * g));
    d.u[0] = q.x; d.u[1] = q.y; d.u[2] = q.z; d.u[3] = q.w;
#else
    for (int q = 0; q < 4; ++q) d.u[q] = ua[4 * g + q];
#endif
  }
  return d;
}

AVR_HD CoarseGroup coarse_group_load(int g, int head, int seg_len, const float* ua, bool vec_ok, const CoarseDraws& req) {
  CoarseGroup c;
  c.i0 = 4 * g - head;
  c.lo = c.i0 < 0 ? 0 : c.i0;
  c.hi = c.i0 + 4 < seg_len ? c.i0 + 4 : seg_len;
  c.full = vec_ok && (c.hi - c.lo == 4);
  if (c.full) {
    for (int q = 0; q < 4; ++q) c.u[q] = req.u[q];
  } else {
    c.u[0] = c.u[1] = c.u[2] = c.u[3] = 0.f;
    for (int q = 0; q < 4; ++q)
      if (c.i0 + q >= c.lo && c.i0 + q < c.hi) c.u[q] = ua[4 * g + q];
  }
  return c;
}

// the general walk of a group, sample by sample, starting from the ray k of its first slot
AVR_HD void coarse_group_walk(const CoarseGroup& c, int g, int k, const CoarseSegment* seg, float* za) {
  CoarseRay p = seg->ray[k];
  int rb = seg->rel[k], re = seg->rel[k + 1];
  float out[4];
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    const int i = c.i0 + q;
    out[q] = 0.f;
    if (i < c.lo || i >= c.hi) continue;
    if (i >= re) {  // on to the next non-empty ray (i < seg_len bounds the walk)
      do {
        ++k;
        re = seg->rel[k + 1];
      } while (i >= re);
      rb = seg->rel[k];
      p = seg->ray[k];
    }
    out[q] = coarse_depth_ray(p, i - rb, c.u[q]);
  }
  if (c.full) {
#if defined(__CUDACC__)
    __stcs(reinterpret_cast<float4*>(za + 4 * g), make_float4(out[0], out[1], out[2], out[3]));
#else
    for (int q = 0; q < 4; ++q) za[4 * g + q] = out[q];
#endif
  } else {
    for (int q = 0; q < 4; ++q)
      if (c.i0 + q >= c.lo && c.i0 + q < c.hi) za[4 * g + q] = out[q];
  }
}

// A group whose first slot lies in ray k.  Whole groups inside ray k, or crossing into a ray k+1
// that reaches the group's end, are computed branch-free; the rest takes coarse_group_walk.
AVR_HD void coarse_group_finish(const CoarseGroup& c, int g, int k, const CoarseSegment* seg, float* za) {
  const int rb = seg->rel[k], re = seg->rel[k + 1], re2 = seg->rel[k + 2];
  const CoarseRay A = seg->ray[k], B = seg->ray[k + 1];
  const int end4 = c.i0 + 4;
  // (no short-circuit: lanes whose group crosses into B must not leave the others behind and run the block twice)
  const bool two_rays = c.full & (A.rcp != 0.f) & ((re >= end4) | ((re2 >= end4) & (B.rcp != 0.f)));
  if (two_rays) {
    const float ja = (float)(c.i0 - rb), jb = (float)(c.i0 - re);  // sample numbers of slot 0 in A / in B (exact)
    const int to_b = re - c.i0;                                     // slots to_b.. of the group lie in B
    float nr[4], sp[4], kf[4], rc[4], jf[4], jit[4];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) {
      const bool in_b = q >= to_b;
      nr[q] = in_b ? B.near : A.near;
      sp[q] = in_b ? B.span : A.span;
      kf[q] = in_b ? B.kf : A.kf;
      rc[q] = in_b ? B.rcp : A.rcp;
      const float ja_q = f_add(ja, (float)q), jb_q = f_add(jb, (float)q);  // exact
      jf[q] = in_b ? jb_q : ja_q;
      jit[q] = f_mul(c.u[q], sp[q]);
    }
    // the division shortcut's range for all four numerators at once (a NaN passes and stays a NaN); zero draws
    // (one in 2^24) walk
    const float a0 = fabsf(jit[0]), a1 = fabsf(jit[1]), a2 = fabsf(jit[2]), a3 = fabsf(jit[3]);
    const float amin = fminf(fminf(a0, a1), fminf(a2, a3)), amax = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
    const bool fast = (amin >= 0x1p-90f) & (amax < 0x1p100f);
    if (fast) {
      float out[4];
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int q = 0; q < 4; ++q) {
        const float bin = div_markstein(jf[q], kf[q], rc[q]);
        const float t = div_markstein(jit[q], kf[q], rc[q]);
        out[q] = f_add(f_add(nr[q], f_mul(sp[q], bin)), t);
      }
#if defined(__CUDACC__)
      __stcs(reinterpret_cast<float4*>(za + 4 * g), make_float4(out[0], out[1], out[2], out[3]));
#else
      for (int q = 0; q < 4; ++q) za[4 * g + q] = out[q];
#endif
      return;
    }
  }
  coarse_group_walk(c, g, k, seg, za);
}

// number of rays of the segment that start at or before sample i: a vote over the lanes' own offsets
// (every lane of the warp calls it with the same i); the host walk counts the table instead
AVR_HD int coarse_starts_le(const CoarseSegment* seg, int my_rel, int i) {
#if defined(__CUDACC__)
  (void)seg;
  return __popc(__ballot_sync(0xffffffffu, my_rel <= i));
#else
  (void)my_rel;
  int n = 0;
  for (int l = 0; l < kSegRays; ++l) n += seg->rel[l] <= i;
  return n;
#endif
}

// One row: groups g_row .. g_row+31, lane `lane` owning group g_row + lane (the draws of a whole group arrive
// in `req`).  Every lane of the warp enters (the vote), lanes without a group leave after it.
AVR_HD void coarse_row_finish(const CoarseDraws& req, int lane, int my_rel, int g_row, int n_groups, int head,
                              int seg_len, const CoarseSegment* seg, const float* ua, float* za, bool vec_ok) {
  int i_first = 4 * g_row - head, i_last = 4 * (g_row + 31) - head + 3;
  i_first = i_first < 0 ? 0 : i_first;
  i_last = i_last < seg_len ? i_last : seg_len - 1;
  const int k_first = coarse_starts_le(seg, my_rel, i_first) - 1;  // rel[0] == 0: at least one
  // the next three ray starts (warp-uniform reads; past the table's 33 offsets they are INT_MAX): when the third
  // lies beyond the row, a group's ray is k_first plus the starts among the first two at or before the group
  const int s1 = seg->rel[k_first + 1], s2 = seg->rel[k_first + 2], s3 = seg->rel[k_first + 3];
  const int g = g_row + lane;
  if (g >= n_groups) return;
  const CoarseGroup c = coarse_group_load(g, head, seg_len, ua, vec_ok, req);
  if (c.hi <= c.lo) return;
  int k;
  if (s3 > i_last) {
    k = k_first + (s1 <= c.lo ? 1 : 0) + (s2 <= c.lo ? 1 : 0);  // rel is non-decreasing
  } else {
    k = coarse_segment_find(seg, c.lo);
  }
  coarse_group_finish(c, g, k, seg, za);
}

#ifndef AVR_COARSE_ROWS
#define AVR_COARSE_ROWS 3   // rows (16-byte loads per lane) per step
#endif
#ifndef AVR_COARSE_AHEAD
#define AVR_COARSE_AHEAD 0  // 1: the next step's draws are requested before the current step is computed
#endif

// `vec_ok`: u and z are 16-byte aligned.
AVR_HD void coarse_segment_run(int lane, const CoarseSegment* seg, int64_t seg_begin, const float* u, float* z,
                               bool vec_ok) {
  constexpr int NR = AVR_COARSE_ROWS;
  const int seg_len = seg->rel[kSegRays];
  const int my_rel = seg->rel[lane];
  const int head = (int)(seg_begin & 3);  // samples between the 16-byte boundary below and the segment
  const int n_groups = (head + seg_len + 3) >> 2;
  const float* ua = u + (seg_begin - head);
  float* za = z + (seg_begin - head);
  CoarseDraws cur[NR], nxt[NR];
  if (AVR_COARSE_AHEAD) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < NR; ++r) cur[r] = coarse_group_request(lane + 32 * r, n_groups, head, seg_len, ua, vec_ok);
  }
  for (int g_row = 0; g_row < n_groups; g_row += 32 * NR) {  // the same trip count on every lane
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < NR; ++r) {
      if (AVR_COARSE_AHEAD) {
        nxt[r] = coarse_group_request(g_row + 32 * (NR + r) + lane, n_groups, head, seg_len, ua, vec_ok);
      } else {
        cur[r] = coarse_group_request(g_row + 32 * r + lane, n_groups, head, seg_len, ua, vec_ok);
      }
    }
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < NR; ++r) {
      if (r == 0 || g_row + 32 * r < n_groups)
        coarse_row_finish(cur[r], lane, my_rel, g_row + 32 * r, n_groups, head, seg_len, seg, ua, za, vec_ok);
    }
    if (AVR_COARSE_AHEAD) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int r = 0; r < NR; ++r) cur[r] = nxt[r];
    }
  }
}

// Per-ray loop for a segment the flat walk cannot take (>= 2^31 samples in 32 rays).
AVR_HD void coarse_segment_slow(int lane, int64_t r0, int64_t R, const int64_t* offsets, const float* near,
                                const float* far, int bound_stride, const float* u, float* z) {
  for (int64_t r = r0; r < r0 + kSegRays && r < R; ++r) {
    const int64_t begin = offsets[r];
    const int64_t cnt = offsets[r + 1] - begin;
    const int64_t b = bound_stride ? r : 0;
    const float nr = near[b], span = f_sub(far[b], nr), kf = (float)cnt;
    for (int64_t j = lane; j < cnt; j += 32) {
      const float zz = f_add(nr, f_mul(span, f_div((float)j, kf)));
      z[begin + j] = f_add(zz, f_div(f_mul(u[begin + j], span), kf));
    }
  }
}

}  // namespace avr
