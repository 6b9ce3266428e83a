// Host walk of the packed coarse sampler's per-lane code (csrc/coarse_packed_core.h): the same
// functions the CUDA kernel calls, driven warp by warp, lane by lane.  Test infrastructure only
// (built by tests/test_host_kernel_cores.py with g++ -ffp-contract=off); never part of the library.
#include <stdint.h>
#include <string.h>

#include "coarse_packed_core.h"

extern "C" {

// Mirrors coarse_fwd_packed_flat_kernel: phase 1 for all 32 lanes, the vote, phase 2 for all lanes.
// `force_slow` != 0 takes the per-ray loop for every segment.  Returns the number of segments that
// took the per-ray loop.
int64_t host_coarse_packed(const float* near, const float* far, int bound_stride, const float* u,
                           const int64_t* offsets, int64_t R, float* z, int vec_ok, int force_slow) {
  int64_t slow = 0;
  const int64_t n_seg = (R + avr::kSegRays - 1) / avr::kSegRays;
  for (int64_t s = 0; s < n_seg; ++s) {
    avr::CoarseSegment seg;
    memset(&seg, 0xff, sizeof seg);
    const int64_t r0 = s * avr::kSegRays;
    bool all_ok = true;
    for (int lane = 0; lane < 32; ++lane)
      all_ok &= avr::coarse_segment_build(lane, r0, R, offsets, near, far, bound_stride, &seg);
    if (all_ok && !force_slow) {
      for (int lane = 0; lane < 32; ++lane) avr::coarse_segment_run(lane, &seg, offsets[r0], u, z, vec_ok != 0);
    } else {
      ++slow;
      for (int lane = 0; lane < 32; ++lane)
        avr::coarse_segment_slow(lane, r0, R, offsets, near, far, bound_stride, u, z);
    }
  }
  return slow;
}

// a / K by the kernel's shortcut (as coarse_depth_ray applies it) and by IEEE division, for
// exhaustive comparisons.  Returns the number of mismatching bit patterns.
int64_t host_markstein_mismatches(const float* a, int64_t n, int k_lo, int k_hi) {
  int64_t bad = 0;
  for (int k = k_lo; k <= k_hi; ++k) {
    const float kf = (float)k, y = avr::f_rcp(kf);
    for (int64_t i = 0; i < n; ++i) {
      const float q = avr::div_markstein(a[i], kf, y), want = a[i] / kf;
      bad += memcmp(&q, &want, 4) != 0;
    }
    for (int j = 0; j < k && j < (1 << 16); ++j) {  // the bin numerators
      const float q = avr::div_markstein((float)j, kf, y), want = (float)j / kf;
      bad += memcmp(&q, &want, 4) != 0;
    }
  }
  return bad;
}

}  // extern "C"
