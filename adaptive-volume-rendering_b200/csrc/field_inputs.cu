// Radiance-field front end kernels (models.py:754-826): the per-lane
// code lives in field_inputs_core.h, shared with the host walk of the CPU test suite.
//
// Bound by the HBM write of the MLP input: (C + code) * 4 bytes per (view, point), 2216 B for
// conf/default.conf's 512 + 42; the feature rows it blends come from the L2-resident map (8 MB at
// 64 x 64 x 512) and stay in registers while consecutive samples of a ray sit in one texel cell.
#include <cstdlib>

#include "avr_common.cuh"
#include "field_inputs_core.h"
#include "kernels.h"

namespace avr {

constexpr int kFieldWarps = 4;
constexpr int kFieldChunk = 16;  // consecutive rows per warp visit: samples of one ray, same view

template <int CPL>
__global__ void __launch_bounds__(kFieldWarps * 32, 4)
field_inputs_fwd_kernel(const FieldInputsArgs a, int row_stride) {
  constexpr int N = CPL > 0 ? CPL : 1;
  const int lane = threadIdx.x & 31;
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + kFieldChunk - 1) / kFieldChunk;
  const int64_t warps = (int64_t)gridDim.x * kFieldWarps;
  const FieldLaneCode lc = field_lane_code(a, lane);
  FieldTapCache<N> cache;
  FieldView view;
  field_cache_reset(&cache);
  field_view_reset(&view);
  for (int64_t ch = blockIdx.x * (int64_t)kFieldWarps + (threadIdx.x >> 5); ch < n_chunks; ch += warps) {
    const int64_t first = ch * kFieldChunk;
    const int n = (int)(first + kFieldChunk < rows ? kFieldChunk : rows - first);
    FieldCursor cur = field_cursor_at(a, first);
    for (int r = 0; r < n; ++r, field_cursor_next(a, &cur)) {
      field_view_fill(a, cur, &view);
      const FieldPoint p = field_point(a, cur, view);
      if (CPL > 0) {
        field_row_lane<N>(a, cur, p, lane, row_stride, lc, &cache);
      } else {
        field_row_lane_generic(a, cur, p, lane, row_stride, lc);
      }
    }
  }
}

// one gradient kind per launch: capped at 170 registers (3 CTAs = 12 warps per SM; the uncapped 180 of
// the feature-map variant left 8 and ran 0.219 instead of 0.194 ms)
template <int CPL, bool kLatent, bool kPoint>
__global__ void __launch_bounds__(kFieldWarps * 32, (kLatent && kPoint) ? 1 : 3)
field_inputs_bwd_kernel(const FieldInputsArgs a, int row_stride) {
  constexpr int N = CPL > 0 ? CPL : 1;
  const int lane = threadIdx.x & 31;
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + kFieldChunk - 1) / kFieldChunk;
  const int64_t warps = (int64_t)gridDim.x * kFieldWarps;
  const FieldLaneCode lc = field_lane_code(a, lane);
  FieldTapCache<N> taps;
  FieldGradCache<N> grads;
  FieldView view;
  field_cache_reset(&taps);
  field_grad_reset(&grads);
  field_view_reset(&view);
  for (int64_t ch = blockIdx.x * (int64_t)kFieldWarps + (threadIdx.x >> 5); ch < n_chunks; ch += warps) {
    const int64_t first = ch * kFieldChunk;
    const int n = (int)(first + kFieldChunk < rows ? kFieldChunk : rows - first);
    FieldCursor cur = field_cursor_at(a, first);
    for (int r = 0; r < n; ++r, field_cursor_next(a, &cur)) {
      field_view_fill(a, cur, &view);
      const FieldPoint p = field_point(a, cur, view);
      FieldRowPartial s;
      if (CPL > 0) {
        s = field_bwd_row_lane<N, kLatent, kPoint>(a, cur, p, lane, row_stride, lc, &taps, &grads);
      } else {
        s = field_bwd_row_lane_generic<kLatent, kPoint>(a, cur, p, lane, row_stride, lc);
      }
      if (kPoint) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          s.gix += __shfl_xor_sync(0xffffffffu, s.gix, d);
          s.giy += __shfl_xor_sync(0xffffffffu, s.giy, d);
          s.enc0 += __shfl_xor_sync(0xffffffffu, s.enc0, d);
          s.enc1 += __shfl_xor_sync(0xffffffffu, s.enc1, d);
          s.enc2 += __shfl_xor_sync(0xffffffffu, s.enc2, d);
          s.vr0 += __shfl_xor_sync(0xffffffffu, s.vr0, d);
          s.vr1 += __shfl_xor_sync(0xffffffffu, s.vr1, d);
          s.vr2 += __shfl_xor_sync(0xffffffffu, s.vr2, d);
        }
        if (lane == 0) field_bwd_row_finish(a, cur, view, p, s);
      }
    }
  }
  if (kLatent && CPL > 0) field_grad_flush<N>(a, lane, &grads);
}

// Experiment knobs (A/B measurements; defaults are what the measurements picked):
//   AVR_FIELD_NOCACHE=1    every channel count takes the generic walk (no register caches)
//   AVR_FIELD_BWD_SPLIT=0  feature-map and point gradients in ONE launch instead of two (234 registers,
//                          8 resident warps/SM: 0.645 ms against 0.56 ms for the two launches)
static bool env_flag(const char* name, bool dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? (*v != '0') : dflt;
}
static bool field_no_cache() { return env_flag("AVR_FIELD_NOCACHE", false); }
static bool field_bwd_split() { return env_flag("AVR_FIELD_BWD_SPLIT", true); }

template <int CPL>
static void launch_bwd_variant(const FieldInputsArgs& a, int row_stride, unsigned g, unsigned t, cudaStream_t stream) {
  const bool latent = a.d_latent != nullptr, point = a.d_xyz != nullptr || a.d_viewdirs != nullptr;
  if (latent && point && field_bwd_split()) {
    field_inputs_bwd_kernel<CPL, true, false><<<g, t, 0, stream>>>(a, row_stride);
    field_inputs_bwd_kernel<CPL, false, true><<<g, t, 0, stream>>>(a, row_stride);
  } else if (latent && point) {
    field_inputs_bwd_kernel<CPL, true, true><<<g, t, 0, stream>>>(a, row_stride);
  } else if (latent) {
    field_inputs_bwd_kernel<CPL, true, false><<<g, t, 0, stream>>>(a, row_stride);
  } else {
    field_inputs_bwd_kernel<CPL, false, true><<<g, t, 0, stream>>>(a, row_stride);
  }
}

int launch_field_inputs_bwd(const FieldInputsArgs& a, int64_t SB, cudaStream_t stream) {
  const int64_t rows = a.NV * a.B;
  if (!a.d_latent && !a.d_xyz && !a.d_viewdirs) return AVR_OK;
  cudaError_t e = cudaSuccess;
  if (a.d_latent) e = cudaMemsetAsync(a.d_latent, 0, sizeof(float) * a.NV * a.H * a.W * a.C, stream);
  if (e == cudaSuccess && a.d_xyz) e = cudaMemsetAsync(a.d_xyz, 0, sizeof(float) * SB * a.B * 3, stream);
  if (e == cudaSuccess && a.d_viewdirs) e = cudaMemsetAsync(a.d_viewdirs, 0, sizeof(float) * SB * a.B * 3, stream);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return AVR_ERR_LAUNCH;
  }
  if (rows == 0) return AVR_OK;
  const int width = a.features_only ? 0 : field_code_width(a);
  const int row_stride = a.C + width;
  const int64_t n_chunks = (rows + kFieldChunk - 1) / kFieldChunk;
  int64_t blocks = (n_chunks + kFieldWarps - 1) / kFieldWarps;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  const unsigned g = (unsigned)blocks, t = kFieldWarps * 32;
  switch (field_no_cache() ? 0 : a.C) {
    case 512: launch_bwd_variant<4>(a, row_stride, g, t, stream); break;
    case 256: launch_bwd_variant<2>(a, row_stride, g, t, stream); break;
    case 128: launch_bwd_variant<1>(a, row_stride, g, t, stream); break;
    default: launch_bwd_variant<0>(a, row_stride, g, t, stream); break;
  }
  return check_launch();
}

int launch_field_inputs_fwd(const FieldInputsArgs& a, cudaStream_t stream) {
  const int64_t rows = a.NV * a.B;
  if (rows == 0) return AVR_OK;
  const int width = a.features_only ? 0 : field_code_width(a);
  const int row_stride = a.C + width;
  const int64_t n_chunks = (rows + kFieldChunk - 1) / kFieldChunk;
  int64_t blocks = (n_chunks + kFieldWarps - 1) / kFieldWarps;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  const unsigned g = (unsigned)blocks, t = kFieldWarps * 32;
  switch (field_no_cache() ? 0 : a.C) {
    case 512: field_inputs_fwd_kernel<4><<<g, t, 0, stream>>>(a, row_stride); break;
    case 256: field_inputs_fwd_kernel<2><<<g, t, 0, stream>>>(a, row_stride); break;
    case 128: field_inputs_fwd_kernel<1><<<g, t, 0, stream>>>(a, row_stride); break;
    default: field_inputs_fwd_kernel<0><<<g, t, 0, stream>>>(a, row_stride); break;
  }
  return check_launch();
}

}  // namespace avr
