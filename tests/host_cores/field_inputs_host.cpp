// Host walk of the radiance-field front end's per-lane code (csrc/field_inputs_core.h), driven
// like field_inputs_fwd_kernel drives it: chunks of consecutive rows per warp, 32 lanes per row,
// the tap cache carried from row to row.  Test infrastructure only.
#include <stdint.h>
#include <string.h>

#include "field_inputs_core.h"

namespace {

template <int CPL>
void walk_fwd(const avr::FieldInputsArgs& a, int row_stride, int chunk, int n_warps) {
  constexpr int N = CPL > 0 ? CPL : 1;
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + chunk - 1) / chunk;
  for (int warp = 0; warp < n_warps; ++warp) {
    // per-lane state that lives in registers on the device
    avr::FieldTapCache<N> cache[32];
    avr::FieldView view[32];
    avr::FieldLaneCode lc[32];
    for (int l = 0; l < 32; ++l) {
      avr::field_cache_reset(&cache[l]);
      avr::field_view_reset(&view[l]);
      lc[l] = avr::field_lane_code(a, l);
    }
    for (int64_t ch = warp; ch < n_chunks; ch += n_warps) {
      const int64_t first = ch * chunk;
      const int n = (int)(first + chunk < rows ? chunk : rows - first);
      for (int lane = 0; lane < 32; ++lane) {
        avr::FieldCursor cur = avr::field_cursor_at(a, first);
        for (int r = 0; r < n; ++r, avr::field_cursor_next(a, &cur)) {
          avr::field_view_fill(a, cur, &view[lane]);
          const avr::FieldPoint p = avr::field_point(a, cur, view[lane]);
          float* out = a.out + cur.row * row_stride;
          if (CPL > 0) {
            avr::field_row_lane<N, true>(a, cur, p, lane, out, lc[lane], &cache[lane]);
          } else {
            avr::field_row_lane_generic<true>(a, cur, p, lane, out, lc[lane]);
          }
        }
      }
    }
  }
}

template <int CPL, bool kLatent, bool kPoint>
void walk_bwd(const avr::FieldInputsArgs& a, int row_stride, int chunk, int n_warps, bool preloaded) {
  constexpr int N = CPL > 0 ? CPL : 1;
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + chunk - 1) / chunk;
  for (int warp = 0; warp < n_warps; ++warp) {
    avr::FieldTapCache<N> taps[32];
    avr::FieldGradCache<N> grads[32];
    avr::FieldView view[32];
    avr::FieldLaneCode lc[32];
    for (int l = 0; l < 32; ++l) {
      avr::field_cache_reset(&taps[l]);
      avr::field_grad_reset(&grads[l]);
      avr::field_view_reset(&view[l]);
      lc[l] = avr::field_lane_code(a, l);
    }
    for (int64_t ch = warp; ch < n_chunks; ch += n_warps) {
      const int64_t first = ch * chunk;
      const int n = (int)(first + chunk < rows ? chunk : rows - first);
      avr::FieldCursor cur = avr::field_cursor_at(a, first);
      for (int r = 0; r < n; ++r, avr::field_cursor_next(a, &cur)) {
        avr::FieldRowPartial sum;
        avr::field_partial_zero(&sum);
        avr::FieldPoint p0;
        for (int lane = 0; lane < 32; ++lane) {
          avr::field_view_fill(a, cur, &view[lane]);
          const avr::FieldPoint p = avr::field_point(a, cur, view[lane]);
          if (lane == 0) p0 = p;
          avr::FieldRowPartial s;
          if (CPL > 0) {
            avr::FieldRowGrad<N> rg;
            if (preloaded) {
              avr::field_load_row_grad<N>(a, cur.row, lane, row_stride, &rg);
              s = avr::field_bwd_row_lane<N, kLatent, kPoint, true>(a, cur, p, lane, row_stride, lc[lane], rg, &taps[lane], &grads[lane]);
            } else {
              s = avr::field_bwd_row_lane<N, kLatent, kPoint, false>(a, cur, p, lane, row_stride, lc[lane], rg, &taps[lane], &grads[lane]);
            }
          } else {
            s = avr::field_bwd_row_lane_generic<kLatent, kPoint>(a, cur, p, lane, row_stride, lc[lane]);
          }
          sum.gix += s.gix;
          sum.giy += s.giy;
          sum.enc0 += s.enc0;
          sum.enc1 += s.enc1;
          sum.enc2 += s.enc2;
          sum.vr0 += s.vr0;
          sum.vr1 += s.vr1;
          sum.vr2 += s.vr2;
        }
        if (kPoint) avr::field_bwd_row_finish(a, cur, view[0], p0, sum);  // lane 0 on the device
      }
    }
    if (kLatent && CPL > 0)
      for (int lane = 0; lane < 32; ++lane) avr::field_grad_flush<N>(a, lane, &grads[lane]);
  }
}

template <int CPL>
void walk_bwd_variant(const avr::FieldInputsArgs& a, int row_stride, int chunk, int n_warps, bool preloaded) {
  const bool latent = a.d_latent != nullptr, point = a.d_xyz != nullptr || a.d_viewdirs != nullptr;
  if (latent && point) {
    walk_bwd<CPL, true, true>(a, row_stride, chunk, n_warps, preloaded);
  } else if (latent) {
    walk_bwd<CPL, true, false>(a, row_stride, chunk, n_warps, preloaded);
  } else if (point) {
    walk_bwd<CPL, false, true>(a, row_stride, chunk, n_warps, preloaded);
  }
}

}  // namespace

extern "C" {

// Backward walk; the caller zeroes d_latent / d_xyz / d_viewdirs (the library's launcher does).
// use_cache: 0 generic walk, 1 register caches, 3 register caches + the row's gradient loaded ahead
int host_field_inputs_bwd(const avr::FieldInputsArgs* a, int use_cache, int chunk, int n_warps) {
  const bool preloaded = (use_cache & 2) != 0;
  const int width = a->features_only ? 0 : avr::field_code_width(*a);
  const int row_stride = a->C + width;
  if (a->C % 4 != 0 || (row_stride & 1)) return -1;
  const int cpl = (use_cache && a->C % 128 == 0) ? a->C / 128 : 0;
  switch (cpl) {
    case 4: walk_bwd_variant<4>(*a, row_stride, chunk, n_warps, preloaded); break;
    case 2: walk_bwd_variant<2>(*a, row_stride, chunk, n_warps, preloaded); break;
    case 1: walk_bwd_variant<1>(*a, row_stride, chunk, n_warps, preloaded); break;
    default: walk_bwd_variant<0>(*a, row_stride, chunk, n_warps, preloaded); break;
  }
  return 0;
}


// `a` is the launch descriptor exactly as the library fills it.  use_cache = 0 forces the generic
// (no tap cache) walk.  n_warps emulates the grid: a warp visits chunks warp, warp + n_warps, ...
int host_field_inputs_fwd(const avr::FieldInputsArgs* a, int use_cache, int chunk, int n_warps) {
  const int width = a->features_only ? 0 : avr::field_code_width(*a);
  const int row_stride = a->C + width;
  if (a->C % 4 != 0 || (row_stride & 1)) return -1;
  const int cpl = (use_cache && a->C % 128 == 0) ? a->C / 128 : 0;
  switch (cpl) {
    case 4: walk_fwd<4>(*a, row_stride, chunk, n_warps); break;
    case 2: walk_fwd<2>(*a, row_stride, chunk, n_warps); break;
    case 1: walk_fwd<1>(*a, row_stride, chunk, n_warps); break;
    default: walk_fwd<0>(*a, row_stride, chunk, n_warps); break;
  }
  return 0;
}

int host_field_args_size() { return (int)sizeof(avr::FieldInputsArgs); }

}  // extern "C"
