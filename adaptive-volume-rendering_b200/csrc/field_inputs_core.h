// Radiance-field front end: what NewPixelNeRFNet.forward does
// between receiving the renderer's sample points and calling its MLP (models.py:754-820) —
//   world -> source-view transform        xyz_rot = R p,  xyz = xyz_rot + t            :757-761
//   positional encoding                   [x, sin(phase_k + x f_k)] of xyz_rot or xyz  :764-781, :45-71
//   view direction into the view frame    R d                                            :783-794
//   projection + pixel-aligned features   uv = -xy/z * focal + c; grid_sample(latent)    :803-815, :256-279
//   concatenation                         mlp_input = [latent | code | viewdir]          :826
// as per-lane code that compiles for the device (nvcc) and for the host (g++; the CPU suite walks
// it lane by lane against the oracle: tests/test_host_kernel_cores.py).
//
// One warp per (source view, point) = one output row; rows are walked in runs of consecutive
// rows (FieldCursor).  The lane-independent coordinate work of a row is field_point(); the kernels
// either let every lane do it for itself or do it once per row and hand it round by shuffles
// (field_inputs.cu) — the functions below take the finished FieldPoint either way.  A lane then
// owns channels {4*lane + 128*i} of the feature vector and code entries {lane, lane+32}.  The
// feature map is channels-last (N, H, W, C): one bilinear tap is one contiguous C*4-byte row.
//
// The operation order reproduces torch-CPU's kernels so that everything except the sine is
// bit-identical to the reference on the same inputs (probed against the reference, see
// oracle/make_golden.py case_field_inputs): the 3x3 product accumulates left to right without
// contraction (bmm's small-matrix loop), the encoding argument is ONE fused multiply-add
// (addcmul), pixel coordinates follow grid_sampler's unnormalize / clip / floor with
// align_corners=True and border padding, and the four taps accumulate nw, ne, sw, se through
// fused multiply-adds.
#pragma once

#include <stdint.h>

#include "avr_b200.h"

#if defined(__CUDACC__)
#define AVR_FI __device__ __forceinline__
#else
#include <math.h>
#define AVR_FI inline
#endif

namespace avr {

#if defined(__CUDACC__)
AVR_FI float fi_mul(float a, float b) { return __fmul_rn(a, b); }
AVR_FI float fi_add(float a, float b) { return __fadd_rn(a, b); }
AVR_FI float fi_sub(float a, float b) { return __fsub_rn(a, b); }
AVR_FI float fi_div(float a, float b) { return __fdiv_rn(a, b); }
AVR_FI float fi_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
AVR_FI float fi_mul(float a, float b) { return a * b; }
AVR_FI float fi_add(float a, float b) { return a + b; }
AVR_FI float fi_sub(float a, float b) { return a - b; }
AVR_FI float fi_div(float a, float b) { return a / b; }
AVR_FI float fi_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

// Launch-constant description of the front end (the public descriptor of include/avr_b200.h,
// passed by value to the kernels).
using FieldInputsArgs = ::avr_field_inputs;

#if defined(__CUDACC__)
__host__ __device__
#endif
inline int field_code_width(const FieldInputsArgs& a) {
  return (a.include_input ? 3 : 0) + 3 * a.n_sin + (a.use_viewdirs ? 3 : 0);
}

// Where a walk stands: row = v*B + b of the output, v = obj*NS + s.  Rows are visited in runs of
// consecutive rows, so the (64-bit) divisions happen once per run, not once per row.
struct FieldCursor {
  int64_t row, v, b, obj;
};

AVR_FI FieldCursor field_cursor_at(const FieldInputsArgs& a, int64_t row) {
  FieldCursor c;
  c.row = row;
  c.v = row / a.B;
  c.b = row - c.v * a.B;
  c.obj = c.v / a.NS;
  return c;
}

AVR_FI void field_cursor_next(const FieldInputsArgs& a, FieldCursor* c) {
  ++c->row;
  if (++c->b == a.B) {
    c->b = 0;
    ++c->v;
    c->obj = c->v / a.NS;
  }
}

// Per-view constants, kept in registers while consecutive rows belong to one source view.
struct FieldView {
  float R[12];          // world -> view, rows of 4 (rotation | translation)
  float fx, fy, cx, cy;
  int64_t v;
};

AVR_FI void field_view_reset(FieldView* w) { w->v = -1; }

AVR_FI void field_view_fill(const FieldInputsArgs& a, const FieldCursor& cur, FieldView* w) {
  if (w->v == cur.v) return;  // warp-uniform
  w->v = cur.v;
  const float* R = a.poses + cur.v * 12;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < 12; ++i) w->R[i] = R[i];
  const float* f = a.focal + (a.focal_per_obj ? cur.obj * 2 : 0);
  const float* c = a.c + (a.c_per_obj ? cur.obj * 2 : 0);
  w->fx = f[0];
  w->fy = f[1];
  w->cx = c[0];
  w->cy = c[1];
}

// Coordinates of one (view, point): everything that does not depend on the lane.
struct FieldPoint {
  float enc0, enc1, enc2;     // what the positional code encodes (R p, or R p + t)
  float vrot0, vrot1, vrot2;  // R d
  float cam0, cam1, cam2;     // R p + t
  float ix, iy;               // clipped pixel coordinates in the feature map
  int x0, y0;                 // floor
  float nw, ne, sw, se;
  bool clip_x, clip_y;        // coordinate was clamped (its gradient is zero)
};

// row i of R times (x, y, z): (R[i][0]*x + R[i][1]*y) + R[i][2]*z, products and sums rounded one
// by one (torch-CPU's small-matrix bmm loop)
AVR_FI float rot_row(const float* R, int i, float x, float y, float z) {
  return fi_add(fi_add(fi_mul(R[4 * i], x), fi_mul(R[4 * i + 1], y)), fi_mul(R[4 * i + 2], z));
}

// The point (x, y, z) with view direction (dx, dy, dz) seen from view `w` (the LSTM march,
// lstm_march.cu, calls this with points it holds in registers).
AVR_FI FieldPoint field_point_xyz(const FieldInputsArgs& a, const FieldView& w, float x, float y, float z, float dx, float dy,
                                  float dz) {
  FieldPoint p;
  const float r0 = rot_row(w.R, 0, x, y, z), r1 = rot_row(w.R, 1, x, y, z), r2 = rot_row(w.R, 2, x, y, z);
  p.cam0 = fi_add(r0, w.R[3]);
  p.cam1 = fi_add(r1, w.R[7]);
  p.cam2 = fi_add(r2, w.R[11]);
  p.enc0 = a.normalize_z ? r0 : p.cam0;
  p.enc1 = a.normalize_z ? r1 : p.cam1;
  p.enc2 = a.normalize_z ? r2 : p.cam2;
  p.vrot0 = p.vrot1 = p.vrot2 = 0.f;
  if (a.use_viewdirs && !a.features_only) {
    p.vrot0 = rot_row(w.R, 0, dx, dy, dz);
    p.vrot1 = rot_row(w.R, 1, dx, dy, dz);
    p.vrot2 = rot_row(w.R, 2, dx, dy, dz);
  }
  // uv = -xy / z; uv *= focal; uv += c; uv = uv * scale - 1          models.py:803-811, :268-270
  const float u = fi_sub(fi_mul(fi_add(fi_mul(fi_div(-p.cam0, p.cam2), w.fx), w.cx), a.scale_x), 1.0f);
  const float t = fi_sub(fi_mul(fi_add(fi_mul(fi_div(-p.cam1, p.cam2), w.fy), w.cy), a.scale_y), 1.0f);
  // grid_sample, align_corners=True, padding_mode="border": x = (u + 1) * ((W-1)/2), clamped
  const float mx = (float)(a.W - 1), my = (float)(a.H - 1);
  const float ux = fi_mul(fi_add(u, 1.0f), mx / 2), uy = fi_mul(fi_add(t, 1.0f), my / 2);
  p.clip_x = !(ux > 0.f && ux < mx);  // grid_sampler's clip has zero gradient on and beyond the border
  p.clip_y = !(uy > 0.f && uy < my);
  p.ix = fminf(mx, fmaxf(ux, 0.f));
  p.iy = fminf(my, fmaxf(uy, 0.f));
  const float fx = floorf(p.ix), fy = floorf(p.iy);
  p.x0 = (int)fx;
  p.y0 = (int)fy;
  const float wx = fi_sub(p.ix, fx), ex = fi_sub(1.0f, wx);
  const float wy = fi_sub(p.iy, fy), sy = fi_sub(1.0f, wy);
  p.nw = fi_mul(sy, ex);
  p.ne = fi_mul(sy, wx);
  p.sw = fi_mul(wy, ex);
  p.se = fi_mul(wy, wx);
  return p;
}

AVR_FI FieldPoint field_point(const FieldInputsArgs& a, const FieldCursor& cur, const FieldView& w) {
  const float* q = a.xyz + (cur.obj * a.B + cur.b) * 3;
  float dx = 0.f, dy = 0.f, dz = 0.f;
  if (a.use_viewdirs && !a.features_only) {
    const float* d = a.viewdirs + (cur.obj * a.B + cur.b) * 3;
    dx = d[0];
    dy = d[1];
    dz = d[2];
  }
  return field_point_xyz(a, w, q[0], q[1], q[2], dx, dy, dz);
}

AVR_FI float pick3(float v0, float v1, float v2, int d) { return d == 0 ? v0 : (d == 1 ? v1 : v2); }

// The code part of a row is [enc (3) | sin rows (3 per row) | view direction (3)]; lane l owns
// entries l and l + 32.  What those entries are does not change from row to row:
struct FieldLaneCode {
  int kind[2];   // 0: none, 1: enc[d], 2: sin(enc[d] * freq + phase), 3: vrot[d]
  int d[2];
  float freq[2], phase[2];
};

AVR_FI FieldLaneCode field_lane_code(const FieldInputsArgs& a, int lane) {
  FieldLaneCode lc;
  const int width = field_code_width(a);
  for (int s = 0; s < 2; ++s) {
    int e = lane + 32 * s;
    lc.kind[s] = 0;
    lc.d[s] = 0;
    lc.freq[s] = lc.phase[s] = 0.f;
    if (e >= width || a.features_only) continue;
    if (a.include_input) {
      if (e < 3) {
        lc.kind[s] = 1;
        lc.d[s] = e;
        continue;
      }
      e -= 3;
    }
    if (e < 3 * a.n_sin) {
      const int k = e / 3;
      lc.kind[s] = 2;
      lc.d[s] = e - 3 * k;
      lc.freq[s] = a.freqs[k];
      lc.phase[s] = a.phases[k];
    } else {
      lc.kind[s] = 3;
      lc.d[s] = e - 3 * a.n_sin;
    }
  }
  return lc;
}

// slot s of this lane's code entries; addcmul(phases, x, freqs) is one fused multiply-add, then sin (:66-67)
AVR_FI float field_code_value(const FieldLaneCode& lc, int s, const FieldPoint& p) {
  const float x = pick3(p.enc0, p.enc1, p.enc2, lc.d[s]);
  if (lc.kind[s] == 2) return sinf(fi_fma(x, lc.freq[s], lc.phase[s]));
  if (lc.kind[s] == 1) return x;
  return pick3(p.vrot0, p.vrot1, p.vrot2, lc.d[s]);
}

// One bilinear tap row: address of channel 0 of texel (x, y) of view v, or null outside the map
// (grid_sample reads such a tap as zero; with border padding its weight is zero as well).
AVR_FI const float* field_tap(const FieldInputsArgs& a, int64_t v, int x, int y) {
  if (x < 0 || y < 0 || x >= a.W || y >= a.H) return nullptr;
  return a.latent + ((v * a.H + y) * (int64_t)a.W + x) * a.C;
}

// nw, ne, sw, se accumulated by fused multiply-adds, in that order
AVR_FI float field_blend(const FieldPoint& p, float t_nw, float t_ne, float t_sw, float t_se) {
  float acc = fi_mul(t_nw, p.nw);
  acc = fi_fma(t_ne, p.ne, acc);
  acc = fi_fma(t_sw, p.sw, acc);
  return fi_fma(t_se, p.se, acc);
}

// ---- forward: one lane's share of one output row ---------------------------------------------
// CPL = float4 groups per lane and tap (C == 128 * CPL).  The four taps stay in registers from
// one point to the next: consecutive samples of a ray usually fall into the same texel cell, so
// the feature rows are re-read only when the cell (or the view) changes.
template <int CPL>
struct FieldTapCache {
  float t[4][CPL][4];  // [nw, ne, sw, se][group][component]
  int x0, y0;
  int64_t v;
};

template <int CPL>
AVR_FI void field_cache_reset(FieldTapCache<CPL>* c) {
  c->x0 = c->y0 = -1;
  c->v = -1;
}

AVR_FI void field_load4(const float* src, float* dst) {
#if defined(__CUDACC__)
  const float4 q = __ldg(reinterpret_cast<const float4*>(src));
  dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
#else
  for (int i = 0; i < 4; ++i) dst[i] = src[i];
#endif
}

// kStream: dst is global memory, written once (streaming hint); otherwise any address space
template <bool kStream>
AVR_FI void field_store2(float* dst, float x, float y) {
#if defined(__CUDACC__)
  if (kStream) {
    __stcs(reinterpret_cast<float2*>(dst), make_float2(x, y));
  } else {
    *reinterpret_cast<float2*>(dst) = make_float2(x, y);
  }
#else
  dst[0] = x;
  dst[1] = y;
#endif
}

// returns true when the cache was (re)filled
template <int CPL>
AVR_FI bool field_cache_fill(const FieldInputsArgs& a, const FieldPoint& p, int64_t v, int lane, FieldTapCache<CPL>* c) {
  if (c->x0 == p.x0 && c->y0 == p.y0 && c->v == v) return false;  // warp-uniform
  c->x0 = p.x0;
  c->y0 = p.y0;
  c->v = v;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    const float* row = field_tap(a, v, p.x0 + (k & 1), p.y0 + (k >> 1));
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 0; i < CPL; ++i) {
      if (row) {
        field_load4(row + 4 * lane + 128 * i, c->t[k][i]);
      } else {
        c->t[k][i][0] = c->t[k][i][1] = c->t[k][i][2] = c->t[k][i][3] = 0.f;
      }
    }
  }
  return true;
}

// The point gradient only needs the four tap DIFFERENCES of a cell,
//   t[0] = ne - nw (along x, north edge), t[1] = se - sw (along x, south edge),
//   t[2] = sw - nw (along y, west edge),  t[3] = se - ne (along y, east edge),
// formed once per cell change instead of once per row (consecutive samples of a ray stay in a cell for
// three rows out of four): the backward kernels keep those in the same registers.
template <int CPL>
AVR_FI void field_diff_fill(const FieldInputsArgs& a, const FieldPoint& p, int64_t v, int lane, FieldTapCache<CPL>* c) {
  if (!field_cache_fill<CPL>(a, p, v, lane, c)) return;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < CPL; ++i) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) {
      const float nw = c->t[0][i][q], ne = c->t[1][i][q], sw = c->t[2][i][q], se = c->t[3][i][q];
      c->t[0][i][q] = ne - nw;
      c->t[1][i][q] = se - sw;
      c->t[2][i][q] = sw - nw;
      c->t[3][i][q] = se - ne;
    }
  }
}

// One row of the output, lane `lane`, written to `out` — the row's place in global memory
// (kStream) or in a shared-memory staging buffer that a bulk copy takes to global memory.  Rows are
// 8-byte aligned (the stride is even, checked by the launcher), not 16: 554 floats per row.
template <int CPL, bool kStream>
AVR_FI void field_row_lane(const FieldInputsArgs& a, const FieldCursor& cur, const FieldPoint& p, int lane, float* out,
                           const FieldLaneCode& lc, FieldTapCache<CPL>* c) {
  field_cache_fill<CPL>(a, p, cur.v, lane, c);
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < CPL; ++i) {
    float o[4];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) o[q] = field_blend(p, c->t[0][i][q], c->t[1][i][q], c->t[2][i][q], c->t[3][i][q]);
    field_store2<kStream>(out + 4 * lane + 128 * i, o[0], o[1]);
    field_store2<kStream>(out + 4 * lane + 128 * i + 2, o[2], o[3]);
  }
  if (lc.kind[0]) out[a.C + lane] = field_code_value(lc, 0, p);
  if (lc.kind[1]) out[a.C + lane + 32] = field_code_value(lc, 1, p);
}

// Any C % 4 == 0 (no tap cache): lanes stride over the float4 groups of the row.
template <bool kStream>
AVR_FI void field_row_lane_generic(const FieldInputsArgs& a, const FieldCursor& cur, const FieldPoint& p, int lane,
                                   float* out, const FieldLaneCode& lc) {
  const float* r0 = field_tap(a, cur.v, p.x0, p.y0);
  const float* r1 = field_tap(a, cur.v, p.x0 + 1, p.y0);
  const float* r2 = field_tap(a, cur.v, p.x0, p.y0 + 1);
  const float* r3 = field_tap(a, cur.v, p.x0 + 1, p.y0 + 1);
  for (int g = lane; g < a.C / 4; g += 32) {
    float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f}, t3[4] = {0.f, 0.f, 0.f, 0.f};
    if (r0) field_load4(r0 + 4 * g, t0);
    if (r1) field_load4(r1 + 4 * g, t1);
    if (r2) field_load4(r2 + 4 * g, t2);
    if (r3) field_load4(r3 + 4 * g, t3);
    float o[4];
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) o[q] = field_blend(p, t0[q], t1[q], t2[q], t3[q]);
    field_store2<kStream>(out + 4 * g, o[0], o[1]);
    field_store2<kStream>(out + 4 * g + 2, o[2], o[3]);
  }
  if (lc.kind[0]) out[a.C + lane] = field_code_value(lc, 0, p);
  if (lc.kind[1]) out[a.C + lane + 32] = field_code_value(lc, 1, p);
}

// ---- backward --------------------------------------------------------------------------------
// Same walk as the forward pass (a warp visits chunks of consecutive rows).  For row (v, b) with
// upstream gradient g (row of g_out):
//   d_latent[tap_k] += g_c * weight_k                     (grid_sample's backward w.r.t. its input)
//   d ix = sum_c g_c [(t_ne - t_nw) sy + (t_se - t_sw) wy],  d iy likewise   (w.r.t. its grid)
//   d enc = g_code (raw entries) + sum_k g_k cos(arg_k) f_k;  d vrot = g_view
//   d p = R^T (d cam + d enc),  d viewdir = R^T d vrot;  accumulated over the NS views of an
//   object by atomic adds (d_xyz / d_viewdirs are zeroed by the launcher).
// The d_latent contributions are summed in registers for as long as consecutive rows stay in one
// texel cell and flushed with vector atomics when the cell changes: samples along a ray would
// otherwise queue on the same four addresses in L2.
template <int CPL>
struct FieldGradCache {
  float acc[4][CPL][4];
  int x0, y0;
  int64_t v;
};

template <int CPL>
AVR_FI void field_grad_reset(FieldGradCache<CPL>* c) {
  c->x0 = c->y0 = -1;
  c->v = -1;
}

AVR_FI void field_atomic_add4(float* dst, const float* v) {
#if defined(__CUDACC__)
  // a reduction (REDG: fire and forget), not an atomic with a discarded result (ATOM ... RZ still waits for L2's answer)
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
#else
  for (int i = 0; i < 4; ++i) dst[i] += v[i];
#endif
}

AVR_FI void field_atomic_add(float* dst, float v) {
#if defined(__CUDACC__)
  atomicAdd(dst, v);
#else
  *dst += v;
#endif
}

AVR_FI void field_load2(const float* src, float* dst) {
#if defined(__CUDACC__)
  const float2 q = __ldcs(reinterpret_cast<const float2*>(src));
  dst[0] = q.x; dst[1] = q.y;
#else
  dst[0] = src[0];
  dst[1] = src[1];
#endif
}

// address of texel (x, y) of view v in d_latent (same layout as latent), or null outside the map
AVR_FI float* field_grad_tap(const FieldInputsArgs& a, int64_t v, int x, int y) {
  if (x < 0 || y < 0 || x >= a.W || y >= a.H) return nullptr;
  return a.d_latent + ((v * a.H + y) * (int64_t)a.W + x) * a.C;
}

template <int CPL>
AVR_FI void field_grad_flush(const FieldInputsArgs& a, int lane, FieldGradCache<CPL>* c) {
  if (c->v < 0) return;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    float* row = field_grad_tap(a, c->v, c->x0 + (k & 1), c->y0 + (k >> 1));
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 0; i < CPL; ++i)
      if (row) field_atomic_add4(row + 4 * lane + 128 * i, c->acc[k][i]);
  }
  c->v = -1;
}

// Per-lane partial sums of one row that the warp has to add up.
struct FieldRowPartial {
  float gix, giy;           // d / d(ix, iy) over this lane's channels
  float enc0, enc1, enc2;   // d / d enc over this lane's code entries
  float vr0, vr1, vr2;      // d / d vrot
};

AVR_FI void field_partial_zero(FieldRowPartial* s) {
  s->gix = s->giy = 0.f;
  s->enc0 = s->enc1 = s->enc2 = 0.f;
  s->vr0 = s->vr1 = s->vr2 = 0.f;
}

// this lane's two code entries: the mirror of field_code_value
AVR_FI void field_code_grad_lane(const FieldLaneCode& lc, const FieldPoint& p, const float* g_code, int lane,
                                 FieldRowPartial* s) {
  for (int slot = 0; slot < 2; ++slot) {
    const int kind = lc.kind[slot], d = lc.d[slot];
    if (kind == 0) continue;
    const float g = g_code[lane + 32 * slot];
    if (kind == 3) {
      s->vr0 += d == 0 ? g : 0.f;
      s->vr1 += d == 1 ? g : 0.f;
      s->vr2 += d == 2 ? g : 0.f;
      continue;
    }
    float t = g;
    if (kind == 2) t = g * cosf(fi_fma(pick3(p.enc0, p.enc1, p.enc2, d), lc.freq[slot], lc.phase[slot])) * lc.freq[slot];
    s->enc0 += d == 0 ? t : 0.f;
    s->enc1 += d == 1 ? t : 0.f;
    s->enc2 += d == 2 ? t : 0.f;
  }
}

// This lane's share of a row of g_out (the latent part): loaded ahead of its use by the kernels
// that prefetch the next row while the current one is processed.
template <int CPL>
struct FieldRowGrad {
  float g[CPL][4];
};

template <int CPL>
AVR_FI void field_load_row_grad(const FieldInputsArgs& a, int64_t row, int lane, int row_stride, FieldRowGrad<CPL>* rg) {
  const float* g_row = a.g_out + row * row_stride;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < CPL; ++i) {
    field_load2(g_row + 4 * lane + 128 * i, rg->g[i]);
    field_load2(g_row + 4 * lane + 128 * i + 2, rg->g[i] + 2);
  }
}

// Lane `lane` of one row: accumulates into the d_latent cache (kLatent) and returns its partial
// sums of the point gradient (kPoint; zero otherwise).  `taps` is the forward tap cache (read only
// when kPoint); with kPreloaded `rg` holds this lane's part of the row's upstream gradient,
// otherwise it is read here, group by group.
// `g_row_in` (may be null): where the row of g_out is to be read from instead of global memory — the
// shared-memory ring of the kernels that prefetch it with bulk copies (any address space).
// kStaged16: `g_row_in` is a staged copy whose channel groups are 16-byte aligned (one LDS.128 per group).
template <int CPL, bool kLatent, bool kPoint, bool kPreloaded, bool kStaged16 = false>
AVR_FI FieldRowPartial field_bwd_row_lane(const FieldInputsArgs& a, const FieldCursor& cur, const FieldPoint& p, int lane,
                                          int row_stride, const FieldLaneCode& lc, const FieldRowGrad<CPL>& rg,
                                          FieldTapCache<CPL>* taps, FieldGradCache<CPL>* grads,
                                          const float* g_row_in = nullptr) {
  const int64_t v = cur.v;
  FieldRowPartial s;
  field_partial_zero(&s);
  const float* g_row = g_row_in ? g_row_in : a.g_out + cur.row * row_stride;
  if (kPoint) field_diff_fill<CPL>(a, p, v, lane, taps);  // `taps` holds the cell's tap differences here
  if (kLatent && !(grads->x0 == p.x0 && grads->y0 == p.y0 && grads->v == v)) {  // warp-uniform
    field_grad_flush<CPL>(a, lane, grads);
    grads->x0 = p.x0;
    grads->y0 = p.y0;
    grads->v = v;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k)
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int i = 0; i < CPL; ++i) grads->acc[k][i][0] = grads->acc[k][i][1] = grads->acc[k][i][2] = grads->acc[k][i][3] = 0.f;
  }
  // bilinear weights again, split into their factors for the coordinate gradient
  const float wx = fi_sub(p.ix, (float)p.x0), ex = fi_sub(1.0f, wx);
  const float wy = fi_sub(p.iy, (float)p.y0), sy = fi_sub(1.0f, wy);
  float dxn = 0.f, dxs = 0.f, dyw = 0.f, dye = 0.f;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < CPL; ++i) {
    float g[4];
    if (kPreloaded) {
      g[0] = rg.g[i][0]; g[1] = rg.g[i][1]; g[2] = rg.g[i][2]; g[3] = rg.g[i][3];
    } else if (kStaged16) {
#if defined(__CUDACC__)
      const float4 gv = *reinterpret_cast<const float4*>(g_row + 4 * lane + 128 * i);
      g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
#else
      for (int q = 0; q < 4; ++q) g[q] = g_row[4 * lane + 128 * i + q];
#endif
    } else if (g_row_in) {  // staged row (shared memory): ordinary 8-byte loads, one group at a time
      g[0] = g_row[4 * lane + 128 * i];
      g[1] = g_row[4 * lane + 128 * i + 1];
      g[2] = g_row[4 * lane + 128 * i + 2];
      g[3] = g_row[4 * lane + 128 * i + 3];
    } else {  // `rg` is not touched: one group of four in flight at a time
      field_load2(g_row + 4 * lane + 128 * i, g);
      field_load2(g_row + 4 * lane + 128 * i + 2, g + 2);
    }
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 0; q < 4; ++q) {
      if (kLatent) {
        grads->acc[0][i][q] = fi_fma(g[q], p.nw, grads->acc[0][i][q]);
        grads->acc[1][i][q] = fi_fma(g[q], p.ne, grads->acc[1][i][q]);
        grads->acc[2][i][q] = fi_fma(g[q], p.sw, grads->acc[2][i][q]);
        grads->acc[3][i][q] = fi_fma(g[q], p.se, grads->acc[3][i][q]);
      }
      if (kPoint) {  // four sums over the channels; the row's bilinear factors multiply them once, below
        dxn += g[q] * taps->t[0][i][q];
        dxs += g[q] * taps->t[1][i][q];
        dyw += g[q] * taps->t[2][i][q];
        dye += g[q] * taps->t[3][i][q];
      }
    }
  }
  if (kPoint) {
    s.gix = dxn * sy + dxs * wy;
    s.giy = dyw * ex + dye * wx;
  }
  if (kPoint) field_code_grad_lane(lc, p, g_row + a.C, lane, &s);
  return s;
}

// Once per row, with the partials summed over the warp: chain rule through the projection and
// the rigid transform, then atomic accumulation over the object's views.
AVR_FI void field_bwd_row_finish(const FieldInputsArgs& a, const FieldCursor& cur, const FieldView& w, const FieldPoint& p,
                                 const FieldRowPartial& s) {
  // ix = clip((u + 1) (W-1)/2): zero gradient where clamped; u = (-x/z * fx + cx) * scale_x - 1
  const float gu = p.clip_x ? 0.f : s.gix * ((float)(a.W - 1) / 2) * a.scale_x * w.fx;
  const float gt = p.clip_y ? 0.f : s.giy * ((float)(a.H - 1) / 2) * a.scale_y * w.fy;
  const float iz = 1.0f / p.cam2;
  // enc is R p (+ t): same Jacobian as cam
  const float d0 = -gu * iz + s.enc0;
  const float d1 = -gt * iz + s.enc1;
  const float d2 = (gu * p.cam0 + gt * p.cam1) * iz * iz + s.enc2;
  if (a.d_xyz) {
    float* dst = a.d_xyz + (cur.obj * a.B + cur.b) * 3;
    for (int j = 0; j < 3; ++j) field_atomic_add(dst + j, w.R[j] * d0 + w.R[4 + j] * d1 + w.R[8 + j] * d2);
  }
  if (a.d_viewdirs && a.use_viewdirs && !a.features_only) {
    float* dst = a.d_viewdirs + (cur.obj * a.B + cur.b) * 3;
    for (int j = 0; j < 3; ++j) field_atomic_add(dst + j, w.R[j] * s.vr0 + w.R[4 + j] * s.vr1 + w.R[8 + j] * s.vr2);
  }
}

// Any C % 4 == 0: no caches, every row adds straight into d_latent.
template <bool kLatent, bool kPoint>
AVR_FI FieldRowPartial field_bwd_row_lane_generic(const FieldInputsArgs& a, const FieldCursor& cur, const FieldPoint& p,
                                                  int lane, int row_stride, const FieldLaneCode& lc) {
  const int64_t v = cur.v;
  FieldRowPartial s;
  field_partial_zero(&s);
  const float* g_row = a.g_out + cur.row * row_stride;
  const float wx = fi_sub(p.ix, (float)p.x0), ex = fi_sub(1.0f, wx);
  const float wy = fi_sub(p.iy, (float)p.y0), sy = fi_sub(1.0f, wy);
  for (int grp = lane; grp < a.C / 4; grp += 32) {
    float g[4], t[4][4];
    field_load2(g_row + 4 * grp, g);
    field_load2(g_row + 4 * grp + 2, g + 2);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
      const int x = p.x0 + (k & 1), y = p.y0 + (k >> 1);
      const float wk = k == 0 ? p.nw : (k == 1 ? p.ne : (k == 2 ? p.sw : p.se));
      if (kLatent) {
        float* dst = field_grad_tap(a, v, x, y);
        if (dst) {
          const float c[4] = {g[0] * wk, g[1] * wk, g[2] * wk, g[3] * wk};
          field_atomic_add4(dst + 4 * grp, c);
        }
      }
      if (kPoint) {
        const float* src = field_tap(a, v, x, y);
        if (src) {
          field_load4(src + 4 * grp, t[k]);
        } else {
          t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0.f;
        }
      }
    }
    if (kPoint) {
#if defined(__CUDACC__)
#pragma unroll
#endif
      for (int q = 0; q < 4; ++q) {
        s.gix += g[q] * ((t[1][q] - t[0][q]) * sy + (t[3][q] - t[2][q]) * wy);
        s.giy += g[q] * ((t[2][q] - t[0][q]) * ex + (t[3][q] - t[1][q]) * wx);
      }
    }
  }
  if (kPoint) field_code_grad_lane(lc, p, g_row + a.C, lane, &s);
  return s;
}

}  // namespace avr
