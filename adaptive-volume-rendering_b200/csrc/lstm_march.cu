// AdaptiveVolumeRenderer's LSTM ray march (renderers.py:411-435) as
// ONE persistent launch, and its backward through time as one more.
//
// The reference marches every ray `steps` times: fetch pixel-aligned encoder features at the
// current point (phi(..., return_features=True) = models.py:757-761, 803-829), feed them to
// LSTMCell(C -> 16), read a signed distance off Linear(16 -> 1), advance along the ray.  In torch
// that is ~25 launches per step with a (rays, C) feature matrix written and re-read in between;
// at train.py's 2048 rays it is pure launch latency.  Here a warp owns RB = 4 rays for the whole
// march and nothing leaves the SM between steps (training additionally saves what backward needs):
//
//   * feature fetch: the front end's own per-lane code (field_inputs_core.h): projection, bilinear
//     blend of four channels-last rows of the L2-resident feature map; a lane owns channels
//     {4*lane + 128*i}; the blended vector goes to a per-warp shared-memory row (and, when saving,
//     to HBM, where the weight-gradient GEMM reads it later);
//   * gates = W_ih v + b_ih + W_hh h + b_hh (torch.nn.LSTMCell, gate order i, f, g, o): W_ih
//     (64 x C, 128 KB at C = 512) sits in shared memory as [C/4][64][4], lanes own GATES (lane l:
//     gates l and l + 32), so one conflict-free LDS.128 of weights meets RB broadcast LDS.128 of
//     features and no cross-lane reduction is needed; fp32 FMAs (the recurrence amplifies TF32's
//     1e-3, so no tensor cores: the whole march is 10 GFLOP for a 128 x 128 frame);
//   * pointwise cell + Linear(16 -> 1) + advance in registers (lanes 0..15 own the hidden units).
//
// Backward (one warp, RB rays, steps in reverse): recompute nothing but the feature taps; per step
//   gh = w_out * (g_world . d) + W_hh^T da_{t+1},  clamped to [-10, 10]   (the hook of :427-428)
//   da = LSTM cell backward;  dv = W_ih^T da  (lanes own CHANNELS here, W_ih as [64][C])
//   d_latent += dv (x) bilinear weights (vector atomics);  g_world += d features / d point
// and writes da / (g_world . d) per (step, ray); the parameter gradients are plain GEMMs over those
// rows (dW_ih = DA^T V, dW_hh = DA^T H_prev, ...), done by the caller with cuBLAS.
#include "avr_common.cuh"
#include "field_inputs_core.h"
#include "kernels.h"

namespace avr {

constexpr int kMarchRB = 4;        // rays per warp
constexpr int kMarchHidden = 16;   // renderers.py:370: hidden_size = 16
constexpr int kMarchGates = 64;

using MarchArgs = ::avr_lstm_march;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// The point of ray `r` in its (single) source view; features_only descriptor.
__device__ __forceinline__ FieldPoint march_point(const FieldInputsArgs& f, const FieldView& w, float x, float y, float z) {
  return field_point_xyz(f, w, x, y, z, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void march_view(const FieldInputsArgs& f, int64_t obj, FieldView* w) {
  FieldCursor cur;
  cur.row = 0; cur.v = obj; cur.b = 0; cur.obj = obj;   // NS == 1: view index == object index
  field_view_fill(f, cur, w);
}

template <int CPL>
__device__ __forceinline__ void march_taps(const FieldInputsArgs& f, const FieldPoint& p, int64_t v, int lane, float (&t)[4][CPL][4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float* row = field_tap(f, v, p.x0 + (k & 1), p.y0 + (k >> 1));
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      if (row) {
        field_load4(row + 4 * lane + 128 * i, t[k][i]);
      } else {
        t[k][i][0] = t[k][i][1] = t[k][i][2] = t[k][i][3] = 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ forward
template <int CPL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
lstm_march_fwd_kernel(const FieldInputsArgs f, const MarchArgs m) {
  constexpr int C = 128 * CPL, RB = kMarchRB;
  extern __shared__ __align__(16) float smem[];
  float4* w4 = reinterpret_cast<float4*>(smem);                 // [C/4][64] float4: W_ih[gate][4*c4 .. 4*c4+3]
  float* whh = smem + C * kMarchGates;                          // [64][17]
  float* vs_all = whh + kMarchGates * 17;                       // [WARPS][RB][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* vs = vs_all + (size_t)warp * RB * C;

  for (int e = threadIdx.x; e < kMarchGates * (C / 4); e += WARPS * 32) {
    const int gate = e / (C / 4), c4 = e - gate * (C / 4);      // coalesced read of row `gate`
    w4[c4 * kMarchGates + gate] = __ldg(reinterpret_cast<const float4*>(m.w_ih + (size_t)gate * C) + c4);
  }
  for (int e = threadIdx.x; e < kMarchGates * kMarchHidden; e += WARPS * 32)
    whh[(e / kMarchHidden) * 17 + (e % kMarchHidden)] = m.w_hh[e];
  __syncthreads();

  const float bias0 = m.b_ih[lane] + m.b_hh[lane], bias1 = m.b_ih[lane + 32] + m.b_hh[lane + 32];
  const float wout = lane < kMarchHidden ? m.w_out[lane] : 0.f;
  const float bout = m.b_out[0];
  const int64_t n_groups = (m.R + RB - 1) / RB;
  FieldView view;
  field_view_reset(&view);

  // groups are dealt CTA-major (consecutive groups to different CTAs): a training batch has fewer groups than
  // the grid has warps, and every SM should get its share
  for (int64_t grp = (int64_t)warp * gridDim.x + blockIdx.x; grp < n_groups; grp += (int64_t)gridDim.x * WARPS) {
    float wx[RB], wy[RB], wz[RB], dx[RB], dy[RB], dz[RB], h[RB], c[RB];
    int64_t ray[RB];
    bool live[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int64_t r = grp * RB + b;
      live[b] = r < m.R;
      ray[b] = live[b] ? r : m.R - 1;
      const float* o = m.ros + ray[b] * 3;
      const float* d = m.rds + ray[b] * 3;
      const float t0 = m.init_dist[ray[b]];
      dx[b] = d[0]; dy[b] = d[1]; dz[b] = d[2];
      wx[b] = __fadd_rn(o[0], __fmul_rn(dx[b], t0));           // renderers.py:415: ros + rds * initial_distance
      wy[b] = __fadd_rn(o[1], __fmul_rn(dy[b], t0));
      wz[b] = __fadd_rn(o[2], __fmul_rn(dz[b], t0));
      h[b] = 0.f;
      c[b] = 0.f;
      if (live[b] && lane < 3) m.world[ray[b] * 3 + lane] = lane == 0 ? wx[b] : (lane == 1 ? wy[b] : wz[b]);
    }
    for (int t = 0; t < m.steps; ++t) {
      // ---- features of the RB current points -> vs (and HBM when saving) --------------------------
      __syncwarp();
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        const int64_t obj = ray[b] / m.rays_per_obj;
        march_view(f, obj, &view);
        const FieldPoint p = march_point(f, view, wx[b], wy[b], wz[b]);
        float tp[4][CPL][4];
        march_taps<CPL>(f, p, obj, lane, tp);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          float o4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) o4[q] = field_blend(p, tp[0][i][q], tp[1][i][q], tp[2][i][q], tp[3][i][q]);
          const float4 v4 = make_float4(o4[0], o4[1], o4[2], o4[3]);
          *reinterpret_cast<float4*>(vs + b * C + 4 * lane + 128 * i) = v4;
          if (m.feats && live[b])
            __stcs(reinterpret_cast<float4*>(m.feats + ((size_t)t * m.R + ray[b]) * C + 4 * lane + 128 * i), v4);
        }
      }
      __syncwarp();
      // ---- gates: lane owns gates `lane` and `lane + 32` -------------------------------------------
      float a0[RB], a1[RB];
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        a0[b] = bias0;
        a1[b] = bias1;
      }
#pragma unroll 4
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 wa = w4[c4 * kMarchGates + lane], wb = w4[c4 * kMarchGates + lane + 32];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          const float4 v4 = *reinterpret_cast<const float4*>(vs + b * C + 4 * c4);
          a0[b] = fmaf(wa.x, v4.x, a0[b]); a0[b] = fmaf(wa.y, v4.y, a0[b]);
          a0[b] = fmaf(wa.z, v4.z, a0[b]); a0[b] = fmaf(wa.w, v4.w, a0[b]);
          a1[b] = fmaf(wb.x, v4.x, a1[b]); a1[b] = fmaf(wb.y, v4.y, a1[b]);
          a1[b] = fmaf(wb.z, v4.z, a1[b]); a1[b] = fmaf(wb.w, v4.w, a1[b]);
        }
      }
#pragma unroll
      for (int k = 0; k < kMarchHidden; ++k) {
        const float u0 = whh[lane * 17 + k], u1 = whh[(lane + 32) * 17 + k];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          const float hk = __shfl_sync(0xffffffffu, h[b], k);
          a0[b] = fmaf(u0, hk, a0[b]);
          a1[b] = fmaf(u1, hk, a1[b]);
        }
      }
      // ---- cell, output layer, advance ------------------------------------------------------------
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        // lanes 0..15: i (gate j), g (gate 32 + j);  lanes 16..31: f (gate 16 + j), o (gate 48 + j)
        const float s0 = sigmoidf_acc(a0[b]);
        const float s1 = lane < kMarchHidden ? tanhf(a1[b]) : sigmoidf_acc(a1[b]);
        const float fj = __shfl_down_sync(0xffffffffu, s0, 16), oj = __shfl_down_sync(0xffffffffu, s1, 16);
        const float cn = fmaf(fj, c[b], s0 * s1);              // c' = f c + i g
        const float hn = oj * tanhf(cn);                       // h' = o tanh(c')
        if (m.gates && live[b]) {
          float* grow = m.gates + ((size_t)t * m.R + ray[b]) * kMarchGates;
          grow[lane] = s0;          // i | f
          grow[lane + 32] = s1;     // g | o
        }
        if (lane < kMarchHidden) {
          c[b] = cn;
          h[b] = hn;
          if (m.cells && live[b]) {
            m.cells[((size_t)t * m.R + ray[b]) * kMarchHidden + lane] = cn;
            m.hidden[((size_t)t * m.R + ray[b]) * kMarchHidden + lane] = hn;
          }
        }
        float part = lane < kMarchHidden ? wout * hn : 0.f;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        const float dist = __shfl_sync(0xffffffffu, part, 0) + bout;   // out_layer(h')  (renderers.py:430)
        wx[b] = __fadd_rn(wx[b], __fmul_rn(dx[b], dist));                // :432
        wy[b] = __fadd_rn(wy[b], __fmul_rn(dy[b], dist));
        wz[b] = __fadd_rn(wz[b], __fmul_rn(dz[b], dist));
        if (live[b] && lane < 3)
          m.world[((size_t)(t + 1) * m.R + ray[b]) * 3 + lane] = lane == 0 ? wx[b] : (lane == 1 ? wy[b] : wz[b]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ backward
template <int CPL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
lstm_march_bwd_kernel(const FieldInputsArgs f, const MarchArgs m) {
  constexpr int C = 128 * CPL, RB = kMarchRB;
  extern __shared__ __align__(16) float smem[];
  float* wih = smem;                                            // [64][C]
  float* whh = smem + C * kMarchGates;                          // [64][16]
  float* das_all = whh + kMarchGates * kMarchHidden;            // [WARPS][RB][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* das = das_all + (size_t)warp * RB * kMarchGates;
  for (int e = threadIdx.x; e < kMarchGates * C / 4; e += WARPS * 32)
    reinterpret_cast<float4*>(wih)[e] = __ldg(reinterpret_cast<const float4*>(m.w_ih) + e);
  for (int e = threadIdx.x; e < kMarchGates * kMarchHidden; e += WARPS * 32) whh[e] = m.w_hh[e];
  __syncthreads();

  const float wout = lane < kMarchHidden ? m.w_out[lane] : 0.f;
  const int64_t n_groups = (m.R + RB - 1) / RB;
  FieldView view;
  field_view_reset(&view);
  const int j = lane & (kMarchHidden - 1);

  // groups are dealt CTA-major (consecutive groups to different CTAs): a training batch has fewer groups than
  // the grid has warps, and every SM should get its share
  for (int64_t grp = (int64_t)warp * gridDim.x + blockIdx.x; grp < n_groups; grp += (int64_t)gridDim.x * WARPS) {
    float gx[RB], gy[RB], gz[RB], dx[RB], dy[RB], dz[RB], gh[RB], gc[RB];
    int64_t ray[RB];
    bool live[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int64_t r = grp * RB + b;
      live[b] = r < m.R;
      ray[b] = live[b] ? r : m.R - 1;
      const float* d = m.rds + ray[b] * 3;
      dx[b] = d[0]; dy[b] = d[1]; dz[b] = d[2];
      const float* g = m.g_world + ray[b] * 3;                  // dL / d world[steps]
      gx[b] = g[0]; gy[b] = g[1]; gz[b] = g[2];
      gh[b] = 0.f;   // gradient reaching h_t from step t+1's gates (lanes 0..15)
      gc[b] = 0.f;   // gradient reaching c_t from step t+1
    }
    for (int t = m.steps - 1; t >= 0; --t) {
      __syncwarp();
      // ---- cell backward: lanes 0..15 own the hidden units ----------------------------------------
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        const size_t row = (size_t)t * m.R + ray[b];
        const float* grow = m.gates + row * kMarchGates;
        const float gi = grow[j], gf = grow[16 + j], gg = grow[32 + j], go = grow[48 + j];
        const float ct = m.cells[row * kMarchHidden + j];
        const float cp = t > 0 ? m.cells[((size_t)(t - 1) * m.R + ray[b]) * kMarchHidden + j] : 0.f;
        const float dd = gx[b] * dx[b] + gy[b] * dy[b] + gz[b] * dz[b];   // d L / d signed_distance_t
        float ght = fmaf(wout, dd, gh[b]);
        ght = fminf(fmaxf(ght, -10.f), 10.f);                               // the hook on state[0] (:427-428)
        const float th = tanhf(ct);
        const float dcell = fmaf(ght * go, 1.0f - th * th, gc[b]);
        const float da_o = ght * th * go * (1.0f - go);
        const float da_i = dcell * gg * gi * (1.0f - gi);
        const float da_g = dcell * gi * (1.0f - gg * gg);
        const float da_f = dcell * cp * gf * (1.0f - gf);
        gc[b] = dcell * gf;
        if (lane < kMarchHidden) {
          float* s = das + b * kMarchGates;
          s[j] = da_i; s[16 + j] = da_f; s[32 + j] = da_g; s[48 + j] = da_o;
          if (live[b]) {
            float* o = m.d_gates + row * kMarchGates;
            o[j] = da_i; o[16 + j] = da_f; o[32 + j] = da_g; o[48 + j] = da_o;
            if (lane == 0) m.d_dist[row] = dd;
          }
        }
      }
      __syncwarp();
      // ---- gh for step t-1: W_hh^T da (lanes 0..15) -------------------------------------------------
#pragma unroll
      for (int b = 0; b < RB; ++b) gh[b] = 0.f;
#pragma unroll 8
      for (int g = 0; g < kMarchGates; ++g) {
        const float u = whh[g * kMarchHidden + j];
#pragma unroll
        for (int b = 0; b < RB; ++b) gh[b] = fmaf(u, das[b * kMarchGates + g], gh[b]);
      }
      // ---- dv = W_ih^T da: lanes own channels {4*lane + 128*i} -----------------------------------------
      float dv[RB][CPL][4];
#pragma unroll
      for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int i = 0; i < CPL; ++i) dv[b][i][0] = dv[b][i][1] = dv[b][i][2] = dv[b][i][3] = 0.f;
#pragma unroll 2
      for (int g = 0; g < kMarchGates; ++g) {
        float4 wv[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) wv[i] = *reinterpret_cast<const float4*>(wih + (size_t)g * C + 4 * lane + 128 * i);
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          const float a = das[b * kMarchGates + g];
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            dv[b][i][0] = fmaf(wv[i].x, a, dv[b][i][0]); dv[b][i][1] = fmaf(wv[i].y, a, dv[b][i][1]);
            dv[b][i][2] = fmaf(wv[i].z, a, dv[b][i][2]); dv[b][i][3] = fmaf(wv[i].w, a, dv[b][i][3]);
          }
        }
      }
      // ---- through the feature fetch: d_latent, and the point gradient into g_world ---------------------
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        const int64_t obj = ray[b] / m.rays_per_obj;
        march_view(f, obj, &view);
        const float* wp = m.world + ((size_t)t * m.R + ray[b]) * 3;
        const FieldPoint p = march_point(f, view, wp[0], wp[1], wp[2]);
        float tp[4][CPL][4];
        march_taps<CPL>(f, p, obj, lane, tp);
        const float wxf = fi_sub(p.ix, (float)p.x0), ex = fi_sub(1.0f, wxf);
        const float wyf = fi_sub(p.iy, (float)p.y0), sy = fi_sub(1.0f, wyf);
        float gix = 0.f, giy = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float g = dv[b][i][q];
            gix += g * ((tp[1][i][q] - tp[0][i][q]) * sy + (tp[3][i][q] - tp[2][i][q]) * wyf);
            giy += g * ((tp[2][i][q] - tp[0][i][q]) * ex + (tp[3][i][q] - tp[1][i][q]) * wxf);
          }
        if (f.d_latent && live[b]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float* row = field_grad_tap(f, obj, p.x0 + (k & 1), p.y0 + (k >> 1));
            const float wk = k == 0 ? p.nw : (k == 1 ? p.ne : (k == 2 ? p.sw : p.se));
            if (row) {
#pragma unroll
              for (int i = 0; i < CPL; ++i) {
                const float cc[4] = {dv[b][i][0] * wk, dv[b][i][1] * wk, dv[b][i][2] * wk, dv[b][i][3] * wk};
                field_atomic_add4(row + 4 * lane + 128 * i, cc);
              }
            }
          }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          gix += __shfl_xor_sync(0xffffffffu, gix, d);
          giy += __shfl_xor_sync(0xffffffffu, giy, d);
        }
        // chain rule of field_bwd_row_finish (features only: no code, no view direction)
        const float gu = p.clip_x ? 0.f : gix * ((float)(f.W - 1) / 2) * f.scale_x * view.fx;
        const float gt = p.clip_y ? 0.f : giy * ((float)(f.H - 1) / 2) * f.scale_y * view.fy;
        const float iz = 1.0f / p.cam2;
        const float d0 = -gu * iz, d1 = -gt * iz, d2 = (gu * p.cam0 + gt * p.cam1) * iz * iz;
        gx[b] += view.R[0] * d0 + view.R[4] * d1 + view.R[8] * d2;
        gy[b] += view.R[1] * d0 + view.R[5] * d1 + view.R[9] * d2;
        gz[b] += view.R[2] * d0 + view.R[6] * d1 + view.R[10] * d2;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ launchers
template <int CPL>
static int march_launch(const FieldInputsArgs& f, const MarchArgs& m, bool backward, cudaStream_t stream) {
  constexpr int C = 128 * CPL;
  constexpr int WARPS = 8;
  const size_t smem = backward
      ? sizeof(float) * ((size_t)C * kMarchGates + kMarchGates * kMarchHidden + (size_t)WARPS * kMarchRB * kMarchGates)
      : sizeof(float) * ((size_t)C * kMarchGates + kMarchGates * 17 + (size_t)WARPS * kMarchRB * C);
  auto kernel = backward ? lstm_march_bwd_kernel<CPL, WARPS> : lstm_march_fwd_kernel<CPL, WARPS>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    (void)cudaGetLastError();
    return AVR_ERR_LAUNCH;
  }
  const int64_t groups = (m.R + kMarchRB - 1) / kMarchRB;
  int64_t blocks = (groups + WARPS - 1) / WARPS;
  // few rays (a training batch): spread the groups over the SMs, one or two warps each, instead of
  // filling a handful of CTAs — every CTA pays the weight fill, but the march itself is latency-bound
  const int64_t sms = num_sms();
  if (blocks < sms) blocks = groups < sms ? groups : sms;
  if (blocks > sms) blocks = sms;
  kernel<<<(unsigned)blocks, WARPS * 32, smem, stream>>>(f, m);
  return check_launch();
}

int launch_lstm_march(const FieldInputsArgs& f, const MarchArgs& m, bool backward, cudaStream_t stream) {
  switch (f.C) {
    case 128: return march_launch<1>(f, m, backward, stream);
    case 256: return march_launch<2>(f, m, backward, stream);
    case 512: return march_launch<4>(f, m, backward, stream);
    default: return AVR_ERR_UNSUPPORTED;
  }
}

}  // namespace avr
