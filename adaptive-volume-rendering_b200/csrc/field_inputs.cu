// Radiance-field front end kernels (models.py:754-826): the per-lane
// code lives in field_inputs_core.h, shared with the host walk of the CPU test suite.
//
// Bound by the HBM write of the MLP input: (C + code) * 4 bytes per (view, point), 2216 B for
// conf/default.conf's 512 + 42; the feature rows it blends come from the L2-resident map (8 MB at
// 64 x 64 x 512) and stay in registers while consecutive samples of a ray sit in one texel cell.
// Measured (profiles/r01_field_inputs.md): forward 58 %, feature-map backward 46 %, point backward 20 % of
// the HBM roofline; all three latency-bound at 10-14 resident warps/SM, DRAM traffic = algorithmic bytes.
#include <cstdlib>

#include "avr_common.cuh"
#include "field_inputs_core.h"
#include "kernels.h"

namespace avr {

constexpr bool kFieldBwdPrefetchDefault = false;  // measured: 0.200 vs 0.147 ms (spills under the 168-register cap)
constexpr bool kFieldStageDefault = false;  // measured: 0.127 ms staged vs 0.118 ms direct (L2 merges the half sectors)
constexpr int kFieldWarps = 4;
constexpr int kFieldBwdAsyncDepth = 4;  // rows of g_out a warp keeps in flight in the feature-map gradient kernel
// consecutive rows per warp visit (samples of one ray, same view): 32 when the lanes share out the
// per-row coordinate work, 16 otherwise
template <bool kShare>
struct FieldChunk {
  static constexpr int value = kShare ? 32 : 16;
};

// kShare: the coordinate work of a row (field_point: two 3x3 products, the projection, two
// divisions, the bilinear weights — ~100 instructions that do not depend on the lane) is done ONCE
// per row instead of once per lane: lane l computes the point of row first + l of a 32-row chunk,
// and the row loop fetches row r's values from lane r by shuffles.
template <bool kFull>
__device__ __forceinline__ FieldPoint point_from_lane(const FieldPoint& m, int src) {
  FieldPoint p;
#define AVR_BCAST(f) p.f = __shfl_sync(0xffffffffu, m.f, src)
  AVR_BCAST(enc0); AVR_BCAST(enc1); AVR_BCAST(enc2);
  AVR_BCAST(vrot0); AVR_BCAST(vrot1); AVR_BCAST(vrot2);
  AVR_BCAST(x0); AVR_BCAST(y0);
  AVR_BCAST(nw); AVR_BCAST(ne); AVR_BCAST(sw); AVR_BCAST(se);
  if (kFull) {  // the backward pass also needs the camera-space point, the pixel coordinates and the clip flags
    AVR_BCAST(cam0); AVR_BCAST(cam1); AVR_BCAST(cam2);
    AVR_BCAST(ix); AVR_BCAST(iy);
    const int clip = __shfl_sync(0xffffffffu, (m.clip_x ? 1 : 0) | (m.clip_y ? 2 : 0), src);
    p.clip_x = (clip & 1) != 0;
    p.clip_y = (clip & 2) != 0;
  } else {
    p.cam0 = p.cam1 = p.cam2 = p.ix = p.iy = 0.f;
    p.clip_x = p.clip_y = false;
  }
#undef AVR_BCAST
  return p;
}

// Sum the eight partials of a row over the warp by a transposing butterfly: in each of the first three steps a
// lane hands half of its values to its partner and keeps the other half (4 + 2 + 1 shuffles), two more steps add
// up the four lanes that then hold the same value: 9 shuffles and
// 9 additions where one butterfly per value takes 40 and 40.  The totals of the row go to `row_totals[0..8)`
// (shared memory; order gix, giy, enc0..2, vr0..2).
__device__ __forceinline__ void field_partial_reduce(const FieldRowPartial* s, int lane, float* row_totals) {
  const unsigned full = 0xffffffffu;
  const float v[8] = {s->gix, s->giy, s->enc0, s->enc1, s->enc2, s->vr0, s->vr1, s->vr2};
  const bool u4 = (lane & 16) != 0, u3 = (lane & 8) != 0, u2 = (lane & 4) != 0;
  float k4[4], k2[2];
#pragma unroll
  for (int j = 0; j < 4; ++j) k4[j] = (u4 ? v[j + 4] : v[j]) + __shfl_xor_sync(full, u4 ? v[j] : v[j + 4], 16);
#pragma unroll
  for (int j = 0; j < 2; ++j) k2[j] = (u3 ? k4[j + 2] : k4[j]) + __shfl_xor_sync(full, u3 ? k4[j] : k4[j + 2], 8);
  float x = (u2 ? k2[1] : k2[0]) + __shfl_xor_sync(full, u2 ? k2[0] : k2[1], 4);
  x += __shfl_xor_sync(full, x, 2);
  x += __shfl_xor_sync(full, x, 1);
  // value 4*b4 + 2*b3 + b2 ended up on the lanes with those bits: value j on lanes 4j .. 4j+3
  if ((lane & 3) == 0) row_totals[lane >> 2] = x;
}

// End of a chunk: lane l finishes row first + l (chain rule through the projection and the rigid transform,
// atomics over the object's views) from the totals its row left in the warp's table — every lane at once, instead
// of lane 0 alone after every row while 31 lanes wait.
__device__ __forceinline__ void field_chunk_finish(const FieldInputsArgs& a, int64_t first, int n, int lane,
                                                   float (*totals)[8], FieldView* view) {
  __syncwarp();
  if (lane < n) {
    const float4 t0 = *reinterpret_cast<const float4*>(&totals[lane][0]);
    const float4 t1 = *reinterpret_cast<const float4*>(&totals[lane][4]);
    FieldRowPartial s;
    s.gix = t0.x; s.giy = t0.y; s.enc0 = t0.z; s.enc1 = t0.w;
    s.enc2 = t1.x; s.vr0 = t1.y; s.vr1 = t1.z; s.vr2 = t1.w;
    const FieldCursor mc = field_cursor_at(a, first + lane);
    field_view_fill(a, mc, view);
    const FieldPoint mp = field_point(a, mc, *view);
    field_bwd_row_finish(a, mc, *view, mp, s);
  }
  __syncwarp();  // the next chunk overwrites the table
}

// the point of row min(first + lane, rows - 1), with this lane's own view constants
__device__ __forceinline__ FieldPoint point_of_my_row(const FieldInputsArgs& a, int64_t first, int lane, int64_t rows,
                                                      FieldView* view) {
  const int64_t mine = first + lane < rows ? first + lane : rows - 1;
  const FieldCursor mc = field_cursor_at(a, mine);
  field_view_fill(a, mc, view);
  return field_point(a, mc, *view);
}

// kStage: rows leave through shared memory.  A lane's 16 channels of a row are 4 separate 16-byte
// pieces, and rows are only 8-byte aligned, so direct stores are 8 bytes per lane at a 16-byte
// stride: every 32-byte sector is written in two halves by two instructions (57 % of the roofline).
// Staged, two consecutive rows (2 * row bytes: a multiple of 16, starting 16-byte aligned at even
// rows) are assembled in a per-warp double buffer and leave as ONE bulk copy
// (cp.async.bulk.global.shared::cta), full lines at a time.
template <int CPL, bool kShare, bool kStage>
__global__ void __launch_bounds__(kFieldWarps * 32, 4)
field_inputs_fwd_kernel(const FieldInputsArgs a, int row_stride) {
  extern __shared__ __align__(16) float s_stage[];  // [warp][2 slots][2 rows][row_stride]
  constexpr int N = CPL > 0 ? CPL : 1;
  constexpr int kChunk = FieldChunk<kShare>::value;
  const int lane = threadIdx.x & 31;
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + kChunk - 1) / kChunk;
  const int64_t warps = (int64_t)gridDim.x * kFieldWarps;
  const FieldLaneCode lc = field_lane_code(a, lane);
  float* stage = s_stage + (size_t)(threadIdx.x >> 5) * 4 * row_stride;
  int slot = 0;
  FieldTapCache<N> cache;
  FieldView view;
  field_cache_reset(&cache);
  field_view_reset(&view);
  for (int64_t ch = blockIdx.x * (int64_t)kFieldWarps + (threadIdx.x >> 5); ch < n_chunks; ch += warps) {
    const int64_t first = ch * kChunk;  // even: pairs (r, r+1) of a chunk start at even rows
    const int n = (int)(first + kChunk < rows ? kChunk : rows - first);
    FieldPoint mine;
    if (kShare) mine = point_of_my_row(a, first, lane, rows, &view);
    FieldCursor cur = field_cursor_at(a, first);
    for (int r = 0; r < n; ++r, field_cursor_next(a, &cur)) {
      FieldPoint p;
      if (kShare) {
        p = point_from_lane<false>(mine, r);
      } else {
        field_view_fill(a, cur, &view);
        p = field_point(a, cur, view);
      }
      const bool paired = kStage && (r | 1) < n;  // both rows of the pair exist (only the very last row can be single)
      float* out = a.out + cur.row * row_stride;
      if (paired) {
        if ((r & 1) == 0) {
          if (lane == 0) bulk_wait_read<1>();  // the copy that last read this slot (two pairs ago) is done
          __syncwarp();
        }
        out = stage + (size_t)(2 * slot + (r & 1)) * row_stride;
      }
      if (paired) {
        if (CPL > 0) {
          field_row_lane<N, false>(a, cur, p, lane, out, lc, &cache);
        } else {
          field_row_lane_generic<false>(a, cur, p, lane, out, lc);
        }
        if (r & 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            bulk_s2g(a.out + (cur.row - 1) * row_stride, stage + (size_t)2 * slot * row_stride, (uint32_t)row_stride * 8u);
            bulk_commit();
          }
          slot ^= 1;
        }
      } else {
        if (CPL > 0) {
          field_row_lane<N, true>(a, cur, p, lane, out, lc, &cache);
        } else {
          field_row_lane_generic<true>(a, cur, p, lane, out, lc);
        }
      }
    }
  }
  if (kStage && lane == 0) bulk_wait_all<0>();
}

// one gradient kind per launch: capped at 170 registers (3 CTAs = 12 warps per SM; the uncapped 180 of
// the feature-map variant left 8 and ran 0.219 instead of 0.194 ms)
// kPre: the next row's upstream gradient is requested before the current row is processed (the
// feature-map variant was waiting on those loads: 4.6 long-scoreboard stalls per issued instruction)
template <int CPL, bool kLatent, bool kPoint, bool kShare, bool kPre>
__global__ void __launch_bounds__(kFieldWarps * 32, (kLatent && kPoint) ? 1 : 3)
field_inputs_bwd_kernel(const FieldInputsArgs a, int row_stride) {
  __shared__ __align__(16) float s_tot[kPoint ? kFieldWarps : 1][32][8];  // per warp: the rows' reduced partial sums
  constexpr int N = CPL > 0 ? CPL : 1;
  constexpr int kChunk = FieldChunk<kShare>::value;
  const int lane = threadIdx.x & 31;
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + kChunk - 1) / kChunk;
  const int64_t warps = (int64_t)gridDim.x * kFieldWarps;
  const FieldLaneCode lc = field_lane_code(a, lane);
  FieldTapCache<N> taps;
  FieldGradCache<N> grads;
  FieldView view;
  field_cache_reset(&taps);
  field_grad_reset(&grads);
  field_view_reset(&view);
  for (int64_t ch = blockIdx.x * (int64_t)kFieldWarps + (threadIdx.x >> 5); ch < n_chunks; ch += warps) {
    const int64_t first = ch * kChunk;
    const int n = (int)(first + kChunk < rows ? kChunk : rows - first);
    FieldRowGrad<N> rg, rg_next;
    if (CPL > 0 && kPre) field_load_row_grad<N>(a, first, lane, row_stride, &rg);
    FieldPoint mine;
    if (kShare) mine = point_of_my_row(a, first, lane, rows, &view);
    FieldCursor cur = field_cursor_at(a, first);
    for (int r = 0; r < n; ++r, field_cursor_next(a, &cur)) {
      if (CPL > 0 && kPre && r + 1 < n) field_load_row_grad<N>(a, cur.row + 1, lane, row_stride, &rg_next);
      FieldPoint p;
      if (kShare) {
        p = point_from_lane<kPoint>(mine, r);  // the feature-map gradient only needs the cell and its weights
      } else {
        field_view_fill(a, cur, &view);
        p = field_point(a, cur, view);
      }
      FieldRowPartial s;
      if (CPL > 0) {
        s = field_bwd_row_lane<N, kLatent, kPoint, kPre>(a, cur, p, lane, row_stride, lc, rg, &taps, &grads);
        if (kPre) rg = rg_next;
      } else {
        s = field_bwd_row_lane_generic<kLatent, kPoint>(a, cur, p, lane, row_stride, lc);
      }
      if (kPoint) field_partial_reduce(&s, lane, s_tot[threadIdx.x >> 5][r]);
    }
    if (kPoint) field_chunk_finish(a, first, n, lane, s_tot[threadIdx.x >> 5], &view);
  }
  if (kLatent && CPL > 0) field_grad_flush<N>(a, lane, &grads);
}

// kRing variant of the above: the rows of g_out (2216 B each, read exactly once) reach the warp through a
// two-stage shared-memory ring filled by 1-D bulk async copies (TMA, cp.async.bulk + mbarrier) one row
// PAIR ahead — rows are 8-byte aligned, pairs starting at an even row are 16-byte aligned and a multiple
// of 16 bytes long.  ncu had these kernels waiting on exactly those loads (4.6 long-scoreboard stall
// cycles per issued instruction at 10-14 resident warps); a register prefetch (kPre) spilled.  The
// ring costs no registers and the copy engine runs ahead across chunk boundaries.
template <int CPL, bool kLatent, bool kPoint, bool kShare>
__global__ void __launch_bounds__(kFieldWarps * 32, (kLatent && kPoint) ? 1 : 3)
field_inputs_bwd_ring_kernel(const FieldInputsArgs a, int row_stride) {
  __shared__ __align__(16) float s_tot[kPoint ? kFieldWarps : 1][32][8];  // per warp: the rows' reduced partial sums
  extern __shared__ __align__(16) unsigned char s_ring_raw[];
  static_assert(CPL > 0, "ring kernels exist for the cached channel counts");
  constexpr int N = CPL;
  constexpr int kChunk = FieldChunk<kShare>::value;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair_floats = 2 * row_stride;              // row_stride is even: a multiple of 16 bytes
  const int stage_floats = (pair_floats + 3) & ~3;
  float* ring = reinterpret_cast<float*>(s_ring_raw) + (size_t)warp * 2 * stage_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring_raw + (size_t)kFieldWarps * 2 * stage_floats * sizeof(float)) + warp * 2;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + kChunk - 1) / kChunk;
  const int64_t warps = (int64_t)gridDim.x * kFieldWarps;
  const FieldLaneCode lc = field_lane_code(a, lane);
  FieldTapCache<N> taps;
  FieldGradCache<N> grads;
  FieldView view;
  field_cache_reset(&taps);
  field_grad_reset(&grads);
  field_view_reset(&view);
  // lane 0: request rows [ch*kChunk + 2*pp, +2) into `stage` (a single last row is read directly instead)
  auto issue = [&](int64_t ch, int pp, int stage) {
    const int64_t first_row = ch * kChunk + 2 * pp;
    if (first_row + 2 <= rows) {
      const uint32_t bytes = (uint32_t)pair_floats * 4u;
      mbar_expect_tx(&bars[stage], bytes);
      bulk_g2s(ring + (size_t)stage * stage_floats, a.g_out + first_row * row_stride, bytes, &bars[stage]);
    }
  };
  int64_t ch = blockIdx.x * (int64_t)kFieldWarps + warp;
  int64_t k = 0;  // row pairs this warp has consumed
  if (lane == 0 && ch < n_chunks) issue(ch, 0, 0);
  for (; ch < n_chunks; ch += warps) {
    const int64_t first = ch * kChunk;
    const int n = (int)(first + kChunk < rows ? kChunk : rows - first);
    const int np = (n + 1) >> 1;
    FieldPoint mine;
    if (kShare) mine = point_of_my_row(a, first, lane, rows, &view);
    FieldCursor cur = field_cursor_at(a, first);
    for (int pp = 0; pp < np; ++pp, ++k) {
      {  // one pair ahead, across chunk boundaries
        int64_t c2 = ch;
        int p2 = pp + 1;
        if (p2 >= np) {
          c2 = ch + warps;
          p2 = 0;
        }
        __syncwarp();  // every lane is done with the stage pair k-1 used
        if (lane == 0 && c2 < n_chunks) issue(c2, p2, (int)((k + 1) & 1));
      }
      const int stage = (int)(k & 1);
      const int nrows = n - 2 * pp < 2 ? n - 2 * pp : 2;
      if (nrows == 2) mbar_wait(&bars[stage], (uint32_t)((k >> 1) & 1));
      const float* srow = ring + (size_t)stage * stage_floats;
      for (int q = 0; q < nrows; ++q, field_cursor_next(a, &cur)) {
        const int r = 2 * pp + q;
        FieldPoint p;
        if (kShare) {
          p = point_from_lane<kPoint>(mine, r);
        } else {
          field_view_fill(a, cur, &view);
          p = field_point(a, cur, view);
        }
        FieldRowGrad<N> rg;
        FieldRowPartial s;
        if (nrows == 2) {
          s = field_bwd_row_lane<N, kLatent, kPoint, false>(a, cur, p, lane, row_stride, lc, rg, &taps, &grads,
                                                            srow + (size_t)q * row_stride);
        } else {  // the single last row of all: straight from global memory
          s = field_bwd_row_lane<N, kLatent, kPoint, false>(a, cur, p, lane, row_stride, lc, rg, &taps, &grads);
        }
        if (kPoint) field_partial_reduce(&s, lane, s_tot[warp][r]);
      }
    }
    if (kPoint) field_chunk_finish(a, first, n, lane, s_tot[warp], &view);
  }
  if (kLatent) field_grad_flush<N>(a, lane, &grads);
}

// Feature-map gradient with a DEEP private prefetch of g_out.  ncu's source page has the plain kernel waiting on
// exactly one instruction, the first use of a row's g_out (half of all stall samples): a warp has one row
// (2 KB) in flight, 10.5 resident warps per SM hold 23 KB against the ~30 KB that cover HBM's latency at full
// rate, and registers for a second row do not exist under the 168-register cap.  Here every lane copies ITS OWN
// pieces of the next kDepth - 1 rows into a per-warp shared-memory ring with cp.async (8-byte copies: rows of
// g_out are only 8-byte aligned; the staged groups are 16-byte aligned) and reads them back itself — no
// barrier, no mbarrier, no cross-lane traffic, no registers held by loads in flight.
template <int CPL, bool kShare, int kDepth, int kMinBlocks>
__global__ void __launch_bounds__(kFieldWarps * 32, kMinBlocks)
field_inputs_bwd_latent_async_kernel(const FieldInputsArgs a, int row_stride) {
  extern __shared__ __align__(16) unsigned char s_ring_raw[];
  static_assert(CPL > 0 && kDepth >= 2, "ring of the cached channel counts");
  constexpr int N = CPL;
  constexpr int kChunk = FieldChunk<kShare>::value;
  constexpr int kRow = 128 * CPL;  // floats of a staged row: the channels only
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(s_ring_raw) + (size_t)warp * kDepth * kRow;
  const int64_t rows = a.NV * a.B;
  const int64_t n_chunks = (rows + kChunk - 1) / kChunk;
  const int64_t warps = (int64_t)gridDim.x * kFieldWarps;
  const int64_t ch0 = blockIdx.x * (int64_t)kFieldWarps + warp;
  const FieldLaneCode lc = field_lane_code(a, lane);
  FieldTapCache<N> taps;
  FieldGradCache<N> grads;
  FieldView view;
  field_cache_reset(&taps);
  field_grad_reset(&grads);
  field_view_reset(&view);
  // prefetch cursor: the next row to request (rows of this warp's chunks, in the order it visits them), kept as a
  // source pointer, a count of rows left in the cursor's chunk and a ring slot — no per-row index arithmetic
  const float* pf_src = a.g_out + ch0 * kChunk * (int64_t)row_stride + 4 * lane;
  const int64_t pf_jump = ((warps - 1) * kChunk) * (int64_t)row_stride;  // from the end of a chunk to this warp's next one
  int64_t pf_chunk = ch0;
  auto rows_in = [&](const int64_t chunk) -> int {
    if (chunk >= n_chunks) return 0;
    const int64_t first = chunk * kChunk;
    return (int)(first + kChunk < rows ? kChunk : rows - first);
  };
  int pf_left = rows_in(pf_chunk);
  uint32_t pf_dst = smem_u32(ring + 4 * lane);
  const uint32_t ring_lo = pf_dst, ring_hi = pf_dst + (uint32_t)(kDepth * kRow * sizeof(float));
  auto issue = [&]() {
    if (pf_left > 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(pf_dst + 512u * i), "l"(pf_src + 128 * i) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(pf_dst + 512u * i + 8u), "l"(pf_src + 128 * i + 2) : "memory");
      }
      pf_src += row_stride;
      if (--pf_left == 0) {
        pf_chunk += warps;
        pf_src += pf_jump;
        pf_left = rows_in(pf_chunk);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // an empty group keeps the count uniform
    pf_dst += (uint32_t)(kRow * sizeof(float));             // the slot advances with every group, used or not
    if (pf_dst == ring_hi) pf_dst = ring_lo;
  };
#pragma unroll
  for (int j = 0; j < kDepth - 1; ++j) issue();
  const float* rd = ring;  // the slot of the row being consumed
  for (int64_t ch = ch0; ch < n_chunks; ch += warps) {
    const int64_t first = ch * kChunk;
    const int n = (int)(first + kChunk < rows ? kChunk : rows - first);
    FieldPoint mine;
    if (kShare) mine = point_of_my_row(a, first, lane, rows, &view);
    FieldCursor cur = field_cursor_at(a, first);
    for (int r = 0; r < n; ++r, field_cursor_next(a, &cur)) {
      issue();
      asm volatile("cp.async.wait_group %0;" ::"n"(kDepth - 1) : "memory");
      FieldPoint p;
      if (kShare) {
        p = point_from_lane<false>(mine, r);
      } else {
        field_view_fill(a, cur, &view);
        p = field_point(a, cur, view);
      }
      FieldRowGrad<N> rg;
      (void)field_bwd_row_lane<N, true, false, false, true>(a, cur, p, lane, row_stride, lc, rg, &taps, &grads, rd);
      rd += kRow;
      if (rd == ring + kDepth * kRow) rd = ring;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  field_grad_flush<N>(a, lane, &grads);
}

// Experiment knobs (A/B measurements; defaults are what the measurements picked):
//   AVR_FIELD_NOCACHE=1      every channel count takes the generic walk (no register caches)
//   AVR_FIELD_BWD_SPLIT=0    feature-map and point gradients in ONE launch instead of two (234 registers,
//                            8 resident warps/SM: 0.645 ms against 0.56 ms for the two launches)
//   AVR_FIELD_SHARE_POINT=x  per-row coordinate work shared out over the lanes (see point_from_lane)
//   AVR_FIELD_STAGE=x        forward rows leave through shared memory and bulk copies (see kStage)
//   AVR_FIELD_BWD_PREFETCH=x backward kernels that share the point work also prefetch the next row of g_out
//   AVR_FIELD_BWD_RING=0     backward without the bulk-copy ring for g_out (see field_inputs_bwd_ring_kernel)
static bool field_no_cache() { return option(OPT_FIELD_NOCACHE, 0) != 0; }
static bool field_bwd_split() { return option(OPT_FIELD_BWD_SPLIT, 1) != 0; }
// unset: the measured defaults (every kernel shares; the point backward did not while lane 0 finished each row by
// itself, 0.346 vs 0.362 ms — with the per-chunk finish sharing wins, 0.195 vs 0.218 ms); 0 / 1 force every kernel one way
static bool field_share_point(bool dflt) { return option(OPT_FIELD_SHARE_POINT, dflt ? 1 : 0) != 0; }
static bool field_bwd_prefetch() { return option(OPT_FIELD_BWD_PREFETCH, kFieldBwdPrefetchDefault ? 1 : 0) != 0; }
static bool field_stage_rows() { return option(OPT_FIELD_STAGE, kFieldStageDefault ? 1 : 0) != 0; }

static unsigned field_grid(int64_t rows, int chunk) {
  const int64_t n_chunks = (rows + chunk - 1) / chunk;
  int64_t blocks = (n_chunks + kFieldWarps - 1) / kFieldWarps;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (unsigned)(blocks > cap ? cap : blocks);
}

template <int CPL, bool kLatent, bool kPoint, bool kShare>
static bool launch_bwd_ring(const FieldInputsArgs& a, int row_stride, cudaStream_t stream) {
  if constexpr (CPL == 0) {
    return false;
  } else {
    // measured (profiles/r02_field_ring.md): the point gradient gains a third (0.343 -> 0.257 ms), the feature-map
    // gradient nothing (0.146 -> 0.149 ms: it waits on its vector atomics, not on g_out), so by default only kernels
    // that form the point gradient take the ring; AVR_FIELD_BWD_RING=2 forces it everywhere, 0 nowhere
    const int mode = option(OPT_FIELD_BWD_RING, 1);
    if (mode == 0 || (mode == 1 && !kPoint) || (row_stride & 1) || !aligned16(a.g_out)) return false;
    const int stage_floats = (2 * row_stride + 3) & ~3;
    const size_t smem = (size_t)kFieldWarps * 2 * stage_floats * sizeof(float) + kFieldWarps * 2 * sizeof(uint64_t);
    auto kern = field_inputs_bwd_ring_kernel<CPL, kLatent, kPoint, kShare>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    kern<<<field_grid(a.NV * a.B, FieldChunk<kShare>::value), kFieldWarps * 32, smem, stream>>>(a, row_stride);
    return true;
  }
}

// feature-map gradient through the cp.async ring (AVR_FIELD_BWD_ASYNC=0: the plain kernel)
template <int CPL, bool kShare>
static bool launch_bwd_latent_async(const FieldInputsArgs& a, int row_stride, cudaStream_t stream) {
  if constexpr (CPL == 0) {
    return false;
  } else {
    // measured on B200 (2048 rays x 96 samples, 512 channels; profiles/r02_field_async.md): plain kernel 0.1458 ms;
    // ring of 2 / 3 / 4 / 6 rows at three CTAs per SM (168 registers, ~100 B of spills) 0.135 / 0.140 / 0.137 /
    // 0.151 ms; at two CTAs per SM (no spills) 3 / 4 / 6 / 8 rows: 0.131 / 0.125 / 0.131 / 0.128 ms -> 4 rows, two CTAs
    if (!option(OPT_FIELD_BWD_ASYNC, 1) || (row_stride & 1) || (reinterpret_cast<uintptr_t>(a.g_out) & 7u)) return false;
    constexpr int kDepth = kFieldBwdAsyncDepth;
    auto kern = field_inputs_bwd_latent_async_kernel<CPL, kShare, kDepth, 2>;
    const size_t smem = (size_t)kFieldWarps * kDepth * 128 * CPL * sizeof(float);
    static_assert((size_t)kFieldWarps * kDepth * 128 * 4 * sizeof(float) <= 48 * 1024, "no opt-in shared memory needed");
    kern<<<field_grid(a.NV * a.B, FieldChunk<kShare>::value), kFieldWarps * 32, smem, stream>>>(a, row_stride);
    return true;
  }
}

template <int CPL, bool kLatent, bool kPoint>
static void launch_bwd_kernel(const FieldInputsArgs& a, int row_stride, bool share, cudaStream_t stream) {
  const unsigned t = kFieldWarps * 32;
  if (kLatent && !kPoint &&
      (share ? launch_bwd_latent_async<CPL, true>(a, row_stride, stream) : launch_bwd_latent_async<CPL, false>(a, row_stride, stream)))
    return;
  if (!(share && field_bwd_prefetch()) &&
      (share ? launch_bwd_ring<CPL, kLatent, kPoint, true>(a, row_stride, stream)
             : launch_bwd_ring<CPL, kLatent, kPoint, false>(a, row_stride, stream)))
    return;
  if (share && field_bwd_prefetch()) {
    field_inputs_bwd_kernel<CPL, kLatent, kPoint, true, true><<<field_grid(a.NV * a.B, FieldChunk<true>::value), t, 0, stream>>>(a, row_stride);
  } else if (share) {
    field_inputs_bwd_kernel<CPL, kLatent, kPoint, true, false><<<field_grid(a.NV * a.B, FieldChunk<true>::value), t, 0, stream>>>(a, row_stride);
  } else {
    field_inputs_bwd_kernel<CPL, kLatent, kPoint, false, false><<<field_grid(a.NV * a.B, FieldChunk<false>::value), t, 0, stream>>>(a, row_stride);
  }
}

template <int CPL>
static void launch_bwd_variant(const FieldInputsArgs& a, int row_stride, cudaStream_t stream) {
  const bool latent = a.d_latent != nullptr, point = a.d_xyz != nullptr || a.d_viewdirs != nullptr;
  if (latent && point && !field_bwd_split()) {
    launch_bwd_kernel<CPL, true, true>(a, row_stride, field_share_point(false), stream);
    return;
  }
  if (latent) launch_bwd_kernel<CPL, true, false>(a, row_stride, field_share_point(true), stream);
  if (point) launch_bwd_kernel<CPL, false, true>(a, row_stride, field_share_point(true), stream);
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
  const int row_stride = a.C + (a.features_only ? 0 : field_code_width(a));
  switch (field_no_cache() ? 0 : a.C) {
    case 512: launch_bwd_variant<4>(a, row_stride, stream); break;
    case 256: launch_bwd_variant<2>(a, row_stride, stream); break;
    case 128: launch_bwd_variant<1>(a, row_stride, stream); break;
    default: launch_bwd_variant<0>(a, row_stride, stream); break;
  }
  return check_launch();
}

template <int CPL, bool kShare>
static int launch_fwd_variant(const FieldInputsArgs& a, int row_stride, cudaStream_t stream) {
  const unsigned g = field_grid(a.NV * a.B, FieldChunk<kShare>::value), t = kFieldWarps * 32;
  // staging buffer: 2 slots x 2 rows per warp; needs 16-byte aligned output and room in shared memory
  const size_t smem = (size_t)kFieldWarps * 4 * row_stride * sizeof(float);
  if (field_stage_rows() && aligned16(a.out) && smem <= 56 * 1024) {
    auto kern = field_inputs_fwd_kernel<CPL, kShare, true>;
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_last_cuda_error(e);
        return AVR_ERR_LAUNCH;
      }
    }
    kern<<<g, t, smem, stream>>>(a, row_stride);
  } else {
    field_inputs_fwd_kernel<CPL, kShare, false><<<g, t, 0, stream>>>(a, row_stride);
  }
  return check_launch();
}

template <bool kShare>
static int launch_fwd_share(const FieldInputsArgs& a, int row_stride, cudaStream_t stream) {
  switch (field_no_cache() ? 0 : a.C) {
    case 512: return launch_fwd_variant<4, kShare>(a, row_stride, stream);
    case 256: return launch_fwd_variant<2, kShare>(a, row_stride, stream);
    case 128: return launch_fwd_variant<1, kShare>(a, row_stride, stream);
    default: return launch_fwd_variant<0, kShare>(a, row_stride, stream);
  }
}

int launch_field_inputs_fwd(const FieldInputsArgs& a, cudaStream_t stream) {
  const int64_t rows = a.NV * a.B;
  if (rows == 0) return AVR_OK;
  const int row_stride = a.C + (a.features_only ? 0 : field_code_width(a));
  return field_share_point(true) ? launch_fwd_share<true>(a, row_stride, stream) : launch_fwd_share<false>(a, row_stride, stream);
}

}  // namespace avr
