// Arguments shared by the register-resident importance samplers
// (importance_reg.cu, importance_grp.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace avr {

struct ImportanceRegArgs {
  const float* weights;
  const float* z_coarse;
  const float* u;
  const float* u2;
  const float* normals;
  const float* near;
  const float* far;
  int bound_stride;
  const int64_t* offsets;       // packed coarse layout or null
  const int64_t* fine_offsets;  // packed fine layout or null
  int64_t R;
  int Kc, n_imp, n_depth;       // dense: exact; packed: Kc/n_imp unused (per-ray counts come from the offsets)
  float depth_std;
  int vec4;  // u/u2 rows are 16-byte aligned (n % 4 == 0 and aligned bases)
  int vecw;  // weight rows are 16-byte aligned
  int vecz;  // z_coarse rows, normals rows and the z_sorted rows are 16-byte aligned
  int skip_grp_classes;  // packed, importance_reg.cu's class kernels: leave rays that importance_grp.cu's ragged classes cover
  float* z_fine;
  float* z_sorted;
  float* cdf;
  int32_t* idx;
};

// importance_grp.cu: static dense shapes, G lanes per ray; AVR_ERR_UNSUPPORTED for other shapes
int launch_importance_grp(const ImportanceRegArgs& a, cudaStream_t stream);
// packed layout: per-class ragged kernels for rays of at most 256 coarse / 128 new samples
int launch_importance_grp_ragged(const ImportanceRegArgs& a, int max_coarse, int max_fine, bool* covers_all,
                                 cudaStream_t stream);

// importance_bins.cu: the same shapes by bucket ranking instead of sorting networks
int launch_importance_bins(const ImportanceRegArgs& a, cudaStream_t stream);
int launch_importance_bins_ragged(const ImportanceRegArgs& a, int max_coarse, int max_fine, bool* covers_all,
                                  cudaStream_t stream);

}  // namespace avr
