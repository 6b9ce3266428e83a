// Shared device/host helpers for the avr_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "avr_b200.h"

namespace avr {

// Constants of the reference's arithmetic (renderers.py:80, :92, :36).
constexpr float kLastDelta = 1e10f;   // delta of the last sample of a ray
constexpr float kTransEps = 1e-10f;   // added to (1 - alpha) before the cumprod
constexpr float kPdfEps = 1e-5f;      // added to the weights before normalising

constexpr int kNumSMs = 148;          // B200

// Per-sample opacity terms exactly as volume_integral forms them (renderers.py:86, :92):
//   e = exp(-(sigma*delta));  alpha = 1 - e;  t = (1 - alpha) + 1e-10
struct Opacity {
  float e, alpha, t;
};

__device__ __forceinline__ Opacity opacity(float sigma, float delta) {
  Opacity o;
  o.e = expf(-(sigma * delta));
  o.alpha = 1.0f - o.e;
  o.t = (1.0f - o.alpha) + kTransEps;
  return o;
}

// Camera depth folded into the compositing kernels (utils.depth_from_world at renderers.py:274-275,
// 508-509): the depth of ros + rds * dist seen from the ray's camera is AFFINE in the composited
// distance, depth = A * dist + B, with (A, B) per ray formed once by the ray-setup kernel
// (geometry.cu).  `aff` == nullptr: the kernels return / differentiate the distance itself.
__device__ __forceinline__ float cam_depth(const float* __restrict__ aff, int64_t ray, float dist) {
  return aff ? fmaf(aff[2 * ray], dist, aff[2 * ray + 1]) : dist;
}
__device__ __forceinline__ float cam_depth_grad(const float* __restrict__ aff, int64_t ray, float g) {
  return aff ? g * aff[2 * ray] : g;
}

// Streaming (read-once / write-once) global accesses: keep them out of L1.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// ---- mbarrier + 1-D bulk async copy (TMA) wrappers --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared, completion counted in bytes on `bar`.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (before bulk_s2g)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- host side -------------------------------------------------------------------
void set_last_cuda_error(cudaError_t e);
int check_launch();  // AVR_OK or AVR_ERR_LAUNCH (records the error string)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace avr
