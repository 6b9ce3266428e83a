// Process-wide runtime state of the library that is NOT on any kernel's data path: tuning/test
// switches (read from the environment ONCE, overridable through avr_set_option), the cached SM
// count per device, dispatch counters (which kernel family served a call — tests and bench assert
// on them) and the wait side of the fused all-gather's completion signal.
#include <atomic>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "avr_common.cuh"
#include "kernels.h"

namespace avr {

// ---- options ---------------------------------------------------------------------------
static const char* const kOptNames[OPT_COUNT] = {
    "AVR_SPAN_L",          "AVR_SPAN_STAGES",      "AVR_SPAN_WARPS",         "AVR_COARSE_PACKED",
    "AVR_PACKED_SPAN",     "AVR_IMPORTANCE_GRP",   "AVR_PACKED_CLASSES",     "AVR_GRP_G",
    "AVR_FIELD_NOCACHE",   "AVR_FIELD_BWD_SPLIT",  "AVR_FIELD_SHARE_POINT",  "AVR_FIELD_BWD_PREFETCH",
    "AVR_FIELD_STAGE",     "AVR_IMPORTANCE_BINS",  "AVR_FIELD_BWD_RING",    "AVR_HOST_SLOTS",         "AVR_HOST_CHUNK_MIB",
    "AVR_FIELD_BWD_ASYNC",
};
constexpr int kUnset = INT_MIN;
static std::atomic<int> g_opt[OPT_COUNT];
static std::once_flag g_opt_once;

static int parse_option(const char* v) {
  if (!v || !*v) return kUnset;
  if ((*v >= '0' && *v <= '9') || *v == '-') return std::atoi(v);
  return (*v == 'r' || *v == 'y' || *v == 't') ? 1 : 0;  // "ray", "yes", "true"
}

static void init_options() {
  for (int i = 0; i < OPT_COUNT; ++i) g_opt[i].store(parse_option(std::getenv(kOptNames[i])));
}

int option(Opt o, int dflt) {
  std::call_once(g_opt_once, init_options);
  const int v = g_opt[o].load(std::memory_order_relaxed);
  return v == kUnset ? dflt : v;
}

// ---- device facts ------------------------------------------------------------------------
int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = kNumSMs;
  cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

// ---- dispatch counters -------------------------------------------------------------------
static std::atomic<int64_t> g_dispatch[AVR_DISPATCH_COUNT];
void count_dispatch(int which) {
  if (which >= 0 && which < AVR_DISPATCH_COUNT) g_dispatch[which].fetch_add(1, std::memory_order_relaxed);
}

// ---- fused all-gather: the waiting side ------------------------------------------------------
// One warp; lane p polls word p until it reaches `value` (serial-number arithmetic, so the 32-bit
// step counter may wrap).  The spin is bounded (~4 s of globaltimer): a peer that died must not
// hang this GPU — *status is set to 1 instead and the consumer finds out.
__global__ void gather_wait_kernel(const uint32_t* flags, int n, uint32_t value, uint32_t* status) {
  const int p = threadIdx.x;
  if (p >= n) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (unsigned it = 0;; ++it) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
    if ((int32_t)(v - value) >= 0) return;
    if ((it & 1023u) == 1023u) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) {
        if (status) atomicExch(status, 1u);
        return;
      }
    }
  }
}

int launch_gather_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* status, cudaStream_t stream) {
  gather_wait_kernel<<<1, 32, 0, stream>>>(flags, n, value, status);
  return check_launch();
}

}  // namespace avr

using namespace avr;

extern "C" {

int avr_set_option(const char* name, int value, int unset) {
  if (!name) return AVR_ERR_BAD_ARG;
  std::call_once(g_opt_once, init_options);
  for (int i = 0; i < OPT_COUNT; ++i) {
    if (std::strcmp(name, kOptNames[i]) == 0) {
      g_opt[i].store(unset ? kUnset : value);
      return AVR_OK;
    }
  }
  return AVR_ERR_BAD_ARG;
}

int avr_dispatch_counters(int64_t* out, int n) {
  if (!out || n < 0) return AVR_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) out[i] = i < AVR_DISPATCH_COUNT ? g_dispatch[i].load(std::memory_order_relaxed) : 0;
  return AVR_DISPATCH_COUNT;
}

void avr_dispatch_reset(void) {
  for (auto& c : g_dispatch) c.store(0);
}

int avr_gather_wait(const uint32_t* flags, int n_sources, uint32_t value, uint32_t* status, avr_stream_t stream) {
  if (!flags || n_sources < 1 || n_sources > 32) return AVR_ERR_BAD_ARG;
  return launch_gather_wait(flags, n_sources, value, status, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
