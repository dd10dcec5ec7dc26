// Shared helpers for libocrpp.so (sm_100a only). No CPU fallbacks live here or anywhere else.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ocrpp.h"

namespace ocrpp {

constexpr int kNumSMs = 148;  // B200

// thread-local last error message (ocrpp_last_error)
char* last_error_buf();
int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launch_count;
int tuning(int key);   // ocrpp_set_tuning value of `key` (0 = default)

#define OCRPP_CHECK_ARG(cond, ...)                                           \
  do {                                                                       \
    if (!(cond)) return ::ocrpp::set_error(OCRPP_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define OCRPP_CUDA(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::ocrpp::set_error(OCRPP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

// counts the launch and checks the launch status; with OCRPP_DEBUG_SYNC=1 in the environment every
// launch is followed by a device synchronisation so that a faulting kernel is reported by name/line
bool debug_sync();
#define OCRPP_LAUNCHED()                                         \
  do {                                                           \
    ::ocrpp::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    OCRPP_CUDA(cudaGetLastError());                              \
    if (::ocrpp::debug_sync()) OCRPP_CUDA(cudaDeviceSynchronize()); \
  } while (0)

// 128-bit streaming load that does not allocate in L1 (data is touched once)
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// the same, pinned in program order (the compiler may otherwise sink a batch of prefetching loads to their uses)
__device__ __forceinline__ uint4 ldg_stream_u4_ordered(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float h2f_lo(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu))); }
__device__ __forceinline__ float h2f_hi(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u >> 16))); }

template <typename T>
__device__ __forceinline__ float load_scalar(const T* p);
template <>
__device__ __forceinline__ float load_scalar<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_scalar<__half>(const __half* p) { return __half2float(__ldg(p)); }

// ---- per-kernel event timing (ocrpp_profile_*) ---------------------------------------------------
constexpr int kProfMaxPhases = 16;
constexpr int kProfMaxCalls = 64;
bool profile_on();
// starts a profiled call; returns a slot (>= 0) or -1 when profiling is off / the ring is full
int profile_begin(cudaStream_t s);
// records the end of phase `phase` (named `name`) of call `slot`
void profile_mark(int slot, int phase, const char* name, cudaStream_t s);

struct ProfileScope {
  int slot;
  int phase = 0;
  cudaStream_t s;
  explicit ProfileScope(cudaStream_t st) : slot(profile_begin(st)), s(st) {}
  void mark(const char* name) {
    if (slot >= 0) profile_mark(slot, phase, name, s);
    ++phase;
  }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ocrpp
