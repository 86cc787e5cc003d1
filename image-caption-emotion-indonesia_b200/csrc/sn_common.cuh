// Shared helpers for libsn100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/sn100.h"

namespace sn {

// per-thread last error text (re-entrant, no locks)
char* err_buf();
int32_t fail(int32_t code, const char* fmt, ...);
int32_t check_launch(const char* what);

struct DevInfo {
  int sm_count;
  int smem_optin;
  int cc_major, cc_minor;
};
// immutable per-device attribute cache (std::call_once per device)
const DevInfo& dev_info();
// Launch mode of the persistent recurrence kernels.  Their CTAs exchange data through flags in global memory, so every
// CTA of the grid must become resident.  The grid is sized to at most one CTA per SM and nothing that can occupy an SM
// in this library waits on these kernels, so a plain launch always makes progress; it avoids the ~8-10 us per launch
// that a cooperative launch costs inside a CUDA graph (timeline, profiles/r1_l).  SN_RECUR_COOP=1 forces cooperative
// launches (the driver then guarantees co-residency, e.g. when other processes share the GPU).
bool recur_cooperative();

#define SN_REQUIRE(cond, ...)                         \
  do {                                                \
    if (!(cond)) return sn::fail(-1, __VA_ARGS__);    \
  } while (0)

#define SN_CUDA(expr)                                                           \
  do {                                                                          \
    cudaError_t e__ = (expr);                                                   \
    if (e__ != cudaSuccess)                                                     \
      return sn::fail((int32_t)e__, "%s: %s", #expr, cudaGetErrorString(e__));  \
  } while (0)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// counter-based RNG for dropout: one 32-bit hash per (seed, row, col)
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint32_t row, uint32_t col) {
  uint64_t x = seed ^ ((uint64_t)row << 32 | col);
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t row, uint32_t col, float p,
                                               float inv_keep) {
  if (p <= 0.f) return 1.f;
  // keep with probability 1-p
  float u = (hash_u32(seed, row, col) >> 8) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace sn
