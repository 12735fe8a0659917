// Shared helpers for libalphafive (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/alphafive.h"

namespace a5 {

void set_error(const char* fmt, ...);

#define A5_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      a5::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return A5_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define A5_ARG(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      a5::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);  \
      return A5_ERR_ARG;                                                    \
    }                                                                       \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011).  Key = (seed, game),
// counter = (event counter lo/hi, lane-specific a, b): any draw is addressable,
// so results do not depend on scheduling or on the number of ranks.
// ---------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed, uint64_t stream) {
    uint64_t k = seed ^ (stream * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull);
    k0 = (uint32_t)k;
    k1 = (uint32_t)(k >> 32);
  }
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// uniform in (0, 1]
__device__ __forceinline__ float u01(uint32_t x) { return ((x >> 8) + 1u) * (1.0f / 16777216.0f); }
// uniform double in [0, 1)
__device__ __forceinline__ double u01d(uint32_t a, uint32_t b) {
  return (((uint64_t)a << 21) ^ (uint64_t)(b >> 11)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_xor(uint32_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v ^= __shfl_xor_sync(FULL, v, o);
  return v;
}

}  // namespace a5
