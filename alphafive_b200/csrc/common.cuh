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
// Programmatic dependent launch.  Every kernel of a search pass lets its successor be scheduled at once
// (griddepcontrol.launch_dependents at its top: the successor's CTAs take SMs as this grid's CTAs exit and run
// their prologue behind its tail) and waits for its predecessor (griddepcontrol.wait: completion + memory
// visibility) right before it first touches the predecessor's output.  Every kernel of the chain executes the
// wait, so "my predecessor is complete" is transitive along the stream.  Without the launch attribute (or after a
// memcpy) both instructions are no-ops.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();                             // A5_TC_PDL=0 turns every programmatic launch attribute off (engine.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_k(void (*kernel)(KArgs...), unsigned grid, unsigned threads, size_t smem, cudaStream_t st,
                                       bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------
// In-situ kernel timing (a5__debug_ktime_*): every kernel of a self-play pass can stamp %globaltimer when
// its first thread starts and when each CTA has finished, into its slot of a device buffer
// [slot][32 sub-slots][start_min, end_max]; a fold kernel at the end of the pass turns the stamps into
// per-slot sums "end of my predecessor -> my end" (these partition the pass exactly, whatever the overlap
// between programmatic-dependent launches) and "my first start -> my end".  Works inside CUDA-graph
// replays; costs one predicated branch per CTA when off (kt == nullptr).
// ---------------------------------------------------------------------------
constexpr int KT_SLOTS = 16;
constexpr int KT_SUB = 32;
enum { KT_C1BITS = 0, KT_CONV1 = 1, KT_CONV0 = 2 /* .. +7: the 8 block-conv launches */, KT_HEADS = 10, KT_STEP = 11, KT_FOLD = 12,
       KT_EC_LOOKUP = 13, KT_EC_COMMIT = 14 /* evaluation cache: before the bitboards / between heads and tree pass */ };
unsigned long long* kt_slot(int slot);          // device pointer of a slot, or nullptr while timing is off (net_tc.cu)

__device__ __forceinline__ unsigned long long kt_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void kt_begin(unsigned long long* kt) {
  if (kt && threadIdx.x == 0) atomicMin(kt + 2 * (blockIdx.x & (KT_SUB - 1)), kt_now());
}
// call after a CTA-wide barrier that follows the CTA's last work
__device__ __forceinline__ void kt_end(unsigned long long* kt) {
  if (kt && threadIdx.x == 0) atomicMax(kt + 2 * (blockIdx.x & (KT_SUB - 1)) + 1, kt_now());
}

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
