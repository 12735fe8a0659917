// Issue-side behaviour: clock after each of 32 back-to-back tcgen05.mma issues (queue depth), commit, completion: M = 256 over a CTA pair, N in {32..256}, SS operands, leader issues.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_bench(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, iters); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  // NW issuing warps (lane 0 of warps 0..NW-1 of the leader), each with its own accumulator columns
  const int NW = iters;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < NW && rank == 0) {
    const uint32_t a0 = smem_u32(smem) + 1024, b0 = smem_u32(smem) + 160 * 1024;
    const uint32_t id = idesc_f16(256, N);
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ad[i] = desc(a0 + i * 2048 + w * 16384, 8960, 128); bd[i] = desc(b0 + (i & 1) * 8192, (N / 2) * 16, 128); }
    long long t0 = clock64();
    for (int it = 0; it < 64; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tc_mma2(tmem + (uint32_t)(w * 128), ad[i], bd[i], id, 1);
    }
    long long t1 = clock64();
    tc_commit2(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[2 * w] = t1 - t0; out[2 * w + 1] = t2 - t0; }
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N>
void run(long long* d_out, int nw) {
  cudaFuncSetAttribute(k_bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int r = 0; r < 2; ++r) k_bench<N><<<148, 128, 200 * 1024>>>(nw, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N %d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  long long h[8]; cudaMemcpy(h, d_out, 8 * 8, cudaMemcpyDeviceToHost);
  printf("N=%3d issuing warps=%d, 512 MMAs each: ", N, nw);
  for (int i = 0; i < nw; ++i) printf("[issue %.1f done %.1f cyc/MMA] ", h[2 * i] / 512.0, h[2 * i + 1] / 512.0);
  printf(" -> %.1f cycles per MMA overall\n", h[1] / 512.0 / nw);
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  for (int nw = 1; nw <= 4; nw *= 2) { run<32>(d_out, nw); run<64>(d_out, nw); run<128>(d_out, nw); }
  return 0;
}
