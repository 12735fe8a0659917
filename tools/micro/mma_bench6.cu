// cta_group::2 MMA rate vs A-operand start alignment (tap shifts move A by 16-byte rows): M = 256 over a CTA pair, N in {32..256}, SS operands, leader issues.
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
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_bench(int iters, long long* out, int aoff, int foldseq, int rnd) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) { uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    ((uint32_t*)smem)[i] = rnd ? ((h & 0x83ff83ffu) | 0x38003800u) : 0x3c003c00u; }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t a0 = smem_u32(smem) + 1024 + aoff, b0 = smem_u32(smem) + 160 * 1024;
    const uint32_t id = idesc_f16(256, N), idh = idesc_f16(256, N / 2);
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ad[i] = desc(a0 + i * 2048, 8960, 128); bd[i] = desc(b0 + (i & 1) * 8192, (N / 2) * 16, 128); }
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tc_mma2(tmem + (uint32_t)((i & 1) * 256), ad[i], bd[i], (foldseq && (i & 2)) ? idh : id, 1);
    }
    tc_commit2(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[blockIdx.x / 2] = t1 - t0; out[74 + blockIdx.x / 2] = (long long)(g1 - g0);
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N>
void run(long long* d_out, int aoff, int foldseq, int rnd, int iters, int reps) {
  cudaFuncSetAttribute(k_bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int r = 0; r < reps; ++r) k_bench<N><<<148, 128, 200 * 1024>>>(iters, d_out, aoff, foldseq, rnd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N %d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, d_out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 74; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("cta_group::2 M=256 N=%3d aoff=%4d foldseq=%d : %7.1f cycles/MMA (floor N/2 = %d)\n", N, aoff, foldseq, (double)mx / iters, N / 2);
  printf("   rnd=%d iters=%d reps=%d : %.1f ns/MMA -> %.0f MHz\n", rnd, iters, reps, (double)h[74] / iters, (double)h[0] / (double)h[74] * 1000.0);
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  run<128>(d_out, 16, 0, 0, 4000, 1);
  run<128>(d_out, 16, 0, 1, 4000, 1);
  run<128>(d_out, 16, 0, 0, 400000, 10);
  run<128>(d_out, 16, 0, 1, 400000, 10);
  run<256>(d_out, 16, 0, 1, 200000, 10);
  run<128>(d_out, 16, 1, 1, 400000, 10);
  run<64>(d_out, 16, 0, 1, 400000, 10);
  return 0;
}
