// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per MMA for M=128, N in {32..256},
// SMEM operands, no-swizzle K-major vs 128B-swizzle K-major, one CTA per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// MODE 0: no-swizzle, A plane stride 8960 (as net_tc.cu), B lbo = N*16
// MODE 1: 128B swizzle both      MODE 2: A from TMEM, B no-swizzle
// MODE 3: no-swizzle, A compact (lbo = 2048)
template <int MODE, int nd, int run_len>
__global__ void __launch_bounds__(128, 1) k_bench(int N, int iters, int distinctA, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // 1.0h
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 160 * 1024;
    const uint32_t id = idesc_f16(128, N);
    uint64_t ad[8], bd[8]; uint32_t at[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ia = i % distinctA;
      if (MODE == 0) { ad[i] = desc(a0 + ia * 2048, 8960, 128, 0); bd[i] = desc(b0 + (i & 1) * 16384, N * 16, 128, 0); }
      else if (MODE == 1) { ad[i] = desc(a0 + ia * 16384 + (i & 3) * 32, 16, 1024, 2); bd[i] = desc(b0 + (i & 3) * 32, 16, 1024, 2); }
      else if (MODE == 4) { ad[i] = desc(a0 + 16 * (1 + ia), 8960, 128, 0); bd[i] = desc(b0 + (i & 1) * 16384, N * 16, 128, 0); }
      else if (MODE == 5) { uint32_t a = a0 + 128 * (1 + ia) + (i & 3) * 32; ad[i] = desc(a, 16, 1024, 2) | ((uint64_t)((a >> 7) & 7) << 49); bd[i] = desc(b0 + (i & 3) * 32, 16, 1024, 2); }
      else if (MODE == 6) { uint32_t a = a0 + 128 * (1 + ia) + (i & 3) * 32; ad[i] = desc(a, 16, 1024, 2); bd[i] = desc(b0 + (i & 3) * 32, 16, 1024, 2); }
      else if (MODE == 3) { ad[i] = desc(a0 + ia * 4096, 2048, 128, 0); bd[i] = desc(b0 + (i & 1) * 16384, N * 16, 128, 0); }
      else { ad[i] = 0; bd[i] = desc(b0 + (i & 1) * 16384, N * 16, 128, 0); }
      at[i] = tmem + 256 + 8 * ia;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t d = tmem + (uint32_t)(((i / run_len) % nd) * (256 / (nd > 2 ? 2 : 1)) % 512);
        if (MODE == 2) tc_mma_ts(d, at[i], bd[i], id, 1);
        else tc_mma(d, ad[i], bd[i], id, 1);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int nd, int rl>
void run2(long long* d_out) {
  cudaFuncSetAttribute(k_bench<0, nd, rl>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int N : {64, 128}) {
    k_bench<0, nd, rl><<<148, 128, 200 * 1024>>>(N, 4000, 8, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: %s\n", N, cudaGetErrorString(e)); exit(1); }
    long long h[148]; cudaMemcpy(h, d_out, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%3d distinct D=%d run_len=%d : %7.1f cycles/MMA\n", N, nd, rl, (double)mx / 4000);
  }
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  run2<1, 1>(d_out); run2<2, 1>(d_out); run2<2, 2>(d_out); run2<2, 4>(d_out); run2<4, 1>(d_out); run2<4, 2>(d_out);
  return 0;
}
