// Does alternating the MMA shape (N) or the accumulator cost issue slots?  Exact fold sequence of net_tc.cu.
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
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// PAT 0: fold order per tile [2N,2N,N,N]   PAT 1: all 2N first for both tiles then all N   PAT 2: only 2N   PAT 3: only N
// PAT 4: as 0 but a commit after every 8 MMAs
template <int PAT, int COUT>
__global__ void __launch_bounds__(128, 1) k_bench(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
    const uint32_t a0 = smem_u32(smem) + 24 * 16 + 16, b0 = smem_u32(smem) + 160 * 1024;
    constexpr uint32_t PLANE = 304 * 16;
    const uint32_t id1 = idesc_f16(128, COUT), id2 = idesc_f16(128, 2 * COUT);
    const uint64_t ad0 = desc(a0, PLANE, 128), bd0 = desc(b0, 2 * COUT * 16, 128);
    constexpr uint64_t A_TILE = 128, A_K16 = 2 * PLANE / 16, A_LO = 4 * PLANE / 16, W_K16 = 2 * (2 * COUT * 16) / 16;
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
      if (PAT == 0 || PAT == 4) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
          for (int k = 0; k < 2; ++k) tc_mma(tmem + m * 2 * COUT, ad0 + m * A_TILE + k * A_K16, bd0 + k * W_K16, id2, 1);
#pragma unroll
          for (int k = 0; k < 2; ++k) tc_mma(tmem + m * 2 * COUT, ad0 + m * A_TILE + k * A_K16 + A_LO, bd0 + k * W_K16, id1, 1);
        }
        if (PAT == 4) tc_commit(&bar2);
      } else if (PAT == 1) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int k = 0; k < 2; ++k) tc_mma(tmem + m * 2 * COUT, ad0 + m * A_TILE + k * A_K16, bd0 + k * W_K16, id2, 1);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int k = 0; k < 2; ++k) tc_mma(tmem + m * 2 * COUT, ad0 + m * A_TILE + k * A_K16 + A_LO, bd0 + k * W_K16, id1, 1);
      } else {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma(tmem + m * 2 * COUT, ad0 + m * A_TILE + (k & 1) * A_K16 + (k >> 1) * A_LO, bd0 + (k & 1) * W_K16, PAT == 2 ? id2 : id1, 1);
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
template <int PAT, int COUT>
void run(const char* name, long long* d_out) {
  cudaFuncSetAttribute(k_bench<PAT, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_bench<PAT, COUT><<<148, 128, 200 * 1024>>>(4000, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, d_out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("cout=%3d %-40s : %7.1f cycles/MMA\n", COUT, name, (double)mx / 4000);
}
int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 8);
  run<0, 64>("fold order [2N,2N,N,N] per tile", d_out);
  run<1, 64>("grouped [2N x4][N x4]", d_out);
  run<2, 64>("only 2N", d_out);
  run<3, 64>("only N", d_out);
  run<4, 64>("fold order + commit per 8", d_out);
  run<0, 32>("fold order [2N,2N,N,N] per tile", d_out);
  run<1, 32>("grouped [2N x4][N x4]", d_out);
  run<2, 32>("only 2N", d_out);
  run<3, 32>("only N", d_out);
  return 0;
}
