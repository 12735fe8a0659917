// Dense heads on the tensor cores (A5_NET_TC): policy fc + softmax (network.py:84-88) and value
// fc1 + ELU + fc2 + tanh(x/2) (network.py:72-76, 163-165) as one launch of k_tc_fc.
//
// A = head-conv activations written by the conv epilogues (net_tc.cu) as fp16 hi/lo, K ordered
// (cell, channel), one M tile = 128 boards: [mtile][stage][hi|lo][kchunk 4][128][8];
// B = dense weights re-ordered to the same K order: [stage][kchunk 4][hi|lo][N][8].
// Same 3-term split as the convs (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, fp32 TMEM accumulate).
// One CTA per (head, M tile): warp 0 bulk-copy producer, warp 1 MMA issuer, warps 2..5 epilogue
// (thread = board: softmax / fc2 run in-thread over the board's TMEM lane).
#include "net.cuh"
#include "tc_ptx.cuh"

namespace a5 {

constexpr int FC_STAGES = 6;                             // at most; the launch passes how many fit (227 KB)
constexpr int FC_A_BYTES = 2 * 4 * 128 * 16;            // 16 KB per stage (K = 32)
constexpr int FC_THREADS = 192;

struct FCHead {
  const __half* A;
  const __half* W;
  const float* bias;     // [N] logits bias (policy) / fc1 bias (value)
  const float* w2;       // value: fc2 kernel [64]
  const float* b2;       // value: fc2 bias [1]
  float* out;            // policy: prob [n][C]; value: [n]
  int nst;               // K stages of 32
  int N;                 // padded outputs (multiple of 16, <= 256)
  int C;                 // real outputs (policy) / 64 (value)
  int fold;              // hi*[Whi|Wlo] as one N = 2N MMA (2N <= 256)
  unsigned long long* kt; // tooling: in-situ kernel timing slot (common.cuh), or null
};

struct FCBarriers {
  uint64_t full[FC_STAGES], empty[FC_STAGES], t_full;
  uint32_t tmem_base, pad;
};

// `nring` stages of (A 16 KB + weights) are in flight per CTA: the kernel is bound by the L2 -> shared-memory stream of
// its 2 MB of operands, i.e. by bytes in flight over the ~2 us L2 latency -- six stages at 11x11 (192 KB), four at 15x15
__global__ void __launch_bounds__(FC_THREADS, 1) k_tc_fc(const __grid_constant__ FCHead P, const __grid_constant__ FCHead V,
                                                         int mtiles, int n, int nring) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_launch_dependents();
  kt_begin(P.kt);
  const bool is_value = (int)blockIdx.x >= mtiles;
  const FCHead& H = is_value ? V : P;
  const int mt = is_value ? blockIdx.x - mtiles : blockIdx.x;
  const int N = H.N;
  const uint32_t w_bytes = 4u * 2u * (uint32_t)N * 16u;
  const uint32_t stride = FC_A_BYTES + ((w_bytes + 1023u) & ~1023u);
  FCBarriers* B = (FCBarriers*)(smem + nring * stride);
  float* s_bias = (float*)((uint8_t*)B + 128);           // [256] bias, then [64] fc2 kernel, [1] fc2 bias
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < N; i += FC_THREADS) s_bias[i] = i < H.C ? H.bias[i] : 0.0f;
  if (is_value) {
    if (threadIdx.x < 64) s_bias[256 + threadIdx.x] = H.w2[threadIdx.x];
    if (threadIdx.x == 64) s_bias[256 + 64] = H.b2[0];
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < nring; ++i) { mbar_init(&B->full[i], 1); mbar_init(&B->empty[i], 1); }
    mbar_init(&B->t_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = B->tmem_base;

  if (warp == 0) {
    // ===================== producer =====================
    const __half* a_src = H.A + (size_t)mt * H.nst * (FC_A_BYTES / 2);
    const __half* w_src = H.W;
    int st = 0, ph = 0;
    pdl_wait();                                              // the conv layers that wrote the A operands are complete
    for (int k = 0; k < H.nst; ++k) {
      mbar_wait(&B->empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&B->full[st], FC_A_BYTES + w_bytes);
        bulk_g2s(smem + st * stride, a_src + (size_t)k * (FC_A_BYTES / 2), FC_A_BYTES, &B->full[st]);
        bulk_g2s(smem + st * stride + FC_A_BYTES, w_src + (size_t)k * (w_bytes / 2), w_bytes, &B->full[st]);
      }
      __syncwarp();
      if (++st == nring) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = instr_desc(128, N), idesc2 = instr_desc(128, 2 * N);
    const uint32_t w_lbo = 2u * (uint32_t)N * 16u;
    const uint64_t w_k16 = (uint64_t)(2u * w_lbo / 16u), w_lo16 = (uint64_t)N;
    constexpr uint64_t A_K16 = 2u * 2048u / 16u, A_LO = 4u * 2048u / 16u;
    int st = 0, ph = 0;
    for (int k = 0; k < H.nst; ++k) {
      mbar_wait(&B->full[st], ph);
      tc_fence_after();
      const uint64_t ad0 = smem_desc(smem_u32(smem + st * stride), 2048, 128);
      const uint64_t bd0 = smem_desc(smem_u32(smem + st * stride + FC_A_BYTES), w_lbo, 128);
      if (elect_one()) {
        if (H.fold) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) tc_mma(tmem, ad0 + kk * A_K16, bd0 + kk * w_k16, idesc2, (k | kk) != 0);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) tc_mma(tmem, ad0 + kk * A_K16 + A_LO, bd0 + kk * w_k16, idesc, 1u);
        } else {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              tc_mma(tmem, ad0 + kk * A_K16 + (pass == 1 ? A_LO : 0), bd0 + kk * w_k16 + (pass == 2 ? w_lo16 : 0), idesc,
                     (k | pass | kk) != 0);
        }
        tc_commit(&B->empty[st]);
        if (k == H.nst - 1) tc_commit(&B->t_full);
      }
      __syncwarp();
      if (++st == nring) { st = 0; ph ^= 1; }
    }
  } else {
    // ===================== epilogue: thread = board =====================
    const int quad = warp & 3;
    const int board = mt * 128 + quad * 32 + lane;
    const uint32_t t0 = tmem + ((uint32_t)(quad * 32) << 16);
    mbar_wait(&B->t_full, 0);
    tc_fence_after();
    auto load16 = [&](int c0, float (&x)[16]) {
      uint32_t v[16];
      tc_ld16(t0 + (uint32_t)c0, v);
      if (H.fold) {
        uint32_t v2[16];
        tc_ld16(t0 + (uint32_t)(N + c0), v2);
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
      } else {
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], OUT_SCALE, s_bias[c0 + j]);
    };
    if (!is_value) {
      // softmax over the C real columns (network.py:88): three sweeps over the TMEM lane
      float mx = -INFINITY;
      for (int c0 = 0; c0 < N; c0 += 16) {
        float x[16];
        load16(c0, x);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < H.C) mx = fmaxf(mx, x[j]);
      }
      float sum = 0.0f;
      for (int c0 = 0; c0 < N; c0 += 16) {
        float x[16];
        load16(c0, x);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < H.C) sum += expf(x[j] - mx);
      }
      const float inv = 1.0f / sum;
      float* row = H.out + (size_t)board * H.C;
      for (int c0 = 0; c0 < N; c0 += 16) {
        float x[16];
        load16(c0, x);
        if (board < n) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < H.C) row[c0 + j] = expf(x[j] - mx) * inv;
        }
      }
    } else {
      float s = s_bias[256 + 64];
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float x[16];
        load16(c0, x);
#pragma unroll
        for (int j = 0; j < 16; ++j) s = fmaf(elu(x[j]), s_bias[256 + c0 + j], s);
      }
      if (board < n) H.out[board] = tanhf(s * 0.5f);
    }
  }
  tc_fence_before();
  __syncthreads();
  kt_end(P.kt);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
  }
}

// TF dense kernel [hc*C][nout] (row = c*C + cell, network.py flatten of NCHW) -> stages
// [stage][kchunk 4][hi|lo][N][8] with K re-ordered to k = cell*hc + c, scaled by 2^10, zero padded.
__global__ void k_fc_pack(const float* __restrict__ w, int hc, int C, int nout, int N, int nst, __half* __restrict__ out) {
  const long long total = (long long)nst * 32 * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int k = (int)(i / N);                 // 0 .. nst*32
    const int stage = k >> 5, kc = (k & 31) >> 3, e = k & 7;
    const int cell = k / hc, c = k - cell * hc;
    float x = 0.0f;
    if (cell < C && n < nout) x = w[((size_t)c * C + cell) * nout + n] * W_SCALE;
    const __half h = __float2half_rn(x);
    __half* base = out + (((size_t)stage * 4 + kc) * 2) * N * 8;
    base[((size_t)0 * N + n) * 8 + e] = h;
    base[((size_t)1 * N + n) * 8 + e] = __float2half_rn(x - __half2float(h));
  }
}

// ------------------------------------------------------------------ host side
struct HeadsState {
  __half* a_pol = nullptr; __half* a_val = nullptr;    // head-conv outputs (A operands)
  __half* w_pol = nullptr; __half* w_val = nullptr;
  float* pconv_w = nullptr; float* pconv_b = nullptr;   // 1x1 32->16 kernel [32][16] and bias
  int nst_pol = 0, nst_val = 0, n_pol = 0, mtiles = 0, smem = 0, nring = 0;
};

int heads_alloc(a5_net* net, HeadsState** out) {
  HeadsState* h = new HeadsState();
  *out = h;
  const int C = net->C;
  h->mtiles = (net->max_batch + 127) / 128;
  h->nst_pol = (16 * C + 31) / 32;
  h->nst_val = (4 * C + 31) / 32;
  h->n_pol = (C + 15) / 16 * 16;
  const size_t ab_pol = (size_t)h->mtiles * h->nst_pol * FC_A_BYTES, ab_val = (size_t)h->mtiles * h->nst_val * FC_A_BYTES;
  A5_CUDA(cudaMalloc(&h->a_pol, ab_pol));
  A5_CUDA(cudaMemset(h->a_pol, 0, ab_pol));             // K padding stays zero
  A5_CUDA(cudaMalloc(&h->a_val, ab_val));
  A5_CUDA(cudaMemset(h->a_val, 0, ab_val));
  A5_CUDA(cudaMalloc(&h->w_pol, (size_t)h->nst_pol * 4 * 2 * h->n_pol * 16));
  A5_CUDA(cudaMalloc(&h->w_val, (size_t)h->nst_val * 4 * 2 * 64 * 16));
  A5_CUDA(cudaMalloc(&h->pconv_w, 32 * 16 * 4));
  A5_CUDA(cudaMalloc(&h->pconv_b, 16 * 4));
  const uint32_t wb = 4u * 2u * (uint32_t)h->n_pol * 16u;
  const int stride = FC_A_BYTES + (int)((wb + 1023u) & ~1023u), misc = 128 + (256 + 64 + 4) * 4 + 128;
  h->nring = (232448 - misc) / stride;
  if (h->nring > FC_STAGES) h->nring = FC_STAGES;
  h->smem = h->nring * stride + misc;
  A5_CUDA(cudaFuncSetAttribute(k_tc_fc, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
  return A5_OK;
}

void heads_free(HeadsState* h) {
  if (!h) return;
  cudaFree(h->a_pol); cudaFree(h->a_val); cudaFree(h->w_pol); cudaFree(h->w_val);
  cudaFree(h->pconv_w); cudaFree(h->pconv_b);
  delete h;
}

int heads_set_weights(a5_net* net, HeadsState* h, const float* const* t, cudaStream_t st) {
  const int C = net->C;
  k_fc_pack<<<256, 256, 0, st>>>(t[T_PFC_K], 16, C, C, h->n_pol, h->nst_pol, h->w_pol);
  A5_CUDA(cudaGetLastError());
  k_fc_pack<<<64, 256, 0, st>>>(t[T_VFC1_K], 4, C, 64, 64, h->nst_val, h->w_val);
  A5_CUDA(cudaGetLastError());
  A5_CUDA(cudaMemcpyAsync(h->pconv_w, t[T_PCONV_K], 32 * 16 * 4, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemcpyAsync(h->pconv_b, t[T_PCONV_B], 16 * 4, cudaMemcpyDeviceToDevice, st));
  return A5_OK;
}

HeadsIO heads_io(const HeadsState* h) { return HeadsIO{h->a_pol, h->a_val, h->pconv_w, h->pconv_b, h->nst_pol, h->nst_val}; }

int heads_forward(a5_net* net, HeadsState* h, int n, float* prob, float* value, cudaStream_t st) {
  FCHead P, V;
  memset(&P, 0, sizeof(P));
  memset(&V, 0, sizeof(V));
  P.A = h->a_pol; P.W = h->w_pol; P.bias = net->bias[12]; P.out = prob;
  P.nst = h->nst_pol; P.N = h->n_pol; P.C = net->C; P.fold = 2 * h->n_pol <= 256;
  V.A = h->a_val; V.W = h->w_val; V.bias = net->vfc1_b; V.w2 = net->vfc2_w; V.b2 = net->vfc2_b; V.out = value;
  V.nst = h->nst_val; V.N = 64; V.C = 64; V.fold = 1;
  P.kt = kt_slot(KT_HEADS);
  const int mtiles = (n + 127) / 128;
  A5_CUDA(launch_pdl_k(k_tc_fc, (unsigned)(2 * mtiles), FC_THREADS, (size_t)h->smem, st, pdl_enabled(), P, V, mtiles, n, h->nring));
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

}  // namespace a5
