// Small-batch latency path of the policy/value net (A5_NET_SMALL): the whole forward of 1..8 boards as ONE
// persistent cooperative kernel, exact fp32 on the CUDA cores.
//
// Why: Player.get_action (player.py:128-147) searches one board at a time -- GUI.py:124-167, self_play.py:94-101,
// choose_best_player.py -- so a move is 500 strictly sequential leaf evaluations of ONE position.  The tensor-core
// path is built for 4096 boards: eleven launches of 148-CTA persistent kernels (TMEM allocation, cluster sync,
// resident weight sets) cost ~13 us each whatever the batch: 145 us per leaf.  Here the 58.7 MMAC of one board are
// spread over all SMs inside one launch and the layers are separated by grid barriers (~1 us) instead of launches.
//
// Work decomposition: a conv layer is an implicit GEMM over the padded position space (net_common.cuh), rows =
// positions, columns = output channels, K = taps x input channels (+ the 1x1 residual projection as one more K
// segment, network.py:52-56).  One *item* = 16 rows x 8 output channels; its 16 + 2 (pitch + 1) source rows and
// its 8 weight rows ([cout][K], K-contiguous copies of the fp32 path's matrices) are staged in shared memory with
// cp.async (L2 only: the rows were written by other SMs before the barrier); the 8 warps of the CTA split K
// (channel quads interleaved), every lane holds a 2 x 2 register tile, and the eight partial sums of an output are
// added in warp order (deterministic) before bias + ELU.  Independent layers share a phase (block3 / block4,
// network.py:68,79; the value head runs beside block5), the policy
// head conv is computed inside its dense layer's input staging, so a forward is 11 phases = 10 barriers.
#include "net.cuh"

namespace a5 {

constexpr int SM_THREADS = 256;
constexpr int SM_RT = 16, SM_CT = 8;
constexpr int SM_MAXH = A5_MAX_BOARD + 2;                 // halo rows on each side: pitch + 1
constexpr int SM_AROWS = SM_RT + 2 * SM_MAXH;
constexpr int SM_CMAX = 128, SM_KMAX = 9 * 128 + 128;
constexpr int SM_LDA = SM_CMAX + 4;
constexpr int SM_SMEM = (SM_AROWS * SM_LDA + SM_RT * SM_LDA + SM_CT * (SM_KMAX + 4) + 8 * 128) * 4;
constexpr int SM_PHASES = 11;

enum { SM_NONE = 0, SM_CONV, SM_CONV1, SM_HEADCONV, SM_DENSE, SM_FINAL };

struct SmTask {
  int type;
  const float* src; int cin;        // CONV: 3x3 source [row][cin]; HEADCONV: features [row][32]; DENSE: x [board][K]
  const float* res; int rcin;       // CONV: source of the 1x1 residual projection, or null
  const float* w; int ldw;          // CONV, DENSE: [cout][K]; CONV1: [75][ldw]; HEADCONV: [32][ldw]
  const float* bias;
  float* out; int cout, K, act, ldo;
  // DENSE with a fused producer: hw = 1x1 head-conv kernel [32][hldw], hb its bias -- the input vector is the head
  // conv (rcin channels) of the feature rows `res`, computed in place
  const float* hw; int hldw; const float* hb;
};
struct SmParams {
  SmTask t[SM_PHASES][2];
  const int8_t* planes;
  int n, S, C, pitch, per_board, guard;
  int nrows;
  const float* logits; int ldl;     // FINAL
  const float* h64; const float* w2; const float* b2;
  float* prob; float* value;
  unsigned long long* bar;
  unsigned long long* dbg;          // tooling: %globaltimer of CTA 0 at kernel start and after every phase, or null
};

__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Every CTA of the (co-resident) grid arrives once per phase with a fire-and-forget release reduction and polls
// the counter (one L2 round trip less than an atomic that returns, then a flag).  The counter only grows; CTA 0
// arrives with the weight SM_BAR - (gridDim.x - 1), so a barrier is worth SM_BAR and a launch
// SM_BAR * (SM_PHASES - 1) whatever the grid: the base of a launch is the counter rounded down to a multiple of that.
constexpr unsigned long long SM_BAR = 1ull << 20;
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long inc = blockIdx.x == 0 ? SM_BAR - (unsigned long long)(gridDim.x - 1) : 1ull;
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(inc) : "memory");
    while (ld_acquire(bar) < target) {}
  }
  __syncthreads();
}

__device__ __forceinline__ int sm_items(const SmTask& T, int n, int row_tiles) {
  switch (T.type) {
    case SM_CONV: return row_tiles * (T.cout / SM_CT);
    case SM_CONV1: case SM_HEADCONV: return row_tiles;
    case SM_DENSE: return n * ((T.cout + 7) / 8);
    case SM_FINAL: return n;
  }
  return 0;
}

__device__ __forceinline__ void fma4(float& acc, const float4& a, const float4& b) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
}

// conv1: 5x5, 3 -> 32, SAME, ELU (network.py:63) of one cell straight from the {0,1} int8 planes; lane = channel.
// Lanes 0..24 look up the 25 taps of the cell in the three planes, a ballot turns them into three warp-uniform tap
// masks, and only the set taps (the stones around the cell) cost a weight add.
__device__ __forceinline__ float sm_conv1_cell(const SmParams& P, const float* __restrict__ w, int ldw, float bias,
                                               int board, int y, int x, int lane) {
  const int S = P.S;
  const int8_t* pb = P.planes + (size_t)board * 3 * P.C;
  const int ky = lane / 5, kx = lane - ky * 5;
  const int yy = y + ky - 2, xx = x + kx - 2;
  const bool ok = lane < 25 && yy >= 0 && yy < S && xx >= 0 && xx < S;
  float acc = 0.0f;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
    unsigned m = __ballot_sync(FULL, ok && __ldg(pb + ci * P.C + yy * S + xx) != 0);
    while (m) {
      const int t = __ffs(m) - 1;
      m &= m - 1;
      acc += __ldg(w + (t * 3 + ci) * ldw + lane);
    }
  }
  return elu(acc + bias);
}
// conv1 rows [r0, r0 + 16) x 32 channels: one row per warp at a time
__device__ void sm_conv1_item(const SmParams& P, const SmTask& T, int item) {
  const int lane = threadIdx.x & 31;
  const float bco = __ldg(T.bias + lane);
#pragma unroll 1
  for (int rr = threadIdx.x >> 5; rr < SM_RT; rr += 8) {
    const int q = item * SM_RT + rr;
    if (q >= P.nrows) continue;
    const int board = q / P.per_board, within = q - board * P.per_board;
    const int y = within / P.pitch, x = within - y * P.pitch;
    const float v = (y < P.S && x < P.S) ? sm_conv1_cell(P, T.w, T.ldw, bco, board, y, x, lane) : 0.0f;
    T.out[(size_t)(P.guard + q) * 32 + lane] = v;
  }
}

// one item of a 3x3 (+1x1 residual) conv layer: rows [r0, r0 + 16) x channels [c0, c0 + 8)
__device__ void sm_conv_item(const SmParams& P, const SmTask& T, int item, float* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nct = T.cout / SM_CT;
  const int ct = item % nct, rt = item / nct;
  const int r0 = rt * SM_RT, c0 = ct * SM_CT;
  const int H = P.pitch + 1, cin = T.cin, rcin = T.res ? T.rcin : 0, K = T.K;
  const int lda = cin + 4, ldr = rcin + 4, ldb = K + 4;
  float* A_s = smem;
  float* R_s = A_s + SM_AROWS * SM_LDA;
  float* B_s = R_s + SM_RT * SM_LDA;
  float* P_s = B_s + SM_CT * (SM_KMAX + 4);
  {
    const int q4 = cin >> 2, arows = SM_RT + 2 * H;
    const float* g = T.src + (size_t)(P.guard + r0 - H) * cin;
    for (int i = tid; i < arows * q4; i += SM_THREADS) {
      const int rr = i / q4, c4 = i - rr * q4;
      cp16(A_s + rr * lda + 4 * c4, g + (size_t)rr * cin + 4 * c4);
    }
    if (rcin) {
      const int rq4 = rcin >> 2;
      const float* gr = T.res + (size_t)(P.guard + r0) * rcin;
      for (int i = tid; i < SM_RT * rq4; i += SM_THREADS) {
        const int rr = i / rq4, c4 = i - rr * rq4;
        cp16(R_s + rr * ldr + 4 * c4, gr + (size_t)rr * rcin + 4 * c4);
      }
    }
    const int k4 = K >> 2;
    const float* gw = T.w + (size_t)c0 * K;
    for (int i = tid; i < SM_CT * k4; i += SM_THREADS) {
      const int col = i / k4, kk = i - col * k4;
      cp16(B_s + col * ldb + 4 * kk, gw + (size_t)col * K + 4 * kk);
    }
    cp_wait_all();
  }
  __syncthreads();
  // lane = (row pair rb, rb + 8) x (column pair cb, cb + 4); warp = K slice
  const int rb = lane >> 2, cb = lane & 3;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const int nq = cin >> 2;
#pragma unroll 1
  for (int t = 0; t < 9; ++t) {
    const int shift = (t / 3 - 1) * P.pitch + (t % 3 - 1);
    const float* a0 = A_s + (rb + H + shift) * lda;
    const float* a1 = a0 + 8 * lda;
    const float* b0 = B_s + cb * ldb + t * cin;
    const float* b1 = b0 + 4 * ldb;
#pragma unroll 2
    for (int q = warp; q < nq; q += 8) {
      const float4 x0 = *(const float4*)(a0 + 4 * q), x1 = *(const float4*)(a1 + 4 * q);
      const float4 w0 = *(const float4*)(b0 + 4 * q), w1 = *(const float4*)(b1 + 4 * q);
      fma4(acc[0][0], x0, w0); fma4(acc[0][1], x0, w1);
      fma4(acc[1][0], x1, w0); fma4(acc[1][1], x1, w1);
    }
  }
  if (rcin) {
    const float* a0 = R_s + rb * ldr;
    const float* a1 = a0 + 8 * ldr;
    const float* b0 = B_s + cb * ldb + 9 * cin;
    const float* b1 = b0 + 4 * ldb;
    for (int q = warp; q < (rcin >> 2); q += 8) {
      const float4 x0 = *(const float4*)(a0 + 4 * q), x1 = *(const float4*)(a1 + 4 * q);
      const float4 w0 = *(const float4*)(b0 + 4 * q), w1 = *(const float4*)(b1 + 4 * q);
      fma4(acc[0][0], x0, w0); fma4(acc[0][1], x0, w1);
      fma4(acc[1][0], x1, w0); fma4(acc[1][1], x1, w1);
    }
  }
  *(float4*)(P_s + warp * 128 + lane * 4) = make_float4(acc[0][0], acc[0][1], acc[1][0], acc[1][1]);
  __syncthreads();
  if (tid < 128) {
    const int l2 = tid >> 2, ij = tid & 3;
    const int row = (l2 >> 2) + 8 * (ij >> 1), col = (l2 & 3) + 4 * (ij & 1);
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += P_s[w * 128 + tid];
    s += __ldg(T.bias + c0 + col);
    if (T.act) s = elu(s);
    const int q = r0 + row;
    if (q < P.nrows) {
      const int within = q % P.per_board;
      const int y = within / P.pitch, x = within - y * P.pitch;
      T.out[(size_t)(P.guard + q) * T.cout + c0 + col] = (y < P.S && x < P.S) ? s : 0.0f;
    }
  }
  __syncthreads();
}

// 1x1 head conv 32 -> cout (4 value, 16 policy) + ELU into the channel-major flat layout the dense layers read
// (network.py:69-71, 81-83): out[board][c * C + cell]
__device__ void sm_headconv_item(const SmParams& P, const SmTask& T, int item) {
  const int hc = T.cout, t = threadIdx.x;
  if (t >= SM_RT * hc) return;
  const int rr = t / hc, c = t - rr * hc;
  const int q = item * SM_RT + rr;
  if (q >= P.nrows) return;
  const int board = q / P.per_board, within = q - board * P.per_board;
  const int y = within / P.pitch, x = within - y * P.pitch;
  if (y >= P.S || x >= P.S) return;
  const float* f = T.src + (size_t)(P.guard + q) * 32;
  float acc = 0.0f;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) acc = fmaf(__ldcg(f + k), __ldg(T.w + k * T.ldw + c), acc);
  T.out[((size_t)board * hc + c) * P.C + y * P.S + x] = elu(acc + __ldg(T.bias + c));
}

// dense layer, 8 outputs of one board: warp = output, lanes along K (weights [cout][K], K-contiguous); the board's
// input vector is staged in shared memory (it was written by other SMs: L2 only)
__device__ void sm_dense_item(const SmParams& P, const SmTask& T, int item, float* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (T.cout + 7) / 8;
  const int b = item / nblk, j = (item - b * nblk) * 8 + warp;
  const int k4 = T.K >> 2;
  if (T.hw) {
    // policy head: x[c * C + cell] = ELU(1x1 conv 32 -> hc of the block-5 row of the cell) (network.py:81-83)
    const int hc = T.rcin;
    float* hws = smem + T.K;
    for (int i = threadIdx.x; i < 32 * hc; i += SM_THREADS) hws[i] = __ldg(T.hw + (i / hc) * T.hldw + (i % hc));
    __syncthreads();
    // thread = (cell, half of the hc output channels)
    for (int i = threadIdx.x; i < 2 * P.C; i += SM_THREADS) {
      const int half = i >= P.C, cell = i - half * P.C;
      const int y = cell / P.S, xx = cell - y * P.S;
      const float4* f4 = (const float4*)(T.res + (size_t)(P.guard + b * P.per_board + y * P.pitch + xx) * 32);
      float f[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float4 v = __ldcg(f4 + q); f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w; }
#pragma unroll 1
      for (int c4 = half * (hc >> 1); c4 < (half + 1) * (hc >> 1); c4 += 4) {
        float a0 = __ldg(T.hb + c4), a1 = __ldg(T.hb + c4 + 1), a2 = __ldg(T.hb + c4 + 2), a3 = __ldg(T.hb + c4 + 3);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float4 w = *(const float4*)(hws + k * hc + c4);
          a0 = fmaf(f[k], w.x, a0); a1 = fmaf(f[k], w.y, a1); a2 = fmaf(f[k], w.z, a2); a3 = fmaf(f[k], w.w, a3);
        }
        smem[(c4 + 0) * P.C + cell] = elu(a0); smem[(c4 + 1) * P.C + cell] = elu(a1);
        smem[(c4 + 2) * P.C + cell] = elu(a2); smem[(c4 + 3) * P.C + cell] = elu(a3);
      }
    }
  } else {
    const float* x = T.src + (size_t)b * T.K;
    for (int i = threadIdx.x; i < k4; i += SM_THREADS) cp16(smem + 4 * i, x + 4 * i);
    cp_wait_all();
  }
  __syncthreads();
  if (j < T.cout) {
    const float4* w4 = (const float4*)(T.w + (size_t)j * T.K);
    const float4* x4 = (const float4*)smem;
    float acc = 0.0f;
#pragma unroll 4
    for (int q = lane; q < k4; q += 32) fma4(acc, x4[q], __ldg(w4 + q));
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += __ldg(T.bias + j);
      T.out[(size_t)b * T.ldo + j] = T.act ? elu(acc) : acc;
    }
  }
  __syncthreads();
}

// softmax of the policy logits (network.py:163-165) and value = tanh(fc2(h) / 2) (network.py:75-76) of one board
__device__ void sm_final_item(const SmParams& P, int b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const float* row = P.logits + (size_t)b * P.ldl;
    float x[8], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = k * 32 + lane;
      x[k] = c < P.C ? __ldcg(row + c) : -INFINITY;
      mx = fmaxf(mx, x[k]);
    }
    mx = warp_max(mx);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[k] = (k * 32 + lane) < P.C ? expf(x[k] - mx) : 0.0f;
      s += x[k];
    }
    s = warp_sum(s);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = k * 32 + lane;
      if (c < P.C) P.prob[(size_t)b * P.C + c] = x[k] / s;
    }
  } else if (warp == 1) {
    float s = 0.0f;
    for (int j = lane; j < 64; j += 32) s += __ldcg(P.h64 + (size_t)b * 64 + j) * __ldg(P.w2 + j);
    s = warp_sum(s);
    if (lane == 0) P.value[b] = tanhf((s + __ldg(P.b2)) * 0.5f);
  }
}

__global__ void __launch_bounds__(SM_THREADS, 1) k_small_net(const __grid_constant__ SmParams P) {
  extern __shared__ __align__(16) float sm_smem[];
  __shared__ unsigned long long s_base;
  const unsigned long long per_launch = SM_BAR * (SM_PHASES - 1);
  if (threadIdx.x == 0) s_base = (ld_acquire(P.bar) / per_launch) * per_launch;
  __syncthreads();
  const unsigned long long base = s_base;
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[0] = kt_now();
  const int row_tiles = (P.nrows + SM_RT - 1) / SM_RT;
#pragma unroll 1
  for (int ph = 0; ph < SM_PHASES; ++ph) {
    const SmTask& T0 = P.t[ph][0];
    const SmTask& T1 = P.t[ph][1];
    const int n0 = sm_items(T0, P.n, row_tiles), n1 = sm_items(T1, P.n, row_tiles);
#pragma unroll 1
    for (int it = blockIdx.x; it < n0 + n1; it += gridDim.x) {
      const SmTask& T = it < n0 ? T0 : T1;
      const int item = it < n0 ? it : it - n0;
      switch (T.type) {
        case SM_CONV: sm_conv_item(P, T, item, sm_smem); break;
        case SM_CONV1: sm_conv1_item(P, T, item); break;
        case SM_HEADCONV: sm_headconv_item(P, T, item); break;
        case SM_DENSE: sm_dense_item(P, T, item, sm_smem); break;
        case SM_FINAL: sm_final_item(P, item); break;
      }
    }
    if (ph + 1 < SM_PHASES) grid_barrier(P.bar, base + (unsigned long long)(ph + 1) * SM_BAR);
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[1 + ph] = kt_now();
  }
}

// fp32 path matrix [K][ldw] -> [cout][K]
__global__ void k_small_transpose(const float* __restrict__ w, int K, int ldw, int cout, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * cout) return;
  const int c = i / K, k = i - c * K;
  out[i] = w[(size_t)k * ldw + c];
}

struct SmallState {
  float* wt[11] = {};
  float* wt_pfc = nullptr;     // [C][16 C]
  float* wt_vfc1 = nullptr;    // [64][4 C]
  float* vflat = nullptr;      // [B][4 C]
  float* h64 = nullptr;        // [B][64]
  unsigned long long* bar = nullptr;
  int num_sms = 0;
  bool cooperative = true;
};

static const int kSmCh[11] = {32, 64, 64, 128, 128, 32, 32, 64, 64, 32, 32};
enum { A32, B1H, B1O, B2H, B2O, B3H, B3O, B4H, B4O, B5H, B5O };
struct SmLayerDef { int src, res_src, out, cin, res_cin, cout; };
static const SmLayerDef kSmLayers[11] = {
    {0, 0, 0, 0, 0, 0},
    {A32, -1, B1H, 32, 0, 64},  {B1H, A32, B1O, 64, 32, 64},
    {B1O, -1, B2H, 64, 0, 128}, {B2H, B1O, B2O, 128, 64, 128},
    {B2O, -1, B3H, 128, 0, 32}, {B3H, B2O, B3O, 32, 128, 32},
    {B2O, -1, B4H, 128, 0, 64}, {B4H, B2O, B4O, 64, 128, 64},
    {B4O, -1, B5H, 64, 0, 32},  {B5H, B4O, B5O, 32, 64, 32}};

int small_alloc(a5_net* net, SmallState** out) {
  SmallState* s = new SmallState();
  *out = s;
  for (int l = 1; l <= 10; ++l) {
    const SmLayerDef& L = kSmLayers[l];
    A5_CUDA(cudaMalloc(&s->wt[l], (size_t)(9 * L.cin + L.res_cin) * L.cout * sizeof(float)));
  }
  const int nb = net->max_batch < A5_NET_SMALL_MAX ? net->max_batch : A5_NET_SMALL_MAX;
  A5_CUDA(cudaMalloc(&s->wt_pfc, (size_t)net->C * 16 * net->C * sizeof(float)));
  A5_CUDA(cudaMalloc(&s->wt_vfc1, (size_t)64 * 4 * net->C * sizeof(float)));
  A5_CUDA(cudaMalloc(&s->vflat, (size_t)nb * 4 * net->C * sizeof(float)));
  A5_CUDA(cudaMalloc(&s->h64, (size_t)nb * 64 * sizeof(float)));
  A5_CUDA(cudaMalloc(&s->bar, 32 * sizeof(unsigned long long)));
  A5_CUDA(cudaMemset(s->bar, 0, 32 * sizeof(unsigned long long)));
  int dev = 0;
  A5_CUDA(cudaGetDevice(&dev));
  A5_CUDA(cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, dev));
  A5_CUDA(cudaFuncSetAttribute(k_small_net, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
  return A5_OK;
}

void small_free(SmallState* s) {
  if (!s) return;
  for (int l = 1; l <= 10; ++l) cudaFree(s->wt[l]);
  cudaFree(s->wt_pfc); cudaFree(s->wt_vfc1);
  cudaFree(s->vflat); cudaFree(s->h64); cudaFree(s->bar);
  delete s;
}

// after fp32_set_weights: K-contiguous copies of its conv matrices
int small_set_weights(a5_net* net, SmallState* s, cudaStream_t st) {
  for (int l = 1; l <= 10; ++l) {
    const SmLayerDef& L = kSmLayers[l];
    const int K = 9 * L.cin + L.res_cin, ldw = (L.cout + 63) / 64 * 64;
    k_small_transpose<<<(K * L.cout + 255) / 256, 256, 0, st>>>(net->w[l], K, ldw, L.cout, s->wt[l]);
    A5_CUDA(cudaGetLastError());
  }
  const int C = net->C;
  k_small_transpose<<<(16 * C * C + 255) / 256, 256, 0, st>>>(net->w[12], 16 * C, net->ldl, C, s->wt_pfc);
  k_small_transpose<<<(4 * C * 64 + 255) / 256, 256, 0, st>>>(net->vfc1_w, 4 * C, 64, 64, s->wt_vfc1);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

static unsigned long long* g_small_dbg = nullptr;

int small_forward(a5_net* net, SmallState* s, const int8_t* planes, int n, float* prob, float* value, cudaStream_t st) {
  PosSpace ps(net->S);
  SmParams P;
  memset(&P, 0, sizeof(P));
  P.planes = planes; P.n = n; P.S = net->S; P.C = net->C; P.pitch = ps.pitch; P.per_board = ps.per_board; P.guard = ps.guard;
  P.nrows = n * ps.per_board;
  P.logits = net->logits; P.ldl = net->ldl; P.h64 = s->h64; P.w2 = net->vfc2_w; P.b2 = net->vfc2_b;
  P.prob = prob; P.value = value; P.bar = s->bar; P.dbg = g_small_dbg;
  auto conv = [&](int l) {
    const SmLayerDef& L = kSmLayers[l];
    SmTask t;
    memset(&t, 0, sizeof(t));
    t.type = SM_CONV; t.src = net->act[L.src]; t.cin = L.cin;
    t.res = L.res_src >= 0 ? net->act[L.res_src] : nullptr; t.rcin = L.res_cin;
    t.w = s->wt[l]; t.bias = net->bias[l]; t.out = net->act[L.out]; t.cout = L.cout; t.K = 9 * L.cin + L.res_cin; t.act = 1;
    return t;
  };
  auto headconv = [&](const float* src, const float* w, int ldw, const float* b, float* out, int hc) {
    SmTask t;
    memset(&t, 0, sizeof(t));
    t.type = SM_HEADCONV; t.src = src; t.w = w; t.ldw = ldw; t.bias = b; t.out = out; t.cout = hc; t.act = 1;
    return t;
  };
  auto dense = [&](const float* x, int K, const float* w, int ldw, const float* b, float* out, int cout, int ldo, int act) {
    SmTask t;
    memset(&t, 0, sizeof(t));
    t.type = SM_DENSE; t.src = x; t.K = K; t.w = w; t.ldw = ldw; t.bias = b; t.out = out; t.cout = cout; t.ldo = ldo; t.act = act;
    return t;
  };
  const int C = net->C;
  P.t[0][0].type = SM_CONV1; P.t[0][0].w = net->w[0]; P.t[0][0].ldw = 64; P.t[0][0].bias = net->bias[0];
  P.t[0][0].out = net->act[A32]; P.t[0][0].cout = 32;
  P.t[1][0] = conv(1); P.t[2][0] = conv(2); P.t[3][0] = conv(3); P.t[4][0] = conv(4);
  P.t[5][0] = conv(7); P.t[5][1] = conv(5);                 // block4-conv1 (the larger) first, block3-conv1 beside it
  P.t[6][0] = conv(8); P.t[6][1] = conv(6);
  P.t[7][0] = conv(9); P.t[7][1] = headconv(net->act[B3O], net->vconv_w, 4, net->vconv_b, s->vflat, 4);
  P.t[8][0] = conv(10); P.t[8][1] = dense(s->vflat, 4 * C, s->wt_vfc1, 0, net->vfc1_b, s->h64, 64, 64, 1);
  P.t[9][0] = dense(nullptr, 16 * C, s->wt_pfc, 0, net->bias[12], net->logits, C, net->ldl, 0);
  P.t[9][0].res = net->act[B5O]; P.t[9][0].rcin = 16;      // policy head conv fused into the dense layer's input staging
  P.t[9][0].hw = net->w[11]; P.t[9][0].hldw = 64; P.t[9][0].hb = net->bias[11];
  P.t[10][0].type = SM_FINAL;
  const int row_tiles = (P.nrows + SM_RT - 1) / SM_RT;
  int grid = row_tiles * (128 / SM_CT);                     // the widest phase (block2: 128 output channels)
  if (grid > s->num_sms) grid = s->num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(SM_THREADS);
  cfg.dynamicSmemBytes = SM_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;               // all CTAs co-resident: the grid barrier cannot deadlock
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = s->cooperative ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_small_net, P);
  if (e != cudaSuccess && s->cooperative) {
    // no cooperative launch here (e.g. under stream capture on this driver): the grid never exceeds the SM count
    // and a CTA needs less than half an SM, so a plain launch is co-resident as well
    (void)cudaGetLastError();
    s->cooperative = false;
    cfg.numAttrs = 0;
    e = cudaLaunchKernelEx(&cfg, k_small_net, P);
  }
  A5_CUDA(e);
  return A5_OK;
}

}  // namespace a5

// internal tooling (not part of alphafive.h): uint64 [1 + 11] device buffer receiving CTA 0's %globaltimer at the
// start of the next A5_NET_SMALL forwards and after each of their phases (null: off)
extern "C" int a5__debug_small_timeline(unsigned long long* d_buf) { a5::g_small_dbg = d_buf; return A5_OK; }
