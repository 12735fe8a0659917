// fp32 CUDA-core forward pass of the policy/value net (A5_NET_FP32).
//
// Exact-fp32 companion of the tensor-core path: every conv is an implicit GEMM over
// the padded position space (net_common.cuh) -- rows = positions, columns = output
// channels, K = taps x input channels, where a tap is a constant row offset -- so the
// kernel is a plain tiled SGEMM whose A rows are fetched at (row + shift).  The 1x1
// projection of a residual block is folded into conv2 as one more K segment reading
// the block input (network.py:52-56), so the skip-add costs nothing.
#include "net.cuh"

namespace a5 {

constexpr int BK = 16;
constexpr int MAXSEG = 10;

struct Seg { const float* src; int lda; int shift; int kc; };

struct GemmArgs {
  Seg seg[MAXSEG];
  int nseg;
  const float* W; int ldw;
  const float* bias;
  float* out; int ldo; int ncols;
  long long row0, nrows;
  int mode;     // 0: conv in padded space (pad rows written as 0)  1: conv -> [board][col][cell] flat
                // 2: dense rows = boards
  int act;      // 1 = ELU
  int S, pitch, per_board, guard, C;
};

template <int BN>
__global__ void __launch_bounds__(256) k_gemm(const __grid_constant__ GemmArgs A) {
  constexpr int TN = BN / 4;          // threads along N
  constexpr int TM = 256 / TN;        // threads along M
  constexpr int BM = TM * 8;
  constexpr int A_F4 = BM * BK / 4 / 256;           // float4 per thread for the A tile
  constexpr int B_F4 = (BK * BN / 4 + 255) / 256;   // float4 per thread for the B tile
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tn = tid % TN, tm = tid / TN;
  const long long m0 = A.row0 + (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long row_end = A.row0 + A.nrows;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  int nk = 0;
  for (int s = 0; s < A.nseg; ++s) nk += A.seg[s].kc / BK;

  float4 ra[A_F4], rb[B_F4];
  int seg = 0, kin = 0, krow = 0;    // iterator over K tiles: segment, offset inside it, packed weight row
  auto load_tile = [&]() {
    const Seg& sg = A.seg[seg];
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int idx = tid + i * 256;
      int row = idx % BM, kq = idx / BM;
      long long r = m0 + row;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < row_end) ra[i] = *(const float4*)(sg.src + (r + sg.shift) * sg.lda + kin + kq * 4);
    }
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int idx = tid + i * 256;
      if (idx < BK * BN / 4) {
        int k = idx / (BN / 4), nq = idx % (BN / 4);
        rb[i] = *(const float4*)(A.W + (size_t)(krow + k) * A.ldw + n0 + nq * 4);
      }
    }
    kin += BK; krow += BK;
    if (kin >= sg.kc) { kin = 0; ++seg; }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int idx = tid + i * 256;
      int row = idx % BM, kq = idx / BM;
      As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
      As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int idx = tid + i * 256;
      if (idx < BK * BN / 4) {
        int k = idx / (BN / 4), nq = idx % (BN / 4);
        *(float4*)&Bs[buf][k][nq * 4] = rb[i];
      }
    }
  };

  load_tile();
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_tile();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *(const float4*)&As[cur][k][tm * 8];
      float4 a1 = *(const float4*)&As[cur][k][tm * 8 + 4];
      float4 b = *(const float4*)&Bs[cur][k][tn * 4];
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(cur ^ 1);
    __syncthreads();
  }

  // epilogue
  const int col = n0 + tn * 4;
  if (col >= A.ncols) return;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = (col + j < A.ncols) ? A.bias[col + j] : 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long r = m0 + tm * 8 + i;
    if (r >= row_end) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] + bv[j];
      if (A.act) v[j] = elu(v[j]);
    }
    if (A.mode == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col + j < A.ncols) A.out[r * A.ldo + col + j] = v[j];
      continue;
    }
    long long q = r - A.guard;
    int board = (int)(q / A.per_board), within = (int)(q % A.per_board);
    int rr = within / A.pitch, cc = within % A.pitch;
    bool real = rr < A.S && cc < A.S;
    if (A.mode == 0) {
      float4 o = real ? make_float4(v[0], v[1], v[2], v[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      *(float4*)(A.out + r * A.ldo + col) = o;       // ncols % 4 == 0 for every conv layer
    } else if (real) {
      int cell = rr * A.S + cc;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col + j < A.ncols) A.out[((size_t)board * A.ncols + col + j) * A.C + cell] = v[j];
    }
  }
}

// conv1: 5x5, 3 -> 32, SAME, ELU (network.py:63) straight from the int8 planes.
// One CTA per board; also writes the zero pad cells of the board's slab.
__global__ void __launch_bounds__(128) k_conv1(const int8_t* __restrict__ planes, const float* __restrict__ W,
                                              const float* __restrict__ bias, float* __restrict__ out, int S,
                                              int pitch, int per_board, int guard) {
  __shared__ float sw[75 * 32];
  __shared__ int8_t sp[3][20][20];
  const int b = blockIdx.x, tid = threadIdx.x, C = S * S;
  for (int i = tid; i < 75 * 32; i += 128) sw[i] = W[(i / 32) * 64 + (i % 32)];   // packed ldw = 64
  for (int i = tid; i < 3 * 20 * 20; i += 128) ((int8_t*)sp)[i] = 0;
  __syncthreads();
  for (int i = tid; i < 3 * C; i += 128) {
    int ch = i / C, cell = i % C;
    sp[ch][cell / S + 2][cell % S + 2] = planes[(size_t)b * 3 * C + i];
  }
  __syncthreads();
  const int co = tid & 31;
  const float bco = bias[co];
  float* ob = out + ((size_t)guard + (size_t)b * per_board) * 32;
  for (int pos = tid >> 5; pos < per_board; pos += 4) {
    int rr = pos / pitch, cc = pos % pitch;
    float v = 0.0f;
    if (rr < S && cc < S) {
      float acc = 0.0f;
      for (int ky = 0; ky < 5; ++ky)
        for (int kx = 0; kx < 5; ++kx)
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
            if (sp[ci][rr + ky][cc + kx]) acc += sw[((ky * 5 + kx) * 3 + ci) * 32 + co];
      v = elu(acc + bco);
    }
    ob[(size_t)pos * 32 + co] = v;
  }
}

// Value head after block3 (network.py:69-76): 1x1 conv 32->4 + ELU, flatten c-major,
// dense 4C->64 + ELU, dense 64->1, tanh(x/2).  One CTA per board.
__global__ void __launch_bounds__(128) k_value_head(const float* __restrict__ feat, const float* __restrict__ cw,
                                                   const float* __restrict__ cb, const float* __restrict__ w1,
                                                   const float* __restrict__ b1, const float* __restrict__ w2,
                                                   const float* __restrict__ b2, float* __restrict__ value, int S,
                                                   int pitch, int per_board, int guard) {
  __shared__ float sfeat[4 * 256];
  __shared__ float scw[32 * 4];
  __shared__ float part[2][64];
  const int b = blockIdx.x, tid = threadIdx.x, C = S * S;
  if (tid < 128) scw[tid] = cw[tid];
  __syncthreads();
  const float* fb = feat + ((size_t)guard + (size_t)b * per_board) * 32;
  for (int i = tid; i < 4 * C; i += 128) {
    int cell = i >> 2, c = i & 3;
    int pos = (cell / S) * pitch + cell % S;
    float acc = 0.0f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) acc = fmaf(fb[(size_t)pos * 32 + k], scw[k * 4 + c], acc);
    sfeat[c * C + cell] = elu(acc + cb[c]);
  }
  __syncthreads();
  {
    int j = tid & 63, half = tid >> 6;
    int K = 4 * C, k0 = half * (K / 2), k1 = half ? K : K / 2;
    float acc = 0.0f;
    for (int k = k0; k < k1; ++k) acc = fmaf(sfeat[k], w1[(size_t)k * 64 + j], acc);
    part[half][j] = acc;
  }
  __syncthreads();
  if (tid < 32) {
    float s = 0.0f;
    for (int j = tid; j < 64; j += 32) s += elu(part[0][j] + part[1][j] + b1[j]) * w2[j];
    s = warp_sum(s);
    if (tid == 0) value[b] = tanhf((s + b2[0]) * 0.5f);
  }
}

__global__ void __launch_bounds__(128) k_softmax(const float* __restrict__ logits, int ld, int C, int n,
                                                float* __restrict__ prob) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + w;
  if (b >= n) return;
  const float* row = logits + (size_t)b * ld;
  float x[8];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = k * 32 + lane;
    x[k] = c < C ? row[c] : -INFINITY;
    mx = fmaxf(mx, x[k]);
  }
  mx = warp_max(mx);
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    x[k] = (k * 32 + lane) < C ? expf(x[k] - mx) : 0.0f;
    s += x[k];
  }
  s = warp_sum(s);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = k * 32 + lane;
    if (c < C) prob[(size_t)b * C + c] = x[k] / s;
  }
}

// --------------------------------------------------------------------------- //
static const int kActCh[11] = {32, 64, 64, 128, 128, 32, 32, 64, 64, 32, 32};
enum { A32, B1H, B1O, B2H, B2O, B3H, B3O, B4H, B4O, B5H, B5O };

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// layer l (1..10): block (l-1)/2, conv1 if odd else conv2(+res)
struct LayerDef { int src, res_src, out, cin, res_cin, cout; };
static const LayerDef kLayers[11] = {
    {0, 0, 0, 0, 0, 0},
    {A32, -1, B1H, 32, 0, 64},  {B1H, A32, B1O, 64, 32, 64},
    {B1O, -1, B2H, 64, 0, 128}, {B2H, B1O, B2O, 128, 64, 128},
    {B2O, -1, B3H, 128, 0, 32}, {B3H, B2O, B3O, 32, 128, 32},
    {B2O, -1, B4H, 128, 0, 64}, {B4H, B2O, B4O, 64, 128, 64},
    {B4O, -1, B5H, 64, 0, 32},  {B5H, B4O, B5O, 32, 64, 32}};

int fp32_alloc(a5_net* net) {
  PosSpace ps(net->S);
  const size_t rows = (size_t)ps.rows(net->max_batch);
  for (int i = 0; i < 11; ++i) {
    A5_CUDA(cudaMalloc(&net->act[i], rows * kActCh[i] * sizeof(float)));
    A5_CUDA(cudaMemset(net->act[i], 0, rows * kActCh[i] * sizeof(float)));   // guard bands stay zero forever
  }
  const int C = net->C;
  A5_CUDA(cudaMalloc(&net->pflat, (size_t)net->max_batch * 16 * C * sizeof(float)));
  net->ldl = round_up(C, 64);
  A5_CUDA(cudaMalloc(&net->logits, (size_t)net->max_batch * net->ldl * sizeof(float)));
  // packed weights
  auto wl = [&](int l, int ktot, int cout) -> int {
    size_t n = (size_t)ktot * round_up(cout, 64);
    A5_CUDA(cudaMalloc(&net->w[l], n * sizeof(float)));
    A5_CUDA(cudaMemset(net->w[l], 0, n * sizeof(float)));
    A5_CUDA(cudaMalloc(&net->bias[l], round_up(cout, 64) * sizeof(float)));
    A5_CUDA(cudaMemset(net->bias[l], 0, round_up(cout, 64) * sizeof(float)));
    return A5_OK;
  };
  int rc;
  if ((rc = wl(0, 75, 32))) return rc;
  for (int l = 1; l <= 10; ++l)
    if ((rc = wl(l, 9 * kLayers[l].cin + kLayers[l].res_cin, kLayers[l].cout))) return rc;
  if ((rc = wl(11, 32, 16))) return rc;
  if ((rc = wl(12, 16 * C, C))) return rc;
  A5_CUDA(cudaMalloc(&net->vconv_w, 128 * 4)); A5_CUDA(cudaMalloc(&net->vconv_b, 4 * 4));
  A5_CUDA(cudaMalloc(&net->vfc1_w, (size_t)4 * C * 64 * 4)); A5_CUDA(cudaMalloc(&net->vfc1_b, 64 * 4));
  A5_CUDA(cudaMalloc(&net->vfc2_w, 64 * 4)); A5_CUDA(cudaMalloc(&net->vfc2_b, 4));
  return A5_OK;
}

void fp32_free(a5_net* net) {
  for (int i = 0; i < 11; ++i) cudaFree(net->act[i]);
  for (int i = 0; i < 13; ++i) { cudaFree(net->w[i]); cudaFree(net->bias[i]); }
  cudaFree(net->pflat); cudaFree(net->logits);
  cudaFree(net->vconv_w); cudaFree(net->vconv_b); cudaFree(net->vfc1_w); cudaFree(net->vfc1_b);
  cudaFree(net->vfc2_w); cudaFree(net->vfc2_b);
}

__global__ void k_add_bias(float* dst, const float* a, const float* b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a[i] + (b ? b[i] : 0.0f);
}

// TF layouts are already K-major x N ([kh][kw][cin][cout] = [(tap, cin)][cout]; dense [in][out]),
// so packing is a pitched copy into the N-padded matrix; the res kernel is appended as K rows.
int fp32_set_weights(a5_net* net, const float* const* t, cudaStream_t st) {
  const int C = net->C;
  auto put = [&](float* dst, int ldw, int row0, const float* src, int rows, int cols) -> int {
    A5_CUDA(cudaMemcpy2DAsync(dst + (size_t)row0 * ldw, (size_t)ldw * 4, src, (size_t)cols * 4, (size_t)cols * 4, rows,
                              cudaMemcpyDeviceToDevice, st));
    return A5_OK;
  };
  int rc;
  if ((rc = put(net->w[0], 64, 0, t[T_CONV1_K], 75, 32))) return rc;
  k_add_bias<<<1, 64, 0, st>>>(net->bias[0], t[T_CONV1_B], nullptr, 32);
  for (int l = 1; l <= 10; ++l) {
    const LayerDef& L = kLayers[l];
    const int blk = (l - 1) / 2, t0 = kBlocks[blk].t0;
    const int ldw = round_up(L.cout, 64);
    if (l & 1) {   // conv1 of the block
      if ((rc = put(net->w[l], ldw, 0, t[t0 + 2], 9 * L.cin, L.cout))) return rc;
      k_add_bias<<<1, 128, 0, st>>>(net->bias[l], t[t0 + 3], nullptr, L.cout);
    } else {       // conv2 + res
      if ((rc = put(net->w[l], ldw, 0, t[t0 + 4], 9 * L.cin, L.cout))) return rc;
      if ((rc = put(net->w[l], ldw, 9 * L.cin, t[t0 + 0], L.res_cin, L.cout))) return rc;
      k_add_bias<<<1, 128, 0, st>>>(net->bias[l], t[t0 + 5], t[t0 + 1], L.cout);
    }
  }
  if ((rc = put(net->w[11], 64, 0, t[T_PCONV_K], 32, 16))) return rc;
  k_add_bias<<<1, 64, 0, st>>>(net->bias[11], t[T_PCONV_B], nullptr, 16);
  if ((rc = put(net->w[12], net->ldl, 0, t[T_PFC_K], 16 * C, C))) return rc;
  k_add_bias<<<(C + 127) / 128, 128, 0, st>>>(net->bias[12], t[T_PFC_B], nullptr, C);
  A5_CUDA(cudaMemcpyAsync(net->vconv_w, t[T_VCONV_K], 128 * 4, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemcpyAsync(net->vconv_b, t[T_VCONV_B], 16, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemcpyAsync(net->vfc1_w, t[T_VFC1_K], (size_t)4 * C * 64 * 4, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemcpyAsync(net->vfc1_b, t[T_VFC1_B], 256, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemcpyAsync(net->vfc2_w, t[T_VFC2_K], 256, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemcpyAsync(net->vfc2_b, t[T_VFC2_B], 4, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

template <int BN>
static int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  constexpr int BM = (256 / (BN / 4)) * 8;
  dim3 grid((unsigned)((a.nrows + BM - 1) / BM), (unsigned)((a.ncols + BN - 1) / BN));
  k_gemm<BN><<<grid, 256, 0, st>>>(a);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

int fp32_forward(a5_net* net, const int8_t* planes, int n, float* prob, float* value, cudaStream_t st) {
  PosSpace ps(net->S);
  const int S = net->S, C = net->C;
  k_conv1<<<n, 128, 0, st>>>(planes, net->w[0], net->bias[0], net->act[A32], S, ps.pitch, ps.per_board, ps.guard);
  A5_CUDA(cudaGetLastError());
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.S = S; a.pitch = ps.pitch; a.per_board = ps.per_board; a.guard = ps.guard; a.C = C;
  int rc;
  for (int l = 1; l <= 10; ++l) {
    const LayerDef& L = kLayers[l];
    a.nseg = 0;
    for (int ky = -1; ky <= 1; ++ky)
      for (int kx = -1; kx <= 1; ++kx)
        a.seg[a.nseg++] = Seg{net->act[L.src], L.cin, ky * ps.pitch + kx, L.cin};
    if (L.res_src >= 0) a.seg[a.nseg++] = Seg{net->act[L.res_src], L.res_cin, 0, L.res_cin};
    a.W = net->w[l]; a.ldw = round_up(L.cout, 64); a.bias = net->bias[l];
    a.out = net->act[L.out]; a.ldo = L.cout; a.ncols = L.cout;
    a.row0 = ps.guard; a.nrows = (long long)n * ps.per_board;
    a.mode = 0; a.act = 1;
    rc = L.cout == 32 ? launch_gemm<32>(a, st) : launch_gemm<64>(a, st);
    if (rc) return rc;
  }
  return fp32_heads(net, n, prob, value, st);
}

// Heads on the fp32 [row][32] outputs of block3 (value) and block5 (policy); shared by both
// compute paths (the tensor-core epilogue of those two layers also emits fp32 rows).
int fp32_heads(a5_net* net, int n, float* prob, float* value, cudaStream_t st) {
  PosSpace ps(net->S);
  const int S = net->S, C = net->C;
  int rc;
  k_value_head<<<n, 128, 0, st>>>(net->act[B3O], net->vconv_w, net->vconv_b, net->vfc1_w, net->vfc1_b,
                                 net->vfc2_w, net->vfc2_b, value, S, ps.pitch, ps.per_board, ps.guard);
  A5_CUDA(cudaGetLastError());
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.S = S; a.pitch = ps.pitch; a.per_board = ps.per_board; a.guard = ps.guard; a.C = C;
  // policy head: 1x1 conv 32->16 + ELU into the c-major flat layout, dense, softmax
  a.nseg = 1; a.seg[0] = Seg{net->act[B5O], 32, 0, 32};
  a.W = net->w[11]; a.ldw = 64; a.bias = net->bias[11];
  a.out = net->pflat; a.ldo = 0; a.ncols = 16; a.mode = 1; a.act = 1;
  a.row0 = ps.guard; a.nrows = (long long)n * ps.per_board;
  if ((rc = launch_gemm<32>(a, st))) return rc;
  a.nseg = 1; a.seg[0] = Seg{net->pflat, 16 * C, 0, 16 * C};
  a.W = net->w[12]; a.ldw = net->ldl; a.bias = net->bias[12];
  a.out = net->logits; a.ldo = net->ldl; a.ncols = C; a.mode = 2; a.act = 0;
  a.row0 = 0; a.nrows = n;
  if ((rc = launch_gemm<64>(a, st))) return rc;
  k_softmax<<<(n + 3) / 4, 128, 0, st>>>(net->logits, net->ldl, C, n, prob);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

}  // namespace a5

// --------------------------------------------------------------------------- //
using namespace a5;

static const char* kTensorNames[A5_NET_NUM_TENSORS] = {
    "bone/conv1/kernel", "bone/conv1/bias",
    "bone/block1_res/kernel", "bone/block1_res/bias", "bone/block1_conv1/kernel", "bone/block1_conv1/bias",
    "bone/block1_conv2/kernel", "bone/block1_conv2/bias",
    "bone/block2_res/kernel", "bone/block2_res/bias", "bone/block2_conv1/kernel", "bone/block2_conv1/bias",
    "bone/block2_conv2/kernel", "bone/block2_conv2/bias",
    "value/block3_res/kernel", "value/block3_res/bias", "value/block3_conv1/kernel", "value/block3_conv1/bias",
    "value/block3_conv2/kernel", "value/block3_conv2/bias",
    "value/conv/kernel", "value/conv/bias", "value/fc1/kernel", "value/fc1/bias", "value/fc2/kernel", "value/fc2/bias",
    "policy/block4_res/kernel", "policy/block4_res/bias", "policy/block4_conv1/kernel", "policy/block4_conv1/bias",
    "policy/block4_conv2/kernel", "policy/block4_conv2/bias",
    "policy/block5_res/kernel", "policy/block5_res/bias", "policy/block5_conv1/kernel", "policy/block5_conv1/bias",
    "policy/block5_conv2/kernel", "policy/block5_conv2/bias",
    "policy/conv/kernel", "policy/conv/bias", "policy/fc/kernel", "policy/fc/bias"};

extern "C" {

const char* a5_net_tensor_name(int i) { return (i >= 0 && i < A5_NET_NUM_TENSORS) ? kTensorNames[i] : nullptr; }

int64_t a5_net_tensor_size(int i, int S) {
  const int64_t C = (int64_t)S * S;
  if (i < 0 || i >= A5_NET_NUM_TENSORS) return -1;
  if (i == T_CONV1_K) return 75 * 32;
  if (i == T_CONV1_B) return 32;
  for (int b = 0; b < 5; ++b) {
    int t0 = kBlocks[b].t0, ci = kBlocks[b].cin, co = kBlocks[b].cout;
    if (i == t0) return (int64_t)ci * co;
    if (i == t0 + 2) return 9LL * ci * co;
    if (i == t0 + 4) return 9LL * co * co;
    if (i == t0 + 1 || i == t0 + 3 || i == t0 + 5) return co;
  }
  switch (i) {
    case T_VCONV_K: return 32 * 4;
    case T_VCONV_B: return 4;
    case T_VFC1_K: return 4 * C * 64;
    case T_VFC1_B: return 64;
    case T_VFC2_K: return 64;
    case T_VFC2_B: return 1;
    case T_PCONV_K: return 32 * 16;
    case T_PCONV_B: return 16;
    case T_PFC_K: return 16 * C * C;
    case T_PFC_B: return C;
  }
  return -1;
}

int a5_net_create(int S, int max_batch, a5_net** out) {
  A5_ARG(out && S >= 5 && S <= A5_MAX_BOARD && max_batch > 0);
  a5_net* net = new a5_net();
  net->S = S; net->C = S * S; net->max_batch = max_batch;
  int rc = fp32_alloc(net);
  if (rc == A5_OK) rc = tc_alloc(net);
  if (rc == A5_OK) rc = small_alloc(net, &net->sm);
  if (rc != A5_OK) { a5_net_destroy(net); return rc; }
  *out = net;
  return A5_OK;
}

int a5_net_destroy(a5_net* net) {
  if (!net) return A5_OK;
  fp32_free(net);
  tc_free(net);
  small_free(net->sm);
  delete net;
  return A5_OK;
}

int a5_net_set_weights(a5_net* net, const float* const* d_tensors, void* stream) {
  A5_ARG(net && d_tensors);
  for (int i = 0; i < A5_NET_NUM_TENSORS; ++i) A5_ARG(d_tensors[i] != nullptr);
  int rc = fp32_set_weights(net, d_tensors, (cudaStream_t)stream);
  if (rc == A5_OK) rc = tc_set_weights(net, d_tensors, (cudaStream_t)stream);
  if (rc == A5_OK) rc = small_set_weights(net, net->sm, (cudaStream_t)stream);
  if (rc == A5_OK) net->has_weights = true;
  return rc;
}

int a5_net_forward(a5_net* net, const int8_t* d_planes, int n, float* d_prob, float* d_value, int mode, void* stream) {
  A5_ARG(net && d_planes && d_prob && d_value && n >= 0 && n <= net->max_batch);
  if (!net->has_weights) { set_error("a5_net_forward: no weights set"); return A5_ERR_STATE; }
  if (n == 0) return A5_OK;
  if (mode == A5_NET_FP32) return fp32_forward(net, d_planes, n, d_prob, d_value, (cudaStream_t)stream);
  if (mode == A5_NET_TC) return tc_forward(net, d_planes, n, d_prob, d_value, (cudaStream_t)stream);
  if (mode == A5_NET_SMALL) {
    if (n > A5_NET_SMALL_MAX) { set_error("a5_net_forward: A5_NET_SMALL takes at most %d boards, got %d", A5_NET_SMALL_MAX, n); return A5_ERR_ARG; }
    return small_forward(net, net->sm, d_planes, n, d_prob, d_value, (cudaStream_t)stream);
  }
  set_error("a5_net_forward: unknown mode %d", mode);
  return A5_ERR_ARG;
}

}  // extern "C"
