// Batched rule / encoding kernels behind the a5_rules_* entry points.
// One warp per board, 4 boards per CTA, board staged in shared memory.
#include <stdarg.h>
#include "rules.cuh"

namespace a5 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int WARPS = 4;
constexpr int CMAX = 256;

__device__ __forceinline__ void stage_board(const int8_t* g, int8_t* s, int C, int lane) {
  for (int c = lane; c < C; c += 32) s[c] = g[c];
  __syncwarp();
}

__global__ void __launch_bounds__(WARPS * 32) k_terminal(const int8_t* boards, int n, int S, int goal, int8_t* codes) {
  __shared__ int8_t sb[WARPS][CMAX];
  __shared__ uint32_t valid[WARPS][4 * (CMAX / 32)];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int i = blockIdx.x * WARPS + w;
  if (i >= n) return;
  constexpr int NCH = CMAX / 32;
  warp_valid_masks<NCH>(S, goal, lane, valid[w]);
  stage_board(boards + (size_t)i * S * S, sb[w], S * S, lane);
  uint32_t own[NCH], opp[NCH];
  warp_board_masks<NCH>(sb[w], S * S, lane, own, opp);
  int code = warp_terminal_bits<NCH>(own, opp, valid[w], S, goal, lane);
  if (lane == 0) codes[i] = (int8_t)code;
}

__global__ void __launch_bounds__(WARPS * 32) k_step(const int8_t* boards, const int32_t* cells, int n, int S, int8_t* out) {
  __shared__ int8_t sb[WARPS][CMAX];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int i = blockIdx.x * WARPS + w;
  if (i >= n) return;
  int C = S * S;
  stage_board(boards + (size_t)i * C, sb[w], C, lane);
  warp_step(sb[w], C, cells[i], lane);
  __syncwarp();
  for (int c = lane; c < C; c += 32) out[(size_t)i * C + c] = sb[w][c];
}

__global__ void __launch_bounds__(WARPS * 32) k_legal(const int8_t* boards, int n, int S, uint8_t* mask, int32_t* count) {
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int i = blockIdx.x * WARPS + w;
  if (i >= n) return;
  int C = S * S, cnt = 0;
  for (int base = 0; base < C; base += 32) {
    int c = base + lane;
    bool e = c < C && boards[(size_t)i * C + c] == 0;
    if (c < C) mask[(size_t)i * C + c] = e;
    cnt += __popc(__ballot_sync(FULL, e));
  }
  if (lane == 0) count[i] = cnt;
}

__global__ void __launch_bounds__(WARPS * 32) k_inputs(const int8_t* boards, const int32_t* last, int n, int S, int8_t* planes) {
  __shared__ int8_t sb[WARPS][CMAX];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int i = blockIdx.x * WARPS + w;
  if (i >= n) return;
  int C = S * S;
  stage_board(boards + (size_t)i * C, sb[w], C, lane);
  warp_write_planes(sb[w], C, last ? last[i] : -1, planes + (size_t)i * 3 * C, lane);
}

// utils.py:156-175.  Lane r encodes row r (<= S+1 chars); an exclusive scan of the row
// lengths gives every row its output offset.
__global__ void __launch_bounds__(WARPS * 32) k_encode(const int8_t* boards, int n, int S, char* states, int stride, int32_t* lens) {
  __shared__ int8_t sb[WARPS][CMAX];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int i = blockIdx.x * WARPS + w;
  if (i >= n) return;
  int C = S * S;
  stage_board(boards + (size_t)i * C, sb[w], C, lane);
  char row[A5_MAX_BOARD + 2];
  int len = 0;
  if (lane < S) {
    int run = 0;
    for (int j = 0; j < S; ++j) {
      int v = sb[w][lane * S + j];
      if (v == 0) { ++run; continue; }
      if (run) { row[len++] = (char)('a' + run); run = 0; }
      row[len++] = (char)('0' + v + 2);
    }
    if (run) row[len++] = (char)('a' + run);
    row[len++] = '/';
  }
  int off = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, off, o);
    if (lane >= o) off += t;
  }
  int total = __shfl_sync(FULL, off, 31);
  off -= len;
  char* dst = states + (size_t)i * stride;
  for (int t = 0; t < len; ++t) dst[off + t] = row[t];
  if (lane == 0) { dst[total] = 0; lens[i] = total; }
}

// utils.py:178-196.  A sequential parse (boundary format conversion, not a hot path).
__global__ void __launch_bounds__(WARPS * 32) k_decode(const char* states, int stride, int n, int S, int8_t* boards) {
  __shared__ int8_t sb[WARPS][CMAX];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int i = blockIdx.x * WARPS + w;
  if (i >= n) return;
  int C = S * S;
  for (int c = lane; c < C; c += 32) sb[w][c] = 0;
  __syncwarp();
  if (lane == 0) {
    const char* s = states + (size_t)i * stride;
    int r = 0, c = 0;
    for (int t = 0; t < stride && s[t]; ++t) {
      char ch = s[t];
      if (ch == '/') { ++r; c = 0; }
      else if (ch >= 'a' && ch <= 'z') c += ch - 'a';
      else { if (r < S && c < S) sb[w][r * S + c] = (int8_t)(ch - '0' - 2); ++c; }
    }
  }
  __syncwarp();
  for (int c = lane; c < C; c += 32) boards[(size_t)i * C + c] = sb[w][c];
}

static inline int nblocks(int n) { return (n + WARPS - 1) / WARPS; }

}  // namespace a5

using namespace a5;

extern "C" {

int a5_version(void) { return A5_VERSION; }
const char* a5_last_error(void) { return a5::g_err; }

#define RULES_PRE()                                            \
  A5_ARG(n >= 0 && S >= 5 && S <= A5_MAX_BOARD);               \
  if (n == 0) return A5_OK;                                    \
  cudaStream_t st = (cudaStream_t)stream

int a5_rules_terminal(const int8_t* d_boards, int n, int S, int goal, int8_t* d_codes, void* stream) {
  RULES_PRE();
  A5_ARG(d_boards && d_codes && goal >= 2 && goal <= S);
  k_terminal<<<nblocks(n), WARPS * 32, 0, st>>>(d_boards, n, S, goal, d_codes);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

int a5_rules_step(const int8_t* d_boards, const int32_t* d_cells, int n, int S, int8_t* d_out, void* stream) {
  RULES_PRE();
  A5_ARG(d_boards && d_cells && d_out);
  k_step<<<nblocks(n), WARPS * 32, 0, st>>>(d_boards, d_cells, n, S, d_out);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

int a5_rules_legal(const int8_t* d_boards, int n, int S, uint8_t* d_mask, int32_t* d_count, void* stream) {
  RULES_PRE();
  A5_ARG(d_boards && d_mask && d_count);
  k_legal<<<nblocks(n), WARPS * 32, 0, st>>>(d_boards, n, S, d_mask, d_count);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

int a5_rules_inputs(const int8_t* d_boards, const int32_t* d_last, int n, int S, int8_t* d_planes, void* stream) {
  RULES_PRE();
  A5_ARG(d_boards && d_planes);
  k_inputs<<<nblocks(n), WARPS * 32, 0, st>>>(d_boards, d_last, n, S, d_planes);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

int a5_rules_encode(const int8_t* d_boards, int n, int S, char* d_states, int stride, int32_t* d_len, void* stream) {
  RULES_PRE();
  A5_ARG(d_boards && d_states && d_len && stride >= S * (S + 1) + 1);
  k_encode<<<nblocks(n), WARPS * 32, 0, st>>>(d_boards, n, S, d_states, stride, d_len);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

int a5_rules_decode(const char* d_states, int stride, int n, int S, int8_t* d_boards, void* stream) {
  RULES_PRE();
  A5_ARG(d_states && d_boards && stride > 0);
  k_decode<<<nblocks(n), WARPS * 32, 0, st>>>(d_states, stride, n, S, d_boards);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

}  // extern "C"
