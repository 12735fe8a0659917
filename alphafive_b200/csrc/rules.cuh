// Warp-cooperative Gomoku rule primitives on a board staged in shared memory.
// One warp owns one board: int8[S*S] (+1 side to move, -1 opponent, 0 empty).
// Lane l handles cells l, l+32, l+64, ... so that ballots enumerate cells in
// row-major order -- the reference's action order (utils.py:238-245).
#pragma once
#include "common.cuh"

namespace a5 {

// utils.py:199-235 on bitboards.  Returns 0 not over, 1 (True,+1.0), 2 (True,-1.0), 3 draw.
// A window can only sum to +-goal if all `goal` cells carry one colour, so a hit at anchor c in a
// direction of cell stride d is  B & B>>d & ... & B>>(goal-1)d  at bit c, masked by the anchors whose
// window stays on the board.  The reference scans anchors row-major and, per anchor, the windows
// down / right / down-right / up-right; all windows of one anchor contain the anchor cell, so the
// colour of the FIRST anchor with any hit decides -- the minimum over (anchor, colour) keys below.
// The up-right window anchored at (i, j) is the down-left window of its top end (i-goal+1, j+goal-1):
// it is found there with stride S-1 and credited to anchor + (goal-1)(S-1).
//
// valid[dir][NCH]: anchor masks of the four directions (warp_valid_masks, once per kernel and warp).
template <int NCH>
__device__ __forceinline__ void warp_valid_masks(int S, int goal, int lane, uint32_t* valid) {
  const int C = S * S;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = k * 32 + lane;
    const int i = c / S, j = c - i * S;
    const bool in = c < C, down = i + goal <= S, right = j + goal <= S, left = j - goal + 1 >= 0;
    const unsigned m0 = __ballot_sync(FULL, in && down);
    const unsigned m1 = __ballot_sync(FULL, in && right);
    const unsigned m2 = __ballot_sync(FULL, in && down && right);
    const unsigned m3 = __ballot_sync(FULL, in && down && left);
    if (lane == 0) { valid[k] = m0; valid[NCH + k] = m1; valid[2 * NCH + k] = m2; valid[3 * NCH + k] = m3; }
  }
  __syncwarp();
}

// own / opp: the position's bitboards (word k = cells 32k..32k+31), uniform in the warp.  Lanes 0..7
// each take one (colour, direction) pair.
template <int NCH>
__device__ __forceinline__ int warp_terminal_bits(const uint32_t (&own)[NCH], const uint32_t (&opp)[NCH],
                                                  const uint32_t* valid, int S, int goal, int lane) {
  uint32_t key = 0xffffffffu;
  if (lane < 8) {
    const int colour = lane >> 2, dir = lane & 3;
    const int delta = dir == 0 ? S : (dir == 1 ? 1 : (dir == 2 ? S + 1 : S - 1));
    uint32_t y[NCH], h[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      y[k] = colour ? opp[k] : own[k];
      h[k] = y[k] & valid[dir * NCH + k];
    }
    for (int t = 1; t < goal; ++t) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {                       // y >>= delta (delta <= S + 1 < 32), ascending k in place
        y[k] = __funnelshift_r(y[k], k + 1 < NCH ? y[k + 1] : 0u, delta);
        h[k] &= y[k];
      }
    }
    int pos = -1;
#pragma unroll
    for (int k = NCH - 1; k >= 0; --k)
      if (h[k]) pos = k * 32 + __ffs(h[k]) - 1;
    if (pos >= 0) key = ((uint32_t)(pos + (dir == 3 ? (goal - 1) * (S - 1) : 0)) << 1) | (uint32_t)colour;
  }
  key = __reduce_min_sync(FULL, key);
  if (key != 0xffffffffu) return (key & 1u) ? 2 : 1;
  const int C = S * S;
  bool empty = false;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int left = C - 32 * k;
    const uint32_t cells = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
    empty = empty || (~(own[k] | opp[k]) & cells) != 0u;
  }
  return empty ? 0 : 3;
}

// Board in shared memory -> bitboards, uniform in the warp.
template <int NCH>
__device__ __forceinline__ void warp_board_masks(const int8_t* b, int C, int lane, uint32_t (&own)[NCH], uint32_t (&opp)[NCH]) {
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = k * 32 + lane;
    const int v = c < C ? b[c] : 0;
    own[k] = __ballot_sync(FULL, v == 1);
    opp[k] = __ballot_sync(FULL, v == -1);
  }
}

// utils.py:275-283: place +1 at `cell`, negate.  Caller syncs the warp afterwards.
__device__ __forceinline__ void warp_step(int8_t* b, int C, int cell, int lane) {
  for (int c = lane; c < C; c += 32) {
    int v = (c == cell) ? 1 : b[c];
    b[c] = (int8_t)(-v);
  }
}

// utils.py:256-272: int8[3][C] planes written straight to global memory.
__device__ __forceinline__ void warp_write_planes(const int8_t* b, int C, int last, int8_t* out, int lane) {
  for (int c = lane; c < C; c += 32) {
    int v = b[c];
    out[c] = v == 1;
    out[C + c] = v == -1;
    out[2 * C + c] = c == last;
  }
}

}  // namespace a5
