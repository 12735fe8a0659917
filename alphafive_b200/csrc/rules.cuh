// Warp-cooperative Gomoku rule primitives on a board staged in shared memory.
// One warp owns one board: int8[S*S] (+1 side to move, -1 opponent, 0 empty).
// Lane l handles cells l, l+32, l+64, ... so that ballots enumerate cells in
// row-major order -- the reference's action order (utils.py:238-245).
#pragma once
#include "common.cuh"

namespace a5 {

// utils.py:199-235.  Returns 0 not over, 1 (True,+1.0), 2 (True,-1.0), 3 draw.
// A window can only sum to +-goal if all `goal` cells carry the anchor's colour, so
// each lane tests the four windows anchored at its cells; the first anchor in
// row-major order decides the colour, exactly the reference's scan order (two hits at
// one anchor share the anchor cell, hence the colour).
__device__ __forceinline__ int warp_terminal(const int8_t* b, int S, int goal, int lane) {
  const int C = S * S;
  unsigned any_empty = 0;
  for (int base = 0; base < C; base += 32) {
    int c = base + lane;
    int col = 0;
    bool hit = false;
    if (c < C) {
      col = b[c];
      int i = c / S, j = c - i * S;
      if (col != 0) {
        bool down = i + goal <= S, right = j + goal <= S, up = i - goal + 1 >= 0;
        bool h0 = down, h1 = right, h2 = down && right, h3 = up && right;
        for (int t = 1; t < goal; ++t) {
          if (h0) h0 = b[c + t * S] == col;
          if (h1) h1 = b[c + t] == col;
          if (h2) h2 = b[c + t * S + t] == col;
          if (h3) h3 = b[c - t * S + t] == col;
        }
        hit = h0 | h1 | h2 | h3;
      }
    }
    unsigned m = __ballot_sync(FULL, hit);
    any_empty |= __ballot_sync(FULL, c < C && col == 0);
    if (m) {
      int first = __ffs(m) - 1;
      int colour = __shfl_sync(FULL, col, first);
      return colour > 0 ? 1 : 2;
    }
  }
  return any_empty ? 0 : 3;
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
