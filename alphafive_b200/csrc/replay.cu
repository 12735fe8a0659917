// Replay-buffer sampling on the device: RandomStack.get_data of the reference (utils.py:118-146).
//
// The buffer is a ring of the engine's fixed-stride ply records (a5_record_header + int8 board +
// f32 policy, alphafive.h) resident in HBM.  A training batch is a gather of `num` records, each
// transformed by one of the eight board symmetries -- np.rot90(k) then optionally np.flip(axis 0),
// with the last move remapped the same way (utils.py:129-140) -- and expanded to the three input
// planes of board_to_inputs (utils.py:256-272).  One warp per sample; every output cell looks its
// source cell up through the inverse symmetry, so both the record read (32 + C + 4C bytes) and the
// batch write (12C + 4C + 8 bytes) are coalesced.  HBM-bound byte work: ~2.6 KB per sample at 11x11.
#include "common.cuh"

namespace a5 {

__global__ void __launch_bounds__(128) k_replay_sample(const uint8_t* __restrict__ records, int stride, int rec_bb, int S,
                                                      const long long* __restrict__ idx, const uint8_t* __restrict__ rot,
                                                      const uint8_t* __restrict__ flip, int num, float* __restrict__ boards,
                                                      float* __restrict__ weights, float* __restrict__ values,
                                                      float* __restrict__ policies) {
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= num) return;
  const int C = S * S;
  const uint8_t* rec = records + (size_t)idx[n] * stride;
  const a5_record_header* hd = (const a5_record_header*)rec;
  const int8_t* board = (const int8_t*)(rec + sizeof(a5_record_header));
  const float* policy = (const float*)(rec + sizeof(a5_record_header) + rec_bb);
  const int k = rot[n] & 3;
  const bool fl = flip[n] != 0;
  const int last = hd->last_action;
  float* ob = boards + (size_t)n * 3 * C;
  float* op = policies + (size_t)n * C;
  for (int c = lane; c < C; c += 32) {
    // output cell (i2, j) <- source cell (si, sj): undo the flip, then the rotation
    const int i2 = c / S, j = c - i2 * S;
    const int i = fl ? S - 1 - i2 : i2;
    int si, sj;
    switch (k) {
      case 0: si = i; sj = j; break;
      case 1: si = j; sj = S - 1 - i; break;                  // rot90: out[i][j] = in[j][S-1-i]
      case 2: si = S - 1 - i; sj = S - 1 - j; break;
      default: si = S - 1 - j; sj = i; break;
    }
    const int src = si * S + sj;
    const int v = board[src];
    ob[c] = v == 1 ? 1.0f : 0.0f;
    ob[C + c] = v == -1 ? 1.0f : 0.0f;
    ob[2 * C + c] = src == last ? 1.0f : 0.0f;
    op[c] = policy[src];
  }
  if (lane == 0) {
    weights[n] = hd->weight;
    values[n] = hd->value;
  }
}

}  // namespace a5

extern "C" int a5_replay_sample(const void* d_records, int S, const int64_t* d_idx, const uint8_t* d_rot,
                                const uint8_t* d_flip, int num, float* d_boards, float* d_weights, float* d_values,
                                float* d_policies, void* stream) {
  A5_ARG(d_records && d_idx && d_rot && d_flip && d_boards && d_weights && d_values && d_policies);
  A5_ARG(S >= 1 && S <= A5_MAX_BOARD && num >= 0);
  if (num == 0) return A5_OK;
  const int C = S * S, rec_bb = (C + 15) / 16 * 16;
  a5::k_replay_sample<<<(num + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)d_records, a5_record_stride(S), rec_bb, S, (const long long*)d_idx, d_rot, d_flip, num, d_boards,
      d_weights, d_values, d_policies);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}
