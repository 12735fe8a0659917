// Policy/value network: graph description shared by the fp32 and tensor-core paths.
// Topology of genData/network.py:58-88 (TF1, channels_first, SAME padding, ELU).
#pragma once
#include "common.cuh"

namespace a5 {

// Canonical tensor order of a5_net_set_weights (TF variable names of the checkpoint).
enum Tensor {
  T_CONV1_K, T_CONV1_B,
  T_B1_RES_K, T_B1_RES_B, T_B1_C1_K, T_B1_C1_B, T_B1_C2_K, T_B1_C2_B,
  T_B2_RES_K, T_B2_RES_B, T_B2_C1_K, T_B2_C1_B, T_B2_C2_K, T_B2_C2_B,
  T_B3_RES_K, T_B3_RES_B, T_B3_C1_K, T_B3_C1_B, T_B3_C2_K, T_B3_C2_B,
  T_VCONV_K, T_VCONV_B, T_VFC1_K, T_VFC1_B, T_VFC2_K, T_VFC2_B,
  T_B4_RES_K, T_B4_RES_B, T_B4_C1_K, T_B4_C1_B, T_B4_C2_K, T_B4_C2_B,
  T_B5_RES_K, T_B5_RES_B, T_B5_C1_K, T_B5_C1_B, T_B5_C2_K, T_B5_C2_B,
  T_PCONV_K, T_PCONV_B, T_PFC_K, T_PFC_B,
  T_COUNT
};
static_assert(T_COUNT == A5_NET_NUM_TENSORS, "tensor table out of sync with alphafive.h");

// Residual blocks (network.py:52-56): name, Cin, Cout, first tensor index.
struct BlockSpec { int cin, cout, t0; };
static const BlockSpec kBlocks[5] = {
    {32, 64, T_B1_RES_K}, {64, 128, T_B2_RES_K}, {128, 32, T_B3_RES_K}, {128, 64, T_B4_RES_K}, {64, 32, T_B5_RES_K}};

// Padded position space.  Every board is laid out with row pitch S+1 (one zero column)
// and S+1 rows (one zero row); boards are stacked, so a 3x3 tap (dy, dx) is the constant
// row offset dy*(S+1)+dx and every out-of-board neighbour reads a zero cell (the zero row
// of the previous board, the zero column of the previous row, or the guard band).
struct PosSpace {
  int S, pitch, per_board, guard;
  __host__ __device__ PosSpace(int S_) : S(S_), pitch(S_ + 1), per_board((S_ + 1) * (S_ + 1)), guard(32) {}
  __host__ __device__ long long rows(int boards) const { return 2LL * guard + (long long)boards * per_board; }
};

__device__ __forceinline__ float elu(float x) { return x > 0.0f ? x : expm1f(x); }

}  // namespace a5
