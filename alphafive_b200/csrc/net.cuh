// a5_net: device state of the policy/value network (both compute paths).
#pragma once
#include <cuda_fp16.h>
#include "net_common.cuh"

struct a5_tc_state;   // tensor-core path (net_tc.cu)
namespace a5 { struct SmallState; }   // small-batch latency path (net_small.cu)

struct a5_net {
  int S = 0, C = 0, max_batch = 0;
  bool has_weights = false;
  // ---- fp32 CUDA-core path (net_fp32.cu) -----------------------------------
  // packed weights: per layer a [Ktot][ldw] matrix (ldw = Cout rounded up to 64)
  float* w[13] = {};        // 0 conv1, 1..10 block convs (res folded into conv2), 11 pconv, 12 pfc
  float* bias[13] = {};
  float* vconv_w = nullptr; float* vconv_b = nullptr;   // 1x1 32->4
  float* vfc1_w = nullptr;  float* vfc1_b = nullptr;    // [4C][64]
  float* vfc2_w = nullptr;  float* vfc2_b = nullptr;    // [64]
  // activations in padded position space, fp32 [rows][channels]
  float* act[11] = {};      // a32, b1h, b1o, b2h, b2o, b3h, b3o, b4h, b4o, b5h, b5o
  float* pflat = nullptr;   // [B][16*C]
  float* logits = nullptr;  // [B][ldl]
  int ldl = 0;
  // ---- tensor-core path ------------------------------------------------------
  a5_tc_state* tc = nullptr;
  // ---- small-batch latency path (shares the fp32 path's matrices and activation buffers) ----
  a5::SmallState* sm = nullptr;
};

namespace a5 {
int fp32_alloc(a5_net* net);
void fp32_free(a5_net* net);
int fp32_set_weights(a5_net* net, const float* const* t, cudaStream_t st);
int fp32_forward(a5_net* net, const int8_t* planes, int n, float* prob, float* value, cudaStream_t st);
int fp32_heads(a5_net* net, int n, float* prob, float* value, cudaStream_t st);

struct HeadsState;
int heads_alloc(a5_net* net, HeadsState** out);
void heads_free(HeadsState* h);
int heads_set_weights(a5_net* net, HeadsState* h, const float* const* t, cudaStream_t st);
int heads_forward(a5_net* net, HeadsState* h, int n, float* prob, float* value, cudaStream_t st);
// conv-epilogue side of the heads: where the fused 1x1 head convs write, and their weights
struct HeadsIO { __half* a_pol; __half* a_val; const float* pconv_w; const float* pconv_b; int nst_pol, nst_val; };
HeadsIO heads_io(const HeadsState* h);

int small_alloc(a5_net* net, SmallState** out);
void small_free(SmallState* s);
int small_set_weights(a5_net* net, SmallState* s, cudaStream_t st);
int small_forward(a5_net* net, SmallState* s, const int8_t* planes, int n, float* prob, float* value, cudaStream_t st);

int tc_alloc(a5_net* net);
void tc_free(a5_net* net);
int tc_set_weights(a5_net* net, const float* const* t, cudaStream_t st);
int tc_forward(a5_net* net, const int8_t* planes, int n, float* prob, float* value, cudaStream_t st, int parts = 7);
}  // namespace a5
