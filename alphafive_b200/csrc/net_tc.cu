// Tensor-core (tcgen05) forward pass of the policy/value net -- placeholder until the
// implicit-GEMM kernels land; A5_NET_TC fails loudly instead of falling back.
#include "net.cuh"

namespace a5 {
int tc_alloc(a5_net*) { return A5_OK; }
void tc_free(a5_net*) {}
int tc_set_weights(a5_net*, const float* const*, cudaStream_t) { return A5_OK; }
int tc_forward(a5_net*, const int8_t*, int, float*, float*, cudaStream_t) {
  set_error("a5_net_forward: A5_NET_TC is not built into this library");
  return A5_ERR_STATE;
}
}  // namespace a5
