// Cross-game evaluation cache for lock-step self-play (a5_evalcache_*).
//
// evaluate_and_expand asks pv_fn for every unseen position (player.py:186-202).  The reference's five workers each
// ask for themselves; with thousands of games in one batch many of them ask for the SAME position -- openings
// above all -- within a few hundred passes (measured at 4096 games, 500 sims: 5 % of the leaves of a pass were
// evaluated during the previous 512 passes, tools/leaf_dups.py).  The network is a deterministic, batch-invariant
// function of the three input planes, so serving such a leaf from a table of earlier results is exact.
//
// Per pass (all stream-ordered, fixed shapes, CUDA-graph capturable):
//   lookup   one warp per game with a pending leaf: planes -> 24-word bitboard key -> 4-slot probe with a FULL key
//            compare (a hash collision can only cost a miss).  Hit: prob / value are copied to the game's rows.
//            Miss: the planes are appended to a compact batch of at most `cap` boards (cap is chosen so that the
//            forward needs one group per CTA less than the full batch would); leaves beyond cap stay pending
//            (served = 0) and are looked up again next pass -- per-game results do not depend on the pass in which
//            a leaf is evaluated (counter-based RNG per game), only the lock-step timing shifts.
//   forward  the network on the compact batch (always `cap` boards: rows past the count hold stale planes)
//   commit   one warp per compact row: prob / value go to the game and into the table (one writer per slot:
//            the row that claimed it during lookup).
#include "common.cuh"

namespace a5 {

constexpr int EC_KEYW = 24;                 // 3 planes x 8 words (C <= 256)
constexpr int EC_PROBE = 4;
constexpr int EC_VSTRIDE = 260;             // floats per entry: value, prob[<= 256], pad

struct ECParams {
  int N, C, cap;
  uint32_t mask;                            // slots - 1
  uint32_t* keys;                           // [slots][24]
  float* vals;                              // [slots][EC_VSTRIDE]: value, prob[C]
  uint32_t* claim;                          // [slots] compact row that may write the slot this pass
  int8_t* cplanes; float* cprob; float* cvalue;   // compact batch [cap]
  int32_t* cgame; uint32_t* cslot; uint32_t* ckey; // per compact row: game, slot to fill (or ~0u), key
  int32_t* count;                           // [0] compact rows requested this pass (may exceed cap); [1] rotation of the game order
  unsigned long long* stats;                // lookups, hits, deferred, inserted
};

__device__ __forceinline__ uint32_t ec_mix(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  return h;
}

__device__ __forceinline__ void ec_lookup_body(const ECParams& P, const int8_t* __restrict__ planes, const uint8_t* __restrict__ need,
                                               float* __restrict__ prob, float* __restrict__ value, uint8_t* __restrict__ served,
                                               const int w, const int lane);
__device__ __forceinline__ void ec_commit_body(const ECParams& P, float* __restrict__ prob, float* __restrict__ value, const int row,
                                               const int lane);

__global__ void __launch_bounds__(128) k_ec_lookup(const ECParams P, const int8_t* __restrict__ planes, const uint8_t* __restrict__ need,
                                                   float* __restrict__ prob, float* __restrict__ value, uint8_t* __restrict__ served,
                                                   unsigned long long* kt) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();                                              // the tree pass that wrote planes / need is complete
  kt_begin(kt);
  ec_lookup_body(P, planes, need, prob, value, served, w, lane);
  if (kt) { __syncthreads(); kt_end(kt); }
}

__device__ __forceinline__ void ec_lookup_body(const ECParams& P, const int8_t* __restrict__ planes, const uint8_t* __restrict__ need,
                                               float* __restrict__ prob, float* __restrict__ value, uint8_t* __restrict__ served,
                                               const int w, const int lane) {
  if (w >= P.N) return;
  // rows of the compact batch go to whoever asks first, and late warps ask last: rotate which games those are from
  // pass to pass, or the same games would be the ones that wait whenever the batch is full
  const int g = (w + P.count[1]) % P.N;
  if (!need[g]) { if (lane == 0) served[g] = 1; return; }
  const int8_t* p = planes + (size_t)g * 3 * P.C;
  uint32_t mine = 0;
#pragma unroll
  for (int pl = 0; pl < 3; ++pl)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = k * 32 + lane;
      const unsigned m = __ballot_sync(FULL, c < P.C && p[pl * P.C + c] != 0);
      if (lane == pl * 8 + k) mine = m;
    }
  uint32_t h = lane < EC_KEYW ? ec_mix(mine + 0x9e3779b9u * (uint32_t)(lane + 1)) : 0u;
#pragma unroll
  for (int o = 16; o; o >>= 1) h ^= __shfl_xor_sync(FULL, h, o);
  h = ec_mix(h);
  int hit = -1, empty = -1;
  uint32_t kk[EC_PROBE];
#pragma unroll
  for (int j = 0; j < EC_PROBE; ++j)                        // the four probes are independent loads: all in flight at once
    kk[j] = lane < EC_KEYW ? P.keys[(size_t)((h + (uint32_t)j) & P.mask) * EC_KEYW + lane] : 0u;
#pragma unroll
  for (int j = 0; j < EC_PROBE; ++j) {
    const uint32_t s = (h + (uint32_t)j) & P.mask;
    const bool same = __all_sync(FULL, lane >= EC_KEYW || kk[j] == mine);
    const bool vacant = __all_sync(FULL, lane >= EC_KEYW || kk[j] == 0u);   // an all-zero key is never stored (see zero_key)
    if (same && hit < 0) hit = (int)s;
    if (vacant && empty < 0) empty = (int)s;
  }
  const bool zero_key = __all_sync(FULL, lane >= EC_KEYW || mine == 0u);   // the empty board: never cached
  if (lane == 0) atomicAdd(&P.stats[0], 1ull);
  if (hit >= 0 && !zero_key) {
    const float* v = P.vals + (size_t)hit * EC_VSTRIDE;
    for (int c = lane; c < P.C; c += 32) prob[(size_t)g * P.C + c] = v[1 + c];
    if (lane == 0) { value[g] = v[0]; served[g] = 1; atomicAdd(&P.stats[1], 1ull); }
    return;
  }
  int row = 0;
  if (lane == 0) row = atomicAdd(P.count, 1);
  row = __shfl_sync(FULL, row, 0);
  if (row >= P.cap) {                                      // the compact batch is full: stay pending, ask again next pass
    if (lane == 0) { served[g] = 0; atomicAdd(&P.stats[2], 1ull); }
    return;
  }
  int8_t* cp = P.cplanes + (size_t)row * 3 * P.C;
  for (int c = lane; c < 3 * P.C; c += 32) cp[c] = p[c];
  // where the result will be stored: a vacant slot of the probe window, else the first one (overwrite)
  const uint32_t slot = zero_key ? 0xffffffffu : (uint32_t)(empty >= 0 ? empty : (int)(h & P.mask));
  if (lane < EC_KEYW) P.ckey[(size_t)row * EC_KEYW + lane] = mine;
  if (lane == 0) {
    P.cgame[row] = g;
    P.cslot[row] = slot;
    if (slot != 0xffffffffu) P.claim[slot] = (uint32_t)row;   // several rows may want the slot: whoever is read back in commit wins
    served[g] = 1;
  }
}

__global__ void __launch_bounds__(128) k_ec_commit(const ECParams P, float* __restrict__ prob, float* __restrict__ value,
                                                   unsigned long long* kt) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();                                              // the heads kernel that wrote the compact prob / value is complete
  kt_begin(kt);
  ec_commit_body(P, prob, value, row, lane);
  if (kt) { __syncthreads(); kt_end(kt); }
}

__device__ __forceinline__ void ec_commit_body(const ECParams& P, float* __restrict__ prob, float* __restrict__ value, const int row,
                                               const int lane) {
  const int n = min(P.count[0], P.cap);
  if (row >= n) return;
  const int g = P.cgame[row];
  const uint32_t slot = P.cslot[row];
  const bool store = slot != 0xffffffffu && P.claim[slot] == (uint32_t)row;
  const float* cp = P.cprob + (size_t)row * P.C;
  float* v = P.vals + (size_t)(store ? slot : 0u) * EC_VSTRIDE;
  for (int c = lane; c < P.C; c += 32) {
    const float x = cp[c];
    prob[(size_t)g * P.C + c] = x;
    if (store) v[1 + c] = x;
  }
  if (lane == 0) {
    const float x = P.cvalue[row];
    value[g] = x;
    if (store) { v[0] = x; atomicAdd(&P.stats[3], 1ull); }
  }
  if (store && lane < EC_KEYW) P.keys[(size_t)slot * EC_KEYW + lane] = P.ckey[(size_t)row * EC_KEYW + lane];
}

__global__ void k_ec_reset_count(int32_t* count, int n) {
  pdl_launch_dependents();
  pdl_wait();
  count[0] = 0;
  count[1] = (count[1] + 1031) % n;
}

}  // namespace a5

using namespace a5;

struct a5_evalcache {
  ECParams p;
  size_t slots;
};

extern "C" {

int a5_evalcache_create(int S, int n_games, int log2_slots, int cap, a5_evalcache** out) {
  A5_ARG(out && S >= 5 && S <= A5_MAX_BOARD && n_games > 0 && log2_slots >= 10 && log2_slots <= 26 && cap > 0 && cap <= n_games);
  a5_evalcache* c = new a5_evalcache();
  memset(&c->p, 0, sizeof(c->p));
  c->slots = (size_t)1 << log2_slots;
  ECParams& p = c->p;
  p.N = n_games; p.C = S * S; p.cap = cap; p.mask = (uint32_t)(c->slots - 1);
  A5_CUDA(cudaMalloc(&p.keys, c->slots * EC_KEYW * 4));
  A5_CUDA(cudaMalloc(&p.vals, c->slots * EC_VSTRIDE * 4));
  A5_CUDA(cudaMalloc(&p.claim, c->slots * 4));
  A5_CUDA(cudaMalloc(&p.cplanes, (size_t)cap * 3 * p.C));
  A5_CUDA(cudaMalloc(&p.cprob, (size_t)cap * p.C * 4));
  A5_CUDA(cudaMalloc(&p.cvalue, (size_t)cap * 4));
  A5_CUDA(cudaMalloc(&p.cgame, (size_t)cap * 4));
  A5_CUDA(cudaMalloc(&p.cslot, (size_t)cap * 4));
  A5_CUDA(cudaMalloc(&p.ckey, (size_t)cap * EC_KEYW * 4));
  A5_CUDA(cudaMalloc(&p.count, 16));
  A5_CUDA(cudaMalloc(&p.stats, 4 * 8));
  A5_CUDA(cudaMemset(p.keys, 0, c->slots * EC_KEYW * 4));
  A5_CUDA(cudaMemset(p.claim, 0xff, c->slots * 4));
  A5_CUDA(cudaMemset(p.cplanes, 0, (size_t)cap * 3 * p.C));
  A5_CUDA(cudaMemset(p.count, 0, 16));
  A5_CUDA(cudaMemset(p.stats, 0, 4 * 8));
  *out = c;
  return A5_OK;
}

int a5_evalcache_destroy(a5_evalcache* c) {
  if (!c) return A5_OK;
  ECParams& p = c->p;
  cudaFree(p.keys); cudaFree(p.vals); cudaFree(p.claim); cudaFree(p.cplanes); cudaFree(p.cprob); cudaFree(p.cvalue);
  cudaFree(p.cgame); cudaFree(p.cslot); cudaFree(p.ckey); cudaFree(p.count); cudaFree(p.stats);
  delete c;
  return A5_OK;
}

int a5_evalcache_clear(a5_evalcache* c, void* stream) {
  A5_ARG(c);
  A5_CUDA(cudaMemsetAsync(c->p.keys, 0, c->slots * EC_KEYW * 4, (cudaStream_t)stream));
  return A5_OK;
}

int a5_evalcache_lookup(a5_evalcache* c, const int8_t* d_planes, const uint8_t* d_need, float* d_prob, float* d_value,
                        uint8_t* d_served, void* stream) {
  A5_ARG(c && d_planes && d_need && d_prob && d_value && d_served);
  cudaStream_t st = (cudaStream_t)stream;
  A5_CUDA(launch_pdl_k(k_ec_reset_count, 1u, 1u, 0, st, pdl_enabled(), c->p.count, c->p.N));
  A5_CUDA(launch_pdl_k(k_ec_lookup, (unsigned)((c->p.N + 3) / 4), 128u, 0, st, pdl_enabled(), c->p, d_planes, d_need, d_prob, d_value,
                       d_served, kt_slot(KT_EC_LOOKUP)));
  return A5_OK;
}

int8_t* a5_evalcache_planes(a5_evalcache* c) { return c ? c->p.cplanes : nullptr; }
float* a5_evalcache_prob(a5_evalcache* c) { return c ? c->p.cprob : nullptr; }
float* a5_evalcache_value(a5_evalcache* c) { return c ? c->p.cvalue : nullptr; }

int a5_evalcache_commit(a5_evalcache* c, float* d_prob, float* d_value, void* stream) {
  A5_ARG(c && d_prob && d_value);
  A5_CUDA(launch_pdl_k(k_ec_commit, (unsigned)((c->p.cap + 3) / 4), 128u, 0, (cudaStream_t)stream, pdl_enabled(), c->p, d_prob, d_value, kt_slot(KT_EC_COMMIT)));
  return A5_OK;
}

int a5_evalcache_stats(a5_evalcache* c, int64_t* h_out, void* stream) {
  A5_ARG(c && h_out);
  A5_CUDA(cudaMemcpyAsync(h_out, c->p.stats, 4 * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  A5_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return A5_OK;
}

}  // extern "C"
