// Lock-step MCTS engine: N concurrent Gomoku games, one warp per game.
//
// Replaces genData/player.py of the reference.  Per game the search graph is a
// transposition table (player.py:28-29): an open-addressing hash table in HBM that
// maps a position (two bitboards, side-to-move relative) to a node holding the edge
// statistics {n, w, p} per cell (player.py:9-20).  One pass of k_step does, per game:
//   1. expand the leaf found by the previous pass with the network output and back
//      its value up (player.py:186-202, 166-184);
//   2. when the budget of the current move is spent (auto_play): calc_policy, play,
//      record, terminal check, garbage-collect the table (player.py:53-126);
//   3. descend from the root -- terminal check, table lookup, PUCT selection with
//      fresh Dirichlet noise, step -- until an unseen position is reached, and write
//      its int8 input planes for the network (player.py:204-279, utils.py:256-272).
// Each game owns its warp, its table and its Philox stream: no inter-game traffic.
//
// HBM layout (per game g):
//   nodes[g][cap]  : {sum_n, hash, stones, pad, own[NCH], opp[NCH]} + n[E] + w[E] + p[E]
//                    (E = 32*NCH cells, SoA inside the node: a warp reads each array
//                    with coalesced 128-byte requests)
//   slots[g][H]    : u32 (tag16 << 16 | node+1), 0 = empty; H = 2*cap
//   root/leaf board: int8[KB] staged into shared memory by the owning warp
#include "rules.cuh"

namespace a5 {

constexpr int GS = 12;          // per-game statistics (see a5_engine_counters)
constexpr int WPB = 4;          // warps (games) per CTA

struct EP {
  int S, C, goal, N, sims, upper, training, random_a, auto_play, cap, H, max_inner, NCH, E;
  int node_bytes, KB, rec_stride, rec_cap, rec_bb;
  float c_puct, alpha, gamma;
  double init_temp, tau_decay, tau_decay_r;
  unsigned long long seed;
  long long gid_base;
  uint8_t* nodes; uint32_t* slots; uint16_t* free_stack;
  int32_t *free_top, *hiwater;
  int8_t *root_board, *leaf_board;
  int32_t *root_last, *leaf_last, *sims_left, *depth, *leaf_slot;
  uint32_t* leaf_hash;
  uint8_t *need_eval, *active;
  uint32_t* path;
  unsigned long long* rng_ctr;
  double* tau;
  int32_t *serial, *rec_len;
  uint8_t* rec_stage;           // [game][ply][rec_stride]: the plies of the game in progress, already in record layout
  uint8_t* out_rec; int32_t* out_count;
  long long* gstat;
  int8_t* planes;
  const uint8_t* served;        // a5_engine_step_served: 0 = this game's pending leaf was not evaluated this pass (or null)
  unsigned long long* kt;       // tooling: in-situ kernel timing slot (common.cuh), or null
};

// --------------------------------------------------------------------------- //
// per-warp context
// --------------------------------------------------------------------------- //
// tooling (a5__debug_step_times): per game, ns spent in the last k_step and what the warp did
// (bit 0 expanded a leaf, bit 1 played a move, bit 2 finished a game; bits 8.. = selection steps)
__device__ unsigned long long g_step_dbg[2 * 8192];
__device__ int g_step_dbg_on = 0;
__device__ __forceinline__ unsigned long long step_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// --------------------------------------------------------------------------- //
// Dirichlet(alpha) numerators (player.py:240), shared by the tree pass and the sampler export
// --------------------------------------------------------------------------- //
// Gamma(alpha < 1): Marsaglia-Tsang for alpha+1, boosted by U^(1/alpha).  One attempt, branch-free
// (acceptance ~96 %): every draw is addressed by (event counter, cell, attempt) in the Philox stream.
// Fast-math intrinsics: the variates only feed exploration noise (parity there is distributional).
__device__ __forceinline__ float gamma_try(const Philox& rng, unsigned long long ctr, float d, float c, float inv_alpha,
                                           int cell, uint32_t attempt, bool* ok) {
  const uint4 r = rng((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)cell, attempt);
  const float x = __fsqrt_rn(-2.0f * __logf(u01(r.x))) * __cosf(6.283185307179586f * u01(r.y));
  const float v1 = 1.0f + c * x;
  const float v = v1 * v1 * v1;
  const float u = u01(r.z);
  const float x2 = x * x;
  // squeeze u < 1 - 0.0331 x^4, else the full log test (NaN for v <= 0 compares false)
  const bool acc = u < 1.0f - 0.0331f * x2 * x2 || __logf(u) < 0.5f * x2 + d - d * v + d * __logf(v);
  *ok = v1 > 0.0f && acc;
  return d * v * exp2f(__log2f(u01(r.w)) * inv_alpha);
}

// Numerators for the NCH cells of this lane.  Attempt 0 of every cell runs as NCH independent
// instruction streams; the ~4 % rejected (lane, cell) pairs of the warp are then compacted -- the j-th
// reject is retried by lane j -- so the retries cost one more stream instead of NCH (a per-lane retry
// loop costs the whole warp a latency chain per round, and some lane nearly always rejects).
template <int NCH>
__device__ __forceinline__ void gamma_cells(const Philox& rng, unsigned long long ctr, int lane, float alpha,
                                            const bool (&legal)[NCH], float (&gam)[NCH]) {
  const float d = alpha + 1.0f - 1.0f / 3.0f;
  const float c = rsqrtf(9.0f * d);
  const float inv_alpha = 1.0f / alpha;
  unsigned rej[NCH];
  int total = 0;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    bool ok0;
    const float g0 = gamma_try(rng, ctr, d, c, inv_alpha, k * 32 + lane, 0u, &ok0);
    gam[k] = legal[k] ? g0 : 0.0f;
    rej[k] = __ballot_sync(FULL, legal[k] && !ok0);
    total += __popc(rej[k]);
  }
  for (int base = 0; base < total; base += 32) {           // rounds of 32 rejects (one round in practice)
    const int j = base + lane;                             // the reject this lane retries
    int cell = -1, cum = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int n = __popc(rej[k]);
      if (cell < 0 && j < cum + n) cell = k * 32 + (int)__fns(rej[k], 0, j - cum + 1);
      cum += n;
    }
    float val = d;
    if (cell >= 0) {
      bool ok = false;
      for (uint32_t attempt = 1; !ok && attempt < 24; ++attempt) {
        const float gv = gamma_try(rng, ctr, d, c, inv_alpha, cell, attempt, &ok);
        if (ok) val = gv;
      }
    }
    cum = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int src = cum + __popc(rej[k] & ((1u << lane) - 1u)) - base;   // who retried (lane, k)
      const float got = __shfl_sync(FULL, val, src & 31);
      if (((rej[k] >> lane) & 1u) && src >= 0 && src < 32) gam[k] = got;
      cum += __popc(rej[k]);
    }
  }
}

template <int NCH>
struct Warp {
  const EP& P;
  int g, lane;
  const uint32_t* valid = nullptr;   // shared: anchor masks of the terminal check (warp_valid_masks)
  int8_t* sb;       // shared: board of the position being processed (KB bytes, zero padded)
  float* sf;        // shared: 256 floats of scratch
  uint8_t* nodes;   // this game's node arena
  uint32_t* slots;
  uint16_t* fstack;
  int free_top, hiwater;
  unsigned long long ctr;
  Philox rng;
  long long st[GS];

  __device__ Warp(const EP& p, int g_, int lane_, int8_t* sb_, float* sf_)
      : P(p), g(g_), lane(lane_), sb(sb_), sf(sf_), rng(p.seed, (unsigned long long)(p.gid_base + g_)) {
    nodes = P.nodes + (size_t)g * P.cap * P.node_bytes;
    slots = P.slots + (size_t)g * P.H;
    fstack = P.free_stack + (size_t)g * P.cap;
    free_top = P.free_top[g];
    hiwater = P.hiwater[g];
    ctr = P.rng_ctr[g];
#pragma unroll
    for (int i = 0; i < GS; ++i) st[i] = 0;
  }
  __device__ void flush() {
    if (lane == 0) {
      P.free_top[g] = free_top;
      P.hiwater[g] = hiwater;
      P.rng_ctr[g] = ctr;
      long long* gs = P.gstat + (size_t)g * GS;
#pragma unroll
      for (int i = 0; i < GS; ++i)
        if (i != 7 && st[i]) gs[i] += st[i];
      int used = P.cap - free_top;
      if (used > gs[7]) gs[7] = used;
    }
  }

  __device__ __forceinline__ uint8_t* node(int idx) const { return nodes + (size_t)idx * P.node_bytes; }
  __device__ __forceinline__ static int32_t* hdr(uint8_t* nd) { return (int32_t*)nd; }
  __device__ __forceinline__ static uint32_t* masks(uint8_t* nd) { return (uint32_t*)(nd + 16); }
  __device__ __forceinline__ int32_t* edge_n(uint8_t* nd) const { return (int32_t*)(nd + 16 + 8 * NCH); }
  __device__ __forceinline__ float* edge_w(uint8_t* nd) const { return (float*)(nd + 16 + 8 * NCH) + P.E; }
  __device__ __forceinline__ float* edge_p(uint8_t* nd) const { return (float*)(nd + 16 + 8 * NCH) + 2 * P.E; }

  // board in shared memory -> bitboards (word k = cells 32k..32k+31), uniform in the warp
  __device__ __forceinline__ void board_masks(uint32_t (&own)[NCH], uint32_t (&opp)[NCH]) const {
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      int c = k * 32 + lane;
      int v = c < P.C ? sb[c] : 0;
      own[k] = __ballot_sync(FULL, v == 1);
      opp[k] = __ballot_sync(FULL, v == -1);
    }
  }
  __device__ __forceinline__ static uint32_t hash_masks(const uint32_t (&own)[NCH], const uint32_t (&opp)[NCH]) {
    uint32_t h = 0x811C9DC5u;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      h = (h ^ own[k]) * 0x01000193u; h ^= h >> 15; h *= 0x2C1B3C6Du;
      h = (h ^ opp[k]) * 0x01000193u; h ^= h >> 13; h *= 0x297A2D39u;
    }
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
  }
  __device__ __forceinline__ static int stones_of(const uint32_t (&own)[NCH], const uint32_t (&opp)[NCH]) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) s += __popc(own[k]) + __popc(opp[k]);
    return s;
  }

  // Linear probing, 32 slots per step.  Returns the node index or -1; *empty_pos receives
  // the first free slot of the probe sequence (where an insert would go).
  __device__ int find(uint32_t h, const uint32_t (&own)[NCH], const uint32_t (&opp)[NCH], int* empty_pos) const {
    const uint32_t tag = h >> 16;
    const int mask = P.H - 1;
    const int pos = (int)(h & (uint32_t)mask);
    for (int probe = 0; probe < P.H; probe += 32) {
      int sp = (pos + probe + lane) & mask;
      uint32_t s = slots[sp];
      unsigned em = __ballot_sync(FULL, s == 0);
      unsigned mt = __ballot_sync(FULL, s != 0 && (s >> 16) == tag);
      if (em) mt &= (1u << (__ffs(em) - 1)) - 1u;
      while (mt) {
        int l = __ffs(mt) - 1;
        mt &= mt - 1;
        int idx = (int)(__shfl_sync(FULL, s, l) & 0xffffu) - 1;
        const uint32_t* mk = masks(node(idx));
        bool eq = true;
        if (lane < 2 * NCH) {
          uint32_t want = 0;                      // own/opp word for this lane (no dynamic register indexing)
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            if (lane == k) want = own[k];
            if (lane == NCH + k) want = opp[k];
          }
          eq = mk[lane] == want;
        }
        if (__all_sync(FULL, eq)) return idx;
      }
      if (em) {
        *empty_pos = (pos + probe + __ffs(em) - 1) & mask;
        return -1;
      }
    }
    *empty_pos = -1;
    return -1;
  }

  __device__ void clear_table() {
    uint4* s4 = (uint4*)slots;
    for (int i = lane; i < P.H / 4; i += 32) s4[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < P.cap; i += 32) fstack[i] = (uint16_t)(P.cap - 1 - i);   // pop order 0,1,2,...
    for (int i = lane; i < hiwater; i += 32) hdr(node(i))[2] = -1;
    free_top = P.cap;
    hiwater = 0;
    __syncwarp();
  }

  // Keep only the nodes still reachable from the new root `sb` (positions that contain
  // all of its stones with matching colours); rebuild the hash table.  Semantically
  // invisible: dropped positions can never be reached again.
  __device__ void collect(const uint32_t (&rown)[NCH], const uint32_t (&ropp)[NCH]) {
    const int q0 = stones_of(rown, ropp);
    uint4* s4 = (uint4*)slots;
    for (int i = lane; i < P.H / 4; i += 32) s4[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    const int mask = P.H - 1;
    int new_hi = 0;
    for (int base = 0; base < hiwater; base += 32) {
      int idx = base + lane;
      bool used = false, keep = false;
      uint32_t h = 0;
      if (idx < hiwater) {
        uint8_t* nd = node(idx);
        const int32_t* hd = hdr(nd);
        int q = hd[2];
        used = q >= 0;
        if (used && q >= q0) {
          h = (uint32_t)hd[1];
          const uint32_t* mk = masks(nd);
          bool flip = (q - q0) & 1;
          keep = true;
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            uint32_t a = mk[flip ? NCH + k : k], b = mk[flip ? k : NCH + k];
            keep = keep && ((a & rown[k]) == rown[k]) && ((b & ropp[k]) == ropp[k]);
          }
        }
        if (used && !keep) hdr(nd)[2] = -1;
      }
      unsigned dead = __ballot_sync(FULL, used && !keep);
      if (used && !keep) fstack[free_top + __popc(dead & ((1u << lane) - 1u))] = (uint16_t)idx;
      free_top += __popc(dead);
      if (keep) {
        uint32_t val = ((h >> 16) << 16) | (uint32_t)(idx + 1);
        int pos = (int)(h & (uint32_t)mask);
        while (atomicCAS(&slots[pos], 0u, val) != 0u) pos = (pos + 1) & mask;
      }
      unsigned kp = __ballot_sync(FULL, keep);
      if (kp) new_hi = base + 32 - __clz(kp);
    }
    hiwater = new_hi;
    __syncwarp();
  }

  // player.py:166-184: walk the path upwards, v = -v; n += 1; w += v (edges are distinct).
  __device__ void backup(int depth, float v) {
    const uint32_t* path = P.path + (size_t)g * P.C;
    for (int i = lane; i < depth; i += 32) {
      uint32_t e = path[i];
      uint8_t* nd = node((int)(e >> 8));
      int cell = (int)(e & 0xffu);
      float val = ((depth - i) & 1) ? -v : v;
      edge_n(nd)[cell] += 1;
      edge_w(nd)[cell] += val;
    }
    __syncwarp();
  }

  __device__ __forceinline__ uint32_t rand_below(uint32_t purpose, uint32_t n) {
    uint4 r = rng((uint32_t)ctr, (uint32_t)(ctr >> 32), 0xFFFFFFFFu, purpose);
    return (uint32_t)(((unsigned long long)r.x * n) >> 32);
  }

  // k-th set predicate in cell order (pred[k] bit of this lane), uniform result
  __device__ int pick_kth(const bool (&pred)[NCH], int kth) const {
    int res = -1;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      unsigned m = __ballot_sync(FULL, pred[k]);
      int cnt = __popc(m);
      if (res < 0) {
        if (kth < cnt) {
          // position of the kth set bit
          unsigned mm = m;
          for (int t = 0; t < kth; ++t) mm &= mm - 1;
          res = k * 32 + __ffs(mm) - 1;
        } else {
          kth -= cnt;
        }
      }
    }
    return res;
  }

  // player.py:230-279.  `nd` is the node of the position in sb.  Returns the chosen cell.
  __device__ int select(uint8_t* nd, bool is_root) {
    int32_t* hd = hdr(nd);
    int sum_n = hd[0] + 1;                       // incremented before use (player.py:237)
    __syncwarp();
    if (lane == 0) hd[0] = sum_n;
    ++ctr;                                       // one RNG event per node visit
    const int32_t* en = edge_n(nd);
    const float* ew = edge_w(nd);
    const float* ep = edge_p(nd);
    int n[NCH];
    float wk[NCH], pk[NCH];                      // edge statistics: loaded up front so the HBM latency hides
    float score[NCH];                            // behind the Dirichlet draw below
    bool legal[NCH];
    float gam[NCH];
    float gsum = 0.0f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      int c = k * 32 + lane;
      legal[k] = c < P.C && sb[c] == 0;
      n[k] = legal[k] ? en[c] : 0;
      wk[k] = legal[k] ? ew[c] : 0.0f;
      pk[k] = legal[k] ? ep[c] : 0.0f;
      gam[k] = 0.0f;
    }
    if (P.training) {
      gamma_cells<NCH>(rng, ctr, lane, P.alpha, legal, gam);
#pragma unroll
      for (int k = 0; k < NCH; ++k) gsum += gam[k];
      gsum = warp_sum(gsum);
    }
    const double sq = sqrt((double)(sum_n + 1));
    float best = -INFINITY;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      int c = k * 32 + lane;
      score[k] = -INFINITY;
      if (legal[k]) {
        float p = pk[k];
        float q = n[k] > 0 ? __fdiv_rn(wk[k], (float)n[k]) : 0.0f;
        double t;
        if (P.training) {
          double eta = gsum > 0.0f ? (double)gam[k] / (double)gsum : 0.0;
          double pm = is_root ? (double)__fmul_rn(0.75f, p) + 0.25 * eta
                              : (double)__fmul_rn(0.9f, p) + 0.1 * eta;
          t = (double)P.c_puct * pm;
        } else {
          t = (double)__fmul_rn(P.c_puct, p);
        }
        double u = __ddiv_rn(__dmul_rn(t, sq), (double)(1 + n[k]));
        score[k] = __double2float_rn(__dadd_rn((double)q, u));
        best = fmaxf(best, score[k]);
      }
    }
    if (is_root && P.training) {                 // forced-visit ladder (player.py:264-276)
      for (int want = 0; want < 2; ++want) {
        bool pred[NCH];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < NCH; ++k) { pred[k] = legal[k] && n[k] == want; cnt += pred[k]; }
        cnt = warp_sum_i(cnt);
        if (cnt > 0) return pick_kth(pred, (int)rand_below(1u + want, (uint32_t)cnt));
      }
    }
    best = warp_max(best);
    bool pred[NCH];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) { pred[k] = legal[k] && score[k] == best; cnt += pred[k]; }
    cnt = warp_sum_i(cnt);
    int kth = cnt > 1 ? (int)rand_below(3u, (uint32_t)cnt) : 0;
    return pick_kth(pred, kth);
  }

  // player.py:186-202 with the network output of the previous pass; the leaf board is
  // already staged in sb.
  __device__ void expand(const float* prob_row, uint32_t h, int slot_pos,
                         const uint32_t (&own)[NCH], const uint32_t (&opp)[NCH]) {
    // priors of the legal cells, +0.0f elsewhere (x + 0.0f == x exactly, so the masked sequential sum below
    // equals the reference's sum over the legal cells only); KB is a multiple of 16, the pad stays zero
    int nlegal = 0;
    for (int c = lane; c < P.KB; c += 32) {
      const bool lg = c < P.C && sb[c] == 0;
      sf[c] = lg ? prob_row[c] : 0.0f;
      nlegal += lg;
    }
    nlegal = warp_sum_i(nlegal);
    __syncwarp();
    float tot = 0.0f;
    if (lane == 0) {
      const float4* s4 = (const float4*)sf;
      for (int c4 = 0; c4 < P.KB / 4; ++c4) {                   // sequential f32 sum, row-major (player.py:198)
        const float4 v = s4[c4];
        tot = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(tot, v.x), v.y), v.z), v.w);
      }
    }
    tot = __shfl_sync(FULL, tot, 0);
    const float denom = 1e-5f > tot ? 1e-5f : tot;
    st[2] += 1;
    st[5] += nlegal;
    if (free_top <= 0 || slot_pos < 0) { st[8] += 1; return; }        // arena full: reported, not stored
    int idx = fstack[free_top - 1];
    --free_top;
    if (idx + 1 > hiwater) hiwater = idx + 1;
    uint8_t* nd = node(idx);
    if (lane == 0) {
      int32_t* hd = hdr(nd);
      hd[0] = 0; hd[1] = (int32_t)h; hd[2] = stones_of(own, opp); hd[3] = 0;
      slots[slot_pos] = ((h >> 16) << 16) | (uint32_t)(idx + 1);
    }
    if (lane < 2 * NCH) {
      uint32_t v = 0;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        if (lane == k) v = own[k];
        if (lane == NCH + k) v = opp[k];
      }
      masks(nd)[lane] = v;
    }
    int32_t* en = edge_n(nd);
    float* ew = edge_w(nd);
    float* ep = edge_p(nd);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      int c = k * 32 + lane;
      bool lg = c < P.C && sb[c] == 0;
      en[c] = 0;
      ew[c] = 0.0f;
      ep[c] = lg ? __fdiv_rn(sf[c], denom) : 0.0f;
    }
    __syncwarp();
  }

  // numpy's pairwise float32 sum for n <= 128 (8 accumulators), recursive above.
  __device__ static float np_sum(const float* a, int n) {
    if (n < 8) {
      float r = 0.0f;
      for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
      return r;
    }
    if (n <= 128) {
      float r[8];
      for (int j = 0; j < 8; ++j) r[j] = a[j];
      int i;
      for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
      float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                            __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
      for (; i < n; ++i) res = __fadd_rn(res, a[i]);
      return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_sum(a, n2), np_sum(a + n2, n - n2));
  }

  // player.py:84-126 on the root node `nd` (may be null if the root was never expanded).
  // Writes the policy row (zeros when the reference returns None) and returns the action.
  __device__ int move_policy(uint8_t* nd, float* policy_out) {
    ++ctr;
    int n[NCH];
    bool legal[NCH];
    int maxn = -1, nleg = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      int c = k * 32 + lane;
      legal[k] = c < P.C && sb[c] == 0;
      n[k] = (legal[k] && nd) ? edge_n(nd)[c] : 0;
      if (legal[k]) { maxn = max(maxn, n[k]); ++nleg; }
    }
    maxn = warp_max_i(maxn);
    nleg = warp_sum_i(nleg);
    bool top[NCH];
    int ntop = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) { top[k] = legal[k] && n[k] == maxn; ntop += top[k]; }
    ntop = warp_sum_i(ntop);
    int best = pick_kth(top, ntop > 1 ? (int)rand_below(4u, (uint32_t)ntop) : 0);
    const bool soft = P.training || P.random_a;
    if (!soft) {
      if (policy_out)
        for (int c = lane; c < P.C; c += 32) policy_out[c] = 0.0f;
      return best;
    }
    const double tau0 = P.tau[g];
    __syncwarp();
    const double tau = tau0 * (P.random_a ? P.tau_decay_r : P.tau_decay);   // decays before use (player.py:108-111)
    if (lane == 0) P.tau[g] = tau;
    if (tau <= 0.01 || maxn <= 0) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        int c = k * 32 + lane;
        if (c < P.C && policy_out) policy_out[c] = top[k] ? __fdiv_rn(1.0f, (float)ntop) : 0.0f;
      }
      return best;
    }
    const float inv_tau = (float)(1.0 / tau);
    float pv[NCH];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      pv[k] = legal[k] ? powf(__fdiv_rn((float)n[k], (float)maxn), inv_tau) : 0.0f;
      s += pv[k];
    }
    s = warp_sum(s);
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      int c = k * 32 + lane;
      pv[k] = legal[k] ? __fdiv_rn(pv[k], s) : 0.0f;
      if (c < P.C && policy_out) policy_out[c] = pv[k];
      tot += (double)pv[k];
    }
    tot = warp_sum_d(tot);
    // np.random.choice(A, p): first index whose normalised cdf exceeds u
    uint4 r = rng((uint32_t)ctr, (uint32_t)(ctr >> 32), 0xFFFFFFFFu, 5u);
    const double target = u01d(r.x, r.y) * tot;
    double carry = 0.0;
    int action = -1, last_legal = -1;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      double x = (double)pv[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += t;
      }
      x += carry;
      unsigned m = __ballot_sync(FULL, legal[k] && x > target);
      unsigned lg = __ballot_sync(FULL, legal[k]);
      if (lg) last_legal = k * 32 + 31 - __clz(lg);
      if (action < 0 && m) action = k * 32 + __ffs(m) - 1;
      carry = __shfl_sync(FULL, x, 31);
    }
    return action >= 0 ? action : last_legal;
  }

  // Look the position in sb up and compute Player.get_action's budget (player.py:140-143).
  __device__ int budget_for_root(const uint32_t (&own)[NCH], const uint32_t (&opp)[NCH]) {
    int ep;
    int idx = find(hash_masks(own, opp), own, opp, &ep);
    if (idx < 0) return P.sims;
    return min(P.sims, P.upper - hdr(node(idx))[0]);
  }

  // player.py:73-82 + main.py:86-93: label the finished game and append its plies to the
  // harvest arena.  `code` is the terminal code of the position after the last move.
  __device__ void emit_game(int L, int code) {
    int base = 0;
    if (lane == 0) {
      // reserve L contiguous records; the counter only ever moves forward by committed games (a reservation
      // that does not fit is never added, so no rollback can hand two warps overlapping ranges)
      int old = *(volatile int32_t*)&P.out_count[0];
      while (true) {
        if (old + L > P.rec_cap) { atomicAdd(&P.out_count[2], L); base = -1; break; }
        const int prev = atomicCAS(&P.out_count[0], old, old + L);
        if (prev == old) { base = old; atomicAdd(&P.out_count[1], 1); break; }
        old = prev;
      }
    }
    base = __shfl_sync(FULL, base, 0);
    st[6] += 1;
    if (base < 0) { st[10] += L; return; }
    float value = code == 3 ? 0.0f : (code == 2 ? -1.0f : 1.0f);
    if (L & 1) value = -value;
    // utils.py:286-296 in float32: w[L-1] = 1, w[i] = w[i+1]*gamma, L*w/sum(w)
    if (lane == 0) {
      sf[L - 1] = 1.0f;
      for (int i = L - 2; i >= 0; --i) sf[i] = __fmul_rn(sf[i + 1], P.gamma);
    }
    __syncwarp();
    float wsum = 0.0f;
    if (lane == 0) wsum = np_sum(sf, L);
    wsum = __shfl_sync(FULL, wsum, 0);
    const int result = (value == 0.0f && code == 3) ? 0 : ((L & 1) ? 1 : -1);
    // the plies were staged in record layout as they were played (board, policy, last_action): complete the
    // headers, lane = ply, then move the L records as one linear 16-byte-vector copy with four loads in flight per
    // lane (the per-ply, per-field loop this replaces was the slowest warp of every pass in which a game ended)
    uint8_t* stage = P.rec_stage + (size_t)g * P.C * P.rec_stride;
    const int serial = P.serial[g];
    for (int t = lane; t < L; t += 32) {
      a5_record_header* hd = (a5_record_header*)(stage + (size_t)t * P.rec_stride);
      hd->game_id = P.gid_base + g;
      hd->game_serial = serial;
      hd->ply = (int16_t)t;
      hd->game_len = (int16_t)L;
      hd->value = (t & 1) ? -value : value;
      hd->weight = __fdiv_rn(__fmul_rn((float)L, sf[t]), wsum);
      hd->result = result;
    }
    __syncwarp();
    const uint4* src = (const uint4*)stage;
    uint4* dst = (uint4*)(P.out_rec + (size_t)base * P.rec_stride);
    const int n = L * (P.rec_stride / 16);
    for (int i0 = 0; i0 < n; i0 += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        if (i < n) v[u] = src[i];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 32 + lane;
        if (i < n) dst[i] = v[u];
      }
    }
    __syncwarp();
  }
};

// --------------------------------------------------------------------------- //
// kernels
// --------------------------------------------------------------------------- //
// Occupancy: the pass is a chain of dependent table reads per game, so resident warps hide it -- 7 CTAs
// (28 games) per SM put all 4096 games of the benchmark batch on the chip in one wave (measured:
// 236 us at 3 CTAs/SM, 165 at 4, 125 at 7, spills stay ~100 bytes).
#ifndef A5_STEP_MINB
#define A5_STEP_MINB (NCH <= 4 ? 7 : 4)
#endif
template <int NCH>
__device__ __forceinline__ void step_body(const EP& P, const float* __restrict__ prob, const float* __restrict__ value,
                                          int8_t (*s_board)[256], float (*s_f)[256], uint32_t (*s_valid)[4 * NCH]) {
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * WPB + wid;
  if (g >= P.N) return;
  if (!P.active[g]) {
    if (lane == 0) P.need_eval[g] = 0;
    return;
  }
  if (P.need_eval[g] && prob && P.served && !P.served[g]) {
    // the evaluation cache's compact batch was full (a5_evalcache_lookup): the leaf stays pending, nothing else
    // happens to this game in this pass
    if (g == 0 && lane == 0) P.gstat[9] += 1;              // (passes are counted through game 0)
    return;
  }
  const unsigned long long dbg_t0 = g_step_dbg_on ? step_now() : 0ull;
  unsigned dbg_flags = 0;
  Warp<NCH> W(P, g, lane, s_board[wid], s_f[wid]);
  warp_valid_masks<NCH>(P.S, P.goal, lane, s_valid[wid]);
  W.valid = s_valid[wid];
  int8_t* sb = W.sb;
  int sims_left = P.sims_left[g];
  uint32_t own[NCH], opp[NCH];
  W.st[9] = (g == 0);

  // ---- 1. expand + backup the pending leaf ---------------------------------
  if (P.need_eval[g] && prob) {
    const int8_t* lb = P.leaf_board + (size_t)g * P.KB;
    for (int c = lane; c < P.KB; c += 32) sb[c] = lb[c];
    __syncwarp();
    W.board_masks(own, opp);
    W.expand(prob + (size_t)g * P.C, P.leaf_hash[g], P.leaf_slot[g], own, opp);
    W.backup(P.depth[g], value[g]);
    --sims_left;
    W.st[1] += 1;
    dbg_flags |= 1u;
  }

  int inner = 0;
  bool pending = false;
  while (true) {
    // ---- 2. budget spent: play the move (auto_play) or go idle ---------------
    if (sims_left <= 0) {
      if (!P.auto_play) break;
      const int8_t* rb = P.root_board + (size_t)g * P.KB;
      for (int c = lane; c < P.KB; c += 32) sb[c] = rb[c];
      __syncwarp();
      W.board_masks(own, opp);
      int ep;
      int ridx = W.find(W.hash_masks(own, opp), own, opp, &ep);
      int L = P.rec_len[g];
      uint8_t* rs = P.rec_stage + ((size_t)g * P.C + L) * P.rec_stride;       // this ply's record (player.py:65)
      float* pol = (float*)(rs + sizeof(a5_record_header) + P.rec_bb);
      int action = W.move_policy(ridx >= 0 ? W.node(ridx) : nullptr, pol);
      int8_t* rbd = (int8_t*)(rs + sizeof(a5_record_header));
      for (int c = lane; c < P.KB; c += 32) rbd[c] = sb[c];
      if (lane == 0) ((a5_record_header*)rs)->last_action = P.root_last[g];
      ++L;
      __syncwarp();
      warp_step(sb, P.C, action, lane);
      __syncwarp();
      W.board_masks(own, opp);
      int code = warp_terminal_bits<NCH>(own, opp, W.valid, P.S, P.goal, lane);
      W.st[0] += 1;
      dbg_flags |= code ? 6u : 2u;
      if (code) {                                   // game over: emit, restart (player.py:73)
        W.emit_game(L, code);
        for (int c = lane; c < P.KB; c += 32) sb[c] = 0;
        __syncwarp();
        W.clear_table();
        if (lane == 0) {
          P.tau[g] = P.init_temp;
          P.serial[g] += 1;
          P.root_last[g] = -1;
          P.rec_len[g] = 0;
        }
        sims_left = P.sims;
      } else {
        W.collect(own, opp);
        sims_left = W.budget_for_root(own, opp);
        if (lane == 0) { P.root_last[g] = action; P.rec_len[g] = L; }
      }
      int8_t* rbw = P.root_board + (size_t)g * P.KB;
      for (int c = lane; c < P.KB; c += 32) rbw[c] = sb[c];
      __syncwarp();
      if (code) break;      // the warp that closed a game is the slowest of its pass: the new game's first descent waits a pass
      continue;
    }
    if (inner >= P.max_inner) break;                // yield: no leaf this pass
    // ---- 3. one simulation: descend from the root ----------------------------
    {
      const int8_t* rb = P.root_board + (size_t)g * P.KB;
      for (int c = lane; c < P.KB; c += 32) sb[c] = rb[c];
      __syncwarp();
    }
    int last = P.root_last[g];
    int depth = 0;
    uint32_t* path = P.path + (size_t)g * P.C;
    while (true) {
      W.board_masks(own, opp);
      int code = warp_terminal_bits<NCH>(own, opp, W.valid, P.S, P.goal, lane);   // before lookup (player.py:214)
      if (code) {
        float v = code == 1 ? 1.0f : (code == 2 ? -1.0f : 0.0f);
        W.backup(depth, v);
        --sims_left;
        W.st[1] += 1;
        W.st[3] += 1;
        ++inner;
        break;
      }
      uint32_t h = W.hash_masks(own, opp);
      int ep;
      int idx = W.find(h, own, opp, &ep);
      if (idx < 0) {                                            // unseen: ask the network
        int8_t* lb = P.leaf_board + (size_t)g * P.KB;
        for (int c = lane; c < P.KB; c += 32) lb[c] = sb[c];
        warp_write_planes(sb, P.C, last, P.planes + (size_t)g * 3 * P.C, lane);
        if (lane == 0) {
          P.leaf_hash[g] = h;
          P.leaf_slot[g] = ep;
          P.leaf_last[g] = last;
          P.depth[g] = depth;
        }
        pending = true;
        break;
      }
      int cell = W.select(W.node(idx), depth == 0);
      W.st[4] += 1;
      dbg_flags += 256u;
      if (lane == 0) path[depth] = ((uint32_t)idx << 8) | (uint32_t)cell;
      ++depth;
      __syncwarp();
      warp_step(sb, P.C, cell, lane);
      __syncwarp();
      last = cell;
    }
    if (pending) break;
  }
  if (lane == 0) {
    P.need_eval[g] = pending;
    P.sims_left[g] = sims_left;
    if (g_step_dbg_on && g < 8192) { g_step_dbg[g] = step_now() - dbg_t0; g_step_dbg[8192 + g] = dbg_flags; }
  }
  W.flush();
}

template <int NCH>
__global__ void __launch_bounds__(WPB * 32, A5_STEP_MINB) k_step(const __grid_constant__ EP P, const float* __restrict__ prob,
                                                  const float* __restrict__ value) {
  __shared__ __align__(16) int8_t s_board[WPB][256];
  __shared__ __align__(16) float s_f[WPB][256];
  __shared__ uint32_t s_valid[WPB][4 * NCH];
  pdl_launch_dependents();
  pdl_wait();                                              // the heads kernel that wrote prob / value is complete
  kt_begin(P.kt);
  step_body<NCH>(P, prob, value, s_board, s_f, s_valid);
  if (P.kt) { __syncthreads(); kt_end(P.kt); }
}

template <int NCH>
__global__ void __launch_bounds__(WPB * 32) k_reset(const __grid_constant__ EP P) {
  __shared__ __align__(16) int8_t s_board[WPB][256];
  __shared__ __align__(16) float s_f[WPB][256];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * WPB + wid;
  if (g >= P.N) return;
  Warp<NCH> W(P, g, lane, s_board[wid], s_f[wid]);
  W.hiwater = P.cap;                               // first reset: mark every node free
  W.clear_table();
  for (int c = lane; c < P.KB; c += 32) P.root_board[(size_t)g * P.KB + c] = 0;
  if (lane == 0) {
    P.root_last[g] = -1; P.leaf_last[g] = -1;
    P.sims_left[g] = P.auto_play ? P.sims : 0;
    P.depth[g] = 0; P.need_eval[g] = 0; P.active[g] = P.auto_play ? 1 : 0;
    P.tau[g] = P.init_temp; P.serial[g] = 0; P.rec_len[g] = 0;
    P.rng_ctr[g] = 0;
    long long* gs = P.gstat + (size_t)g * GS;
    for (int i = 0; i < GS; ++i) gs[i] = 0;
    P.free_top[g] = W.free_top;
    P.hiwater[g] = 0;
  }
}

template <int NCH>
__global__ void __launch_bounds__(WPB * 32) k_set_roots(const __grid_constant__ EP P, const int8_t* boards,
                                                       const int32_t* last, const uint8_t* act,
                                                       const uint8_t* clr) {
  __shared__ __align__(16) int8_t s_board[WPB][256];
  __shared__ __align__(16) float s_f[WPB][256];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * WPB + wid;
  if (g >= P.N) return;
  Warp<NCH> W(P, g, lane, s_board[wid], s_f[wid]);
  if (clr && clr[g]) {
    W.clear_table();
    if (lane == 0) P.tau[g] = P.init_temp;
  }
  const bool on = act ? act[g] != 0 : true;
  if (lane == 0) { P.active[g] = on; P.need_eval[g] = 0; P.depth[g] = 0; }
  if (on) {
    int8_t* sb = W.sb;
    for (int c = lane; c < P.KB; c += 32) sb[c] = c < P.C ? boards[(size_t)g * P.C + c] : 0;
    __syncwarp();
    for (int c = lane; c < P.KB; c += 32) P.root_board[(size_t)g * P.KB + c] = sb[c];
    uint32_t own[NCH], opp[NCH];
    W.board_masks(own, opp);
    W.collect(own, opp);
    int b = W.budget_for_root(own, opp);
    if (lane == 0) { P.root_last[g] = last ? last[g] : -1; P.sims_left[g] = b; }
  } else if (lane == 0) {
    P.sims_left[g] = 0;
  }
  W.flush();
}

template <int NCH>
__global__ void __launch_bounds__(WPB * 32) k_finish_move(const __grid_constant__ EP P, float* policy, int32_t* action) {
  __shared__ __align__(16) int8_t s_board[WPB][256];
  __shared__ __align__(16) float s_f[WPB][256];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * WPB + wid;
  if (g >= P.N) return;
  if (!P.active[g]) {
    if (lane == 0 && action) action[g] = -1;
    return;
  }
  Warp<NCH> W(P, g, lane, s_board[wid], s_f[wid]);
  int8_t* sb = W.sb;
  for (int c = lane; c < P.KB; c += 32) sb[c] = P.root_board[(size_t)g * P.KB + c];
  __syncwarp();
  uint32_t own[NCH], opp[NCH];
  W.board_masks(own, opp);
  int ep;
  int idx = W.find(W.hash_masks(own, opp), own, opp, &ep);
  int a = W.move_policy(idx >= 0 ? W.node(idx) : nullptr, policy ? policy + (size_t)g * P.C : nullptr);
  if (lane == 0 && action) action[g] = a;
  W.st[0] += 1;
  W.flush();
}

// Continuous batching of get_action calls (auto_play = 0; a5_engine_collect_moves / a5_engine_submit_roots).
// The searches of a batch end in different passes (budget rule of player.py:140-143 on re-used trees, terminal
// simulations that run two to a pass, leaves deferred by the evaluation cache).  k_collect_moves closes every search that has ended since the last call -- calc_policy
// (player.py:84-126), utils.step and utils.is_game_over of the played position -- into a compact row and parks the
// game (active = 0) until k_submit_roots gives it its next root; the other games keep searching in between.
template <int NCH>
__global__ void __launch_bounds__(WPB * 32) k_collect_moves(const __grid_constant__ EP P, int cap, int32_t* count, int32_t* game,
                                                           float* policy, int32_t* action, int8_t* next, int8_t* code) {
  __shared__ __align__(16) int8_t s_board[WPB][256];
  __shared__ __align__(16) float s_f[WPB][256];
  __shared__ uint32_t s_valid[WPB][4 * NCH];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * WPB + wid;
  if (g >= P.N) return;
  if (!P.active[g] || P.sims_left[g] > 0 || P.need_eval[g]) return;
  int row = 0;
  if (lane == 0) row = atomicAdd(count, 1);
  row = __shfl_sync(FULL, row, 0);
  if (row >= cap) return;                                  // no room this time: the search stays closed-but-uncollected
  Warp<NCH> W(P, g, lane, s_board[wid], s_f[wid]);
  warp_valid_masks<NCH>(P.S, P.goal, lane, s_valid[wid]);
  int8_t* sb = W.sb;
  for (int c = lane; c < P.KB; c += 32) sb[c] = P.root_board[(size_t)g * P.KB + c];
  __syncwarp();
  uint32_t own[NCH], opp[NCH];
  W.board_masks(own, opp);
  int ep;
  const int idx = W.find(W.hash_masks(own, opp), own, opp, &ep);
  const int a = W.move_policy(idx >= 0 ? W.node(idx) : nullptr, policy + (size_t)row * P.C);
  __syncwarp();
  if (a >= 0) warp_step(sb, P.C, a, lane);
  __syncwarp();
  W.board_masks(own, opp);
  const int cd = warp_terminal_bits<NCH>(own, opp, s_valid[wid], P.S, P.goal, lane);
  for (int c = lane; c < P.C; c += 32) next[(size_t)row * P.C + c] = sb[c];
  if (lane == 0) {
    game[row] = g;
    action[row] = a;
    code[row] = (int8_t)cd;
    P.active[g] = 0;
  }
  W.st[0] += 1;
  W.flush();
}

template <int NCH>
__global__ void __launch_bounds__(WPB * 32) k_submit_roots(const __grid_constant__ EP P, int n, const int32_t* game,
                                                          const int8_t* boards, const int32_t* last, const uint8_t* clr) {
  __shared__ __align__(16) int8_t s_board[WPB][256];
  __shared__ __align__(16) float s_f[WPB][256];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * WPB + wid;
  if (row >= n) return;
  const int g = game[row];
  if (g < 0 || g >= P.N) return;
  Warp<NCH> W(P, g, lane, s_board[wid], s_f[wid]);
  if (clr && clr[row]) {
    W.clear_table();
    if (lane == 0) P.tau[g] = P.init_temp;
  }
  int8_t* sb = W.sb;
  for (int c = lane; c < P.KB; c += 32) sb[c] = c < P.C ? boards[(size_t)row * P.C + c] : 0;
  __syncwarp();
  for (int c = lane; c < P.KB; c += 32) P.root_board[(size_t)g * P.KB + c] = sb[c];
  uint32_t own[NCH], opp[NCH];
  W.board_masks(own, opp);
  W.collect(own, opp);
  const int b = W.budget_for_root(own, opp);
  if (lane == 0) {
    P.active[g] = 1; P.need_eval[g] = 0; P.depth[g] = 0;
    P.root_last[g] = last ? last[row] : -1;
    P.sims_left[g] = b;
  }
  W.flush();
}

template <int NCH>
__global__ void __launch_bounds__(WPB * 32) k_node_stats(const __grid_constant__ EP P, const int8_t* boards, int32_t* dn,
                                                        float* dw, float* dp, int32_t* dsum) {
  __shared__ __align__(16) int8_t s_board[WPB][256];
  __shared__ __align__(16) float s_f[WPB][256];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * WPB + wid;
  if (g >= P.N) return;
  Warp<NCH> W(P, g, lane, s_board[wid], s_f[wid]);
  int8_t* sb = W.sb;
  for (int c = lane; c < P.KB; c += 32)
    sb[c] = boards ? (c < P.C ? boards[(size_t)g * P.C + c] : 0) : P.root_board[(size_t)g * P.KB + c];
  __syncwarp();
  uint32_t own[NCH], opp[NCH];
  W.board_masks(own, opp);
  int ep;
  int idx = W.find(W.hash_masks(own, opp), own, opp, &ep);
  uint8_t* nd = idx >= 0 ? W.node(idx) : nullptr;
  for (int c = lane; c < P.C; c += 32) {
    bool lg = nd && sb[c] == 0;
    if (dn) dn[(size_t)g * P.C + c] = lg ? W.edge_n(nd)[c] : 0;
    if (dw) dw[(size_t)g * P.C + c] = lg ? W.edge_w(nd)[c] : 0.0f;
    if (dp) dp[(size_t)g * P.C + c] = lg ? W.edge_p(nd)[c] : 0.0f;
  }
  if (lane == 0 && dsum) dsum[g] = nd ? W.hdr(nd)[0] : -1;
}

// Sampler export: draw i = one Dirichlet(alpha) vector over the first n_legal cells from Philox stream i,
// through the same gamma_cells the tree pass calls (distribution tests; player.py:240).
template <int NCH>
__global__ void __launch_bounds__(128) k_dirichlet(unsigned long long seed, float alpha, int n_legal, int n_draws, float* eta) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= n_draws) return;
  Philox rng(seed, (unsigned long long)i);
  bool legal[NCH];
  float gam[NCH];
  float gsum = 0.0f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) legal[k] = k * 32 + lane < n_legal;
  gamma_cells<NCH>(rng, 1ull, lane, alpha, legal, gam);
#pragma unroll
  for (int k = 0; k < NCH; ++k) gsum += gam[k];
  gsum = warp_sum(gsum);
#pragma unroll
  for (int k = 0; k < NCH; ++k)
    if (legal[k]) eta[(size_t)i * n_legal + k * 32 + lane] = (float)((double)gam[k] / (double)gsum);
}

template <int NCH>
__global__ void k_table_dump(const __grid_constant__ EP P, int g, int8_t* boards, int32_t* sum_n, int max_nodes,
                             int32_t* count) {
  // one warp; serial over nodes (test helper)
  const int lane = threadIdx.x & 31;
  const uint8_t* nodes = P.nodes + (size_t)g * P.cap * P.node_bytes;
  int hi = P.hiwater[g], cnt = 0;
  for (int idx = 0; idx < hi; ++idx) {
    const uint8_t* nd = nodes + (size_t)idx * P.node_bytes;
    const int32_t* hd = (const int32_t*)nd;
    if (hd[2] < 0) continue;
    if (cnt < max_nodes) {
      const uint32_t* mk = (const uint32_t*)(nd + 16);
      for (int c = lane; c < P.C; c += 32) {
        int k = c >> 5, b = c & 31;
        int v = ((mk[k] >> b) & 1u) ? 1 : (((mk[NCH + k] >> b) & 1u) ? -1 : 0);
        boards[(size_t)cnt * P.C + c] = (int8_t)v;
      }
      if (lane == 0) sum_n[cnt] = hd[0];
    }
    ++cnt;
  }
  if (lane == 0) *count = cnt;
}

__global__ void k_busy(const int32_t* sims_left, const uint8_t* need_eval, const uint8_t* active, int n, int32_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool b = i < n && active[i] && (sims_left[i] > 0 || need_eval[i]);
  unsigned m = __ballot_sync(FULL, b);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, __popc(m));
}

__global__ void k_reduce_stats(const long long* gstat, int n, long long* out) {
  // out[GS]; [7] is a max, the rest are sums
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int s = 0; s < GS; ++s) {
    long long v = i < n ? gstat[(size_t)i * GS + s] : 0;
    for (int o = 16; o; o >>= 1) {
      long long t = __shfl_xor_sync(FULL, v, o);
      v = s == 7 ? (t > v ? t : v) : v + t;
    }
    if ((threadIdx.x & 31) == 0 && v) {
      if (s == 7) atomicMax((long long*)&out[s], v);
      else atomicAdd((unsigned long long*)&out[s], (unsigned long long)v);
    }
  }
}

}  // namespace a5

// --------------------------------------------------------------------------- //
// host side
// --------------------------------------------------------------------------- //
using namespace a5;

struct a5_engine {
  a5_config cfg;
  EP p;
  void* arena = nullptr;           // one allocation for all per-game state
  size_t arena_bytes = 0;
  int32_t* d_scratch = nullptr;    // small device scratch (busy count, dump count)
  long long* d_stats = nullptr;
  int32_t* h_pinned = nullptr;
  long long* h_stats = nullptr;
  long long harvest_dropped = 0;   // records a harvest could not return (max_records too small)
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

namespace a5 {
bool pdl_enabled() {
  static const bool on = [] { const char* ev = getenv("A5_TC_PDL"); return !(ev && atoi(ev) == 0); }();
  return on;
}
}  // namespace a5

extern "C" {

int a5_record_stride(int S) {
  int C = S * S;
  return (int)align_up(sizeof(a5_record_header) + align_up(C, 16) + 4 * (size_t)C, 16);
}

int a5_engine_create(const a5_config* cfg, a5_engine** out) {
  A5_ARG(cfg && out);
  A5_ARG(cfg->board_size >= 5 && cfg->board_size <= A5_MAX_BOARD);
  A5_ARG(cfg->goal >= 2 && cfg->goal <= cfg->board_size);
  A5_ARG(cfg->n_games > 0 && cfg->sims > 0 && cfg->upper_sims >= 0);
  a5_engine* e = new a5_engine();
  e->cfg = *cfg;
  EP& p = e->p;
  memset(&p, 0, sizeof(p));
  p.S = cfg->board_size; p.C = p.S * p.S; p.goal = cfg->goal; p.N = cfg->n_games;
  p.sims = cfg->sims; p.upper = cfg->upper_sims; p.training = cfg->training; p.random_a = cfg->random_a;
  p.auto_play = cfg->auto_play;
  p.NCH = p.C <= 128 ? 4 : 8;
  p.E = 32 * p.NCH;
  int cap = cfg->node_capacity > 0 ? cfg->node_capacity : 2 * cfg->sims + 1024;
  if (cap > 32768) cap = 32768;
  p.cap = cap;
  int H = 64;
  while (H < 2 * cap) H <<= 1;
  p.H = H;
  // lock-step batches: a game that hits terminal positions yields after two NN-free simulations -- the pass
  // lasts as long as its slowest warp (measured at 4096 games: 2 already collects the +4 % moves per pass
  // that 16 gives, at a third of the tail); small batches (the B = 1 Player path) keep going to save
  // host round trips
  p.max_inner = cfg->max_inner > 0 ? cfg->max_inner : (cfg->n_games >= 256 ? 2 : 16);
  p.node_bytes = 16 + 8 * p.NCH + 12 * p.E;
  p.KB = (int)align_up(p.C, 16);
  p.rec_bb = p.KB;
  p.rec_stride = a5_record_stride(p.S);
  p.rec_cap = cfg->record_capacity > 0 ? cfg->record_capacity : (cfg->auto_play ? p.N * p.C : 1);
  p.c_puct = cfg->c_puct; p.alpha = cfg->dirichlet_alpha; p.gamma = cfg->gamma;
  p.init_temp = cfg->init_temp; p.tau_decay = cfg->tau_decay; p.tau_decay_r = cfg->tau_decay_r;
  p.seed = cfg->seed; p.gid_base = cfg->game_id_base;

  // carve one arena
  size_t off = 0;
  const size_t N = p.N;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  size_t o_nodes = take(N * p.cap * (size_t)p.node_bytes);
  size_t o_slots = take(N * p.H * 4);
  size_t o_fstack = take(N * p.cap * 2);
  size_t o_ftop = take(N * 4), o_hi = take(N * 4);
  size_t o_rb = take(N * p.KB), o_lb = take(N * p.KB);
  size_t o_rl = take(N * 4), o_ll = take(N * 4), o_sl = take(N * 4), o_dp = take(N * 4), o_ls = take(N * 4);
  size_t o_lh = take(N * 4);
  size_t o_ne = take(N), o_ac = take(N);
  size_t o_path = take(N * p.C * 4);
  size_t o_ctr = take(N * 8), o_tau = take(N * 8);
  size_t o_ser = take(N * 4), o_rlen = take(N * 4);
  size_t recN = p.auto_play ? N : 1;
  size_t o_rst = take(recN * p.C * (size_t)p.rec_stride);
  size_t o_out = take((size_t)p.rec_cap * p.rec_stride), o_oc = take(16);
  size_t o_gs = take(N * GS * 8);
  size_t o_pl = take(N * 3 * p.C);
  e->arena_bytes = off;
  cudaError_t err = cudaMalloc(&e->arena, off);
  if (err != cudaSuccess) {
    set_error("a5_engine_create: cudaMalloc(%zu bytes) failed: %s", off, cudaGetErrorString(err));
    delete e;
    return A5_ERR_CUDA;
  }
  uint8_t* b = (uint8_t*)e->arena;
  p.nodes = b + o_nodes; p.slots = (uint32_t*)(b + o_slots); p.free_stack = (uint16_t*)(b + o_fstack);
  p.free_top = (int32_t*)(b + o_ftop); p.hiwater = (int32_t*)(b + o_hi);
  p.root_board = (int8_t*)(b + o_rb); p.leaf_board = (int8_t*)(b + o_lb);
  p.root_last = (int32_t*)(b + o_rl); p.leaf_last = (int32_t*)(b + o_ll); p.sims_left = (int32_t*)(b + o_sl);
  p.depth = (int32_t*)(b + o_dp); p.leaf_slot = (int32_t*)(b + o_ls); p.leaf_hash = (uint32_t*)(b + o_lh);
  p.need_eval = b + o_ne; p.active = b + o_ac;
  p.path = (uint32_t*)(b + o_path);
  p.rng_ctr = (unsigned long long*)(b + o_ctr); p.tau = (double*)(b + o_tau);
  p.serial = (int32_t*)(b + o_ser); p.rec_len = (int32_t*)(b + o_rlen);
  p.rec_stage = b + o_rst;
  p.out_rec = b + o_out; p.out_count = (int32_t*)(b + o_oc);
  p.gstat = (long long*)(b + o_gs);
  p.planes = (int8_t*)(b + o_pl);
  A5_CUDA(cudaMalloc(&e->d_scratch, 64));
  A5_CUDA(cudaMalloc(&e->d_stats, GS * 8));
  A5_CUDA(cudaMallocHost(&e->h_pinned, 64));
  A5_CUDA(cudaMallocHost(&e->h_stats, GS * 8));
  // zero the small state (not the node arena)
  A5_CUDA(cudaMemset(b + o_slots, 0, off - o_slots));
  *out = e;
  int rc = a5_engine_reset(e, nullptr);
  if (rc) return rc;
  A5_CUDA(cudaDeviceSynchronize());
  return A5_OK;
}

int a5_engine_destroy(a5_engine* e) {
  if (!e) return A5_OK;
  cudaFree(e->arena); cudaFree(e->d_scratch); cudaFree(e->d_stats);
  cudaFreeHost(e->h_pinned); cudaFreeHost(e->h_stats);
  delete e;
  return A5_OK;
}

#define DISPATCH(kern, grid, block, st, ...)                                   \
  do {                                                                         \
    if (e->p.NCH == 4) kern<4><<<grid, block, 0, st>>>(__VA_ARGS__);           \
    else kern<8><<<grid, block, 0, st>>>(__VA_ARGS__);                         \
    A5_CUDA(cudaGetLastError());                                               \
  } while (0)

static inline int ngrid(const a5_engine* e) { return (e->p.N + WPB - 1) / WPB; }

int a5_engine_reset(a5_engine* e, void* stream) {
  A5_ARG(e);
  cudaStream_t st = (cudaStream_t)stream;
  A5_CUDA(cudaMemsetAsync(e->p.out_count, 0, 16, st));
  e->harvest_dropped = 0;
  DISPATCH(k_reset, ngrid(e), WPB * 32, st, e->p);
  return A5_OK;
}

int a5_engine_set_roots(a5_engine* e, const int8_t* d_boards, const int32_t* d_last, const uint8_t* d_active,
                        const uint8_t* d_clear, void* stream) {
  A5_ARG(e && d_boards);
  if (e->p.auto_play) { set_error("a5_engine_set_roots: engine is in auto_play mode"); return A5_ERR_STATE; }
  DISPATCH(k_set_roots, ngrid(e), WPB * 32, (cudaStream_t)stream, e->p, d_boards, d_last, d_active, d_clear);
  return A5_OK;
}

int a5_engine_step_served(a5_engine* e, const float* d_prob, const float* d_value, const uint8_t* d_served, void* stream) {
  A5_ARG(e && d_prob && d_value && d_served);
  e->p.served = d_served;
  int rc = a5_engine_step(e, d_prob, d_value, stream);
  e->p.served = nullptr;
  return rc;
}

int a5_engine_step(a5_engine* e, const float* d_prob, const float* d_value, void* stream) {
  A5_ARG(e && ((d_prob == nullptr) == (d_value == nullptr)));
  e->p.kt = kt_slot(KT_STEP);
  // programmatic dependent launch (common.cuh): scheduled behind the tail of the heads kernel
  if (e->p.NCH == 4) A5_CUDA(launch_pdl_k(k_step<4>, (unsigned)ngrid(e), WPB * 32, 0, (cudaStream_t)stream, pdl_enabled(), e->p, d_prob, d_value));
  else A5_CUDA(launch_pdl_k(k_step<8>, (unsigned)ngrid(e), WPB * 32, 0, (cudaStream_t)stream, pdl_enabled(), e->p, d_prob, d_value));
  return A5_OK;
}

int a5_engine_set_mode(a5_engine* e, int training, int random_a) {
  A5_ARG(e);
  e->p.training = training != 0;
  e->p.random_a = random_a != 0;
  return A5_OK;
}

int a5_engine_set_budget(a5_engine* e, int sims, int upper_sims) {
  A5_ARG(e && sims > 0 && upper_sims >= 0);
  e->p.sims = sims;
  e->p.upper = upper_sims;
  return A5_OK;
}

int8_t* a5_engine_planes(a5_engine* e) { return e ? e->p.planes : nullptr; }
uint8_t* a5_engine_need_eval(a5_engine* e) { return e ? e->p.need_eval : nullptr; }
int32_t* a5_engine_sims_left(a5_engine* e) { return e ? e->p.sims_left : nullptr; }

int a5_engine_busy(a5_engine* e, int32_t* h_busy, void* stream) {
  A5_ARG(e && h_busy);
  cudaStream_t st = (cudaStream_t)stream;
  A5_CUDA(cudaMemsetAsync(e->d_scratch, 0, 4, st));
  k_busy<<<(e->p.N + 255) / 256, 256, 0, st>>>(e->p.sims_left, e->p.need_eval, e->p.active, e->p.N, e->d_scratch);
  A5_CUDA(cudaGetLastError());
  A5_CUDA(cudaMemcpyAsync(e->h_pinned, e->d_scratch, 4, cudaMemcpyDeviceToHost, st));
  A5_CUDA(cudaStreamSynchronize(st));
  *h_busy = e->h_pinned[0];
  return A5_OK;
}

int a5_engine_finish_move(a5_engine* e, float* d_policy, int32_t* d_action, void* stream) {
  A5_ARG(e);
  if (e->p.auto_play) { set_error("a5_engine_finish_move: engine is in auto_play mode"); return A5_ERR_STATE; }
  DISPATCH(k_finish_move, ngrid(e), WPB * 32, (cudaStream_t)stream, e->p, d_policy, d_action);
  return A5_OK;
}

int a5_engine_collect_moves(a5_engine* e, int cap, int32_t* d_count, int32_t* d_game, float* d_policy, int32_t* d_action,
                            int8_t* d_next, int8_t* d_code, void* stream) {
  A5_ARG(e && cap > 0 && d_count && d_game && d_policy && d_action && d_next && d_code);
  if (e->p.auto_play) { set_error("a5_engine_collect_moves: engine is in auto_play mode"); return A5_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  A5_CUDA(cudaMemsetAsync(d_count, 0, 4, st));
  DISPATCH(k_collect_moves, ngrid(e), WPB * 32, st, e->p, cap, d_count, d_game, d_policy, d_action, d_next, d_code);
  return A5_OK;
}

int a5_engine_submit_roots(a5_engine* e, int n, const int32_t* d_game, const int8_t* d_boards, const int32_t* d_last,
                           const uint8_t* d_clear, void* stream) {
  A5_ARG(e && n >= 0 && (n == 0 || (d_game && d_boards)));
  if (e->p.auto_play) { set_error("a5_engine_submit_roots: engine is in auto_play mode"); return A5_ERR_STATE; }
  if (n == 0) return A5_OK;
  DISPATCH(k_submit_roots, (n + WPB - 1) / WPB, WPB * 32, (cudaStream_t)stream, e->p, n, d_game, d_boards, d_last, d_clear);
  return A5_OK;
}

int a5_engine_root_stats(a5_engine* e, int32_t* d_n, float* d_w, float* d_p, int32_t* d_sum_n, void* stream) {
  A5_ARG(e);
  DISPATCH(k_node_stats, ngrid(e), WPB * 32, (cudaStream_t)stream, e->p, (const int8_t*)nullptr, d_n, d_w, d_p, d_sum_n);
  return A5_OK;
}

int a5_engine_node_stats(a5_engine* e, const int8_t* d_boards, int32_t* d_n, float* d_w, float* d_p, int32_t* d_sum_n,
                         void* stream) {
  A5_ARG(e && d_boards);
  DISPATCH(k_node_stats, ngrid(e), WPB * 32, (cudaStream_t)stream, e->p, d_boards, d_n, d_w, d_p, d_sum_n);
  return A5_OK;
}

double* a5_engine_tau(a5_engine* e) { return e ? e->p.tau : nullptr; }

int a5_engine_get_roots(a5_engine* e, int8_t* d_boards, int32_t* d_last, void* stream) {
  A5_ARG(e && d_boards);
  cudaStream_t st = (cudaStream_t)stream;
  A5_CUDA(cudaMemcpy2DAsync(d_boards, e->p.C, e->p.root_board, e->p.KB, e->p.C, e->p.N, cudaMemcpyDeviceToDevice, st));
  if (d_last) A5_CUDA(cudaMemcpyAsync(d_last, e->p.root_last, sizeof(int32_t) * e->p.N, cudaMemcpyDeviceToDevice, st));
  return A5_OK;
}

int a5_dirichlet_sample(uint64_t seed, float alpha, int n_legal, int n_draws, float* d_eta, void* stream) {
  A5_ARG(d_eta && n_legal > 0 && n_legal <= 256 && n_draws > 0 && alpha > 0.0f);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_legal <= 128) k_dirichlet<4><<<(n_draws + 3) / 4, 128, 0, st>>>(seed, alpha, n_legal, n_draws, d_eta);
  else k_dirichlet<8><<<(n_draws + 3) / 4, 128, 0, st>>>(seed, alpha, n_legal, n_draws, d_eta);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}

int a5_engine_table_dump(a5_engine* e, int game, int8_t* d_boards, int32_t* d_sum_n, int max_nodes, int32_t* h_count,
                         void* stream) {
  A5_ARG(e && game >= 0 && game < e->p.N && d_boards && d_sum_n && h_count && max_nodes >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH(k_table_dump, 1, 32, st, e->p, game, d_boards, d_sum_n, max_nodes, e->d_scratch + 4);
  A5_CUDA(cudaMemcpyAsync(e->h_pinned + 4, e->d_scratch + 4, 4, cudaMemcpyDeviceToHost, st));
  A5_CUDA(cudaStreamSynchronize(st));
  *h_count = e->h_pinned[4];
  return A5_OK;
}

int a5_engine_harvest(a5_engine* e, void* d_out, int max_records, int32_t* h_count, int32_t* h_games, void* stream) {
  A5_ARG(e && h_count && (d_out || max_records == 0) && max_records >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  A5_CUDA(cudaMemcpyAsync(e->h_pinned + 8, e->p.out_count, 12, cudaMemcpyDeviceToHost, st));
  A5_CUDA(cudaStreamSynchronize(st));
  const int cnt = e->h_pinned[8], games = e->h_pinned[9], arena_full = e->h_pinned[10];
  const int ncopy = cnt < max_records ? cnt : max_records;
  if (ncopy > 0)
    A5_CUDA(cudaMemcpyAsync(d_out, e->p.out_rec, (size_t)ncopy * e->p.rec_stride, cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemsetAsync(e->p.out_count, 0, 12, st));
  const int lost = arena_full + (cnt - ncopy);     // plies of games that found the arena full + plies cut off here
  e->harvest_dropped += lost;
  *h_count = ncopy;
  if (h_games) *h_games = games;
  if (lost > 0) {
    // whole games are lost, never parts of one (emit_game reserves a game's plies atomically), but the
    // replay buffer would silently miss them: the caller must harvest more often or raise record_capacity
    set_error("a5_engine_harvest: %d finished-ply records lost (%d did not fit the record arena of %d, %d beyond "
              "max_records = %d)", lost, arena_full, e->p.rec_cap, cnt - ncopy, max_records);
    return A5_ERR_CAPACITY;
  }
  return A5_OK;
}

int a5_engine_counters(a5_engine* e, int64_t* h_out, void* stream) {
  A5_ARG(e && h_out);
  cudaStream_t st = (cudaStream_t)stream;
  A5_CUDA(cudaMemsetAsync(e->d_stats, 0, GS * 8, st));
  k_reduce_stats<<<(e->p.N + 255) / 256, 256, 0, st>>>(e->p.gstat, e->p.N, e->d_stats);
  A5_CUDA(cudaGetLastError());
  A5_CUDA(cudaMemcpyAsync(e->h_stats, e->d_stats, GS * 8, cudaMemcpyDeviceToHost, st));
  A5_CUDA(cudaMemcpyAsync(e->h_pinned + 12, e->p.out_count + 2, 4, cudaMemcpyDeviceToHost, st));
  A5_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < A5_NUM_COUNTERS; ++i) h_out[i] = i < GS ? e->h_stats[i] : 0;
  h_out[10] += e->h_pinned[12] + e->harvest_dropped;
  if (h_out[8] > 0) {
    set_error("a5_engine: %lld leaf expansions did not fit the per-game node arena (node_capacity=%d)",
              (long long)h_out[8], e->p.cap);
    return A5_ERR_CAPACITY;
  }
  return A5_OK;
}

// internal tooling (not part of alphafive.h): switch the per-game k_step timing on/off, read it back
// (h_out: uint64 [2][8192] = ns, flags of the last pass).
int a5__debug_step_times(int on, unsigned long long* h_out) {
  A5_CUDA(cudaMemcpyToSymbol(a5::g_step_dbg_on, &on, sizeof(int)));
  if (h_out) A5_CUDA(cudaMemcpyFromSymbol(h_out, a5::g_step_dbg, sizeof(unsigned long long) * 2 * 8192));
  return A5_OK;
}
}  // extern "C"
