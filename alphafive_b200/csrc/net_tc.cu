// Tensor-core forward pass of the policy/value net (A5_NET_TC): tcgen05.mma + TMEM
// accumulators + bulk-async (TMA engine) operand staging, fp32-faithful.
//
// Why not plain bf16: the reference graph is fp32 and parity is 1e-4 on policy/value;
// bf16 or tf32 operands miss that by 10-100x on trained weights.  Every
// operand is therefore split x = hi + lo with hi, lo fp16 (22 significant bits) and each
// product runs as three MMAs  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  into one fp32 TMEM
// accumulator.  Activations are pre-scaled by 2^4 and weights by 2^10 (exact) so that the
// lo halves stay in fp16's normal range; the epilogue multiplies by 2^-14.
//
// Conv as implicit GEMM over the padded position space (net_common.cuh):
//   D[128 positions x Cout] += A[128 x 16] * B[16 x Cout]      per tcgen05.mma (M=128, K=16)
//   A = activations, K-major, no swizzle: HBM/SMEM layout [hi|lo][channel/8][position][8],
//       so one core matrix is 8 consecutive positions x 8 channels = 128 contiguous bytes
//       and a 3x3 tap is just a different (16-byte aligned) start address in the SAME staged
//       slab -- the slab is loaded once with a halo and reused by all 9 taps;
//   B = weights [hi|lo][channel/8][Cout][8], pre-packed per (32-channel slab, tap) stage.
// The 1x1 projection of a residual block is one more K segment of conv2 reading the block
// input, so skip-add + bias + ELU + hi/lo re-split are all in the epilogue.
//
// One persistent CTA per SM, clusters of two (tcgen05 cta_group::2): see k_tc_conv2 below for the warp
// roles.  conv1 (5x5 on {0,1} planes) has its own kernel (k_tc_conv1m); the dense heads are in
// net_heads_tc.cu.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "net.cuh"
#include "tc_ptx.cuh"

namespace a5 {

constexpr int TC_HALO = 24;             // largest slab halo: >= pitch + 1 for S <= 15, multiple of 8
constexpr int TC_KS = 32;               // channels per slab
constexpr int TC_EPI_WARPS = 16;        // four per TMEM lane quadrant
constexpr int TC_W_WARP = 2 + TC_EPI_WARPS;   // weight producer (its own warp: slab prefetch must not wait on weight-ring credits)
// T = M tiles (of 128 positions) per CTA and group
template <int T>
struct TCfg {
  static constexpr int ROWS = T * 128;                 // positions per group
};

struct TCLayer {
  const __half* src; int src_ch;
  const __half* src2; int src2_ch;  // optional second 3x3 segment (two convs of different inputs merged into one layer)
  int n0;                          // output columns of the first segment (== cout without src2); src2 fills [n0, cout)
  const __half* res; int res_ch;
  const __half* wpk;
  const float* bias;
  __half* out;
  __half* out2; int split;       // merged layers: channels [0, split) -> out, [split, cout) -> out2
  float* out_f32;
  int cout, ntaps;
  int fold;                      // 1: hi*[Whi|Wlo] as one N = 2*cout MMA (cout <= 64)
  int reverse;                   // 1: walk the groups from the last to the first.  Layers alternate direction so
                                 // that each starts on the rows its predecessor touched last: ~100 MB of every
                                 // 150-450 MB activation tensor are then still in the 126 MB L2
  uint32_t lo_add, lo_mask;      // rounding of the lo halves to fewer mantissa bits (a5_tc_state::lo_drop), both fp16 lanes
  int nslab_buf, w_bytes;        // pair kernel: A slab buffers, bytes of the weight region (ring or resident set)
  unsigned long long* dbg;       // tooling: clock64 timeline of CTA 0 (4 roles x 256 slots), or null
  unsigned long long* kt;        // tooling: in-situ kernel timing slot (common.cuh), or null
  unsigned long long* clk;       // tooling: {clock64, globaltimer} at start and end of CTA 0 (SM clock of this launch), or null
  // fused 1x1 head conv + ELU in the epilogue (network.py:69-70 value, :81-82 policy): the
  // layer's own activation is then not stored; the head output goes out as the A operand of
  // the dense layer that follows (k_tc_fc), K ordered (cell, channel).
  int head_ch;                   // 0 none, 4 value head, 16 policy head
  const float* head_w;           // f32 [32][head_ch]
  const float* head_b;           // f32 [head_ch]
  __half* head_out;              // [mtile][stage][hi|lo][kchunk 4][128 boards][8]
  int head_nst;                  // K stages (of 32) per board row
  int shifts[9];
  long long plane_rows;          // rows per channel-chunk plane in HBM (incl. guards)
  long long row0;                // first valid row (guard)
  long long nrows;               // valid rows (boards * per_board)
  int ngroups;
  int S, pitch, per_board;
};

struct __align__(8) TCBarriers {
  uint64_t a_full[4], a_empty[4], w_full[8], w_empty[8], t_full[2], t_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(TCBarriers) <= 256, "barrier block outgrew its smem reservation");

__device__ __forceinline__ unsigned long long dbg_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));   // ns, comparable across SMs
  return t;
}
__device__ __forceinline__ void dbg_mark(unsigned long long* dbg, int role, int& n) {
  if (dbg && blockIdx.x == 0 && n < 256) dbg[role * 256 + n++] = dbg_now();
}
// roles 4..6 (pair kernel): W producer of the leader, relay of the peer, W producer of the peer
__device__ __forceinline__ void dbg_mark_cta(unsigned long long* dbg, int role, int& n, int cta) {
  if (dbg && (int)blockIdx.x == cta && n < 256) dbg[role * 256 + n++] = dbg_now();
}

// ------------------------------------------------------------------ epilogue (shared by the conv kernels)
// One group: TMEM -> bias / ELU / hi-lo split -> HBM (or the fused 1x1 head conv) for the T tiles whose accumulators
// start at TMEM column `tmem_buf`; q0 = row index (from row0) of the group's first position, rows >= qlimit are not
// stored.  Four warps per TMEM lane quadrant (a warp may only read lanes 32*(warp%4)..+31); the (M tile, 16-column)
// units of a group are dealt round-robin to them.  All arithmetic is on values pre-scaled by ACT_SCALE.
template <int T>
__device__ __forceinline__ void tc_epilogue_group(const TCLayer& L, const uint32_t tmem_buf, const float* s_bias, const float* s_hw,
                                                  const int warp, const int lane, const int cpt, const uint32_t q0,
                                                  const uint32_t qlimit) {
  const int cout = L.cout;
  const int quad = warp & 3;
  const int sub = (warp - 2) >> 2;
  constexpr float K_ACC = OUT_SCALE * ACT_SCALE;          // accumulator -> scaled activation
  constexpr float K_L2E = 1.4426950408889634f / ACT_SCALE;
  if (L.head_ch) {
    // head layers: columns [0, 32) of every tile feed the fused 1x1 head conv -- a warp takes whole rows
    // of one M tile, so the head conv sees all 32 channels of its position (cout > 32: the remaining
    // columns are plain units, below)
    // the four warps of a lane quadrant share the T tiles; with T = 2 two warps split the 16 policy-head
    // outputs of a tile between them (each recomputes the 32 activations it needs)
    constexpr int NPART = (TC_EPI_WARPS / 4) / T > 0 ? (TC_EPI_WARPS / 4) / T : 1;
    const int part = sub / T;
    for (int m = sub % T; m < T && (part == 0 || L.head_ch == 16); m += T * NPART) {
      const uint32_t q = q0 + (uint32_t)(m * 128 + quad * 32 + lane);
      const uint32_t board = q / (uint32_t)L.per_board, within = q - board * (uint32_t)L.per_board;
      const uint32_t rr = within / (uint32_t)L.pitch, cc = within - rr * (uint32_t)L.pitch;
      const bool real = q < qlimit && rr < (uint32_t)L.S && cc < (uint32_t)L.S;
      float f[32];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t v[16];
        const uint32_t taddr = tmem_buf + ((uint32_t)(quad * 32) << 16) + (uint32_t)(m * cpt + h2 * 16);
        tc_ld16(taddr, v);
        if (L.fold) {
          uint32_t v2[16];
          tc_ld16(taddr + (uint32_t)cout, v2);
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          tc_ld_wait();
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b4 = *(const float4*)&s_bias[h2 * 16 + 4 * e];
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = fmaf(__uint_as_float(v[4 * e + j]), K_ACC, bb[j]);
            const float neg = fmaf(ex2_approx(x * K_L2E), ACT_SCALE, -ACT_SCALE);
            f[h2 * 16 + 4 * e + j] = x > 0.0f ? x : neg;       // ACT_SCALE * activation
          }
        }
      }
      if (real) {
        const uint32_t cell = rr * (uint32_t)L.S + cc;
        const uint32_t mt = board >> 7, brow = board & 127u;
        if (L.head_ch == 16) {
          // k = cell*16 + c -> stage = cell/2, kchunk = (cell%2)*2 + c/8; eight outputs at a time
          __half* base = L.head_out + ((((size_t)mt * L.head_nst + (cell >> 1)) * 2) * 4 + (cell & 1u) * 2) * 128 * 8 + brow * 8;
#pragma unroll 1
          for (int c8 = (NPART >= 2 ? part : 0); c8 < (NPART >= 2 ? part + 1 : 2); ++c8) {
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = s_hw[32 * 16 + c8 * 8 + c];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float4 w0 = *(const float4*)&s_hw[k * 16 + c8 * 8];
              const float4 w1 = *(const float4*)&s_hw[k * 16 + c8 * 8 + 4];
              acc[0] = fmaf(f[k], w0.x, acc[0]); acc[1] = fmaf(f[k], w0.y, acc[1]);
              acc[2] = fmaf(f[k], w0.z, acc[2]); acc[3] = fmaf(f[k], w0.w, acc[3]);
              acc[4] = fmaf(f[k], w1.x, acc[4]); acc[5] = fmaf(f[k], w1.y, acc[5]);
              acc[6] = fmaf(f[k], w1.z, acc[6]); acc[7] = fmaf(f[k], w1.w, acc[7]);
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a0 = acc[2 * e], a1 = acc[2 * e + 1];
              const float x0 = a0 > 0.0f ? a0 : fmaf(ex2_approx(a0 * K_L2E), ACT_SCALE, -ACT_SCALE);
              const float x1 = a1 > 0.0f ? a1 : fmaf(ex2_approx(a1 * K_L2E), ACT_SCALE, -ACT_SCALE);
              const __half2 h = __floats2half2_rn(x0, x1);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
              hi[e] = *(const uint32_t*)&h;
              lo[e] = *(const uint32_t*)&l;
            }
            *(uint4*)(base + (size_t)c8 * 128 * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *(uint4*)(base + (size_t)(4 + c8) * 128 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        } else {
          float acc[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[c] = s_hw[32 * 16 + c];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float4 w4 = *(const float4*)&s_hw[k * 4];
            acc[0] = fmaf(f[k], w4.x, acc[0]);
            acc[1] = fmaf(f[k], w4.y, acc[1]);
            acc[2] = fmaf(f[k], w4.z, acc[2]);
            acc[3] = fmaf(f[k], w4.w, acc[3]);
          }
          // k = cell*4 + c -> stage = cell/8, kchunk = (cell%8)/2, element = (cell%2)*4 + c
          __half* base = L.head_out + ((((size_t)mt * L.head_nst + (cell >> 3)) * 2) * 4 + ((cell & 7u) >> 1)) * 128 * 8 +
                         brow * 8 + (cell & 1u) * 4;
          uint32_t hi[2], lo[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float a0 = acc[2 * e], a1 = acc[2 * e + 1];
            const float x0 = a0 > 0.0f ? a0 : fmaf(ex2_approx(a0 * K_L2E), ACT_SCALE, -ACT_SCALE);
            const float x1 = a1 > 0.0f ? a1 : fmaf(ex2_approx(a1 * K_L2E), ACT_SCALE, -ACT_SCALE);
            const __half2 h = __floats2half2_rn(x0, x1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
            hi[e] = *(const uint32_t*)&h;
            lo[e] = *(const uint32_t*)&l;
          }
          *(uint2*)base = make_uint2(hi[0], hi[1]);
          *(uint2*)(base + (size_t)4 * 128 * 8) = make_uint2(lo[0], lo[1]);
        }
      }
    }
  }
  int cur_m = -1;
  bool real = false;
  uint32_t q = 0;
  const int hcols = L.head_ch ? 32 : 0;                    // columns consumed by the head path above
  const int ncp = (cout - hcols) >> 4;                     // plain 16-column units per tile
  for (int u = sub; u < T * ncp; u += TC_EPI_WARPS / 4) {
    const int m = u / ncp, c0 = hcols + ((u - m * ncp) << 4);
    if (m != cur_m) {
      cur_m = m;
      q = q0 + (uint32_t)(m * 128 + quad * 32 + lane);      // row index from row0
      const uint32_t within = q % (uint32_t)L.per_board;
      const uint32_t rr = within / (uint32_t)L.pitch, cc = within - rr * (uint32_t)L.pitch;
      real = q < qlimit && rr < (uint32_t)L.S && cc < (uint32_t)L.S;
    }
    const long long row = L.row0 + q;
    uint32_t v[16];
    const uint32_t taddr = tmem_buf + ((uint32_t)(quad * 32) << 16) + (uint32_t)(m * cpt + c0);
    tc_ld16(taddr, v);
    if (L.fold) {                                         // + a_hi * w_lo partial sums
      uint32_t v2[16];
      tc_ld16(taddr + (uint32_t)cout, v2);
      tc_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
    } else {
      tc_ld_wait();
    }
    float f[16];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 b4 = *(const float4*)&s_bias[c0 + 4 * e];
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = fmaf(__uint_as_float(v[4 * e + j]), K_ACC, bb[j]);
        const float neg = fmaf(ex2_approx(x * K_L2E), ACT_SCALE, -ACT_SCALE);
        f[4 * e + j] = x > 0.0f ? x : neg;
      }
    }
    if (L.out || L.out2) {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = f[kc * 8 + 2 * e], x1 = f[kc * 8 + 2 * e + 1];
          const __half2 h = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          hi[e] = real ? *(const uint32_t*)&h : 0u;
          lo[e] = real ? ((*(const uint32_t*)&l + L.lo_add) & L.lo_mask) : 0u;
        }
        const bool second = L.out2 && c0 >= L.split;
        __half* ob = second ? L.out2 : L.out;
        const int och = L.out2 ? (second ? cout - L.split : L.split) : cout;
        const long long chunk = ((second ? c0 - L.split : c0) >> 3) + kc;
        __half* ph = ob + (chunk * L.plane_rows + row) * 8;
        __half* pl = ob + (((long long)(och >> 3) + chunk) * L.plane_rows + row) * 8;
        if (q < qlimit) {                                    // rows past the limit belong to another chunk / CTA
          *(uint4*)ph = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *(uint4*)pl = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    if (L.out_f32 && q < qlimit) {
      constexpr float inv = 1.0f / ACT_SCALE;
      float4* po = (float4*)(L.out_f32 + row * cout + c0);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        po[e] = real ? make_float4(f[4 * e] * inv, f[4 * e + 1] * inv, f[4 * e + 2] * inv, f[4 * e + 3] * inv)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// The groups g_first, g_first + g_step, ... < g_end of this CTA (one layer per launch).  REMOTE: the "TMEM drained"
// arrival goes to the leader CTA of the pair.
template <int T, bool REMOTE>
__device__ __forceinline__ void tc_epilogue(const TCLayer& L, TCBarriers* B, const uint32_t tmem, const float* s_bias,
                                            const float* s_hw, const int warp, const int lane, const int cpt,
                                            const int nbuf, const int g_first, const int g_end, const int g_step) {
  using Cfg = TCfg<T>;
  int tb = 0, tph = 0, dn = 0;
  unsigned long long* edbg = (warp == 2 && lane == 0) ? L.dbg : nullptr;
  for (int g = g_first; g < g_end; g += g_step) {
    mbar_wait(&B->t_full[tb], tph);
    tc_fence_after();
    dbg_mark(edbg, 2, dn);                               // accumulators ready
    tc_epilogue_group<T>(L, tmem + (uint32_t)(tb * T * cpt), s_bias, s_hw, warp, lane, cpt,
                         (uint32_t)(L.reverse ? g_end - 1 - g : g) * Cfg::ROWS, (uint32_t)L.nrows);
    tc_fence_before();
    __syncwarp();
    dbg_mark(edbg, 2, dn);                               // tile(s) drained
    if (lane == 0) {
      if (REMOTE) mbar_arrive_remote(&B->t_empty[tb], 0);   // the leader CTA's barrier (cta_group::2)
      else mbar_arrive(&B->t_empty[tb]);
    }
    if (++tb == nbuf) { tb = 0; tph ^= 1; }
  }
}

// ------------------------------------------------------------------ the CTA-pair conv kernel
// Same computation as k_tc_conv on a pair of SMs (cluster of 2, tcgen05 cta_group::2): one
// tcgen05.mma covers M = 256 -- the T tiles of this CTA and the T tiles of its peer -- while the
// B operand (weights) is split along N between the two CTAs' shared memories.  Per SM that
//   * halves the weight bytes staged from L2 and read by the tensor core (an N = 128 SS-MMA
//     otherwise needs A 4 KB + B 4 KB per 64 cycles = all 128 B/clk of shared memory),
//   * lets a weight stage serve 2T tiles while each CTA's TMEM holds only T (double-buffered).
// Only the leader CTA (cluster rank 0) issues MMAs.  Data arrival in the peer is relayed to the
// leader's "full" barriers by the peer's otherwise idle warp 1 (mapa + remote mbarrier.arrive);
// MMA completion is multicast to both CTAs' "empty" / "accumulator ready" barriers
// (tcgen05.commit ... multicast::cluster); the peer's epilogue warps report "TMEM drained" on the
// leader's barrier.
//
// Weight stage per CTA [kchunk 4][X rows | S rows][8]:
//   fold (cout <= 64): X = cout rows (leader: w_hi, peer: w_lo)  -> a_hi * [w_hi | w_lo], N = 2 cout
//                      S = cout/2 rows (leader: w_hi[0:cout/2], peer: w_hi[cout/2:]) -> a_lo * w_hi
//   else             : X = cout/2 rows of w_hi, S = cout/2 rows of w_lo (this CTA's half of N)
//
// Issue side (measured, tools/micro/mma_bench7/8.cu): one thread issues a tcgen05.mma every ~50 cycles
// and the tensor core needs >= ~40 cycles per instruction (the 4 KB A operand at 128 B/clk), so with
// N <= 128 a single issuer that also polls a weight barrier and commits per (slab, tap) stage cannot
// keep the pipe fed.  Hence (i) TWO issuer warps, each owning half of the CTA's M tiles (separate
// accumulators: no ordering between them is needed; every "empty"/"ready" barrier counts both
// commits), and (ii) RESW: when the layer's whole weight set fits beside the slabs it is loaded once
// per CTA and stays resident -- no per-stage wait / commit, and 1/16 of the L2 -> SMEM weight traffic.
constexpr int TC2_THREADS = 32 * (2 + TC_EPI_WARPS + 2);   // producer, issuer/relay, epilogue, W producer, 2nd issuer
constexpr int TC2_MMA_WARP1 = TC_W_WARP + 1;              // second MMA issuer (leader CTA only)
constexpr int TC2_WSTAGES = 8;
constexpr int TC2_WSTAGE_MAX = 4 * 128 * 16;              // 8 KB (cout = 128: 64 + 64 rows)
constexpr int TC2_MISC = 128 * 4 + 256 + (32 * 16 + 16) * 4 + 128;
constexpr int TC2_SMEM_LIMIT = 232448;                    // 227 KB opt-in maximum per CTA

__device__ __forceinline__ void tc_commit2(uint64_t* bar) {   // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// slab geometry of the pair kernel: the halo is a template parameter (>= pitch + 1, multiple of 8: 16 for
// boards up to 14x14, 24 for 15x15) -- the 2 KB per slab it saves at 11x11 buy a third slab buffer
// beside the resident weights of block1-conv2
template <int T, int HALO>
struct TCfgH {
  static constexpr int ROWS = T * 128;
  static constexpr int SROWS = ROWS + 2 * HALO;
  static constexpr int PLANE = SROWS * 16;
  static constexpr int SLAB = 2 * (TC_KS / 8) * PLANE;
};

// Descriptors as (lo, hi) words: only the 14-bit start-address field in the low word changes between
// MMAs, so the issuer's address arithmetic is 32-bit and the high words stay loop-invariant.
__device__ __forceinline__ void tc_mma2s(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum) : "memory");
}

// Order in which a group's K slabs are consumed: the one-tap residual slabs are spread evenly between the
// nine-tap main slabs (m0 r0 r1 m1 r2 r3 for 2 + 4), so that every window of (slab buffers - 1)
// consecutive slabs holds a main slab's worth of MMA time for the prefetch of the slab after it.
struct SlabSeq {
  int M, R, rfirst, mi, ri;
  // rfirst (layers with two 3x3 segments): the group starts with a residual slab, whose first MMA initialises
  // ALL output columns (the segments then only accumulate), and the rest alternate r0 m0 r1 m1 r2 m2 r3 --
  // otherwise a residual MMA could hit columns whose segment has not started yet
  __device__ __forceinline__ SlabSeq(int m, int r, int rfirst_ = 0) : M(m), R(r), rfirst(rfirst_), mi(0), ri(0) {}
  // returns true for a residual slab; idx = its index within its kind
  __device__ __forceinline__ bool next(int& idx) {
    const bool res = mi == M || (rfirst ? ri < 1 + (mi * (R - 1)) / M : ri < (mi * R) / M);
    idx = res ? ri++ : mi++;
    return res;
  }
};

// What one K slab of a layer is: where its activations come from, how many taps it carries, which output
// columns its MMAs produce (a layer may merge two 3x3 convs of different inputs: columns [0, n0) and
// [n0, cout); the 1x1 residual slabs always cover all cout columns), and where its weight stages start --
// `wstage` in units of this segment's per-CTA stage size, `wbase` in bytes of per-CTA weights before it.
struct SlabInfo {
  const __half* X; int xch, kc0, ntap, n, col0;
  uint32_t sbytes, wbase; int wstage;
};
__device__ __forceinline__ uint32_t seg_stage_bytes(const TCLayer& L, int n) {
  return L.fold ? 4u * (uint32_t)(n + n / 2) * 16u : 4u * (uint32_t)n * 16u;   // [kchunk 4][X rows | S rows][8] halves
}
__device__ __forceinline__ SlabInfo slab_info(const TCLayer& L, bool is_res, int idx) {
  SlabInfo s;
  const int m0 = L.src_ch / TC_KS, m1 = L.src2 ? L.src2_ch / TC_KS : 0;
  const int n1 = L.cout - L.n0;
  const uint32_t sb0 = seg_stage_bytes(L, L.n0), sb1 = L.src2 ? seg_stage_bytes(L, n1) : 0u;
  if (is_res) {
    s.X = L.res; s.xch = L.res_ch; s.kc0 = idx * (TC_KS / 8); s.ntap = 1; s.n = L.cout; s.col0 = 0;
    s.sbytes = seg_stage_bytes(L, L.cout);
    s.wbase = sb0 * (uint32_t)(m0 * L.ntaps) + sb1 * (uint32_t)(m1 * L.ntaps);
    s.wstage = idx;
  } else if (idx < m0) {
    s.X = L.src; s.xch = L.src_ch; s.kc0 = idx * (TC_KS / 8); s.ntap = L.ntaps; s.n = L.n0; s.col0 = 0;
    s.sbytes = sb0; s.wbase = 0; s.wstage = idx * L.ntaps;
  } else {
    const int j = idx - m0;
    s.X = L.src2; s.xch = L.src2_ch; s.kc0 = j * (TC_KS / 8); s.ntap = L.ntaps; s.n = n1; s.col0 = L.n0;
    s.sbytes = sb1; s.wbase = sb0 * (uint32_t)(m0 * L.ntaps); s.wstage = j * L.ntaps;
  }
  return s;
}

template <int T, bool RESW, int HALO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1) k_tc_conv2(const __grid_constant__ TCLayer L) {
  using Cfg = TCfgH<T, HALO>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int nsb = L.nslab_buf;                             // A slab buffers (2..4)
  uint8_t* a_buf = smem;
  uint8_t* w_buf = smem + nsb * Cfg::SLAB;
  float* s_bias = (float*)(w_buf + L.w_bytes);
  TCBarriers* B = (TCBarriers*)(s_bias + 128);
  float* s_hw = (float*)((uint8_t*)B + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cout = L.cout;
  const bool fold = L.fold != 0;
  const int cpt = fold ? 2 * cout : cout;                  // TMEM columns per M tile
  const int nbuf = (T * cpt <= 256) ? 2 : 1;
  const int main_slabs = L.src_ch / TC_KS + (L.src2 ? L.src2_ch / TC_KS : 0), res_slabs = L.res ? L.res_ch / TC_KS : 0;
  const int nslabs = main_slabs + res_slabs;
  const int lead = L.src2 ? 1 : 0;                         // SlabSeq::rfirst
  // pair-groups: the pair handles groups 2 pg (leader) and 2 pg + 1 (peer); both CTAs run the
  // same number of iterations (rows past nrows are padding and masked in the epilogue)
  const int npairs = (L.ngroups + 1) / 2;
  const int g_first = 2 * (int)(blockIdx.x >> 1) + (int)rank, g_end = 2 * npairs, g_step = (int)gridDim.x;

  pdl_launch_dependents();
  kt_begin(L.kt);
  if (L.clk && blockIdx.x == 0 && threadIdx.x == 0) { L.clk[0] = (unsigned long long)clock64(); L.clk[1] = kt_now(); }
  if (threadIdx.x < cout) s_bias[threadIdx.x] = L.bias[threadIdx.x] * ACT_SCALE;
  if (L.head_ch) {
    for (int i = threadIdx.x; i < 32 * L.head_ch; i += TC2_THREADS) s_hw[i] = L.head_w[i];
    if (threadIdx.x < L.head_ch) s_hw[32 * 16 + threadIdx.x] = L.head_b[threadIdx.x] * ACT_SCALE;
  }
  if (warp == 0 && lane == 0) {
    const uint32_t full_count = rank == 0 ? 2u : 1u;       // leader: own producer + the peer's relay
    for (int i = 0; i < 4; ++i) { mbar_init(&B->a_full[i], full_count); mbar_init(&B->a_empty[i], 2); }
    for (int i = 0; i < TC2_WSTAGES; ++i) { mbar_init(&B->w_full[i], full_count); mbar_init(&B->w_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&B->t_full[i], 2); mbar_init(&B->t_empty[i], 2 * TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // the peer's barriers are initialised too
  tc_fence_after();
  const uint32_t tmem = B->tmem_base;

  if (warp == 0) {
    // ===================== A producer (each CTA loads its own slabs) =====================
    int ab = 0, aph = 0, dn = 0;
    pdl_wait();                                              // the layer(s) that wrote src / res are complete
    for (int g = g_first; g < g_end; g += g_step) {
      const long long r0 = L.row0 + (long long)(L.reverse ? g_end - 1 - g : g) * Cfg::ROWS - HALO;
      SlabSeq seq(main_slabs, res_slabs, lead);
      for (int s = 0; s < nslabs; ++s) {
        int sidx;
        const bool is_res = seq.next(sidx);
        const SlabInfo si = slab_info(L, is_res, sidx);
        const __half* X = si.X;
        const int xch = si.xch, kc0 = si.kc0;
        mbar_wait(&B->a_empty[ab], aph ^ 1);
        if (lane == 0) dbg_mark(L.dbg, 0, dn);
        if (elect_one()) {
          mbar_expect_tx(&B->a_full[ab], Cfg::SLAB);
          uint8_t* dst = a_buf + ab * Cfg::SLAB;
#pragma unroll
          for (int hl = 0; hl < 2; ++hl)
#pragma unroll
            for (int j = 0; j < TC_KS / 8; ++j) {
              const __half* p = X + ((long long)(hl * (xch / 8) + kc0 + j) * L.plane_rows + r0) * 8;
              bulk_g2s(dst + (hl * (TC_KS / 8) + j) * Cfg::PLANE, p, Cfg::PLANE, &B->a_full[ab]);
            }
        }
        __syncwarp();
        if (++ab == nsb) { ab = 0; aph ^= 1; }
      }
    }
  } else if (warp == TC_W_WARP) {
    // ===================== W producer (each CTA loads its half of every weight stage) =====================
    // packed weights: per segment, per stage, the shares of the two CTAs back to back
    const uint8_t* wsrc0 = (const uint8_t*)L.wpk;
    if (RESW) {
      if (elect_one()) {
        uint32_t total = 0;
        for (int kind = 0; kind < 2; ++kind)
          for (int sidx = 0; sidx < (kind ? res_slabs : main_slabs); ++sidx) {
            const SlabInfo si = slab_info(L, kind != 0, sidx);
            total += si.sbytes * (uint32_t)si.ntap;
          }
        mbar_expect_tx(&B->w_full[0], total);
        for (int kind = 0; kind < 2; ++kind)
          for (int sidx = 0; sidx < (kind ? res_slabs : main_slabs); ++sidx) {
            const SlabInfo si = slab_info(L, kind != 0, sidx);
            for (int t = 0; t < si.ntap; ++t) {
              const uint32_t off = si.wbase + (uint32_t)(si.wstage + t) * si.sbytes;      // per-CTA byte offset
              bulk_g2s(w_buf + off, wsrc0 + 2u * (size_t)off + (size_t)rank * si.sbytes, si.sbytes, &B->w_full[0]);
            }
          }
      }
      __syncwarp();
    } else {
      int ws = 0, wph = 0;
      for (int g = g_first; g < g_end; g += g_step) {
        SlabSeq seq(main_slabs, res_slabs, lead);
        for (int s = 0; s < nslabs; ++s) {
          int sidx;
          const bool is_res = seq.next(sidx);
          const SlabInfo si = slab_info(L, is_res, sidx);
          for (int t = 0; t < si.ntap; ++t) {
            const uint32_t off = si.wbase + (uint32_t)(si.wstage + t) * si.sbytes;
            mbar_wait(&B->w_empty[ws], wph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&B->w_full[ws], si.sbytes);
              bulk_g2s(w_buf + ws * TC2_WSTAGE_MAX, wsrc0 + 2u * (size_t)off + (size_t)rank * si.sbytes, si.sbytes, &B->w_full[ws]);
            }
            __syncwarp();
            if (++ws == TC2_WSTAGES) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer relay: "my slab / stage landed" -> the leader's full barriers =====================
    int ab = 0, aph = 0, ws = 0, wph = 0;
    if (RESW) {
      mbar_wait(&B->w_full[0], 0);
      if (lane == 0) mbar_arrive_remote(&B->w_full[0], 0);
      __syncwarp();
    }
    for (int g = g_first; g < g_end; g += g_step) {
      SlabSeq seq(main_slabs, res_slabs, lead);
      for (int s = 0; s < nslabs; ++s) {
        int sidx;
        const int ntap = seq.next(sidx) ? 1 : L.ntaps;
        mbar_wait(&B->a_full[ab], aph);
        if (lane == 0) mbar_arrive_remote(&B->a_full[ab], 0);
        __syncwarp();
        if (++ab == nsb) { ab = 0; aph ^= 1; }
        if (!RESW) {
          for (int t = 0; t < ntap; ++t) {
            mbar_wait(&B->w_full[ws], wph);
            if (lane == 0) mbar_arrive_remote(&B->w_full[ws], 0);
            __syncwarp();
            if (++ws == TC2_WSTAGES) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if ((warp == 1 || warp == TC2_MMA_WARP1) && rank == 0) {
    // ===================== MMA issuers (leader): M = 256 over both CTAs, T/2 tiles each =====================
    const int m0 = (warp == 1) ? 0 : T / 2;
    unsigned long long* dbg = (warp == 1) ? L.dbg : nullptr;
    // descriptor low-word deltas (the start-address field counts 16-byte units)
    constexpr uint32_t A_TILE = 128u * 16u / 16u;
    constexpr uint32_t A_K16 = 2u * Cfg::PLANE / 16u;
    constexpr uint32_t A_LO = (TC_KS / 8) * Cfg::PLANE / 16u;
    const uint64_t ad64 = smem_desc(smem_u32(a_buf) + (uint32_t)HALO * 16u, Cfg::PLANE, 128);
    const uint32_t a_hi = (uint32_t)(ad64 >> 32), b_hi = (uint32_t)(smem_desc(0, 0, 128) >> 32);
    const uint32_t ad_base = (uint32_t)ad64 + (uint32_t)m0 * A_TILE;
    const uint32_t w_addr16 = (smem_u32(w_buf) & 0x3FFFFu) >> 4;
    int sh[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) sh[t] = L.shifts[t];
    int ab = 0, aph = 0, ws = 0, wph = 0, tb = 0, tph = 0, dn = 0;
    if (RESW) { mbar_wait_cluster(&B->w_full[0], 0); tc_fence_after(); }
    for (int g = g_first; g < g_end; g += g_step) {
      if (lane == 0) dbg_mark(dbg, 1, dn);
      mbar_wait_cluster(&B->t_empty[tb], tph ^ 1);
      tc_fence_after();
      if (lane == 0) dbg_mark(dbg, 1, dn);
      const uint32_t d0 = tmem + (uint32_t)(tb * T * cpt + m0 * cpt);
      SlabSeq seq(main_slabs, res_slabs, lead);
      for (int s = 0; s < nslabs; ++s) {
        int sidx;
        const bool is_res = seq.next(sidx);
        const SlabInfo si = slab_info(L, is_res, sidx);
        const int ntap = si.ntap;
        // this slab's B operand: stage layout [kchunk 4][X rows | S rows][8]; fold: X = n rows (w_hi in the
        // leader, w_lo in the peer), S = n/2 rows of w_hi; else X = n/2 rows of w_hi, S = n/2 rows of w_lo
        const uint32_t rows = fold ? (uint32_t)(si.n + si.n / 2) : (uint32_t)si.n;
        const uint32_t w_k16 = 2u * rows, w_s16 = fold ? (uint32_t)si.n : (uint32_t)(si.n / 2);
        const uint32_t idesc = instr_desc(256, si.n), idesc2 = instr_desc(256, 2 * si.n);
        const uint32_t w_step = RESW ? si.sbytes / 16u : (uint32_t)(TC2_WSTAGE_MAX / 16);
        const uint32_t bd_lbo = rows << 16;                  // K-adjacent core matrices are rows * 16 bytes apart
        uint32_t bd = (w_addr16 + (si.wbase + (uint32_t)si.wstage * si.sbytes) / 16u) | bd_lbo;   // RESW: first resident stage
        const uint32_t dcol = d0 + (uint32_t)si.col0;
        mbar_wait_cluster(&B->a_full[ab], aph);
        tc_fence_after();
        if (lane == 0) dbg_mark(dbg, 1, dn);
        const uint32_t ad_slab = ad_base + (uint32_t)(ab * (Cfg::SLAB / 16));
        // one (slab, tap) stage: the MMAs of this issuer's tiles
        auto issue_stage = [&](const uint32_t ad0, const uint32_t bd0, const uint32_t first) {
          if (fold) {
#pragma unroll
            for (int m = 0; m < T / 2; ++m) {
#pragma unroll
              for (int k = 0; k < TC_KS / 16; ++k)
                tc_mma2s(dcol + (uint32_t)(m * cpt), ad0 + m * A_TILE + k * A_K16, a_hi, bd0 + k * w_k16, b_hi, idesc2, (first | k) != 0);
#pragma unroll
              for (int k = 0; k < TC_KS / 16; ++k)
                tc_mma2s(dcol + (uint32_t)(m * cpt), ad0 + m * A_TILE + k * A_K16 + A_LO, a_hi, bd0 + k * w_k16 + w_s16, b_hi, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int m = 0; m < T / 2; ++m) {
#pragma unroll
              for (int pass = 0; pass < 3; ++pass) {   // hi*hi, lo*hi, hi*lo
#pragma unroll
                for (int k = 0; k < TC_KS / 16; ++k)
                  tc_mma2s(dcol + (uint32_t)(m * cpt), ad0 + m * A_TILE + k * A_K16 + (pass == 1 ? A_LO : 0), a_hi,
                           bd0 + k * w_k16 + (pass == 2 ? w_s16 : 0), b_hi, idesc, (first | pass | k) != 0);
              }
            }
          }
        };
        // the group's first MMA initialises the accumulator columns: the first stage of the (single) 3x3
        // segment, or -- two segments -- the first residual slab, which spans all columns
        const uint32_t first0 = lead ? (is_res ? (uint32_t)sidx : 1u) : ((uint32_t)si.wstage | (is_res ? 1u : 0u));
        if (RESW) {
          // resident weights: nothing to wait for between the taps -- the whole slab is issued from one
          // elected region (no per-tap elect / branch / warp sync), one commit at its end
          if (elect_one()) {
            if (is_res) {
              issue_stage(ad_slab, bd, first0);
            } else {
#pragma unroll
              for (int t = 0; t < 9; ++t) issue_stage(ad_slab + (uint32_t)sh[t], bd + (uint32_t)t * w_step, first0 | (uint32_t)t);
            }
            tc_commit2(&B->a_empty[ab]);
            if (s == nslabs - 1) tc_commit2(&B->t_full[tb]);
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (t < ntap) {
              mbar_wait_cluster(&B->w_full[ws], wph);
              tc_fence_after();
              const uint32_t ad0 = ad_slab + (uint32_t)(is_res ? 0 : sh[t]);
              const uint32_t bd0 = (w_addr16 + (uint32_t)ws * w_step) | bd_lbo;
              const bool last_tap = t == ntap - 1;
              if (elect_one()) {
                issue_stage(ad0, bd0, first0 | (uint32_t)t);
                tc_commit2(&B->w_empty[ws]);
                if (last_tap) tc_commit2(&B->a_empty[ab]);
                if (last_tap && s == nslabs - 1) tc_commit2(&B->t_full[tb]);
              }
              __syncwarp();
              if (++ws == TC2_WSTAGES) { ws = 0; wph ^= 1; }
            }
          }
        }
        if (++ab == nsb) { ab = 0; aph ^= 1; }
      }
      if (lane == 0) dbg_mark(dbg, 1, dn);
      if (++tb == nbuf) { tb = 0; tph ^= 1; }
    }
  } else if (warp >= 2 && warp < 2 + TC_EPI_WARPS) {
    if (rank == 0) tc_epilogue<T, false>(L, B, tmem, s_bias, s_hw, warp, lane, cpt, nbuf, g_first, g_end, g_step);
    else tc_epilogue<T, true>(L, B, tmem, s_bias, s_hw, warp, lane, cpt, nbuf, g_first, g_end, g_step);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                      // nobody leaves while the peer may still touch its barriers / smem
  kt_end(L.kt);
  if (L.clk && blockIdx.x == 0 && threadIdx.x == 0) { L.clk[2] = (unsigned long long)clock64(); L.clk[3] = kt_now(); }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// ------------------------------------------------------------------ the chunk-major megakernel (A5_TC_MEGA)
// All block-conv layers of a forward in ONE persistent launch, depth first.  Every CTA owns a fixed range of whole
// boards and walks it in chunks of CB boards (gpc groups of 256 positions); for each chunk it runs layer 0, 1, ...,
// nl - 1 before it moves to the next chunk.  Boards are independent and the position space puts zero rows between
// them, so everything a CTA reads in layer l + 1 was written by ITSELF in layer l: there is no dependency between
// CTAs, no grid barrier and no launch boundary between the layers -- only program order inside the CTA (the
// epilogue publishes the units it has stored, the activation producer waits for the unit that covers its slab and
// its halo).  A chunk of 7 boards x 148 CTAs keeps a layer's output (<= 77 MB) in the 126 MB L2 until the next layer
// reads it, the slab / weight / TMEM pipelines never drain between layers, and the last-wave quantisation is per
// forward (28 boards = 31.5 tiles per CTA) instead of per layer.  Weights always stream through a ring of three-tap
// stages.
constexpr int TCM_MAXL = 8;
struct TCMega {
  TCLayer L[TCM_MAXL];
  int dep_src[TCM_MAXL], dep_src2[TCM_MAXL], dep_res[TCM_MAXL];   // layer of this launch that writes the tensor, or -1
  int nl;
  int B, n_hi;             // CTAs [0, n_hi) own B + 1 boards, the others B
  int CB, nchunks, gpc;    // boards per chunk; chunks per CTA and groups per chunk (the same for every CTA)
  unsigned long long* kt;
  unsigned long long* dbg; // tooling: [nchunks * nl + 1][2] = {%globaltimer, clock64} when CTA 0's issuer starts a (chunk, layer)
};
constexpr int TCM_MISC = TCM_MAXL * 128 * 4 + 256 + 2 * (32 * 16 + 16) * 4 + 128 + 64;
// weight ring: a stage holds up to three (slab, tap) blocks -- one full / empty handshake and one commit per three
// taps keeps the two issuing threads ahead of the tensor pipe also for the N <= 64 layers (a handshake per tap does
// not: ~200 cycles of waits, elect and commits per 4 MMAs)
constexpr int TCM_TPS = 3;
constexpr int TCM_WST = TCM_TPS * TC2_WSTAGE_MAX;         // 24 KB
constexpr int TCM_NWS = 4;
constexpr int TCM_WBYTES = TCM_NWS * TCM_WST;

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// rows of chunk k of this CTA: first row index (from row0) and the first row past its boards
__device__ __forceinline__ void tcm_chunk(const TCMega& P, int cta, int k, uint32_t& q0, uint32_t& qlim) {
  const int per_board = P.L[0].per_board;
  const int first = cta * P.B + min(cta, P.n_hi), nb = P.B + (cta < P.n_hi ? 1 : 0);
  const int b0 = min(k * P.CB, nb), b1 = min((k + 1) * P.CB, nb);
  const uint32_t nrows = (uint32_t)P.L[0].nrows;
  q0 = min((uint32_t)(first + b0) * (uint32_t)per_board, nrows);
  qlim = min((uint32_t)(first + b1) * (uint32_t)per_board, nrows);
}

template <int HALO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1) k_tc_mega(const __grid_constant__ TCMega P) {
  constexpr int T = 2;
  using Cfg = TCfgH<T, HALO>;
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NSB = (TC2_SMEM_LIMIT - TCM_MISC - TCM_WBYTES) / Cfg::SLAB > 4 ? 4 : (TC2_SMEM_LIMIT - TCM_MISC - TCM_WBYTES) / Cfg::SLAB;
  uint8_t* a_buf = smem;
  uint8_t* w_buf = smem + NSB * Cfg::SLAB;
  float* s_bias = (float*)(w_buf + TCM_WBYTES);                              // [nl][128]
  TCBarriers* B = (TCBarriers*)(s_bias + TCM_MAXL * 128);
  float* s_hw = (float*)((uint8_t*)B + 256);                                 // [2][32 * 16 + 16]
  volatile int* s_done = (volatile int*)(s_hw + 2 * (32 * 16 + 16));         // [16] units stored, per epilogue warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cta = (int)blockIdx.x;
  const int nl = P.nl, gpc = P.gpc, nchunks = P.nchunks;

  pdl_launch_dependents();
  kt_begin(P.kt);
  for (int i = threadIdx.x; i < nl * 128; i += TC2_THREADS) {
    const int li = i >> 7, c = i & 127;
    s_bias[i] = c < P.L[li].cout ? P.L[li].bias[c] * ACT_SCALE : 0.0f;
  }
  {
    int hs = 0;
    for (int li = 0; li < nl; ++li)
      if (P.L[li].head_ch) {
        float* hw = s_hw + hs * (32 * 16 + 16);
        for (int i = threadIdx.x; i < 32 * P.L[li].head_ch; i += TC2_THREADS) hw[i] = P.L[li].head_w[i];
        if (threadIdx.x < P.L[li].head_ch) hw[32 * 16 + threadIdx.x] = P.L[li].head_b[threadIdx.x] * ACT_SCALE;
        ++hs;
      }
  }
  if (threadIdx.x < 16) s_done[threadIdx.x] = 0;
  if (warp == 0 && lane == 0) {
    const uint32_t full_count = rank == 0 ? 2u : 1u;       // leader: own producer + the peer's relay
    for (int i = 0; i < 4; ++i) { mbar_init(&B->a_full[i], full_count); mbar_init(&B->a_empty[i], 2); }
    for (int i = 0; i < TCM_NWS; ++i) { mbar_init(&B->w_full[i], full_count); mbar_init(&B->w_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&B->t_full[i], 2); mbar_init(&B->t_empty[i], 2 * TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = B->tmem_base;

  if (warp == 0) {
    // ===================== A producer =====================
    int ab = 0, aph = 0, seen = 0;
    pdl_wait();                                              // conv1 (the layer before this launch) is complete
    for (int k = 0; k < nchunks; ++k) {
      uint32_t qc, qlim;
      tcm_chunk(P, cta, k, qc, qlim);
      for (int li = 0; li < nl; ++li) {
        const TCLayer& L = P.L[li];
        const int main_slabs = L.src_ch / TC_KS + (L.src2 ? L.src2_ch / TC_KS : 0), res_slabs = L.res ? L.res_ch / TC_KS : 0;
        const int m0s = L.src_ch / TC_KS;
        for (int g = 0; g < gpc; ++g) {
          const long long r0 = L.row0 + (long long)qc + (long long)g * Cfg::ROWS - HALO;
          SlabSeq seq(main_slabs, res_slabs, L.src2 ? 1 : 0);
          for (int s = 0; s < main_slabs + res_slabs; ++s) {
            int sidx;
            const bool is_res = seq.next(sidx);
            const SlabInfo si = slab_info(L, is_res, sidx);
            // the unit (of this CTA, this chunk) that wrote the last rows this slab reads: group g + 1 holds the halo
            const int dl = is_res ? P.dep_res[li] : (sidx < m0s ? P.dep_src[li] : P.dep_src2[li]);
            if (dl >= 0) {
              const int need = (k * nl + dl) * gpc + min(g + 1, gpc - 1) + 1;
              for (uint32_t spin = 0; seen < need; ++spin) {
                int v = lane < TC_EPI_WARPS ? s_done[lane] : 0x7fffffff;
                seen = __reduce_min_sync(FULL, v);
                if (spin > (1u << 26)) __trap();             // a lost unit must fail loudly, not hang the GPU
              }
              fence_proxy_async_all();                       // the epilogue's generic-proxy stores before the bulk reads
            }
            mbar_wait(&B->a_empty[ab], aph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&B->a_full[ab], Cfg::SLAB);
              uint8_t* dst = a_buf + ab * Cfg::SLAB;
#pragma unroll
              for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                for (int j = 0; j < TC_KS / 8; ++j) {
                  const __half* p = si.X + ((long long)(hl * (si.xch / 8) + si.kc0 + j) * L.plane_rows + r0) * 8;
                  bulk_g2s(dst + (hl * (TC_KS / 8) + j) * Cfg::PLANE, p, Cfg::PLANE, &B->a_full[ab]);
                }
            }
            __syncwarp();
            if (++ab == NSB) { ab = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == TC_W_WARP) {
    // ===================== W producer: every (slab, tap) stage of every group through the ring =====================
    int ws = 0, wph = 0;
    for (int k = 0; k < nchunks; ++k)
      for (int li = 0; li < nl; ++li) {
        const TCLayer& L = P.L[li];
        const uint8_t* wsrc0 = (const uint8_t*)L.wpk;
        const int main_slabs = L.src_ch / TC_KS + (L.src2 ? L.src2_ch / TC_KS : 0), res_slabs = L.res ? L.res_ch / TC_KS : 0;
        for (int g = 0; g < gpc; ++g) {
          SlabSeq seq(main_slabs, res_slabs, L.src2 ? 1 : 0);
          for (int s = 0; s < main_slabs + res_slabs; ++s) {
            int sidx;
            const bool is_res = seq.next(sidx);
            const SlabInfo si = slab_info(L, is_res, sidx);
            for (int t0 = 0; t0 < si.ntap; t0 += TCM_TPS) {
              const int nt = min(TCM_TPS, si.ntap - t0);
              mbar_wait(&B->w_empty[ws], wph ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&B->w_full[ws], (uint32_t)nt * si.sbytes);
                for (int j = 0; j < nt; ++j) {
                  const uint32_t off = si.wbase + (uint32_t)(si.wstage + t0 + j) * si.sbytes;
                  bulk_g2s(w_buf + ws * TCM_WST + j * si.sbytes, wsrc0 + 2u * (size_t)off + (size_t)rank * si.sbytes, si.sbytes,
                           &B->w_full[ws]);
                }
              }
              __syncwarp();
              if (++ws == TCM_NWS) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer relay =====================
    int ab = 0, aph = 0, ws = 0, wph = 0;
    for (int k = 0; k < nchunks; ++k)
      for (int li = 0; li < nl; ++li) {
        const TCLayer& L = P.L[li];
        const int main_slabs = L.src_ch / TC_KS + (L.src2 ? L.src2_ch / TC_KS : 0), res_slabs = L.res ? L.res_ch / TC_KS : 0;
        for (int g = 0; g < gpc; ++g) {
          SlabSeq seq(main_slabs, res_slabs, L.src2 ? 1 : 0);
          for (int s = 0; s < main_slabs + res_slabs; ++s) {
            int sidx;
            const int ntap = seq.next(sidx) ? 1 : L.ntaps;
            mbar_wait(&B->a_full[ab], aph);
            if (lane == 0) mbar_arrive_remote(&B->a_full[ab], 0);
            __syncwarp();
            if (++ab == NSB) { ab = 0; aph ^= 1; }
            for (int t0 = 0; t0 < ntap; t0 += TCM_TPS) {
              mbar_wait(&B->w_full[ws], wph);
              if (lane == 0) mbar_arrive_remote(&B->w_full[ws], 0);
              __syncwarp();
              if (++ws == TCM_NWS) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
  } else if ((warp == 1 || warp == TC2_MMA_WARP1) && rank == 0) {
    // ===================== MMA issuers (leader) =====================
    const int m0 = (warp == 1) ? 0 : T / 2;
    constexpr uint32_t A_TILE = 128u * 16u / 16u;
    constexpr uint32_t A_K16 = 2u * Cfg::PLANE / 16u;
    constexpr uint32_t A_LO = (TC_KS / 8) * Cfg::PLANE / 16u;
    const uint64_t ad64 = smem_desc(smem_u32(a_buf) + (uint32_t)HALO * 16u, Cfg::PLANE, 128);
    const uint32_t a_hi = (uint32_t)(ad64 >> 32), b_hi = (uint32_t)(smem_desc(0, 0, 128) >> 32);
    const uint32_t ad_base = (uint32_t)ad64 + (uint32_t)m0 * A_TILE;
    const uint32_t w_addr16 = (smem_u32(w_buf) & 0x3FFFFu) >> 4;
    constexpr uint32_t w_step = (uint32_t)(TCM_WST / 16);
    int ab = 0, aph = 0, ws = 0, wph = 0, tb = 0, tph = 0;
    for (int k = 0; k < nchunks; ++k)
      for (int li = 0; li < nl; ++li) {
        const TCLayer& L = P.L[li];
        const bool fold = L.fold != 0;
        const int cpt = fold ? 2 * L.cout : L.cout;
        const int main_slabs = L.src_ch / TC_KS + (L.src2 ? L.src2_ch / TC_KS : 0), res_slabs = L.res ? L.res_ch / TC_KS : 0;
        const int lead = L.src2 ? 1 : 0;
        int sh[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) sh[t] = L.shifts[t];
        if (P.dbg && cta == 0 && warp == 1 && lane == 0) {
          P.dbg[2 * (k * nl + li)] = kt_now();
          P.dbg[2 * (k * nl + li) + 1] = (unsigned long long)clock64();
        }
        for (int g = 0; g < gpc; ++g) {
          mbar_wait_cluster(&B->t_empty[tb], tph ^ 1);
          tc_fence_after();
          const uint32_t d0 = tmem + (uint32_t)(tb * 256 + m0 * cpt);
          SlabSeq seq(main_slabs, res_slabs, lead);
          for (int s = 0; s < main_slabs + res_slabs; ++s) {
            int sidx;
            const bool is_res = seq.next(sidx);
            const SlabInfo si = slab_info(L, is_res, sidx);
            const int ntap = si.ntap;
            const uint32_t rows = fold ? (uint32_t)(si.n + si.n / 2) : (uint32_t)si.n;
            const uint32_t w_k16 = 2u * rows, w_s16 = fold ? (uint32_t)si.n : (uint32_t)(si.n / 2);
            const uint32_t idesc = instr_desc(256, si.n), idesc2 = instr_desc(256, 2 * si.n);
            const uint32_t bd_lbo = rows << 16;
            const uint32_t dcol = d0 + (uint32_t)si.col0;
            mbar_wait_cluster(&B->a_full[ab], aph);
            tc_fence_after();
            const uint32_t ad_slab = ad_base + (uint32_t)(ab * (Cfg::SLAB / 16));
            const uint32_t first0 = lead ? (is_res ? (uint32_t)sidx : 1u) : ((uint32_t)si.wstage | (is_res ? 1u : 0u));
            const uint32_t sb16 = si.sbytes / 16u;
            const bool last_slab = s == main_slabs + res_slabs - 1;
#pragma unroll
            for (int t0 = 0; t0 < 9; t0 += TCM_TPS) {
              if (t0 < ntap) {
                mbar_wait_cluster(&B->w_full[ws], wph);
                tc_fence_after();
                const uint32_t bst = (w_addr16 + (uint32_t)ws * w_step) | bd_lbo;
                const bool last_stage = t0 + TCM_TPS >= ntap;
                if (elect_one()) {
#pragma unroll
                  for (int j = 0; j < TCM_TPS; ++j) {
                    const int t = t0 + j;
                    if (t < ntap) {
                      const uint32_t ad0 = ad_slab + (uint32_t)(is_res ? 0 : sh[t]);
                      const uint32_t bd0 = bst + (uint32_t)j * sb16;
                      const uint32_t first = first0 | (uint32_t)t;
                      if (fold) {
#pragma unroll
                        for (int kk = 0; kk < TC_KS / 16; ++kk)
                          tc_mma2s(dcol, ad0 + kk * A_K16, a_hi, bd0 + kk * w_k16, b_hi, idesc2, (first | kk) != 0);
#pragma unroll
                        for (int kk = 0; kk < TC_KS / 16; ++kk)
                          tc_mma2s(dcol, ad0 + kk * A_K16 + A_LO, a_hi, bd0 + kk * w_k16 + w_s16, b_hi, idesc, 1u);
                      } else {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                          for (int kk = 0; kk < TC_KS / 16; ++kk)
                            tc_mma2s(dcol, ad0 + kk * A_K16 + (pass == 1 ? A_LO : 0), a_hi,
                                     bd0 + kk * w_k16 + (pass == 2 ? w_s16 : 0), b_hi, idesc, (first | pass | kk) != 0);
                      }
                    }
                  }
                  tc_commit2(&B->w_empty[ws]);
                  if (last_stage) tc_commit2(&B->a_empty[ab]);
                  if (last_stage && last_slab) tc_commit2(&B->t_full[tb]);
                }
                __syncwarp();
                if (++ws == TCM_NWS) { ws = 0; wph ^= 1; }
              }
            }
            if (++ab == NSB) { ab = 0; aph ^= 1; }
          }
          if (++tb == 2) { tb = 0; tph ^= 1; }
        }
      }
  } else if (warp >= 2 && warp < 2 + TC_EPI_WARPS) {
    // ===================== epilogue =====================
    int tb = 0, tph = 0, unit = 0;
    for (int k = 0; k < nchunks; ++k) {
      uint32_t qc, qlim;
      tcm_chunk(P, cta, k, qc, qlim);
      int hs = 0;
      for (int li = 0; li < nl; ++li) {
        const TCLayer& L = P.L[li];
        const int cpt = L.fold ? 2 * L.cout : L.cout;
        const float* hw = s_hw + hs * (32 * 16 + 16);
        if (L.head_ch) ++hs;
        for (int g = 0; g < gpc; ++g) {
          mbar_wait(&B->t_full[tb], tph);
          tc_fence_after();
          tc_epilogue_group<T>(L, tmem + (uint32_t)(tb * 256), s_bias + li * 128, hw, warp, lane, cpt,
                               qc + (uint32_t)(g * Cfg::ROWS), qlim);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank != 0) mbar_arrive_remote(&B->t_empty[tb], 0);
            else mbar_arrive(&B->t_empty[tb]);
          }
          // publish: this warp's stores of the unit are visible to the bulk-copy (async) proxy of this SM
          __threadfence();
          fence_proxy_async_all();
          __syncwarp();
          ++unit;
          if (lane == 0) s_done[warp - 2] = unit;
          if (++tb == 2) { tb = 0; tph ^= 1; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  kt_end(P.kt);
  if (P.dbg && cta == 0 && threadIdx.x == 0) {
    P.dbg[2 * (nchunks * nl)] = kt_now();
    P.dbg[2 * (nchunks * nl) + 1] = (unsigned long long)clock64();
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// One 3x3 / 1x1 segment of a merged layer, non-fold layout: per stage (32-channel slab, tap) the shares of the two
// CTAs, each [kchunk 4][X: n/2 rows of w_hi | S: n/2 rows of w_lo][8].  `wb` (optional) continues `w` along N.
__device__ __forceinline__ __half lo_round(__half l, int drop) {
  if (!drop) return l;
  unsigned short b = *(unsigned short*)&l;
  b = (unsigned short)((b + (1u << (drop - 1))) & ~((1u << drop) - 1u));
  return *(__half*)&b;
}
__global__ void k_tc_pack2_seg(const float* __restrict__ w, const float* __restrict__ wb, int na, int ntaps, int cin, int n,
                               __half* __restrict__ out, int drop) {
  const int nstages = (cin / TC_KS) * ntaps;
  const long long per_cta = (long long)4 * n * 8;          // halfs
  const long long total = (long long)nstages * 2 * per_cta;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int stage = (int)(i / (2 * per_cta));
    const int r = (int)(i % (2 * per_cta));
    const int cta = r / (int)per_cta, r2 = r % (int)per_cta;
    const int j = r2 / (n * 8), row = (r2 / 8) % n, e = r2 % 8;
    const int half_n = n / 2;
    const int lo = row >= half_n;
    const int col = cta * half_n + (lo ? row - half_n : row);
    const int slab = stage / ntaps, tap = stage % ntaps;
    const size_t kk = (size_t)tap * cin + slab * TC_KS + j * 8 + e;
    float x = !wb ? w[kk * n + col] : (col < na ? w[kk * na + col] : wb[kk * (n - na) + (col - na)]);
    x *= W_SCALE;
    const __half h = __float2half_rn(x);
    out[i] = lo ? lo_round(__float2half_rn(x - __half2float(h)), drop) : h;
  }
}

// Weights for the CTA-pair kernel: [stage][cta 2][kchunk 4][X rows | S rows][8] (see k_tc_conv2).
// `wb` (optional): a second kernel over the same input whose output channels follow the `na` of `w`
// (two convs of one tensor merged into one layer).
__global__ void k_tc_pack2(const float* __restrict__ w, const float* __restrict__ wb, int na, const float* __restrict__ wres,
                           int ntaps, int cin, int rcin, int cout, int fold, __half* __restrict__ out, int drop) {
  const int main_stages = (cin / TC_KS) * ntaps;
  const int nstages = main_stages + (wres ? rcin / TC_KS : 0);
  const int xr = fold ? cout : cout / 2, sr = cout / 2, rows = xr + sr;
  const long long per_cta = (long long)4 * rows * 8;      // halfs
  const long long total = (long long)nstages * 2 * per_cta;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int stage = (int)(i / (2 * per_cta));
    const int r = (int)(i % (2 * per_cta));
    const int cta = r / (int)per_cta, r2 = r % (int)per_cta;
    const int j = r2 / (rows * 8), row = (r2 / 8) % rows, e = r2 % 8;
    // which weight row / half this smem row holds
    int n, lo;
    if (fold) {
      if (row < xr) { n = row; lo = cta; }                              // X: leader w_hi, peer w_lo
      else { n = cta * sr + (row - xr); lo = 0; }                        // S: halves of w_hi
    } else {
      if (row < xr) { n = cta * xr + row; lo = 0; }                      // X: this CTA's half of w_hi
      else { n = cta * sr + (row - xr); lo = 1; }                        // S: this CTA's half of w_lo
    }
    const int kin = j * 8 + e;
    float x;
    if (stage < main_stages) {
      const int slab = stage / ntaps, tap = stage % ntaps;
      const size_t kk = (size_t)tap * cin + slab * TC_KS + kin;
      x = !wb ? w[kk * cout + n] : (n < na ? w[kk * na + n] : wb[kk * (cout - na) + (n - na)]);
    } else {
      const int slab = stage - main_stages;
      x = wres[(size_t)(slab * TC_KS + kin) * cout + n];
    }
    x *= W_SCALE;
    const __half h = __float2half_rn(x);
    out[i] = lo ? lo_round(__float2half_rn(x - __half2float(h)), drop) : h;
  }
}

// ------------------------------------------------------------------ support kernels
// ------------------------------------------------------------------ conv1 on the tensor cores
// conv1 (5x5, 3 -> 32, ELU; network.py:63) as a GEMM with an explicit A tile: the inputs are exactly
// {0, 1} (utils.py:256-272), so A = im2col(planes) is exact in fp16 and only the weights are split:
//   D[128 positions x 64] = A[128 x 128] * [w_hi | w_lo][128 x 64]     (8 MMAs of K = 16, N = 64)
// with k = (ky*3 + plane)*8 + kx: one 8-element K chunk (16 bytes) per (kernel row, plane), kx = 5..7 and
// the 16th chunk zero, so a chunk of the A tile is a 32-entry table lookup on the 5-bit window pattern
// of that board row.  k_c1_bits turns the int8 planes into bitboards; builder warps extract the 15
// patterns of their position and write the tile as fp16 {0, 1.0} straight in the tcgen05 K-major
// layout; one warp issues the MMAs; four warps run the epilogue (hi + lo, bias, ELU, split, store).
constexpr int C1M_BUILD_WARPS = 4;
constexpr int C1M_THREADS = 32 * (C1M_BUILD_WARPS + 1 + 4);
constexpr int C1M_KCH = 16;                                   // K = 128 as 16 chunks of 8 (15 used)
constexpr int C1M_ATILE = C1M_KCH * 128 * 16;                 // 32 KB
constexpr int C1M_WBYTES = C1M_KCH * 64 * 16;                 // 16 KB
constexpr int C1M_BW = 24;                                    // words per board: 8 per input plane
#ifndef C1M_NA
#define C1M_NA 2                                              // A tile buffers per CTA
#endif
#ifndef C1M_CTAS
#define C1M_CTAS 2                                            // resident CTAs per SM the grid is sized for
#endif
constexpr int C1M_SMEM = C1M_NA * C1M_ATILE + C1M_WBYTES + 32 * 16 + 128 + 128;

struct __align__(8) C1MBarriers {
  uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2];
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// int8 planes [n][3][C] -> bitboards [n][plane 3][8 words]: one warp per board.
__global__ void __launch_bounds__(128) k_c1_bits(const int8_t* __restrict__ planes, int n, int C, uint32_t* __restrict__ bits,
                                                 unsigned long long* kt) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();                                              // the tree pass that wrote the planes is complete
  kt_begin(kt);
  if (kt && b >= n) { __syncthreads(); return; }           // (timing on: everybody meets at the barrier below)
  if (b >= n) return;
  const int8_t* p = planes + (size_t)b * 3 * C;
  uint32_t mine = 0;
#pragma unroll
  for (int pl = 0; pl < 3; ++pl)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = k * 32 + lane;
      const unsigned m = __ballot_sync(FULL, c < C && p[pl * C + c] != 0);
      if (lane == pl * 8 + k) mine = m;
    }
  if (lane < C1M_BW) bits[(size_t)b * C1M_BW + lane] = mine;
  if (kt) { __syncthreads(); kt_end(kt); }
}

__global__ void __launch_bounds__(C1M_THREADS) k_tc_conv1m(const uint32_t* __restrict__ bits, const __half* __restrict__ wpk,
                                                          const float* __restrict__ bias, __half* __restrict__ out,
                                                          long long plane_rows, int S, int pitch, int per_board, int guard,
                                                          int n, int ntiles, uint32_t lo_add, uint32_t lo_mask,
                                                          unsigned long long* kt) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_launch_dependents();
  kt_begin(kt);
  uint8_t* a_buf = smem;                                       // 2 A tiles [kchunk 10][row 128][8]
  uint8_t* w_buf = smem + C1M_NA * C1M_ATILE;                  // [kchunk 16][hi 32 | lo 32][8]
  uint4* s_lut = (uint4*)(w_buf + C1M_WBYTES);                 // window pattern -> 8 halves (5 used)
  float* s_bias = (float*)(s_lut + 32);
  C1MBarriers* B = (C1MBarriers*)(s_bias + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const uint32_t nrows = (uint32_t)n * (uint32_t)per_board;

  for (int i = tid; i < C1M_WBYTES / 16; i += C1M_THREADS) ((uint4*)w_buf)[i] = ((const uint4*)wpk)[i];
  if (tid < 32) {
    const uint32_t b = (uint32_t)tid;
    s_lut[tid] = make_uint4(((b & 1u) ? 0x3C00u : 0u) | ((b & 2u) ? 0x3C000000u : 0u),
                            ((b & 4u) ? 0x3C00u : 0u) | ((b & 8u) ? 0x3C000000u : 0u), (b & 16u) ? 0x3C00u : 0u, 0u);
  }
  for (int i = tid; i < C1M_NA * 128; i += C1M_THREADS)        // the 16th K chunk of every tile buffer stays zero
    *(uint4*)(a_buf + (i >> 7) * C1M_ATILE + (15 * 128 + (i & 127)) * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 32) s_bias[tid] = bias[tid];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&B->a_full[i], C1M_BUILD_WARPS); mbar_init(&B->a_empty[i], 1);
      mbar_init(&B->t_full[i], 1); mbar_init(&B->t_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == C1M_BUILD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&B->tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async_smem();                                    // the weights were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = B->tmem_base;

  if (warp < C1M_BUILD_WARPS) {
    // ===================== builders: bitboards -> 15 window patterns -> A row =====================
    // Software-pipelined: the 30 bitboard words of the NEXT tile's position are in flight (one L2 round
    // trip, row index clamped, result masked later) while the current tile is expanded and stored.
    const int t = tid;                                         // row of the tile
    const uint32_t rowbits = (1u << S) - 1u;
    struct Pos { int rr, cc; bool real; uint32_t lo[5][3], hi[5][3]; };
    auto fetch = [&](int tile, Pos& P) {
      const uint32_t q = (uint32_t)tile * 128u + (uint32_t)t;
      const uint32_t board = q / (uint32_t)per_board, within = q - board * (uint32_t)per_board;
      P.rr = (int)(within / (uint32_t)pitch);
      P.cc = (int)within - P.rr * pitch;
      P.real = tile < ntiles && q < nrows && P.rr < S && P.cc < S;
      const uint32_t* bw = bits + (size_t)(P.real ? board : 0u) * C1M_BW;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        const int r = min(max(P.rr + ky - 2, 0), S - 1);
        const int wd = (r * S) >> 5;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          P.lo[ky][p] = __ldg(bw + p * 8 + wd);
          P.hi[ky][p] = __ldg(bw + p * 8 + min(wd + 1, 7));
        }
      }
    };
    int ab = 0, aph = 0;
    Pos cur, nxt;
    pdl_wait();                                                // k_c1_bits is complete (weights, LUT, TMEM were set up meanwhile)
    fetch(blockIdx.x, cur);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      fetch(tile + gridDim.x, nxt);
      uint32_t pat[3][5];
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        const int r = cur.rr + ky - 2;
        const bool ok = cur.real && r >= 0 && r < S;
        const int sh = (min(max(r, 0), S - 1) * S) & 31;        // row r = bits [r S, r S + S) of the 8-word bitboard
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint32_t row = __funnelshift_r(cur.lo[ky][p], cur.hi[ky][p], sh) & rowbits;
          pat[p][ky] = ok ? ((row << 2) >> cur.cc) & 31u : 0u;  // columns cc-2 .. cc+2
        }
      }
      mbar_wait(&B->a_empty[ab], aph ^ 1);
      uint8_t* arow = a_buf + ab * C1M_ATILE + t * 16;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky)
#pragma unroll
        for (int p = 0; p < 3; ++p) *(uint4*)(arow + (ky * 3 + p) * 128 * 16) = s_lut[pat[p][ky]];
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->a_full[ab]);
      if (++ab == C1M_NA) { ab = 0; aph ^= 1; }
      cur = nxt;
    }
  } else if (warp == C1M_BUILD_WARPS) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = instr_desc(128, 64);
    const uint64_t ad_base = smem_desc(smem_u32(a_buf), 128 * 16, 128);
    const uint64_t bd_base = smem_desc(smem_u32(w_buf), 64 * 16, 128);
    int ab = 0, aph = 0, tb = 0, tph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(&B->t_empty[tb], tph ^ 1);
      mbar_wait(&B->a_full[ab], aph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < C1M_KCH / 2; ++k)
          tc_mma(tmem + (uint32_t)(tb * 64), ad_base + (uint64_t)(ab * (C1M_ATILE / 16) + k * (2 * 128 * 16 / 16)),
                 bd_base + (uint64_t)(k * (2 * 64 * 16 / 16)), idesc, k != 0);
        tc_commit(&B->a_empty[ab]);
        tc_commit(&B->t_full[tb]);
      }
      __syncwarp();
      if (++ab == C1M_NA) { ab = 0; aph ^= 1; }
      if (++tb == 2) { tb = 0; tph ^= 1; }
    }
  } else {
    // ===================== epilogue: hi + lo, bias, ELU, hi/lo split, store =====================
    const int quad = warp & 3;
    int tb = 0, tph = 0;
    constexpr float K_ACC = 1.0f / W_SCALE;
    constexpr float K_L2E = 1.4426950408889634f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const uint32_t q = (uint32_t)tile * 128u + (uint32_t)(quad * 32 + lane);
      const uint32_t within = q % (uint32_t)per_board;
      const int rr = (int)(within / (uint32_t)pitch), cc = (int)within - rr * pitch;
      const bool real = q < nrows && rr < S && cc < S;
      const long long row = guard + (long long)q;
      mbar_wait(&B->t_full[tb], tph);
      tc_fence_after();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t v[16], v2[16];
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(tb * 64 + h2 * 16);
        tc_ld16(taddr, v);
        tc_ld16(taddr + 32u, v2);
        tc_ld_wait();
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = kc * 8 + e * 2 + h;
              const float a = fmaf(__uint_as_float(v[c]) + __uint_as_float(v2[c]), K_ACC, s_bias[h2 * 16 + c]);
              x[h] = (a > 0.0f ? a : ex2_approx(a * K_L2E) - 1.0f) * ACT_SCALE;
            }
            const __half2 hh = __floats2half2_rn(x[0], x[1]);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(x[0] - hf.x, x[1] - hf.y);
            hi[e] = real ? *(const uint32_t*)&hh : 0u;
            lo[e] = real ? ((*(const uint32_t*)&ll + lo_add) & lo_mask) : 0u;
          }
          const long long chunk = h2 * 2 + kc;
          *(uint4*)(out + (chunk * plane_rows + row) * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *(uint4*)(out + ((4 + chunk) * plane_rows + row) * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->t_empty[tb]);
      if (++tb == 2) { tb = 0; tph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  kt_end(kt);
  if (warp == C1M_BUILD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
  }
}

// conv1 weights [75][ldw] f32 (row (ky*5 + kx)*3 + plane) -> [kchunk 16 = ky*3 + plane][w_hi ch 0..31 | w_lo ch 0..31][kx 8]
// fp16, scaled by 2^10; kx >= 5 and chunk 15 are zero
__global__ void k_c1m_pack(const float* __restrict__ w, int ldw, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C1M_KCH * 64 * 8) return;
  const int j = i / (64 * 8), r = (i / 8) % 64, e = i % 8;
  const int ky = j / 3, p = j - ky * 3, ch = r & 31;
  const float x = (j < 15 && e < 5) ? w[(size_t)((ky * 5 + e) * 3 + p) * ldw + ch] * W_SCALE : 0.0f;
  const __half h = __float2half_rn(x);
  out[i] = (r >> 5) ? __float2half_rn(x - __half2float(h)) : h;
}

// TC activation (hi/lo fp16 planes) -> fp32 [row][ch]   (debug / parity tooling)
__global__ void k_tc_unpack(const __half* __restrict__ x, int ch, long long plane_rows, long long nrows,
                            float* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nrows * ch) return;
  long long row = i / ch;
  int c = (int)(i % ch);
  float hi = __half2float(x[(((long long)(c >> 3)) * plane_rows + row) * 8 + (c & 7)]);
  float lo = __half2float(x[(((long long)((ch >> 3) + (c >> 3))) * plane_rows + row) * 8 + (c & 7)]);
  out[i] = (hi + lo) * (1.0f / ACT_SCALE);
}

}  // namespace a5

// ------------------------------------------------------------------ in-situ kernel timing (common.cuh)
namespace a5 {
static unsigned long long* g_kt_dev = nullptr;       // [KT_SLOTS][KT_SUB][2] stamps + [KT_SLOTS][3] sums + [1] previous pass end
static bool g_kt_on = false;
unsigned long long* kt_slot(int slot) { return (g_kt_on && g_kt_dev) ? g_kt_dev + (size_t)slot * KT_SUB * 2 : nullptr; }

// one warp: fold the stamps of the pass that just ended into the sums, reset the stamps
__global__ void k_kt_fold(unsigned long long* kt) {
  unsigned long long* acc = kt + (size_t)KT_SLOTS * KT_SUB * 2;
  unsigned long long* prev_end = acc + KT_SLOTS * 3;
  const int lane = threadIdx.x;
  unsigned long long prev = *prev_end;
  // slots in the order their kernels run in a pass
  const int order[KT_SLOTS] = {KT_EC_LOOKUP, KT_C1BITS, KT_CONV1, KT_CONV0, KT_CONV0 + 1, KT_CONV0 + 2, KT_CONV0 + 3, KT_CONV0 + 4,
                               KT_CONV0 + 5, KT_CONV0 + 6, KT_CONV0 + 7, KT_HEADS, KT_EC_COMMIT, KT_STEP, KT_FOLD, 15};
  for (int oi = 0; oi < KT_SLOTS; ++oi) {
    const int s = order[oi];
    unsigned long long* p = kt + (size_t)s * KT_SUB * 2 + 2 * lane;
    unsigned long long st = p[0], en = p[1];
    for (int o = 16; o; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, st, o), b = __shfl_xor_sync(0xffffffffu, en, o);
      st = a < st ? a : st;
      en = b > en ? b : en;
    }
    p[0] = ~0ull; p[1] = 0ull;
    if (s == KT_FOLD) { en = kt_now(); st = prev; }          // the fold itself closes the pass
    if (en == 0ull) continue;                                // slot not used in this pass
    if (lane == 0 && prev != 0ull) {
      acc[s * 3 + 0] += en > prev ? en - prev : 0ull;       // predecessor's end -> my end
      acc[s * 3 + 1] += en - st;                            // my first start -> my end
      acc[s * 3 + 2] += 1ull;
    }
    prev = en > prev ? en : prev;
  }
  if (lane == 0) *prev_end = prev;
}
}  // namespace a5

// ------------------------------------------------------------------ host side
using namespace a5;

static const int kTcActCh[11] = {32, 64, 64, 128, 128, 32, 32, 64, 64, 32, 32};
enum { A32, B1H, B1O, B2H, B2O, B3H, B3O, B4H, B4O, B5H, B5O };
struct TcLayerDef { int src, res_src, out, cin, res_cin, cout; };
static const TcLayerDef kTcLayers[11] = {
    {0, 0, 0, 0, 0, 0},
    {A32, -1, B1H, 32, 0, 64},  {B1H, A32, B1O, 64, 32, 64},
    {B1O, -1, B2H, 64, 0, 128}, {B2H, B1O, B2O, 128, 64, 128},
    {B2O, -1, B3H, 128, 0, 32}, {B3H, B2O, B3O, 32, 128, 32},
    {B2O, -1, B4H, 128, 0, 64}, {B4H, B2O, B4O, 64, 128, 64},
    {B4O, -1, B5H, 64, 0, 32},  {B5H, B4O, B5O, 32, 64, 32}};

struct a5_tc_state {
  __half* act[11] = {};
  __half* wpk2[11] = {};        // per layer: [stage][cta 2][kchunk 4][X rows | S rows][8]
  // block3-conv1 and block4-conv1 both read block2's output (network.py:68,79): the CTA-pair path
  // runs them as ONE layer of cout = 32 + 64 (N = 96 MMAs instead of N = 32/64 ones below the
  // ~44-cycle MMA floor, and the 128-channel input is read once)
  __half* wpk2_m = nullptr;
  float* bias_m = nullptr;
  // block3-conv2 (32->32 + r128) and block4-conv2 (64->64 + r128) both add a 1x1 projection of block2's
  // output (network.py:52-56,68,79): as ONE layer with two 3x3 segments (columns [0,32) and [32,96)) and a
  // shared N = 96 residual segment the 302 MB tensor is read once instead of twice
  __half* wpk2_m2 = nullptr;
  float* bias_m2 = nullptr;
  int merge2 = 1;               // A5_TC_MERGE2=0: run the two layers separately
  __half* wpk_c1 = nullptr;     // conv1 weights for k_tc_conv1m
  uint32_t* c1_bits = nullptr;  // bitboards of the input planes (k_c1_bits)
  int zigzag = 1;               // A5_TC_ZIGZAG=0: every layer walks the groups in ascending order
  int resw = 1;                 // A5_TC_RESW=0: always stream weights through the stage ring
  int pdl = 1;                  // A5_TC_PDL=0: plain stream-ordered launches
  // The lo halves of activations and conv weights are rounded to 10 - lo_drop mantissa bits (default 4: 6 bits, an
  // operand = 11 + 7 significant bits).  The chip is power-capped under tensor load and the clock follows the
  // switching activity of the multiplier arrays: fewer significant bits in two of the three passes = +3 % clock,
  // with no measurable change of the output error (profiles/r02_lo_bits.txt).  A5_TC_LO_DROP=0..8 overrides.
  int lo_drop = 4;
  uint32_t lo_add = 0, lo_mask = 0xFFFFFFFFu;   // both fp16 lanes of a packed pair
  int mega = 0;                 // all block convs as one chunk-major launch (k_tc_mega); default: S >= 13, A5_TC_MEGA=0/1
  int merge = 1;                // A5_TC_MERGE=0: run block3-conv1 / block4-conv1 separately
  long long plane_rows = 0;
  int fold = 1;
  int num_sms = 0;               // SMs the block convs may use (one persistent CTA each)
  int front_sms = 0;             // SMs conv1 may use (a5_net_set_sm_limit: SM-partitioned streams)
  a5::HeadsState* heads = nullptr;
};

namespace a5 {

// launch with the programmatic-stream-serialization attribute (see pdl_wait)
template <typename K>
static cudaError_t launch_pdl(K kernel, int grid, int threads, size_t smem, cudaStream_t st, const TCLayer& L, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, L);
}

static long long tc_plane_rows(const a5_net* net) {
  PosSpace ps(net->S);
  long long valid = (long long)net->max_batch * ps.per_board;
  long long padded = (valid + 1023) / 1024 * 1024;       // whole pair-groups for every T
  // + one chunk of slack: the megakernel's CTAs all run the same number of groups, the surplus ones read (never
  // write) rows past the last board
  return ps.guard + padded + ps.guard + TC_HALO + 1024 + 64;
}

int tc_alloc(a5_net* net) {
  a5_tc_state* tc = new a5_tc_state();
  net->tc = tc;
  tc->plane_rows = tc_plane_rows(net);
  for (int i = 0; i < 11; ++i) {
    size_t bytes = (size_t)2 * kTcActCh[i] * tc->plane_rows * sizeof(__half);
    A5_CUDA(cudaMalloc(&tc->act[i], bytes));
    A5_CUDA(cudaMemset(tc->act[i], 0, bytes));            // guard bands / pad cells stay zero
  }
  for (int l = 1; l <= 10; ++l) {
    const TcLayerDef& L = kTcLayers[l];
    size_t stages = (size_t)(L.cin / TC_KS) * 9 + (L.res_src >= 0 ? L.res_cin / TC_KS : 0);
    A5_CUDA(cudaMalloc(&tc->wpk2[l], stages * 2 * 4 * (L.cout + L.cout / 2) * 8 * sizeof(__half)));
  }
  A5_CUDA(cudaMalloc(&tc->wpk2_m, (size_t)(128 / TC_KS) * 9 * 2 * 4 * 96 * 8 * sizeof(__half)));
  A5_CUDA(cudaMalloc(&tc->bias_m, 96 * sizeof(float)));
  A5_CUDA(cudaMalloc(&tc->wpk2_m2, (size_t)2 * (9 * 2048 + 18 * 4096 + 4 * 6144)));
  A5_CUDA(cudaMalloc(&tc->bias_m2, 96 * sizeof(float)));
  A5_CUDA(cudaMalloc(&tc->wpk_c1, C1M_WBYTES));
  A5_CUDA(cudaMalloc(&tc->c1_bits, (size_t)net->max_batch * C1M_BW * sizeof(uint32_t)));
  A5_CUDA(cudaFuncSetAttribute(k_tc_conv1m, cudaFuncAttributeMaxDynamicSharedMemorySize, C1M_SMEM));
  A5_CUDA(cudaFuncSetAttribute(k_tc_conv2<2, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_LIMIT));
  A5_CUDA(cudaFuncSetAttribute(k_tc_conv2<2, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_LIMIT));
  A5_CUDA(cudaFuncSetAttribute(k_tc_conv2<2, false, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_LIMIT));
  A5_CUDA(cudaFuncSetAttribute(k_tc_conv2<2, true, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_LIMIT));
  A5_CUDA(cudaFuncSetAttribute(k_tc_mega<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_LIMIT));
  A5_CUDA(cudaFuncSetAttribute(k_tc_mega<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_LIMIT));
  // tuning knobs: M tiles per group for Cout = 128 / 64, and N-folding of the hi/lo weight halves
  const char* ev;
  tc->fold = ((ev = getenv("A5_TC_FOLD")) && atoi(ev) == 0) ? 0 : 1;
  tc->zigzag = ((ev = getenv("A5_TC_ZIGZAG")) && atoi(ev) == 0) ? 0 : 1;
  tc->resw = ((ev = getenv("A5_TC_RESW")) && atoi(ev) == 0) ? 0 : 1;
  tc->pdl = ((ev = getenv("A5_TC_PDL")) && atoi(ev) == 0) ? 0 : 1;
  tc->merge2 = ((ev = getenv("A5_TC_MERGE2")) && atoi(ev) == 0) ? 0 : 1;
  tc->merge = ((ev = getenv("A5_TC_MERGE")) && atoi(ev) == 0) ? 0 : 1;
  // chunk-major megakernel: measured +3 % at 15x15 (256 positions per board = whole tiles, chunks of 4 boards) and
  // +0.7 % at 11x11 (within run-to-run noise): on by default for boards of 13x13 and more, A5_TC_MEGA=0/1 forces it
  tc->mega = (ev = getenv("A5_TC_MEGA")) ? (atoi(ev) != 0) : (net->S >= 13);
  if ((ev = getenv("A5_TC_LO_DROP")) && atoi(ev) >= 0 && atoi(ev) <= 8) tc->lo_drop = atoi(ev);
  {
    const uint32_t d = (uint32_t)tc->lo_drop, r = d ? (1u << (d - 1)) : 0u, m = ~((1u << d) - 1u) & 0xFFFFu;
    tc->lo_add = r | (r << 16);
    tc->lo_mask = m | (m << 16);
  }
  int hrc = heads_alloc(net, &tc->heads);
  if (hrc) return hrc;
  int dev = 0;
  A5_CUDA(cudaGetDevice(&dev));
  A5_CUDA(cudaDeviceGetAttribute(&tc->num_sms, cudaDevAttrMultiProcessorCount, dev));
  tc->front_sms = tc->num_sms;
  return A5_OK;
}

void tc_free(a5_net* net) {
  if (!net->tc) return;
  for (int i = 0; i < 11; ++i) cudaFree(net->tc->act[i]);
  for (int i = 0; i < 11; ++i) cudaFree(net->tc->wpk2[i]);
  cudaFree(net->tc->wpk2_m);
  cudaFree(net->tc->bias_m);
  cudaFree(net->tc->wpk2_m2);
  cudaFree(net->tc->bias_m2);
  cudaFree(net->tc->wpk_c1);
  cudaFree(net->tc->c1_bits);
  heads_free(net->tc->heads);
  delete net->tc;
  net->tc = nullptr;
}

int tc_set_weights(a5_net* net, const float* const* t, cudaStream_t st) {
  a5_tc_state* tc = net->tc;
  for (int l = 1; l <= 10; ++l) {
    const TcLayerDef& L = kTcLayers[l];
    const int t0 = kBlocks[(l - 1) / 2].t0;
    const float* w = (l & 1) ? t[t0 + 2] : t[t0 + 4];
    const float* wres = (l & 1) ? nullptr : t[t0 + 0];
    k_tc_pack2<<<256, 256, 0, st>>>(w, nullptr, 0, wres, 9, L.cin, L.res_cin, L.cout, (L.cout <= 64) ? tc->fold : 0, tc->wpk2[l], tc->lo_drop);
    A5_CUDA(cudaGetLastError());
  }
  k_c1m_pack<<<(C1M_KCH * 64 * 8 + 255) / 256, 256, 0, st>>>(net->w[0], 64, tc->wpk_c1);
  A5_CUDA(cudaGetLastError());
  k_tc_pack2<<<256, 256, 0, st>>>(t[T_B3_C1_K], t[T_B4_C1_K], 32, nullptr, 9, 128, 0, 96, 0, tc->wpk2_m, tc->lo_drop);
  A5_CUDA(cudaGetLastError());
  A5_CUDA(cudaMemcpyAsync(tc->bias_m, net->bias[5], 32 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  A5_CUDA(cudaMemcpyAsync(tc->bias_m + 32, net->bias[7], 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  {
    // merged conv2 layer: [block3-conv2 3x3 | block4-conv2 3x3 | (block3-res | block4-res) 1x1]
    uint8_t* base = (uint8_t*)tc->wpk2_m2;
    k_tc_pack2_seg<<<64, 256, 0, st>>>(t[T_B3_C2_K], nullptr, 0, 9, 32, 32, (__half*)base, tc->lo_drop);
    k_tc_pack2_seg<<<64, 256, 0, st>>>(t[T_B4_C2_K], nullptr, 0, 9, 64, 64, (__half*)(base + 2 * 9 * 2048), tc->lo_drop);
    k_tc_pack2_seg<<<64, 256, 0, st>>>(t[T_B3_RES_K], t[T_B4_RES_K], 32, 1, 128, 96, (__half*)(base + 2 * (9 * 2048 + 18 * 4096)), tc->lo_drop);
    A5_CUDA(cudaGetLastError());
    A5_CUDA(cudaMemcpyAsync(tc->bias_m2, net->bias[6], 32 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    A5_CUDA(cudaMemcpyAsync(tc->bias_m2 + 32, net->bias[8], 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return heads_set_weights(net, tc->heads, t, st);
}

// profiling hook (a5__debug_layer_times): when set, an event is recorded after every launch group
static cudaEvent_t* g_tc_events = nullptr;
static bool g_tc_keep_head_acts = false;           // a5__debug_activation wants block3/5 outputs stored too
static unsigned long long* g_tc_dbg = nullptr;   // a5__debug_timeline: [10 layers][4 roles][256]
static unsigned long long* g_tc_clk = nullptr;   // a5__debug_clk: [8 conv launches][4]
static unsigned long long* g_tc_mega_dbg = nullptr;   // a5__debug_mega_clk: [nchunks * layers + 1][2]
#define TC_MARK(i) do { if (g_tc_events) cudaEventRecord(g_tc_events[i], st); } while (0)

// heads + biases come from the fp32 path's packed copies (fp32_set_weights runs first)
int tc_forward(a5_net* net, const int8_t* planes, int n, float* prob, float* value, cudaStream_t st, int parts) {
  a5_tc_state* tc = net->tc;
  PosSpace ps(net->S);
  TC_MARK(0);
  const long long nrows1 = (long long)n * ps.per_board;
  if (parts & A5_NET_PART_FRONT) {
    const int ntiles = (int)((nrows1 + 127) / 128);
    const int grid1 = ntiles < C1M_CTAS * tc->front_sms ? ntiles : C1M_CTAS * tc->front_sms;
    A5_CUDA(launch_pdl_k(k_c1_bits, (unsigned)((n + 3) / 4), 128, 0, st, tc->pdl != 0, planes, n, net->C, tc->c1_bits,
                         kt_slot(KT_C1BITS)));
    A5_CUDA(launch_pdl_k(k_tc_conv1m, (unsigned)grid1, C1M_THREADS, C1M_SMEM, st, tc->pdl != 0, (const uint32_t*)tc->c1_bits,
                         (const __half*)tc->wpk_c1, (const float*)net->bias[0], tc->act[A32], tc->plane_rows, net->S, ps.pitch,
                         ps.per_board, ps.guard, n, ntiles, tc->lo_add, tc->lo_mask, kt_slot(KT_CONV1)));
  }
  A5_CUDA(cudaGetLastError());
  TC_MARK(1);
  const long long nrows = (long long)n * ps.per_board;
  int nexec = 0;                                   // conv layers launched so far (alternating walk direction)
  const bool mega = tc->mega && tc->merge && tc->merge2 && !g_tc_dbg && !g_tc_events && !g_tc_keep_head_acts;
  TCMega M;
  if (mega) memset(&M, 0, sizeof(M));
  for (int l = 1; l <= 10 && (parts & A5_NET_PART_BODY); ++l) {
    TcLayerDef D = kTcLayers[l];
    TCLayer L;
    memset(&L, 0, sizeof(L));
    const bool merged = tc->merge && l == 5;      // block3-conv1 + block4-conv1 as one cout = 96 layer
    const bool merged2 = tc->merge2 && l == 6;    // block3-conv2 + block4-conv2 (+ both residual projections)
    if ((tc->merge && l == 7) || (tc->merge2 && l == 8)) { TC_MARK(1 + l); continue; }
    if (merged || merged2) D.cout = 96;
    L.src = tc->act[D.src]; L.src_ch = D.cin;
    L.res = D.res_src >= 0 ? tc->act[D.res_src] : nullptr; L.res_ch = D.res_cin;
    L.bias = net->bias[l];
    L.out = tc->act[D.out];
    L.out_f32 = nullptr;
    if (l == 6 || l == 10) {          // block3 / block5 outputs feed only the heads: fused 1x1 conv, nothing else stored
      const HeadsIO io = heads_io(tc->heads);
      L.out = g_tc_keep_head_acts ? L.out : nullptr;
      L.head_ch = l == 6 ? 4 : 16;
      L.head_w = l == 6 ? net->vconv_w : io.pconv_w;
      L.head_b = l == 6 ? net->vconv_b : io.pconv_b;
      L.head_out = l == 6 ? io.a_val : io.a_pol;
      L.head_nst = l == 6 ? io.nst_val : io.nst_pol;
    }
    if (merged) { L.bias = tc->bias_m; L.out2 = tc->act[B4H]; L.split = 32; }
    L.cout = D.cout; L.ntaps = 9;
    L.n0 = D.cout;
    if (merged2) {
      L.src2 = tc->act[B4H]; L.src2_ch = 64; L.n0 = 32;
      L.bias = tc->bias_m2; L.out2 = tc->act[B4O]; L.split = 32;
    }
    int k = 0;
    for (int ky = -1; ky <= 1; ++ky)
      for (int kx = -1; kx <= 1; ++kx) L.shifts[k++] = ky * ps.pitch + kx;
    L.fold = (D.cout <= 64) ? tc->fold : 0;
    // conv1 writes ascending; from there on every layer starts where its inputs were touched last
    L.reverse = tc->zigzag && (nexec++ % 2 == 0);
    L.dbg = g_tc_dbg ? g_tc_dbg + (size_t)(l - 1) * 8 * 256 : nullptr;
    L.kt = kt_slot(KT_CONV0 + (nexec - 1));
    L.clk = g_tc_clk ? g_tc_clk + 4 * (nexec - 1) : nullptr;
    L.plane_rows = tc->plane_rows; L.row0 = ps.guard; L.nrows = nrows;
    L.lo_add = tc->lo_add; L.lo_mask = tc->lo_mask;
    L.S = net->S; L.pitch = ps.pitch; L.per_board = ps.per_board;
    if (mega) {
      L.wpk = merged ? tc->wpk2_m : (merged2 ? tc->wpk2_m2 : tc->wpk2[l]);
      L.reverse = 0; L.kt = nullptr; L.clk = nullptr;
      M.L[M.nl++] = L;
      continue;
    }
    {
      // CTA pairs: T = 2 tiles per CTA, 4 per weight stage; TMEM double-buffers for every layer
      constexpr int T = 2;
      const int ngroups = (int)((nrows + T * 128 - 1) / (T * 128));
      const int npairs = (ngroups + 1) / 2;
      L.ngroups = ngroups;
      L.wpk = merged ? tc->wpk2_m : (merged2 ? tc->wpk2_m2 : tc->wpk2[l]);
      const int maxpairs = tc->num_sms / 2;
      const int grid = 2 * (npairs < maxpairs ? npairs : maxpairs);
      const bool h16 = ps.pitch + 1 <= 16;
      const int slab = h16 ? TCfgH<T, 16>::SLAB : TCfgH<T, 24>::SLAB;
      const int xr = L.fold ? D.cout : D.cout / 2, sr = D.cout / 2;
      const int res_slabs = D.res_src >= 0 ? D.res_cin / TC_KS : 0;
      const int nstage = (D.cin / TC_KS) * 9 + res_slabs;
      // weights resident in shared memory when the whole set fits beside enough slab buffers: two if
      // every slab carries nine taps of MMAs, three if one-tap residual slabs must be prefetched past
      const int wres = merged2 ? 9 * 2048 + 18 * 4096 + 4 * 6144 : (nstage * 4 * (xr + sr) * 16 + 127) & ~127;
      const int nsb_res = (TC2_SMEM_LIMIT - TC2_MISC - wres) / slab;
      const bool resw = tc->resw && nsb_res >= (res_slabs ? 3 : 2);
      L.w_bytes = resw ? wres : TC2_WSTAGES * TC2_WSTAGE_MAX;
      const int nsb = (TC2_SMEM_LIMIT - TC2_MISC - L.w_bytes) / slab;
      L.nslab_buf = nsb > 4 ? 4 : nsb;
      const size_t smem = (size_t)L.nslab_buf * slab + L.w_bytes + TC2_MISC;
      if (h16 && resw) A5_CUDA(launch_pdl(k_tc_conv2<T, true, 16>, grid, TC2_THREADS, smem, st, L, tc->pdl));
      else if (h16) A5_CUDA(launch_pdl(k_tc_conv2<T, false, 16>, grid, TC2_THREADS, smem, st, L, tc->pdl));
      else if (resw) A5_CUDA(launch_pdl(k_tc_conv2<T, true, 24>, grid, TC2_THREADS, smem, st, L, tc->pdl));
      else A5_CUDA(launch_pdl(k_tc_conv2<T, false, 24>, grid, TC2_THREADS, smem, st, L, tc->pdl));
    }
    A5_CUDA(cudaGetLastError());
    TC_MARK(1 + l);
  }
  if (mega && (parts & A5_NET_PART_BODY)) {
    // launch order: b1c1 b1c2 b2c1 b2c2 (b3c1+b4c1) (b3c2+b4c2) b5c1 b5c2; who wrote each layer's inputs
    static const int dsrc[8] = {-1, 0, 1, 2, 3, 4, 5, 6}, dsrc2[8] = {-1, -1, -1, -1, -1, 4, -1, -1}, dres[8] = {-1, -1, -1, 1, -1, 3, -1, 5};
    for (int i = 0; i < 8; ++i) { M.dep_src[i] = dsrc[i]; M.dep_src2[i] = dsrc2[i]; M.dep_res[i] = dres[i]; }
    int grid = tc->num_sms & ~1;
    if (n < grid) grid = (n + 1) & ~1;
    M.B = n / grid; M.n_hi = n % grid;
    const int bmax = M.B + (M.n_hi ? 1 : 0);
    // boards per chunk: every CTA runs nchunks x gpc groups whatever it owns, so pick the chunk (<= 2048 positions:
    // a layer's output of 148 such chunks must stay in L2) that wastes the fewest groups; ties -> the smaller chunk
    int cb = 1, best = 1 << 30;
    for (int c = 1; c <= bmax && (c == 1 || c * ps.per_board <= 2048); ++c) {
      if ((c * ps.per_board + 255) / 256 < 3 && c < bmax) continue;      // < 3 groups per layer: the pipelines drain at every layer
      const int cost = ((bmax + c - 1) / c) * ((c * ps.per_board + 255) / 256);
      if (cost < best) { best = cost; cb = c; }
    }
    { const char* ev = getenv("A5_TC_MEGA_CB"); if (ev && atoi(ev) >= 1 && atoi(ev) <= bmax) cb = atoi(ev); }
    M.CB = cb;
    M.gpc = (cb * ps.per_board + 255) / 256;
    M.nchunks = (bmax + cb - 1) / cb;
    M.dbg = g_tc_mega_dbg;                             // a5__debug_mega_clk: [(nchunks * nl + 1)][2]
    M.kt = kt_slot(KT_CONV0);
    const bool h16 = ps.pitch + 1 <= 16;
    const int slab = h16 ? TCfgH<2, 16>::SLAB : TCfgH<2, 24>::SLAB;
    int nsb = (TC2_SMEM_LIMIT - TCM_MISC - TCM_WBYTES) / slab;
    if (nsb > 4) nsb = 4;
    const size_t smem = (size_t)nsb * slab + TCM_WBYTES + TCM_MISC;
    if (h16) A5_CUDA(launch_pdl_k(k_tc_mega<16>, (unsigned)grid, TC2_THREADS, smem, st, tc->pdl != 0, M));
    else A5_CUDA(launch_pdl_k(k_tc_mega<24>, (unsigned)grid, TC2_THREADS, smem, st, tc->pdl != 0, M));
  }
  int rc = (parts & A5_NET_PART_HEADS) ? heads_forward(net, tc->heads, n, prob, value, st) : A5_OK;
  TC_MARK(12);
  return rc;
}

}  // namespace a5

extern "C" {
int a5_net_tc_available(void) { return 1; }

int a5_net_set_sm_limit(a5_net* net, int body_sms, int front_sms) {
  A5_ARG(net && net->tc && body_sms >= 2 && front_sms >= 1);
  int dev = 0, total = 0;
  A5_CUDA(cudaGetDevice(&dev));
  A5_CUDA(cudaDeviceGetAttribute(&total, cudaDevAttrMultiProcessorCount, dev));
  net->tc->num_sms = body_sms < total ? body_sms : total;
  net->tc->front_sms = front_sms < total ? front_sms : total;
  return A5_OK;
}

int a5_net_forward_parts(a5_net* net, const int8_t* d_planes, int n, float* d_prob, float* d_value, int parts, void* stream) {
  A5_ARG(net && net->tc && d_planes && n >= 0 && n <= net->max_batch && (parts & ~A5_NET_PART_ALL) == 0);
  A5_ARG(!(parts & A5_NET_PART_HEADS) || (d_prob && d_value));
  if (!net->has_weights) { set_error("a5_net_forward_parts: no weights set"); return A5_ERR_STATE; }
  if (n == 0 || parts == 0) return A5_OK;
  return tc_forward(net, d_planes, n, d_prob, d_value, (cudaStream_t)stream, parts);
}

// internal tooling (not part of alphafive.h): in-situ kernel timing.  enable(1) arms the stamps for kernels
// launched (or captured into a graph) from now on and clears the sums; fold(stream) closes a pass (launch it
// after the tree pass, also inside the captured graph); read copies double[KT_SLOTS][3] =
// {ns predecessor-end -> end, ns start -> end, passes}.
int a5__debug_ktime_enable(int on) {
  const size_t bytes = ((size_t)KT_SLOTS * KT_SUB * 2 + KT_SLOTS * 3 + 1) * sizeof(unsigned long long);
  if (on && !g_kt_dev) A5_CUDA(cudaMalloc(&g_kt_dev, bytes));
  if (on) {
    A5_CUDA(cudaDeviceSynchronize());
    A5_CUDA(cudaMemset(g_kt_dev, 0, bytes));
    k_kt_fold<<<1, 32>>>(g_kt_dev);                      // sets every stamp to (start_min = ~0, end_max = 0)
    A5_CUDA(cudaGetLastError());
    A5_CUDA(cudaDeviceSynchronize());
    // sums and "previous pass end" cleared: the first pass after this only sets the reference point
    A5_CUDA(cudaMemset(g_kt_dev + (size_t)KT_SLOTS * KT_SUB * 2, 0, (KT_SLOTS * 3 + 1) * sizeof(unsigned long long)));
    A5_CUDA(cudaDeviceSynchronize());
  }
  g_kt_on = on != 0;
  return A5_OK;
}
int a5__debug_ktime_fold(void* stream) {
  if (!g_kt_on || !g_kt_dev) return A5_OK;
  k_kt_fold<<<1, 32, 0, (cudaStream_t)stream>>>(g_kt_dev);
  A5_CUDA(cudaGetLastError());
  return A5_OK;
}
int a5__debug_ktime_read(double* h_out) {
  A5_ARG(h_out && g_kt_dev);
  unsigned long long acc[KT_SLOTS * 3];
  A5_CUDA(cudaDeviceSynchronize());
  A5_CUDA(cudaMemcpy(acc, g_kt_dev + (size_t)KT_SLOTS * KT_SUB * 2, sizeof(acc), cudaMemcpyDeviceToHost));
  for (int i = 0; i < KT_SLOTS * 3; ++i) h_out[i] = (double)acc[i];
  return A5_OK;
}

// internal tooling: device buffer uint64 [8][4] receiving {clock64, globaltimer} at start / end of CTA 0 of every
// block-conv launch from now on (null: off) -- the SM clock each layer actually ran at
int a5__debug_clk(unsigned long long* d_buf) { g_tc_clk = d_buf; return A5_OK; }
// the same for the chunk-major megakernel: uint64 [nchunks * 8 + 1][2] = {%globaltimer, clock64} when CTA 0's issuer starts
// each (chunk, layer), and at the kernel's end
int a5__debug_mega_clk(unsigned long long* d_buf) { g_tc_mega_dbg = d_buf; return A5_OK; }

// internal tooling: also store the block3 / block5 activations (normally consumed in-register by
// the fused head convs) so a5__debug_activation can show them.
int a5__debug_keep_head_acts(int on) { g_tc_keep_head_acts = on != 0; return A5_OK; }

// internal tooling: clock64 timeline of CTA 0 for every conv layer of one forward;
// d_dbg = uint64 [10][8][256] (zeroed by the caller): roles 0 producer, 1 MMA issuer (groups/slabs),
// 2 epilogue warp 2, 3 MMA issuer (weight stages: wait, landed).
int a5__debug_timeline(a5_net* net, const int8_t* d_planes, int n, float* d_prob, float* d_value,
                       unsigned long long* d_dbg, void* stream) {
  A5_ARG(net && d_planes && d_dbg);
  g_tc_dbg = d_dbg;
  int rc = tc_forward(net, d_planes, n, d_prob, d_value, (cudaStream_t)stream, A5_NET_PART_ALL);
  g_tc_dbg = nullptr;
  return rc;
}

// internal tooling (not part of alphafive.h): average milliseconds of each launch group of the
// tensor-core forward over `reps` runs: h_ms[0] conv1, h_ms[1..10] block convs, h_ms[11] heads.
int a5__debug_layer_times(a5_net* net, const int8_t* d_planes, int n, int reps, float* d_prob, float* d_value,
                          float* h_ms, void* stream) {
  A5_ARG(net && d_planes && h_ms && reps > 0);
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t ev[13];
  for (int i = 0; i < 13; ++i) A5_CUDA(cudaEventCreate(&ev[i]));
  for (int i = 0; i < 12; ++i) h_ms[i] = 0.0f;
  for (int r = 0; r < reps; ++r) {
    g_tc_events = ev;
    int rc = tc_forward(net, d_planes, n, d_prob, d_value, st, A5_NET_PART_ALL);
    g_tc_events = nullptr;
    if (rc) return rc;
    A5_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 12; ++i) {
      float ms = 0.0f;
      A5_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      h_ms[i] += ms / reps;
    }
  }
  for (int i = 0; i < 13; ++i) cudaEventDestroy(ev[i]);
  return A5_OK;
}

// internal tooling (not part of alphafive.h): activation `idx` of the last forward of `mode`
// as fp32 [rows][channels] with rows = n * (S+1)^2, for layer-by-layer parity debugging.
int a5__debug_activation(a5_net* net, int mode, int idx, int n, float* d_out, void* stream) {
  A5_ARG(net && idx >= 0 && idx < 11 && d_out);
  PosSpace ps(net->S);
  const long long nrows = (long long)n * ps.per_board;
  const int ch = kTcActCh[idx];
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == A5_NET_FP32) {
    A5_CUDA(cudaMemcpyAsync(d_out, net->act[idx] + (size_t)ps.guard * ch, (size_t)nrows * ch * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    const __half* x = net->tc->act[idx];
    long long total = nrows * ch;
    // rows are offset by the guard inside each plane
    k_tc_unpack<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x + (size_t)ps.guard * 8, ch, net->tc->plane_rows, nrows, d_out);
    A5_CUDA(cudaGetLastError());
  }
  return A5_OK;
}
}
