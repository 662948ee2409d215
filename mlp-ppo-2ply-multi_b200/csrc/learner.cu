// learner.cu -- the TD(0) learner: Trainer.update (reference src/agents/trainer.py:48-166) as ONE persistent thread-block cluster.
//
// The reference trains SEQUENTIALLY: for every episode of the 200-episode batch one forward pass over its T observations,
// TD(0) targets r_t + gamma V(x_{t+1}) (last target = r_T), mse loss, backward, clip_grad_norm_(1.0), one Adam step; the next
// episode sees the updated weights.  That chain cannot be batched without changing the result, so the B200 design attacks its
// latency instead: one cluster of CL <= 8 CTAs keeps the whole optimiser state in shared memory for the entire batch (hidden
// unit u lives in CTA u / UPC: its fc1 row, b1, w2 and their Adam moments), every episode costs two cluster barriers (value
// partial sums and gradient-norm partial sums are pushed into every peer's shared memory over DSMEM), and no kernel launch,
// host sync or HBM round trip happens between the 200 optimiser steps.  Observations arrive as compact int8[52] boards (or
// straight from the arena's episode records) and are expanded on chip into a sparse feature list per row (forward) and a
// row bitmask per feature (backward), so neither the [T,198] feature matrix nor any activation ever exists in HBM.
// Everything is deterministic: all sums run in a fixed order (no atomics).
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include <new>

#include "learner.cuh"

namespace cg = cooperative_groups;

namespace bg {
namespace {

constexpr int NF = 198;
constexpr int NROW = 200;  // per hidden unit: 198 fc1 weights, b1, w2 (row f of the packed [W1t | b1 | w2] matrix)
constexpr int TMAX = 320;  // experiences per episode the kernel accepts (reference MAX_TIMESTEPS = 300)
constexpr int NT = 256;
constexpr int MAXCL = 8;    // portable cluster size (k_td0_update<32>)
constexpr int MAXCL_TC = 16;  // k_td0_update_tc: 16 units per CTA, so H = 256 needs 16 CTAs (non-portable cluster size, opt-in)

#ifdef BG_LEARNER_PROFILE
#define PH(k)                                      \
  do {                                             \
    const long long now_ = clock64();              \
    if (tid == 0 && rank == 0) ph[k] += now_ - t_; \
    t_ = now_;                                     \
  } while (0)
#else
#define PH(k)
#endif

struct LearnerArgs {
  const int8_t* boards;
  const uint8_t* flags;  // records == 0: observation flag; records == 1: arena meta byte (bit0 = mover)
  const float* reward;
  const int64_t* ep_offsets;
  const int32_t* ep_len;  // optional: explicit episode lengths (ep_offsets then only gives the first row of each episode)
  int64_t n_eps;
  int32_t records;
  float* params;  // packed [W1t(198,H) | b1 | w2 | b2]
  float* m;
  float* v;
  int64_t* step;
  float lr, gamma, grad_clip;
  float* metrics;
  int32_t* status;
  int32_t H;
};

constexpr int EC = 512;   // episodes whose offsets are staged in shared memory at a time
constexpr int NSC = NT;   // Adam bias-correction scalars precomputed for this many steps at a time
constexpr int LROW = 40;  // sparse-row capacity: <= 15 + 15 point features, 2 bar, 2 off, 1 flag = 35
constexpr int TW = TMAX / 32;

template <int UPC>
struct Smem {
  float P[NROW * UPC], M[NROW * UPC], V[NROW * UPC];
  alignas(16) float hs[TMAX * UPC];  // dL/dz of the episode, [t][unit]
  float ypart[MAXCL * TMAX];   // ypart[c][t]: CTA c's partial of V(x_t), pushed by CTA c
  float Y[TMAX], dY[TMAX], rew[TMAX];
  alignas(16) float pr[2][NT / 32][UPC];  // per-warp partials of the w2 / b1 gradients
  float2 sc[NSC];              // (lr / bias_correction1, 1 / sqrt(bias_correction2)) for the next NSC optimiser steps
  int64_t offs[EC + 1], lens[EC];
  float npart[MAXCL];          // per-CTA sums of squared gradients, pushed by each CTA
  float red[8][8];
  float red2[8];
  float vtab[32];              // list value codes: 0 -> 1, k (1..15) -> k/2, 16+k -> k/15
  float ftab[4][16];           // feature value by board byte: [0] 1, [1] max(c-3,0)/2, [2] c/2, [3] c/15
  float b2[4];                 // b2 and its Adam moments (replicated in every CTA, updated identically)
  uint32_t brd[2][TMAX * 13];  // observation boards, double buffered: episode e+1 is fetched while e is processed
  uint32_t nz[NF * TW];        // nz[f][w] bit b: feature f of row 32w+b is non-zero
  uint16_t list[TMAX * LROW];  // sparse rows: feature index | value code << 8
  uint8_t lcnt[TMAX];
  uint8_t flg[TMAX];
};

static_assert(sizeof(Smem<32>) <= 232448, "k_td0_update<32> exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ float adam1(float p, float& m, float& v, float g, float step_size, float inv_bc2_sqrt) {
  // torch/optim/adam.py _single_tensor_adam: lerp, mul + addcmul, sqrt / bias_correction2_sqrt + eps, addcdiv
  // (sqrt and the two divisions through the SFU approximations: relative error ~2e-7 of a 1e-3 step)
  m = m + 0.1f * (g - m);
  v = v * 0.999f + 0.001f * g * g;
  const float denom = fmaf(v > 0.f ? v * rsqrtf(v) : 0.f, inv_bc2_sqrt, 1e-8f);
  return p - step_size * __fdividef(m, denom);
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// number of non-zero features of 4 board bytes packed in a word: sum over bytes of min(c, 4)
__device__ __forceinline__ int word_entries(uint32_t w) {
  const uint32_t s = __vsetgtu4(w, 0u) + __vsetgtu4(w, 0x01010101u) + __vsetgtu4(w, 0x02020202u) + __vsetgtu4(w, 0x03030303u);
  return (int)__dp4a(s, 0x01010101u, 0u);
}

// Thread layout.  The CTA's UPC hidden units are split into QJ = UPC / 4 quads; a thread owns one quad (float4 weights) and
// one of FL = NT / QJ lanes, which index ROWS in the forward pass and FEATURES in the backward pass and the Adam step.
// Backward features are visited in the order o -> f(o): the 48 points' ">= 1" features first, then ">= 2", ">= 3", the
// excess features, then bar/off/flag, b1 and w2, so that the 8 (or 4) lanes of a warp scan bitmasks of similar density.
__device__ __forceinline__ int feature_of(int o) { return o < 192 ? (o % 48) * 4 + o / 48 : o; }

template <int UPC>
__global__ void __launch_bounds__(NT, 1) k_td0_update(const LearnerArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CL = (int)cluster.num_blocks();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<UPC>& S = *reinterpret_cast<Smem<UPC>*>(smem_raw);
  constexpr int QJ = UPC / 4;
  constexpr int FL = NT / QJ;
  constexpr int MAXR = TMAX / FL;             // rows per thread (forward)
  constexpr int KF = (NROW + FL - 1) / FL;    // features per thread (backward / Adam)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jq = tid % QJ, fl = tid / QJ;
  const int H = a.H;
  const int64_t iB2 = (int64_t)NROW * H;

  for (int s = tid; s < NROW * UPC; s += NT) {
    const int64_t gi = (int64_t)(s / UPC) * H + rank * UPC + (s % UPC);
    S.P[s] = a.params[gi];
    S.M[s] = a.m[gi];
    S.V[s] = a.v[gi];
  }
  if (tid == 0) {
    S.b2[0] = a.params[iB2];
    S.b2[1] = a.m[iB2];
    S.b2[2] = a.v[iB2];
  }
  if (tid < 32) {
    const int k = tid & 15;
    // reference immutable_board.py:99-112: (c-3)/2 and bar/2 are exact; off/15 is a double division stored to fp32
    S.vtab[tid] = tid == 0 ? 1.0f : tid < 16 ? 0.5f * (float)k : (float)((double)k / 15.0);
  }
  if (tid < 64) {
    const int k = tid & 15, kind = tid >> 4;
    S.ftab[kind][k] = kind == 0 ? 1.0f : kind == 1 ? (k > 3 ? 0.5f * (float)(k - 3) : 0.0f) : kind == 2 ? 0.5f * (float)k : (float)((double)k / 15.0);
  }
  const int64_t step0 = *a.step;
  int64_t kstep = 0;  // optimiser steps taken by this launch (thread-uniform)
  // torch/optim/adam.py: bias_correction = 1 - beta ** step in double; step_size = lr / bc1; bias_correction2_sqrt = sqrt(bc2)
  auto fill_scalars = [&]() {
    const double n = (double)(step0 + kstep + tid + 1);
    S.sc[tid] = make_float2((float)((double)a.lr / (1.0 - pow(0.9, n))), 1.0f / (float)sqrt(1.0 - pow(0.999, n)));
  };
  fill_scalars();

  // per-thread feature slots (constant over the launch): packed-row index, board byte and value table of each
  int fF[KF], fBo[KF], fKind[KF];
#pragma unroll
  for (int k = 0; k < KF; ++k) {
    const int o = fl + FL * k;
    const int f = o < NROW ? feature_of(o) : -1;
    fF[k] = f;
    fBo[k] = f < 0 ? 0 : f < 192 ? (f >> 2) : f == 192 ? 48 : f == 193 ? 50 : f == 194 ? 49 : f == 195 ? 51 : 0;
    fKind[k] = f < 0 ? 0 : f < 192 ? ((f & 3) == 3 ? 1 : 0) : f < 196 ? ((f & 1) ? 3 : 2) : 0;
  }
  __syncthreads();
  cluster.sync();  // every CTA of the cluster is resident before the first DSMEM store

  const uint32_t* gb32 = reinterpret_cast<const uint32_t*>(a.boards);

  // first episode at or after e (within the staged chunk [c0, c1)) that takes an optimiser step; empty and over-long ones are
  // reported and skipped.  Uniform over the cluster.
  int64_t c0 = 0, c1 = 0;
  auto next_valid = [&](int64_t e) {
    for (; e < c1; ++e) {
      const int64_t Tl = S.lens[e - c0];
      if (Tl > 0 && Tl <= TMAX) break;
      if (rank == 0 && tid == 0) {
        if (Tl > TMAX && a.status) *a.status = BG_ERR_CAPACITY;
        if (a.metrics)
          for (int k = 0; k < 6; ++k) a.metrics[e * 6 + k] = 0.0f;
      }
    }
    return e;
  };
  // start fetching episode e: boards by cp.async into buffer `buf`, flag / reward of rows tid and tid + NT into registers
  uint8_t pf_flag[2] = {0, 0};
  float pf_rew[2] = {0.f, 0.f};
  auto prefetch = [&](int64_t e, int buf) {
    const int64_t lo = S.offs[e - c0];
    const int T = (int)S.lens[e - c0];
    for (int w = tid; w < T * 13; w += NT) {
      const int t = w / 13, k = w - t * 13;
      if (a.records && t == 0)
        S.brd[buf][w] = initial_board_word(k);  // records mode: x_t's board is record t-1's afterstate, x_0's the start position
      else
        cp_async4(&S.brd[buf][w], gb32 + (a.records ? lo + t - 1 : lo + t) * 13 + k);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int t = tid + i * NT;
      if (t < T) {
        pf_flag[i] = a.flags[lo + t] & 1;
        pf_rew[i] = a.reward[lo + t];
      }
    }
  };

#ifdef BG_LEARNER_PROFILE
  long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long t_ = clock64();
#endif
  for (c0 = 0; c0 < a.n_eps; c0 += EC) {
    c1 = c0 + EC < a.n_eps ? c0 + EC : a.n_eps;
    __syncthreads();
    for (int i = tid; i < (int)(c1 - c0); i += NT) {  // episode e: rows [offs[e], offs[e] + len); len from ep_len or the next CSR offset
      S.offs[i] = a.ep_offsets[c0 + i];
      S.lens[i] = a.ep_len ? (int64_t)a.ep_len[c0 + i] : a.ep_offsets[c0 + i + 1] - a.ep_offsets[c0 + i];
    }
    __syncthreads();
    int buf = 0;
    int64_t e = next_valid(c0);
    if (e < c1) prefetch(e, buf);
    while (e < c1) {
      PH(9);
      const int T = (int)S.lens[e - c0];
      const int nch = (T + 31) >> 5;
      const uint32_t* brd32 = S.brd[buf];
      const uint8_t* brd8 = reinterpret_cast<const uint8_t*>(S.brd[buf]);

      // ---- A: land the staged episode; start fetching the next one ----
      cp_async_wait_all();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int t = tid + i * NT;
        if (t < T) {
          S.flg[t] = pf_flag[i];
          S.rew[t] = pf_rew[i];
        }
      }
      __syncthreads();
      const int64_t e_next = next_valid(e + 1);
      if (e_next < c1) prefetch(e_next, buf ^ 1);
      PH(0);
      // sparse feature rows for the forward pass: one (row, player half) task per thread, predicated stores, no branches;
      // player 2's half starts where player 1's ends (its length is counted with byte-SIMD compares)
      for (int task = tid; task < 2 * T; task += NT) {
        const int t = task >> 1, p = task & 1;
        const uint8_t* b = brd8 + t * 52;
        const int fg = S.flg[t];
        int cnt = 0;
        if (p) {
#pragma unroll
          for (int k = 0; k < 6; ++k) cnt += word_entries(brd32[t * 13 + k]);
          cnt += (b[48] > 0) + (b[50] > 0) + (fg == 0);
        }
        uint16_t* li = S.list + t * LROW;
#pragma unroll 4
        for (int pt = 0; pt < 24; ++pt) {
          const int c = b[p * 24 + pt];
          const int f0 = p * 96 + pt * 4;
          if (c > 0) li[cnt] = (uint16_t)f0;
          if (c > 1) li[cnt + 1] = (uint16_t)(f0 + 1);
          if (c > 2) li[cnt + 2] = (uint16_t)(f0 + 2);
          if (c > 3) li[cnt + 3] = (uint16_t)((f0 + 3) | ((c - 3) << 8));
          cnt += min(c, 4);
        }
        const int bar = b[48 + p], off = b[50 + p];
        if (bar > 0) li[cnt++] = (uint16_t)((192 + 2 * p) | (bar << 8));
        if (off > 0) li[cnt++] = (uint16_t)((193 + 2 * p) | ((16 + off) << 8));
        if (fg == p) li[cnt++] = (uint16_t)(196 + p);
        if (p) S.lcnt[t] = (uint8_t)cnt;
      }
      // row bitmasks per feature for the backward pass: one (32-row chunk, board byte) task per warp, lane == row
      for (int task = warp; task < nch * 53; task += NT / 32) {
        const int ch = task / 53, q = task - ch * 53;  // q < 48 points, 48..51 bar/off bytes, 52 the flag
        const int t = ch * 32 + lane;
        const bool valid = t < T;
        if (q < 48) {
          const int c = valid ? brd8[t * 52 + q] : 0;
          const uint32_t m0 = __ballot_sync(BG_FULL, c > 0), m1 = __ballot_sync(BG_FULL, c > 1), m2 = __ballot_sync(BG_FULL, c > 2),
                         m3 = __ballot_sync(BG_FULL, c > 3);
          if (lane < 4) S.nz[(q * 4 + lane) * TW + ch] = lane == 0 ? m0 : lane == 1 ? m1 : lane == 2 ? m2 : m3;
        } else if (q < 52) {
          const int c = valid ? brd8[t * 52 + q] : 0;
          const uint32_t m0 = __ballot_sync(BG_FULL, c > 0);
          const int f = q == 48 ? 192 : q == 49 ? 194 : q == 50 ? 193 : 195;
          if (lane == 0) S.nz[f * TW + ch] = m0;
        } else {
          const int fg = valid ? S.flg[t] : 2;
          const uint32_t m0 = __ballot_sync(BG_FULL, fg == 0), m1 = __ballot_sync(BG_FULL, fg == 1);
          if (lane < 2) S.nz[(196 + lane) * TW + ch] = lane == 0 ? m0 : m1;
        }
      }
      __syncthreads();
      PH(1);

      // ---- B: forward for this thread's quad of hidden units over its rows (kept in registers); value partials to every CTA ----
      float4 hreg[MAXR];
      const float4 b1q = *reinterpret_cast<const float4*>(&S.P[NF * UPC + jq * 4]);
      const float4 w2q = *reinterpret_cast<const float4*>(&S.P[(NF + 1) * UPC + jq * 4]);
#pragma unroll
      for (int k = 0; k < MAXR; k += 2) {
        if (k * FL < T) {  // uniform
          const int tA = fl + k * FL, tB = tA + FL;
          const bool vA = tA < T, vB = tB < T;
          const int nA = vA ? S.lcnt[tA] : 0, nB = vB ? S.lcnt[tB] : 0;
          const uint16_t* lA = S.list + (vA ? tA : 0) * LROW;
          const uint16_t* lB = S.list + (vB ? tB : 0) * LROW;
          float4 zA = b1q, zB = b1q;
          const int nmax = max(nA, nB);
#pragma unroll 2
          for (int q = 0; q < nmax; ++q) {
            // unconditional loads (a stale entry past a row's count stays inside the shared-memory block) + select
            const uint32_t eA = lA[q], eB = lB[q];
            const float4 wA = *reinterpret_cast<const float4*>(&S.P[(eA & 255u) * UPC + jq * 4]);
            const float4 wB = *reinterpret_cast<const float4*>(&S.P[(eB & 255u) * UPC + jq * 4]);
            const float xA = q < nA ? S.vtab[(eA >> 8) & 31u] : 0.f, xB = q < nB ? S.vtab[(eB >> 8) & 31u] : 0.f;
            if (q < nA) {
              zA.x = fmaf(xA, wA.x, zA.x);
              zA.y = fmaf(xA, wA.y, zA.y);
              zA.z = fmaf(xA, wA.z, zA.z);
              zA.w = fmaf(xA, wA.w, zA.w);
            }
            if (q < nB) {
              zB.x = fmaf(xB, wB.x, zB.x);
              zB.y = fmaf(xB, wB.y, zB.y);
              zB.z = fmaf(xB, wB.z, zB.z);
              zB.w = fmaf(xB, wB.w, zB.w);
            }
          }
          float4 hA, hB;
          hA.x = 1.0f / (1.0f + expf(-zA.x));
          hA.y = 1.0f / (1.0f + expf(-zA.y));
          hA.z = 1.0f / (1.0f + expf(-zA.z));
          hA.w = 1.0f / (1.0f + expf(-zA.w));
          hB.x = 1.0f / (1.0f + expf(-zB.x));
          hB.y = 1.0f / (1.0f + expf(-zB.y));
          hB.z = 1.0f / (1.0f + expf(-zB.z));
          hB.w = 1.0f / (1.0f + expf(-zB.w));
          hreg[k] = hA;
          if (k + 1 < MAXR) hreg[k + 1] = hB;
          float pA = vA ? fmaf(w2q.w, hA.w, fmaf(w2q.z, hA.z, fmaf(w2q.y, hA.y, w2q.x * hA.x))) : 0.f;
          float pB = vB ? fmaf(w2q.w, hB.w, fmaf(w2q.z, hB.z, fmaf(w2q.y, hB.y, w2q.x * hB.x))) : 0.f;
#pragma unroll
          for (int o = QJ / 2; o > 0; o >>= 1) {
            pA += __shfl_xor_sync(BG_FULL, pA, o);
            pB += __shfl_xor_sync(BG_FULL, pB, o);
          }
          for (int c = jq; c < CL; c += QJ) {
            float* dst = cluster.map_shared_rank(S.ypart, c) + rank * TMAX;
            if (vA) dst[tA] = pA;
            if (vB) dst[tB] = pB;
          }
        }
      }
      PH(2);
      cluster.sync();
      PH(3);

      // ---- C: values, TD(0) targets (trainer.py:110-115), dL/dY of the mse loss (:118), metric sums ----
      for (int t = tid; t < T; t += NT) {
        float y = S.b2[0];
        for (int c = 0; c < CL; ++c) y += S.ypart[c * TMAX + t];
        S.Y[t] = y;
      }
      __syncthreads();
      {
        float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int t = tid; t < T; t += NT) {
          const float y = S.Y[t];
          float tg = S.rew[t];
          if (t + 1 < T) tg = __fadd_rn(tg, __fmul_rn(a.gamma, S.Y[t + 1]));
          const float d = y - tg;
          const float dy = 2.0f * d / (float)T;
          S.dY[t] = dy;
          q[0] += d * d;
          q[1] += fabsf(d);
          q[2] += y;
          q[3] += S.rew[t];
          q[4] += dy;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q[k] += __shfl_xor_sync(BG_FULL, q[k], o);
        }
        if (lane == 0)
          for (int k = 0; k < 5; ++k) S.red[warp][k] = q[k];
      }
      __syncthreads();
      PH(4);

      // ---- D: dL/dz from the activations still in registers; w2 / b1 gradient partials, reduced over the warp's row lanes ----
      {
        float4 gw2 = make_float4(0.f, 0.f, 0.f, 0.f), gb1 = gw2;
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const int t = fl + k * FL;
          if (t < T) {
            const float dy = S.dY[t];
            const float4 h = hreg[k];
            float4 dz;
            gw2.x = fmaf(dy, h.x, gw2.x);
            gw2.y = fmaf(dy, h.y, gw2.y);
            gw2.z = fmaf(dy, h.z, gw2.z);
            gw2.w = fmaf(dy, h.w, gw2.w);
            dz.x = dy * w2q.x * h.x * (1.0f - h.x);
            dz.y = dy * w2q.y * h.y * (1.0f - h.y);
            dz.z = dy * w2q.z * h.z * (1.0f - h.z);
            dz.w = dy * w2q.w * h.w * (1.0f - h.w);
            *reinterpret_cast<float4*>(&S.hs[t * UPC + jq * 4]) = dz;
            gb1.x += dz.x;
            gb1.y += dz.y;
            gb1.z += dz.z;
            gb1.w += dz.w;
          }
        }
#pragma unroll
        for (int o = QJ; o < 32; o <<= 1) {
          gw2.x += __shfl_xor_sync(BG_FULL, gw2.x, o);
          gw2.y += __shfl_xor_sync(BG_FULL, gw2.y, o);
          gw2.z += __shfl_xor_sync(BG_FULL, gw2.z, o);
          gw2.w += __shfl_xor_sync(BG_FULL, gw2.w, o);
          gb1.x += __shfl_xor_sync(BG_FULL, gb1.x, o);
          gb1.y += __shfl_xor_sync(BG_FULL, gb1.y, o);
          gb1.z += __shfl_xor_sync(BG_FULL, gb1.z, o);
          gb1.w += __shfl_xor_sync(BG_FULL, gb1.w, o);
        }
        if (lane < QJ) {
          *reinterpret_cast<float4*>(&S.pr[0][warp][jq * 4]) = gw2;
          *reinterpret_cast<float4*>(&S.pr[1][warp][jq * 4]) = gb1;
        }
      }
      float msum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (tid == 0) {
        for (int k = 0; k < 5; ++k) {
          float s = 0.f;
          for (int w = 0; w < NT / 32; ++w) s += S.red[w][k];
          msum[k] = s;
        }
      }
      __syncthreads();
      PH(5);

      // ---- E: gradients of this thread's feature slots: a walk over the set bits of the feature's row mask, two rows per
      //         iteration; b1 / w2 slots sum the warp partials.  Kept in registers for the Adam step ----
      float4 g[KF];
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < KF; ++k) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = acc;
        const int f = fF[k];
        if (f >= 0 && f < NF) {
          const uint32_t* nzf = S.nz + f * TW;
          const float* tab = S.ftab[fKind[k]];
          const uint8_t* bb = brd8 + fBo[k];
          for (int w = 0; w < nch; ++w) {
            uint32_t bits = nzf[w];
            while (bits) {
              const int t0 = w * 32 + __ffs(bits) - 1;
              bits &= bits - 1;
              const bool two = bits != 0;
              const int t1 = two ? w * 32 + __ffs(bits) - 1 : t0;
              bits &= bits - 1;
              const float4 d0 = *reinterpret_cast<const float4*>(&S.hs[t0 * UPC + jq * 4]);
              const float4 d1 = *reinterpret_cast<const float4*>(&S.hs[t1 * UPC + jq * 4]);
              const float x0 = tab[bb[t0 * 52]];
              const float x1 = two ? tab[bb[t1 * 52]] : 0.f;
              acc.x = fmaf(x0, d0.x, acc.x);
              acc.y = fmaf(x0, d0.y, acc.y);
              acc.z = fmaf(x0, d0.z, acc.z);
              acc.w = fmaf(x0, d0.w, acc.w);
              acc2.x = fmaf(x1, d1.x, acc2.x);
              acc2.y = fmaf(x1, d1.y, acc2.y);
              acc2.z = fmaf(x1, d1.z, acc2.z);
              acc2.w = fmaf(x1, d1.w, acc2.w);
            }
          }
          acc.x += acc2.x;
          acc.y += acc2.y;
          acc.z += acc2.z;
          acc.w += acc2.w;
        } else if (f >= NF) {  // f == 198: b1, f == 199: w2 (fixed-order sum of the warp partials)
          for (int w = 0; w < NT / 32; ++w) {
            const float4 x = *reinterpret_cast<const float4*>(&S.pr[199 - f][w][jq * 4]);
            acc.x += x.x;
            acc.y += x.y;
            acc.z += x.z;
            acc.w += x.w;
          }
        }
        g[k] = acc;
        ss = fmaf(acc.x, acc.x, fmaf(acc.y, acc.y, fmaf(acc.z, acc.z, fmaf(acc.w, acc.w, ss))));
      }
      if (tid == 0 && rank == 0) ss = fmaf(msum[4], msum[4], ss);  // b2 gradient, counted once
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(BG_FULL, ss, o);
      if (lane == 0) S.red2[warp] = ss;
      __syncthreads();
      if (tid < CL) {
        float s = 0.f;
        for (int w = 0; w < NT / 32; ++w) s += S.red2[w];
        cluster.map_shared_rank(S.npart, tid)[rank] = s;
      }
      PH(6);
      cluster.sync();
      PH(7);

      // ---- F: clip_grad_norm_ (trainer.py:125-128) and the Adam step (:139) on this thread's slots ----
      float tot2 = 0.f;
      for (int c = 0; c < CL; ++c) tot2 += S.npart[c];
      const float total = sqrtf(tot2);
      float coef = 1.0f;
      if (a.grad_clip > 0.f) coef = fminf(1.0f, a.grad_clip / (total + 1e-6f));
      const float2 sc = S.sc[kstep & (NSC - 1)];
#pragma unroll
      for (int k = 0; k < KF; ++k) {
        if (fF[k] >= 0) {
          const int s = fF[k] * UPC + jq * 4;
          float4 p = *reinterpret_cast<float4*>(&S.P[s]), m = *reinterpret_cast<float4*>(&S.M[s]), v = *reinterpret_cast<float4*>(&S.V[s]);
          p.x = adam1(p.x, m.x, v.x, g[k].x * coef, sc.x, sc.y);
          p.y = adam1(p.y, m.y, v.y, g[k].y * coef, sc.x, sc.y);
          p.z = adam1(p.z, m.z, v.z, g[k].z * coef, sc.x, sc.y);
          p.w = adam1(p.w, m.w, v.w, g[k].w * coef, sc.x, sc.y);
          *reinterpret_cast<float4*>(&S.P[s]) = p;
          *reinterpret_cast<float4*>(&S.M[s]) = m;
          *reinterpret_cast<float4*>(&S.V[s]) = v;
        }
      }
      if (tid == 0) {
        S.b2[0] = adam1(S.b2[0], S.b2[1], S.b2[2], msum[4] * coef, sc.x, sc.y);
        if (rank == 0 && a.metrics) {
          float* mt = a.metrics + e * 6;
          mt[0] = msum[0] / (float)T;  // loss.item()
          mt[1] = msum[1] / (float)T;  // TD_error.abs().mean()
          mt[2] = total * coef;        // norm of the clipped gradients (trainer.py:131-136)
          mt[3] = msum[2] / (float)T;  // Y_values.mean()
          mt[4] = msum[3];             // rewards.sum()
          mt[5] = (float)T;
        }
      }
      kstep += 1;
      __syncthreads();
      if ((kstep & (NSC - 1)) == 0) {
        fill_scalars();
        __syncthreads();
      }
      PH(8);
      e = e_next;
      buf ^= 1;
    }
  }
#ifdef BG_LEARNER_PROFILE
  if (tid == 0 && rank == 0 && kstep)
    printf("k_td0_update cycles/episode: stage %lld decode %lld fwd %lld csync1 %lld targets %lld dz %lld gradW %lld csync2 %lld adam %lld head %lld\n",
           ph[0] / kstep, ph[1] / kstep, ph[2] / kstep, ph[3] / kstep, ph[4] / kstep, ph[5] / kstep, ph[6] / kstep, ph[7] / kstep,
           ph[8] / kstep, ph[9] / kstep);
#endif

  for (int s = tid; s < NROW * UPC; s += NT) {
    const int64_t gi = (int64_t)(s / UPC) * H + rank * UPC + (s % UPC);
    a.params[gi] = S.P[s];
    a.m[gi] = S.M[s];
    a.v[gi] = S.V[s];
  }
  if (rank == 0 && tid == 0) {
    a.params[iB2] = S.b2[0];
    a.m[iB2] = S.b2[1];
    a.v[iB2] = S.b2[2];
    *a.step = step0 + kstep;
  }
  cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

// ================================================================================================================
// H <= 128: the same algorithm with the two contractions on the tensor cores.
//
// Per CTA (16 hidden units) an episode is two small GEMMs: Z[T,16] = X[T,208] W[208,16] and G[208,16] = X^T[208,T] dZ[T,16].
// X (the 198 features, a ones column that carries b1, zero padding) is exact in bf16 -- every feature is 0, 1, k/2 or a borne-off
// count 0..15 whose 1/15 is folded into W -- and the fp32 operand of each product (W, dZ) is split into three bf16 pieces
// (hi + mid + lo = the fp32 value), so every tensor-core product is exact and the fp32 accumulation matches a CUDA-core FMA
// loop to rounding (parity with the reference trainer is tested to 1e-5 on every weight after 400 steps).  This replaces the
// sparse, divergent, latency-bound row/bitmask walks of k_td0_update with ~250 warp-level mma.sync per episode and CTA; the
// dense X tile costs one pass over the boards.  mma.sync (not tcgen05): the tiles are 16 units wide and the critical resource
// is latency inside a 200-step sequential chain, not tensor throughput.
// ================================================================================================================
constexpr int KP = 208;  // K of the forward / M of the backward GEMM: 198 features, ones column (198), zero padding
constexpr int XS = 216;  // X row stride in bf16 (432 B): the 8 row addresses of an ldmatrix fall in distinct 16-byte bank groups
constexpr int WS = 24;   // row stride of the split W / dZ operands in bf16 (48 B), same property
constexpr int RT = 128;  // rows per X tile; longer episodes take several tiles
constexpr int UT = 16;   // hidden units per CTA

struct SmemTC {
  alignas(16) float w2s[UT];   // value-head weights of this CTA's units (every other parameter lives in its owner's registers)
  float hs[TMAX * UT];         // sigmoid activations [t][unit]
  float ypart[MAXCL_TC * TMAX];
  float Y[TMAX], rew[TMAX];
  float2 sc[NSC];
  int64_t offs[EC + 1], lens[EC];
  alignas(16) float pr[NT / 32][UT];  // per-warp partials of the w2 gradient
  uint2 xtab[16];              // the four bf16 thermometer features of a point holding c checkers
  float npart[MAXCL_TC];
  float red[8][8];
  float red2[8];
  float b2[4];
  alignas(16) uint32_t brd[2][TMAX * 13];
  alignas(16) __nv_bfloat16 X[RT * XS];
  alignas(16) __nv_bfloat16 Ws[3][KP * WS];  // W1 (+ b1 in row 198), three bf16 pieces, [feature][unit]
  alignas(16) __nv_bfloat16 Ds[3][RT * WS];  // dL/dz of the current tile, three bf16 pieces, [row][unit]
  uint8_t flg[TMAX];
};
static_assert(sizeof(SmemTC) <= 232448, "k_td0_update_tc exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + mid + lo with three bf16 pieces (24 mantissa bits), for a pair, through the packed converts: p[0..2] = (hi, mid, lo) pieces of (x0, x1) as bf16x2 words
__device__ __forceinline__ void split3x2(float x0, float x1, uint32_t (&p)[3]) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(x0, x1);
    p[s] = *reinterpret_cast<const uint32_t*>(&b);
    x0 -= __uint_as_float(p[s] << 16);
    x1 -= __uint_as_float(p[s] & 0xffff0000u);
  }
}
template <int N>
__device__ __forceinline__ void store_split2(__nv_bfloat16 (*dst)[N], int idx, float x0, float x1) {
  uint32_t p[3];
  split3x2(x0, x1, p);
#pragma unroll
  for (int s = 0; s < 3; ++s) *reinterpret_cast<uint32_t*>(&dst[s][idx]) = p[s];
}

__global__ void __launch_bounds__(NT, 1) k_td0_update_tc(const LearnerArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CL = (int)cluster.num_blocks();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemTC& S = *reinterpret_cast<SmemTC*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;  // mma fragment coordinates: row group, column pair
  const int H = a.H;
  const int64_t iB2 = (int64_t)NROW * H;
  const float k15 = 1.0f / 15.0f;

  // Every parameter of the slice is owned by exactly one thread for the whole launch -- the thread whose mma accumulator fragment
  // receives its gradient: slot (i, hh, nt, k) is feature row f = (warp + 8 i) * 16 + g + 8 hh, unit nt * 8 + 2 q + k.  Weight and
  // both Adam moments stay in that thread's registers; shared memory only holds the bf16 operand image of W1/b1 and w2.
  float pW[2][2][2][2], pM[2][2][2][2], pV[2][2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int f = (warp + 8 * i) * 16 + g + 8 * hh;
          const int64_t gi = (int64_t)f * H + rank * UT + nt * 8 + 2 * q + k;
          const bool own = f < NROW;
          pW[i][hh][nt][k] = own ? a.params[gi] : 0.f;
          pM[i][hh][nt][k] = own ? a.m[gi] : 0.f;
          pV[i][hh][nt][k] = own ? a.v[gi] : 0.f;
        }
  if (tid == 0) {
    S.b2[0] = a.params[iB2];
    S.b2[1] = a.m[iB2];
    S.b2[2] = a.v[iB2];
  }
  for (int i = tid; i < 3 * KP * WS / 2; i += NT) reinterpret_cast<uint32_t*>(&S.Ws[0][0])[i] = 0u;
  if (tid < 16) {
    const uint32_t one = 0x3F80u;  // bf16 1.0
    const uint32_t ex = tid > 3 ? (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(0.5f * (float)(tid - 3))) : 0u;
    S.xtab[tid] = make_uint2((tid > 0 ? one : 0u) | (tid > 1 ? one << 16 : 0u), (tid > 2 ? one : 0u) | (ex << 16));
  }
  const int64_t step0 = *a.step;
  int64_t kstep = 0;
  auto fill_scalars = [&]() {
    const double n = (double)(step0 + kstep + tid + 1);
    S.sc[tid] = make_float2((float)((double)a.lr / (1.0 - pow(0.9, n))), 1.0f / (float)sqrt(1.0 - pow(0.999, n)));
  };
  fill_scalars();
  __syncthreads();
  // operand image of the weights: rows 0..197 = W1^T (the two borne-off rows scaled by 1/15: X holds the count), row 198 = b1;
  // row 199 (w2) is not a GEMM operand, it goes to w2s
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int f = (warp + 8 * i) * 16 + g + 8 * hh;
        const int n = nt * 8 + 2 * q;
        if (f <= NF) {
          const float wsc = (f == 193 || f == 195) ? k15 : 1.0f;
          store_split2(S.Ws, f * WS + n, pW[i][hh][nt][0] * wsc, pW[i][hh][nt][1] * wsc);
        } else if (f == NF + 1) {
          *reinterpret_cast<float2*>(&S.w2s[n]) = make_float2(pW[i][hh][nt][0], pW[i][hh][nt][1]);
        }
      }
  __syncthreads();
  cluster.sync();

  const uint32_t* gb32 = reinterpret_cast<const uint32_t*>(a.boards);
  int64_t c0 = 0, c1 = 0;
  auto next_valid = [&](int64_t e) {
    for (; e < c1; ++e) {
      const int64_t Tl = S.lens[e - c0];
      if (Tl > 0 && Tl <= TMAX) break;
      if (rank == 0 && tid == 0) {
        if (Tl > TMAX && a.status) *a.status = BG_ERR_CAPACITY;
        if (a.metrics)
          for (int k = 0; k < 6; ++k) a.metrics[e * 6 + k] = 0.0f;
      }
    }
    return e;
  };
  uint8_t pf_flag[2] = {0, 0};
  float pf_rew[2] = {0.f, 0.f};
  auto prefetch = [&](int64_t e, int buf) {
    const int64_t lo = S.offs[e - c0];
    const int T = (int)S.lens[e - c0];
    for (int w = tid; w < T * 13; w += NT) {
      const int t = w / 13, k = w - t * 13;
      if (a.records && t == 0)
        S.brd[buf][w] = initial_board_word(k);
      else
        cp_async4(&S.brd[buf][w], gb32 + (a.records ? lo + t - 1 : lo + t) * 13 + k);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int t = tid + i * NT;
      if (t < T) {
        pf_flag[i] = a.flags[lo + t] & 1;
        pf_rew[i] = a.reward[lo + t];
      }
    }
  };

#ifdef BG_LEARNER_PROFILE
  long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long t_ = clock64();
#endif
  for (c0 = 0; c0 < a.n_eps; c0 += EC) {
    c1 = c0 + EC < a.n_eps ? c0 + EC : a.n_eps;
    __syncthreads();
    for (int i = tid; i < (int)(c1 - c0); i += NT) {  // episode e: rows [offs[e], offs[e] + len); len from ep_len or the next CSR offset
      S.offs[i] = a.ep_offsets[c0 + i];
      S.lens[i] = a.ep_len ? (int64_t)a.ep_len[c0 + i] : a.ep_offsets[c0 + i + 1] - a.ep_offsets[c0 + i];
    }
    __syncthreads();
    int buf = 0;
    int64_t e = next_valid(c0);
    if (e < c1) prefetch(e, buf);
    while (e < c1) {
      PH(9);
      const int T = (int)S.lens[e - c0];
      const uint8_t* brd8 = reinterpret_cast<const uint8_t*>(S.brd[buf]);

      // ---- A: land the staged episode; start fetching the next one ----
      cp_async_wait_all();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int t = tid + i * NT;
        if (t < T) {
          S.flg[t] = pf_flag[i];
          S.rew[t] = pf_rew[i];
        }
      }
      __syncthreads();
      const int64_t e_next = next_valid(e + 1);
      if (e_next < c1) prefetch(e_next, buf ^ 1);
      PH(0);

      // dense bf16 feature rows of the tile starting at row tb (rows >= T are zero); returns the tile's row count (multiple of 16)
      auto decode_tile = [&](int tb) {
        const int rows = min(RT, ((T - tb) + 15) & ~15);
        for (int i = tid; i < rows * 12; i += NT) {  // (row, board word): four points = 16 thermometer features, two 16-byte stores
          const int tr = i / 12, w = i - tr * 12;
          const int t = tb + tr;
          const uint32_t c4 = t < T ? S.brd[buf][t * 13 + w] : 0u;
          const uint2 v0 = S.xtab[c4 & 15u], v1 = S.xtab[(c4 >> 8) & 15u], v2 = S.xtab[(c4 >> 16) & 15u], v3 = S.xtab[(c4 >> 24) & 15u];
          uint4* dst = reinterpret_cast<uint4*>(&S.X[tr * XS + w * 16]);
          dst[0] = make_uint4(v0.x, v0.y, v1.x, v1.y);
          dst[1] = make_uint4(v2.x, v2.y, v3.x, v3.y);
        }
        for (int tr = tid; tr < rows; tr += NT) {  // bar/2, off count, flags, ones column, zero padding
          const int t = tb + tr;
          const bool vd = t < T;
          const uint8_t* b = brd8 + t * 52;
          auto bf = [](float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); };
          uint4 v;
          v.x = vd ? bf(0.5f * (float)b[48]) | (bf((float)b[50]) << 16) : 0u;
          v.y = vd ? bf(0.5f * (float)b[49]) | (bf((float)b[51]) << 16) : 0u;
          v.z = vd ? (S.flg[t] == 0 ? 0x3F80u : 0x3F80u << 16) : 0u;
          v.w = vd ? 0x3F80u : 0u;
          *reinterpret_cast<uint4*>(&S.X[tr * XS + 192]) = v;
          *reinterpret_cast<uint4*>(&S.X[tr * XS + 200]) = make_uint4(0u, 0u, 0u, 0u);
        }
        return rows;
      };

      // ---- B: forward, tile by tile: Z = X W on the tensor cores (three bf16 pieces of W), sigmoid, value partials to every CTA ----
      const float2 w2a = *reinterpret_cast<const float2*>(&S.w2s[2 * q]);      // units 2q, 2q+1
      const float2 w2b = *reinterpret_cast<const float2*>(&S.w2s[8 + 2 * q]);  // units 8+2q, 9+2q
      for (int tb = 0; tb < T; tb += RT) {
        if (tb) __syncthreads();  // previous tile's X fully consumed
        const int rows = decode_tile(tb);
        __syncthreads();
        PH(1);
        if (warp * 16 < rows) {
          float ac3[3][2][4];  // one accumulator chain per bf16 piece and unit half: six independent mma chains
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) ac3[s][nt][0] = ac3[s][nt][1] = ac3[s][nt][2] = ac3[s][nt][3] = 0.f;
          const __nv_bfloat16* xa = &S.X[(warp * 16 + (lane & 15)) * XS + (lane >> 4) * 8];
          const int brow = ((lane >> 3) & 1) * 8 + (lane & 7), bcol = (lane >> 4) * 8;
#pragma unroll
          for (int ks = 0; ks < KP / 16; ++ks) {
            uint32_t af[4];
            ldsm_x4(af, xa + ks * 16);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              uint32_t bfr[4];
              ldsm_x4_t(bfr, &S.Ws[s][(ks * 16 + brow) * WS + bcol]);
              mma_bf16(ac3[s][0], af, bfr[0], bfr[1]);
              mma_bf16(ac3[s][1], af, bfr[2], bfr[3]);
            }
          }
          float acc[2][4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[nt][k] = ac3[0][nt][k] + (ac3[1][nt][k] + ac3[2][nt][k]);  // small pieces first
          // C fragment: rows g and g+8 of the warp's 16, units 2q,2q+1 (acc[0]) and 8+2q,9+2q (acc[1])
          const int tA = tb + warp * 16 + g, tB = tA + 8;
          float h[2][4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int k = 0; k < 4; ++k) h[nt][k] = __fdividef(1.0f, 1.0f + __expf(-acc[nt][k]));  // as in k_eval (|dV| <= 1e-5 holds)
          float pA = fmaf(w2b.y, h[1][1], fmaf(w2b.x, h[1][0], fmaf(w2a.y, h[0][1], w2a.x * h[0][0])));
          float pB = fmaf(w2b.y, h[1][3], fmaf(w2b.x, h[1][2], fmaf(w2a.y, h[0][3], w2a.x * h[0][2])));
          pA += __shfl_xor_sync(BG_FULL, pA, 1);
          pB += __shfl_xor_sync(BG_FULL, pB, 1);
          pA += __shfl_xor_sync(BG_FULL, pA, 2);
          pB += __shfl_xor_sync(BG_FULL, pB, 2);
          if (tA < T) {
            *reinterpret_cast<float2*>(&S.hs[tA * UT + 2 * q]) = make_float2(h[0][0], h[0][1]);
            *reinterpret_cast<float2*>(&S.hs[tA * UT + 8 + 2 * q]) = make_float2(h[1][0], h[1][1]);
          }
          if (tB < T) {
            *reinterpret_cast<float2*>(&S.hs[tB * UT + 2 * q]) = make_float2(h[0][2], h[0][3]);
            *reinterpret_cast<float2*>(&S.hs[tB * UT + 8 + 2 * q]) = make_float2(h[1][2], h[1][3]);
          }
          for (int c = q; c < CL; c += 4) {
            float* dst = cluster.map_shared_rank(S.ypart, c) + rank * TMAX;
            if (tA < T) dst[tA] = pA;
            if (tB < T) dst[tB] = pB;
          }
        }
      }
      PH(2);
      cluster.sync();
      PH(3);

      // ---- C: values ----
      for (int t = tid; t < T; t += NT) {
        float y = S.b2[0];
        for (int c = 0; c < CL; ++c) y += S.ypart[c * TMAX + t];
        S.Y[t] = y;
      }
      __syncthreads();
      PH(4);

      // ---- D + E: per tile, dL/dz split into three bf16 pieces, then G += X^T dZ on the tensor cores.  Warp w owns the feature
      //             m-tiles w and w + 8 (16 features each) for both 8-unit halves; the result stays in registers for Adam ----
      float G3[3][2][2][4];  // [piece][m-tile slot][unit half][fragment]
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) G3[s][i][nt][0] = G3[s][i][nt][1] = G3[s][i][nt][2] = G3[s][i][nt][3] = 0.f;
      float2 gw2 = make_float2(0.f, 0.f);  // this thread's unit pair (2 * (tid % 8)), summed over its rows
      const int np = (tid & 7) * 2;
      const float2 w2p = *reinterpret_cast<const float2*>(&S.w2s[np]);
      float r5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // loss, |td|, value, reward, dL/dY sums (counted by the np == 0 thread of each row)
      for (int tb = 0; tb < T; tb += RT) {
        int rows = min(RT, ((T - tb) + 15) & ~15);
        if (T > RT) {  // the forward pass left another tile in X
          __syncthreads();
          rows = decode_tile(tb);
        }
        // TD(0) target (trainer.py:110-115), dL/dY of the mse loss (:118), dL/dz and its three bf16 pieces, one (row, unit pair) per thread
        for (int i = tid; i < rows * 8; i += NT) {
          const int tr = i >> 3, t = tb + tr;
          float2 dz = make_float2(0.f, 0.f);
          if (t < T) {
            const float y = S.Y[t], rw = S.rew[t];
            float tg = rw;
            if (t + 1 < T) tg = __fadd_rn(tg, __fmul_rn(a.gamma, S.Y[t + 1]));
            const float d = y - tg;
            const float dy = 2.0f * d / (float)T;
            if (np == 0) {
              r5[0] = fmaf(d, d, r5[0]);
              r5[1] += fabsf(d);
              r5[2] += y;
              r5[3] += rw;
              r5[4] += dy;
            }
            const float2 h = *reinterpret_cast<const float2*>(&S.hs[t * UT + np]);
            gw2.x = fmaf(dy, h.x, gw2.x);
            gw2.y = fmaf(dy, h.y, gw2.y);
            dz.x = dy * w2p.x * h.x * (1.0f - h.x);
            dz.y = dy * w2p.y * h.y * (1.0f - h.y);
          }
          store_split2(S.Ds, tr * WS + np, dz.x, dz.y);
        }
        __syncthreads();
        PH(5);
        const int brow = ((lane >> 3) & 1) * 8 + (lane & 7), bcol = (lane >> 4) * 8;
        // A = X^T: stored [row t][feature]; sub-matrices (features 0-7 | 8-15) x (rows 0-7 | 8-15), loaded transposed.  The dZ
        // fragments of a k-step are shared by the warp's two m-tiles.
        const __nv_bfloat16* xa0 = &S.X[((lane >> 4) * 8 + (lane & 7)) * XS + warp * 16 + ((lane >> 3) & 1) * 8];
        const bool two = warp + 8 < KP / 16;
#pragma unroll 2
        for (int ks = 0; ks < rows / 16; ++ks) {
          uint32_t af0[4], af1[4] = {0u, 0u, 0u, 0u};
          ldsm_x4_t(af0, xa0 + ks * 16 * XS);
          if (two) ldsm_x4_t(af1, xa0 + ks * 16 * XS + 8 * 16);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            uint32_t bfr[4];
            ldsm_x4_t(bfr, &S.Ds[s][(ks * 16 + brow) * WS + bcol]);
            mma_bf16(G3[s][0][0], af0, bfr[0], bfr[1]);
            mma_bf16(G3[s][0][1], af0, bfr[2], bfr[3]);
            if (two) {
              mma_bf16(G3[s][1][0], af1, bfr[0], bfr[1]);
              mma_bf16(G3[s][1][1], af1, bfr[2], bfr[3]);
            }
          }
        }
      }
      // w2 gradient: reduce the unit-pair partials over the lanes that share tid % 8, then over the warps
      gw2.x += __shfl_xor_sync(BG_FULL, gw2.x, 8);
      gw2.y += __shfl_xor_sync(BG_FULL, gw2.y, 8);
      gw2.x += __shfl_xor_sync(BG_FULL, gw2.x, 16);
      gw2.y += __shfl_xor_sync(BG_FULL, gw2.y, 16);
      if (lane < 8) *reinterpret_cast<float2*>(&S.pr[warp][np]) = gw2;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r5[k] += __shfl_xor_sync(BG_FULL, r5[k], o);
      }
      if (lane == 0)
        for (int k = 0; k < 5; ++k) S.red[warp][k] = r5[k];
      __syncthreads();
      float msum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (tid == 0) {
        for (int k = 0; k < 5; ++k) {
          float sres = 0.f;
          for (int w = 0; w < NT / 32; ++w) sres += S.red[w][k];
          msum[k] = sres;
        }
      }
      // fragment ownership: G[i][nt][2 * hh + k] is feature f = (warp + 8 i) * 16 + g + 8 hh, unit nt * 8 + 2 q + k
      float G[2][2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int k = 0; k < 4; ++k) G[i][nt][k] = G3[0][i][nt][k] + (G3[1][i][nt][k] + G3[2][i][nt][k]);  // small pieces first
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int f = (warp + 8 * i) * 16 + g + 8 * hh;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              float x = G[i][nt][2 * hh + k];
              if (f == 193 || f == 195) x *= k15;
              if (f == NF + 1) {  // w2: fixed-order sum of the warp partials
                x = 0.f;
                for (int w = 0; w < NT / 32; ++w) x += S.pr[w][nt * 8 + 2 * q + k];
              }
              if (f >= NROW) x = 0.f;
              G[i][nt][2 * hh + k] = x;
              ss = fmaf(x, x, ss);
            }
        }
      if (tid == 0 && rank == 0) ss = fmaf(msum[4], msum[4], ss);  // b2 gradient, counted once
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(BG_FULL, ss, o);
      if (lane == 0) S.red2[warp] = ss;
      __syncthreads();
      if (tid < CL) {
        float s = 0.f;
        for (int w = 0; w < NT / 32; ++w) s += S.red2[w];
        cluster.map_shared_rank(S.npart, tid)[rank] = s;
      }
      PH(6);
      cluster.sync();
      PH(7);

      // ---- F: clip_grad_norm_ (trainer.py:125-128), the Adam step (:139) and the new operand image of the touched weights ----
      float tot2 = 0.f;
      for (int c = 0; c < CL; ++c) tot2 += S.npart[c];
      const float total = sqrtf(tot2);
      float coef = 1.0f;
      if (a.grad_clip > 0.f) coef = fminf(1.0f, a.grad_clip / (total + 1e-6f));
      const float2 sc = S.sc[kstep & (NSC - 1)];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int f = (warp + 8 * i) * 16 + g + 8 * hh;
          if (f < NROW) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              float(&w)[2] = pW[i][hh][nt];
              w[0] = adam1(w[0], pM[i][hh][nt][0], pV[i][hh][nt][0], G[i][nt][2 * hh] * coef, sc.x, sc.y);
              w[1] = adam1(w[1], pM[i][hh][nt][1], pV[i][hh][nt][1], G[i][nt][2 * hh + 1] * coef, sc.x, sc.y);
              const int n = nt * 8 + 2 * q;
              if (f <= NF) {  // the thread that owns a weight rewrites its three bf16 pieces
                const float wsc = (f == 193 || f == 195) ? k15 : 1.0f;
                store_split2(S.Ws, f * WS + n, w[0] * wsc, w[1] * wsc);
              } else {
                *reinterpret_cast<float2*>(&S.w2s[n]) = make_float2(w[0], w[1]);
              }
            }
          }
        }
      if (tid == 0) {
        S.b2[0] = adam1(S.b2[0], S.b2[1], S.b2[2], msum[4] * coef, sc.x, sc.y);
        if (rank == 0 && a.metrics) {
          float* mt = a.metrics + e * 6;
          mt[0] = msum[0] / (float)T;
          mt[1] = msum[1] / (float)T;
          mt[2] = total * coef;
          mt[3] = msum[2] / (float)T;
          mt[4] = msum[3];
          mt[5] = (float)T;
        }
      }
      kstep += 1;
      __syncthreads();
      if ((kstep & (NSC - 1)) == 0) {
        fill_scalars();
        __syncthreads();
      }
      PH(8);
      e = e_next;
      buf ^= 1;
    }
  }
#ifdef BG_LEARNER_PROFILE
  if (tid == 0 && rank == 0 && kstep)
    printf("k_td0_update_tc cycles/episode: stage %lld decode %lld fwd %lld csync1 %lld targets %lld dzsplit %lld gradW %lld csync2 %lld adam %lld head %lld\n",
           ph[0] / kstep, ph[1] / kstep, ph[2] / kstep, ph[3] / kstep, ph[4] / kstep, ph[5] / kstep, ph[6] / kstep, ph[7] / kstep,
           ph[8] / kstep, ph[9] / kstep);
#endif

#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int f = (warp + 8 * i) * 16 + g + 8 * hh;
          if (f < NROW) {
            const int64_t gi = (int64_t)f * H + rank * UT + nt * 8 + 2 * q + k;
            a.params[gi] = pW[i][hh][nt][k];
            a.m[gi] = pM[i][hh][nt][k];
            a.v[gi] = pV[i][hh][nt][k];
          }
        }
  if (rank == 0 && tid == 0) {
    a.params[iB2] = S.b2[0];
    a.m[iB2] = S.b2[1];
    a.v[iB2] = S.b2[2];
    *a.step = step0 + kstep;
  }
  cluster.sync();
}

struct OptScalars {
  int64_t step;
};

}  // namespace

struct Learner {
  int32_t device;
  int32_t H;
  int64_t n_params;
  float lr, gamma, grad_clip;
  float *params, *m, *v;
  OptScalars* opt;
  int path;  // 0 undecided, 1 tensor-core kernel (16 units per CTA), 2 sparse CUDA-core kernel (32 units per CTA)
};

int32_t learner_create(Learner** out, int32_t device, int32_t H, float lr, float gamma, float grad_clip) {
  if (H < 32 || H > 256 || H % 32) {
    set_error("bg_learner_create: H must be a multiple of 32 in [32,256]");
    return BG_ERR_ARG;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return check_cuda(e, "cudaSetDevice");
  Learner* L = new (std::nothrow) Learner();
  if (!L) return BG_ERR_ARG;
  L->device = device;
  L->H = H;
  L->n_params = (int64_t)NROW * H + 1;
  L->lr = lr;
  L->gamma = gamma;
  L->grad_clip = grad_clip;
  L->path = 0;
  float* buf = nullptr;
  e = cudaMalloc(&buf, sizeof(float) * 3 * (size_t)L->n_params + sizeof(OptScalars) + 16);
  if (e != cudaSuccess) {
    delete L;
    return check_cuda(e, "cudaMalloc(learner)");
  }
  L->params = buf;
  L->m = buf + L->n_params;
  L->v = buf + 2 * L->n_params;
  L->opt = reinterpret_cast<OptScalars*>((reinterpret_cast<uintptr_t>(buf + 3 * L->n_params) + 15) & ~(uintptr_t)15);
  e = cudaMemset(buf, 0, sizeof(float) * 3 * (size_t)L->n_params + sizeof(OptScalars) + 16);
  const OptScalars init{0};
  if (e == cudaSuccess) e = cudaMemcpy(L->opt, &init, sizeof(init), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(buf);
    delete L;
    return check_cuda(e, "learner init");
  }
  *out = L;
  return BG_OK;
}

int learner_device(const Learner* L) { return L ? L->device : 0; }

int32_t learner_destroy(Learner* L) {
  if (!L) return BG_OK;
  cudaSetDevice(L->device);
  cudaFree(L->params);
  delete L;
  return BG_OK;
}

int32_t learner_set_parameters(Learner* L, const float* packed_dev, int32_t reset_optimizer, cudaStream_t s) {
  cudaError_t e = cudaSetDevice(L->device);
  if (e == cudaSuccess) e = cudaMemcpyAsync(L->params, packed_dev, sizeof(float) * (size_t)L->n_params, cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess && reset_optimizer) {
    e = cudaMemsetAsync(L->m, 0, sizeof(float) * 2 * (size_t)L->n_params, s);
    static const OptScalars init{0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(L->opt, &init, sizeof(init), cudaMemcpyHostToDevice, s);
  }
  return check_cuda(e, "bg_learner_set_parameters");
}

int32_t learner_get_parameters(Learner* L, float* packed_dev, cudaStream_t s) {
  cudaError_t e = cudaSetDevice(L->device);
  if (e == cudaSuccess) e = cudaMemcpyAsync(packed_dev, L->params, sizeof(float) * (size_t)L->n_params, cudaMemcpyDeviceToDevice, s);
  return check_cuda(e, "bg_learner_get_parameters");
}

int32_t learner_get_optimizer(Learner* L, float* m_dev, float* v_dev, int64_t* step_dev, cudaStream_t s) {
  cudaError_t e = cudaSetDevice(L->device);
  const size_t nb = sizeof(float) * (size_t)L->n_params;
  if (e == cudaSuccess && m_dev) e = cudaMemcpyAsync(m_dev, L->m, nb, cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess && v_dev) e = cudaMemcpyAsync(v_dev, L->v, nb, cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess && step_dev) e = cudaMemcpyAsync(step_dev, &L->opt->step, sizeof(int64_t), cudaMemcpyDeviceToDevice, s);
  return check_cuda(e, "bg_learner_get_optimizer");
}

template <typename K>
static void fill_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, size_t smem, unsigned CL, cudaStream_t s) {
  cfg = {};
  cfg.gridDim = dim3(CL);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
}

template <typename K>
static int32_t launch_update(K kernel, size_t smem, unsigned CL, const LearnerArgs& a, cudaStream_t s) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  fill_config<K>(cfg, at, smem, CL, s);
  return check_cuda(cudaLaunchKernelEx(&cfg, kernel, a), "k_td0_update launch");
}

// One-time choice of the kernel.  H <= 128: tensor-core kernel, portable cluster of H / 16 CTAs.  H > 128: the same kernel needs
// 10-16 CTAs per cluster (non-portable size, allowed on B200 when a GPC can co-schedule them); if the device cannot place such a
// cluster the CUDA-core kernel with 32 units per CTA (cluster of H / 32 <= 8) is used instead.
static int32_t choose_path(Learner* L) {
  if (L->path) return BG_OK;
  cudaError_t e = cudaFuncSetAttribute(k_td0_update_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemTC));
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_td0_update_tc)");
  e = cudaFuncSetAttribute(k_td0_update<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<32>));
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_td0_update)");
  const unsigned CL = (unsigned)(L->H / UT);
  L->path = 1;
  const char* force = getenv("BG_LEARNER_PATH");  // "cuda-core": force the 32-unit kernel (tests; only meaningful for H > 128)
  if (CL > 8) {
    L->path = 2;
    if (force && !strcmp(force, "cuda-core")) return BG_OK;
    e = cudaFuncSetAttribute(k_td0_update_tc, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute at[1];
      fill_config<decltype(&k_td0_update_tc)>(cfg, at, sizeof(SmemTC), CL, nullptr);
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, k_td0_update_tc, &cfg);
      if (e == cudaSuccess && n >= 1) L->path = 1;
    }
    cudaGetLastError();  // a refused opt-in is not an error: the 32-unit kernel handles these sizes
  }
  return BG_OK;
}

int32_t learner_update(Learner* L, const int8_t* boards, const uint8_t* flags_or_meta, const float* reward, const int64_t* ep_offsets,
                       const int32_t* ep_len, int64_t n_eps, int32_t records, float* out_metrics, int32_t* out_status, cudaStream_t s) {
  cudaError_t e = cudaSetDevice(L->device);
  if (e != cudaSuccess) return check_cuda(e, "cudaSetDevice");
  if ((reinterpret_cast<uintptr_t>(boards) & 3) != 0) {
    set_error("bg_learner_update: boards must be 4-byte aligned");
    return BG_ERR_ARG;
  }
  if (out_status) {
    e = cudaMemsetAsync(out_status, 0, sizeof(int32_t), s);
    if (e != cudaSuccess) return check_cuda(e, "cudaMemsetAsync(status)");
  }
  if (n_eps == 0) return BG_OK;
  LearnerArgs a{boards, flags_or_meta, reward, ep_offsets, ep_len, n_eps, records, L->params, L->m, L->v, &L->opt->step,
                L->lr, L->gamma, L->grad_clip, out_metrics, out_status, L->H};
  int32_t rc = choose_path(L);
  if (rc != BG_OK) return rc;
  if (L->path == 1) return launch_update(k_td0_update_tc, sizeof(SmemTC), (unsigned)(L->H / UT), a, s);
  return launch_update(k_td0_update<32>, sizeof(Smem<32>), (unsigned)(L->H / 32), a, s);
}

}  // namespace bg
