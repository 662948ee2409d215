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
#include <math.h>
#include <stdio.h>

#include <new>

#include "learner.cuh"

namespace cg = cooperative_groups;

namespace bg {
namespace {

constexpr int NF = 198;
constexpr int NROW = 200;  // per hidden unit: 198 fc1 weights, b1, w2 (row f of the packed [W1t | b1 | w2] matrix)
constexpr int TMAX = 320;  // experiences per episode the kernel accepts (reference MAX_TIMESTEPS = 300)
constexpr int LHALF = 20;  // sparse-list capacity per (row, player half): <= 15 point features + bar + off + flag = 18
constexpr int NT = 256;
constexpr int MAXCL = 8;

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
constexpr int FI = 4;     // rows in flight per thread in the forward pass

template <int UPC>
struct Smem {
  float P[NROW * UPC], M[NROW * UPC], V[NROW * UPC], G[NROW * UPC];
  float hs[TMAX * UPC];        // sigmoid activations, then dL/dz
  float ypart[MAXCL * TMAX];   // ypart[c][t]: CTA c's partial of V(x_t), pushed by CTA c
  float Y[TMAX], dY[TMAX], rew[TMAX];
  float pr[2][NT];             // per-thread partials of the w2 / b1 gradients
  float2 sc[NSC];              // (lr / bias_correction1, sqrt(bias_correction2)) for the next NSC optimiser steps
  int64_t offs[EC + 1];
  float npart[MAXCL];          // per-CTA sums of squared gradients, pushed by each CTA
  float red[8][8];
  float red2[8];
  float vtab[32];              // feature value codes: 0 -> 1, k (1..15) -> k/2, 16+k -> k/15
  float extab[16];             // value of a point's 4th feature by checker count: max(c - 3, 0) / 2
  float b2[4];                 // b2 and its Adam moments (replicated in every CTA, updated identically)
  uint32_t brd[2][TMAX * 13];  // observation boards, double buffered: episode e+1 is fetched while e is processed
  uint16_t list[TMAX * 2 * LHALF];  // sparse rows: feature index | value code << 8, one sub-list per player half
  uint8_t lcnt[TMAX * 2];
  uint8_t flg[TMAX];
};

static_assert(sizeof(Smem<32>) <= 232448, "k_td0_update<32> exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ void adam_step(float& p, float& m, float& v, float g, float step_size, float bc2_sqrt) {
  // torch/optim/adam.py _single_tensor_adam: lerp, mul + addcmul, sqrt / bias_correction2_sqrt + eps, addcdiv
  m = m + 0.1f * (g - m);
  v = v * 0.999f + 0.001f * g * g;
  const float denom = __fdividef(sqrtf(v), bc2_sqrt) + 1e-8f;
  p = p - step_size * __fdividef(m, denom);
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int UPC>
__global__ void __launch_bounds__(NT, 1) k_td0_update(const LearnerArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CL = (int)cluster.num_blocks();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<UPC>& S = *reinterpret_cast<Smem<UPC>*>(smem_raw);
  constexpr int RL = NT / UPC;   // row lanes (forward) / point lanes (backward)
  constexpr int NPT = 48 / RL;   // board points per thread in the fc1 gradient
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = tid % UPC, r = tid / UPC;
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
    if (tid < 16) S.extab[tid] = tid > 3 ? 0.5f * (float)(tid - 3) : 0.0f;
  }
  const int64_t step0 = *a.step;
  int64_t kstep = 0;  // optimiser steps taken by this launch (thread-uniform)
  // torch/optim/adam.py: bias_correction = 1 - beta ** step in double; step_size = lr / bc1; bias_correction2_sqrt = sqrt(bc2)
  auto fill_scalars = [&]() {
    const double n = (double)(step0 + kstep + tid + 1);
    S.sc[tid] = make_float2((float)((double)a.lr / (1.0 - pow(0.9, n))), (float)sqrt(1.0 - pow(0.999, n)));
  };
  fill_scalars();
  __syncthreads();
  cluster.sync();  // every CTA of the cluster is resident before the first DSMEM store

  const uint32_t* gb32 = reinterpret_cast<const uint32_t*>(a.boards);

  // first episode at or after e (within the staged chunk [c0, c1)) that takes an optimiser step; empty and over-long ones are
  // reported and skipped.  Uniform over the cluster.
  int64_t c0 = 0, c1 = 0;
  auto next_valid = [&](int64_t e) {
    for (; e < c1; ++e) {
      const int64_t Tl = S.offs[e - c0 + 1] - S.offs[e - c0];
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
    const int T = (int)(S.offs[e - c0 + 1] - lo);
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
    for (int i = tid; i <= (int)(c1 - c0); i += NT) S.offs[i] = a.ep_offsets[c0 + i];
    __syncthreads();
    int buf = 0;
    int64_t e = next_valid(c0);
    if (e < c1) prefetch(e, buf);
    while (e < c1) {
      PH(9);
      const int T = (int)(S.offs[e - c0 + 1] - S.offs[e - c0]);
      const int8_t* brd8 = reinterpret_cast<const int8_t*>(S.brd[buf]);

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
      // sparse feature rows for the forward pass: one (row, player half) task per thread
      for (int task = tid; task < 2 * T; task += NT) {
        const int t = task >> 1, p = task & 1;
        uint16_t* li = S.list + task * LHALF;
        const int8_t* b = brd8 + t * 52;
        int cnt = 0;
        for (int pt = 0; pt < 24; ++pt) {
          const int c = b[p * 24 + pt];
          if (c > 0) {
            const int f0 = p * 96 + pt * 4;
            li[cnt++] = (uint16_t)f0;
            if (c > 1) li[cnt++] = (uint16_t)(f0 + 1);
            if (c > 2) li[cnt++] = (uint16_t)(f0 + 2);
            if (c > 3) li[cnt++] = (uint16_t)((f0 + 3) | ((c - 3) << 8));
          }
        }
        const int bar = b[48 + p], off = b[50 + p];
        if (bar > 0) li[cnt++] = (uint16_t)((192 + 2 * p) | (bar << 8));
        if (off > 0) li[cnt++] = (uint16_t)((193 + 2 * p) | ((16 + off) << 8));
        if (S.flg[t] == p) li[cnt++] = (uint16_t)(196 + p);
        S.lcnt[task] = (uint8_t)cnt;
      }
      __syncthreads();
      PH(1);

      // ---- B: forward for this CTA's UPC hidden units (FI rows in flight per thread); push the value partials to every CTA ----
      {
        const float b1j = S.P[NF * UPC + j], w2j = S.P[(NF + 1) * UPC + j];
        for (int t0 = 0; t0 < T; t0 += FI * RL) {
          int tr[FI];
          float z[FI];
#pragma unroll
          for (int i = 0; i < FI; ++i) {
            tr[i] = t0 + r + i * RL;
            z[i] = b1j;
          }
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            int n[FI], nmax = 0;
            const uint16_t* l[FI];
#pragma unroll
            for (int i = 0; i < FI; ++i) {
              const bool vd = tr[i] < T;
              n[i] = vd ? S.lcnt[tr[i] * 2 + p] : 0;
              l[i] = S.list + ((vd ? tr[i] : 0) * 2 + p) * LHALF;
              nmax = max(nmax, n[i]);
            }
            for (int q = 0; q < nmax; ++q) {
              // unconditional loads (a stale entry past a row's count stays inside the shared-memory block) + select: the FI
              // chains overlap instead of serialising behind predicated branches
              uint32_t en[FI];
              float w[FI], x[FI];
#pragma unroll
              for (int i = 0; i < FI; ++i) en[i] = l[i][q];
#pragma unroll
              for (int i = 0; i < FI; ++i) {
                w[i] = S.P[(en[i] & 255u) * UPC + j];
                x[i] = S.vtab[(en[i] >> 8) & 31u];
              }
#pragma unroll
              for (int i = 0; i < FI; ++i) z[i] = q < n[i] ? fmaf(x[i], w[i], z[i]) : z[i];
            }
          }
#pragma unroll
          for (int i = 0; i < FI; ++i) {
            const bool vd = tr[i] < T;
            float part = 0.0f;
            if (vd) {
              const float h = 1.0f / (1.0f + expf(-z[i]));
              S.hs[tr[i] * UPC + j] = h;
              part = w2j * h;
            }
#pragma unroll
            for (int o = UPC / 2; o > 0; o >>= 1) part += __shfl_xor_sync(BG_FULL, part, o);
            if (vd && j < CL) cluster.map_shared_rank(S.ypart, j)[rank * TMAX + tr[i]] = part;
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

      // ---- D: dL/dz in place of the activations; w2 / b1 gradient partials per row lane ----
      {
        float gw2 = 0.f, gb1 = 0.f;
        const float w2j = S.P[(NF + 1) * UPC + j];
#pragma unroll 4
        for (int t = r; t < T; t += RL) {
          const float h = S.hs[t * UPC + j], dy = S.dY[t];
          gw2 = fmaf(dy, h, gw2);
          const float dz = dy * w2j * h * (1.0f - h);
          S.hs[t * UPC + j] = dz;
          gb1 += dz;
        }
        S.pr[0][tid] = gw2;
        S.pr[1][tid] = gb1;
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

      // ---- E: fc1 gradient, dense over the rows: this thread owns unit j of the 4 features of NPT board points (the thermometer
      //         code makes each a predicated add of dL/dz), rows in ascending order; then the 6 bar/off/flag features ----
      float ss = 0.f;
      {
        float acc[NPT][4];
#pragma unroll
        for (int i = 0; i < NPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        const int8_t* bp = brd8 + r * NPT;
#pragma unroll 4
        for (int t = 0; t < T; ++t) {
          const float dz = S.hs[t * UPC + j];
#pragma unroll
          for (int i = 0; i < NPT; ++i) {
            const int c = bp[t * 52 + i];
            acc[i][0] += c > 0 ? dz : 0.f;
            acc[i][1] += c > 1 ? dz : 0.f;
            acc[i][2] += c > 2 ? dz : 0.f;
            acc[i][3] = fmaf(S.extab[c], dz, acc[i][3]);  // exact 0 for c <= 3
          }
        }
#pragma unroll
        for (int i = 0; i < NPT; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            S.G[((r * NPT + i) * 4 + k) * UPC + j] = acc[i][k];
            ss = fmaf(acc[i][k], acc[i][k], ss);
          }
      }
      if (r < 6) {
        const int f = 192 + r;
        const int bo = f == 192 ? 48 : f == 193 ? 50 : f == 194 ? 49 : 51;
        float acc = 0.f;
        if (f < 196) {
          const int base = (f & 1) ? 16 : 0;  // off/15 : bar/2
          for (int t = 0; t < T; ++t) {
            const int c = brd8[t * 52 + bo];
            if (c > 0) acc = fmaf(S.vtab[base + c], S.hs[t * UPC + j], acc);
          }
        } else {
          for (int t = 0; t < T; ++t) acc += S.flg[t] == f - 196 ? S.hs[t * UPC + j] : 0.f;
        }
        S.G[f * UPC + j] = acc;
        ss = fmaf(acc, acc, ss);
      } else if (r < 8) {  // r == 6: b1 gradient, r == 7: w2 gradient (fixed-order sum of the row-lane partials)
        float g = 0.f;
        for (int rr = 0; rr < RL; ++rr) g += S.pr[7 - r][rr * UPC + j];
        S.G[(NF + r - 6) * UPC + j] = g;
        ss = fmaf(g, g, ss);
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

      // ---- F: clip_grad_norm_ (trainer.py:125-128) and the Adam step (:139) on this CTA's slice ----
      float tot2 = 0.f;
      for (int c = 0; c < CL; ++c) tot2 += S.npart[c];
      const float total = sqrtf(tot2);
      float coef = 1.0f;
      if (a.grad_clip > 0.f) coef = fminf(1.0f, a.grad_clip / (total + 1e-6f));
      const float2 sc = S.sc[kstep & (NSC - 1)];
      for (int s = tid; s < NROW * UPC; s += NT) adam_step(S.P[s], S.M[s], S.V[s], S.G[s] * coef, sc.x, sc.y);
      if (tid == 0) {
        adam_step(S.b2[0], S.b2[1], S.b2[2], msum[4] * coef, sc.x, sc.y);
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
  bool attr_set;
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
  L->attr_set = false;
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

template <int UPC>
static int32_t launch_update(Learner* L, const LearnerArgs& a, cudaStream_t s) {
  const size_t smem = sizeof(Smem<UPC>);
  if (!L->attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_td0_update<UPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_td0_update)");
    L->attr_set = true;
  }
  const unsigned CL = (unsigned)(L->H / UPC);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, k_td0_update<UPC>, a), "k_td0_update launch");
}

int32_t learner_update(Learner* L, const int8_t* boards, const uint8_t* flags_or_meta, const float* reward, const int64_t* ep_offsets,
                       int64_t n_eps, int32_t records, float* out_metrics, int32_t* out_status, cudaStream_t s) {
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
  LearnerArgs a{boards, flags_or_meta, reward, ep_offsets, n_eps, records, L->params, L->m, L->v, &L->opt->step,
                L->lr, L->gamma, L->grad_clip, out_metrics, out_status, L->H};
  return L->H <= 128 ? launch_update<16>(L, a, s) : launch_update<32>(L, a, s);
}

}  // namespace bg
