// arena.cu -- GPU-resident self-play arena (sm_100a): tens of thousands of concurrent games per GPU.
//
// Replaces the reference's per-process worker loop: Worker.play_episode (src/multi/worker.py:78-174) driving
// BackgammonEnv.reset/step (src/environments/backgammon_env.py:92-329) with the terminal / shaping predicates of
// src/environments/env_helper.py:113-242, Experience/Episode recording (src/environments/episode.py) and the
// ExperienceQueue hand-off (src/multi/experience_queue.py).  One arena step == one ply of EVERY game:
//   movegen (movegen.cu) -> fused eval of all afterstates + current observations (eval.cu) ->
//   k_apply: softmax(V/T)/argmax pick, apply, rewards, episode record, terminal/cap handling, auto-reset, dice.
// Everything stays in HBM/L2; the host only enqueues kernels.  Finished episodes are copied into a device ring
// (compact records: 52-byte afterstate + scalars) that bg_arena_drain_episodes hands to the learner.
#include "arena.cuh"

#include <new>

#include "eval.cuh"
#include "movegen.cuh"
#include "select.cuh"
#include "two_ply.cuh"

namespace bg {

namespace {

constexpr uint64_t KEY_DICE = 0x6a09e667f3bcc908ull, KEY_ACT = 0xbb67ae8584caa73bull;

struct ArenaDev {
  // per-game state
  uint32_t* board;   // [G][13]
  uint8_t* player;   // [G]
  uint8_t* roll;     // [G][2]
  int32_t* step;     // [G] env steps in the current episode (passes included)
  int32_t* nexp;     // [G] experiences recorded in the current episode
  int32_t* npass;    // [G]
  uint8_t* gflags;   // [G] bit0/1 close-out given P1/P2, bit2/3 prime given P1/P2, bit4/5 player seen
  uint8_t* gstate;   // [G]
  int32_t* epcnt;    // [G][4] close_out_counts[2], prime_reward_counts[2]
  int64_t* serial;   // [G] episodes started by this slot
  int32_t* dice_ctr; // [G] dice draws in the current episode (Philox) / absolute tape position (tape mode)
  // per-game experience buffers [G][P]
  uint32_t* xb_after;  // [G][P][13]
  float* xb_v;
  float* xb_vnext;
  float* xb_reward;
  uint8_t* xb_meta;
  int16_t* xb_nmoves;
  int16_t* xb_action;
  uint8_t* xb_roll;  // [G][P][2]
  // finished-episode ring
  uint32_t* fp_after;
  float* fp_v;
  float* fp_vnext;
  float* fp_reward;
  uint8_t* fp_meta;
  int16_t* fp_nmoves;
  int16_t* fp_action;
  uint8_t* fp_roll;
  int64_t* ep_start;  // [EPS] absolute experience cursor of the episode's first record
  int32_t* ep_len;    // [EPS]
  int32_t* ep_info;   // [EPS][BG_EP_INFO_INTS]
  unsigned long long* ring;  // [0] head (eps<<40 | exps), [1] tail_eps, [2] tail_exps, [3] plan n_take, [4] plan end_exps,
                             // [5] end of the successful reservations, [6] limit eps, [7] limit exps
  // stats
  unsigned long long* stats;  // [BG_ARENA_NSTATS]
  // tape
  const uint8_t* tape;  // [G][L][2] or null
};

struct ArenaCfg {
  int64_t G;
  int32_t P;         // max plies per episode (reference MAX_TIMESTEPS = 300)
  int32_t move_cap;  // reference max_legal_moves = 500
  int64_t E;         // ring capacity in experiences
  int64_t EPS;       // ring capacity in episodes
  uint64_t seed;
  int64_t gid_base;
  int64_t tape_len;
  int32_t auto_reset;
};

// one dice roll (reference backgammon_env.py:310-311); all lanes compute the same value
__device__ __forceinline__ void roll_dice(const ArenaDev& D, const ArenaCfg& C, int64_t g, int64_t serial, int32_t& ctr, int& d0, int& d1,
                                          bool& exhausted) {
  if (D.tape) {
    exhausted = ctr >= C.tape_len;
    const int64_t p = exhausted ? C.tape_len - 1 : ctr;
    d0 = D.tape[(g * C.tape_len + p) * 2];
    d1 = D.tape[(g * C.tape_len + p) * 2 + 1];
  } else {
    uint32_t r[4];
    Philox::gen(C.seed ^ KEY_DICE, (uint64_t)(C.gid_base + g), ((uint64_t)serial << 20) | (uint64_t)(uint32_t)ctr, r);
    d0 = 1 + (int)__umulhi(r[0], 6u);
    d1 = 1 + (int)__umulhi(r[1], 6u);
    exhausted = false;
  }
  ctr += 1;
}

// reference BackgammonEnv.reset (backgammon_env.py:92-128): roll until non-double -> starter; roll until non-double again
__device__ __forceinline__ void reset_game(const ArenaDev& D, const ArenaCfg& C, int64_t g, int lane) {
  const int64_t serial = D.serial[g] + 1;
  int32_t ctr = D.tape ? D.dice_ctr[g] : 0;
  int d0, d1;
  bool ex = false, any_ex = false;
  do {
    roll_dice(D, C, g, serial, ctr, d0, d1, ex);
    any_ex |= ex;
  } while (d0 == d1 && !ex);
  const int starter = d0 < d1 ? 1 : 0;
  do {
    roll_dice(D, C, g, serial, ctr, d0, d1, ex);
    any_ex |= ex;
  } while (d0 == d1 && !ex);
  if (lane < 13) D.board[g * 13 + lane] = initial_board_word(lane);
  if (lane == 0) {
    D.serial[g] = serial;
    D.dice_ctr[g] = ctr;
    D.player[g] = (uint8_t)starter;
    D.roll[2 * g] = (uint8_t)d0;
    D.roll[2 * g + 1] = (uint8_t)d1;
    D.step[g] = 0;
    D.nexp[g] = 0;
    D.npass[g] = 0;
    D.gflags[g] = 0;
    D.epcnt[4 * g + 0] = D.epcnt[4 * g + 1] = D.epcnt[4 * g + 2] = D.epcnt[4 * g + 3] = 0;
    D.gstate[g] = any_ex ? BG_GAME_STOPPED : BG_GAME_ACTIVE;
  }
}

__global__ void __launch_bounds__(256) k_reset(ArenaDev D, ArenaCfg C) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)gridDim.x * 8;
  for (int64_t g = warp; g < C.G; g += nwarps) reset_game(D, C, g, lane);
}

// try to move the finished episode of game g into the ring; returns false when there is no room
__device__ __forceinline__ bool finalize_episode(const ArenaDev& D, const ArenaCfg& C, int64_t g, int lane, int win_type, int winner) {
  const int n = D.nexp[g];
  unsigned long long slot = 0;
  int ok = 1;
  if (lane == 0) {
    if (n > 0) {
      // One fetch-add per finished episode (a CAS loop serialised ~700 finishing warps per step).  The capacity test is
      // a threshold on the pre-add cursor (room for one more episode of the MAXIMUM length), hence monotone: the
      // successful reservations are a prefix of the cursor sequence, and ring[5] tracks where that prefix ends;
      // ring_repair() rolls the cursor back over the failed suffix before the ring is used again.
      const unsigned long long inc = (1ull << 40) | (unsigned long long)n;
      const unsigned long long cur = atomicAdd(&D.ring[0], inc);
      const unsigned long long eps = cur >> 40, exps = cur & ((1ull << 40) - 1);
      if (eps + 1 > D.ring[6] || exps + (unsigned long long)C.P > D.ring[7]) {
        ok = 0;
      } else {
        slot = cur;
        atomicMax(&D.ring[5], cur + inc);
      }
    }
  }
  ok = __shfl_sync(BG_FULL, ok, 0);
  if (!ok) return false;
  // stats
  if (lane == 0) {
    atomicAdd(&D.stats[BG_STAT_GAMES], 1ull);
    atomicAdd(&D.stats[BG_STAT_STEPS], (unsigned long long)D.step[g]);
    atomicAdd(&D.stats[BG_STAT_PASSES], (unsigned long long)D.npass[g]);
    atomicAdd(&D.stats[BG_STAT_DECISIONS], (unsigned long long)n);
    if (win_type > 0) {
      atomicAdd(&D.stats[BG_STAT_WIN_REGULAR + win_type - 1], 1ull);
      if (winner == 0) atomicAdd(&D.stats[BG_STAT_P1_WINS], 1ull);
    } else {
      atomicAdd(&D.stats[BG_STAT_TRUNCATED], 1ull);
    }
  }
  if (n == 0) return true;
  const unsigned long long s_lo = __shfl_sync(BG_FULL, (unsigned)(slot & 0xffffffffull), 0);
  const unsigned long long s_hi = __shfl_sync(BG_FULL, (unsigned)(slot >> 32), 0);
  slot = (s_hi << 32) | s_lo;
  const unsigned long long ep = slot >> 40, x0 = slot & ((1ull << 40) - 1);
  const int64_t eslot = (int64_t)(ep % (unsigned long long)C.EPS);
  if (lane == 0) {
    D.ep_start[eslot] = (int64_t)x0;
    D.ep_len[eslot] = n;
    int32_t* inf = D.ep_info + eslot * BG_EP_INFO_INTS;
    inf[0] = win_type;
    inf[1] = win_type > 0 ? winner : -1;
    inf[2] = D.step[g];
    inf[3] = D.npass[g];
    inf[4] = D.epcnt[4 * g + 0];
    inf[5] = D.epcnt[4 * g + 1];
    inf[6] = D.epcnt[4 * g + 2];
    inf[7] = D.epcnt[4 * g + 3];
    inf[8] = (D.gflags[g] >> 4) & 3;
    inf[9] = (int32_t)(C.gid_base + g);
    inf[10] = (int32_t)D.serial[g];
    inf[11] = 0;
  }
  const int64_t src = g * C.P;
  const int64_t d0 = (int64_t)(x0 % (unsigned long long)C.E);
  // The copy sits on the critical path of the ply (a finishing game's warp is the slowest of its CTA): the loads of several
  // records are issued before the first store, so their latencies overlap instead of adding up (source and destination never
  // alias, which the compiler cannot know).
  for (int t0 = lane; t0 < n * 13; t0 += 32 * 4) {
    uint32_t v[4];
    int64_t di[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = t0 + 32 * k;
      const int e = t / 13, w = t - e * 13;
      int64_t dst = d0 + e;
      if (dst >= C.E) dst -= C.E;
      di[k] = dst * 13 + w;
      v[k] = t < n * 13 ? D.xb_after[(src + e) * 13 + w] : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (t0 + 32 * k < n * 13) D.fp_after[di[k]] = v[k];
  }
  for (int e = lane; e < n; e += 32) {
    int64_t dst = d0 + e;
    if (dst >= C.E) dst -= C.E;
    const auto x_v = D.xb_v[src + e];
    const auto x_vn = D.xb_vnext[src + e];
    const auto x_r = D.xb_reward[src + e];
    const auto x_m = D.xb_meta[src + e];
    const auto x_nm = D.xb_nmoves[src + e];
    const auto x_a = D.xb_action[src + e];
    const auto x_r0 = D.xb_roll[2 * (src + e)];
    const auto x_r1 = D.xb_roll[2 * (src + e) + 1];
    D.fp_v[dst] = x_v;
    D.fp_vnext[dst] = x_vn;
    D.fp_reward[dst] = x_r;
    D.fp_meta[dst] = x_m;
    D.fp_nmoves[dst] = x_nm;
    D.fp_action[dst] = x_a;
    D.fp_roll[2 * dst] = x_r0;
    D.fp_roll[2 * dst + 1] = x_r1;
  }
  __threadfence();
  return true;
}

// One ply of every game.  Warp per game.
__global__ void __launch_bounds__(256) k_apply(ArenaDev D, ArenaCfg C, const uint32_t* __restrict__ pool, const long long* __restrict__ offsets,
                                               const int32_t* __restrict__ counts, const float* __restrict__ v_pool,
                                               const float* __restrict__ v_cur, const int32_t* __restrict__ forced_action,
                                               float temperature, const float* __restrict__ sel_score, const int32_t* __restrict__ sel_idx,
                                               const int32_t* __restrict__ sel_n, int sel_c, const float* __restrict__ pool_score) {
  __shared__ uint32_t sb[8][16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* bw = sb[wib];
  const int64_t warp = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)gridDim.x * 8;
  for (int64_t g = warp; g < C.G; g += nwarps) {
    __syncwarp();
    int st = D.gstate[g];
    if (st == BG_GAME_STOPPED) continue;
    int win_type = 0, winner = -1;
    bool finished = false;
    if (st == BG_GAME_WAIT) {  // episode finished earlier but the ring was full
      finished = true;
      win_type = (D.gflags[g] >> 6) & 3;
      winner = D.player[g];
    } else {
      // every per-game scalar the ply needs is loaded here, before the first store of the iteration: the compiler cannot move a
      // load above a store through another pointer, and each late load would cost a full memory latency on the critical path
      const int n_true = counts[g];
      const int n = n_true < C.move_cap ? n_true : C.move_cap;
      const long long off = offsets[g];
      int player = D.player[g];
      int step = D.step[g];
      const int64_t serial = D.serial[g];
      int32_t dctr = D.dice_ctr[g];
      const uint32_t gf_in = D.gflags[g];
      const int e_in = D.nexp[g];
      const uint8_t roll_in0 = D.roll[2 * g], roll_in1 = D.roll[2 * g + 1];
      const float vcur_in = v_cur[g];
      if (n_true < 0 || (n > 0 && off < 0)) {  // generator overflow for this item: surface, stop the game
        if (lane == 0) {
          atomicAdd(&D.stats[BG_STAT_ERRORS], 1ull);
          D.gstate[g] = BG_GAME_STOPPED;
        }
        continue;
      }
      bool done = false;
      if (n == 0) {
        // pass (backgammon_env.py:139-151; worker.py:103-113 records no Experience)
        player ^= 1;
        int d0, d1;
        bool ex;
        roll_dice(D, C, g, serial, dctr, d0, d1, ex);
        step += 1;
        if (lane == 0) {
          D.player[g] = (uint8_t)player;
          D.roll[2 * g] = (uint8_t)d0;
          D.roll[2 * g + 1] = (uint8_t)d1;
          D.npass[g] += 1;
          D.dice_ctr[g] = dctr;
          D.step[g] = step;
          if (ex) D.gstate[g] = BG_GAME_STOPPED;
        }
      } else {
        int a;
        if (forced_action && forced_action[g] >= 0) {
          a = forced_action[g] < n ? forced_action[g] : n - 1;
        } else {
          uint32_t r[4];
          Philox::gen(C.seed ^ KEY_ACT, (uint64_t)(C.gid_base + g), ((uint64_t)serial << 20) | (uint64_t)(uint32_t)step, r);
          const float u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
          if (sel_n && sel_n[g] > 0) {  // 2-ply over the top candidates: pick among their scores, map back to the action index
            const int al = warp_select(sel_score + g * sel_c, sel_n[g], temperature, u, lane);
            a = sel_idx[g * sel_c + al];
          } else {
            a = warp_select((pool_score ? pool_score : v_pool) + off, n, temperature, u, lane);
          }
        }
        if (lane < 13) bw[lane] = pool[(off + a) * 13 + lane];
        __syncwarp();
        const int mover = player, opp = 1 - player;
        const uint32_t w12 = bw[12];
        const int own_off = (w12 >> (16 + 8 * mover)) & 0xff, opp_off = (w12 >> (16 + 8 * opp)) & 0xff;
        const int opp_bar = (w12 >> (8 * opp)) & 0xff;
        uint32_t oc = 0, pc = 0;
        if (lane < 24) {
          pc = (bw[mover * 6 + (lane >> 2)] >> ((lane & 3) * 8)) & 0xffu;
          oc = (bw[opp * 6 + (lane >> 2)] >> ((lane & 3) * 8)) & 0xffu;
        }
        const uint32_t made = __ballot_sync(BG_FULL, pc >= 2) & 0xffffffu;
        const uint32_t opp_occ = __ballot_sync(BG_FULL, oc > 0) & 0xffffffu;
        const uint32_t home = mover == 0 ? 0xfc0000u : 0x00003fu;
        float reward = 0.f;
        uint32_t meta = (uint32_t)mover;
        int next_flag = mover;
        if (own_off >= 15) {  // terminal (backgammon_env.py:167-193, env_helper.py:113-163)
          const bool backgammon = opp_off == 0 && ((opp_occ & home) != 0 || opp_bar > 0);
          const bool gammon = opp_off == 0;
          win_type = backgammon ? 3 : gammon ? 2 : 1;
          reward = backgammon ? 2.5f : gammon ? 2.0f : 1.0f;
          winner = mover;
          done = true;
          meta |= 4u;
        } else {  // shaping rewards, once per player per game (backgammon_env.py:195-218, env_helper.py:167-242)
          uint32_t gf = gf_in;
          const bool closed = opp_bar > 0 && (made & home) == home;
          const uint32_t m5 = made & (made >> 1) & (made >> 2) & (made >> 3) & (made >> 4);
          bool prime = false;
          if (m5) {
            if (mover == 0) {
              const int idx = __ffs(m5) - 1 + 4;
              prime = idx < 23 && (opp_occ >> (idx + 1)) != 0;
            } else {
              const int idx = 31 - __clz(m5);
              prime = (opp_occ & ((1u << idx) - 1u)) != 0;
            }
          }
          if (closed && !((gf >> mover) & 1u)) {
            reward += 0.30f;
            gf |= 1u << mover;
            meta |= 8u;
          }
          if (prime && !((gf >> (2 + mover)) & 1u)) {
            reward += 0.20f;
            gf |= 1u << (2 + mover);
            meta |= 16u;
          }
          gf |= 1u << (4 + mover);
          player ^= 1;
          next_flag = player;
          int d0, d1;
          bool ex;
          roll_dice(D, C, g, serial, dctr, d0, d1, ex);
          if (lane == 0) {
            D.gflags[g] = (uint8_t)gf;
            if (meta & 8u) D.epcnt[4 * g + mover] += 1;
            if (meta & 16u) D.epcnt[4 * g + 2 + mover] += 1;
            D.player[g] = (uint8_t)player;
            // roll recorded below needs the OLD roll: write after recording
            D.dice_ctr[g] = dctr;
            if (ex) D.gstate[g] = BG_GAME_STOPPED;
          }
          // stash new roll in registers until the experience is recorded
          meta |= (uint32_t)d0 << 8 | (uint32_t)d1 << 12;
        }
        meta |= (uint32_t)next_flag << 1;
        const int e = e_in;
        const int64_t xi = g * C.P + e;
        if (lane < 13) {
          D.xb_after[xi * 13 + lane] = bw[lane];
          D.board[g * 13 + lane] = bw[lane];
        }
        if (lane == 0) {
          D.xb_v[xi] = vcur_in;
          D.xb_vnext[xi] = v_pool[off + a];
          D.xb_reward[xi] = reward;
          D.xb_meta[xi] = (uint8_t)(meta & 0xffu);
          D.xb_nmoves[xi] = (int16_t)n;
          D.xb_action[xi] = (int16_t)a;
          D.xb_roll[2 * xi] = roll_in0;
          D.xb_roll[2 * xi + 1] = roll_in1;
          D.nexp[g] = e + 1;
          step += 1;
          D.step[g] = step;
          if (done) {
            D.gflags[g] = (uint8_t)((gf_in & 0x3fu) | ((uint32_t)win_type << 6) | (1u << (4 + mover)));
          } else {
            D.roll[2 * g] = (uint8_t)((meta >> 8) & 15u);
            D.roll[2 * g + 1] = (uint8_t)((meta >> 12) & 15u);
          }
          atomicAdd(&D.stats[BG_STAT_AFTERSTATES], (unsigned long long)n);
        }
        step = __shfl_sync(BG_FULL, step, 0);
      }
      __syncwarp();
      finished = done || step >= C.P;  // worker.py:101 `while not done and step_count < max_steps`
      if (finished && !done) win_type = 0;
      if (D.gstate[g] == BG_GAME_STOPPED && !finished) continue;
    }
    if (finished) {
      __syncwarp();
      if (finalize_episode(D, C, g, lane, win_type, winner)) {
        if (C.auto_reset)
          reset_game(D, C, g, lane);
        else if (lane == 0)
          D.gstate[g] = BG_GAME_STOPPED;
      } else if (lane == 0) {
        D.gstate[g] = BG_GAME_WAIT;
        atomicAdd(&D.stats[BG_STAT_WAIT_STEPS], 1ull);
      }
    }
  }
}

// 2-ply, reference setting (src/multi/two_ply.py:153-193, the commented-out integration): when a decision has at least
// `C` legal moves, the C best 1-ply candidates (torch.topk order: descending value, lowest index first on ties) are rescored
// by bg_two_ply and the action is drawn among them; otherwise the 1-ply policy is used.  Warp per game.
__global__ void __launch_bounds__(256) k_pick_candidates(ArenaDev D, ArenaCfg Cc, const uint32_t* __restrict__ pool, const long long* __restrict__ offsets,
                                                         const int32_t* __restrict__ counts, const float* __restrict__ v_pool, int C,
                                                         uint32_t* __restrict__ cand_boards, uint8_t* __restrict__ cand_mover,
                                                         uint8_t* __restrict__ cand_active, float* __restrict__ cand_S, int32_t* __restrict__ cand_idx,
                                                         int32_t* __restrict__ sel_n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t g = warp; g < Cc.G; g += nwarps) {
    int n = counts[g];
    if (n > Cc.move_cap) n = Cc.move_cap;
    const long long off = offsets[g];
    const bool use = D.gstate[g] == BG_GAME_ACTIVE && n >= C && off >= 0;
    if (!use) {
      if (lane == 0) sel_n[g] = 0;
      if (lane < C) cand_active[g * C + lane] = 0;
      continue;
    }
    int chosen[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) chosen[c] = -1;
    for (int c = 0; c < C; ++c) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = lane; i < n; i += 32) {
        bool taken = false;
#pragma unroll
        for (int q = 0; q < 8; ++q) taken |= chosen[q] == i;
        const float x = v_pool[off + i];
        if (!taken && (bi == 0x7fffffff || x > best)) {
          best = x;
          bi = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(BG_FULL, best, o);
        const int oi = __shfl_xor_sync(BG_FULL, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) {
          best = ob;
          bi = oi;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q == c) chosen[q] = bi;
      const int64_t slot = g * C + c;
      if (lane < 13) cand_boards[slot * 13 + lane] = pool[(off + bi) * 13 + lane];
      if (lane == 0) {
        cand_mover[slot] = D.player[g];
        cand_active[slot] = 1;
        cand_S[slot] = best;
        cand_idx[slot] = bi;
      }
    }
    if (lane == 0) sel_n[g] = C;
  }
}

// ring cursor repair + per-step limits (see finalize_episode); executed by ONE thread before the ring is used
__device__ __forceinline__ void ring_repair(const ArenaDev& D, const ArenaCfg& C) {
  unsigned long long head = D.ring[0];
  const unsigned long long good = D.ring[5];
  if (head > good) head = good;  // drop reservations that failed the capacity test
  D.ring[0] = head;
  D.ring[5] = head;
  D.ring[6] = D.ring[1] + (unsigned long long)C.EPS;  // tails only move in drain (stream ordered)
  D.ring[7] = D.ring[2] + (unsigned long long)C.E;
}

__global__ void k_prologue(ArenaDev D, ArenaCfg C, uint8_t* __restrict__ active) {
  if (blockIdx.x == 0 && threadIdx.x == 0) ring_repair(D, C);
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < C.G; g += (int64_t)gridDim.x * blockDim.x)
    active[g] = D.gstate[g] == BG_GAME_ACTIVE;
}

// ---- drain ------------------------------------------------------------------------------------------------------
__global__ void k_drain_plan(ArenaDev D, ArenaCfg C, int64_t max_eps, int64_t max_exps) {
  __shared__ unsigned long long s_n, s_end;
  if (threadIdx.x == 0) ring_repair(D, C);
  __syncthreads();
  const unsigned long long head = D.ring[0];
  const unsigned long long head_eps = head >> 40;
  const unsigned long long tail_eps = D.ring[1], tail_exps = D.ring[2];
  if (threadIdx.x == 0) {
    s_n = 0;
    s_end = tail_exps;
  }
  __syncthreads();
  unsigned long long avail = head_eps - tail_eps;
  if (avail > (unsigned long long)max_eps) avail = (unsigned long long)max_eps;
  unsigned long long my_n = 0, my_end = tail_exps;
  for (unsigned long long k = threadIdx.x; k < avail; k += blockDim.x) {
    const int64_t slot = (int64_t)((tail_eps + k) % (unsigned long long)C.EPS);
    const unsigned long long end = (unsigned long long)D.ep_start[slot] + (unsigned long long)D.ep_len[slot];
    if (end - tail_exps <= (unsigned long long)max_exps) {  // monotonic in k
      my_n += 1;
      if (end > my_end) my_end = end;
    }
  }
  atomicAdd(&s_n, my_n);
  atomicMax(&s_end, my_end);
  __syncthreads();
  if (threadIdx.x == 0) {
    D.ring[3] = s_n;
    D.ring[4] = s_end;
  }
}

struct DrainOut {
  int8_t* after;      // [max_exps][52]
  uint8_t* meta;      // [max_exps]
  float* reward;      // [max_exps]
  float* v;           // [max_exps]
  float* vnext;       // [max_exps]
  int16_t* nmoves;    // [max_exps] or null
  int16_t* action;    // [max_exps] or null
  uint8_t* roll;      // [max_exps][2] or null
  int64_t* ep_offsets;  // [max_eps + 1]
  int32_t* ep_info;     // [max_eps][BG_EP_INFO_INTS]
  int64_t* out_n;       // [2] episodes, experiences
};

__global__ void __launch_bounds__(256) k_drain_copy(ArenaDev D, ArenaCfg C, DrainOut O) {
  const unsigned long long n_take = D.ring[3], tail_eps = D.ring[1], tail_exps = D.ring[2];
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    O.out_n[0] = (int64_t)n_take;
    O.out_n[1] = (int64_t)(D.ring[4] - tail_exps);
    O.ep_offsets[n_take] = (int64_t)(D.ring[4] - tail_exps);
  }
  for (int64_t k = warp; k < (int64_t)n_take; k += nwarps) {
    const int64_t slot = (int64_t)((tail_eps + k) % (unsigned long long)C.EPS);
    const unsigned long long x0 = (unsigned long long)D.ep_start[slot];
    const int n = D.ep_len[slot];
    const int64_t o0 = (int64_t)(x0 - tail_exps);
    if (lane == 0) O.ep_offsets[k] = o0;
    if (lane < BG_EP_INFO_INTS) O.ep_info[k * BG_EP_INFO_INTS + lane] = D.ep_info[slot * BG_EP_INFO_INTS + lane];
    uint32_t* oa = reinterpret_cast<uint32_t*>(O.after);
    const int64_t s0 = (int64_t)(x0 % (unsigned long long)C.E);
    for (int t = lane; t < n * 13; t += 32) {
      const int e = t / 13, w = t - e * 13;
      int64_t src = s0 + e;
      if (src >= C.E) src -= C.E;
      oa[(o0 + e) * 13 + w] = D.fp_after[src * 13 + w];
    }
    for (int e = lane; e < n; e += 32) {
      int64_t src = s0 + e;
      if (src >= C.E) src -= C.E;
      O.meta[o0 + e] = D.fp_meta[src];
      O.reward[o0 + e] = D.fp_reward[src];
      O.v[o0 + e] = D.fp_v[src];
      O.vnext[o0 + e] = D.fp_vnext[src];
      if (O.nmoves) O.nmoves[o0 + e] = D.fp_nmoves[src];
      if (O.action) O.action[o0 + e] = D.fp_action[src];
      if (O.roll) {
        O.roll[2 * (o0 + e)] = D.fp_roll[2 * src];
        O.roll[2 * (o0 + e) + 1] = D.fp_roll[2 * src + 1];
      }
    }
  }
}

__global__ void k_drain_commit(ArenaDev D) {
  D.ring[1] += D.ring[3];
  D.ring[2] = D.ring[4];
}

__global__ void k_export_state(ArenaDev D, int64_t G, int8_t* boards, uint8_t* players, uint8_t* rolls, uint8_t* gstate) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
    if (boards)
      for (int w = 0; w < 13; ++w) reinterpret_cast<uint32_t*>(boards)[g * 13 + w] = D.board[g * 13 + w];
    if (players) players[g] = D.player[g];
    if (rolls) {
      rolls[2 * g] = D.roll[2 * g];
      rolls[2 * g + 1] = D.roll[2 * g + 1];
    }
    if (gstate) gstate[g] = D.gstate[g];
  }
}

}  // namespace

struct Arena {
  int device = 0;
  ArenaCfg C{};
  ArenaDev D{};
  int32_t H = 0;
  float* prepared[2] = {nullptr, nullptr};
  int cur_w = -1;
  float temperature = 1.5f;  // reference INITIAL_TEMPERATURE (src/config/configuration.py:23)
  int64_t version = 0;
  // per-step scratch
  int8_t* pool = nullptr;
  int64_t pool_cap = 0;
  uint8_t* pflags = nullptr;
  int64_t* offsets = nullptr;
  int32_t* counts = nullptr;
  int64_t* total = nullptr;
  int32_t* status = nullptr;
  float* v_pool = nullptr;
  float* v_cur = nullptr;
  uint8_t* active = nullptr;
  void* ws = nullptr;
  int64_t ws_bytes = 0;
  uint8_t* tape_dev = nullptr;
  // 2-ply lookahead (bg_arena_set_lookahead); scratch is allocated on first use
  int32_t la_cands = 4, la_topk = 5;  // reference: top-4 candidates, mean of the top-5 replies, alpha 1.0, beta 0.9
  float la_alpha = 1.0f, la_beta = 0.9f;
  int8_t* cand_boards = nullptr;
  uint8_t* cand_mover = nullptr;
  uint8_t* cand_active = nullptr;
  float* cand_S = nullptr;
  float* cand_score = nullptr;
  int32_t* cand_idx = nullptr;
  int32_t* sel_n = nullptr;
  float* pool_score = nullptr;
  int32_t* tp_status = nullptr;
  void* tp_ws = nullptr;
  int64_t tp_ws_bytes = 0;
  int64_t tp_cands_alloc = 0;
  // side stream: the pool evaluation of the bulk move-generator tier and the evaluation of the current positions overlap the tail tiers
  SideCtx side{};
  cudaEvent_t ev_fork = nullptr, ev_cur = nullptr;
  std::vector<void*> allocs;
};

template <typename T>
static int32_t dmalloc(Arena* A, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
  if (e != cudaSuccess) return check_cuda(e, "cudaMalloc(arena)");
  if (zero) {
    e = cudaMemset(q, 0, count * sizeof(T) + 16);
    if (e != cudaSuccess) return check_cuda(e, "cudaMemset(arena)");
  }
  A->allocs.push_back(q);
  *p = (T*)q;
  return BG_OK;
}

#define TRY(x)                  \
  do {                          \
    int32_t _rc = (x);          \
    if (_rc != BG_OK) return _rc; \
  } while (0)

int32_t arena_create(Arena** out, int32_t device, int64_t n_games, int32_t H, int32_t max_plies, int32_t move_cap, uint64_t seed,
                     int64_t game_id_base, int64_t ring_exps, int64_t ring_eps, int32_t auto_reset) {
  if (n_games <= 0 || n_games >= (1ll << 31) || max_plies <= 0 || max_plies > 32767 || move_cap <= 0 || move_cap > BG_MAX_ITEM_MOVES) {
    set_error("bg_arena_create: bad sizes");
    return BG_ERR_ARG;
  }
  if (H < 32 || H > 256 || H % 32) {
    set_error("bg_arena_create: H must be a multiple of 32 in [32,256]");
    return BG_ERR_ARG;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return check_cuda(e, "cudaSetDevice");
  Arena* A = new (std::nothrow) Arena();
  if (!A) return BG_ERR_ARG;
  A->device = device;
  A->H = H;
  ArenaCfg& C = A->C;
  C.G = n_games;
  C.P = max_plies;
  C.move_cap = move_cap;
  C.E = ring_exps > 0 ? ring_exps : (n_games * 32 > (1 << 20) ? n_games * 32 : (1 << 20));
  if (C.E < max_plies) C.E = max_plies;
  C.EPS = ring_eps > 0 ? ring_eps : (n_games > 4096 ? n_games : 4096);
  C.seed = seed;
  C.gid_base = game_id_base;
  C.tape_len = 0;
  C.auto_reset = auto_reset;
  ArenaDev& D = A->D;
  const size_t G = (size_t)n_games, GP = G * (size_t)max_plies;
  int32_t rc = BG_OK;
  auto fail = [&](int32_t r) {
    for (void* p : A->allocs) cudaFree(p);
    delete A;
    return r;
  };
#define AL(ptr, count)                                    \
  if ((rc = dmalloc(A, &(ptr), (count))) != BG_OK) return fail(rc)
  AL(D.board, G * 13);
  AL(D.player, G);
  AL(D.roll, G * 2);
  AL(D.step, G);
  AL(D.nexp, G);
  AL(D.npass, G);
  AL(D.gflags, G);
  AL(D.gstate, G);
  AL(D.epcnt, G * 4);
  AL(D.serial, G);
  AL(D.dice_ctr, G);
  AL(D.xb_after, GP * 13);
  AL(D.xb_v, GP);
  AL(D.xb_vnext, GP);
  AL(D.xb_reward, GP);
  AL(D.xb_meta, GP);
  AL(D.xb_nmoves, GP);
  AL(D.xb_action, GP);
  AL(D.xb_roll, GP * 2);
  AL(D.fp_after, (size_t)C.E * 13);
  AL(D.fp_v, (size_t)C.E);
  AL(D.fp_vnext, (size_t)C.E);
  AL(D.fp_reward, (size_t)C.E);
  AL(D.fp_meta, (size_t)C.E);
  AL(D.fp_nmoves, (size_t)C.E);
  AL(D.fp_action, (size_t)C.E);
  AL(D.fp_roll, (size_t)C.E * 2);
  AL(D.ep_start, (size_t)C.EPS);
  AL(D.ep_len, (size_t)C.EPS);
  AL(D.ep_info, (size_t)C.EPS * BG_EP_INFO_INTS);
  AL(D.ring, 8);
  AL(D.stats, BG_ARENA_NSTATS);
  A->pool_cap = n_games * 128 < n_games * (int64_t)move_cap ? n_games * 128 : n_games * (int64_t)move_cap;
  if (A->pool_cap < 65536) A->pool_cap = n_games * (int64_t)move_cap < 65536 ? n_games * (int64_t)move_cap : 65536;
  AL(A->pool, (size_t)A->pool_cap * BG_BOARD_BYTES);
  AL(A->pflags, (size_t)A->pool_cap);
  AL(A->offsets, G);
  AL(A->counts, G);
  AL(A->total, 2);  // [0] rows of the whole pool, [1] rows written by the move generator's first tier
  AL(A->status, 2);
  AL(A->v_pool, (size_t)A->pool_cap);
  AL(A->v_cur, G);
  AL(A->active, G);
  A->ws_bytes = movegen_workspace_bytes(n_games);
  {
    uint8_t* w = nullptr;
    AL(w, (size_t)A->ws_bytes);
    A->ws = w;
  }
  const size_t pw = (size_t)prepared_weights_bytes(H) / 4;
  AL(A->prepared[0], pw);
  AL(A->prepared[1], pw);
#undef AL
  *out = A;
  return BG_OK;
}

int arena_device(const Arena* A) { return A ? A->device : 0; }

int32_t arena_destroy(Arena* A) {
  if (!A) return BG_OK;
  cudaSetDevice(A->device);
  side_ctx_destroy(&A->side);
  for (cudaEvent_t ev : {A->ev_fork, A->ev_cur})
    if (ev) cudaEventDestroy(ev);
  for (void* p : A->allocs) cudaFree(p);
  delete A;
  return BG_OK;
}

int32_t arena_set_weights(Arena* A, const float* packed_dev, int64_t version, float temperature, cudaStream_t s) {
  const int nxt = A->cur_w < 0 ? 0 : 1 - A->cur_w;  // double buffer: kernels already enqueued keep the old table
  TRY(prepare_weights_launch(packed_dev, A->H, A->prepared[nxt], s));
  A->cur_w = nxt;
  A->version = version;
  A->temperature = temperature;
  return BG_OK;
}

int32_t arena_set_dice_tape(Arena* A, const uint8_t* tape_host_or_dev, int64_t L, cudaStream_t s) {
  if (!tape_host_or_dev || L <= 0) {
    A->D.tape = nullptr;
    A->C.tape_len = 0;
    return BG_OK;
  }
  uint8_t* t = nullptr;
  TRY(dmalloc(A, &t, (size_t)A->C.G * (size_t)L * 2, false));
  cudaError_t e = cudaMemcpyAsync(t, tape_host_or_dev, (size_t)A->C.G * (size_t)L * 2, cudaMemcpyDefault, s);
  if (e != cudaSuccess) return check_cuda(e, "copy dice tape");
  A->D.tape = t;
  A->C.tape_len = L;
  return BG_OK;
}

static int grid_for(int64_t warps_needed) {
  int64_t blocks = (warps_needed + 7) / 8;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < 148 * 8 ? blocks : 148 * 8);
}

int32_t arena_reset(Arena* A, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(A->D.ring, 0, 8 * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return check_cuda(e, "memset ring");
  e = cudaMemsetAsync(A->D.stats, 0, BG_ARENA_NSTATS * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return check_cuda(e, "memset stats");
  e = cudaMemsetAsync(A->D.serial, 0, (size_t)A->C.G * sizeof(int64_t), s);
  if (e != cudaSuccess) return check_cuda(e, "memset serial");
  e = cudaMemsetAsync(A->D.dice_ctr, 0, (size_t)A->C.G * sizeof(int32_t), s);
  if (e != cudaSuccess) return check_cuda(e, "memset dice_ctr");
  k_reset<<<grid_for(A->C.G), 256, 0, s>>>(A->D, A->C);
  e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_reset launch");
  return BG_OK;
}

int32_t arena_set_lookahead(Arena* A, int32_t n_candidates, int32_t top_k, float alpha, float beta) {
  if (n_candidates < 0 || n_candidates > 8 || n_candidates == 1 || top_k < 1 || top_k > 8) {
    set_error("bg_arena_set_lookahead: n_candidates must be 0 (all) or 2..8, top_k in 1..8");
    return BG_ERR_ARG;
  }
  A->la_cands = n_candidates;
  A->la_topk = top_k;
  A->la_alpha = alpha;
  A->la_beta = beta;
  return BG_OK;
}

static int32_t ensure_two_ply_scratch(Arena* A) {
  const int64_t G = A->C.G;
  const int64_t n_cand = A->la_cands > 0 ? G * A->la_cands : A->pool_cap;
  if (A->tp_cands_alloc >= n_cand && A->tp_ws) return BG_OK;
  int32_t rc;
#define AL2(ptr, count) \
  if ((rc = dmalloc(A, &(ptr), (count))) != BG_OK) return rc
  AL2(A->cand_boards, (size_t)G * 8 * BG_BOARD_BYTES);
  AL2(A->cand_mover, (size_t)G * 8);
  AL2(A->cand_active, (size_t)G * 8);
  AL2(A->cand_S, (size_t)G * 8);
  AL2(A->cand_score, (size_t)G * 8);
  AL2(A->cand_idx, (size_t)G * 8);
  AL2(A->sel_n, (size_t)G);
  AL2(A->pool_score, (size_t)A->pool_cap);
  AL2(A->tp_status, 4);
  A->tp_ws_bytes = two_ply_workspace_bytes(n_cand);
  uint8_t* w = nullptr;
  AL2(w, (size_t)A->tp_ws_bytes);
  A->tp_ws = w;
#undef AL2
  A->tp_cands_alloc = A->pool_cap > G * 8 ? A->pool_cap : G * 8;
  return BG_OK;
}

int32_t arena_step(Arena* A, int32_t n_plies, int32_t lookahead, const int32_t* forced_action, cudaStream_t s) {
  if (A->cur_w < 0) {
    set_error("bg_arena_step: no weights set (call bg_arena_set_weights first)");
    return BG_ERR_ARG;
  }
  if (lookahead != 1 && lookahead != 2) {
    set_error("bg_arena_step: lookahead must be 1 or 2");
    return BG_ERR_ARG;
  }
  const ArenaCfg& C = A->C;
  if (lookahead == 2) TRY(ensure_two_ply_scratch(A));
  for (int ply = 0; ply < n_plies; ++ply) {
    k_prologue<<<(int)((C.G + 255) / 256 < 1184 ? (C.G + 255) / 256 : 1184), 256, 0, s>>>(A->D, A->C, A->active);
    MovegenArgs m{};
    m.boards = reinterpret_cast<const int8_t*>(A->D.board);
    m.players = A->D.player;
    m.rolls = A->D.roll;
    m.B = C.G;
    m.item_cap = C.move_cap;
    m.pool_cap = A->pool_cap;
    m.out_boards = A->pool;
    m.out_submoves = nullptr;
    m.out_owner = nullptr;
    m.out_flags = A->pflags;
    m.out_offsets = A->offsets;
    m.out_count = A->counts;
    m.out_total = A->total;
    m.out_status = A->status;
    m.workspace = A->ws;
    m.workspace_bytes = A->ws_bytes;
    m.active = A->active;
    // fork: the side stream evaluates the current positions (independent of move generation) and, as soon as the bulk tier
    // is done, the afterstates it produced; the main stream meanwhile runs the tail tiers (a few very wide doubles trees whose
    // latency used to sit on the ply's critical path) and then evaluates only the rows they added
    if (!A->side.stream) {
      TRY(side_ctx_create(&A->side, true));
      cudaError_t ce = cudaEventCreateWithFlags(&A->ev_fork, cudaEventDisableTiming);
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&A->ev_cur, cudaEventDisableTiming);
      if (ce != cudaSuccess) return check_cuda(ce, "arena events");
    }
    cudaError_t fe = cudaEventRecord(A->ev_fork, s);
    if (fe == cudaSuccess) fe = cudaStreamWaitEvent(A->side.stream, A->ev_fork, 0);
    if (fe != cudaSuccess) return check_cuda(fe, "arena fork");
    EvalArgs ec{reinterpret_cast<const int8_t*>(A->D.board), A->D.player, nullptr, nullptr, C.G, nullptr, C.G, A->prepared[A->cur_w], A->H, A->v_cur};
    TRY(eval_launch(ec, A->side.stream));
    TRY(movegen_eval_overlapped(m, A->total, A->prepared[A->cur_w], A->H, A->v_pool, &A->side, s));  // joins the side stream
    const float* sel_score = nullptr;
    const int32_t* sel_idx = nullptr;
    const int32_t* sel_n = nullptr;
    const float* pool_score = nullptr;
    if (lookahead == 2) {
      TwoPlyArgs t{};
      t.prepared = A->prepared[A->cur_w];
      t.H = A->H;
      t.top_k = A->la_topk;
      t.alpha = A->la_alpha;
      t.beta = A->la_beta;
      t.out_replies = nullptr;
      t.out_status = A->tp_status;
      t.workspace = A->tp_ws;
      t.workspace_bytes = A->tp_ws_bytes;
      t.reply_counter = A->D.stats + BG_STAT_REPLIES;
      t.side = &A->side;
      if (A->la_cands > 0) {  // reference setting: rescore the top-C candidates of each decision with >= C moves
        k_pick_candidates<<<grid_for(C.G), 256, 0, s>>>(A->D, A->C, reinterpret_cast<const uint32_t*>(A->pool), (const long long*)A->offsets,
                                                        A->counts, A->v_pool, A->la_cands, reinterpret_cast<uint32_t*>(A->cand_boards),
                                                        A->cand_mover, A->cand_active, A->cand_S, A->cand_idx, A->sel_n);
        t.cand_boards = A->cand_boards;
        t.mover = A->cand_mover;
        t.S = A->cand_S;
        t.N = C.G * A->la_cands;
        t.out_score = A->cand_score;
        t.cand_active = A->cand_active;
        TRY(two_ply_launch(t, s));
        sel_score = A->cand_score;
        sel_idx = A->cand_idx;
        sel_n = A->sel_n;
      } else {  // north-star setting: every legal afterstate is a candidate (needs this ply's afterstate count on the host)
        int64_t total = 0;
        cudaError_t e = cudaMemcpyAsync(&total, A->total, sizeof(total), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return check_cuda(e, "read afterstate count");
        if (total > A->pool_cap) total = A->pool_cap;
        t.cand_boards = A->pool;
        t.mover = A->pflags;
        t.S = A->v_pool;
        t.N = total;
        t.out_score = A->pool_score;
        TRY(two_ply_launch(t, s));
        pool_score = A->pool_score;
      }
    }
    k_apply<<<(int)((C.G + 7) / 8), 256, 0, s>>>(A->D, A->C, reinterpret_cast<const uint32_t*>(A->pool), (const long long*)A->offsets, A->counts,
                                                  A->v_pool, A->v_cur, forced_action, A->temperature, sel_score, sel_idx, sel_n, A->la_cands,
                                                  pool_score);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return check_cuda(e, "k_apply launch");
  }
  return BG_OK;
}

int32_t arena_drain(Arena* A, int64_t max_eps, int64_t max_exps, int8_t* after, uint8_t* meta, float* reward, float* v, float* vnext,
                    int16_t* nmoves, int16_t* action, uint8_t* roll, int64_t* ep_offsets, int32_t* ep_info, int64_t* out_n, cudaStream_t s) {
  if (max_eps <= 0 || max_exps <= 0 || !after || !meta || !reward || !v || !vnext || !ep_offsets || !ep_info || !out_n) {
    set_error("bg_arena_drain_episodes: bad arguments");
    return BG_ERR_ARG;
  }
  DrainOut O{after, meta, reward, v, vnext, nmoves, action, roll, ep_offsets, ep_info, out_n};
  k_drain_plan<<<1, 1024, 0, s>>>(A->D, A->C, max_eps, max_exps);
  k_drain_copy<<<grid_for(max_eps), 256, 0, s>>>(A->D, A->C, O);
  k_drain_commit<<<1, 1, 0, s>>>(A->D);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "drain launch");
  return BG_OK;
}

int32_t arena_stats(Arena* A, int64_t* out_dev_or_host, cudaStream_t s) {
  cudaError_t e = cudaMemcpyAsync(out_dev_or_host, A->D.stats, BG_ARENA_NSTATS * sizeof(int64_t), cudaMemcpyDefault, s);
  if (e != cudaSuccess) return check_cuda(e, "copy stats");
  return BG_OK;
}

int32_t arena_export_state(Arena* A, int8_t* boards, uint8_t* players, uint8_t* rolls, uint8_t* gstate, cudaStream_t s) {
  k_export_state<<<(int)((A->C.G + 255) / 256 < 1184 ? (A->C.G + 255) / 256 : 1184), 256, 0, s>>>(A->D, A->C.G, boards, players, rolls, gstate);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_export_state launch");
  return BG_OK;
}

}  // namespace bg
