// two_ply.cu -- 2-ply lookahead scorer (sm_100a).
//
// Replaces compute_scores_for_boards / compute_weighted_opponent_response (reference src/multi/two_ply.py:44-150):
//   for every candidate afterstate, over the 21 unordered opponent rolls (weights 1/36 doubles, 2/36 otherwise,
//   two_ply.py:10-35): generate the opponent's legal replies (movegen.cu), evaluate them with the opponent's flag
//   (eval.cu), take the mean of the top_k reply values (all of them if fewer; a roll with no reply adds 0), and
//   W = sum_r p_r * mean_r;  score = alpha * S - beta * W.
// top_k = 5, alpha = 1, beta = 0.9 is the reference's setting; top_k = 1 is north_star's "best reply" expectimax.
// The reference's random.sample(replies, 50) on 1-1/2-2/3-3 (two_ply.py:119-121) is OFF by default (every reply is evaluated; it is
// nondeterministic there) and available as an option (two_ply_reply_sampling): the replies of those three rolls are
// cut to a uniform sample WITHOUT replacement of `cap` of them BEFORE they are evaluated -- the "as shipped" cost of the lookahead --, keyed
// by (seed, candidate, roll) so that a run is reproducible: row j of the sample is reply perm(j), perm = a 4-round Feistel permutation of
// [0, 2^b) (b = the even number of bits that covers the reply count), cycle-walked into [0, n), round keys = Philox4x32-10(seed ^ KEY, item).
// Composition: position-major movegen (one warp per candidate, all 21 rolls; overflow -> per-item tiers) -> k_eval -> k_reduce (warp per candidate, top-k
// selection and the 21-term expectation entirely in registers).  Candidates are processed in workspace-sized chunks.
#include "two_ply.cuh"

#include <atomic>

#include "eval.cuh"
#include "movegen.cuh"

namespace bg {

namespace {

constexpr int ROWS_PER_ITEM = 48;  // pool rows provisioned per (candidate, roll) item (mean is ~22)
constexpr int MAX_TOPK = 8;

// the opponent (the player who replies) and the active flag of each candidate; the 21 rolls of a candidate are expanded by the
// position-major move generator itself (movegen21.cu), so the candidate boards are read in place instead of being replicated 21 times
__global__ void __launch_bounds__(256) k_repliers(const uint8_t* __restrict__ mover, const uint8_t* __restrict__ cand_active, int64_t c0, int64_t nc,
                                                  uint8_t* __restrict__ ip, uint8_t* __restrict__ iactive) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= nc) return;
  ip[c] = 1 - (mover[c0 + c] & 1);
  iactive[c] = cand_active ? cand_active[c0 + c] : (uint8_t)1;
}

__global__ void __launch_bounds__(256) k_reduce(const float* __restrict__ v, const long long* __restrict__ offsets, const int32_t* __restrict__ counts,
                                                const float* __restrict__ S, int64_t c0, int64_t nc, int top_k, float alpha, float beta,
                                                float* __restrict__ out_score, long long* __restrict__ out_replies,
                                                const int32_t* __restrict__ chunk_status, int32_t* __restrict__ out_status,
                                                unsigned long long* __restrict__ reply_counter) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_status && *chunk_status != 0) atomicMin(out_status, *chunk_status);
  for (int64_t c = warp; c < nc; c += nwarps) {
    double W = 0.0;
    long long nrep = 0;
    bool bad = false;
    // the 21 (count, offset) pairs of this candidate are fetched by 21 lanes at once: one memory latency instead of 21 dependent ones
    const int my_n = lane < 21 ? counts[c * 21 + lane] : 0;
    const long long my_off = lane < 21 ? offsets[c * 21 + lane] : 0;
    for (int r = 0; r < 21; ++r) {
      const int n = __shfl_sync(BG_FULL, my_n, r);
      const long long off = __shfl_sync(BG_FULL, my_off, r);
      if (n < 0 || (n > 0 && off < 0)) {
        bad = true;
        continue;
      }
      if (n == 0) continue;  // two_ply.py:123 `if opponent_moves:`
      nrep += n;
      float loc[MAX_TOPK];
#pragma unroll
      for (int q = 0; q < MAX_TOPK; ++q) loc[q] = -INFINITY;
      for (int i = lane; i < n; i += 32) {
        float x = v[off + i];
#pragma unroll
        for (int q = 0; q < MAX_TOPK; ++q) {  // insertion into the lane-local descending list
          const float hi = fmaxf(loc[q], x);
          x = fminf(loc[q], x);
          loc[q] = hi;
        }
      }
      const int k = n < top_k ? n : top_k;
      float sum = 0.f;
      for (int j = 0; j < k; ++j) {
        float m = loc[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(BG_FULL, m, o));
        const uint32_t who = __ballot_sync(BG_FULL, loc[0] == m);
        if (lane == __ffs(who) - 1) {
#pragma unroll
          for (int q = 0; q < MAX_TOPK - 1; ++q) loc[q] = loc[q + 1];
          loc[MAX_TOPK - 1] = -INFINITY;
        }
        sum += m;  // descending order, as torch.sort(...)[:k].mean() sums them
      }
      const float mean = sum / (float)k;
      const double p = (r == 0 || r == 6 || r == 11 || r == 15 || r == 18 || r == 20) ? 1.0 / 36.0 : 2.0 / 36.0;
      W += (double)mean * p;
    }
    if (lane == 0) {
      out_score[c0 + c] = bad ? NAN : (float)((double)alpha * (double)S[c0 + c] - (double)beta * W);
      if (out_replies) out_replies[c0 + c] = nrep;
      if (reply_counter && nrep) atomicAdd(reply_counter, (unsigned long long)nrep);
    }
  }
}

constexpr uint64_t KEY_SAMPLE = 0x3c6ef372fe94f82bull;

// Optional reply sampling (see the header comment): warp per (candidate, roll) item; copies the item's rows into a second pool -- all of them,
// or, for 1-1 / 2-2 / 3-3 with more than `cap` replies, the rows perm(0) .. perm(cap - 1) -- and writes the item's new offset / count.
__global__ void __launch_bounds__(256) k_sample_replies(const uint2* __restrict__ pool, const long long* __restrict__ off, const int32_t* __restrict__ cnt,
                                                        int64_t n_items, int cap, uint64_t seed, int64_t item_id_base, uint2* __restrict__ pool2,
                                                        long long* __restrict__ off2, int32_t* __restrict__ cnt2, unsigned long long* __restrict__ cursor2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t it = warp; it < n_items; it += nwarps) {
    const int n = cnt[it];
    const long long o = off[it];
    if (n <= 0 || o < 0) {  // no reply, or an item the generator could not hold: k_reduce sees the same (count, offset)
      if (lane == 0) {
        cnt2[it] = n;
        off2[it] = o;
      }
      continue;
    }
    const int r = (int)(it % 21);
    const bool sampled = (r == 0 || r == 6 || r == 11) && n > cap;  // two_ply.py:119-121: dice_roll in ([1, 1], [2, 2], [3, 3])
    const int m = sampled ? cap : n;
    unsigned long long dst = 0;
    if (lane == 0) dst = atomicAdd(cursor2, (unsigned long long)m);
    dst = __shfl_sync(BG_FULL, dst, 0);
    if (!sampled) {
      for (int i = lane; i < n; i += 32) pool2[dst + i] = pool[o + i];
    } else {
      uint32_t k[4];
      Philox::gen(seed ^ KEY_SAMPLE, (uint64_t)(item_id_base + it), 0ull, k);
      int bits = 2;
      while ((1 << bits) < n) bits += 2;
      const int half = bits >> 1;
      const uint32_t mask = (1u << half) - 1u;
      for (int j = lane; j < cap; j += 32) {
        uint32_t x = (uint32_t)j;
        do {
          uint32_t L = x >> half, R = x & mask;
#pragma unroll
          for (int rd = 0; rd < 4; ++rd) {
            const uint32_t t = L ^ (mix32(R, k[rd], (uint32_t)rd, 0x9E3779B9u) & mask);
            L = R;
            R = t;
          }
          x = (L << half) | R;
        } while (x >= (uint32_t)n);
        pool2[dst + j] = pool[o + x];
      }
    }
    if (lane == 0) {
      cnt2[it] = m;
      off2[it] = (long long)dst;
    }
  }
}

std::atomic<int32_t> g_sample_cap{0};
std::atomic<uint64_t> g_sample_seed{0};

struct Layout {
  int64_t C;  // candidates per chunk
  int64_t items, rows;
  int64_t o_ip, o_act, o_off, o_cnt, o_tot, o_pool, o_owner, o_val, o_ws, ws_bytes, total;
  int64_t o_pool2, o_off2, o_cnt2;  // reply sampling only
};

int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

Layout make_layout(int64_t C, bool sampled = false) {
  Layout L;
  L.C = C;
  L.items = C * 21;
  L.rows = L.items * ROWS_PER_ITEM;
  int64_t o = 0;
  L.o_ip = o;
  o += align256(L.C);
  L.o_act = o;
  o += align256(L.C);
  L.o_off = o;
  o += align256(L.items * 8);
  L.o_cnt = o;
  o += align256(L.items * 4);
  L.o_tot = o;
  o += 256;
  L.o_pool = o;
  o += align256(L.rows * 8);  // compact pool: (code, candidate index) per reply -- the reply boards are never materialised
  L.o_owner = o;
  o += align256(L.rows);
  L.o_val = o;
  o += align256(L.rows * 4);
  L.o_ws = o;
  L.ws_bytes = movegen_workspace_bytes(L.items);
  o += align256(L.ws_bytes);
  L.o_pool2 = L.o_off2 = L.o_cnt2 = 0;
  if (sampled) {  // the second (sampled) pool and its item table
    L.o_pool2 = o;
    o += align256(L.rows * 8);
    L.o_off2 = o;
    o += align256(L.items * 8);
    L.o_cnt2 = o;
    o += align256(L.items * 4);
  }
  L.total = o;
  return L;
}

constexpr int64_t DEFAULT_CHUNK = 262144;

}  // namespace

int64_t two_ply_workspace_bytes(int64_t N) {
  if (N < 1) N = 1;
  return make_layout(N < DEFAULT_CHUNK ? N : DEFAULT_CHUNK, g_sample_cap.load() > 0).total;  // (the second pool of the sampling option, if it is on)
}

int32_t two_ply_reply_sampling(int32_t cap, uint64_t seed) {
  g_sample_seed.store(seed);
  return g_sample_cap.exchange(cap > 0 ? cap : 0);
}

int32_t two_ply_launch(const TwoPlyArgs& a, cudaStream_t s) {
  if (a.N <= 0) return BG_OK;
  const int32_t sample_cap = g_sample_cap.load();
  const uint64_t sample_seed = g_sample_seed.load();
  const bool sampled = sample_cap > 0;
  if (a.top_k < 1 || a.top_k > MAX_TOPK) {
    set_error("bg_two_ply: top_k must be in [1,%d]", MAX_TOPK);
    return BG_ERR_ARG;
  }
  // largest chunk that fits the caller's workspace
  int64_t C = a.N < DEFAULT_CHUNK ? a.N : DEFAULT_CHUNK;
  while (C > 1 && make_layout(C, sampled).total > a.workspace_bytes) C = (C + 1) / 2;  // (sampling needs a second pool: smaller chunks)
  Layout L = make_layout(C, sampled);
  if (L.total > a.workspace_bytes) {
    set_error("bg_two_ply: workspace too small (%lld bytes; need >= %lld)", (long long)a.workspace_bytes, (long long)L.total);
    return BG_ERR_ARG;
  }
  char* w = (char*)a.workspace;
  if (a.out_status) {
    cudaError_t e = cudaMemsetAsync(a.out_status, 0, 4, s);
    if (e != cudaSuccess) return check_cuda(e, "memset status");
  }
  for (int64_t c0 = 0; c0 < a.N; c0 += C) {
    const int64_t nc = a.N - c0 < C ? a.N - c0 : C;
    int64_t blocks = (nc + 7) / 8;
    const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
    k_repliers<<<(int)((nc + 255) / 256), 256, 0, s>>>(a.mover, a.cand_active, c0, nc, (uint8_t*)(w + L.o_ip), (uint8_t*)(w + L.o_act));
    MovegenArgs m{};
    m.boards = a.cand_boards + c0 * 52;
    m.players = (const uint8_t*)(w + L.o_ip);
    m.rolls = nullptr;
    m.all_rolls = 1;  // items = candidate * 21 + roll index
    m.B = nc;
    m.item_cap = BG_MAX_ITEM_MOVES;
    m.pool_cap = L.rows;
    m.out_boards = nullptr;
    m.out_codes = (uint2*)(w + L.o_pool);
    m.out_submoves = nullptr;
    m.out_owner = nullptr;
    m.out_flags = nullptr;
    m.out_offsets = (int64_t*)(w + L.o_off);
    m.out_count = (int32_t*)(w + L.o_cnt);
    m.out_status = (int32_t*)(w + L.o_tot + 16);
    m.workspace = w + L.o_ws;
    m.workspace_bytes = L.ws_bytes;
    m.active = (const uint8_t*)(w + L.o_act);
    const long long* red_off = (const long long*)(w + L.o_off);
    const int32_t* red_cnt = (const int32_t*)(w + L.o_cnt);
    if (!sampled) {
      int32_t rc = movegen_eval_overlapped(m, (int64_t*)(w + L.o_tot), a.prepared, a.H, (float*)(w + L.o_val), a.side, s);
      if (rc != BG_OK) return rc;
    } else {
      // generate every reply, cut 1-1 / 2-2 / 3-3 to the sample, evaluate only what is left
      m.out_total = (int64_t*)(w + L.o_tot);
      int32_t rc = movegen_launch(m, s);
      if (rc != BG_OK) return rc;
      unsigned long long* cursor2 = (unsigned long long*)(w + L.o_tot + 32);
      cudaError_t ce = cudaMemsetAsync(cursor2, 0, 8, s);
      if (ce != cudaSuccess) return check_cuda(ce, "memset sample cursor");
      const int64_t n_items = nc * 21;
      const int64_t sb = (n_items + 7) / 8;
      k_sample_replies<<<(int)(sb < 148 * 16 ? sb : 148 * 16), 256, 0, s>>>((const uint2*)(w + L.o_pool), red_off, red_cnt, n_items, sample_cap, sample_seed,
                                                                            c0 * 21, (uint2*)(w + L.o_pool2), (long long*)(w + L.o_off2),
                                                                            (int32_t*)(w + L.o_cnt2), cursor2);
      EvalArgs ev{m.boards, m.players, nullptr, nullptr, 0, (const int64_t*)cursor2, L.rows, a.prepared, a.H, (float*)(w + L.o_val)};
      ev.codes = (const uint2*)(w + L.o_pool2);
      rc = eval_launch(ev, s);
      if (rc != BG_OK) return rc;
      red_off = (const long long*)(w + L.o_off2);
      red_cnt = (const int32_t*)(w + L.o_cnt2);
    }
    k_reduce<<<grid, 256, 0, s>>>((const float*)(w + L.o_val), red_off, red_cnt, a.S, c0, nc,
                                  a.top_k, a.alpha, a.beta, a.out_score, (long long*)a.out_replies, (const int32_t*)(w + L.o_tot + 16),
                                  a.out_status, a.reply_counter);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return check_cuda(e, "two_ply launch");
  }
  return BG_OK;
}

}  // namespace bg
