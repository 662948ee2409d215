// movegen.cu -- legal-move generation for batches of (board, player, roll) items on sm_100a.
//
// Replaces get_all_possible_moves (reference src/backgammon/moves/generate_all_moves.py:7-90,
// handle_move_types.py:7-221, get_moves_one_die.py:13-251, conditions.py) + execute_full_move_on_board_copy
// (src/environments/env_helper.py:27-91).  Output order == the reference's (first-occurrence DFS order).
//
// Design (one warp per item, nothing but the result leaves the SM):
//   * a node is a 128-bit exact key: mover's 24 point counts as nibbles (3 words) + {hit-mask of opponent
//     blots (24b), mover bar (4b), mover off (4b)}; plus 2 words of sub-move history.  The opponent side is
//     root-constant except for hit blots, so the key identifies the resulting board exactly (no hash-only dedup).
//   * breadth-first by ply: the frontier of unique nodes after k sub-moves is kept ordered by the
//     lexicographically smallest sub-move index sequence that reaches it.  Expanding frontier nodes in order,
//     children in the reference's one-die order (move slots: 0..23 point moves ascending, 24 bar entry,
//     25 bear-off of the farthest checker, 26 exact bear-off), with first-occurrence dedup through a
//     shared-memory hash set, reproduces the reference's DFS first-occurrence order while visiting each
//     distinct intermediate board once (the reference re-expands ~5x-8x duplicates).
//   * inside a ply the work is lane-per-candidate: move sets are slot bitmasks computed lane-per-parent, and the
//     (parent, move) pairs are flattened over the warp so that every lane builds one child per round.
//   * non-doubles keep the reference's literal control flow (both die orders, singles only when an order has
//     no two-move play, quirk Q1 skip, shared seen-set, max-length filter).
//   * five capacity tiers (128 / 256 / 512 / 2048 nodes per ply in shared memory, 4096 in L2-resident global scratch);
//     an item overflowing a tier is queued for the next one.  Overflowing the last tier is BG_ERR_CAPACITY.
#include "movegen.cuh"
#include "movegen_dev.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace bg {

namespace {

// dynamic shared memory, addressed through this accessor everywhere (also inside the out-of-line level expansion) so that
// the compiler keeps shared-space loads/stores (LDS/STS) instead of generic ones
__device__ __forceinline__ uint32_t* dyn_smem() {
  extern __shared__ uint32_t bg_dyn_smem[];
  return bg_dyn_smem;
}

template <int CAP, bool GLOBAL, bool MOVES>
struct Frontier {
  static constexpr int NF = MOVES ? 6 : 4;  // words per node: k0..k3 key (+ m0 m1 sub-move history)
  uint32_t* gbase;                          // GLOBAL: [2][NF][CAP] in L2-resident scratch
  uint32_t soff;                            // !GLOBAL: word offset of [2][NF][CAP] in dynamic shared memory
  uint32_t toff;                            // word offset of the hash table [2 * CAP] in dynamic shared memory
  __device__ __forceinline__ uint32_t ld(int lvl, int f, int pos) const {
    if constexpr (GLOBAL)
      return __ldcg(gbase + (lvl * NF + f) * CAP + pos);
    else
      return dyn_smem()[soff + (lvl * NF + f) * CAP + pos];
  }
  __device__ __forceinline__ void st(int lvl, int f, int pos, uint32_t v) const {
    if constexpr (GLOBAL)
      __stcg(gbase + (lvl * NF + f) * CAP + pos, v);
    else
      dyn_smem()[soff + (lvl * NF + f) * CAP + pos] = v;
  }
  __device__ __forceinline__ uint32_t* table() const { return dyn_smem() + toff; }
};

template <int CAP, bool GLOBAL, bool MOVES>
__device__ __forceinline__ Node load_node(const Frontier<CAP, GLOBAL, MOVES>& F, int lvl, int pos) {
  Node p;
  p.k0 = F.ld(lvl, 0, pos);
  p.k1 = F.ld(lvl, 1, pos);
  p.k2 = F.ld(lvl, 2, pos);
  p.k3 = F.ld(lvl, 3, pos);
  if constexpr (MOVES) {
    p.m0 = F.ld(lvl, 4, pos);
    p.m1 = F.ld(lvl, 5, pos);
  } else {
    p.m0 = p.m1 = 0u;
  }
  return p;
}

template <int CAP, bool GLOBAL, bool MOVES>
__device__ __forceinline__ void store_node(const Frontier<CAP, GLOBAL, MOVES>& F, int lvl, int pos, const Node& c) {
  F.st(lvl, 0, pos, c.k0);
  F.st(lvl, 1, pos, c.k1);
  F.st(lvl, 2, pos, c.k2);
  F.st(lvl, 3, pos, c.k3);
  if constexpr (MOVES) {
    F.st(lvl, 4, pos, c.m0);
    F.st(lvl, 5, pos, c.m1);
  }
}

template <int CAP>
__device__ __forceinline__ void clear_table(uint32_t* tab, int lane) {
#pragma unroll 4
  for (int i = lane; i < 2 * CAP; i += 32) tab[i] = 0u;
  __syncwarp();
}

enum { ITEM_OK = 0, ITEM_OVERFLOW = 1, ITEM_BAD = 2 };
enum { MODE_EXPAND = 0, MODE_IDENTITY = 1 };

// THE inner loop (kept out of line so the kernel has exactly one copy of it: the fully inlined variant was 55 KB of
// SASS and stalled on instruction fetch).  Expands the parents src[src_off .. src_off+n_src) by `die` and appends their
// children to dst[.. n_dst) IN CANONICAL ORDER (parent order, then the reference's one-die move order).
//   1. lane-per-parent: each lane derives its parent's move set as a slot bitmask (pure bit arithmetic);
//   2. lane-per-candidate: the (parent, move) pairs of up to 32 parents are flattened (warp scan + shuffle binary
//      search + n-th-set-bit), so every lane builds one child per round regardless of how few moves a parent has;
//   3. dedup, first occurrence wins: equal children inside a round are resolved with match.any on the key hash
//      (verified on the full 128-bit key; a fingerprint collision falls back to one-lane-at-a-time rounds), earlier rounds
//      and earlier calls through the shared-memory open-addressing set `tab`;
//   4. survivors are appended by ballot compaction (order preserving) and inserted with one CAS each.
// MODE_IDENTITY re-emits the parents themselves (the reference's single-move fallback).
// Returns false on capacity overflow.  flags: bit0 some child existed, bit1 some child was appended.
template <int CAP, bool GLOBAL, bool MOVES>
__device__ __noinline__ bool expand_level(const Frontier<CAP, GLOBAL, MOVES> F, const Root r, int src_lvl, int src_off,
                                          int n_src, int dst_lvl, int& n_dst, bool dedup, int mode, int die, int depth, uint32_t& flags) {
  constexpr uint32_t TMASK = 2 * CAP - 1;
  uint32_t* const tab = F.table();
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  int n = n_dst;
  uint32_t fl = 0;
  for (int pb = 0; pb < n_src; pb += 32) {
    // ---- 1. move sets of up to 32 parents ---------------------------------------------------------------------------
    uint32_t vmw = 0;
    if (pb + lane < n_src) {
      if (mode == MODE_EXPAND) {
        Node p;
        p.k0 = F.ld(src_lvl, 0, src_off + pb + lane);
        p.k1 = F.ld(src_lvl, 1, src_off + pb + lane);
        p.k2 = F.ld(src_lvl, 2, src_off + pb + lane);
        p.k3 = F.ld(src_lvl, 3, src_off + pb + lane);
        vmw = move_mask(p, r, die);
      } else {
        vmw = 1u;
      }
    }
    const int cnt = __popc(vmw & 0x7ffffffu);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(BG_FULL, inc, o);
      if (lane >= o) inc += t;
    }
    const int total = __shfl_sync(BG_FULL, inc, 31);
    const int exc = inc - cnt;
    if (total) fl |= 1u;
    // ---- 2..4. rounds of 32 candidates ------------------------------------------------------------------------------------
    for (int b0 = 0; b0 < total; b0 += 32) {
      const int t = b0 + lane;
      const bool cv = t < total;
      int q = 0;  // owner parent lane: number of parents whose inclusive end <= t
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        const int v = __shfl_sync(BG_FULL, inc, (q + sft - 1) & 31);
        if (v <= t && q + sft <= 32) q += sft;
      }
      q &= 31;
      const int rank = t - __shfl_sync(BG_FULL, exc, q);
      const uint32_t vq = __shfl_sync(BG_FULL, vmw, q);
      Node c;
      c.k0 = c.k1 = c.k2 = c.k3 = c.m0 = c.m1 = 0u;
      if (cv) {
        const Node p = load_node(F, src_lvl, src_off + pb + q);
        if (mode == MODE_EXPAND)
          make_child(p, r, nth_set_bit(vq & 0x7ffffffu, rank), vq >> 27, die, depth, c);
        else
          c = p;
      }
      uint32_t h = 0, fp = 0;
      bool dup = false, collide = false;
      if (dedup) {
        h = mix32(c.k0, c.k1, c.k2, c.k3);
        fp = h >> 16;
        // equal children inside this round: lowest lane of each hash group leads, the others compare against it
        const uint32_t grp = __match_any_sync(BG_FULL, cv ? h : (0x80000000u ^ (uint32_t)lane ^ (h & 0u)));
        const uint32_t cvm = __ballot_sync(BG_FULL, cv);
        const int leader = __ffs(grp & cvm) - 1;
        const int ls = leader < 0 ? lane : leader;
        const uint32_t l0 = __shfl_sync(BG_FULL, c.k0, ls), l1 = __shfl_sync(BG_FULL, c.k1, ls), l2 = __shfl_sync(BG_FULL, c.k2, ls),
                       l3 = __shfl_sync(BG_FULL, c.k3, ls);
        const bool same = l0 == c.k0 && l1 == c.k1 && l2 == c.k2 && l3 == c.k3;
        dup = cv && ls != lane && same;
        collide = cv && ls != lane && !same;
      }
      // normally one pass over all lanes; after a 32-bit hash collision inside the round, one lane per pass (exact)
      const bool slow = __any_sync(BG_FULL, collide);
      const int npass = slow ? 32 : 1;
      for (int pass = 0; pass < npass; ++pass) {
        const bool act = cv && (slow ? lane == pass : !dup);
        uint32_t idx = h & TMASK;
        bool found = false;
        if (dedup && act) {
          while (true) {
            const uint32_t sw = tab[idx];
            if (!sw) break;
            if ((sw >> 16) == fp) {
              const int pp = (int)(sw & 0xffffu) - 1;
              if (F.ld(dst_lvl, 0, pp) == c.k0 && F.ld(dst_lvl, 1, pp) == c.k1 && F.ld(dst_lvl, 2, pp) == c.k2 &&
                  F.ld(dst_lvl, 3, pp) == c.k3) {
                found = true;
                break;
              }
            }
            idx = (idx + 1) & TMASK;
          }
        }
        const bool isnew = act && !found;
        const uint32_t bal = __ballot_sync(BG_FULL, isnew);
        const int add = __popc(bal);
        if (add) fl |= 2u;
        if (n + add > CAP) return false;
        if (isnew) {
          const int pos = n + __popc(bal & lt);
          store_node(F, dst_lvl, pos, c);
          if (dedup) {
            const uint32_t word = (fp << 16) | (uint32_t)(pos + 1);
            while (atomicCAS(&tab[idx], 0u, word) != 0u) idx = (idx + 1) & TMASK;
          }
        }
        n += add;
        __syncwarp();
      }
    }
  }
  n_dst = n;
  flags = fl;
  return true;
}

// Children of the ROOT for one die, lane == move slot (no flattening needed for a single parent, no dedup: the children of one
// node are pairwise distinct).  Stored compactly, in slot order, at dst[dst_pos ...); returns their number.
template <int CAP, bool GLOBAL, bool MOVES>
__device__ __forceinline__ int expand_root(const Frontier<CAP, GLOBAL, MOVES>& F, const Root& r, const Node& root, int die, int dst_lvl,
                                           int dst_pos, int lane) {
  const uint32_t vm = move_mask(root, r, die);
  const bool valid = lane < 27 && ((vm >> lane) & 1u);
  const uint32_t bal = __ballot_sync(BG_FULL, valid);
  if (valid) {
    Node c;
    make_child(root, r, lane, vm >> 27, die, 0, c);
    store_node(F, dst_lvl, dst_pos + __popc(bal & ((1u << lane) - 1u)), c);
  }
  __syncwarp();
  return __popc(bal);
}

// Result of one item: up to two index ranges [a0,a1) ++ [b0,b1) of frontier level `lvl`, in output order.
struct ItemOut {
  int lvl, a0, a1, b0, b1;
};

// Generates the ordered legal afterstate list of one item.
template <int CAP, bool GLOBAL, bool MOVES>
__device__ __forceinline__ int generate(const Frontier<CAP, GLOBAL, MOVES>& F, const uint32_t* rootw, int player, int d0,
                                        int d1, int lane, ItemOut& out) {
  // ---- root ------------------------------------------------------------------------------------------
  const int ob = player * 6, pb = (1 - player) * 6;
  uint32_t bad = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) bad |= rootw[i] & 0xf0f0f0f0u;
  const uint32_t w12 = rootw[12];
  bad |= w12 & 0xf0f0f0f0u;
  if (bad || d0 < 1 || d0 > 6 || d1 < 1 || d1 > 6) return ITEM_BAD;
  Node root;
  root.k0 = bytes_to_nib4(rootw[ob + 0]) | (bytes_to_nib4(rootw[ob + 1]) << 16);
  root.k1 = bytes_to_nib4(rootw[ob + 2]) | (bytes_to_nib4(rootw[ob + 3]) << 16);
  root.k2 = bytes_to_nib4(rootw[ob + 4]) | (bytes_to_nib4(rootw[ob + 5]) << 16);
  const uint32_t own_bar = (w12 >> (8 * player)) & 15u, own_off = (w12 >> (16 + 8 * player)) & 15u;
  root.k3 = (own_bar << 24) | (own_off << 28);
  root.m0 = root.m1 = 0u;
  Root r;
  r.player = player;
  r.dirsign = player == 0 ? 1 : -1;
  r.home = player == 0 ? 0xfc0000u : 0x00003fu;
  {
    uint32_t oc = 0;
    if (lane < 24) oc = (rootw[pb + (lane >> 2)] >> ((lane & 3) * 8)) & 0xffu;
    r.blocked = __ballot_sync(BG_FULL, oc >= 2) & 0xffffffu;
    r.blot = __ballot_sync(BG_FULL, oc == 1) & 0xffffffu;
    uint32_t total = own_bar + own_off;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const uint32_t w = rootw[ob + i];
      total += (w & 0xff) + ((w >> 8) & 0xff) + ((w >> 16) & 0xff) + (w >> 24);
    }
    r.valid15 = total == 15u;
  }
  if (lane == 0) store_node(F, 0, 0, root);
  __syncwarp();
  uint32_t fl;
  uint32_t* const tab = F.table();
  const bool doubles = d0 == d1;
  const int hi = d0 > d1 ? d0 : d1, lo = d0 > d1 ? d1 : d0;
  clear_table<CAP>(tab, lane);
  if (doubles) {
    // ---- doubles: BFS by ply with per-ply dedup (handle_move_types.py:84-193) ------------------------
    int cur = 1, depth_done = 0;
    int n_cur = expand_root(F, r, root, d0, 1, 0, lane);
    if (n_cur > 0) {
      depth_done = 1;
      for (int depth = 1; depth < 4; ++depth) {
        if (depth > 1) clear_table<CAP>(tab, lane);
        int n_next = 0;
        if (!expand_level(F, r, cur, 0, n_cur, cur ^ 1, n_next, true, MODE_EXPAND, d0, depth, fl)) return ITEM_OVERFLOW;
        if (n_next == 0) break;
        cur ^= 1;
        n_cur = n_next;
        depth_done = depth + 1;
      }
    }
    out.lvl = cur;
    out.a0 = 0;
    out.a1 = depth_done == 0 ? 0 : n_cur;
    out.b0 = out.b1 = 0;
    return ITEM_OK;
  }
  // ---- non-doubles: literal two-order control flow (generate_all_moves.py:25-53, handle_move_types.py:7-81) ----
  // level 0: root at [0], the first-die boards at [1, 1+n1); level 1: the result list (shared seen-set `tab`).
  // Entries are appended in two phases (one per die order), each of uniform length (2 sub-moves, or 1 when the order
  // has no two-move play), so the max-sub-moves filter (generate_all_moves.py:69-90) is a choice of phase ranges.
  int n_res = 0, n_a = 0, len_a = 0, len_b = 0;
  for (int order = 0; order < 2; ++order) {
    const int dA = order == 0 ? hi : lo, dB = order == 0 ? lo : hi;
    const int n1 = expand_root(F, r, root, dA, 0, 1, lane);  // stored after the root; pairwise distinct -> no dedup
    if (!expand_level(F, r, 0, 1, n1, 1, n_res, true, MODE_EXPAND, dB, 1, fl)) return ITEM_OVERFLOW;
    int len = 2;
    if (!(fl & 1u)) {  // no two-move sequence in this order: singles, in first-die order (handle_move_types.py:70-81)
      len = 1;
      if (!expand_level(F, r, 0, 1, n1, 1, n_res, true, MODE_IDENTITY, dA, 0, fl)) return ITEM_OVERFLOW;
    }
    if (order == 0) {
      n_a = n_res;
      len_a = len;
      // quirk Q1 (generate_all_moves.py:40-50): reverse order skipped iff exactly one 1-sub-move result
      if (n_res == 1 && len == 1) break;
    } else {
      len_b = len;
    }
  }
  // a phase that appended nothing has no entries and does not define the maximum
  const int la = n_a > 0 ? len_a : 0, lb = n_res > n_a ? len_b : 0;
  const int maxlen = la > lb ? la : lb;
  out.lvl = 1;
  out.a0 = 0;
  out.a1 = la == maxlen ? n_a : 0;
  out.b0 = n_a;
  out.b1 = lb == maxlen ? n_res : n_a;
  return ITEM_OK;
}

template <int CAP, bool GLOBAL, bool MOVES, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_movegen(MovegenParams P) {
  constexpr int NF = MOVES ? 6 : 4;
  constexpr int FRONT_WORDS = GLOBAL ? 0 : 2 * NF * CAP;
  constexpr int PER_WARP = FRONT_WORDS + 2 * CAP + 16;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  Frontier<CAP, GLOBAL, MOVES> F;
  F.soff = (uint32_t)(wib * PER_WARP);
  F.toff = F.soff + FRONT_WORDS;
  F.gbase = nullptr;
  if constexpr (GLOBAL) F.gbase = P.gfront + (size_t)(blockIdx.x * WARPS + wib) * (2 * NF * CAP);
  uint32_t* const rootw = dyn_smem() + F.toff + 2 * CAP;

  const int64_t n_items = P.in_list ? (int64_t)(*P.in_count) : P.B;
  while (true) {
    int t0 = 0;
    if (lane == 0) t0 = atomicAdd(P.item_counter, P.grab);
    t0 = __shfl_sync(BG_FULL, t0, 0);
    if (t0 >= n_items) break;
    for (int g = 0; g < P.grab; ++g) {
      const int64_t t = (int64_t)t0 + g;
      if (t >= n_items) break;
      const int item = P.in_list ? P.in_list[t] : (int)t;
      const int src = P.all_rolls ? item / 21 : item;  // position-major: 21 items share one board
      if (P.active && !P.active[src]) {
        if (lane == 0) {
          P.out_count[item] = 0;
          P.out_offsets[item] = 0;
        }
        continue;
      }
      __syncwarp();
      if (lane < 13) rootw[lane] = reinterpret_cast<const uint32_t*>(P.boards)[(int64_t)src * 13 + lane];
      __syncwarp();
      const int player = P.players[src] & 1;
      int d0, d1;
      if (P.all_rolls) {  // roll index -> (d0 <= d1), lexicographic (src/multi/two_ply.py:10-32)
        int ri = item - src * 21;
        d0 = 1;
        while (ri >= 7 - d0) {
          ri -= 7 - d0;
          ++d0;
        }
        d1 = d0 + ri;
      } else {
        d0 = P.rolls[2 * (int64_t)item];
        d1 = P.rolls[2 * (int64_t)item + 1];
      }
      ItemOut io;
      const int rc = generate<CAP, GLOBAL, MOVES>(F, rootw, player, d0, d1, lane, io);
      if (rc == ITEM_OVERFLOW || (rc == ITEM_OK && P.out_codes)) {  // compact mode: this kernel writes boards, not codes -> the item fails
        if (lane == 0) {
          if (P.ovf_list && !P.out_codes) {
            const int q = atomicAdd(P.ovf_count, 1);
            P.ovf_list[q] = item;
          } else {
            P.out_count[item] = -1;
            P.out_offsets[item] = -1;
            atomicMin(P.status, BG_ERR_CAPACITY);
          }
        }
        continue;
      }
      if (rc == ITEM_BAD) {
        if (lane == 0) {
          P.out_count[item] = 0;
          P.out_offsets[item] = -1;
          atomicMin(P.status, BG_ERR_INVARIANT);
        }
        continue;
      }
      // ---- emit --------------------------------------------------------------------------------------
      const int lvl = io.lvl, na = io.a1 - io.a0;
      const int n = na + (io.b1 - io.b0);
      const int n_keep = n < P.item_cap ? n : P.item_cap;
      long long base = 0;
      if (lane == 0) {
        base = n_keep ? (long long)atomicAdd(P.pool_cursor, (unsigned long long)n_keep) : 0ll;
        if (base + n_keep > P.pool_cap) {
          base = -1;
          atomicMin(P.status, BG_ERR_CAPACITY);
        }
        P.out_count[item] = n;
        P.out_offsets[item] = base;
      }
      base = __shfl_sync(BG_FULL, base, 0);
      if (base < 0 || n_keep == 0) continue;
      // Boards are rebuilt lane-per-board into a shared staging area (13-word rows: odd stride, conflict free) and then
      // copied out with fully coalesced word stores.  Staging reuses memory that is dead after generation: the frontier
      // level that does not hold the result (shared tiers) or the hash table (global tier).
      uint32_t* ob = reinterpret_cast<uint32_t*>(P.out_boards) + base * 13;
      uint32_t* const stage = GLOBAL ? F.table() : dyn_smem() + F.soff + (lvl ^ 1) * NF * CAP;
      const uint32_t w12 = rootw[12];
      const uint32_t opp_bar = (w12 >> (8 * (1 - player))) & 0xffu, opp_off = (w12 >> (16 + 8 * (1 - player))) & 0xffu;
      const int ownw = player * 6, oppw = (1 - player) * 6;
      for (int b0 = 0; b0 < n_keep; b0 += 32) {
        const int bo = b0 + lane;
        __syncwarp();
        if (bo < n_keep) {
          const int b = bo < na ? io.a0 + bo : io.b0 + (bo - na);
          const uint32_t k0 = F.ld(lvl, 0, b), k1 = F.ld(lvl, 1, b), k2 = F.ld(lvl, 2, b), k3 = F.ld(lvl, 3, b);
          uint32_t* row = stage + lane * 13;
          row[ownw + 0] = nib4_to_bytes(k0 & 0xffffu);
          row[ownw + 1] = nib4_to_bytes(k0 >> 16);
          row[ownw + 2] = nib4_to_bytes(k1 & 0xffffu);
          row[ownw + 3] = nib4_to_bytes(k1 >> 16);
          row[ownw + 4] = nib4_to_bytes(k2 & 0xffffu);
          row[ownw + 5] = nib4_to_bytes(k2 >> 16);
#pragma unroll
          for (int q = 0; q < 6; ++q) row[oppw + q] = rootw[oppw + q] - bits4_to_bytes((k3 >> (4 * q)) & 0xfu);
          const uint32_t ownbar = (k3 >> 24) & 15u, ownoff = k3 >> 28, ob2 = opp_bar + __popc(k3 & 0xffffffu);
          row[12] = player == 0 ? (ownbar | (ob2 << 8) | (ownoff << 16) | (opp_off << 24))
                                : (ob2 | (ownbar << 8) | (opp_off << 16) | (ownoff << 24));
        }
        __syncwarp();
        const int nw = (n_keep - b0 < 32 ? n_keep - b0 : 32) * 13;
        for (int tt = lane; tt < nw; tt += 32) ob[b0 * 13 + tt] = stage[tt];
      }
      if (P.out_owner)
        for (int b = lane; b < n_keep; b += 32) P.out_owner[base + b] = item;
      if (P.out_flags)
        for (int b = lane; b < n_keep; b += 32) P.out_flags[base + b] = (uint8_t)player;
      if constexpr (MOVES) {
        uint32_t* os = reinterpret_cast<uint32_t*>(P.out_submoves) + base * 3;
        for (int bo = lane; bo < n_keep; bo += 32) {
          const int b = bo < na ? io.a0 + bo : io.b0 + (bo - na);
          const uint32_t m[2] = {F.ld(lvl, 4, b), F.ld(lvl, 5, b)};
          uint8_t by[12];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t sm = (m[q >> 1] >> (16 * (q & 1))) & 0xffffu;
            const bool ok = (sm >> 11) & 1u;
            by[3 * q] = ok ? (uint8_t)(sm & 31u) : 255;
            by[3 * q + 1] = ok ? (uint8_t)((sm >> 5) & 31u) : 255;
            by[3 * q + 2] = ok ? (uint8_t)((sm >> 10) & 1u) : 0;
          }
#pragma unroll
          for (int q = 0; q < 3; ++q)
            os[bo * 3 + q] = by[4 * q] | (by[4 * q + 1] << 8) | (by[4 * q + 2] << 16) | ((uint32_t)by[4 * q + 3] << 24);
        }
      }
    }
  }
}

// capacity tiers: nodes per ply.  T1-T4 keep the frontiers in shared memory, T5 in L2-resident global scratch.  The small tiers buy
// occupancy (more resident warps for the many light items), the 2048 tier buys latency for the rare very wide doubles trees.
constexpr int T1_CAP = 128, T1_WARPS = 4, T1_CTAS_PER_SM = 8;
constexpr int T2_CAP = 256, T2_WARPS = 1, T2_CTAS_PER_SM = 20;
constexpr int T3_CAP = 512, T3_WARPS = 1, T3_CTAS_PER_SM = 10;
constexpr int T4_CAP = 2048, T4_WARPS = 1, T4_CTAS_PER_SM = 2;
constexpr int T5_CAP = 4096, T5_WARPS = 2, T5_CTAS_PER_SM = 3;
constexpr int NUM_SMS = 148;

constexpr size_t smem_bytes(int cap, bool global, bool moves, int warps) {
  return (size_t)warps * ((global ? 0 : 2 * (moves ? 6 : 4) * cap) + 2 * cap + 16) * 4;
}

constexpr int64_t HDR_BYTES = 256;
constexpr int64_t GFRONT_BYTES = (int64_t)NUM_SMS * T5_CTAS_PER_SM * T5_WARPS * 2 * 6 * T5_CAP * 4;
constexpr int N_OVF = 4;  // overflow lists chaining the five tiers

template <int CAP, bool GLOBAL, bool MOVES, int WARPS, int CTAS>
int32_t prepare_tier() {
  static DeviceOnce once;
  return once.run([]() -> int32_t {
    return check_cuda(opt_in_shared(k_movegen<CAP, GLOBAL, MOVES, WARPS, CTAS>, smem_bytes(CAP, GLOBAL, MOVES, WARPS)), "cudaFuncSetAttribute(k_movegen)");
  });
}

// a tail tier: consumes overflow list `in`, produces list `out` (out < 0: last tier), uses work counter `in + 1`
template <int CAP, bool GLOBAL, bool MOVES, int WARPS, int CTAS>
int32_t launch_tail_tier(MovegenParams P, int in, int out, int ctas, int32_t* ctr, int32_t* const (&ovf)[N_OVF], int32_t* const (&ovf_n)[N_OVF],
                         cudaStream_t stream) {
  int32_t rc = prepare_tier<CAP, GLOBAL, MOVES, WARPS, CTAS>();
  if (rc != BG_OK) return rc;
  P.item_counter = ctr + in + 1;
  P.in_list = ovf[in];
  P.in_count = ovf_n[in];
  P.ovf_list = out >= 0 ? ovf[out] : nullptr;
  P.ovf_count = out >= 0 ? ovf_n[out] : nullptr;
  P.grab = 1;
  k_movegen<CAP, GLOBAL, MOVES, WARPS, CTAS><<<NUM_SMS * ctas, WARPS * 32, smem_bytes(CAP, GLOBAL, MOVES, WARPS), stream>>>(P);
  return BG_OK;
}

// Longest-first order for SMALL per-item batches (one self-play ply): a double's tree costs several times a non-double's, and the warps take
// items dynamically, so the last items decide when the bulk tier ends.  The items with d0 == d1 go to the front of the list, the others
// fill it from the back.  cursors[0] / cursors[1] (zeroed with the workspace header) count the two classes; *count = B.
__global__ void __launch_bounds__(256) k_order_items(const uint8_t* __restrict__ rolls, int64_t B, int32_t* __restrict__ list, int32_t* __restrict__ count,
                                                     int32_t* __restrict__ cursors) {
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = (int32_t)B;
  for (int64_t base = ((int64_t)blockIdx.x * 256 + threadIdx.x) & ~31ll; base < B; base += (int64_t)gridDim.x * 256) {
    const int64_t i = base + lane;
    const bool in = i < B;
    const bool dbl = in && rolls[2 * i] == rolls[2 * i + 1];
    const uint32_t md = __ballot_sync(BG_FULL, dbl), mn = __ballot_sync(BG_FULL, in && !dbl);
    int bd = 0, bn = 0;
    if (lane == 0) {
      if (md) bd = atomicAdd(&cursors[0], __popc(md));
      if (mn) bn = atomicAdd(&cursors[1], __popc(mn));
    }
    bd = __shfl_sync(BG_FULL, bd, 0);
    bn = __shfl_sync(BG_FULL, bn, 0);
    const uint32_t below = (1u << lane) - 1u;
    if (dbl) list[bd + __popc(md & below)] = (int32_t)i;
    else if (in) list[B - 1 - (bn + __popc(mn & below))] = (int32_t)i;
  }
}

template <bool MOVES>
int32_t launch_tiers(MovegenParams P, int64_t B, int32_t* ctr, int32_t* const (&ovf)[N_OVF], int32_t* const (&ovf_n)[N_OVF],
                     int64_t* tier1_total, cudaEvent_t tier1_event, int32_t tier2_ctas, cudaStream_t stream) {
  int32_t rc = BG_OK;
  // tier 1: every item
  P.item_counter = ctr + 0;
  P.in_list = nullptr;
  P.in_count = nullptr;
  P.ovf_list = ovf[0];
  P.ovf_count = ovf_n[0];
  // Position-major batches: the bulk tier is the code-based generator of movegen21.cu, one warp per position expanding all 21 rolls.
  // Per-item batches: the 128-node frontier kernel stays the bulk tier (a warp per item amortises nothing of the code-based kernel's set-up),
  // but what overflows it -- the wide doubles trees -- goes to the code-based generator in single-roll mode, which prunes their
  // duplicate candidates and holds ~670 results per item, instead of the 512- and 2048-node frontier tiers.  Only the FullMove sub-move
  // histories (MOVES) still use the frontier tiers throughout.
  const bool code_tiers = !MOVES && P.item_cap >= MOVEGEN21_MIN_ITEM_CAP && !getenv("BG_MOVEGEN_LEGACY");
  const bool fast21 = code_tiers && P.all_rolls;
  if (fast21) {
    P.grab = 1;
    if (const char* dbg = getenv("BG_MG21_DEBUG")) P.grab = atoi(dbg) | 1;  // development: bit 1 skip non-doubles, bit 2 skip doubles, bit 3 skip board stores
    P.B = B / 21;
    rc = movegen21_launch_kernel(P, stream);
    P.B = B;
    if (rc != BG_OK) return rc;
  } else {
    if ((rc = prepare_tier<T1_CAP, false, MOVES, T1_WARPS, T1_CTAS_PER_SM>()) != BG_OK) return rc;
    P.grab = B > (1 << 20) ? 8 : 1;
    if (B < (1 << 20) && B >= 4096 && P.rolls && !getenv("BG_MOVEGEN_NO_ORDER")) {
      // overflow list 1 is only used by the 256-node tier of batches >= 2^20 items: free here for the ordered item list
      k_order_items<<<(int)((B + 255) / 256 < 592 ? (B + 255) / 256 : 592), 256, 0, stream>>>(P.rolls, B, ovf[1], ovf_n[1], ctr + 12);
      P.in_list = ovf[1];
      P.in_count = ovf_n[1];
    }
    int64_t want = (B + T1_WARPS - 1) / T1_WARPS;
    int grid = (int)(want < (int64_t)NUM_SMS * T1_CTAS_PER_SM ? want : (int64_t)NUM_SMS * T1_CTAS_PER_SM);
    k_movegen<T1_CAP, false, MOVES, T1_WARPS, T1_CTAS_PER_SM><<<grid, T1_WARPS * 32, smem_bytes(T1_CAP, false, MOVES, T1_WARPS), stream>>>(P);
  }
  if (tier1_total) {
    cudaError_t e = cudaMemcpyAsync(tier1_total, P.pool_cursor, 8, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return check_cuda(e, "copy tier1_total");
  }
  if (tier1_event) {
    cudaError_t e = cudaEventRecord(tier1_event, stream);
    if (e != cudaSuccess) return check_cuda(e, "record tier1_event");
  }
  // tail tiers: the items that overflowed 128 / 256 / 512 / 2048 nodes in some ply.  The 256 tier (twice the resident warps of the 512
  // tier) pays when the overflow lists are long; for small batches (one self-play ply) every extra tier is one more kernel whose
  // slowest item sits on the critical path, so the chain goes 128 -> 512 directly.
  const bool use256 = B >= (1 << 20) && !fast21;  // what the position-major tier hands over is wider than 224 nodes per ply
  const int c2 = tier2_ctas > 0 && tier2_ctas < T2_CTAS_PER_SM ? tier2_ctas : T2_CTAS_PER_SM;
  if (use256 && (rc = launch_tail_tier<T2_CAP, false, MOVES, T2_WARPS, T2_CTAS_PER_SM>(P, 0, 1, c2, ctr, ovf, ovf_n, stream)) != BG_OK) return rc;
  auto code_tier = [&](int in, int out, int big) -> int32_t {  // the code-based generator (movegen21.cu) consuming overflow list `in`
    MovegenParams Q = P;
    Q.item_counter = ctr + in + 1;
    Q.in_list = ovf[in];
    Q.in_count = ovf_n[in];
    Q.ovf_list = ovf[out];
    Q.ovf_count = ovf_n[out];
    Q.grab = 1;
    return movegen21_launch_kernel(Q, stream, big);
  };
  if (code_tiers) {
    // per-item batches: the standard configuration as the middle tier (many resident warps for the long overflow list of the 128-node tier);
    // then, for both batch shapes, the big configuration (one warp per CTA, ~6,000 results per item) for the few trees that are wider still
    if (!P.all_rolls && (rc = code_tier(use256 ? 1 : 0, 2, 0)) != BG_OK) return rc;
    if ((rc = code_tier(P.all_rolls ? 0 : 2, 3, 1)) != BG_OK) return rc;
  } else {
    if ((rc = launch_tail_tier<T3_CAP, false, MOVES, T3_WARPS, T3_CTAS_PER_SM>(P, use256 ? 1 : 0, 2, T3_CTAS_PER_SM, ctr, ovf, ovf_n, stream)) != BG_OK)
      return rc;
    if ((rc = launch_tail_tier<T4_CAP, false, MOVES, T4_WARPS, T4_CTAS_PER_SM>(P, 2, 3, T4_CTAS_PER_SM, ctr, ovf, ovf_n, stream)) != BG_OK) return rc;
  }
  if ((rc = launch_tail_tier<T5_CAP, true, MOVES, T5_WARPS, T5_CTAS_PER_SM>(P, 3, -1, T5_CTAS_PER_SM, ctr, ovf, ovf_n, stream)) != BG_OK) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_movegen launch");
  return BG_OK;
}

}  // namespace

int64_t movegen_workspace_bytes(int64_t B) {
  int64_t lists = ((N_OVF * B * 4 + 255) / 256) * 256;
  return HDR_BYTES + lists + GFRONT_BYTES;
}

// workspace header layout (first HDR_BYTES): [0] u64 pool cursor, [8] i32 status, [12] i32 counter1,
// [16] .. [28] i32 counters of tiers 2-5, [32] [36] [40] [44] i32 overflow-list counts of tiers 1-4
int32_t movegen_launch(const MovegenArgs& a, cudaStream_t stream) {
  if (a.B < 0 || a.B >= (1ll << 31) || a.item_cap < 0 || a.pool_cap < 0) {
    set_error("bg_movegen: bad sizes (B=%lld item_cap=%d pool_cap=%lld)", (long long)a.B, a.item_cap, (long long)a.pool_cap);
    return BG_ERR_ARG;
  }
  if (a.out_codes && (!a.all_rolls || a.out_submoves || getenv("BG_MOVEGEN_LEGACY"))) {
    set_error("bg_movegen: compact (code) output needs a position-major batch without sub-moves");
    return BG_ERR_ARG;
  }
  const int64_t n_items = a.all_rolls ? a.B * 21 : a.B;
  if (n_items >= (1ll << 31)) {
    set_error("bg_movegen: too many items (%lld)", (long long)n_items);
    return BG_ERR_ARG;
  }
  if (a.workspace_bytes < movegen_workspace_bytes(n_items)) {
    set_error("bg_movegen: workspace too small (%lld < %lld)", (long long)a.workspace_bytes,
              (long long)movegen_workspace_bytes(n_items));
    return BG_ERR_ARG;
  }
  char* ws = (char*)a.workspace;
  cudaError_t e = cudaMemsetAsync(ws, 0, HDR_BYTES, stream);
  if (e != cudaSuccess) return check_cuda(e, "cudaMemsetAsync(workspace)");
  int64_t lists = ((N_OVF * n_items * 4 + 255) / 256) * 256;
  MovegenParams P;
  P.boards = a.boards;
  P.players = a.players;
  P.rolls = a.rolls;
  P.B = n_items;
  P.item_cap = a.item_cap;
  P.pool_cap = a.pool_cap;
  P.out_boards = a.out_boards;
  P.out_submoves = a.out_submoves;
  P.out_owner = a.out_owner;
  P.out_flags = a.out_flags;
  P.out_offsets = (long long*)a.out_offsets;
  P.out_count = a.out_count;
  P.pool_cursor = (unsigned long long*)ws;
  P.status = (int32_t*)(ws + 8);
  P.gfront = (uint32_t*)(ws + HDR_BYTES + lists);
  P.active = a.active;
  P.all_rolls = a.all_rolls;
  P.out_codes = a.out_codes;
  int32_t* ctr = (int32_t*)(ws + 12);
  int32_t* const l0 = (int32_t*)(ws + HDR_BYTES);
  int32_t* const ovf[N_OVF] = {l0, l0 + n_items, l0 + 2 * n_items, l0 + 3 * n_items};
  int32_t* const ovf_n[N_OVF] = {(int32_t*)(ws + 32), (int32_t*)(ws + 36), (int32_t*)(ws + 40), (int32_t*)(ws + 44)};
  if (a.B > 0) {
    // the sub-move history (2 extra words per node) is only carried when the caller asks for the FullMove sequences
    int32_t rc = a.out_submoves ? launch_tiers<true>(P, n_items, ctr, ovf, ovf_n, a.tier1_total, a.tier1_event, a.tier2_ctas_per_sm, stream)
                                : launch_tiers<false>(P, n_items, ctr, ovf, ovf_n, a.tier1_total, a.tier1_event, a.tier2_ctas_per_sm, stream);
    if (rc != BG_OK) return rc;
  }
  if (a.out_total) {
    e = cudaMemcpyAsync(a.out_total, ws, 8, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return check_cuda(e, "copy out_total");
  }
  if (a.out_status) {
    e = cudaMemcpyAsync(a.out_status, ws + 8, 4, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return check_cuda(e, "copy out_status");
  }
  return BG_OK;
}

}  // namespace bg
