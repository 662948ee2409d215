// movegen21.cu -- position-major legal-move generation on sm_100a: ONE WARP PER POSITION, ALL 21 ROLLS.
//
// Same contract as movegen.cu (get_all_possible_moves for every roll, reference src/backgammon/moves/generate_all_moves.py:7-90,
// handle_move_types.py:7-221, + execute_full_move_on_board_copy, src/environments/env_helper.py:27-91; output order == the
// reference's action order), for the workloads that ask for every roll of a position: BASELINE config 2 (positions x 21 rolls) and
// the 2-ply lookahead (src/multi/two_ply.py:93-150: 21 opponent rolls per candidate).  Item index = position * 21 + roll index
// (roll order of src/multi/two_ply.py:10-32: (1,1), (1,2), ... (6,6)).
//
// What the position-major formulation buys over one warp per (board, roll) item:
//   * the root is decoded once, the one-die move sets of the six dice are computed once, and each of the <= 64 first-move
//     children gets its six second-die move sets in one lane-per-child pass: every non-double (hi, lo) reads its two die orders
//     out of that table, every double d-d its first two plies;
//   * a result is a 30-bit CODE, not a board: for a non-double the sorted sources / sorted destinations of the two sub-moves
//     after cancelling a point that is both, plus the one intermediate point whose blot was hit; for a double the sorted
//     multiset of (up to four) source points.  Equal codes <=> equal boards (the mover's configuration is root - sources +
//     destinations and every landing point that holds a blot is hit), so first-occurrence dedup is an exact one-word hash set,
//     and boards are only materialised for the survivors, at emit time;
//   * candidates that provably duplicate an EARLIER candidate are never generated (tests/tools/proto_movegen21.py checks both rules
//     against the oracle): in a non-double's reverse order, the second sub-move from a point that could already move first
//     (the commuting pair was produced by the forward order); in a doubles tree, a slot below the parent's last slot whose move
//     was already legal one ply earlier (the same multiset is reached through an earlier parent).  What is left is ~1.1x the
//     unique results instead of ~2x / ~4x;
//   * candidates of all rolls are flattened over the warp (parent batches of 32, one lane per candidate in canonical order), so
//     lanes stay busy even though a single (position, roll) item has ~14 results;
//   * the six doubles trees advance ply by ply TOGETHER (their nodes carry the roll id), in groups sized from the ply-2 counts so that
//     frontier + results fit the warp's arena; a group that still overflows is retried one die at a time;
//   * results are buffered per warp and emitted in bulk: one pool reservation per flush (~3 per position instead of 21),
//     rows rebuilt from the root bytes + code in shared memory and copied out with 16-byte stores.
// Anything that does not fit the fast path's per-warp capacities (a single doubles tree wider than the arena, more than
// C1CAP first moves) is queued, per item, for the generic capacity tiers of movegen.cu.
#include "movegen.cuh"
#include "movegen_dev.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace bg {

namespace {

constexpr int C1CAP = 64;      // first-move children over the six dice (observed max 70 in 20,000 positions, p99 44)
constexpr int DCAP = 128;      // candidate descriptors per sub-batch
constexpr uint32_t NONE5 = 31u;
// -DBG_CHECKED=1: index / capacity assertions on every shared-memory write of this file (compute-sanitizer is closed on the GPU pool this was
// developed on, so the checked build is the memcheck stand-in: the GPU suite is run once with it, profiles/r02_checked_build_gpu_tests.txt)
#ifndef BG_CHECKED
#define BG_CHECKED 0
#endif
#if BG_CHECKED
#define BG_CHECK(cond, what)                                                                                         \
  do {                                                                                                               \
    if (!(cond)) {                                                                                                   \
      printf("BG_CHECK failed: %s (%s:%d) block %d thread %d\n", what, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
      asm volatile("trap;");                                                                                          \
    }                                                                                                                \
  } while (0)
#else
#define BG_CHECK(cond, what) \
  do {                       \
  } while (0)
#endif
#ifndef BG_EMIT_NIBBLE
#define BG_EMIT_NIBBLE 0
#endif

// per-warp shared memory (words), common part
constexpr int O_ROOT = 0;                       // 16: the 13 root words
constexpr int O_C1INFO = O_ROOT + 16;           // C1CAP: slot | src << 5 | dst << 10 | die0 << 15 | lone << 18
constexpr int O_C1MASK = O_C1INFO + C1CAP;      // 6 x C1CAP: second-die move sets of each child, [die0][child]
constexpr int O_DESC = O_C1MASK + 6 * C1CAP;    // DCAP u16: parent lane | slot << 5
constexpr int O_TAB = O_DESC + DCAP / 2;        // TCAP: hash set; doubles as the emit staging area (32 rows x 13 words + 3)

// Capacity configuration of one instantiation of the kernel.  `Std` is the bulk tier (many resident warps, ~10 KB each); `Big` is the tail
// tier for the few items whose doubles tree does not fit it (one warp per CTA, 68 KB): up to ~6,000 results per item, i.e. everything
// below BG_MAX_ITEM_MOVES that the 512- / 2048-node frontier tiers of movegen.cu used to take, at a third of their candidates.
template <int TLOG2, int BCAP_, int F2CAP_, int RESERVE_, int WARPS_, int CTAS_>
struct Cfg21 {
  static constexpr int TCAP_LOG2 = TLOG2;
  static constexpr int TCAP = 1 << TLOG2;          // hash-set slots
  static constexpr int TLIVE = TCAP * 5 / 8;       // most live keys the hash set is allowed to hold
  static constexpr int BCAP = BCAP_;               // arena: buffered result codes (non-doubles) / ply-2 frontier + ply-3 frontier + results (doubles)
  static constexpr int F2CAP = F2CAP_;             // most ply-2 nodes of the doubles trees together
  static constexpr int RES_RESERVE = RESERVE_;     // arena words kept for results behind the frontiers
  static constexpr int WARPS = WARPS_, CTAS = CTAS_;
  static constexpr int O_BUF = O_TAB + TCAP;       // BCAP: the arena
  static constexpr int O_ISTART = O_BUF + BCAP;    // 24 + 24: per-item result ranges of a flush
  static constexpr int WARP_WORDS = O_ISTART + 48;
  static_assert(TCAP >= 32 * 13 + 4, "staging area");
  static_assert(O_TAB % 4 == 0 && WARP_WORDS % 4 == 0, "16-byte alignment of the staging area");
  static constexpr size_t SMEM_BYTES = (size_t)WARPS * WARP_WORDS * 4;
};
#ifndef BG21_WARPS
#define BG21_WARPS 4
#endif
#ifndef BG21_CTAS
#define BG21_CTAS 5
#endif
#ifndef BG21_TLOG2
#define BG21_TLOG2 10
#define BG21_BCAP 1024
#define BG21_F2 512
#define BG21_RES 352
#endif
using Std21 = Cfg21<BG21_TLOG2, BG21_BCAP, BG21_F2, BG21_RES, BG21_WARPS, BG21_CTAS>;
using Big21 = Cfg21<13, 8192, 2048, 2560, 1, 3>;

// the 15 non-double rolls in roll-index order: 0-based (lo, hi) dice, 3 bits each
constexpr uint64_t pack15(const int (&v)[15]) {
  uint64_t r = 0;
  for (int i = 0; i < 15; ++i) r |= (uint64_t)v[i] << (3 * i);
  return r;
}
constexpr int ND_LO_V[15] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4};
constexpr int ND_HI_V[15] = {1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};
constexpr uint64_t ND_LO = pack15(ND_LO_V), ND_HI = pack15(ND_HI_V);
constexpr uint32_t DBL_IDS = (1u << 0) | (1u << 6) | (1u << 11) | (1u << 15) | (1u << 18) | (1u << 20);  // roll indices of d-d

template <class C>
__device__ __forceinline__ uint32_t* wsm() {
  extern __shared__ uint32_t bg_dyn_smem21[];
  return bg_dyn_smem21 + (threadIdx.x >> 5) * C::WARP_WORDS;
}

// occupancy / state summary of a node, from which the move set of any die is a few bit operations (same rules as move_mask)
struct View {
  uint32_t occ, bar, off, last;
  bool bear;
};

__device__ __forceinline__ View make_view(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, const Root& r) {
  View v;
  v.occ = nib_occupancy(k0) | (nib_occupancy(k1) << 8) | (nib_occupancy(k2) << 16);
  v.bar = (k3 >> 24) & 15u;
  v.off = k3 >> 28;
  v.bear = r.valid15 && (v.occ & ~r.home) == 0;
  v.last = r.player == 0 ? (v.occ ? __ffs(v.occ) - 1 : 18) : (v.occ ? 31 - __clz(v.occ) : 5);
  return v;
}

// slot mask of one die (bits 0..23 point moves, 24 bar entry, 25 farthest bear-off, 26 exact bear-off; bits 27..31 `last`)
__device__ __forceinline__ uint32_t view_mask(const View& v, const Root& r, int die) {
  if (v.off == 15u) return 0u;
  if (v.bar > 0) {
    const int e = r.player == 0 ? die - 1 : 24 - die;
    return ((r.blocked >> e) & 1u) ? 0u : (1u << 24);
  }
  uint32_t vm = r.player == 0 ? (v.occ & ~(r.blocked >> die) & ((1u << (24 - die)) - 1u))
                              : (v.occ & ~(r.blocked << die) & (0xffffffu & ~((1u << die) - 1u)));
  uint32_t last = 0;
  if (v.bear) {
    last = v.last;
    const bool far_off = r.player == 0 ? ((int)last + die >= 24) : ((int)last - die < 0);
    const uint32_t ps = r.player == 0 ? 24 - die : die - 1;
    if (far_off) vm |= 1u << 25;
    if (ps != last && ((v.occ >> ps) & 1u)) vm |= 1u << 26;
  }
  return vm | (last << 27);
}

// (source, destination) of a slot: points 0..23, BAR = 24 as a source, BEAR_OFF = 25 as a destination
__device__ __forceinline__ void slot_se(const Root& r, uint32_t slot, uint32_t last, int die, uint32_t& s, uint32_t& e) {
  if (slot < 24u) {
    s = slot;
    e = slot + r.dirsign * die;
  } else if (slot == 24u) {
    s = 24u;
    e = r.player == 0 ? die - 1 : 24 - die;
  } else {
    s = slot == 25u ? last : (r.player == 0 ? 24 - die : die - 1);
    e = 25u;
  }
}

// move one checker s -> e on the packed key (immutable_board.py:183-258)
__device__ __forceinline__ void key_move(uint32_t& k0, uint32_t& k1, uint32_t& k2, uint32_t& k3, const Root& r, uint32_t s, uint32_t e) {
  if (s == 24u) {
    k3 -= 1u << 24;
  } else {
    const uint32_t ds = 1u << ((s & 7u) * 4u);
    const uint32_t ws = s >> 3;
    k0 -= ws == 0u ? ds : 0u;
    k1 -= ws == 1u ? ds : 0u;
    k2 -= ws == 2u ? ds : 0u;
  }
  if (e == 25u) {
    k3 += 1u << 28;
  } else {
    const uint32_t de = 1u << ((e & 7u) * 4u);
    const uint32_t we = e >> 3;
    k0 += we == 0u ? de : 0u;
    k1 += we == 1u ? de : 0u;
    k2 += we == 2u ? de : 0u;
    k3 |= ((r.blot >> e) & 1u) << e;
  }
}

__device__ __forceinline__ uint32_t key_count(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t p) {
  const uint32_t w = p < 8u ? k0 : (p < 16u ? k1 : k2);
  return (w >> ((p & 7u) * 4u)) & 15u;
}

// destination of a doubles source point for die (0-based point or BAR)
__device__ __forceinline__ uint32_t dbl_dest(const Root& r, uint32_t s, int die) {
  if (s == 24u) return r.player == 0 ? die - 1 : 24 - die;
  const int e = (int)s + r.dirsign * die;
  return (e < 0 || e > 23) ? 25u : (uint32_t)e;
}

#if BG_EMIT_NIBBLE
// 4 nibbles (low 16 bits) -> 4 bytes
__device__ __forceinline__ uint32_t nib4_spread(uint32_t x) {
  const uint32_t y = (x | (x << 8)) & 0x00ff00ffu;
  return (y | (y << 4)) & 0x0f0f0f0fu;
}
#endif

template <class C>
__device__ __forceinline__ uint32_t hash_key(uint32_t k) {
  return (k * 0x9E3779B1u) >> (32 - C::TCAP_LOG2);
}

__device__ __forceinline__ bool is_double_id(uint32_t id) { return (DBL_IDS >> (id - 1u)) & 1u; }
__device__ __forceinline__ int die_of_id(uint32_t id) { return __popc(DBL_IDS & ((1u << (id - 1u)) - 1u)) + 1; }
__device__ __forceinline__ uint32_t id_of_die0(int d0) { return (uint32_t)(d0 * 6 - d0 * (d0 - 1) / 2) + 1u; }

template <class C>
__device__ __forceinline__ void clear_tab(uint32_t* tab, int lane) {
#pragma unroll 4
  for (int i = 0; i < C::TCAP / 32; ++i) tab[i * 32 + lane] = 0u;
  __syncwarp();
}

// first index of buf[0 .. n) whose id is >= id (ids are non-decreasing along a doubles frontier)
__device__ __forceinline__ int lower_bound_id(const uint32_t* buf, int n, uint32_t id) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((buf[mid] >> 25) < id) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// rebuild the hash set from the results buf[from .. n)
template <class C>
__device__ __noinline__ void rehash21(int buf_off, int from, int n) {
  uint32_t* const W = wsm<C>();
  const int lane = threadIdx.x & 31;
  clear_tab<C>(W + O_TAB, lane);
  for (int i = from + lane; i < n; i += 32) {
    const uint32_t w = W[buf_off + i];
    const uint32_t key = is_double_id(w >> 25) ? (w & ~(31u << 20)) : w;
    uint32_t idx = hash_key<C>(key);
    while (atomicCAS(&W[O_TAB + idx], 0u, key) != 0u) idx = (idx + 1) & (C::TCAP - 1);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------------------
// flush: emit the complete items buf[0 .. n_emit) -- reserve pool rows, write each item's offset / count, rebuild the boards from
// the root bytes + codes in the staging area (the hash set's memory: the caller rehashes afterwards if it still needs the set) and
// copy them out -- then move the tail buf[n_emit .. n_res) to the front.  Returns the new emitted-items mask.
// ---------------------------------------------------------------------------------------------------------------------------
template <class C>
__device__ __noinline__ uint32_t flush21(const MovegenParams* __restrict__ Pp, long long pos, int player, uint32_t blot, uint4 rk, int buf_off,
                                         int n_emit, int n_res, uint32_t emitted, bool single) {
  const MovegenParams& P = *Pp;
  uint32_t* const W = wsm<C>();
  uint32_t* const res = W + buf_off;
  const int lane = threadIdx.x & 31;
  if (n_emit <= 0) return emitted;
  // per-item ranges: the results of an item are contiguous
  if (lane < 24) {
    W[C::O_ISTART + lane] = 0u;
    W[C::O_ISTART + 24 + lane] = 0u;
  }
  __syncwarp();
  for (int i0 = 0; i0 < n_emit; i0 += 32) {
    const int i = i0 + lane;
    if (i < n_emit) {
      const uint32_t id = res[i] >> 25;
      const uint32_t pid = i > 0 ? res[i - 1] >> 25 : 0u;
      const uint32_t nid = i + 1 < n_emit ? res[i + 1] >> 25 : 0u;
      BG_CHECK(id >= 1 && id <= 21, "roll id of a result");
      if (id != pid) W[C::O_ISTART + id] = (uint32_t)i;
      if (id != nid) W[C::O_ISTART + 24 + id] = (uint32_t)(i + 1);
    }
  }
  __syncwarp();
  int rs = 0, re = 0;  // lane = roll id (1..21): its range
  if (lane >= 1 && lane <= 21) {
    rs = (int)W[C::O_ISTART + lane];
    re = (int)W[C::O_ISTART + 24 + lane];
  }
  const int true_count = re - rs;
  int n_rows = n_emit;
  // an item with more than item_cap results keeps its first item_cap rows (out_count stays the true count); rare: squeeze the rest out
  uint32_t overm = __ballot_sync(BG_FULL, true_count > P.item_cap);
  while (overm) {
    const int l = __ffs(overm) - 1;
    overm &= overm - 1u;
    const int e_l = __shfl_sync(BG_FULL, re, l), drop = e_l - __shfl_sync(BG_FULL, rs, l) - P.item_cap;
    for (int i0 = e_l; i0 < n_rows; i0 += 32) {
      const int i = i0 + lane;
      const uint32_t w = i < n_rows ? res[i] : 0u;
      __syncwarp();
      if (i < n_rows) res[i - drop] = w;
      __syncwarp();
    }
    n_rows -= drop;
    if (rs >= e_l) {
      rs -= drop;
      re -= drop;
    }
    if (lane == l) re -= drop;
  }
  long long base = 0;
  if (lane == 0) {
    base = (long long)atomicAdd(P.pool_cursor, (unsigned long long)n_rows);
    if (base + n_rows > P.pool_cap) {
      base = -1;
      atomicMin(P.status, BG_ERR_CAPACITY);
    }
  }
  base = __shfl_sync(BG_FULL, base, 0);
  if (true_count > 0) {
    const long long item = single ? pos : pos * 21 + (lane - 1);  // single-roll mode: `pos` IS the item
    P.out_count[item] = (int32_t)true_count;
    P.out_offsets[item] = base < 0 ? -1ll : base + (long long)rs;
  }
  emitted |= __ballot_sync(BG_FULL, true_count > 0) >> 1;
  if (base >= 0 && P.out_codes) {
    // compact mode: the code and the position it applies to; nothing is materialised
    const uint32_t src_pos = (uint32_t)((single && P.all_rolls) ? pos / 21 : pos);
    for (int i = lane; i < n_rows; i += 32) P.out_codes[base + i] = make_uint2(res[i], src_pos);
    if (P.out_flags)
      for (int i = lane; i < n_rows; i += 32) P.out_flags[base + i] = (uint8_t)player;
  } else if (base >= 0 && !(P.grab & 8)) {
    uint32_t* const stage = W + O_TAB;
    Root r;
    r.player = player;
    r.dirsign = player == 0 ? 1 : -1;
    r.blot = blot;
    const int ow = player * 6, pw = (1 - player) * 6;
#if BG_EMIT_NIBBLE
    uint32_t oppw[6];  // the opponent's six point words and the bar / off word of the root
#pragma unroll
    for (int q = 0; q < 6; ++q) oppw[q] = W[O_ROOT + pw + q];
    const uint32_t w12 = W[O_ROOT + 12];
    const uint32_t opp_bar = (w12 >> (8 * (1 - player))) & 0xffu, opp_off = (w12 >> (16 + 8 * (1 - player))) & 0xffu;
#else
    uint32_t rw[13];
#pragma unroll
    for (int q = 0; q < 13; ++q) rw[q] = W[O_ROOT + q];
    const int own = player * 24, opp = (1 - player) * 24;
    (void)ow;
    (void)pw;
#endif
    const unsigned long long gw_base = ((unsigned long long)(uintptr_t)P.out_boards >> 2) + (unsigned long long)base * 13ull;
    uint32_t* const gp_base = reinterpret_cast<uint32_t*>(P.out_boards) + base * 13;
    for (int i0 = 0; i0 < n_rows; i0 += 32) {
      const int i = i0 + lane;
      // rows are staged with the same 16-byte phase as their destination, so that the body of the copy is 16-byte loads and stores
      const int ph = (int)((gw_base + (unsigned long long)i0 * 13ull) & 3ull);
      __syncwarp();
      if (i < n_rows) {
#if BG_EMIT_NIBBLE
        // the afterstate's packed key = root key with the code's moves applied; then nibbles -> bytes (as movegen.cu's emit)
        const uint32_t w = res[i];
        const bool dbl = is_double_id(w >> 25);
        const int die = die_of_id(w >> 25);
        uint32_t k0 = rk.x, k1 = rk.y, k2 = rk.z, k3 = rk.w;
        if (!dbl && ((w >> 20) & 31u) != NONE5) k3 |= 1u << ((w >> 20) & 31u);
        const int nq = dbl ? 4 : 2;
#pragma unroll 1
        for (int q = 0; q < nq; ++q) {
          const uint32_t s = (w >> (5 * q)) & 31u;
          if (s == NONE5) break;
          key_move(k0, k1, k2, k3, r, s, dbl ? dbl_dest(r, s, die) : (w >> (10 + 5 * q)) & 31u);
        }
        uint32_t* const row = stage + ph + lane * 13;
        row[ow + 0] = nib4_spread(k0 & 0xffffu);
        row[ow + 1] = nib4_spread(k0 >> 16);
        row[ow + 2] = nib4_spread(k1 & 0xffffu);
        row[ow + 3] = nib4_spread(k1 >> 16);
        row[ow + 4] = nib4_spread(k2 & 0xffffu);
        row[ow + 5] = nib4_spread(k2 >> 16);
#pragma unroll
        for (int q = 0; q < 6; ++q) row[pw + q] = oppw[q] - ((((k3 >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u);
        const uint32_t ownbar = (k3 >> 24) & 15u, ownoff = k3 >> 28, ob2 = opp_bar + __popc(k3 & 0xffffffu);
        row[12] = player == 0 ? (ownbar | (ob2 << 8) | (ownoff << 16) | (opp_off << 24)) : (ob2 | (ownbar << 8) | (opp_off << 16) | (ownoff << 24));
#else
        // the root's 13 words, then one byte decrement / increment per checker moved or hit
        uint32_t* const row = stage + ph + lane * 13;
        BG_CHECK(ph + lane * 13 + 13 <= C::TCAP, "staging row");
#pragma unroll
        for (int q = 0; q < 13; ++q) row[q] = rw[q];
        uint8_t* const rb = reinterpret_cast<uint8_t*>(row);
        const uint32_t w = res[i];
        const bool dbl = is_double_id(w >> 25);
        const int die = die_of_id(w >> 25);
        uint32_t hit = (!dbl && ((w >> 20) & 31u) != NONE5) ? 1u << ((w >> 20) & 31u) : 0u;
        const int nq = dbl ? 4 : 2;
#pragma unroll 1
        for (int q = 0; q < nq; ++q) {
          const uint32_t s = (w >> (5 * q)) & 31u;
          if (s == NONE5) break;
          const uint32_t e = dbl ? dbl_dest(r, s, die) : (w >> (10 + 5 * q)) & 31u;
          rb[s == 24u ? 48 + player : own + (int)s] -= 1;
          rb[e == 25u ? 50 + player : own + (int)e] += 1;
          hit |= blot & (1u << e);  // e == 25: bit 25 of a 24-bit mask
        }
        rb[48 + 1 - player] += (uint8_t)__popc(hit);
        while (hit) {
          const int p = __ffs(hit) - 1;
          hit &= hit - 1u;
          rb[opp + p] -= 1;
        }
#endif
        if (P.out_flags) P.out_flags[base + i] = (uint8_t)player;
        if (P.out_owner) P.out_owner[base + i] = (int32_t)(single ? pos : pos * 21 + (w >> 25) - 1);
      }
      __syncwarp();
      const int nw = (n_rows - i0 < 32 ? n_rows - i0 : 32) * 13;
      const uint32_t* const sp = stage + ph;
      uint32_t* const gp = gp_base + i0 * 13;
      int head = (4 - ph) & 3;
      head = head < nw ? head : nw;
      if (lane < head) gp[lane] = sp[lane];
      const int nb = (nw - head) >> 2;
      const uint4* const s4 = reinterpret_cast<const uint4*>(sp + head);
      uint4* const g4 = reinterpret_cast<uint4*>(gp + head);
#if BG_COPY_UNROLL
#pragma unroll
      for (int it = 0; it < 4; ++it) {  // nb <= 104
        const int v = lane + 32 * it;
        if (v < nb) g4[v] = s4[v];
      }
#else
      for (int v = lane; v < nb; v += 32) g4[v] = s4[v];
#endif
      const int t0 = head + 4 * nb;
      if (lane < nw - t0) gp[t0 + lane] = sp[t0 + lane];
    }
    __syncwarp();
  }
  // keep the tail
  const int keep = n_res - n_emit;
  for (int i0 = 0; i0 < keep; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t w = i < keep ? res[n_emit + i] : 0u;
    __syncwarp();
    if (i < keep) res[i] = w;
  }
  __syncwarp();
  return emitted;
}

// make room in the result buffer: every result before the first one with roll id `id0` (the item the pending candidates belong
// to) is a complete item -- emit those, keep the rest, rebuild the hash set from it.  Returns emitted | n_res << 32, or bit 63 set
// if the pending item alone fills the buffer.
template <class C>
__device__ __noinline__ unsigned long long make_room21(const MovegenParams* __restrict__ Pp, long long pos, int player, uint32_t blot, uint4 rk,
                                                       int buf_off, int n_res, uint32_t id0, uint32_t emitted, bool single) {
  uint32_t* const W = wsm<C>();
  const int lane = threadIdx.x & 31;
  int keep_from = n_res;
  for (int i0 = 0; i0 < n_res; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t b = __ballot_sync(BG_FULL, i < n_res && (W[buf_off + i] >> 25) == id0);
    if (b) {
      keep_from = i0 + __ffs(b) - 1;
      break;
    }
  }
  if (keep_from == 0) return 1ull << 63;
  emitted = flush21<C>(Pp, pos, player, blot, rk, buf_off, keep_from, n_res, emitted, single);
  n_res -= keep_from;
  rehash21<C>(buf_off, 0, n_res);
  return (unsigned long long)emitted | ((unsigned long long)n_res << 32);
}

// ---------------------------------------------------------------------------------------------------------------------------
// One batch of <= 32 parents (one per lane: `pinfo` describes the parent, the low 27 bits of `m` are its candidate slots) -> its
// candidates in canonical order (parent order, then slot order), first-occurrence dedup through the hash set, survivors appended
// to dst[n ..).  mode 0: non-doubles (pinfo = src1 | dst1 << 5 | die0 of the second sub-move << 10 | single << 13 | last << 14 |
// id << 25); mode 1: doubles (pinfo = the parent's sorted sources (20 bits) | last << 20 | id << 25; the child records its slot).
// Sub-batches of <= DCAP candidates; returns n | (resume + 1) << 16 when the next sub-batch would not fit below `cap` or would
// push the hash set past C::TLIVE keys (dst[key_base .. n) are the keys it holds); the caller makes room and calls again with
// start = resume.  High half 0: the batch is done.
// ---------------------------------------------------------------------------------------------------------------------------
template <class C>
__device__ __noinline__ uint32_t expand21(int mode, uint32_t pinfo, uint32_t m, int start, int dst_off, int n, int cap, int key_base, int player,
                                          uint32_t blot) {
  uint32_t* const W = wsm<C>();
  uint16_t* const desc = reinterpret_cast<uint16_t*>(W + O_DESC);
  uint32_t* const tab = W + O_TAB;
  uint32_t* const dst = W + dst_off;
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  Root r;
  r.player = player;
  r.dirsign = player == 0 ? 1 : -1;
  const uint32_t mm0 = m & 0x7ffffffu;
  const int cnt = __popc(mm0);
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(BG_FULL, inc, o);
    if (lane >= o) inc += t;
  }
  const int total = __shfl_sync(BG_FULL, inc, 31);
  const int exc = inc - cnt;
  while (start < total) {
    const int limit = start + DCAP;
    const uint32_t nofit = __ballot_sync(BG_FULL, inc > limit);
    const int next_start = nofit ? __shfl_sync(BG_FULL, exc, __ffs(nofit) - 1) : total;
    const int ncand = next_start - start;
    if (n + ncand > cap || n - key_base + ncand > C::TLIVE) return (uint32_t)n | ((uint32_t)(start + 1) << 16);
    if (exc >= start && inc <= limit) {
      uint32_t mm = mm0;
      int o = exc - start;
      while (mm) {
        const int slot = __ffs(mm) - 1;
        mm &= mm - 1u;
        BG_CHECK(o >= 0 && o < DCAP, "descriptor index");
        desc[o++] = (uint16_t)(lane | (slot << 5));
      }
    }
    __syncwarp();
    for (int t0 = 0; t0 < ncand; t0 += 32) {
      const int t = t0 + lane;
      const bool cv = t < ncand;
      const uint32_t dsc = cv ? desc[t] : 0u;
      const uint32_t pi = __shfl_sync(BG_FULL, pinfo, dsc & 31u);
      const uint32_t slot = dsc >> 5;
      uint32_t word, key;
      if (mode == 0) {
        const uint32_t s1 = pi & 31u, e1 = (pi >> 5) & 31u;
        uint32_t code;
        if ((pi >> 13) & 1u) {  // single sub-move result (handle_move_types.py:70-81)
          code = s1 | (NONE5 << 5) | (e1 << 10) | (NONE5 << 15) | (NONE5 << 20);
        } else {
          uint32_t s2, e2;
          slot_se(r, slot, (pi >> 14) & 31u, (int)((pi >> 10) & 7u) + 1, s2, e2);
          if (s2 == e1) {  // a checker lands on e1 and a checker leaves it
            const uint32_t in = ((blot >> e1) & 1u) ? e1 : NONE5;
            code = s1 | (NONE5 << 5) | (e2 << 10) | (NONE5 << 15) | (in << 20);
          } else if (s1 == e2) {  // the second sub-move lands on the point the first one left
            code = s2 | (NONE5 << 5) | (e1 << 10) | (NONE5 << 15) | (NONE5 << 20);
          } else {
            code = min(s1, s2) | (max(s1, s2) << 5) | (min(e1, e2) << 10) | (max(e1, e2) << 15) | (NONE5 << 20);
          }
        }
        word = code | (pi & (31u << 25));
        key = word;
      } else {
        uint32_t s, e;
        slot_se(r, slot, (pi >> 20) & 31u, die_of_id(cv ? pi >> 25 : 1u), s, e);
        // sorted insert of s into the parent's (at most three) sources
        const uint32_t a0 = pi & 31u, a1 = (pi >> 5) & 31u, a2 = (pi >> 10) & 31u;
        const uint32_t l0 = min(a0, s), t1 = max(a0, s);
        const uint32_t l1 = min(a1, t1), t2 = max(a1, t1);
        const uint32_t l2 = min(a2, t2), l3 = max(a2, t2);
        key = l0 | (l1 << 5) | (l2 << 10) | (l3 << 15) | (pi & (31u << 25));
        word = key | (slot << 20);
      }
      // first occurrence wins: equal keys inside the round elect their lowest lane, earlier rounds are in the hash set
      const uint32_t grp = __match_any_sync(BG_FULL, cv ? key : (0x80000000u | (uint32_t)lane));
      bool isnew = false;
      if (cv && (__ffs(grp) - 1) == lane) {
        uint32_t idx = hash_key<C>(key);
        int probes = 0;
        (void)probes;
        while (true) {
          const uint32_t old = atomicCAS(&tab[idx], 0u, key);
          if (old == 0u) {
            isnew = true;
            break;
          }
          if (old == key) break;
          idx = (idx + 1) & (C::TCAP - 1);
          BG_CHECK(++probes <= C::TCAP, "hash set full");
        }
      }
      const uint32_t bal = __ballot_sync(BG_FULL, isnew);
      BG_CHECK(n + __popc(bal) <= cap && dst_off + n + __popc(bal) <= C::O_BUF + C::BCAP, "append beyond the arena");
      if (isnew) dst[n + __popc(bal & lt)] = word;
      n += __popc(bal);
    }
    __syncwarp();
    start = next_start;
  }
  return (uint32_t)n;
}

// roll id of the parent that candidate `st` of a batch belongs to (parents are in item order)
__device__ __noinline__ uint32_t pending_id(uint32_t pinfo, uint32_t m, int st, int lane) {
  int inl = __popc(m & 0x7ffffffu);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(BG_FULL, inl, o);
    if (lane >= o) inl += t;
  }
  const uint32_t pend = __ballot_sync(BG_FULL, inl > st);
  return __shfl_sync(BG_FULL, pinfo >> 25, pend ? __ffs(pend) - 1 : 0);
}

__device__ __forceinline__ void push_overflow(const MovegenParams& P, long long item) {
  const int q = atomicAdd(P.ovf_count, 1);
  P.ovf_list[q] = (int32_t)item;
}

// move set of a doubles node (its sorted sources in w) for its die, pruned: a slot below the slot `t` that created the node is dropped
// when its move was already legal one ply earlier, i.e. unless it moves the checker that has just landed alone on t's destination
__device__ __noinline__ uint32_t node_mask(uint32_t w, const Root r, uint32_t rk0, uint32_t rk1, uint32_t rk2, uint32_t rk3) {
  const int die = die_of_id(w >> 25);
  uint32_t k0 = rk0, k1 = rk1, k2 = rk2, k3 = rk3;
#pragma unroll 1
  for (int q = 0; q < 3; ++q) {
    const uint32_t s = (w >> (5 * q)) & 31u;
    if (s == NONE5) break;
    key_move(k0, k1, k2, k3, r, s, dbl_dest(r, s, die));
  }
  const View v = make_view(k0, k1, k2, k3, r);
  uint32_t m = view_mask(v, r, die);
  const uint32_t t = (w >> 20) & 31u;
  if (t < 24u) {
    uint32_t keep = ~((1u << t) - 1u);
    const int lt = (int)t + r.dirsign * die;
    if (lt >= 0 && lt < (int)t && key_count(k0, k1, k2, (uint32_t)lt) == 1u) keep |= 1u << lt;
    m &= keep | 0xff000000u;
  }
  return m;
}

template <bool ALL, class C>
__global__ void __launch_bounds__(C::WARPS * 32, C::CTAS) k_movegen21(const __grid_constant__ MovegenParams P) {
  uint32_t* const W = wsm<C>();
  const int lane = threadIdx.x & 31;
  const uint32_t* const boards32 = reinterpret_cast<const uint32_t*>(P.boards);
  const long long n_work = (!ALL && P.in_list) ? (long long)(*P.in_count) : P.B;
  while (true) {
    long long pos = 0;
    if (lane == 0) pos = atomicAdd(P.item_counter, 1);
    pos = __shfl_sync(BG_FULL, pos, 0);
    if (pos >= n_work) break;
    if constexpr (!ALL) {
      if (P.in_list) pos = P.in_list[pos];  // single-roll mode as a tail tier: the items another tier handed over
    }
    // single-roll mode: `pos` is the ITEM; in a position-major batch its board is shared by 21 items
    const long long src = (!ALL && P.all_rolls) ? pos / 21 : pos;
    __syncwarp();
    if (P.active && !P.active[src]) {
      if (lane < (ALL ? 21 : 1)) {
        P.out_count[ALL ? pos * 21 + lane : pos] = 0;
        P.out_offsets[ALL ? pos * 21 + lane : pos] = 0;
      }
      continue;
    }
    if (lane < 13) W[O_ROOT + lane] = boards32[src * 13 + lane];
    __syncwarp();
    const int player = P.players[src] & 1;
    // ---- root (as movegen.cu generate()) ------------------------------------------------------------------------------
    const int ob = player * 6, pb = (1 - player) * 6;
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < 13; ++i) bad |= W[O_ROOT + i] & 0xf0f0f0f0u;
    // single-roll mode: the item's own roll; its roll index selects the one active lane of the non-doubles section / the one die
    int my_rid = -1, need = 0x3f;
    bool dbl_item = true;
    if constexpr (!ALL) {
      int d0, d1;
      if (P.all_rolls) {  // roll index -> (d0 <= d1), lexicographic
        int ri = (int)(pos - src * 21);
        d0 = 1;
        while (ri >= 7 - d0) {
          ri -= 7 - d0;
          ++d0;
        }
        d1 = d0 + ri;
      } else {
        d0 = P.rolls[2 * pos];
        d1 = P.rolls[2 * pos + 1];
      }
      if (d0 < 1 || d0 > 6 || d1 < 1 || d1 > 6) bad = 1u;
      const int lo0 = (d0 < d1 ? d0 : d1) - 1, hi0 = (d0 < d1 ? d1 : d0) - 1;
      my_rid = lo0 * 6 - lo0 * (lo0 - 1) / 2 + (hi0 - lo0);  // index in the (1,1), (1,2), ... (6,6) order
      need = (1 << lo0) | (1 << hi0);
      dbl_item = lo0 == hi0;
    }
    if (bad) {
      if (lane < (ALL ? 21 : 1)) {
        P.out_count[ALL ? pos * 21 + lane : pos] = 0;
        P.out_offsets[ALL ? pos * 21 + lane : pos] = -1;
      }
      if (lane == 0) atomicMin(P.status, BG_ERR_INVARIANT);
      continue;
    }
    const uint32_t w12 = W[O_ROOT + 12];
    const uint32_t rk0 = bytes_to_nib4(W[O_ROOT + ob + 0]) | (bytes_to_nib4(W[O_ROOT + ob + 1]) << 16);
    const uint32_t rk1 = bytes_to_nib4(W[O_ROOT + ob + 2]) | (bytes_to_nib4(W[O_ROOT + ob + 3]) << 16);
    const uint32_t rk2 = bytes_to_nib4(W[O_ROOT + ob + 4]) | (bytes_to_nib4(W[O_ROOT + ob + 5]) << 16);
    const uint32_t own_bar = (w12 >> (8 * player)) & 15u, own_off = (w12 >> (16 + 8 * player)) & 15u;
    const uint32_t rk3 = (own_bar << 24) | (own_off << 28);
    const uint4 rk = make_uint4(rk0, rk1, rk2, rk3);
    Root r;
    r.player = player;
    r.dirsign = player == 0 ? 1 : -1;
    r.home = player == 0 ? 0xfc0000u : 0x00003fu;
    {
      uint32_t oc = 0;
      if (lane < 24) oc = (W[O_ROOT + pb + (lane >> 2)] >> ((lane & 3) * 8)) & 0xffu;
      r.blocked = __ballot_sync(BG_FULL, oc >= 2) & 0xffffffu;
      r.blot = __ballot_sync(BG_FULL, oc == 1) & 0xffffffu;
      uint32_t total = own_bar + own_off;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const uint32_t w = W[O_ROOT + ob + i];
        total += (w & 0xff) + ((w >> 8) & 0xff) + ((w >> 16) & 0xff) + (w >> 24);
      }
      r.valid15 = total == 15u;
    }
    // ---- ply 1: the six dice (lane = die - 1) ------------------------------------------------------------------------------
    const View rv = make_view(rk0, rk1, rk2, rk3, r);
    const uint32_t m1 = (lane < 6 && ((need >> lane) & 1)) ? view_mask(rv, r, lane + 1) : 0u;
    const int n1 = __popc(m1 & 0x7ffffffu);
    int inc1 = n1;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const int t = __shfl_up_sync(BG_FULL, inc1, o);
      if (lane >= o) inc1 += t;
    }
    const int exc1 = inc1 - n1;
    const int N1 = __shfl_sync(BG_FULL, inc1, 5);
    if (N1 > C1CAP) {  // too many first moves for the table: every roll of this position goes to the generic tiers
      if (lane < (ALL ? 21 : 1)) push_overflow(P, ALL ? pos * 21 + lane : pos);
      continue;
    }
    // ---- children of the root and their six second-die move sets (lane = child) ------------------------------------------------
    uint32_t has2a = 0, has2b = 0;  // bit 6a+b (a < 5) / bit b (a == 5): some child of die a has a move with die b
    for (int c0 = 0; c0 < N1; c0 += 32) {
      const int c = c0 + lane;
      const bool valid = c < N1;
      int d = 0;
#pragma unroll
      for (int q = 0; q < 5; ++q) d += __shfl_sync(BG_FULL, inc1, q) <= c ? 1 : 0;
      const uint32_t md = __shfl_sync(BG_FULL, m1, d);
      const int bd = __shfl_sync(BG_FULL, exc1, d);
      uint32_t nz = 0;
      if (valid) {
        const uint32_t slot = nth_set_bit(md & 0x7ffffffu, c - bd);
        uint32_t s, e;
        slot_se(r, slot, md >> 27, d + 1, s, e);
        uint32_t k0 = rk0, k1 = rk1, k2 = rk2, k3 = rk3;
        key_move(k0, k1, k2, k3, r, s, e);
        const View v = make_view(k0, k1, k2, k3, r);
#pragma unroll 1
        for (int b = 0; b < 6; ++b) {
          const uint32_t m = view_mask(v, r, b + 1);
          W[O_C1MASK + b * C1CAP + c] = m;
          nz |= ((m & 0x7ffffffu) != 0u ? 1u : 0u) << b;
        }
        const uint32_t lone = (e < 24u && key_count(k0, k1, k2, e) == 1u) ? 1u : 0u;
        BG_CHECK(c < C1CAP && s <= 24u && e <= 25u, "children table");
        W[O_C1INFO + c] = slot | (s << 5) | (e << 10) | ((uint32_t)d << 15) | (lone << 18);
      }
      has2a |= __reduce_or_sync(BG_FULL, (valid && d < 5) ? nz << (6 * d) : 0u);
      has2b |= __reduce_or_sync(BG_FULL, (valid && d == 5) ? nz : 0u);
    }
    __syncwarp();

    uint32_t emitted = 0;  // bit i: item i (roll index) has been written (or handed to the generic tiers)
    int n_res = 0;
    // ---- non-doubles (generate_all_moves.py:25-53, handle_move_types.py:7-81) ------------------------------------------------
    // lane < 15 = roll; which die orders contribute which kind of entries:
    //   forward order has two-move plays: they are results; the reverse order's two-move plays too (its singles would be filtered);
    //   otherwise the forward order's singles are results, unless exactly one (quirk Q1: reverse order skipped), none (reverse
    //   order alone), or the reverse order has two-move plays (the singles are filtered by the max-length rule).
    if (!(P.grab & 2)) {
      int cnt0 = 0, cnt1 = 0;
      uint32_t sg0 = 0, sg1 = 0;
      const int lo = (int)((ND_LO >> (3 * (lane < 15 ? lane : 0))) & 7u), hi = (int)((ND_HI >> (3 * (lane < 15 ? lane : 0))) & 7u);
      const int nh = __shfl_sync(BG_FULL, n1, hi), nl = __shfl_sync(BG_FULL, n1, lo);
      if (lane < 15 && (ALL || lane + 1 + lo == my_rid)) {  // roll index of non-double lane = lane + 1 + lo
        const bool H0 = hi < 5 ? (has2a >> (6 * hi + lo)) & 1u : (has2b >> lo) & 1u;
        const bool H1 = (has2a >> (6 * lo + hi)) & 1u;  // lo < 5 always
        if (H0) {
          cnt0 = nh;
          cnt1 = H1 ? nl : 0;
        } else if (nh == 1) {
          cnt0 = 1;
          sg0 = 1;
        } else if (nh == 0) {
          cnt1 = nl;
          sg1 = H1 ? 0u : 1u;
        } else if (H1) {
          cnt1 = nl;
        } else {
          cnt0 = nh;
          cnt1 = nl;
          sg0 = sg1 = 1;
        }
      }
      const int pe = cnt0 + cnt1;
      int pinc = pe;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int t = __shfl_up_sync(BG_FULL, pinc, o);
        if (lane >= o) pinc += t;
      }
      const int PE = __shfl_sync(BG_FULL, pinc, 14);
      if (lane >= 15) pinc = 0x7fffffff;
      const int pexc = pinc - pe;
      const uint32_t rinfo = (uint32_t)cnt0 | (sg0 << 6) | (sg1 << 7);
      if (PE > 0) clear_tab<C>(W + O_TAB, lane);
      bool nd_abort = false;
      for (int e0 = 0; e0 < PE && !nd_abort; e0 += 32) {
        const int e = e0 + lane;
        const bool valid = e < PE;
        int q = 0;  // roll of entry e: number of rolls whose inclusive end <= e
#pragma unroll
        for (int sft = 8; sft > 0; sft >>= 1) {
          const int v = __shfl_sync(BG_FULL, pinc, q + sft - 1);
          if (v <= e) q += sft;
        }
        q = q > 14 ? 14 : q;
        const int j = e - __shfl_sync(BG_FULL, pexc, q);
        const uint32_t ri = __shfl_sync(BG_FULL, rinfo, q);
        const int c0n = (int)(ri & 63u);
        const int order = j >= c0n ? 1 : 0;
        const int qlo = (int)((ND_LO >> (3 * q)) & 7u), qhi = (int)((ND_HI >> (3 * q)) & 7u);
        const int fa = order ? qlo : qhi, fb = order ? qhi : qlo;
        const int c = __shfl_sync(BG_FULL, exc1, fa) + (order ? j - c0n : j);
        const uint32_t single = (ri >> (6 + order)) & 1u;
        const uint32_t mhi = __shfl_sync(BG_FULL, m1, qhi);
        uint32_t pinfo = 0, m = 0;
        if (valid) {
          const uint32_t info = W[O_C1INFO + c];
          m = single ? 1u : W[O_C1MASK + fb * C1CAP + c];
          // reverse order: a second sub-move from a point that could move first with the high die commutes with an earlier pair
          if (!single && order == 1 && own_bar == 0u && (info & 31u) < 24u) m &= ~(mhi & 0xffffffu);
          pinfo = ((info >> 5) & 1023u) | ((uint32_t)fb << 10) | (single << 13) | ((m >> 27) << 14) | ((uint32_t)(q + 1 + qlo + 1) << 25);
        }
        int st = 0;
        while (true) {
          const uint32_t rc = expand21<C>(0, pinfo, m, st, C::O_BUF, n_res, C::BCAP, 0, player, r.blot);
          n_res = (int)(rc & 0xffffu);
          if ((rc >> 16) == 0u) break;
          st = (int)(rc >> 16) - 1;
          const unsigned long long mr = make_room21<C>(&P, pos, player, r.blot, rk, C::O_BUF, n_res, pending_id(pinfo, m, st, lane), emitted, !ALL);
          if (mr >> 63) {  // one roll alone exceeds the buffer (not observed: a non-double has <= ~120 results)
            nd_abort = true;
            break;
          }
          emitted = (uint32_t)mr;
          n_res = (int)(mr >> 32);
        }
      }
      if (nd_abort) {  // hand every non-double that has not been written to the generic tiers
        n_res = 0;
        if constexpr (ALL) {
          if (lane < 21 && !((DBL_IDS >> lane) & 1u) && !((emitted >> lane) & 1u)) push_overflow(P, pos * 21 + lane);
        } else {
          if (lane == 0) push_overflow(P, pos);
        }
        emitted |= ~DBL_IDS & 0x1fffffu;
      }
    }
    emitted = flush21<C>(&P, pos, player, r.blot, rk, C::O_BUF, n_res, n_res, emitted, !ALL);
    n_res = 0;

    // ---- doubles (handle_move_types.py:84-193): breadth-first by ply, the six trees together ------------------------------------
    if (N1 > 0 && !(P.grab & 4) && dbl_item) {
      // ply 1 -> 2 for every die: parents are the children table (their move sets are in it), nodes go to the bottom of the arena
      int n2 = 0;
      bool fail = false;
      clear_tab<C>(W + O_TAB, lane);
      for (int c0 = 0; c0 < N1 && !fail; c0 += 32) {
        const int c = c0 + lane;
        uint32_t pinfo = 0, m = 0;
        if (c < N1) {
          const uint32_t info = W[O_C1INFO + c];
          const uint32_t d0 = (info >> 15) & 7u;
          m = W[O_C1MASK + d0 * C1CAP + c];
          const uint32_t t = info & 31u, e = (info >> 10) & 31u;
          if (t < 24u) {
            uint32_t keep = ~((1u << t) - 1u);
            if (e < t && ((info >> 18) & 1u)) keep |= 1u << e;
            m &= keep | 0xff000000u;
          }
          pinfo = ((info >> 5) & 31u) | (NONE5 << 5) | (NONE5 << 10) | (NONE5 << 15) | ((m >> 27) << 20) | (id_of_die0((int)d0) << 25);
        }
        const uint32_t rc = expand21<C>(1, pinfo, m, 0, C::O_BUF, n2, C::F2CAP, 0, player, r.blot);
        n2 = (int)(rc & 0xffffu);
        fail = (rc >> 16) != 0u;
      }
      if (fail) {  // the second plies alone overflow the arena: all doubles to the generic tiers
        if (lane < 6 && n1 > 0) push_overflow(P, ALL ? pos * 21 + (id_of_die0(lane) - 1) : pos);
        emitted |= DBL_IDS;
      } else {
        // per-die ranges of the ply-2 frontier (lane = die - 1; lane 6 = end)
        const int st2 = lower_bound_id(W + C::O_BUF, n2, lane < 6 ? id_of_die0(lane) : 31u);
        const int cnt2 = __shfl_down_sync(BG_FULL, st2, 1) - st2;
        const int fb_off = C::O_BUF + n2;
        const int avail = C::BCAP - n2;
        int d = 0;
        bool singles = false;
        while (d < 6) {
          // group [d, g): as many consecutive dice as are expected to fit (ply 3 ~ 2x, ply 4 ~ 3.5x the ply-2 count)
          int g = d, sum = 0;
          while (g < 6) {
            const int cg = __shfl_sync(BG_FULL, cnt2, g);
            if (g > d && (singles || 11 * (sum + cg) > 2 * avail - 256)) break;
            sum += cg;
            ++g;
          }
          const int a0 = __shfl_sync(BG_FULL, st2, d), a1 = __shfl_sync(BG_FULL, st2, g);
          const uint32_t gmask = ((1u << g) - 1u) & ~((1u << d) - 1u);  // dice of the group (bit = die - 1)
          // dice of the group that still need results: have a first move, not written yet
          const uint32_t todo = __ballot_sync(BG_FULL, lane < 6 && ((gmask >> lane) & 1u) && n1 > 0 && !((emitted >> (id_of_die0(lane < 6 ? lane : 0) - 1)) & 1u));
          if (!todo) {
            d = g;
            continue;
          }
          bool ok = true;
          int n3 = 0, cnt3 = 0, st3 = 0;
          int res_off = fb_off, res_cap = avail;
          n_res = 0;
          // ply 2 -> 3 (into the frontier behind the ply-2 nodes), then ply 3 -> 4 (straight into the result buffer behind that)
          for (int ply = 2; ply <= 3 && ok; ++ply) {
            const bool last = ply == 3;
            const int src_off = last ? fb_off : C::O_BUF + a0, n_src = last ? n3 : a1 - a0;
            const int dst_off = last ? res_off : fb_off, cap = last ? res_cap : avail - C::RES_RESERVE;
            int n_dst = last ? n_res : 0;
            if (n_src > 0) {
              clear_tab<C>(W + O_TAB, lane);
              int key_base = n_dst;
              for (int p0 = 0; p0 < n_src && ok; p0 += 32) {
                uint32_t pinfo = 0, m = 0;
                if (p0 + lane < n_src) {
                  const uint32_t w = W[src_off + p0 + lane];
                  if ((todo >> (die_of_id(w >> 25) - 1)) & 1u) {
                    m = node_mask(w, r, rk0, rk1, rk2, rk3);
                    pinfo = (w & 0xfffffu) | ((m >> 27) << 20) | (w & (31u << 25));
                  }
                }
                int st = 0;
                while (true) {
                  const uint32_t rc = expand21<C>(1, pinfo, m, st, dst_off, n_dst, cap, key_base, player, r.blot);
                  n_dst = (int)(rc & 0xffffu);
                  if ((rc >> 16) == 0u) break;
                  st = (int)(rc >> 16) - 1;
                  // a frontier that does not fit fails the group; results make room by emitting the dice that are complete
                  const unsigned long long mr = last ? make_room21<C>(&P, pos, player, r.blot, rk, res_off, n_dst, pending_id(pinfo, m, st, lane), emitted, !ALL)
                                                     : 1ull << 63;
                  if (mr >> 63) {
                    ok = false;
                    break;
                  }
                  emitted = (uint32_t)mr;
                  n_dst = (int)(mr >> 32);
                  key_base = 0;
                }
              }
            }
            if (!ok) break;
            // dice whose tree ended one ply earlier: that ply's nodes are their results (after ply 2 -> 3: the ones that stop at ply 1 or 2;
            // after ply 3 -> 4: the ones that stop at ply 3, i.e. have ply-3 nodes but neither results in the buffer nor written ones)
            uint32_t present = emitted;
            if (!last) {
              n3 = n_dst;
              res_off = fb_off + n3;
              res_cap = avail - n3;
              st3 = lower_bound_id(W + fb_off, n3, lane < 6 ? id_of_die0(lane) : 31u);
              cnt3 = __shfl_down_sync(BG_FULL, st3, 1) - st3;
            } else {
              n_res = n_dst;
              for (int i0 = 0; i0 < n_res; i0 += 32) {
                const int i = i0 + lane;
                present |= __reduce_or_sync(BG_FULL, i < n_res ? 1u << ((W[res_off + i] >> 25) - 1u) : 0u);
              }
            }
            for (int e = d; e < g && ok; ++e) {
              if (!((todo >> e) & 1u)) continue;
              const int c2 = __shfl_sync(BG_FULL, cnt2, e), c3 = __shfl_sync(BG_FULL, cnt3, e);
              int n_add, src;
              uint32_t child_id = 0;
              if (!last) {
                if (c2 > 0 && c3 > 0) continue;
                n_add = c2 == 0 ? __shfl_sync(BG_FULL, n1, e) : c2;
                src = c2 == 0 ? O_C1INFO + __shfl_sync(BG_FULL, exc1, e) : C::O_BUF + __shfl_sync(BG_FULL, st2, e);
                child_id = c2 == 0 ? id_of_die0(e) : 0u;
              } else {
                if (c3 == 0 || ((present >> (id_of_die0(e) - 1)) & 1u)) continue;
                n_add = c3;
                src = fb_off + __shfl_sync(BG_FULL, st3, e);
              }
              if (n_res + n_add > res_cap) {  // everything buffered is complete: emit it
                emitted = flush21<C>(&P, pos, player, r.blot, rk, res_off, n_res, n_res, emitted, !ALL);
                n_res = 0;
              }
              if (n_add > res_cap) {
                ok = false;
                break;
              }
              BG_CHECK(res_off + n_res + n_add <= C::O_BUF + C::BCAP, "results beyond the arena");
              for (int i = lane; i < n_add; i += 32) {
                const uint32_t w = W[src + i];
                W[res_off + n_res + i] = child_id ? (((w >> 5) & 31u) | (NONE5 << 5) | (NONE5 << 10) | (NONE5 << 15) | (child_id << 25)) : w;
              }
              n_res += n_add;
              __syncwarp();
            }
          }
          if (ok) {
            emitted = flush21<C>(&P, pos, player, r.blot, rk, res_off, n_res, n_res, emitted, !ALL);
            n_res = 0;
            d = g;
          } else {
            n_res = 0;  // unwritten results of this attempt are dropped; dice already written stay written
            if (g - d > 1) {
              singles = true;  // retry this group one die at a time
            } else {
              if (lane == 0 && !((emitted >> (id_of_die0(d) - 1)) & 1u)) push_overflow(P, ALL ? pos * 21 + (id_of_die0(d) - 1) : pos);
              emitted |= 1u << (id_of_die0(d) - 1);
              d = g;
            }
          }
        }
      }
    }
    if constexpr (ALL) {
      if (lane < 21 && !((emitted >> lane) & 1u)) {
        P.out_count[pos * 21 + lane] = 0;
        P.out_offsets[pos * 21 + lane] = 0;
      }
    } else {
      if (lane == 0 && !((emitted >> my_rid) & 1u)) {
        P.out_count[pos] = 0;
        P.out_offsets[pos] = 0;
      }
    }
  }
}

template <bool ALL, class C>
int32_t launch21(const MovegenParams& P, cudaStream_t stream) {
  static DeviceOnce once;
  int32_t rc0 = once.run([]() -> int32_t { return check_cuda(opt_in_shared(k_movegen21<ALL, C>, C::SMEM_BYTES), "cudaFuncSetAttribute(k_movegen21)"); });
  if (rc0 != BG_OK) return rc0;
  int64_t want = P.in_list ? (int64_t)148 * C::CTAS : (P.B + C::WARPS - 1) / C::WARPS;
  int ctas_per_sm = C::CTAS;
  if (const char* lim = getenv("BG_MG21_CTAS")) {  // development: fewer resident CTAs per SM (co-residency experiments)
    const int v = atoi(lim);
    if (v >= 1 && v < C::CTAS) ctas_per_sm = v;
  }
  const int64_t full = (int64_t)148 * ctas_per_sm;
  const int grid = (int)(want < full ? want : full);
  k_movegen21<ALL, C><<<grid, C::WARPS * 32, C::SMEM_BYTES, stream>>>(P);
  return check_cuda(cudaGetLastError(), "k_movegen21 launch");
}

}  // namespace

// big == 0: the bulk tier (Std21); big != 0: the tail tier for items too wide for it (Big21; always consumes P.in_list)
int32_t movegen21_launch_kernel(const MovegenParams& P, cudaStream_t stream, int big) {
  if (big) return launch21<false, Big21>(P, stream);  // one warp per ITEM (of a per-item or of a position-major batch)
  return P.all_rolls ? launch21<true, Std21>(P, stream) : launch21<false, Std21>(P, stream);
}

}  // namespace bg
