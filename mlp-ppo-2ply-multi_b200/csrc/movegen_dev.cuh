// movegen_dev.cuh -- device helpers shared by the two move generators (movegen.cu: one warp per (board, roll) item;
// movegen21.cu: one warp per position, all 21 rolls).  The packed node key and the one-die move set of the reference
// (src/backgammon/moves/get_moves_one_die.py:13-251, conditions.py:5-194, board/immutable_board.py:183-258).
#pragma once
#include "bg_common.cuh"

namespace bg {

struct Root {
  int player;
  int dirsign;       // +1 / -1
  uint32_t blocked;  // opponent >= 2 (24 bits)
  uint32_t blot;     // opponent == 1 (24 bits)
  uint32_t home;     // mover's home mask
  bool valid15;      // mover has exactly 15 checkers (conditions.py:191-194)
};

struct Node {
  uint32_t k0, k1, k2, k3, m0, m1;
};

// 24 nibbles (3 words) -> 24-bit occupancy mask (bit p set iff nibble p != 0)
__device__ __forceinline__ uint32_t nib_occupancy(uint32_t w) {  // 8 nibbles -> 8 bits
  uint32_t t = w | (w >> 1);
  t |= t >> 2;
  t &= 0x11111111u;                    // bit 4i = nibble i non-zero
  t = (t | (t >> 3)) & 0x03030303u;    // byte b: bits 0,1 = nibbles 2b, 2b+1
  t = (t | (t >> 6)) & 0x000f000fu;    // half h: bits 0..3 = nibbles 4h..4h+3
  return (t | (t >> 12)) & 0xffu;
}

// One-die move set of node `p` as a slot mask, in the reference's get_moves_with_one_die order (slot order):
//   bits 0..23 in-board move from that point, 24 bar entry, 25 bear-off of the farthest checker, 26 exact bear-off;
//   bits 27..31 carry `last` (the farthest checker's point) for slot 25.
__device__ __forceinline__ uint32_t move_mask(const Node& p, const Root& r, int die) {
  const uint32_t occ = nib_occupancy(p.k0) | (nib_occupancy(p.k1) << 8) | (nib_occupancy(p.k2) << 16);
  const uint32_t bar = (p.k3 >> 24) & 15u, off = p.k3 >> 28;
  if (off == 15u) return 0u;  // GAME_OVER (conditions.py:16-17)
  if (bar > 0) {              // ON_BAR (get_moves_one_die.py:86-130)
    const int e = r.player == 0 ? die - 1 : 24 - die;
    return ((r.blocked >> e) & 1u) ? 0u : (1u << 24);
  }
  // NORMAL (:40-83) / in-home moves of BEAR_OFF (:164-189): destination on the board and not blocked
  uint32_t vm = r.player == 0 ? (occ & ~(r.blocked >> die) & ((1u << (24 - die)) - 1u))
                              : (occ & ~(r.blocked << die) & (0xffffffu & ~((1u << die) - 1u)));
  uint32_t last = 0;
  if (r.valid15 && (occ & ~r.home) == 0) {  // BEAR_OFF (:192-249)
    last = r.player == 0 ? (occ ? __ffs(occ) - 1 : 18) : (occ ? 31 - __clz(occ) : 5);
    const bool far_off = r.player == 0 ? ((int)last + die >= 24) : ((int)last - die < 0);
    const uint32_t ps = r.player == 0 ? 24 - die : die - 1;
    if (far_off) vm |= 1u << 25;
    if (ps != last && ((occ >> ps) & 1u)) vm |= 1u << 26;
  }
  return vm | (last << 27);
}

// n-th (0-based) set bit of m
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
  int pos = 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const int c = __popc((m >> pos) & ((1u << s) - 1u));
    if (c <= n) {
      n -= c;
      pos += s;
    }
  }
  return pos;
}

// apply the move in `slot` of parent p (immutable_board.py:183-258 on the packed key)
__device__ __forceinline__ void make_child(const Node& p, const Root& r, int slot, uint32_t last, int die, int depth, Node& c) {
  int s, e;
  if (slot < 24) {
    s = slot;
    e = slot + r.dirsign * die;
  } else if (slot == 24) {
    s = 24;
    e = r.player == 0 ? die - 1 : 24 - die;
  } else {
    s = slot == 25 ? (int)last : (r.player == 0 ? 24 - die : die - 1);
    e = 25;
  }
  uint32_t k[3] = {p.k0, p.k1, p.k2};
  uint32_t k3 = p.k3;
  uint32_t hit = 0;
  if (s == 24) {
    k3 -= 1u << 24;
  } else {
    const uint32_t ds = 1u << ((s & 7) * 4);
    const int ws = s >> 3;
    k[0] -= ws == 0 ? ds : 0u;
    k[1] -= ws == 1 ? ds : 0u;
    k[2] -= ws == 2 ? ds : 0u;
  }
  if (e == 25) {
    k3 += 1u << 28;
  } else {
    const uint32_t de = 1u << ((e & 7) * 4);
    const int we = e >> 3;
    k[0] += we == 0 ? de : 0u;
    k[1] += we == 1 ? de : 0u;
    k[2] += we == 2 ? de : 0u;
    hit = ((r.blot & ~p.k3) >> e) & 1u;
    k3 |= hit << e;
  }
  const uint32_t sm = (uint32_t)s | ((uint32_t)e << 5) | (hit << 10) | (1u << 11);
  c.k0 = k[0];
  c.k1 = k[1];
  c.k2 = k[2];
  c.k3 = k3;
  c.m0 = depth < 2 ? (p.m0 | (sm << (16 * depth))) : p.m0;
  c.m1 = depth < 2 ? p.m1 : (p.m1 | (sm << (16 * (depth - 2))));
}

}  // namespace bg
