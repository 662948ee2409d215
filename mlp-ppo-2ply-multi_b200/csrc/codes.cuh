// codes.cuh -- the result codes of the code-based move generator (movegen21.cu) and how to apply one to a board.
//
// A legal afterstate of (position, roll) is identified by one 32-bit word (bits 25..29: roll index + 1; movegen21.cu proves equal codes <=>
// equal boards):
//   non-double : sorted sources [0:5) [5:10), sorted destinations [10:15) [15:20) of the two sub-moves after cancelling a point that is both,
//                [20:25) the intermediate point whose blot was hit by a checker that moved on (31 = none); a single sub-move has 31 in the
//                second source / destination;
//   double d-d : the sorted multiset of up to four source points [0:20) (31 = none), [20:25) the slot of the last sub-move (unused here).
// Points are 0..23, BAR = 24 as a source, BEAR_OFF = 25 as a destination.  In COMPACT mode the pool holds (code, position index) pairs instead
// of 52-byte boards; consumers rebuild a board from the position's board + the code with apply_code_bytes (the evaluator does it while it
// builds its feature rows, bg_afterstates_from_codes for the rows a caller wants to see).
#pragma once
#include "bg_common.cuh"

namespace bg {

constexpr uint32_t CODE_NONE = 31u;
constexpr uint32_t CODE_DBL_IDS = (1u << 0) | (1u << 6) | (1u << 11) | (1u << 15) | (1u << 18) | (1u << 20);  // roll indices of d-d

__device__ __forceinline__ bool code_is_double(uint32_t code) { return (CODE_DBL_IDS >> ((code >> 25) - 1u)) & 1u; }
__device__ __forceinline__ int code_die(uint32_t code) { return __popc(CODE_DBL_IDS & ((1u << ((code >> 25) - 1u)) - 1u)) + 1; }

// rb: the 52 bytes of the POSITION's board (positions_0[24] | positions_1[24] | bar[2] | off[2]), rewritten in place into the afterstate of
// `player` described by `code` (immutable_board.py:183-258 per sub-move: a landing on a point where the opponent has exactly one checker hits it)
__device__ __forceinline__ void apply_code_bytes(uint8_t* rb, uint32_t code, int player) {
  const int own = player * 24, opp = 24 - own;
  const bool dbl = code_is_double(code);
  const int die = code_die(code);
  const int dirsign = player == 0 ? 1 : -1;
  const int nq = dbl ? 4 : 2;
#pragma unroll 1
  for (int q = 0; q < nq; ++q) {
    const uint32_t s = (code >> (5 * q)) & 31u;
    if (s == CODE_NONE) break;
    uint32_t e;
    if (dbl) {
      const int ee = (int)s + dirsign * die;
      e = s == 24u ? (uint32_t)(player == 0 ? die - 1 : 24 - die) : ((ee < 0 || ee > 23) ? 25u : (uint32_t)ee);
    } else {
      e = (code >> (10 + 5 * q)) & 31u;
    }
    rb[s == 24u ? 48 + player : own + (int)s] -= 1;
    if (e == 25u) {
      rb[50 + player] += 1;
    } else {
      rb[own + (int)e] += 1;
      if (rb[opp + (int)e] == 1) {  // blot hit (a second checker landing here finds the point already cleared)
        rb[opp + (int)e] = 0;
        rb[48 + 1 - player] += 1;
      }
    }
  }
  if (!dbl) {
    const uint32_t in = (code >> 20) & 31u;
    if (in != CODE_NONE) {  // the blot on the point a checker only passed through
      rb[opp + (int)in] -= 1;
      rb[48 + 1 - player] += 1;
    }
  }
}

}  // namespace bg
