// movegen.cuh -- host-side launch interface of the move generator (see movegen.cu)
#pragma once
#include "bg_common.cuh"

namespace bg {

struct MovegenArgs {
  const int8_t* boards;
  const uint8_t* players;
  const uint8_t* rolls;
  int64_t B;
  int32_t item_cap;
  int64_t pool_cap;
  int8_t* out_boards;
  uint8_t* out_submoves;
  int32_t* out_owner;
  uint8_t* out_flags;  // optional [pool_cap]: mover (== feature flag player) of each row
  int64_t* out_offsets;
  int32_t* out_count;
  int64_t* out_total;
  int32_t* out_status;
  void* workspace;
  int64_t workspace_bytes;
  const uint8_t* active;  // optional [B]: items with active[i] == 0 are skipped (count 0)
  // optional hook after the first (bulk) tier: the pool rows [0, *tier1_total) are final once tier1_event has fired, so a consumer on
  // another stream can start on them while the tail tiers (few, heavy items) are still running
  int64_t* tier1_total = nullptr;
  cudaEvent_t tier1_event = nullptr;
  // optional: resident CTAs per SM of the second tier (0 = default).  A consumer that overlaps the tail tiers with another kernel
  // lowers this so that both fit on an SM at the same time
  int32_t tier2_ctas_per_sm = 0;
  // position-major mode (movegen21.cu): boards / players / active are per POSITION (B of them), `rolls` is ignored, and the items are
  // position * 21 + roll index in the roll order of src/multi/two_ply.py:10-32; out_offsets / out_count have 21 * B entries
  int32_t all_rolls = 0;
  // compact mode (position-major batches only): the pool holds (code, position index) pairs (codes.cuh) instead of boards; out_boards is
  // ignored.  8 bytes per afterstate instead of 53, and no board is materialised unless a consumer asks for it.
  uint2* out_codes = nullptr;
};

// kernel parameter block
struct MovegenParams {
  const int8_t* boards;
  const uint8_t* players;
  const uint8_t* rolls;
  int64_t B;
  int32_t item_cap;
  int64_t pool_cap;
  int8_t* out_boards;
  uint8_t* out_submoves;
  int32_t* out_owner;
  uint8_t* out_flags;
  long long* out_offsets;
  int32_t* out_count;
  unsigned long long* pool_cursor;
  int32_t* status;
  int32_t* item_counter;
  const int32_t* in_list;
  const int32_t* in_count;
  int32_t* ovf_list;
  int32_t* ovf_count;
  uint32_t* gfront;
  int32_t grab;
  const uint8_t* active;
  int32_t all_rolls;  // items are position * 21 + roll index; boards / players / active are indexed by position
  uint2* out_codes;   // compact mode: (code, position index) per pool row instead of a board
};

int64_t movegen_workspace_bytes(int64_t B);  // B = number of ITEMS (21 per position in position-major mode)
// movegen21.cu: the position-major bulk tier (one warp per position, all 21 rolls); overflowing items go to P.ovf_list
constexpr int MOVEGEN21_MIN_ITEM_CAP = 1;  // the position-major tier truncates an item to item_cap rows at flush time
int32_t movegen21_launch_kernel(const MovegenParams& P, cudaStream_t stream, int big = 0);
int32_t movegen_launch(const MovegenArgs& a, cudaStream_t stream);

}  // namespace bg
