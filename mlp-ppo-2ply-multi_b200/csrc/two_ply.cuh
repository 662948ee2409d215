// two_ply.cuh -- host-side interface of the 2-ply scorer (see two_ply.cu)
#pragma once
#include "bg_common.cuh"
#include "eval.cuh"

namespace bg {

struct TwoPlyArgs {
  const int8_t* cand_boards;
  const uint8_t* mover;
  const float* S;
  int64_t N;
  const float* prepared;
  int32_t H;
  int32_t top_k;
  float alpha, beta;
  float* out_score;
  int64_t* out_replies;
  int32_t* out_status;
  void* workspace;
  int64_t workspace_bytes;
  const uint8_t* cand_active = nullptr;            // optional [N]: inactive candidates are skipped (score = alpha * S)
  unsigned long long* reply_counter = nullptr;     // optional device counter: += replies evaluated
  SideCtx* side = nullptr;                         // optional: evaluate each chunk's bulk-tier replies while its tail tiers run
};

int64_t two_ply_workspace_bytes(int64_t N);
// process-wide option of every 2-ply scorer call (bg_two_ply, bg_arena_step(lookahead = 2)): cap > 0 cuts the replies of 1-1 / 2-2 / 3-3 to a
// uniform sample of `cap` before they are evaluated (reference two_ply.py:119-121 with cap = 50); returns the previous cap
int32_t two_ply_reply_sampling(int32_t cap, uint64_t seed);
int32_t two_ply_launch(const TwoPlyArgs& a, cudaStream_t s);

}  // namespace bg
