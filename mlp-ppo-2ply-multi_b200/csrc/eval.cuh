// eval.cuh -- host-side launch interface of the fused evaluator / encoder (see eval.cu)
#pragma once
#include "bg_common.cuh"
#include "movegen.cuh"

namespace bg {

struct EvalArgs {
  const int8_t* boards;
  const uint8_t* flags;
  const int32_t* owner;
  const uint8_t* owner_players;
  int64_t N;
  const int64_t* N_dev;
  int64_t max_N;
  const float* prepared;
  int32_t H;
  float* out_v;
  const int64_t* start_dev = nullptr;  // optional device counter: evaluate rows [*start_dev, N) only (rows before it are left untouched)
  // compact pool (codes.cuh): rows are (code, position index) pairs; `boards` / `flags` are then the POSITIONS' boards and players.  Always
  // evaluated by the tensor-core kernel.
  const uint2* codes = nullptr;
  // tensor-core kernel: -1 = choose by size (dynamic tile schedule up to EVAL_TC_DYNAMIC_MAX_ROWS rows, see eval_tc.cu), 0 = static, 1 = dynamic
  int32_t dynamic_tiles = -1;
};
constexpr int64_t EVAL_TC_DYNAMIC_MAX_ROWS = 1ll << 23;

int64_t prepared_weights_bytes(int32_t H);
int32_t prepare_weights_launch(const float* packed, int32_t H, float* prepared, cudaStream_t stream);
int32_t eval_launch(const EvalArgs& a, cudaStream_t stream);
// H == 128 production kernel (eval128.cu); its table is appended to the generic prepared table
int64_t eval128_table_floats();
int32_t eval128_prepare(const float* packed, float* t128, cudaStream_t stream);
int32_t eval128_launch(const EvalArgs& a, const float* t128, cudaStream_t stream);
// tensor-core (tcgen05 / TMEM) kernel (eval_tc.cu): 128 hidden units per pass.  Smaller nets are zero-padded to 128 units; wider nets
// (H <= 256) take two passes over the boards, the second one accumulating the other half of the units into out_v.  The operand
// image(s) follow the generic (and, for H == 128, the eval128) table
int64_t eval_tc_image_bytes();
int32_t eval_tc_prepare(const float* packed, int32_t H_src, int32_t unit0, uint8_t* img, cudaStream_t stream);
int32_t eval_tc_launch(const EvalArgs& a, const uint8_t* img, int32_t* err_flag, cudaStream_t stream, int accumulate = 0);
int32_t eval_tc_tile_schedule(int32_t mode);  // -1 by size, 0 static, 1 dynamic (process-wide default of EvalArgs::dynamic_tiles < 0); returns the previous one
int32_t eval_tc_status();  // synchronising: 0 ok, != 0 a bounded mbarrier wait timed out in k_eval_tc
// Move generation + evaluation of the whole afterstate pool with the tail tiers overlapped: the rows written by the move generator's
// bulk tier are evaluated on `side->stream` as soon as that tier is done, while `stream` runs the tail tiers (a few very wide doubles
// trees) and then evaluates only the rows they added; `stream` joins before returning.  `total2` is a device int64[2]
// ({rows used, rows of the bulk tier}); m.out_total / m.tier1_* are set here.  side == nullptr: plain sequential launch.
struct SideCtx {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_t1 = nullptr, ev_side = nullptr;
};
int32_t side_ctx_create(SideCtx* c, bool high_priority);
void side_ctx_destroy(SideCtx* c);
int32_t movegen_eval_overlapped(MovegenArgs m, int64_t* total2, const float* prepared, int32_t H, float* out_v, SideCtx* side,
                                cudaStream_t stream);
int32_t materialize_launch(const int8_t* boards, const uint8_t* players, const uint2* codes, const int64_t* rows, int64_t n, int8_t* out,
                           cudaStream_t stream);
int32_t encode_launch(const int8_t* boards, const uint8_t* flags, int64_t N, float* out, cudaStream_t stream);

}  // namespace bg
