// arena.cuh -- host-side interface of the self-play arena (see arena.cu)
#pragma once
#include <vector>

#include "bg_common.cuh"

namespace bg {

struct Arena;
int arena_device(const Arena* A);  // the device the handle lives on

int32_t arena_create(Arena** out, int32_t device, int64_t n_games, int32_t H, int32_t max_plies, int32_t move_cap, uint64_t seed,
                     int64_t game_id_base, int64_t ring_exps, int64_t ring_eps, int32_t auto_reset);
int32_t arena_destroy(Arena* A);
int32_t arena_set_weights(Arena* A, const float* packed_dev, int64_t version, float temperature, cudaStream_t s);
int32_t arena_set_dice_tape(Arena* A, const uint8_t* tape, int64_t L, cudaStream_t s);
int32_t arena_reset(Arena* A, cudaStream_t s);
int32_t arena_set_lookahead(Arena* A, int32_t n_candidates, int32_t top_k, float alpha, float beta);
int32_t arena_step(Arena* A, int32_t n_plies, int32_t lookahead, const int32_t* forced_action, cudaStream_t s);
int32_t arena_drain(Arena* A, int64_t max_eps, int64_t max_exps, int8_t* after, uint8_t* meta, float* reward, float* v, float* vnext,
                    int16_t* nmoves, int16_t* action, uint8_t* roll, int64_t* ep_offsets, int32_t* ep_info, int64_t* out_n, cudaStream_t s);
int32_t arena_stats(Arena* A, int64_t* out, cudaStream_t s);
int32_t arena_export_state(Arena* A, int8_t* boards, uint8_t* players, uint8_t* rolls, uint8_t* gstate, cudaStream_t s);

}  // namespace bg
