/*
 * bgarena.h -- C ABI of libbgarena.so, the B200-native (sm_100a) backgammon self-play hot path.
 *
 * Drop-in boundary for the per-decision hot path of Nick-qsv/MLP-PPO-2PLY-MULTI (a pure-Python
 * reference with no FFI of its own).  Each entry point names the reference
 * interface it replaces (file:line relative to the reference root).  A maintainer binds these with
 * ctypes (INTEGRATION.md shows the stub); mlp-ppo-2ply-multi_b200/ is that binding plus a host-side
 * mirror of the reference's Python API.
 *
 * Conventions
 *   - plain C symbols; no torch / C++ types; all buffers are CALLER-OWNED raw pointers
 *     (device pointers unless the parameter says "host"); `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream).  Calls are asynchronous with respect to the host
 *     unless documented otherwise.
 *   - every function returns a status: BG_OK, or < 0 (see below).  Nothing throws or aborts.
 *     bg_last_error() gives a thread-local description of the last failure.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns BG_ERR_CUDA.
 *
 * Data model (reference: src/backgammon/board/immutable_board.py:16-24, src/backgammon/types/moves.py:7-46)
 *   board  : int8[52]  = positions_0[24] | positions_1[24] | bar[2] | borne_off[2]   (absolute points)
 *   player : uint8      0 = PLAYER1 (moves 0->23, home 18..23), 1 = PLAYER2 (moves 23->0, home 0..5)
 *   roll   : uint8[2]   two dice 1..6 (any order)
 *   submove: uint8[3]   (start, end, hits_blot); BAR = 24, BEAR_OFF = 25; unused slots are (255,255,0)
 *   weights: float32    packed [ W1^T (198 x H) | b1 (H) | w2 (H) | b2 (1) ]  from
 *                       BackgammonPolicyNetwork.state_dict()  (src/agents/policy_network.py:36-51);
 *                       H must be a multiple of 32, 32 <= H <= 256.
 */
#ifndef BGARENA_H
#define BGARENA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_ABI_VERSION 1

#define BG_OK 0
#define BG_ERR_ARG (-1)       /* bad argument */
#define BG_ERR_CUDA (-2)      /* CUDA runtime error / no device */
#define BG_ERR_CAPACITY (-3)  /* an output pool or per-item capacity was exceeded (results incomplete) */
#define BG_ERR_INVARIANT (-4) /* an input violated a board invariant */

#define BG_BOARD_BYTES 52
#define BG_NUM_FEATURES 198
#define BG_MAX_ITEM_MOVES 4096 /* hard per-(board,roll) capacity of the move generator */

int32_t bg_abi_version(void);
const char* bg_last_error(void);
/* number of visible CUDA devices (0 on a CPU-only host); never fails */
int32_t bg_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Stateless batch operators
 * ---------------------------------------------------------------------------------------------- */

/* Bytes of device scratch bg_movegen needs for a batch of B items. */
int64_t bg_movegen_workspace_bytes(int64_t B);

/*
 * Legal-move generation for B independent (board, player, roll) items.
 * Replaces get_all_possible_moves(player, board, roll_result)  (src/backgammon/moves/generate_all_moves.py:7-66)
 * followed by execute_full_move_on_board_copy per move           (src/environments/env_helper.py:27-91).
 *
 * Item i's legal afterstates are written, IN THE REFERENCE'S ORDER (first-occurrence DFS order after the
 * max-sub-moves filter; the index is the reference's action index), to the contiguous pool segment
 *     out_boards[out_offsets[i] .. out_offsets[i] + min(out_count[i], item_cap))
 * Segments are packed into the pool in an unspecified (run-dependent) order; segment CONTENT is deterministic.
 *   item_cap     : keep at most this many afterstates per item (the reference env truncates at 500,
 *                  src/environments/backgammon_env.py:262-272); out_count still holds the true count.
 *   pool_cap     : capacity of out_boards / out_submoves / out_owner in afterstates.
 *   out_submoves : optional (may be NULL) [pool_cap,4,3] the FullMove sub-move sequences.
 *   out_owner    : optional (may be NULL) [pool_cap] item index owning each pool slot.
 *   out_total    : [1] device int64, number of pool slots used.
 *   out_status   : [1] device int32, BG_OK or BG_ERR_CAPACITY (pool or BG_MAX_ITEM_MOVES exceeded; the
 *                  offending items get out_offsets = -1).
 */
int32_t bg_movegen(const int8_t* boards /*[B,52]*/, const uint8_t* players /*[B]*/, const uint8_t* rolls /*[B,2]*/,
                   int64_t B, int32_t item_cap, int64_t pool_cap, int8_t* out_boards /*[pool_cap,52]*/,
                   uint8_t* out_submoves /*[pool_cap,4,3] or NULL*/, int32_t* out_owner /*[pool_cap] or NULL*/,
                   int64_t* out_offsets /*[B]*/, int32_t* out_count /*[B]*/, int64_t* out_total /*[1]*/,
                   int32_t* out_status /*[1]*/, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * 198-feature Tesauro-style encoding.  Replaces ImmutableBoard.get_board_features(player)
 * (src/backgammon/board/immutable_board.py:86-128) / generate_all_board_features (src/environments/env_helper.py:7-24).
 * flag_player is the player whose indicator feature (196/197) is set.  Bit-exact fp32.
 */
int32_t bg_encode(const int8_t* boards /*[N,52]*/, const uint8_t* flag_player /*[N]*/, int64_t N,
                  float* out /*[N,198]*/, void* stream);

/* Bytes of the device-side prepared weight table built by bg_prepare_weights (depends on H only). */
int64_t bg_prepared_weights_bytes(int32_t H);
/* Re-layout packed weights for the fused evaluator (cumulative per-point rows; done once per weight set). */
int32_t bg_prepare_weights(const float* packed /*device*/, int32_t H, float* prepared /*device*/, void* stream);

/*
 * Fused encode + value network: V = w2 . sigmoid(W1 x + b1) + b2 straight from boards; the 198 features
 * are never materialised.  Replaces generate_all_board_features + BackgammonPolicyNetwork.forward
 * (src/agents/policy_network.py:53-70).  fp32, |V - reference| <= 1e-5.
 * flag_player may be given per row (flags != NULL) or through owner indirection
 * (flags == NULL: flag = owner_players[owner[i]]), which is how afterstate pools are evaluated.
 */
int32_t bg_eval(const int8_t* boards /*[N,52]*/, const uint8_t* flags /*[N] or NULL*/, const int32_t* owner /*[N] or NULL*/,
                const uint8_t* owner_players /*[.] or NULL*/, int64_t N, const float* prepared, int32_t H,
                float* out_v /*[N]*/, void* stream);
/* same, N read from a device counter (e.g. bg_movegen's out_total) so no host sync is needed; N <= max_N */
int32_t bg_eval_indirect(const int8_t* boards, const uint8_t* flags, const int32_t* owner, const uint8_t* owner_players,
                         const int64_t* N_dev, int64_t max_N, const float* prepared, int32_t H, float* out_v, void* stream);

/*
 * Action selection over ragged value segments.  Replaces softmax(V/T) + Categorical.sample
 * (src/multi/worker.py:136-143) and, for temperature <= 0, torch.argmax (lowest index on ties,
 * src/play/play_versus_ai.py:188-195).  Randomness: Philox4x32-10 keyed by (seed, item id, ctr);
 * sampling is distributionally (not bitwise) equivalent to the reference, which uses an unseeded global RNG.
 * Items with count == 0 get action -1.
 */
int32_t bg_select(const float* v, const int64_t* offsets /*[B]*/, const int32_t* counts /*[B]*/, int32_t item_cap,
                  int64_t B, float temperature, uint64_t seed, uint64_t ctr, int64_t item_id_base,
                  int32_t* out_action /*[B]*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BGARENA_H */
