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
 *   - bg_movegen_eval, bg_two_ply and bg_arena_step overlap independent kernels on library-owned non-blocking streams; these are
 *     forked from and joined back into `stream` inside the call, so the caller sees ordinary stream ordering on `stream`.
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
 *   out_flags    : optional (may be NULL) [pool_cap] the item's player for each slot (the flag bg_eval needs for afterstates).
 *   out_total    : [1] device int64, number of pool slots used.
 *   out_status   : [1] device int32, BG_OK or BG_ERR_CAPACITY (pool or BG_MAX_ITEM_MOVES exceeded; the
 *                  offending items get out_offsets = -1).
 */
int32_t bg_movegen(const int8_t* boards /*[B,52]*/, const uint8_t* players /*[B]*/, const uint8_t* rolls /*[B,2]*/,
                   int64_t B, int32_t item_cap, int64_t pool_cap, int8_t* out_boards /*[pool_cap,52]*/,
                   uint8_t* out_submoves /*[pool_cap,4,3] or NULL*/, int32_t* out_owner /*[pool_cap] or NULL*/,
                   uint8_t* out_flags /*[pool_cap] or NULL*/, int64_t* out_offsets /*[B]*/, int32_t* out_count /*[B]*/, int64_t* out_total /*[1]*/,
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
 * Fused bg_movegen + bg_eval_indirect over the whole afterstate pool (what the per-decision hot path always does:
 * worker.py:101-125 evaluates every legal afterstate).  Same results as the two calls; internally the evaluation of the rows
 * written by the move generator's bulk tier runs on a library-owned high-priority stream while the tail tiers (the few very wide
 * doubles trees) are still being generated on `stream`, which then evaluates only the rows they added and joins.
 *   out_flags : required [pool_cap];  out_total : device int64[2] = {rows used, rows written by the bulk tier};  out_v : [pool_cap].
 */
int32_t bg_movegen_eval(const int8_t* boards /*[B,52]*/, const uint8_t* players /*[B]*/, const uint8_t* rolls /*[B,2]*/, int64_t B,
                        int32_t item_cap, int64_t pool_cap, int8_t* out_boards /*[pool_cap,52]*/, uint8_t* out_flags /*[pool_cap]*/,
                        int64_t* out_offsets /*[B]*/, int32_t* out_count /*[B]*/, int64_t* out_total /*[2]*/, int32_t* out_status /*[1]*/,
                        void* workspace, int64_t workspace_bytes, const float* prepared, int32_t H, float* out_v /*[pool_cap]*/,
                        void* stream);

/*
 * Position-major forms of bg_movegen / bg_movegen_eval: EVERY one of the 21 unordered rolls of each of P positions, i.e. the batch
 * shape of BASELINE config 2 (positions x 21 rolls) and of the 2-ply lookahead (compute_weighted_opponent_response iterates
 * DICE_ROLLS for each candidate board, src/multi/two_ply.py:93-150).  Item index = position * 21 + r with r indexing DICE_ROLLS
 * (src/multi/two_ply.py:10-32: (1,1), (1,2), ... (1,6), (2,2), ... (6,6)); out_offsets / out_count have 21 * P entries; results,
 * order and pool layout rules are those of bg_movegen (get_all_possible_moves per item).  One warp expands a whole position
 * (csrc/movegen21.cu): the inputs are 1/21 of the replicated item form and the first two plies are shared by all rolls.
 * Workspace: bg_movegen_workspace_bytes(21 * P).  out_submoves != NULL or item_cap < 320 routes through the per-item kernels.
 */
int32_t bg_movegen_all_rolls(const int8_t* boards /*[P,52]*/, const uint8_t* players /*[P]*/, int64_t P, int32_t item_cap,
                             int64_t pool_cap, int8_t* out_boards /*[pool_cap,52]*/, uint8_t* out_submoves /*[pool_cap,4,3] or NULL*/,
                             int32_t* out_owner /*[pool_cap] or NULL*/, uint8_t* out_flags /*[pool_cap] or NULL*/,
                             int64_t* out_offsets /*[21*P]*/, int32_t* out_count /*[21*P]*/, int64_t* out_total /*[1]*/,
                             int32_t* out_status /*[1]*/, void* workspace, int64_t workspace_bytes, void* stream);
int32_t bg_movegen_eval_all_rolls(const int8_t* boards /*[P,52]*/, const uint8_t* players /*[P]*/, int64_t P, int32_t item_cap,
                                  int64_t pool_cap, int8_t* out_boards /*[pool_cap,52]*/, uint8_t* out_flags /*[pool_cap]*/,
                                  int64_t* out_offsets /*[21*P]*/, int32_t* out_count /*[21*P]*/, int64_t* out_total /*[2]*/,
                                  int32_t* out_status /*[1]*/, void* workspace, int64_t workspace_bytes, const float* prepared, int32_t H,
                                  float* out_v /*[pool_cap]*/, void* stream);

/*
 * COMPACT forms of the position-major operators: the pool holds one uint64 per legal afterstate -- low word = the move's CODE (sorted sources /
 * destinations of the sub-moves, see csrc/codes.cuh), high word = the index of the position it applies to -- instead of a 52-byte board.
 * This is what callers that only need values / actions want (worker.py:101-143 looks at the afterstates' VALUES and then steps the env
 * with the chosen index; two_ply.py:93-150 looks at reply values only): move generation does not materialise 25 GB of boards per 10^6
 * positions and the evaluator rebuilds each afterstate on chip while it builds its feature row (bg_eval_codes / the fused call).
 * Order, counts, offsets and values are exactly those of the board forms.  bg_afterstates_from_codes materialises the boards of selected
 * pool rows (rows[j] < 0 -> zeros; rows == NULL -> rows 0..n-1), e.g. the afterstate of each item's chosen action: rows[i] = out_offsets[i] + action[i].
 * Items whose tree exceeds the code-based tiers (more than ~6,000 afterstates; none is known for a legal position) get BG_ERR_CAPACITY.
 */
int32_t bg_movegen_all_rolls_compact(const int8_t* boards /*[P,52]*/, const uint8_t* players /*[P]*/, int64_t P, int32_t item_cap,
                                     int64_t pool_cap, uint64_t* out_codes /*[pool_cap]*/, int64_t* out_offsets /*[21*P]*/,
                                     int32_t* out_count /*[21*P]*/, int64_t* out_total /*[1]*/, int32_t* out_status /*[1]*/, void* workspace,
                                     int64_t workspace_bytes, void* stream);
/* values of the rows of a compact pool; N rows, or *N_dev (<= max_N) when N_dev != NULL */
int32_t bg_eval_codes(const int8_t* boards /*[P,52]*/, const uint8_t* players /*[P]*/, const uint64_t* codes /*[N]*/, int64_t N,
                      const int64_t* N_dev /*or NULL*/, int64_t max_N, const float* prepared, int32_t H, float* out_v /*[N]*/, void* stream);
int32_t bg_movegen_eval_all_rolls_compact(const int8_t* boards /*[P,52]*/, const uint8_t* players /*[P]*/, int64_t P, int32_t item_cap,
                                          int64_t pool_cap, uint64_t* out_codes /*[pool_cap]*/, int64_t* out_offsets /*[21*P]*/,
                                          int32_t* out_count /*[21*P]*/, int64_t* out_total /*[2]*/, int32_t* out_status /*[1]*/,
                                          void* workspace, int64_t workspace_bytes, const float* prepared, int32_t H,
                                          float* out_v /*[pool_cap]*/, void* stream);
int32_t bg_afterstates_from_codes(const int8_t* boards /*[P,52]*/, const uint8_t* players /*[P]*/, const uint64_t* codes /*pool*/,
                                  const int64_t* rows /*[n] or NULL*/, int64_t n, int8_t* out_boards /*[n,52]*/, void* stream);

/* Diagnostic for the tcgen05 evaluator (batches >= 32768 rows with per-row flags, any H: 128 hidden units per pass, smaller nets
 * zero-padded, wider nets in two passes; set BG_EVAL_PATH=ffma to force the CUDA-core kernels): synchronises and returns 0, or non-zero if one of its bounded mbarrier waits ever timed out. */
int32_t bg_eval_tc_status(void);
/* Tile schedule of the tcgen05 evaluator: -1 (default) = by size -- launches of up to 2^23 rows (a self-play ply, a 2-ply candidate set: they
 * run next to other kernels) claim their 128-row tiles from a grid-wide counter, larger (bulk) passes split them statically --, 0 = always
 * static, 1 = always dynamic.  Process-wide; returns the previous setting.  Results do not depend on it. */
int32_t bg_eval_tc_tile_schedule(int32_t mode);

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

/*
 * 2-ply lookahead score of N candidate afterstates.  Replaces compute_scores_for_boards /
 * compute_weighted_opponent_response (src/multi/two_ply.py:44-150): for each candidate (board after `mover` moved),
 * W = sum over the 21 unordered opponent rolls (1/36 doubles, 2/36 others) of mean(top_k opponent-reply values)
 * (all replies if fewer than top_k; 0 for a roll without reply), score = alpha * S - beta * W.
 * Reference setting: top_k = 5, alpha = 1.0, beta = 0.9; north_star's best-reply expectimax: top_k = 1.
 * The reference's random.sample(replies, 50) on 1-1/2-2/3-3 is not reproduced (every reply is evaluated).
 *   S           : [N] 1-ply value of each candidate (mover's flag)
 *   out_replies : optional [N] number of replies evaluated per candidate
 *   out_status  : optional [1] device int32, BG_OK or BG_ERR_CAPACITY (such candidates get score NaN)
 *   workspace   : device scratch, bg_two_ply_workspace_bytes(N) recommended; smaller is legal (more, smaller chunks)
 */
int64_t bg_two_ply_workspace_bytes(int64_t N);
/* Optional: the reference cuts the reply list of the rolls 1-1, 2-2 and 3-3 to random.sample(opponent_moves, 50) before evaluating it
 * (src/multi/two_ply.py:119-121; nondeterministic there).  cap > 0 switches the same cut on for every later bg_two_ply / 2-ply bg_arena_step of
 * the process: a uniform sample WITHOUT replacement of `cap` replies, reproducible: keyed by (seed, candidate index, roll).  Row j of an
 * item's sample is its reply perm(j): perm = 4-round Feistel permutation over the even number of bits that covers the reply count n,
 * cycle-walked into [0, n), round function mix32(R, key[round], round, 0x9E3779B9) of csrc/bg_common.cuh, keys =
 * Philox4x32-10(seed ^ 0x3c6ef372fe94f82b, counter = candidate * 21 + roll index).  A second reply pool lives in the workspace while the
 * option is on: bg_two_ply_workspace_bytes includes it when called with the option on; a workspace sized before is still legal (smaller
 * chunks).  cap <= 0 (default): every reply is evaluated.  Returns the previous cap. */
int32_t bg_two_ply_reply_sampling(int32_t cap, uint64_t seed);
int32_t bg_two_ply(const int8_t* cand_boards /*[N,52]*/, const uint8_t* mover /*[N]*/, const float* S /*[N]*/, int64_t N,
                   const float* prepared, int32_t H, int32_t top_k, float alpha, float beta, float* out_score /*[N]*/,
                   int64_t* out_replies /*[N] or NULL*/, int32_t* out_status /*[1] or NULL*/, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-resident batches: the same hot path with HOST pointers in and out (what a CPU-side caller of get_all_possible_moves +
 * generate_all_board_features + policy_network.forward + argmax / Categorical.sample makes, src/multi/worker.py:101-143).
 * The batch is cut into chunks of chunk_units that rotate over n_streams library-owned streams, so that the host->device copy of
 * chunk k+1 and the device->host copy of chunk k-1 overlap the kernels of chunk k (bg_movegen_eval[_all_rolls] + bg_select); all
 * device buffers are allocated once, here.  Pinned host memory makes the copies asynchronous; pageable memory works, serialised.
 *   all_rolls != 0 : units are POSITIONS, every one expanded to the 21 rolls of DICE_ROLLS (bg_movegen_all_rolls); h_rolls is
 *                    ignored and h_actions / h_counts have 21 * n_units entries (item = position * 21 + roll index);
 *   all_rolls == 0 : units are (board, player, roll) items (bg_movegen).
 *   rows_per_item  : pool rows provisioned per item of a chunk (the mean is ~22; 26 + a fixed slack is what bench.py uses).
 * bg_hostpipe_run enqueues and returns; `stream` is ordered after every chunk (synchronise it before reading the outputs).
 * h_actions[i] = selected action (argmax for temperature <= 0, else softmax(V/T) sample keyed by (seed, i)), -1 without a legal move;
 * h_counts[i] = true legal-move count.  bg_hostpipe_status synchronises the pipe's streams and returns the worst status of every chunk
 * since the previous call (BG_OK, BG_ERR_CAPACITY, BG_ERR_INVARIANT) in *out_status.
 * ---------------------------------------------------------------------------------------------- */
typedef struct bg_hostpipe bg_hostpipe;
int32_t bg_hostpipe_create(bg_hostpipe** out, int32_t device, int32_t H, int64_t chunk_units, int32_t all_rolls, int32_t item_cap,
                           int32_t rows_per_item, int32_t n_streams);
int32_t bg_hostpipe_destroy(bg_hostpipe* p);
int32_t bg_hostpipe_run(bg_hostpipe* p, const int8_t* h_boards /*host [n,52]*/, const uint8_t* h_players /*host [n]*/,
                        const uint8_t* h_rolls /*host [n,2] or NULL*/, int64_t n_units, const float* prepared /*device*/, float temperature,
                        uint64_t seed, int32_t* h_actions /*host*/, int32_t* h_counts /*host*/, void* stream);
int32_t bg_hostpipe_status(bg_hostpipe* p, int32_t* out_status /*host*/);

/* ------------------------------------------------------------------------------------------------
 * Stateful self-play arena: n_games concurrent games resident on one GPU.
 * Replaces the reference's worker processes: Worker.play_episode (src/multi/worker.py:78-174) over
 * BackgammonEnv.reset/step (src/environments/backgammon_env.py:92-329), Experience/Episode recording
 * (src/environments/episode.py:5-84), the ExperienceQueue (src/multi/experience_queue.py:5-13) and the
 * ParameterManager polling (src/multi/worker.py:66-76).  One opaque handle per device, one host thread per handle.
 * ---------------------------------------------------------------------------------------------- */
typedef struct bg_arena bg_arena;

#define BG_GAME_ACTIVE 0  /* playing */
#define BG_GAME_WAIT 1    /* episode finished, waiting for room in the episode ring (drain it) */
#define BG_GAME_STOPPED 2 /* no auto-reset / dice tape exhausted / per-game error */

/* bg_arena_stats layout (int64[BG_ARENA_NSTATS]); all counters are cumulative since bg_arena_reset */
#define BG_ARENA_NSTATS 16
#define BG_STAT_GAMES 0        /* finished episodes */
#define BG_STAT_STEPS 1        /* env steps (plies incl. passes) of finished episodes */
#define BG_STAT_DECISIONS 2    /* recorded experiences of finished episodes */
#define BG_STAT_PASSES 3
#define BG_STAT_AFTERSTATES 4  /* afterstates evaluated at decisions (sum of num_moves), all episodes */
#define BG_STAT_WIN_REGULAR 5
#define BG_STAT_WIN_GAMMON 6
#define BG_STAT_WIN_BACKGAMMON 7
#define BG_STAT_TRUNCATED 8    /* episodes that hit max_plies without a winner */
#define BG_STAT_P1_WINS 9
#define BG_STAT_WAIT_STEPS 10  /* game-steps spent waiting for ring space (actor idle) */
#define BG_STAT_ERRORS 11      /* games stopped by a per-item move-generator overflow */
#define BG_STAT_REPLIES 12     /* opponent replies evaluated by the 2-ply lookahead */

/* ep_info layout (int32[BG_EP_INFO_INTS] per drained episode) */
#define BG_EP_INFO_INTS 12
/* [0] win_type (0 none/truncated, 1 regular, 2 gammon, 3 backgammon)  [1] winner (-1 none)  [2] env steps  [3] passes
 * [4],[5] close_out_counts P1,P2  [6],[7] prime_reward_counts P1,P2  [8] players-seen mask  [9] global game id
 * [10] episode serial of that game slot  [11] reserved */

/* per-experience meta byte: bit0 mover (== observation flag), bit1 next-observation flag, bit2 done,
 * bit3 close_out_reward, bit4 prime_reward */

/*
 * max_plies: reference MAX_TIMESTEPS = 300 (src/config/configuration.py:4); move_cap: max_legal_moves = 500
 * (src/environments/backgammon_env.py:35).  game_id_base offsets the Philox streams so that sharding games over GPUs
 * does not change any game's dice or sampled actions.  ring_*: capacity of the finished-episode ring (0 = default).
 * auto_reset != 0: a finished game immediately starts its next episode.
 */
int32_t bg_arena_create(bg_arena** out, int32_t device, int64_t n_games, int32_t H, int32_t max_plies, int32_t move_cap,
                        uint64_t seed, int64_t game_id_base, int64_t ring_experiences, int64_t ring_episodes,
                        int32_t auto_reset);
int32_t bg_arena_destroy(bg_arena* a);
/* Publish a new packed weight blob (device pointer) + its version and sampling temperature
 * (ParameterManager.set_parameters / get_temperature, src/multi/parameter_manager.py:79-111). Double buffered. */
int32_t bg_arena_set_weights(bg_arena* a, const float* packed, int64_t version, float temperature /* <=0: greedy */,
                             void* stream);
/* Parity mode: dice come from tape[n_games][L][2] (host or device pointer, copied) in order, including the doubles
 * rejected by the reset protocol; a game whose tape runs out stops.  NULL restores Philox dice. */
int32_t bg_arena_set_dice_tape(bg_arena* a, const uint8_t* tape, int64_t L, void* stream);
/* (Re)start every game: BackgammonEnv.reset (src/environments/backgammon_env.py:92-128); clears stats and the ring. */
int32_t bg_arena_reset(bg_arena* a, void* stream);
/* 2-ply policy parameters used by bg_arena_step(lookahead = 2) (src/multi/two_ply.py:44-90 and its commented-out integration
 * :153-193).  n_candidates = 4, top_k = 5, alpha = 1.0, beta = 0.9 is the reference's setting (decisions with fewer than
 * n_candidates legal moves fall back to the 1-ply policy); n_candidates = 0 scores EVERY legal afterstate (north_star's
 * expectimax with top_k = 1), which reads one counter back to the host per ply. */
int32_t bg_arena_set_lookahead(bg_arena* a, int32_t n_candidates, int32_t top_k, float alpha, float beta);
/* Advance every active game by n_plies env steps: movegen + fused eval (+ 2-ply rescoring when lookahead == 2) + select +
 * apply + record.  lookahead is 1 or 2.
 * forced_action: optional device int32[n_games]; entries >= 0 override the policy's choice (env.step(action)). */
int32_t bg_arena_step(bg_arena* a, int32_t n_plies, int32_t lookahead, const int32_t* forced_action, void* stream);
/*
 * Hand finished episodes to the learner (replaces ExperienceQueue.get, src/main.py:117): copies up to max_episodes whole
 * episodes / max_experiences records, oldest first, into caller buffers in CSR form and frees their ring space.
 * Record t of an episode: after_boards = board after the mover's move (== next_observation's board; the observation's
 * board is the previous record's after_board, or the initial board for t = 0), Experience fields state_value = v,
 * next_state_value = v_next, reward, done (meta).  n_moves/action/roll are optional traces (may be NULL).
 * out_n: device int64[2] = {episodes, experiences} written.
 */
int32_t bg_arena_drain_episodes(bg_arena* a, int64_t max_episodes, int64_t max_experiences, int8_t* after_boards /*[.,52]*/,
                                uint8_t* meta, float* reward, float* v, float* v_next, int16_t* n_moves, int16_t* action,
                                uint8_t* roll /*[.,2]*/, int64_t* ep_offsets /*[max_episodes+1]*/,
                                int32_t* ep_info /*[max_episodes,BG_EP_INFO_INTS]*/, int64_t* out_n /*[2]*/, void* stream);
/* out: int64[BG_ARENA_NSTATS], host or device pointer (cudaMemcpyAsync on `stream`). */
int32_t bg_arena_stats(bg_arena* a, int64_t* out, void* stream);
/* Snapshot of the live games (any pointer may be NULL): boards[n,52], players[n], rolls[n,2], game_state[n]. */
int32_t bg_arena_export_state(bg_arena* a, int8_t* boards, uint8_t* players, uint8_t* rolls, uint8_t* game_state,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * TD(0) learner: the consumer of the arena's episodes.
 * Replaces Trainer.__init__ / Trainer.update (src/agents/trainer.py:11-46, 48-166): for each episode IN ORDER one forward
 * pass over its T observations, targets r_t + gamma * V(x_{t+1}) (last target r_T), mse loss, backward,
 * clip_grad_norm_(grad_clip), one torch.optim.Adam step (betas 0.9/0.999, eps 1e-8).  The whole batch runs as one
 * persistent thread-block cluster with the optimiser state in shared memory; results agree with the reference trainer to
 * fp32 rounding (tests/golden/learner.npz: |dW| <= 1e-5 after 400 sequential steps).  16 hidden units per CTA: H <= 128 uses a
 * portable cluster (<= 8 CTAs), larger H a 10-16 CTA cluster (opt-in size); if the device refuses that, or with
 * BG_LEARNER_PATH=cuda-core in the environment, a CUDA-core kernel with 32 units per CTA is used for H > 128.
 * ---------------------------------------------------------------------------------------------- */
typedef struct bg_learner bg_learner;

#define BG_LEARNER_NMETRICS 6
/* per-episode metrics row (the quantities trainer.py:141-151 accumulates): [0] loss  [1] mean |TD error|
 * [2] gradient norm after clipping  [3] mean predicted value  [4] reward sum  [5] length T */
#define BG_LEARNER_MAX_T 320 /* longest episode accepted (reference MAX_TIMESTEPS = 300) */

/* reference defaults: lr = 1e-3, gamma = 0.99, grad_clip = 1.0 (src/config/configuration.py:17-21); grad_clip <= 0 disables clipping.
 * Parameters start at zero: call bg_learner_set_parameters before the first update. */
int32_t bg_learner_create(bg_learner** out, int32_t device, int32_t H, float lr, float gamma, float grad_clip);
int32_t bg_learner_destroy(bg_learner* l);
/* Load a packed weight blob (device pointer; Trainer.__init__'s load_state_dict, trainer.py:21-24).
 * reset_optimizer != 0 also zeroes the Adam moments and the step count (a fresh torch.optim.Adam). */
int32_t bg_learner_set_parameters(bg_learner* l, const float* packed, int32_t reset_optimizer, void* stream);
/* Copy the current packed weights out (device pointer): what Trainer hands to ParameterManager.set_parameters (trainer.py:166);
 * feed it to bg_arena_set_weights. */
int32_t bg_learner_get_parameters(bg_learner* l, float* packed_out, void* stream);
/* Adam moments in packed order and the step count (device pointers, any may be NULL). */
int32_t bg_learner_get_optimizer(bg_learner* l, float* exp_avg, float* exp_avg_sq, int64_t* step, void* stream);
/*
 * One Trainer.update over n_episodes episodes in CSR form (ep_offsets[n_episodes+1], device).
 *   ep_len      : optional device int32[n_episodes].  When given, episode e is the rows [ep_offsets[e], ep_offsets[e] + ep_len[e]) and only
 *                 ep_offsets[0..n_episodes) is read, so the episodes need not be contiguous (padded per-rank segments of an all-gather).
 *   records == 0: boards[N,52] / flags[N] are the OBSERVATIONS (board the decision was made on, player to move).
 *   records == 1: boards / flags are bg_arena_drain_episodes' after_boards / meta exactly as drained: experience t's
 *                 observation board is record t-1's after_board (the initial board for t = 0), its flag is meta bit 0.
 *   out_metrics : optional [n_episodes, BG_LEARNER_NMETRICS].
 *   out_status  : optional [1] device int32: BG_OK, or BG_ERR_CAPACITY if an episode was longer than BG_LEARNER_MAX_T
 *                 (such episodes, and empty ones, are skipped without an optimiser step).
 */
int32_t bg_learner_update(bg_learner* l, const int8_t* boards, const uint8_t* flags, const float* reward, const int64_t* ep_offsets,
                          const int32_t* ep_len /*or NULL*/, int64_t n_episodes, int32_t records, float* out_metrics, int32_t* out_status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BGARENA_H */
