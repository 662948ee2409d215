// api.cu -- the extern "C" boundary of libbgarena.so (include/bgarena.h).  No torch / C++ types cross it.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "arena.cuh"
#include "eval.cuh"
#include "learner.cuh"
#include "movegen.cuh"
#include "select.cuh"
#include "two_ply.cuh"
#include "hostpath.cuh"

namespace bg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int32_t check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return BG_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return BG_ERR_CUDA;
}

static int32_t require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (libbgarena has no CPU fallback)");
    return BG_ERR_CUDA;
  }
  return BG_OK;
}

}  // namespace bg

using namespace bg;

namespace {
// per-device library-owned side stream (highest priority, so that the evaluator's CTAs are placed before the tail tiers' when both
// become runnable at the end of the bulk tier)
std::mutex g_fused_mu;
SideCtx g_fused[64];

// Callers hold g_fused_mu from here until their launches are enqueued: the context's events are recorded and waited on in pairs,
// and two host threads interleaving those pairs on one device would wait on each other's records.
int32_t fused_ctx(SideCtx** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) return check_cuda(e == cudaSuccess ? cudaErrorInvalidDevice : e, "cudaGetDevice");
  int32_t rc = side_ctx_create(&g_fused[dev], true);
  *out = &g_fused[dev];
  return rc;
}
}  // namespace

#define BG_REQUIRE(cond, msg)   \
  do {                          \
    if (!(cond)) {              \
      set_error("%s", msg);     \
      return BG_ERR_ARG;        \
    }                           \
  } while (0)

extern "C" {

int32_t bg_abi_version(void) { return BG_ABI_VERSION; }

const char* bg_last_error(void) { return g_err; }

int32_t bg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t bg_movegen_workspace_bytes(int64_t B) { return movegen_workspace_bytes(B < 0 ? 0 : B); }

int32_t bg_movegen(const int8_t* boards, const uint8_t* players, const uint8_t* rolls, int64_t B, int32_t item_cap,
                   int64_t pool_cap, int8_t* out_boards, uint8_t* out_submoves, int32_t* out_owner, uint8_t* out_flags,
                   int64_t* out_offsets, int32_t* out_count, int64_t* out_total, int32_t* out_status, void* workspace, int64_t workspace_bytes,
                   void* stream) {
  BG_REQUIRE(B >= 0, "bg_movegen: B < 0");
  BG_REQUIRE(B == 0 || (boards && players && rolls && out_offsets && out_count), "bg_movegen: null input/output pointer");
  BG_REQUIRE(pool_cap == 0 || out_boards, "bg_movegen: out_boards is null");
  BG_REQUIRE(workspace, "bg_movegen: workspace is null");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  MovegenArgs a{boards,    players,     rolls,     B,         item_cap,   pool_cap,  out_boards,      out_submoves,
                out_owner, out_flags, out_offsets, out_count, out_total, out_status, workspace, workspace_bytes, nullptr};
  return movegen_launch(a, (cudaStream_t)stream);
}

int32_t bg_encode(const int8_t* boards, const uint8_t* flag_player, int64_t N, float* out, void* stream) {
  BG_REQUIRE(N >= 0, "bg_encode: N < 0");
  BG_REQUIRE(N == 0 || (boards && flag_player && out), "bg_encode: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  return encode_launch(boards, flag_player, N, out, (cudaStream_t)stream);
}

int64_t bg_prepared_weights_bytes(int32_t H) { return prepared_weights_bytes(H); }

int32_t bg_prepare_weights(const float* packed, int32_t H, float* prepared, void* stream) {
  BG_REQUIRE(packed && prepared, "bg_prepare_weights: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  return prepare_weights_launch(packed, H, prepared, (cudaStream_t)stream);
}

int32_t bg_eval(const int8_t* boards, const uint8_t* flags, const int32_t* owner, const uint8_t* owner_players, int64_t N,
                const float* prepared, int32_t H, float* out_v, void* stream) {
  BG_REQUIRE(N >= 0, "bg_eval: N < 0");
  BG_REQUIRE(N == 0 || (boards && prepared && out_v), "bg_eval: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  EvalArgs a{boards, flags, owner, owner_players, N, nullptr, N, prepared, H, out_v};
  return eval_launch(a, (cudaStream_t)stream);
}

int32_t bg_eval_indirect(const int8_t* boards, const uint8_t* flags, const int32_t* owner, const uint8_t* owner_players,
                         const int64_t* N_dev, int64_t max_N, const float* prepared, int32_t H, float* out_v, void* stream) {
  BG_REQUIRE(max_N >= 0 && N_dev, "bg_eval_indirect: bad N");
  BG_REQUIRE(max_N == 0 || (boards && prepared && out_v), "bg_eval_indirect: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  EvalArgs a{boards, flags, owner, owner_players, 0, N_dev, max_N, prepared, H, out_v};
  return eval_launch(a, (cudaStream_t)stream);
}

int32_t bg_select(const float* v, const int64_t* offsets, const int32_t* counts, int32_t item_cap, int64_t B, float temperature,
                  uint64_t seed, uint64_t ctr, int64_t item_id_base, int32_t* out_action, void* stream) {
  BG_REQUIRE(B >= 0, "bg_select: B < 0");
  BG_REQUIRE(B == 0 || (v && offsets && counts && out_action), "bg_select: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  SelectArgs a{v, offsets, counts, item_cap, B, temperature, seed, ctr, item_id_base, out_action};
  return select_launch(a, (cudaStream_t)stream);
}

int32_t bg_eval_tc_status(void) { return eval_tc_status(); }
int32_t bg_eval_tc_tile_schedule(int32_t mode) { return eval_tc_tile_schedule(mode); }

int64_t bg_two_ply_workspace_bytes(int64_t N) { return two_ply_workspace_bytes(N); }
int32_t bg_two_ply_reply_sampling(int32_t cap, uint64_t seed) { return two_ply_reply_sampling(cap, seed); }

int32_t bg_two_ply(const int8_t* cand_boards, const uint8_t* mover, const float* S, int64_t N, const float* prepared, int32_t H,
                   int32_t top_k, float alpha, float beta, float* out_score, int64_t* out_replies, int32_t* out_status, void* workspace,
                   int64_t workspace_bytes, void* stream) {
  BG_REQUIRE(N >= 0, "bg_two_ply: N < 0");
  BG_REQUIRE(N == 0 || (cand_boards && mover && S && prepared && out_score && workspace), "bg_two_ply: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  TwoPlyArgs a{cand_boards, mover, S, N, prepared, H, top_k, alpha, beta, out_score, out_replies, out_status, workspace, workspace_bytes, nullptr, nullptr};
  std::lock_guard<std::mutex> lk(g_fused_mu);
  SideCtx* c = nullptr;
  if ((rc = fused_ctx(&c)) != BG_OK) return rc;
  a.side = c;
  return two_ply_launch(a, (cudaStream_t)stream);
}

/* ---- fused move generation + evaluation ---- */


int32_t bg_movegen_eval(const int8_t* boards, const uint8_t* players, const uint8_t* rolls, int64_t B, int32_t item_cap, int64_t pool_cap,
                        int8_t* out_boards, uint8_t* out_flags, int64_t* out_offsets, int32_t* out_count, int64_t* out_total /*[2]*/,
                        int32_t* out_status, void* workspace, int64_t workspace_bytes, const float* prepared, int32_t H, float* out_v,
                        void* stream) {
  BG_REQUIRE(B >= 0, "bg_movegen_eval: B < 0");
  BG_REQUIRE(B == 0 || (boards && players && rolls && out_offsets && out_count), "bg_movegen_eval: null input/output pointer");
  BG_REQUIRE(out_boards && out_flags && out_total && out_v && prepared && workspace, "bg_movegen_eval: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  std::lock_guard<std::mutex> lk(g_fused_mu);
  SideCtx* c = nullptr;
  if ((rc = fused_ctx(&c)) != BG_OK) return rc;
  MovegenArgs a{boards,  players,   rolls,       B,         item_cap,  pool_cap,   out_boards, nullptr,
                nullptr, out_flags, out_offsets, out_count, out_total, out_status, workspace,  workspace_bytes, nullptr};
  return movegen_eval_overlapped(a, out_total, prepared, H, out_v, c, (cudaStream_t)stream);
}

int32_t bg_movegen_all_rolls(const int8_t* boards, const uint8_t* players, int64_t P, int32_t item_cap, int64_t pool_cap, int8_t* out_boards,
                             uint8_t* out_submoves, int32_t* out_owner, uint8_t* out_flags, int64_t* out_offsets, int32_t* out_count,
                             int64_t* out_total, int32_t* out_status, void* workspace, int64_t workspace_bytes, void* stream) {
  BG_REQUIRE(P >= 0, "bg_movegen_all_rolls: P < 0");
  BG_REQUIRE(P == 0 || (boards && players && out_offsets && out_count), "bg_movegen_all_rolls: null input/output pointer");
  BG_REQUIRE(pool_cap == 0 || out_boards, "bg_movegen_all_rolls: out_boards is null");
  BG_REQUIRE(workspace, "bg_movegen_all_rolls: workspace is null");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  MovegenArgs a{boards,    players,   nullptr,     P,         item_cap,  pool_cap,   out_boards, out_submoves,
                out_owner, out_flags, out_offsets, out_count, out_total, out_status, workspace,  workspace_bytes, nullptr};
  a.all_rolls = 1;
  return movegen_launch(a, (cudaStream_t)stream);
}

int32_t bg_movegen_eval_all_rolls(const int8_t* boards, const uint8_t* players, int64_t P, int32_t item_cap, int64_t pool_cap, int8_t* out_boards,
                                  uint8_t* out_flags, int64_t* out_offsets, int32_t* out_count, int64_t* out_total /*[2]*/, int32_t* out_status,
                                  void* workspace, int64_t workspace_bytes, const float* prepared, int32_t H, float* out_v, void* stream) {
  BG_REQUIRE(P >= 0, "bg_movegen_eval_all_rolls: P < 0");
  BG_REQUIRE(P == 0 || (boards && players && out_offsets && out_count), "bg_movegen_eval_all_rolls: null input/output pointer");
  BG_REQUIRE(out_boards && out_flags && out_total && out_v && prepared && workspace, "bg_movegen_eval_all_rolls: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  std::lock_guard<std::mutex> lk(g_fused_mu);
  SideCtx* c = nullptr;
  if ((rc = fused_ctx(&c)) != BG_OK) return rc;
  MovegenArgs a{boards,  players,   nullptr,     P,         item_cap,  pool_cap,   out_boards, nullptr,
                nullptr, out_flags, out_offsets, out_count, out_total, out_status, workspace,  workspace_bytes, nullptr};
  a.all_rolls = 1;
  return movegen_eval_overlapped(a, out_total, prepared, H, out_v, c, (cudaStream_t)stream);
}

/* ---- compact (code) pools ---- */

int32_t bg_movegen_all_rolls_compact(const int8_t* boards, const uint8_t* players, int64_t P, int32_t item_cap, int64_t pool_cap, uint64_t* out_codes,
                                     int64_t* out_offsets, int32_t* out_count, int64_t* out_total, int32_t* out_status, void* workspace,
                                     int64_t workspace_bytes, void* stream) {
  BG_REQUIRE(P >= 0, "bg_movegen_all_rolls_compact: P < 0");
  BG_REQUIRE(P == 0 || (boards && players && out_offsets && out_count), "bg_movegen_all_rolls_compact: null input/output pointer");
  BG_REQUIRE(out_codes && workspace, "bg_movegen_all_rolls_compact: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  MovegenArgs a{boards,  players, nullptr,     P,         item_cap,  pool_cap,   nullptr,   nullptr,
                nullptr, nullptr, out_offsets, out_count, out_total, out_status, workspace, workspace_bytes, nullptr};
  a.all_rolls = 1;
  a.out_codes = reinterpret_cast<uint2*>(out_codes);
  return movegen_launch(a, (cudaStream_t)stream);
}

int32_t bg_eval_codes(const int8_t* boards, const uint8_t* players, const uint64_t* codes, int64_t N, const int64_t* N_dev, int64_t max_N,
                      const float* prepared, int32_t H, float* out_v, void* stream) {
  BG_REQUIRE(boards && players && codes && prepared && out_v, "bg_eval_codes: null pointer");
  BG_REQUIRE(N_dev ? max_N >= 0 : N >= 0, "bg_eval_codes: bad N");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  EvalArgs a{boards, players, nullptr, nullptr, N_dev ? 0 : N, N_dev, N_dev ? max_N : N, prepared, H, out_v};
  a.codes = reinterpret_cast<const uint2*>(codes);
  return eval_launch(a, (cudaStream_t)stream);
}

int32_t bg_movegen_eval_all_rolls_compact(const int8_t* boards, const uint8_t* players, int64_t P, int32_t item_cap, int64_t pool_cap,
                                          uint64_t* out_codes, int64_t* out_offsets, int32_t* out_count, int64_t* out_total /*[2]*/,
                                          int32_t* out_status, void* workspace, int64_t workspace_bytes, const float* prepared, int32_t H,
                                          float* out_v, void* stream) {
  BG_REQUIRE(P >= 0, "bg_movegen_eval_all_rolls_compact: P < 0");
  BG_REQUIRE(P == 0 || (boards && players && out_offsets && out_count), "bg_movegen_eval_all_rolls_compact: null input/output pointer");
  BG_REQUIRE(out_codes && out_total && out_v && prepared && workspace, "bg_movegen_eval_all_rolls_compact: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  std::lock_guard<std::mutex> lk(g_fused_mu);
  SideCtx* c = nullptr;
  if ((rc = fused_ctx(&c)) != BG_OK) return rc;
  MovegenArgs a{boards,  players, nullptr,     P,         item_cap,  pool_cap,   nullptr,   nullptr,
                nullptr, nullptr, out_offsets, out_count, out_total, out_status, workspace, workspace_bytes, nullptr};
  a.all_rolls = 1;
  a.out_codes = reinterpret_cast<uint2*>(out_codes);
  return movegen_eval_overlapped(a, out_total, prepared, H, out_v, c, (cudaStream_t)stream);
}

int32_t bg_afterstates_from_codes(const int8_t* boards, const uint8_t* players, const uint64_t* codes, const int64_t* rows, int64_t n,
                                  int8_t* out_boards, void* stream) {
  BG_REQUIRE(n >= 0, "bg_afterstates_from_codes: n < 0");
  BG_REQUIRE(n == 0 || (boards && players && codes && out_boards), "bg_afterstates_from_codes: null pointer");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  return materialize_launch(boards, players, reinterpret_cast<const uint2*>(codes), rows, n, out_boards, (cudaStream_t)stream);
}

/* ---- host-resident batches ---- */

int32_t bg_hostpipe_create(bg_hostpipe** out, int32_t device, int32_t H, int64_t chunk_units, int32_t all_rolls, int32_t item_cap,
                           int32_t rows_per_item, int32_t n_streams) {
  BG_REQUIRE(out, "bg_hostpipe_create: out is null");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  HostPipe* p = nullptr;
  rc = hostpipe_create(&p, device, H, chunk_units, all_rolls, item_cap, rows_per_item, n_streams);
  *out = reinterpret_cast<bg_hostpipe*>(p);
  return rc;
}

int32_t bg_hostpipe_destroy(bg_hostpipe* p) { return hostpipe_destroy(reinterpret_cast<HostPipe*>(p)); }

int32_t bg_hostpipe_run(bg_hostpipe* p, const int8_t* h_boards, const uint8_t* h_players, const uint8_t* h_rolls, int64_t n_units,
                        const float* prepared, float temperature, uint64_t seed, int32_t* h_actions, int32_t* h_counts, void* stream) {
  BG_REQUIRE(p, "bg_hostpipe_run: null pipe");
  return hostpipe_run(reinterpret_cast<HostPipe*>(p), h_boards, h_players, h_rolls, n_units, prepared, temperature, seed, h_actions, h_counts,
                      (cudaStream_t)stream);
}

int32_t bg_hostpipe_status(bg_hostpipe* p, int32_t* out_status) {
  BG_REQUIRE(p, "bg_hostpipe_status: null pipe");
  return hostpipe_status(reinterpret_cast<HostPipe*>(p), out_status);
}

/* ---- arena ---- */

int32_t bg_arena_create(bg_arena** out, int32_t device, int64_t n_games, int32_t H, int32_t max_plies, int32_t move_cap, uint64_t seed,
                        int64_t game_id_base, int64_t ring_experiences, int64_t ring_episodes, int32_t auto_reset) {
  BG_REQUIRE(out, "bg_arena_create: out is null");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  Arena* A = nullptr;
  DeviceGuard guard(device);  // the caller's current device is restored on return
  rc = arena_create(&A, device, n_games, H, max_plies, move_cap, seed, game_id_base, ring_experiences, ring_episodes, auto_reset);
  *out = reinterpret_cast<bg_arena*>(A);
  return rc;
}

int32_t bg_arena_destroy(bg_arena* a) {
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_destroy(reinterpret_cast<Arena*>(a));
}

int32_t bg_arena_set_weights(bg_arena* a, const float* packed, int64_t version, float temperature, void* stream) {
  BG_REQUIRE(a && packed, "bg_arena_set_weights: null pointer");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_set_weights(reinterpret_cast<Arena*>(a), packed, version, temperature, (cudaStream_t)stream);
}

int32_t bg_arena_set_dice_tape(bg_arena* a, const uint8_t* tape, int64_t L, void* stream) {
  BG_REQUIRE(a, "bg_arena_set_dice_tape: null arena");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_set_dice_tape(reinterpret_cast<Arena*>(a), tape, L, (cudaStream_t)stream);
}

int32_t bg_arena_reset(bg_arena* a, void* stream) {
  BG_REQUIRE(a, "bg_arena_reset: null arena");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_reset(reinterpret_cast<Arena*>(a), (cudaStream_t)stream);
}

int32_t bg_arena_set_lookahead(bg_arena* a, int32_t n_candidates, int32_t top_k, float alpha, float beta) {
  BG_REQUIRE(a, "bg_arena_set_lookahead: null arena");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_set_lookahead(reinterpret_cast<Arena*>(a), n_candidates, top_k, alpha, beta);
}

int32_t bg_arena_step(bg_arena* a, int32_t n_plies, int32_t lookahead, const int32_t* forced_action, void* stream) {
  BG_REQUIRE(a && n_plies >= 0, "bg_arena_step: bad arguments");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_step(reinterpret_cast<Arena*>(a), n_plies, lookahead, forced_action, (cudaStream_t)stream);
}

int32_t bg_arena_drain_episodes(bg_arena* a, int64_t max_episodes, int64_t max_experiences, int8_t* after_boards, uint8_t* meta,
                                float* reward, float* v, float* v_next, int16_t* n_moves, int16_t* action, uint8_t* roll,
                                int64_t* ep_offsets, int32_t* ep_info, int64_t* out_n, void* stream) {
  BG_REQUIRE(a, "bg_arena_drain_episodes: null arena");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_drain(reinterpret_cast<Arena*>(a), max_episodes, max_experiences, after_boards, meta, reward, v, v_next, n_moves, action,
                     roll, ep_offsets, ep_info, out_n, (cudaStream_t)stream);
}

int32_t bg_arena_stats(bg_arena* a, int64_t* out, void* stream) {
  BG_REQUIRE(a && out, "bg_arena_stats: null pointer");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_stats(reinterpret_cast<Arena*>(a), out, (cudaStream_t)stream);
}

int32_t bg_arena_export_state(bg_arena* a, int8_t* boards, uint8_t* players, uint8_t* rolls, uint8_t* game_state, void* stream) {
  BG_REQUIRE(a, "bg_arena_export_state: null arena");
  DeviceGuard guard(arena_device(reinterpret_cast<Arena*>(a)));
  return arena_export_state(reinterpret_cast<Arena*>(a), boards, players, rolls, game_state, (cudaStream_t)stream);
}

/* ---- learner ---- */

int32_t bg_learner_create(bg_learner** out, int32_t device, int32_t H, float lr, float gamma, float grad_clip) {
  BG_REQUIRE(out, "bg_learner_create: out is null");
  BG_REQUIRE(lr > 0.f && gamma >= 0.f, "bg_learner_create: bad hyper-parameters");
  int32_t rc = require_device();
  if (rc != BG_OK) return rc;
  Learner* L = nullptr;
  DeviceGuard guard(device);
  rc = learner_create(&L, device, H, lr, gamma, grad_clip);
  *out = reinterpret_cast<bg_learner*>(L);
  return rc;
}

int32_t bg_learner_destroy(bg_learner* l) {
  DeviceGuard guard(learner_device(reinterpret_cast<Learner*>(l)));
  return learner_destroy(reinterpret_cast<Learner*>(l));
}

int32_t bg_learner_set_parameters(bg_learner* l, const float* packed, int32_t reset_optimizer, void* stream) {
  BG_REQUIRE(l && packed, "bg_learner_set_parameters: null pointer");
  DeviceGuard guard(learner_device(reinterpret_cast<Learner*>(l)));
  return learner_set_parameters(reinterpret_cast<Learner*>(l), packed, reset_optimizer, (cudaStream_t)stream);
}

int32_t bg_learner_get_parameters(bg_learner* l, float* packed_out, void* stream) {
  BG_REQUIRE(l && packed_out, "bg_learner_get_parameters: null pointer");
  DeviceGuard guard(learner_device(reinterpret_cast<Learner*>(l)));
  return learner_get_parameters(reinterpret_cast<Learner*>(l), packed_out, (cudaStream_t)stream);
}

int32_t bg_learner_get_optimizer(bg_learner* l, float* exp_avg, float* exp_avg_sq, int64_t* step, void* stream) {
  BG_REQUIRE(l, "bg_learner_get_optimizer: null learner");
  DeviceGuard guard(learner_device(reinterpret_cast<Learner*>(l)));
  return learner_get_optimizer(reinterpret_cast<Learner*>(l), exp_avg, exp_avg_sq, step, (cudaStream_t)stream);
}

int32_t bg_learner_update(bg_learner* l, const int8_t* boards, const uint8_t* flags, const float* reward, const int64_t* ep_offsets,
                          const int32_t* ep_len, int64_t n_episodes, int32_t records, float* out_metrics, int32_t* out_status, void* stream) {
  BG_REQUIRE(l && n_episodes >= 0, "bg_learner_update: bad arguments");
  BG_REQUIRE(n_episodes == 0 || (boards && flags && reward && ep_offsets), "bg_learner_update: null pointer");
  DeviceGuard guard(learner_device(reinterpret_cast<Learner*>(l)));
  return learner_update(reinterpret_cast<Learner*>(l), boards, flags, reward, ep_offsets, ep_len, n_episodes, records, out_metrics, out_status,
                        (cudaStream_t)stream);
}

}  // extern "C"
