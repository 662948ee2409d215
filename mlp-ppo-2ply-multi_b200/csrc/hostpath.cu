// hostpath.cu -- the per-decision hot path for HOST-resident batches, behind the C ABI (bg_hostpipe_*).
//
// What a CPU-side caller of get_all_possible_moves + generate_all_board_features + policy_network.forward + argmax / sample
// (reference src/multi/worker.py:101-143) does when its positions live in host memory: (boards, players[, rolls]) in, one action and
// the legal-move count per item out.  The batch is cut into chunks that rotate over n_streams library-owned streams, so chunk k+1's
// host->device copy and chunk k-1's device->host copy overlap chunk k's kernels (move generation + fused evaluation + selection);
// every device buffer is allocated once at create time.  Each chunk's status is folded into one device word (nothing is overwritten),
// read by bg_hostpipe_status.
#include <new>
#include <vector>

#include "eval.cuh"
#include "movegen.cuh"
#include "select.cuh"

namespace bg {

namespace {

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  SideCtx side;
  int8_t* boards = nullptr;
  uint8_t* players = nullptr;
  uint8_t* rolls = nullptr;
  int8_t* pool = nullptr;   // boards (per-item batches)
  uint2* codes = nullptr;   // compact pool (position-major batches): 8 bytes per afterstate
  uint8_t* flags = nullptr;
  float* v = nullptr;
  int64_t* offsets = nullptr;
  int32_t* counts = nullptr;
  int32_t* actions = nullptr;
  int64_t* total = nullptr;  // [2]
  int32_t* status = nullptr;
  void* ws = nullptr;
};

__global__ void k_fold_status(const int32_t* __restrict__ chunk_status, int32_t* __restrict__ acc) {
  if (*chunk_status != 0) atomicMin(acc, *chunk_status);
}

}  // namespace

struct HostPipe {
  int device = 0;
  int32_t H = 0, item_cap = 0, all_rolls = 0;
  int64_t chunk_units = 0;  // positions (all_rolls) or items per chunk
  int64_t chunk_items = 0, pool_cap = 0, ws_bytes = 0;
  cudaEvent_t fork = nullptr;
  int32_t* status_acc = nullptr;
  std::vector<Slot> slots;
};

#define HP_TRY(expr, what)                               \
  do {                                                   \
    cudaError_t e__ = (expr);                            \
    if (e__ != cudaSuccess) return check_cuda(e__, what); \
  } while (0)

int32_t hostpipe_destroy(HostPipe* p) {
  if (!p) return BG_OK;
  DeviceGuard g(p->device);
  for (Slot& s : p->slots) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    side_ctx_destroy(&s.side);
    cudaFree(s.boards);
    cudaFree(s.players);
    cudaFree(s.rolls);
    cudaFree(s.pool);
    cudaFree(s.codes);
    cudaFree(s.flags);
    cudaFree(s.v);
    cudaFree(s.offsets);
    cudaFree(s.counts);
    cudaFree(s.actions);
    cudaFree(s.total);
    cudaFree(s.status);
    cudaFree(s.ws);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (p->fork) cudaEventDestroy(p->fork);
  cudaFree(p->status_acc);
  delete p;
  return BG_OK;
}

int32_t hostpipe_create(HostPipe** out, int32_t device, int32_t H, int64_t chunk_units, int32_t all_rolls, int32_t item_cap, int32_t rows_per_item,
                        int32_t n_streams) {
  *out = nullptr;
  if (chunk_units <= 0 || item_cap <= 0 || rows_per_item <= 0 || n_streams < 1 || n_streams > 8 || H < 32 || H > 256 || (H % 32)) {
    set_error("bg_hostpipe_create: bad arguments");
    return BG_ERR_ARG;
  }
  HostPipe* p = new (std::nothrow) HostPipe();
  if (!p) {
    set_error("bg_hostpipe_create: out of host memory");
    return BG_ERR_ARG;
  }
  p->device = device;
  DeviceGuard g(device);
  p->H = H;
  p->item_cap = item_cap;
  p->all_rolls = all_rolls ? 1 : 0;
  p->chunk_units = chunk_units;
  p->chunk_items = all_rolls ? chunk_units * 21 : chunk_units;
  p->pool_cap = p->chunk_items * rows_per_item + (1 << 18);
  p->ws_bytes = movegen_workspace_bytes(p->chunk_items);
  p->slots.resize(n_streams);
  cudaError_t e = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&p->status_acc, 4);
  if (e == cudaSuccess) e = cudaMemset(p->status_acc, 0, 4);
  for (Slot& s : p->slots) {
    if (e != cudaSuccess) break;
    e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    if (e == cudaSuccess && side_ctx_create(&s.side, true) != BG_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMalloc(&s.boards, (size_t)chunk_units * 52);
    if (e == cudaSuccess) e = cudaMalloc(&s.players, (size_t)chunk_units);
    if (e == cudaSuccess && !all_rolls) e = cudaMalloc(&s.rolls, (size_t)chunk_units * 2);
    // position-major batches only need actions and counts back: compact pool, the evaluator rebuilds the afterstates on chip
    if (e == cudaSuccess && all_rolls) e = cudaMalloc(&s.codes, (size_t)p->pool_cap * 8);
    if (e == cudaSuccess && !all_rolls) e = cudaMalloc(&s.pool, (size_t)p->pool_cap * 52);
    if (e == cudaSuccess && !all_rolls) e = cudaMalloc(&s.flags, (size_t)p->pool_cap);
    if (e == cudaSuccess) e = cudaMalloc(&s.v, (size_t)p->pool_cap * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.offsets, (size_t)p->chunk_items * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s.counts, (size_t)p->chunk_items * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.actions, (size_t)p->chunk_items * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.total, 16);
    if (e == cudaSuccess) e = cudaMalloc(&s.status, 4);
    if (e == cudaSuccess) e = cudaMalloc(&s.ws, (size_t)p->ws_bytes);
  }
  if (e != cudaSuccess) {
    int32_t rc = check_cuda(e, "bg_hostpipe_create");
    hostpipe_destroy(p);
    return rc;
  }
  *out = p;
  return BG_OK;
}

int32_t hostpipe_run(HostPipe* p, const int8_t* h_boards, const uint8_t* h_players, const uint8_t* h_rolls, int64_t n_units, const float* prepared,
                     float temperature, uint64_t seed, int32_t* h_actions, int32_t* h_counts, cudaStream_t stream) {
  if (n_units < 0 || (n_units > 0 && (!h_boards || !h_players || !h_actions || !h_counts || !prepared)) || (!p->all_rolls && n_units > 0 && !h_rolls)) {
    set_error("bg_hostpipe_run: bad arguments");
    return BG_ERR_ARG;
  }
  DeviceGuard g(p->device);
  HP_TRY(cudaEventRecord(p->fork, stream), "record fork");
  for (Slot& s : p->slots) HP_TRY(cudaStreamWaitEvent(s.stream, p->fork, 0), "fork");
  const int per_unit = p->all_rolls ? 21 : 1;
  int k = 0;
  for (int64_t lo = 0; lo < n_units; lo += p->chunk_units, ++k) {
    const int64_t n = n_units - lo < p->chunk_units ? n_units - lo : p->chunk_units;
    Slot& s = p->slots[k % p->slots.size()];
    HP_TRY(cudaMemcpyAsync(s.boards, h_boards + lo * 52, (size_t)n * 52, cudaMemcpyHostToDevice, s.stream), "H2D boards");
    HP_TRY(cudaMemcpyAsync(s.players, h_players + lo, (size_t)n, cudaMemcpyHostToDevice, s.stream), "H2D players");
    if (!p->all_rolls) HP_TRY(cudaMemcpyAsync(s.rolls, h_rolls + lo * 2, (size_t)n * 2, cudaMemcpyHostToDevice, s.stream), "H2D rolls");
    MovegenArgs m{};
    m.boards = s.boards;
    m.players = s.players;
    m.rolls = s.rolls;
    m.B = n;
    m.all_rolls = p->all_rolls;
    m.item_cap = p->item_cap;
    m.pool_cap = p->pool_cap;
    m.out_boards = s.pool;
    m.out_flags = s.flags;
    m.out_codes = s.codes;
    m.out_offsets = s.offsets;
    m.out_count = s.counts;
    m.out_status = s.status;
    m.workspace = s.ws;
    m.workspace_bytes = p->ws_bytes;
    int32_t rc = movegen_eval_overlapped(m, s.total, prepared, p->H, s.v, &s.side, s.stream);
    if (rc != BG_OK) return rc;
    k_fold_status<<<1, 1, 0, s.stream>>>(s.status, p->status_acc);
    const int64_t items = n * per_unit;
    SelectArgs sel{s.v, s.offsets, s.counts, p->item_cap, items, temperature, seed, 0, lo * per_unit, s.actions};
    if ((rc = select_launch(sel, s.stream)) != BG_OK) return rc;
    HP_TRY(cudaMemcpyAsync(h_actions + lo * per_unit, s.actions, (size_t)items * 4, cudaMemcpyDeviceToHost, s.stream), "D2H actions");
    HP_TRY(cudaMemcpyAsync(h_counts + lo * per_unit, s.counts, (size_t)items * 4, cudaMemcpyDeviceToHost, s.stream), "D2H counts");
  }
  for (Slot& s : p->slots) {
    HP_TRY(cudaEventRecord(s.done, s.stream), "record join");
    HP_TRY(cudaStreamWaitEvent(stream, s.done, 0), "join");
  }
  return BG_OK;
}

int32_t hostpipe_status(HostPipe* p, int32_t* out_status) {
  DeviceGuard g(p->device);
  for (Slot& s : p->slots) HP_TRY(cudaStreamSynchronize(s.stream), "bg_hostpipe_status: synchronize");
  int32_t st = 0;
  HP_TRY(cudaMemcpy(&st, p->status_acc, 4, cudaMemcpyDeviceToHost), "bg_hostpipe_status: read");
  HP_TRY(cudaMemset(p->status_acc, 0, 4), "bg_hostpipe_status: reset");
  if (out_status) *out_status = st;
  return BG_OK;
}

}  // namespace bg
