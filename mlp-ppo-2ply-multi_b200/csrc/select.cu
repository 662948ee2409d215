// select.cu -- temperature-softmax sampling / greedy argmax over ragged value segments (sm_100a).
// Replaces softmax(V/T) + Categorical.sample (reference src/multi/worker.py:136-143) and torch.argmax
// (src/play/play_versus_ai.py:188-195).  One warp per item, warp-shuffle max / sum / prefix scan,
// counter-based Philox4x32-10 keyed by (seed, global item id, ctr).
#include "select.cuh"

namespace bg {

namespace {

__global__ void __launch_bounds__(256) k_select(SelectArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t i = warp; i < a.B; i += nwarps) {
    int n = a.counts[i];
    if (n > a.item_cap) n = a.item_cap;
    const int64_t off = a.offsets[i];
    int act = -1;
    if (n > 0 && off >= 0) {
      uint32_t r[4];
      Philox::gen(a.seed, (uint64_t)(a.item_id_base + i), a.ctr, r);
      const float u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
      act = warp_select(a.v + off, n, a.temperature, u, lane);
    }
    if (lane == 0) a.out_action[i] = act;
  }
}

}  // namespace

int32_t select_launch(const SelectArgs& a, cudaStream_t stream) {
  if (a.B <= 0) return BG_OK;
  int64_t want = (a.B + 7) / 8;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  k_select<<<grid, 256, 0, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_select launch");
  return BG_OK;
}

}  // namespace bg
