// select.cuh -- host-side launch interface of the action selector (see select.cu)
#pragma once
#include "bg_common.cuh"

namespace bg {

struct SelectArgs {
  const float* v;
  const int64_t* offsets;
  const int32_t* counts;
  int32_t item_cap;
  int64_t B;
  float temperature;
  uint64_t seed, ctr;
  int64_t item_id_base;
  int32_t* out_action;
};

int32_t select_launch(const SelectArgs& a, cudaStream_t stream);

// warp-cooperative selection over v[0..n): temperature <= 0 -> first-index argmax, else softmax(v/T) sample with u in [0,1)
__device__ __forceinline__ int warp_select(const float* __restrict__ v, int n, float temperature, float u, int lane) {
  if (n <= 0) return -1;
  if (temperature <= 0.f) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      const float x = v[i];
      if (x > best || bi == 0x7fffffff) {  // strict > keeps the lowest index within a lane
        best = x;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(BG_FULL, best, o);
      const int oi = __shfl_xor_sync(BG_FULL, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) {
        best = ob;
        bi = oi;
      }
    }
    return bi;
  }
  const float invT = 1.0f / temperature;
  float m = -INFINITY;
  for (int i = lane; i < n; i += 32) m = fmaxf(m, v[i] * invT);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(BG_FULL, m, o));
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += __expf(v[i] * invT - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(BG_FULL, s, o);
  const float target = u * s;
  float running = 0.f;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const float e = i < n ? __expf(v[i] * invT - m) : 0.f;
    float inc = e;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(BG_FULL, inc, o);
      if (lane >= o) inc += t;
    }
    const uint32_t hit = __ballot_sync(BG_FULL, i < n && target < running + inc);
    if (hit) return base + __ffs(hit) - 1;
    running += __shfl_sync(BG_FULL, inc, 31);
  }
  return n - 1;
}

}  // namespace bg
