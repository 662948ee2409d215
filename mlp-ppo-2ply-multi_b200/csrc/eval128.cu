// eval128.cu -- the H = 128 production evaluator: fused 198-feature encode + sigmoid-MLP value, two boards per warp.
//
// Same maths as eval.cu (reference src/backgammon/board/immutable_board.py:86-128 + src/agents/policy_network.py:53-70),
// restructured after the r01 ncu profiles (issue-slot bound at ~380 instructions / board, shared-memory pipe at 79 %):
//   * a board is evaluated by a HALF-warp (16 lanes x 8 hidden units), so every warp instruction serves two boards;
//     a row read is two 16-byte loads per lane at [4*gl, 4*gl+4) and [64+4*gl, ...): each quarter-warp touches one
//     contiguous 128-byte span -> conflict-free, 4 wavefronts per 512-byte row exactly as in the one-board layout;
//   * the per-point table is cumulative up to SIX checkers (T[c] = r0 + r1 + r2 + (c-3)/2 * r3), so a point costs one row
//     for c <= 6 (the old 3-deep table paid a second, shuffle-fed row for every stack of 4+, ~2 per board);
//   * the flag row is folded into two bias rows (b1 + W[196], b1 + W[197]); sigmoid uses ex2.approx / rcp.approx directly.
// One persistent CTA of 1024 threads per SM keeps the 176 KB table in shared memory.
#include "eval.cuh"

namespace bg {

namespace {

constexpr int H = 128, K = 6, RP = K + 1, MISC = 48 * RP;  // MISC rows: bar0', off0, bar1', off1, bias(flag 0), bias(flag 1), w2
constexpr int TABLE_ROWS = MISC + 7;
constexpr int THREADS = 1024, WARPS = THREADS / 32, LIST_WORDS = 56, ELIST_WORDS = 8;
constexpr int NUM_SMS = 148;

__constant__ float c_off15b[16];

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// table for this kernel, appended to the generic one by bg_prepare_weights (H == 128 only)
__global__ void k_prepare128(const float* __restrict__ packed, float* __restrict__ t) {
  const int total = TABLE_ROWS * H + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i == TABLE_ROWS * H) {
      t[i] = packed[200 * H];  // b2
      continue;
    }
    const int row = i / H, h = i - row * H;
    const float b1 = packed[198 * H + h];
    float v;
    if (row < MISC) {
      const int sp = row / RP, k = row - sp * RP, f0 = sp * 4;  // sp = side * 24 + point
      const float r0 = packed[(f0 + 0) * H + h], r1 = packed[(f0 + 1) * H + h], r2 = packed[(f0 + 2) * H + h], r3 = packed[(f0 + 3) * H + h];
      const float cum3 = (r0 + r1) + r2;
      if (k == 0)
        v = r0;
      else if (k == 1)
        v = r0 + r1;
      else if (k == 2)
        v = cum3;
      else if (k < K)
        v = fmaf((float)(k - 2) * 0.5f, r3, cum3);  // c = k + 1 checkers: + (c - 3) / 2 * r3
      else
        v = 0.5f * r3;  // per extra checker above K
    } else {
      switch (row - MISC) {
        case 0: v = 0.5f * packed[192 * H + h]; break;
        case 1: v = packed[193 * H + h]; break;
        case 2: v = 0.5f * packed[194 * H + h]; break;
        case 3: v = packed[195 * H + h]; break;
        case 4: v = b1 + packed[196 * H + h]; break;
        case 5: v = b1 + packed[197 * H + h]; break;
        default: v = packed[199 * H + h]; break;  // w2
      }
    }
    t[i] = v;
  }
}

struct Acc8 {
  float a[8];
};

__device__ __forceinline__ void add2(Acc8& z, const float* row, int gl) {
  const float4 x = *reinterpret_cast<const float4*>(row + gl * 4);
  const float4 y = *reinterpret_cast<const float4*>(row + 64 + gl * 4);
  z.a[0] += x.x;
  z.a[1] += x.y;
  z.a[2] += x.z;
  z.a[3] += x.w;
  z.a[4] += y.x;
  z.a[5] += y.y;
  z.a[6] += y.z;
  z.a[7] += y.w;
}
__device__ __forceinline__ void fma2(Acc8& z, float s, const float* row, int gl) {
  const float4 x = *reinterpret_cast<const float4*>(row + gl * 4);
  const float4 y = *reinterpret_cast<const float4*>(row + 64 + gl * 4);
  z.a[0] = fmaf(s, x.x, z.a[0]);
  z.a[1] = fmaf(s, x.y, z.a[1]);
  z.a[2] = fmaf(s, x.z, z.a[2]);
  z.a[3] = fmaf(s, x.w, z.a[3]);
  z.a[4] = fmaf(s, y.x, z.a[4]);
  z.a[5] = fmaf(s, y.y, z.a[5]);
  z.a[6] = fmaf(s, y.z, z.a[6]);
  z.a[7] = fmaf(s, y.w, z.a[7]);
}

__global__ void __launch_bounds__(THREADS, 1)
    k_eval128(const int8_t* __restrict__ boards, const uint8_t* __restrict__ flags, const int32_t* __restrict__ owner,
              const uint8_t* __restrict__ owner_players, int64_t N_host, const int64_t* __restrict__ N_dev, int64_t max_N,
              const float* __restrict__ t128, float* __restrict__ out_v, const int64_t* __restrict__ start_dev) {
  extern __shared__ __align__(16) float sT[];
  {
    const float4* src = reinterpret_cast<const float4*>(t128);
    float4* dst = reinterpret_cast<float4*>(sT);
    for (int i = threadIdx.x; i < TABLE_ROWS * H / 4; i += THREADS) dst[i] = src[i];
  }
  uint32_t* lists = reinterpret_cast<uint32_t*>(sT + TABLE_ROWS * H);
  __syncthreads();
  const float b2 = t128[TABLE_ROWS * H];
  int64_t N = N_dev ? *N_dev : N_host;
  if (N > max_N) N = max_N;
  {  // optional device-side start row: shift the row-indexed arrays once, everything below is unchanged
    int64_t start = start_dev ? *start_dev : 0;
    if (start > N) start = N;
    if (start < 0) start = 0;
    boards += start * BG_BOARD_BYTES;
    if (flags) flags += start;
    if (owner) owner += start;
    out_v += start;
    N -= start;
  }
  const int lane = threadIdx.x & 31, gl = lane & 15, grp = lane >> 4, gb = grp << 4;
  const int wib = threadIdx.x >> 5;
  uint32_t* list = lists + (wib * 2 + grp) * (LIST_WORDS + ELIST_WORDS);
  uint32_t* elist = list + LIST_WORDS;
  const int64_t gid = ((int64_t)blockIdx.x * WARPS + wib) * 2 + grp;
  const int64_t ngroups = (int64_t)gridDim.x * WARPS * 2;
  const uint32_t* b32 = reinterpret_cast<const uint32_t*>(boards);
  float w2r[8];
  {
    const float* w2 = sT + (MISC + 6) * H;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      w2r[q] = w2[gl * 4 + q];
      w2r[4 + q] = w2[64 + gl * 4 + q];
    }
  }
  // group lanes 0..12 fetch the board words, lane 13 the flag; the next board is prefetched while this one is evaluated
  auto fetch = [&](int64_t j) -> uint32_t {
    if (j >= N) return 0u;
    if (gl < 13) return __ldg(b32 + j * 13 + gl);
    if (gl == 13) return flags ? (uint32_t)flags[j] : (uint32_t)owner_players[owner[j]];
    return 0u;
  };
  uint32_t nxt = fetch(gid);
  const uint32_t ltg = (1u << gl) - 1u;
  for (int64_t i = gid; __any_sync(BG_FULL, i < N); i += ngroups) {
    const bool valid = i < N;
    const uint32_t myw = nxt;
    nxt = fetch(i + ngroups);
    // ---- checker counts: lane gl owns points gl and (gl < 8) 16 + gl of both sides ----------------------------------
    const uint32_t flag = __shfl_sync(BG_FULL, myw, gb + 13) & 1u;
    const uint32_t w12 = __shfl_sync(BG_FULL, myw, gb + 12);
    const uint32_t wa0 = __shfl_sync(BG_FULL, myw, gb + (gl >> 2));
    const uint32_t wb0 = __shfl_sync(BG_FULL, myw, gb + 4 + ((gl & 7) >> 2));
    const uint32_t wa1 = __shfl_sync(BG_FULL, myw, gb + 6 + (gl >> 2));
    const uint32_t wb1 = __shfl_sync(BG_FULL, myw, gb + 10 + ((gl & 7) >> 2));
    const int sh = (gl & 3) * 8;
    const uint32_t c0a = (wa0 >> sh) & 0xffu, c1a = (wa1 >> sh) & 0xffu;
    const uint32_t c0b = gl < 8 ? (wb0 >> sh) & 0xffu : 0u, c1b = gl < 8 ? (wb1 >> sh) & 0xffu : 0u;
    const uint32_t o0 = ((__ballot_sync(BG_FULL, c0a > 0) >> gb) & 0xffffu) | (((__ballot_sync(BG_FULL, c0b > 0) >> gb) & 0xffu) << 16);
    const uint32_t o1 = ((__ballot_sync(BG_FULL, c1a > 0) >> gb) & 0xffffu) | (((__ballot_sync(BG_FULL, c1b > 0) >> gb) & 0xffu) << 16);
    const uint32_t anyx = __ballot_sync(BG_FULL, c0a > K || c0b > K || c1a > K || c1b > K);
    const int n0 = __popc(o0), n = n0 + __popc(o1);
    __syncwarp();
    // ---- compacted row list (one row per occupied point: cumulative table up to K checkers) -------------------------------
    if (c0a) list[__popc(o0 & ltg)] = (uint32_t)((gl * RP + (int)min(c0a, (uint32_t)K) - 1) * H);
    if (c0b) list[__popc(o0 & ((ltg << 16) | 0xffffu))] = (uint32_t)(((16 + gl) * RP + (int)min(c0b, (uint32_t)K) - 1) * H);
    if (c1a) list[n0 + __popc(o1 & ltg)] = (uint32_t)(((24 + gl) * RP + (int)min(c1a, (uint32_t)K) - 1) * H);
    if (c1b) list[n0 + __popc(o1 & ((ltg << 16) | 0xffffu))] = (uint32_t)(((40 + gl) * RP + (int)min(c1b, (uint32_t)K) - 1) * H);
    int ne = 0;
    if (anyx) {  // stacks above K checkers (rare): (extra count, excess row) pairs
      const uint32_t x0a = (__ballot_sync(BG_FULL, c0a > K) >> gb) & 0xffffu, x0b = (__ballot_sync(BG_FULL, c0b > K) >> gb) & 0xffu;
      const uint32_t x1a = (__ballot_sync(BG_FULL, c1a > K) >> gb) & 0xffffu, x1b = (__ballot_sync(BG_FULL, c1b > K) >> gb) & 0xffu;
      const int e0 = __popc(x0a), e1 = e0 + __popc(x0b), e2 = e1 + __popc(x1a);
      ne = e2 + __popc(x1b);
      if (ne > ELIST_WORDS) ne = ELIST_WORDS;  // > 8 stacks of 7+ checkers cannot occur with 30 checkers
      if (c0a > K) elist[__popc(x0a & ltg)] = ((c0a - K) << 24) | (uint32_t)((gl * RP + K) * H);
      if (c0b > K) elist[e0 + __popc(x0b & ltg)] = ((c0b - K) << 24) | (uint32_t)(((16 + gl) * RP + K) * H);
      if (c1a > K) elist[e1 + __popc(x1a & ltg)] = ((c1a - K) << 24) | (uint32_t)(((24 + gl) * RP + K) * H);
      if (c1b > K) elist[e2 + __popc(x1b & ltg)] = ((c1b - K) << 24) | (uint32_t)(((40 + gl) * RP + K) * H);
    }
    __syncwarp();
    // ---- gather-sum (per half-warp; the two boards of a warp may have different row counts) ---------------------------
    Acc8 z;
    {
      const float* bias = sT + (MISC + 4 + flag) * H;
      const float4 x = *reinterpret_cast<const float4*>(bias + gl * 4);
      const float4 y = *reinterpret_cast<const float4*>(bias + 64 + gl * 4);
      z.a[0] = x.x; z.a[1] = x.y; z.a[2] = x.z; z.a[3] = x.w;
      z.a[4] = y.x; z.a[5] = y.y; z.a[6] = y.z; z.a[7] = y.w;
    }
    int j = 0;
    for (; j + 4 <= n; j += 4) {
      const uint4 r = *reinterpret_cast<const uint4*>(list + j);
      add2(z, sT + r.x, gl);
      add2(z, sT + r.y, gl);
      add2(z, sT + r.z, gl);
      add2(z, sT + r.w, gl);
    }
    for (; j < n; ++j) add2(z, sT + list[j], gl);
    for (int e = 0; e < ne; ++e) {
      const uint32_t w = elist[e];
      fma2(z, (float)(w >> 24), sT + (w & 0xffffffu), gl);
    }
    const uint32_t bar0 = w12 & 0xffu, bar1 = (w12 >> 8) & 0xffu, off0 = (w12 >> 16) & 0xffu, off1 = w12 >> 24;
    if (bar0) fma2(z, (float)bar0, sT + (MISC + 0) * H, gl);
    if (off0) fma2(z, c_off15b[off0 & 15u], sT + (MISC + 1) * H, gl);
    if (bar1) fma2(z, (float)bar1, sT + (MISC + 2) * H, gl);
    if (off1) fma2(z, c_off15b[off1 & 15u], sT + (MISC + 3) * H, gl);
    __syncwarp();
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float s = rcp_approx(1.0f + ex2_approx(z.a[q] * -1.4426950408889634f));
      v = fmaf(w2r[q], s, v);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(BG_FULL, v, o);
    if (gl == 0 && valid) out_v[i] = v + b2;
  }
}

}  // namespace

int64_t eval128_table_floats() { return (int64_t)TABLE_ROWS * H + 1; }

int32_t eval128_prepare(const float* packed, float* t128, cudaStream_t stream) {
  k_prepare128<<<128, 256, 0, stream>>>(packed, t128);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_prepare128 launch");
  return BG_OK;
}

int32_t eval128_launch(const EvalArgs& a, const float* t128, cudaStream_t stream) {
  static DeviceOnce once;
  constexpr size_t smem = (size_t)TABLE_ROWS * H * 4 + (size_t)WARPS * 2 * (LIST_WORDS + ELIST_WORDS) * 4;
  int32_t rc0 = once.run([&]() -> int32_t {
    float h[16];
    for (int n = 0; n < 16; ++n) h[n] = (float)((double)n / 15.0);
    cudaError_t e = cudaMemcpyToSymbol(c_off15b, h, sizeof(h));
    if (e != cudaSuccess) return check_cuda(e, "cudaMemcpyToSymbol(c_off15b)");
    return check_cuda(opt_in_shared(k_eval128, smem), "cudaFuncSetAttribute(k_eval128)");
  });
  if (rc0 != BG_OK) return rc0;
  const int64_t bound = a.N_dev ? a.max_N : a.N;
  int64_t want = (bound + WARPS * 2 - 1) / (WARPS * 2);
  if (want < 1) want = 1;
  const int grid = (int)(want < NUM_SMS ? want : NUM_SMS);
  k_eval128<<<grid, THREADS, smem, stream>>>(a.boards, a.flags, a.owner, a.owner_players, a.N, a.N_dev, a.max_N, t128, a.out_v, a.start_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_eval128 launch");
  return BG_OK;
}

}  // namespace bg
