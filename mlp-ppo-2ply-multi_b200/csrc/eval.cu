// eval.cu -- fused 198-feature encode + sigmoid-MLP value, and the stand-alone encoder (sm_100a).
//
// Replaces ImmutableBoard.get_board_features (reference src/backgammon/board/immutable_board.py:86-128),
// generate_all_board_features (src/environments/env_helper.py:7-24) and BackgammonPolicyNetwork.forward
// (src/agents/policy_network.py:53-70).
//
// The 198-vector of a board has <= ~31 non-zeros and the four per-point features are a thermometer code of the
// checker count, so layer 1 is evaluated as a gather-sum of pre-accumulated weight rows held in shared memory:
//   z = b1 + sum_{occupied (player, point)} Tcum[player][point][min(c,3)] + (c-3) * Thalf[player][point] + bar/off/flag rows
// i.e. one 512-byte (H=128) conflict-free row read per occupied point instead of a dense 198xH contraction; features
// never exist in memory.  One warp per board, lane owns H/32 hidden units; fp32 FFMA (meets the 1e-5 contract; a bf16
// tensor-core path would not).  Persistent CTAs keep the 100 KB table resident in shared memory.
#include "codes.cuh"
#include "eval.cuh"

#include <stdlib.h>

namespace bg {

namespace {

__constant__ float c_off15[16];  // (float)(n / 15.0), reference immutable_board.py:117,120

constexpr int NUM_SMS = 148;
constexpr int EVAL_THREADS = 512;

// prepared layout: rows [0,192): (pl*24+pt)*4 + k  (k=0: r0, 1: r0+r1, 2: r0+r1+r2, 3: 0.5*r3)
//                  rows 192..197: 0.5*W[192], W[193], 0.5*W[194], W[195], W[196], W[197]; then b1[H], w2[H], b2
__global__ void k_prepare(const float* __restrict__ packed, int H, float* __restrict__ prep) {
  const int total = 198 * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + 2 * H + 1; i += gridDim.x * blockDim.x) {
    if (i >= total) {
      prep[i] = packed[i];
      continue;
    }
    const int row = i / H, h = i - row * H;
    float v;
    if (row < 192) {
      const int k = row & 3, f0 = row - k;
      const float r0 = packed[(f0 + 0) * H + h], r1 = packed[(f0 + 1) * H + h], r2 = packed[(f0 + 2) * H + h];
      v = k == 0 ? r0 : k == 1 ? r0 + r1 : k == 2 ? (r0 + r1) + r2 : 0.5f * packed[(f0 + 3) * H + h];
    } else {
      v = packed[i];
      if (row == 192 || row == 194) v *= 0.5f;
    }
    prep[i] = v;
  }
}

template <int HPL>
struct Acc {
  float a[HPL];
};

template <int HPL>
__device__ __forceinline__ void add_row(Acc<HPL>& z, const float* row, int lane) {
  if constexpr (HPL == 4) {
    const float4 r = *reinterpret_cast<const float4*>(row + lane * 4);
    z.a[0] += r.x;
    z.a[1] += r.y;
    z.a[2] += r.z;
    z.a[3] += r.w;
  } else if constexpr (HPL == 8) {
    const float4 r = *reinterpret_cast<const float4*>(row + lane * 4);
    const float4 s = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
    z.a[0] += r.x;
    z.a[1] += r.y;
    z.a[2] += r.z;
    z.a[3] += r.w;
    z.a[4] += s.x;
    z.a[5] += s.y;
    z.a[6] += s.z;
    z.a[7] += s.w;
  } else {
#pragma unroll
    for (int q = 0; q < HPL; ++q) z.a[q] += row[q * 32 + lane];
  }
}

template <int HPL>
__device__ __forceinline__ void fma_row(Acc<HPL>& z, float s, const float* row, int lane) {
  if constexpr (HPL == 4) {
    const float4 r = *reinterpret_cast<const float4*>(row + lane * 4);
    z.a[0] = fmaf(s, r.x, z.a[0]);
    z.a[1] = fmaf(s, r.y, z.a[1]);
    z.a[2] = fmaf(s, r.z, z.a[2]);
    z.a[3] = fmaf(s, r.w, z.a[3]);
  } else if constexpr (HPL == 8) {
    const float4 r = *reinterpret_cast<const float4*>(row + lane * 4);
    const float4 t = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
    z.a[0] = fmaf(s, r.x, z.a[0]);
    z.a[1] = fmaf(s, r.y, z.a[1]);
    z.a[2] = fmaf(s, r.z, z.a[2]);
    z.a[3] = fmaf(s, r.w, z.a[3]);
    z.a[4] = fmaf(s, t.x, z.a[4]);
    z.a[5] = fmaf(s, t.y, z.a[5]);
    z.a[6] = fmaf(s, t.z, z.a[6]);
    z.a[7] = fmaf(s, t.w, z.a[7]);
  } else {
#pragma unroll
    for (int q = 0; q < HPL; ++q) z.a[q] = fmaf(s, row[q * 32 + lane], z.a[q]);
  }
}

// hidden-unit ownership of a lane must match add_row's addressing:
//   HPL==4: units lane*4..+3 ; HPL==8: lane*4..+3 and 128+lane*4..+3 ; else q*32+lane
template <int HPL>
__device__ __forceinline__ int unit_index(int lane, int q) {
  if constexpr (HPL == 4)
    return lane * 4 + q;
  else if constexpr (HPL == 8)
    return (q >> 2) * 128 + lane * 4 + (q & 3);
  else
    return q * 32 + lane;
}

// Row gather for one board.  `list` (per-warp shared scratch, 16-byte aligned) receives the table-row offsets (in floats)
// of every occupied (player, point) pair, compacted across lanes with ballot/popc; the warp then streams them four rows at a
// time (one broadcast 16-byte read of four offsets, then four independent 512-byte conflict-free row reads at H=128).
template <int HPL>
__device__ __forceinline__ void gather_rows(Acc<HPL>& z, const float* sT, uint32_t* list, uint32_t c0, uint32_t c1, int lane) {
  constexpr int H = HPL * 32;
  const uint32_t occ0 = __ballot_sync(BG_FULL, c0 > 0), occ1 = __ballot_sync(BG_FULL, c1 > 0);
  const uint32_t gt0 = __ballot_sync(BG_FULL, c0 > 3), gt1 = __ballot_sync(BG_FULL, c1 > 3);
  const uint32_t lt = (1u << lane) - 1u;
  const int n0 = __popc(occ0), n = n0 + __popc(occ1);
  __syncwarp();
  if (c0 > 0) list[__popc(occ0 & lt)] = (uint32_t)((lane * 4 + (int)min(c0, 3u) - 1) * H);
  if (c1 > 0) list[n0 + __popc(occ1 & lt)] = (uint32_t)((96 + lane * 4 + (int)min(c1, 3u) - 1) * H);
  __syncwarp();
  int j = 0;
  for (; j + 4 <= n; j += 4) {
    const uint4 r = *reinterpret_cast<const uint4*>(list + j);
    add_row<HPL>(z, sT + r.x, lane);
    add_row<HPL>(z, sT + r.y, lane);
    add_row<HPL>(z, sT + r.z, lane);
    add_row<HPL>(z, sT + r.w, lane);
  }
  for (; j < n; ++j) add_row<HPL>(z, sT + list[j], lane);
  // counts above 3: (c - 3) * (0.5 * W[.,3])
  uint32_t m = gt0;
  while (m) {
    const int p = __ffs(m) - 1;
    m &= m - 1;
    const int c = __shfl_sync(BG_FULL, (int)c0, p);
    fma_row<HPL>(z, (float)(c - 3), sT + (p * 4 + 3) * H, lane);
  }
  m = gt1;
  while (m) {
    const int p = __ffs(m) - 1;
    m &= m - 1;
    const int c = __shfl_sync(BG_FULL, (int)c1, p);
    fma_row<HPL>(z, (float)(c - 3), sT + (96 + p * 4 + 3) * H, lane);
  }
}

template <int HPL>
__global__ void __launch_bounds__(EVAL_THREADS, (HPL <= 4 ? 2 : 1))
    k_eval(const int8_t* __restrict__ boards, const uint8_t* __restrict__ flags, const int32_t* __restrict__ owner,
           const uint8_t* __restrict__ owner_players, int64_t N_host, const int64_t* __restrict__ N_dev, int64_t max_N,
           const float* __restrict__ prep, float* __restrict__ out_v, const int64_t* __restrict__ start_dev) {
  constexpr int H = HPL * 32;
  extern __shared__ __align__(16) float sT[];
  const int n_floats = 200 * H + 1;
  for (int i = threadIdx.x; i < n_floats; i += blockDim.x) sT[i] = prep[i];
  uint32_t* lists = reinterpret_cast<uint32_t*>(sT + 202 * H);
  __syncthreads();
  const float* sb1 = sT + 198 * H;
  const float* sw2 = sb1 + H;
  const float b2 = sw2[H];
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
  const int lane = threadIdx.x & 31;
  uint32_t* list = lists + (threadIdx.x >> 5) * 56;
  const int64_t warp = (int64_t)blockIdx.x * (EVAL_THREADS / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (EVAL_THREADS / 32);
  const uint32_t* b32 = reinterpret_cast<const uint32_t*>(boards);
  // per-lane constants: w2 and the two possible initial accumulators b1 + W[196 + flag] (the flag row is always present)
  float w2r[HPL], bf0[HPL], bf1[HPL];
#pragma unroll
  for (int q = 0; q < HPL; ++q) {
    const int u = unit_index<HPL>(lane, q);
    w2r[q] = sw2[u];
    bf0[q] = sb1[u] + sT[196 * H + u];
    bf1[q] = sb1[u] + sT[197 * H + u];
  }
  // lanes 0..12 fetch the board words, lane 13 the flag; the next board is prefetched while this one is evaluated
  auto fetch = [&](int64_t j) -> uint32_t {
    if (lane < 13) return __ldg(b32 + j * 13 + lane);
    if (lane == 13) return flags ? (uint32_t)flags[j] : (uint32_t)owner_players[owner[j]];
    return 0u;
  };
  uint32_t nxt = warp < N ? fetch(warp) : 0u;
  for (int64_t i = warp; i < N; i += nwarps) {
    const uint32_t myw = nxt;
    if (i + nwarps < N) nxt = fetch(i + nwarps);
    const int flag = (int)__shfl_sync(BG_FULL, myw, 13);
    const uint32_t wa = __shfl_sync(BG_FULL, myw, lane >> 2);
    const uint32_t wb = __shfl_sync(BG_FULL, myw, 6 + (lane >> 2));
    const uint32_t w12 = __shfl_sync(BG_FULL, myw, 12);
    const uint32_t c0 = lane < 24 ? (wa >> ((lane & 3) * 8)) & 0xffu : 0u;
    const uint32_t c1 = lane < 24 ? (wb >> ((lane & 3) * 8)) & 0xffu : 0u;
    Acc<HPL> z;
#pragma unroll
    for (int q = 0; q < HPL; ++q) z.a[q] = (flag & 1) ? bf1[q] : bf0[q];
    gather_rows<HPL>(z, sT, list, c0, c1, lane);
    const uint32_t bar0 = w12 & 0xffu, bar1 = (w12 >> 8) & 0xffu, off0 = (w12 >> 16) & 0xffu, off1 = w12 >> 24;
    if (bar0) fma_row<HPL>(z, (float)bar0, sT + 192 * H, lane);
    if (off0) fma_row<HPL>(z, c_off15[off0 & 15u], sT + 193 * H, lane);
    if (bar1) fma_row<HPL>(z, (float)bar1, sT + 194 * H, lane);
    if (off1) fma_row<HPL>(z, c_off15[off1 & 15u], sT + 195 * H, lane);
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < HPL; ++q) {
      const float s = __fdividef(1.0f, 1.0f + __expf(-z.a[q]));
      v = fmaf(w2r[q], s, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BG_FULL, v, o);
    if (lane == 0) out_v[i] = v + b2;
  }
}

// Materialised 198-feature rows (immutable_board.py:86-128), bit-exact fp32.  Purely bandwidth-bound (52 + 1 bytes in, 792 out per row), so
// it is written around 16-byte stores: a warp encodes TWO consecutive rows (2 x 792 B = 99 float4, 16-byte aligned for an even first row);
// a float4 of the first row is one point's thermometer code (a 16-entry table read), of the second row the tail of one point and the
// head of the next.
__device__ __forceinline__ float4 thermo4(int c) {
  return make_float4(c >= 1 ? 1.f : 0.f, c >= 2 ? 1.f : 0.f, c >= 3 ? 1.f : 0.f, c > 3 ? (float)(c - 3) * 0.5f : 0.f);
}

__global__ void __launch_bounds__(256) k_encode(const int8_t* __restrict__ boards, const uint8_t* __restrict__ flags, int64_t N,
                                                float* __restrict__ out) {
  __shared__ uint32_t sb[8][32];
  __shared__ float4 s_t4[16];
  if (threadIdx.x < 16) s_t4[threadIdx.x] = thermo4((int)threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)gridDim.x * 8;
  const uint32_t* b32 = reinterpret_cast<const uint32_t*>(boards);
  const int64_t n_pairs = (N + 1) / 2;
  for (int64_t pr = warp; pr < n_pairs; pr += nwarps) {
    const int64_t i0 = pr * 2;
    const bool two = i0 + 1 < N;
    __syncwarp();
    if (lane < (two ? 26 : 13)) sb[wib][lane] = __ldg(b32 + i0 * 13 + lane);  // the two boards are 26 consecutive words
    __syncwarp();
    const uint8_t* b0 = reinterpret_cast<const uint8_t*>(sb[wib]);
    const uint8_t* b1 = b0 + 52;
    const int flag0 = flags[i0] & 1, flag1 = two ? flags[i0 + 1] & 1 : 0;
    float4* o4 = reinterpret_cast<float4*>(out + i0 * 198);  // i0 even: 16-byte aligned
    if (two) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int j = lane + 32 * it;
        if (j >= 99) break;
        float4 x;
        if (j < 48) {
          x = s_t4[b0[j] & 15u];
        } else if (j == 48) {
          x = make_float4((float)(int8_t)b0[48] * 0.5f, c_off15[b0[50] & 15u], (float)(int8_t)b0[49] * 0.5f, c_off15[b0[51] & 15u]);
        } else if (j == 49) {
          const float4 n = s_t4[b1[0] & 15u];
          x = make_float4(flag0 == 0 ? 1.f : 0.f, flag0 == 1 ? 1.f : 0.f, n.x, n.y);
        } else if (j < 97) {
          const float4 a = s_t4[b1[j - 50] & 15u], n = s_t4[b1[j - 49] & 15u];
          x = make_float4(a.z, a.w, n.x, n.y);
        } else if (j == 97) {
          const float4 a = s_t4[b1[47] & 15u];
          x = make_float4(a.z, a.w, (float)(int8_t)b1[48] * 0.5f, c_off15[b1[50] & 15u]);
        } else {
          x = make_float4((float)(int8_t)b1[49] * 0.5f, c_off15[b1[51] & 15u], flag1 == 0 ? 1.f : 0.f, flag1 == 1 ? 1.f : 0.f);
        }
        o4[j] = x;
      }
    } else {  // the last row of an odd batch: 49 float4 + one float2
      for (int j = lane; j < 50; j += 32) {
        if (j < 48) {
          o4[j] = s_t4[b0[j] & 15u];
        } else if (j == 48) {
          o4[j] = make_float4((float)(int8_t)b0[48] * 0.5f, c_off15[b0[50] & 15u], (float)(int8_t)b0[49] * 0.5f, c_off15[b0[51] & 15u]);
        } else {
          reinterpret_cast<float2*>(o4)[98] = make_float2(flag0 == 0 ? 1.f : 0.f, flag0 == 1 ? 1.f : 0.f);
        }
      }
    }
  }
}

int32_t init_constants() {
  static DeviceOnce once;
  return once.run([]() -> int32_t {
    float h[16];
    for (int n = 0; n < 16; ++n) h[n] = (float)((double)n / 15.0);
    return check_cuda(cudaMemcpyToSymbol(c_off15, h, sizeof(h)), "cudaMemcpyToSymbol(c_off15)");
  });
}

template <int HPL>
int32_t launch_eval_t(const EvalArgs& a, cudaStream_t stream) {
  constexpr int H = HPL * 32;
  const size_t smem = (size_t)(202 * H) * 4 + (EVAL_THREADS / 32) * 56 * 4;
  static DeviceOnce once;
  int32_t rc0 = once.run([&]() -> int32_t {
    return check_cuda(opt_in_shared(k_eval<HPL>, smem), "cudaFuncSetAttribute(k_eval)");
  });
  if (rc0 != BG_OK) return rc0;
  const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int64_t bound = a.N_dev ? a.max_N : a.N;
  int64_t want = (bound + (EVAL_THREADS / 32) - 1) / (EVAL_THREADS / 32);
  if (want < 1) want = 1;
  const int grid = (int)(want < (int64_t)NUM_SMS * ctas_per_sm ? want : (int64_t)NUM_SMS * ctas_per_sm);
  k_eval<HPL><<<grid, EVAL_THREADS, smem, stream>>>(a.boards, a.flags, a.owner, a.owner_players, a.N, a.N_dev, a.max_N, a.prepared,
                                                    a.out_v, a.start_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_eval launch");
  return BG_OK;
}

}  // namespace

// generic table (200 * H + 1 floats, padded to 256 B) followed, for H == 128, by the table of the two-boards-per-warp kernel
static int64_t generic_table_bytes(int32_t H) { return ((int64_t)(200 * H + 1) * 4 + 255) / 256 * 256; }

static int64_t table128_bytes() { return (eval128_table_floats() * 4 + 255) / 256 * 256; }

int64_t prepared_weights_bytes(int32_t H) {
  return generic_table_bytes(H) + (H == 128 ? table128_bytes() : 0) + (H <= 128 ? 1 : 2) * eval_tc_image_bytes();
}

// which evaluator for H <= 128: BG_EVAL_PATH = "tc" (tcgen05 for batches >= 32768 rows, default), "ffma" (always the CUDA-core kernels)
static int tc_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("BG_EVAL_PATH");
    mode = (e && e[0] == 'f') ? 0 : 1;
  }
  return mode;
}

static int32_t* g_tc_err = nullptr;

int32_t eval_tc_status() {
  if (!g_tc_err) return 0;
  int32_t h = 0;
  if (cudaMemcpy(&h, g_tc_err, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return h;
}

int32_t prepare_weights_launch(const float* packed, int32_t H, float* prepared, cudaStream_t stream) {
  if (H < 32 || H > 256 || H % 32) {
    set_error("H must be a multiple of 32 in [32,256], got %d", H);
    return BG_ERR_ARG;
  }
  k_prepare<<<64, 256, 0, stream>>>(packed, H, prepared);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_prepare launch");
  if (H == 128) {
    int32_t rc = eval128_prepare(packed, prepared + generic_table_bytes(H) / 4, stream);
    if (rc != BG_OK) return rc;
    return eval_tc_prepare(packed, H, 0, reinterpret_cast<uint8_t*>(prepared) + generic_table_bytes(H) + table128_bytes(), stream);
  }
  uint8_t* img = reinterpret_cast<uint8_t*>(prepared) + generic_table_bytes(H);
  int32_t rc = eval_tc_prepare(packed, H, 0, img, stream);
  if (rc == BG_OK && H > 128) rc = eval_tc_prepare(packed, H, 128, img + eval_tc_image_bytes(), stream);
  return rc;
}

int32_t eval_launch(const EvalArgs& a, cudaStream_t stream) {
  if (a.H < 32 || a.H > 256 || a.H % 32) {
    set_error("bg_eval: H must be a multiple of 32 in [32,256] (got %d)", a.H);
    return BG_ERR_ARG;
  }
  if (!a.flags && !(a.owner && a.owner_players)) {
    set_error("bg_eval: need flags or (owner, owner_players)");
    return BG_ERR_ARG;
  }
  if ((a.N_dev ? a.max_N : a.N) <= 0) return BG_OK;
  {
    // large batches go to the tensor-core kernel, 128 hidden units per pass: smaller nets zero-padded, wider nets in two passes
    const int64_t bound = a.N_dev ? a.max_N : a.N;
    if (a.codes && !a.flags) {
      set_error("bg_eval: the compact form needs the positions' players");
      return BG_ERR_ARG;
    }
    if (a.codes || (tc_mode() == 1 && a.flags && bound >= 32768)) {
      if (!g_tc_err) {
        cudaError_t e = cudaMalloc(&g_tc_err, 4);
        if (e == cudaSuccess) e = cudaMemset(g_tc_err, 0, 4);
        if (e != cudaSuccess) return check_cuda(e, "cudaMalloc(tc status)");
      }
      const uint8_t* img = reinterpret_cast<const uint8_t*>(a.prepared) + generic_table_bytes(a.H) + (a.H == 128 ? table128_bytes() : 0);
      int32_t rc = eval_tc_launch(a, img, g_tc_err, stream, 0);
      if (rc == BG_OK && a.H > 128) rc = eval_tc_launch(a, img + eval_tc_image_bytes(), g_tc_err, stream, 1);
      return rc;
    }
    if (a.H == 128) return eval128_launch(a, a.prepared + generic_table_bytes(128) / 4, stream);
  }
  int32_t rc = init_constants();
  if (rc != BG_OK) return rc;
  switch (a.H / 32) {
    case 1:
      return launch_eval_t<1>(a, stream);
    case 2:
      return launch_eval_t<2>(a, stream);
    case 3:
      return launch_eval_t<3>(a, stream);
    case 4:
      return launch_eval_t<4>(a, stream);
    case 5:
      return launch_eval_t<5>(a, stream);
    case 6:
      return launch_eval_t<6>(a, stream);
    case 7:
      return launch_eval_t<7>(a, stream);
    default:
      return launch_eval_t<8>(a, stream);
  }
}

int32_t encode_launch(const int8_t* boards, const uint8_t* flags, int64_t N, float* out, cudaStream_t stream) {
  if (N <= 0) return BG_OK;
  int32_t rc = init_constants();
  if (rc != BG_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(out) & 15u) || (reinterpret_cast<uintptr_t>(boards) & 3u)) {
    set_error("bg_encode: out must be 16-byte aligned and boards 4-byte aligned");
    return BG_ERR_ARG;
  }
  int64_t want = ((N + 1) / 2 + 7) / 8;
  const int grid = (int)(want < (int64_t)NUM_SMS * 8 ? want : (int64_t)NUM_SMS * 8);
  k_encode<<<grid, 256, 0, stream>>>(boards, flags, N, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_encode launch");
  return BG_OK;
}

// rows of a compact pool -> boards: out[j] = position board of codes[rows[j]] with the code applied (rows == nullptr: rows 0 .. n-1)
__global__ void __launch_bounds__(256) k_materialize(const int8_t* __restrict__ boards, const uint8_t* __restrict__ players, const uint2* __restrict__ codes,
                                                     const long long* __restrict__ rows, int64_t n, int8_t* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const long long rrow = rows ? rows[j] : j;
  uint32_t w[13];
  if (rrow < 0) {  // an item without a legal move (action -1): zeros
#pragma unroll
    for (int q = 0; q < 13; ++q) reinterpret_cast<uint32_t*>(out)[j * 13 + q] = 0u;
    return;
  }
  const uint2 e = codes[rrow];
  const uint32_t* src = reinterpret_cast<const uint32_t*>(boards) + (int64_t)e.y * 13;
#pragma unroll
  for (int q = 0; q < 13; ++q) w[q] = src[q];
  apply_code_bytes(reinterpret_cast<uint8_t*>(w), e.x, players[e.y] & 1);
#pragma unroll
  for (int q = 0; q < 13; ++q) reinterpret_cast<uint32_t*>(out)[j * 13 + q] = w[q];
}

int32_t materialize_launch(const int8_t* boards, const uint8_t* players, const uint2* codes, const int64_t* rows, int64_t n, int8_t* out,
                           cudaStream_t stream) {
  if (n <= 0) return BG_OK;
  k_materialize<<<(int)((n + 255) / 256), 256, 0, stream>>>(boards, players, codes, reinterpret_cast<const long long*>(rows), n, out);
  return check_cuda(cudaGetLastError(), "k_materialize launch");
}

int32_t side_ctx_create(SideCtx* c, bool high_priority) {
  if (c->stream) return BG_OK;
  int lo = 0, hi = 0;
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, high_priority ? hi : lo);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_t1, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming);
  return check_cuda(e, "side stream");
}

void side_ctx_destroy(SideCtx* c) {
  if (c->stream) {
    cudaStreamSynchronize(c->stream);
    cudaStreamDestroy(c->stream);
  }
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->ev_side) cudaEventDestroy(c->ev_side);
  *c = SideCtx{};
}

int32_t movegen_eval_overlapped(MovegenArgs m, int64_t* total2, const float* prepared, int32_t H, float* out_v, SideCtx* side,
                                cudaStream_t stream) {
  m.out_total = total2;
  int32_t rc;
  if (!side || !side->stream) {
    if ((rc = movegen_launch(m, stream)) != BG_OK) return rc;
    EvalArgs ev{m.out_boards, m.out_flags, nullptr, nullptr, 0, total2, m.pool_cap, prepared, H, out_v};
    if (m.out_codes) {
      ev.boards = m.boards;
      ev.flags = m.players;
      ev.codes = m.out_codes;
    }
    return eval_launch(ev, stream);
  }
  m.tier1_total = total2 + 1;
  m.tier1_event = side->ev_t1;
  if ((rc = movegen_launch(m, stream)) != BG_OK) return rc;
  cudaError_t e = cudaStreamWaitEvent(side->stream, side->ev_t1, 0);
  if (e != cudaSuccess) return check_cuda(e, "wait bulk tier");
  EvalArgs e1{m.out_boards, m.out_flags, nullptr, nullptr, 0, total2 + 1, m.pool_cap, prepared, H, out_v};
  if (m.out_codes) {  // compact pool: the evaluator rebuilds each afterstate from its position + code
    e1.boards = m.boards;
    e1.flags = m.players;
    e1.codes = m.out_codes;
  }
  if ((rc = eval_launch(e1, side->stream)) != BG_OK) return rc;
  e = cudaEventRecord(side->ev_side, side->stream);
  if (e != cudaSuccess) return check_cuda(e, "record side");
  EvalArgs e2 = e1;
  e2.N_dev = total2;
  e2.start_dev = total2 + 1;
  if ((rc = eval_launch(e2, stream)) != BG_OK) return rc;
  return check_cuda(cudaStreamWaitEvent(stream, side->ev_side, 0), "join side stream");
}

}  // namespace bg
