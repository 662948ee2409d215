// eval_tc.cu -- tcgen05 / TMEM evaluator for H = 128 (sm_100a): fused 198-feature encode + value MLP on the 5th-gen tensor cores.
//
// Same contract as eval.cu / eval128.cu (reference src/backgammon/board/immutable_board.py:86-128 +
// src/agents/policy_network.py:53-70, |dV| <= 1e-5), for LARGE afterstate batches where layer 1 is a real GEMM:
//   Z[128 boards, 128 hidden] = X[128, 208] * W^T[208, 128]
// * fp32-equivalent on fp16 tensor cores: every feature value is EXACT in fp16 ({0, 1, k/2}); the two off/15 features are
//   carried as three fp16 terms (hi + mid + lo) in spare K columns; the weights (and b1, as the column of a constant-1
//   feature) are split into TWO fp16 pieces W = hi + lo: 11 + 11 significand bits, so the residual is <= 2^-24 |W|, the
//   rounding error fp32 itself has (valid for |W| < 65504; fp16 subnormals keep the low piece of tiny weights to 3e-8 abs).
//   The 2 x 13 tcgen05.mma (kind::f16, M128 N128 K16, fp32 accumulate in TMEM) form every product exactly and only the
//   accumulation rounds.  (Round 1 first used three bf16 pieces: same accuracy, 1.5x the tensor work.)
// * A never touches shared memory: each of 128 worker threads owns one TMEM lane (= one board), builds its 208-entry
//   fp16 feature row in registers and tcgen05.st's it into TMEM (A-from-TMEM "TS" MMA); B (2 x 128 x 208 fp16 = 104 KB,
//   no-swizzle K-major core-matrix layout) stays resident in shared memory; D is read back with tcgen05.ld and the
//   sigmoid / w2 dot product / +b2 epilogue runs one thread per board.
// * two worker groups (2 x 4 warps) alternate tiles against one MMA-issuer warp, TMEM = [A0 | D0 | A1 | D1] (512 columns),
//   so one group's epilogue + next feature build overlaps the other group's MMAs.  mbarriers: full[g] (128 arrivals) ->
//   MMA thread; tcgen05.commit -> done[g] -> workers.  All waits are bounded (an error flag instead of a hang).
#include <cuda_fp16.h>

#include "codes.cuh"
#include "eval.cuh"

namespace bg {

namespace {

constexpr int H = 128, KP = 208, KSTEPS = KP / 16, NSPLIT = 2;
constexpr int KCHUNK_BYTES = 16 * 128;                      // one 8-wide K chunk of all 128 rows: 16 N-groups x 128 B
constexpr int SPLIT_BYTES = (KP / 8) * KCHUNK_BYTES;        // 26 chunks = 53,248 B
constexpr int B_BYTES = NSPLIT * SPLIT_BYTES;               // 106,496 B
constexpr int THREADS = 288;                                // warps 0-3 group 0, 4-7 group 1, warp 8 MMA issuer
constexpr int TMEM_COLS = 512, A_COLS = 128, D_COLS = 128;  // per group: A at +0 (104 used), D at +128
constexpr int NUM_SMS = 148;
// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = F16 (format fields 0), N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__constant__ uint32_t c_off15_split[16][2];  // n/15 = hi + mid + lo in fp16: [n][0] = hi | mid << 16, [n][1] = lo

// feature index -> source row of the packed weights for the K-padded operand (198..201: off/15 mid/lo terms, 202: bias)
__device__ __forceinline__ int krow_source(int k) {
  if (k < 198) return k;
  if (k == 198 || k == 199) return 193;
  if (k == 200 || k == 201) return 195;
  if (k == 202) return 198;  // b1 lives right after W1^T in the packed blob
  return -1;
}

// B operand image (fp16, 2 pieces, canonical no-swizzle K-major layout), built once per weight set
// Hs: hidden size of the packed source weights; the image holds its units [u0, u0 + 128), units >= Hs are zero padding (zero weights and
// zero w2: they add nothing).  b2 goes into the image of the first 128 units only.
__global__ void k_prepare_tc(const float* __restrict__ packed, int Hs, int u0, uint8_t* __restrict__ img) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < KP * H; i += gridDim.x * blockDim.x) {
    const int k = i / H, n = i - k * H;
    const int src = krow_source(k);
    // the operand is W scaled by -log2(e): the accumulator then holds y = -z log2(e) and the epilogue's sigmoid is 1 / (1 + 2^y) with no
    // multiply per unit
    const float w = (src < 0 || u0 + n >= Hs) ? 0.f : packed[src * Hs + u0 + n] * -1.4426950408889634f;
    // two fp16 pieces carry 22+ mantissa bits: w - (hi + lo) <= 2^-24 |w|, the rounding error of fp32 itself (|w| < 65504)
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const int off = (k >> 3) * KCHUNK_BYTES + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(img + 0 * SPLIT_BYTES + off) = hi;
    *reinterpret_cast<__half*>(img + 1 * SPLIT_BYTES + off) = lo;
  }
  // trailer: w2[128] (zero for the padding units), b2
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= H; i += gridDim.x * blockDim.x)
    reinterpret_cast<float*>(img + B_BYTES)[i] = i == H ? (u0 == 0 ? packed[200 * Hs] : 0.f) : u0 + i < Hs ? packed[199 * Hs + u0 + i] : 0.f;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
// bounded wait: returns false on timeout (so a descriptor / protocol bug raises an error instead of hanging the GPU)
__device__ __forceinline__ bool mbar_wait(uint32_t addr, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
    if (it > 64) __nanosleep(64);
  }
  return false;
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

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

// thermometer code of one checker count as two packed fp16x2 words: (c>=1, c>=2), (c>=3, (c-3)/2)
__device__ __forceinline__ void point_words(uint32_t c, uint32_t& w0, uint32_t& w1) {
  const uint32_t ONE = 0x3c00u;  // fp16 1.0
  w0 = (c >= 1 ? ONE : 0u) | (c >= 2 ? ONE << 16 : 0u);
  uint32_t ex = 0;
  if (c > 3) ex = (uint32_t)__half_as_ushort(__float2half_rn((float)(c - 3) * 0.5f));  // 0.5 .. 6.0: exact in fp16
  w1 = (c >= 3 ? ONE : 0u) | (ex << 16);
}

// CODES: the rows are (code, position index) pairs of the compact pool (codes.cuh); `boards` / `flags` are then the POSITIONS' boards and
// players, and each worker thread rebuilds its afterstate in a 52-byte shared-memory scratch row (13 word stores, a few byte updates, 13 word
// loads: ~6 % more instructions in a kernel that is tensor-bound) -- the afterstate boards never exist in HBM.
constexpr int SCRATCH_BYTES = 2 * 128 * 13 * 4;
template <bool CODES>
__global__ void __launch_bounds__(THREADS, 1)
    k_eval_tc(const int8_t* __restrict__ boards, const uint8_t* __restrict__ flags, int64_t N_host, const int64_t* __restrict__ N_dev,
              int64_t max_N, const uint8_t* __restrict__ img, float* __restrict__ out_v, int32_t* __restrict__ err,
              const int64_t* __restrict__ start_dev, int accumulate, const uint2* __restrict__ codes) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;                                             // B operand image
  float* sW2 = reinterpret_cast<float*>(smem + B_BYTES);          // w2[128], b2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_BYTES + 1024);  // full[2], done[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B_BYTES + 1024 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the four fp16 thermometer features of a point by checker count: one 8-byte table read instead of ~10 ALU instructions per point
  // (16 entries x 8 B cover the 32 banks exactly once: conflict free for any mix of counts)
  __shared__ uint2 s_xtab[16];
  if (tid < 16) {
    uint32_t w0, w1;
    point_words((uint32_t)tid, w0, w1);
    s_xtab[tid] = make_uint2(w0, w1);
  }

  for (int i = tid; i < B_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(img)[i];
  for (int i = tid; i <= H; i += THREADS) sW2[i] = reinterpret_cast<const float*>(img + B_BYTES)[i];
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 128);
    mbar_init(smem_u32(&bars[1]), 128);
    mbar_init(smem_u32(&bars[2]), 1);
    mbar_init(smem_u32(&bars[3]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // make the generic-proxy writes of B visible to the async (tensor-core) proxy, then publish the TMEM base
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  int64_t N = N_dev ? *N_dev : N_host;
  if (N > max_N) N = max_N;
  {  // optional device-side start row: shift the row-indexed arrays once, everything below is unchanged
    int64_t start = start_dev ? *start_dev : 0;
    if (start > N) start = N;
    if (start < 0) start = 0;
    if constexpr (CODES) {
      codes += start;
    } else {
      boards += start * BG_BOARD_BYTES;
      if (flags) flags += start;
    }
    out_v += start;
    N -= start;
  }
  const int64_t n_tiles = (N + 127) / 128;
  // tile t of the grid is owned by CTA (t / 2) % gridDim.x, group t & 1
  if (warp < 8) {
    // ================= workers: one thread per TMEM lane / board =================
    const int g = warp >> 2, q = warp & 3, row = q * 32 + lane;
    const uint32_t tA = tmem_base + (uint32_t)(g * (A_COLS + D_COLS)) + ((uint32_t)(q * 32) << 16);
    const uint32_t tD = tA + A_COLS;
    const uint32_t full = smem_u32(&bars[g]), done = smem_u32(&bars[2 + g]);
    const uint32_t* b32 = reinterpret_cast<const uint32_t*>(boards);
    uint32_t it = 0;
    // board words of the CURRENT tile live in registers; the next tile's are fetched before waiting on the tensor core
    uint32_t bw[13];
    uint32_t flag = 0, code = 0;
    uint32_t* const scr = reinterpret_cast<uint32_t*>(smem + B_BYTES + 1024 + 128) + (g * 128 + row) * 13;  // CODES: this thread's scratch row
    auto fetch = [&](int64_t tile) {
      const int64_t i = tile * 128 + row;
      if (tile < n_tiles && i < N) {
        int64_t src = i;
        if constexpr (CODES) {
          const uint2 e = __ldg(codes + i);
          code = e.x;
          src = (int64_t)e.y;  // consecutive rows share their position: these loads hit L1
        }
#pragma unroll
        for (int w = 0; w < 13; ++w) bw[w] = __ldg(b32 + src * 13 + w);
        flag = flags[src] & 1u;
        if constexpr (CODES) {  // position board -> afterstate board, in this thread's scratch row (13-word stride: conflict free)
#pragma unroll
          for (int w = 0; w < 13; ++w) scr[w] = bw[w];
          apply_code_bytes(reinterpret_cast<uint8_t*>(scr), code, (int)flag);
#pragma unroll
          for (int w = 0; w < 13; ++w) bw[w] = scr[w];
        }
      } else {
#pragma unroll
        for (int w = 0; w < 13; ++w) bw[w] = 0u;
        flag = 0u;
      }
    };
    const int64_t tstride = (int64_t)gridDim.x * 2;
    fetch((int64_t)blockIdx.x * 2 + g);
    for (int64_t t = (int64_t)blockIdx.x * 2 + g; t < n_tiles; t += tstride, ++it) {
      const int64_t i = t * 128 + row;
      const bool valid = i < N;
      // ---- build this board's fp16 feature row, 8 TMEM columns (= one K16 step = four points) at a time ----
#pragma unroll
      for (int wd = 0; wd < 12; ++wd) {  // board word wd: 4 points -> 16 features -> 8 columns
        uint32_t r[8];
        const uint2 p0 = s_xtab[bw[wd] & 15u], p1 = s_xtab[(bw[wd] >> 8) & 15u], p2 = s_xtab[(bw[wd] >> 16) & 15u], p3 = s_xtab[(bw[wd] >> 24) & 15u];
        r[0] = p0.x;
        r[1] = p0.y;
        r[2] = p1.x;
        r[3] = p1.y;
        r[4] = p2.x;
        r[5] = p2.y;
        r[6] = p3.x;
        r[7] = p3.y;
        tmem_st8(tA + wd * 8, r);
      }
      {
        const uint32_t ONE = 0x3c00u;
        const uint32_t w12 = bw[12];
        const uint32_t bar0 = w12 & 0xffu, bar1 = (w12 >> 8) & 0xffu, off0 = (w12 >> 16) & 15u, off1 = (w12 >> 24) & 15u;
        const uint32_t hb0 = __half_as_ushort(__float2half_rn((float)bar0 * 0.5f));
        const uint32_t hb1 = __half_as_ushort(__float2half_rn((float)bar1 * 0.5f));
        const uint32_t s0a = c_off15_split[off0][0], s0b = c_off15_split[off0][1];
        const uint32_t s1a = c_off15_split[off1][0], s1b = c_off15_split[off1][1];
        uint32_t r[8];
        r[0] = hb0 | ((s0a & 0xffffu) << 16);                       // 192 bar0/2, 193 off0 hi
        r[1] = hb1 | ((s1a & 0xffffu) << 16);                       // 194 bar1/2, 195 off1 hi
        r[2] = valid ? (flag ? ONE << 16 : ONE) : 0u;               // 196, 197 flag one-hot
        r[3] = (s0a >> 16) | (s0b << 16);                           // 198 off0 mid, 199 off0 lo
        r[4] = (s1a >> 16) | (s1b << 16);                           // 200 off1 mid, 201 off1 lo
        r[5] = valid ? ONE : 0u;                                    // 202 constant 1 (bias column), 203 = 0
        r[6] = 0u;
        r[7] = 0u;
        tmem_st8(tA + 96, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(full);
      fetch(t + tstride);  // overlaps the MMAs
      // ---- wait for the 26 MMAs of this tile, then the epilogue straight out of TMEM ----
      if (!mbar_wait(done, it & 1u)) {
        if (lane == 0) atomicExch(err, 1);
        break;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < H; c0 += 32) {
        uint32_t z[32];
        tmem_ld32(tD + c0, z);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          // the epilogue is SFU-bound (ex2 and rcp share the 16-lane MUFU pipe), so FOUR sigmoids share one reciprocal:
          // with p = (1+a)(1+b), q = (1+c)(1+d), r = 1/(pq):  1/(1+a) = (1+b) q r, ...  The accumulator is y = -z log2(e) (scaled
          // weights); y is clamped at 28.85 (z >= -20: sigmoid < 2.1e-9, far below the 1e-5 contract) so that the product of four terms
          // stays below 5.5e34
          const float a1 = 1.0f + ex2_approx(fminf(__uint_as_float(z[c]), 28.853901f));
          const float b1 = 1.0f + ex2_approx(fminf(__uint_as_float(z[c + 1]), 28.853901f));
          const float c1 = 1.0f + ex2_approx(fminf(__uint_as_float(z[c + 2]), 28.853901f));
          const float d1 = 1.0f + ex2_approx(fminf(__uint_as_float(z[c + 3]), 28.853901f));
          const float p = a1 * b1, q = c1 * d1;
          const float r = rcp_approx(p * q);
          const float rp = r * q, rq = r * p;  // 1/p, 1/q
          const float4 w4 = *reinterpret_cast<const float4*>(&sW2[c0 + c]);  // one broadcast 16-byte read per four units
          v = fmaf(w4.x, b1 * rp, v);
          v = fmaf(w4.y, a1 * rp, v);
          v = fmaf(w4.z, d1 * rq, v);
          v = fmaf(w4.w, c1 * rq, v);
        }
      }
      if (valid) out_v[i] = (accumulate ? out_v[i] : 0.f) + (v + sW2[H]);  // accumulate: second half of the units of a wider net
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // D / A of this group are free again after our loads
    }
  } else if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    const uint32_t sB_addr = smem_u32(sB);
    uint32_t it = 0;
    bool ok = true;
    for (int64_t t0 = (int64_t)blockIdx.x * 2; t0 < n_tiles && ok; t0 += (int64_t)gridDim.x * 2, ++it) {
      for (int g = 0; g < 2 && ok; ++g) {
        if (t0 + g >= n_tiles) break;
        if (!mbar_wait(smem_u32(&bars[g]), it & 1u)) {
          atomicExch(err, 2);
          ok = false;
          break;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tA = tmem_base + (uint32_t)(g * (A_COLS + D_COLS)), tD = tA + A_COLS;
#pragma unroll 1
        for (int s = 0; s < KSTEPS; ++s) {
#pragma unroll
          for (int j = 0; j < NSPLIT; ++j) {
            const uint32_t baddr = sB_addr + j * SPLIT_BYTES + s * 2 * KCHUNK_BYTES;
            // K-major, no swizzle: LBO = distance between the two 8-wide K chunks, SBO = distance between 8-row groups
            const uint64_t bdesc = (uint64_t)((baddr >> 4) & 0x3fffu) | ((uint64_t)(KCHUNK_BYTES >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
                                   (1ull << 46);
            const uint32_t acc = (s | j) ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                ::"r"(tD), "r"(tA + s * 8), "l"(bdesc), "r"(IDESC), "r"(acc), "r"(0u)
                : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[2 + g])) : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

int64_t eval_tc_image_bytes() { return ((int64_t)B_BYTES + (H + 1) * 4 + 255) / 256 * 256; }

int32_t eval_tc_prepare(const float* packed, int32_t H_src, int32_t unit0, uint8_t* img, cudaStream_t stream) {
  k_prepare_tc<<<104, 256, 0, stream>>>(packed, H_src, unit0, img);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_prepare_tc launch");
  return BG_OK;
}

int32_t eval_tc_launch(const EvalArgs& a, const uint8_t* img, int32_t* err_flag, cudaStream_t stream, int accumulate) {
  static DeviceOnce once;
  constexpr size_t smem = (size_t)B_BYTES + 1024 + 128;
  int32_t rc0 = once.run([&]() -> int32_t {
    uint32_t h[16][2];
    for (int n = 0; n < 16; ++n) {
      const float x = (float)((double)n / 15.0);
      const __half hi = __float2half_rn(x);
      const float r1 = x - __half2float(hi);
      const __half mid = __float2half_rn(r1);
      const __half lo = __float2half_rn(r1 - __half2float(mid));
      unsigned short uh, um, ul;
      memcpy(&uh, &hi, 2);
      memcpy(&um, &mid, 2);
      memcpy(&ul, &lo, 2);
      h[n][0] = (uint32_t)uh | ((uint32_t)um << 16);
      h[n][1] = (uint32_t)ul;
    }
    cudaError_t e = cudaMemcpyToSymbol(c_off15_split, h, sizeof(h));
    if (e != cudaSuccess) return check_cuda(e, "cudaMemcpyToSymbol(c_off15_split)");
    e = opt_in_shared(k_eval_tc<false>, smem);
    if (e == cudaSuccess) e = opt_in_shared(k_eval_tc<true>, smem + SCRATCH_BYTES);
    return check_cuda(e, "cudaFuncSetAttribute(k_eval_tc)");
  });
  if (rc0 != BG_OK) return rc0;
  const int64_t bound = a.N_dev ? a.max_N : a.N;
  int64_t want = (bound + 255) / 256;
  if (want < 1) want = 1;
  const int grid = (int)(want < NUM_SMS ? want : NUM_SMS);
  if (a.codes)
    k_eval_tc<true><<<grid, THREADS, smem + SCRATCH_BYTES, stream>>>(a.boards, a.flags, a.N, a.N_dev, a.max_N, img, a.out_v, err_flag, a.start_dev, accumulate,
                                                                     a.codes);
  else
    k_eval_tc<false><<<grid, THREADS, smem, stream>>>(a.boards, a.flags, a.N, a.N_dev, a.max_N, img, a.out_v, err_flag, a.start_dev, accumulate, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_eval_tc launch");
  return BG_OK;
}

}  // namespace bg
