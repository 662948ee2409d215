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
// * A never touches shared memory: a builder thread owns one TMEM lane (= one board), builds its 208-entry fp16 feature row in
//   registers and tcgen05.st's it into TMEM (A-from-TMEM "TS" MMA); B (2 x 128 x 208 fp16 = 104 KB, no-swizzle K-major core-matrix
//   layout) stays resident in shared memory; D is read back with tcgen05.ld and the sigmoid / w2 dot product / +b2 epilogue runs one
//   thread per board.
// * persistent CTA (one per SM), warp roles apart: 8 builder warps, 8 epilogue warps, one MMA-issuer thread, one tile-scheduler thread, two
//   TMEM slots [A0 | D0 | A1 | D1] (512 columns) handed round by mbarriers -- see the comment at k_eval_tc.  All waits are bounded (a
//   status word instead of a hang).  Development switches: BG_TC_STAMPS (clock64 phase stamps), BG_TC_EXP (energy-split experiments,
//   DESIGN.md 4.4), BG_TC_BUILDER_SETS.
#include <cuda_fp16.h>

#include <atomic>

#include "codes.cuh"
#include "eval.cuh"

namespace bg {

namespace {

constexpr int H = 128, KP = 208, KSTEPS = KP / 16, NSPLIT = 2;
constexpr int KCHUNK_BYTES = 16 * 128;                      // one 8-wide K chunk of all 128 rows: 16 N-groups x 128 B
constexpr int SPLIT_BYTES = (KP / 8) * KCHUNK_BYTES;        // 26 chunks = 53,248 B
constexpr int B_BYTES = NSPLIT * SPLIT_BYTES;               // 106,496 B
#ifndef BG_TC_EXP
#define BG_TC_EXP 0
#endif
#ifndef BG_TC_BUILDER_SETS
#define BG_TC_BUILDER_SETS 2
#endif
static_assert(BG_TC_BUILDER_SETS == 1 || BG_TC_BUILDER_SETS == 2, "a builder set may lag its slot's mbarrier by one phase only");
constexpr int NB = BG_TC_BUILDER_SETS;                      // builder warp sets (4 warps each); set b builds local tiles b, b + NB, ...
constexpr int EPI_WARP0 = 4 * NB, MMA_WARP = 4 * NB + 8;    // warps [0, 4 NB) builders, 8 epilogue warps, one MMA issuer (see k_eval_tc)
constexpr int SCHED_WARP = MMA_WARP + 1;                      // lane 0: tile scheduler (claims tile pairs from the grid-wide counter)
constexpr int THREADS = 32 * (SCHED_WARP + 1);
constexpr int RING = 32, RUNAHEAD = 24, END_PAD = 8;          // tile ring entries, scheduler run-ahead over the slowest epilogue, END entries
constexpr uint32_t TILE_END = 0xffffffffu;
constexpr int TMEM_COLS = 512, A_COLS = 128, D_COLS = 128;  // per slot: A at +0 (104 columns used), D at +128
constexpr int NUM_SMS = 148;
// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = F16 (format fields 0), N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__constant__ uint32_t c_off15_split[16][2];  // n/15 = hi + mid + lo in fp16: [n][0] = hi | mid << 16, [n][1] = lo

// -DBG_TC_STAMPS=1 (development builds only; see scripts/dev_tc_stamps.py): clock64 stamps of CTA 0's phases for the 64 local tiles
// starting at BG_TC_STAMP_FIRST -- [0] builder warp 0, [1] the epilogue warp of lane quarter 0 that owns the tile, [2] the MMA issuer
#ifdef BG_TC_STAMPS
#ifndef BG_TC_STAMP_FIRST
#define BG_TC_STAMP_FIRST 0
#endif
__device__ long long g_tc_stamps[3][64][8];
#define TC_STAMP(cond, who, it, slot)                                                                               \
  do {                                                                                                             \
    if ((cond) && blockIdx.x == 0 && (it) >= BG_TC_STAMP_FIRST && (it) < BG_TC_STAMP_FIRST + 64u)                  \
      g_tc_stamps[who][(it) - BG_TC_STAMP_FIRST][slot] = clock64();                                                \
  } while (0)
#else
#define TC_STAMP(cond, who, it, slot) \
  do {                                \
  } while (0)
#endif

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
// bounded wait: returns false on timeout (so a descriptor / protocol bug raises an error instead of hanging the GPU).  try_wait suspends the
// thread in hardware for up to ~1 us per call and wakes it when the phase completes (no polling back-off on the tile hand-offs)
__device__ __forceinline__ bool mbar_wait(uint32_t addr, uint32_t parity) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 20); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(1000u)
        : "memory");
    if (ok) return true;
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

// tcgen05.wait::ld that also names the 32 destination registers of the load it covers, so the compiler cannot schedule their
// consumers above the wait while another load is already in flight behind it
__device__ __forceinline__ void tmem_wait_ld32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                 "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                 "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                 "+r"(r[31])
               :
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

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2: one issue slot for two lanes of the epilogue's arithmetic) ----
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Position board + result code -> afterstate board (codes.cuh; same result as apply_code_bytes for every code the move generator emits),
// shaped for a builder thread: words 0..11 (the 48 points) go through this thread's shared-memory scratch row because the points a code touches
// are data dependent; word 12 (bars / borne-off counts) stays in a register.  The only state a sub-move reads -- "does the landing point hold
// exactly one opposing checker" -- is taken from the ORIGINAL board (a point is hit by the first checker that lands on it, later landings on
// the same point find it empty), so all reads are issued together and every change is an independent fire-and-forget shared-memory add of
// +-1 into the right byte: no load -> modify -> store chains (the byte-wise form spent 1,200-2,500 cycles per tile on them).
// DOUBLES = false: the caller knows that no row of the warp carries a double's code (two sub-moves, explicit destinations).
template <bool DOUBLES>
__device__ __forceinline__ void rebuild_afterstate(uint32_t* scr, const uint32_t (&raw)[13], uint32_t (&cur)[13], uint32_t code, uint32_t player) {
#pragma unroll
  for (int w = 0; w < 12; ++w) scr[w] = raw[w];
  const uint32_t own = player * 24u, opp = 24u - own;
  constexpr int NQ = DOUBLES ? 4 : 2;
  const bool dbl = DOUBLES && code_is_double(code);
  const int die = DOUBLES ? code_die(code) : 0;
  const int step = player == 0 ? die : -die;
  const uint32_t enter = player == 0 ? (uint32_t)(die - 1) : (uint32_t)(24 - die);
  uint32_t src[NQ], dst[NQ], oc[NQ];
  bool ok[NQ];
  bool alive = true;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    src[q] = (code >> (5 * q)) & 31u;
    alive = alive && src[q] != CODE_NONE && (dbl || q < 2);
    ok[q] = alive;
    const int ee = (int)src[q] + step;
    const uint32_t ed = src[q] == 24u ? enter : ((ee < 0 || ee > 23) ? 25u : (uint32_t)ee);
    dst[q] = dbl ? ed : ((code >> (10 + 5 * (q & 1))) & 31u);
  }
  const uint8_t* rb = reinterpret_cast<const uint8_t*>(scr);
#pragma unroll
  for (int q = 0; q < NQ; ++q) oc[q] = (ok[q] && dst[q] < 24u) ? rb[opp + dst[q]] : 0u;
  auto bump = [&](uint32_t byte, uint32_t delta) { atomicAdd(&scr[byte >> 2], delta << ((byte & 3u) * 8u)); };
  uint32_t d12 = 0, hits = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if (ok[q]) {
      if (src[q] == 24u)
        d12 -= 1u << (8u * player);  // from the bar
      else
        bump(own + src[q], 0xffffffffu);
      if (dst[q] < 24u) {
        bump(own + dst[q], 1u);
        bool first = true;
#pragma unroll
        for (int j = 0; j < q; ++j) first = first && !(ok[j] && dst[j] == dst[q]);
        if (oc[q] == 1u && first) {  // blot hit
          bump(opp + dst[q], 0xffffffffu);
          ++hits;
        }
      } else {
        d12 += 1u << (8u * (2u + player));  // borne off
      }
    }
  }
  if (!dbl) {
    const uint32_t in = (code >> 20) & 31u;
    if (in != CODE_NONE) {  // the blot on the point a checker only passed through
      bump(opp + in, 0xffffffffu);
      ++hits;
    }
  }
  d12 += hits << (8u * (1u - player));  // hit checkers go to the opponent's bar
#pragma unroll
  for (int w = 0; w < 12; ++w) cur[w] = scr[w];
  cur[12] = raw[12] + d12;
}

// Roles of the 4 NB + 10 = 18 warps of the persistent CTA (one CTA per SM; TMEM lane quarter of a warp = warp % 4 = its scheduler):
//   warps 0-7   BUILDERS, two sets of four (set b = local tiles b, b + 2, ...): thread (q, lane) owns row q * 32 + lane: a three-stage register
//               pipeline loads the row's (code, position) two of its tiles ahead, the position's board words one tile ahead, (CODES) rebuilds
//               the afterstate, and writes the fp16 feature row into TMEM operand A[slot] as soon as the MMAs that read it are done;
//   warps 8-15  EPILOGUE: warps 8-11 read accumulator D[0] (even local tiles), warps 12-15 D[1] (odd ones): sigmoid, w2 dot product, store;
//   warp 16     lane 0 issues the 26 tcgen05.mma of a tile when A[slot] is full and D[slot] has been drained;
//   warp 17     lane 0 is the tile scheduler of the dynamic schedule (idle under the static one; see tile_of below).
// Local tile n of a CTA uses slot n & 1; TMEM = [A0 | D0 | A1 | D1].  mbarriers: a_full[s] (128 builder arrivals), mma_done[s] (tcgen05.commit;
// frees A[s] for the builders and hands D[s] to its epilogue warps), d_free[s] (128 epilogue arrivals, made as soon as the accumulator sits in
// registers).  A role lags its slot's barrier by at most one phase (parity waits): hence exactly two builder sets.  With the roles apart the
// tensor pipe only waits when a builder or an epilogue falls behind a whole tile; in the round-1 form (the same threads built, waited and ran
// the epilogue) the next tile's fetch + rebuild sat on the critical path: clock64 stamps showed 2,300-4,000 of the 6,100 cycles of a
// two-tile period there (profiles/r02_eval_tc_phase_stamps.txt).
constexpr int SCRATCH_BYTES = NB * 128 * 13 * 4;
constexpr int BAR_OFF = B_BYTES + 1024, RING_OFF = BAR_OFF + 128, SCRATCH_OFF = RING_OFF + RING * 8 + 64;
constexpr size_t SMEM_BYTES = (size_t)SCRATCH_OFF + SCRATCH_BYTES;

// CODES: the rows are (code, position index) pairs of the compact pool (codes.cuh); `boards` / `flags` are then the POSITIONS' boards and
// players, and each builder thread rebuilds its afterstate in its shared-memory scratch row (rebuild_afterstate: 12 word stores, a few
// independent +-1 adds, 12 word loads) -- the afterstate boards never exist in HBM.
template <bool CODES>
__global__ void __launch_bounds__(THREADS, 1)
    k_eval_tc(const int8_t* __restrict__ boards, const uint8_t* __restrict__ flags, int64_t N_host, const int64_t* __restrict__ N_dev,
              int64_t max_N, const uint8_t* __restrict__ img, float* __restrict__ out_v, int32_t* __restrict__ err,
              const int64_t* __restrict__ start_dev, int accumulate, const uint2* __restrict__ codes, unsigned int* __restrict__ tile_ctr) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;                                                   // B operand image
  float* sW2 = reinterpret_cast<float*>(smem + B_BYTES);                // w2 as (w0, w2, w1, w3) per four units, then b2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);         // a_full[2], mma_done[2], d_free[2], b_loaded
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 64);
  volatile uint64_t* ring = reinterpret_cast<volatile uint64_t*>(smem + RING_OFF);               // [RING] (local tile index << 32) | grid tile
  volatile uint32_t* progress = reinterpret_cast<volatile uint32_t*>(smem + RING_OFF + RING * 8);  // [2] local tiles finished per slot (+1)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < RING) ring[tid] = ~0ull;
  if (tid < 2) progress[tid] = 0u;
  // the 2 x 4 fp16 thermometer features of TWO neighbouring points by their checker counts (c0 + 16 c1): one 16-byte table read per pair of
  // points instead of ~10 ALU instructions per point
  __shared__ uint4 s_xtab[256];
  if (tid < 256) {
    uint32_t a0, a1, b0, b1;
    point_words((uint32_t)tid & 15u, a0, a1);
    point_words((uint32_t)tid >> 4, b0, b1);
    s_xtab[tid] = make_uint4(a0, a1, b0, b1);
  }

  for (int i = tid; i <= H; i += THREADS) {
    const int j = i == H ? H : (i & ~3) | ((i & 1) << 1) | ((i & 2) >> 1);  // units (0, 1, 2, 3) of a group sit at (0, 2, 1, 3)
    sW2[j] = reinterpret_cast<const float*>(img + B_BYTES)[i];
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 128);
    mbar_init(smem_u32(&bars[1]), 128);
    mbar_init(smem_u32(&bars[2]), 1);
    mbar_init(smem_u32(&bars[3]), 1);
    mbar_init(smem_u32(&bars[4]), 128);
    mbar_init(smem_u32(&bars[5]), 128);
    mbar_init(smem_u32(&bars[6]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the 104 KB B operand image: two bulk copies by the TMA engine (one per weight piece), completion counted in bytes on bars[6]; they
    // run while the other threads build the tables and the MMA warp allocates TMEM
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bars[6])), "r"((uint32_t)B_BYTES) : "memory");
#pragma unroll
    for (int j = 0; j < NSPLIT; ++j)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB + j * SPLIT_BYTES)),
                   "l"(img + j * SPLIT_BYTES), "r"((uint32_t)SPLIT_BYTES), "r"(smem_u32(&bars[6]))
                   : "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // make the generic-proxy writes of B visible to the async (tensor-core) proxy, then publish the TMEM base
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // the B image has landed (async-proxy writes: visible to the tensor core without a proxy fence); on a timeout the status word is set and
  // the CTA runs through with zero rows (so that TMEM is still released below)
  const bool b_ok = mbar_wait(smem_u32(&bars[6]), 0u);
  if (!b_ok && tid == 0) atomicExch(err, 7);

  int64_t N = N_dev ? *N_dev : N_host;
  if (N > max_N) N = max_N;
  if (!b_ok) N = 0;
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
  // Two tile schedules.  STATIC (tile_ctr == nullptr): local tile n of CTA b is grid tile (n >> 1) * 2 * gridDim.x + 2 b + (n & 1).  DYNAMIC: the
  // scheduler lane claims PAIRS of neighbouring 128-row tiles from a grid-wide counter and publishes them, in order, as this CTA's local
  // tiles n = 0, 1, 2, ... in a shared-memory ring that every role reads.  Either way pairs of neighbouring tiles stay on one SM (compact
  // pools: neighbouring rows share their position's board, so its 13 words come out of L1).  The static split is ~4 % faster when the
  // evaluator owns the GPU (no ring reads in the builders' loop; 36.3 vs 37.9 ms at the full configuration); the dynamic one is for launches
  // that run NEXT TO other kernels -- a self-play ply's tail tiers, the learner's 8-CTA cluster in the training loop -- where CTAs that
  // cannot become resident would leave a static share undone until the other kernel ends (one 65,536-game ply 0.517 -> 0.495 ms).
  const bool dyn = tile_ctr != nullptr;
  const int64_t tstride = (int64_t)gridDim.x * 2, tfirst = (int64_t)blockIdx.x * 2;
  auto tile_of = [&](uint32_t n) -> int64_t {
    if (!dyn) return (int64_t)(n >> 1) * tstride + tfirst + (n & 1u);
    volatile uint64_t* e = &ring[n % RING];
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
      const uint64_t v = *e;
      if ((uint32_t)(v >> 32) == n) return (uint32_t)v == TILE_END ? n_tiles : (int64_t)(uint32_t)v;
      if (spin > 8) __nanosleep(40);
    }
    atomicExch(err, 5);  // the scheduler never published local tile n: surface it and stop
    return n_tiles;
  };

  if (warp < EPI_WARP0) {
    // ================= builders: one thread per TMEM lane / row =================
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t bset = (uint32_t)warp >> 2;
    const uint32_t tA_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t* b32 = reinterpret_cast<const uint32_t*>(boards);
    uint32_t* const scr = reinterpret_cast<uint32_t*>(smem + SCRATCH_OFF) + ((int)bset * 128 + row) * 13;  // CODES: this thread's scratch row (13-word stride)
    // software pipeline over local tiles: `cur` = afterstate words of tile n, `raw` = board words of tile n + 1 (loads in flight since the
    // previous iteration), `pos2` = the row -> position index of tile n + 2 (CODES: loaded one iteration before its board words)
    uint32_t cur[13], raw[13];
    uint32_t cur_flag = 0, raw_flag = 0, raw_code = 0, code2 = 0;
    int64_t pos2 = -1;  // -1: no such row
    bool cur_valid = false, raw_valid = false;
    bool idx_tile = false, raw_tile = false, cur_tile = false;  // does the staged local tile exist at all (loop condition)
    auto load_index = [&](uint32_t n) {  // stage 1: which board does row `row` of local tile n read?
      const int64_t t = tile_of(n), i = t * 128 + row;
      idx_tile = t < n_tiles;
      pos2 = -1;
      code2 = 0;
      if (t < n_tiles && i < N) {
        if constexpr (CODES) {
          const uint2 e = __ldg(codes + i);
          code2 = e.x;
          pos2 = (int64_t)e.y;
        } else {
          pos2 = i;
        }
      }
    };
    auto load_raw = [&]() {  // stage 2: issue the loads of the 13 board words + player flag of (pos2, code2)
      raw_valid = pos2 >= 0;
      raw_tile = idx_tile;
      raw_code = code2;
      if (raw_valid) {
#pragma unroll
        for (int w = 0; w < 13; ++w) raw[w] = __ldg(b32 + pos2 * 13 + w);
        raw_flag = flags[pos2] & 1u;
      } else {
#pragma unroll
        for (int w = 0; w < 13; ++w) raw[w] = 0u;
        raw_flag = 0u;
      }
    };
    auto finish_raw = [&]() {  // stage 3: raw -> cur (CODES: position board -> afterstate board in the scratch row)
      cur_valid = raw_valid;
      cur_tile = raw_tile;
      cur_flag = raw_flag;
      if constexpr (CODES) {
        // neighbouring rows mostly share their (position, roll): the four-sub-move form only runs for warps that hold a double's row
        const bool warp_has_double = __any_sync(BG_FULL, raw_valid && code_is_double(raw_code));
        if (raw_valid) {
          if (warp_has_double)
            rebuild_afterstate<true>(scr, raw, cur, raw_code, raw_flag);
          else
            rebuild_afterstate<false>(scr, raw, cur, raw_code, raw_flag);
        } else {
#pragma unroll
          for (int w = 0; w < 13; ++w) cur[w] = 0u;
        }
      } else {
#pragma unroll
        for (int w = 0; w < 13; ++w) cur[w] = raw[w];
      }
    };
    load_index(bset);
    load_raw();
    load_index(bset + NB);
    finish_raw();  // this set's first tile is ready
    load_raw();    // its second one in flight
    load_index(bset + 2 * NB);
    for (uint32_t n = bset; cur_tile; n += NB) {
      const uint32_t s = n & 1u, k = n >> 1;
      const uint32_t tA = tA_lane + s * (uint32_t)(A_COLS + D_COLS);
      TC_STAMP(lane == 0 && q == 0, 0, n, 0);
      if (k >= 1) {  // A[s] is free once the MMAs of local tile n - 2 are done
        if (!mbar_wait(smem_u32(&bars[2 + s]), (k - 1) & 1u)) {
          if (lane == 0) atomicExch(err, 1);
          break;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      TC_STAMP(lane == 0 && q == 0, 0, n, 1);
      // ---- this row's fp16 features, 8 TMEM columns (= one K16 step = four points) at a time ----
#pragma unroll
      for (int wd = 0; wd < 12; ++wd) {  // board word wd: 4 points -> 16 features -> 8 columns
        uint32_t r[8];
        const uint32_t w = cur[wd] & 0x0f0f0f0fu, t = w | (w >> 4);  // bytes 0 and 2 of t: c0 + 16 c1, c2 + 16 c3
        const uint4 p01 = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(s_xtab) + ((t << 4) & 0xff0u));
        const uint4 p23 = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(s_xtab) + ((t >> 12) & 0xff0u));
        r[0] = p01.x;
        r[1] = p01.y;
        r[2] = p01.z;
        r[3] = p01.w;
        r[4] = p23.x;
        r[5] = p23.y;
        r[6] = p23.z;
        r[7] = p23.w;
        tmem_st8(tA + wd * 8, r);
      }
      {
        const uint32_t ONE = 0x3c00u;
        const uint32_t w12 = cur[12];
        const uint32_t bar0 = w12 & 0xffu, bar1 = (w12 >> 8) & 0xffu, off0 = (w12 >> 16) & 15u, off1 = (w12 >> 24) & 15u;
        const uint32_t hb0 = __half_as_ushort(__float2half_rn((float)bar0 * 0.5f));
        const uint32_t hb1 = __half_as_ushort(__float2half_rn((float)bar1 * 0.5f));
        const uint32_t s0a = c_off15_split[off0][0], s0b = c_off15_split[off0][1];
        const uint32_t s1a = c_off15_split[off1][0], s1b = c_off15_split[off1][1];
        uint32_t r[8];
        r[0] = hb0 | ((s0a & 0xffffu) << 16);                       // 192 bar0/2, 193 off0 hi
        r[1] = hb1 | ((s1a & 0xffffu) << 16);                       // 194 bar1/2, 195 off1 hi
        r[2] = cur_valid ? (cur_flag ? ONE << 16 : ONE) : 0u;       // 196, 197 flag one-hot
        r[3] = (s0a >> 16) | (s0b << 16);                           // 198 off0 mid, 199 off0 lo
        r[4] = (s1a >> 16) | (s1b << 16);                           // 200 off1 mid, 201 off1 lo
        r[5] = cur_valid ? ONE : 0u;                                // 202 constant 1 (bias column), 203 = 0
        r[6] = 0u;
        r[7] = 0u;
        tmem_st8(tA + 96, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&bars[s]));
      TC_STAMP(lane == 0 && q == 0, 0, n, 2);
      finish_raw();  // this set's next tile: its loads were issued one iteration ago
      TC_STAMP(lane == 0 && q == 0, 0, n, 3);
      load_raw();    // the one after: its index was loaded one iteration ago
      TC_STAMP(lane == 0 && q == 0, 0, n, 4);
      load_index(n + 3 * NB);
      TC_STAMP(lane == 0 && q == 0, 0, n, 5);
    }
  } else if (warp < MMA_WARP) {
    // ================= epilogue: the first four warps drain D[0] (even local tiles), the other four D[1] =================
    const uint32_t s = (uint32_t)(warp - EPI_WARP0) >> 2;
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t tD = tmem_base + ((uint32_t)(q * 32) << 16) + s * (uint32_t)(A_COLS + D_COLS) + A_COLS;
    const uint32_t done = smem_u32(&bars[2 + s]), dfree = smem_u32(&bars[4 + s]);
    const float b2 = sW2[H];
    const u64 one2 = pack2(1.0f, 1.0f);
    for (uint32_t k = 0;; ++k) {
      const int64_t t = tile_of(2 * k + s);
      if (t >= n_tiles) break;
      const int64_t i = t * 128 + row;
      TC_STAMP(lane == 0 && q == 0, 1, 2 * k + s, 0);
      if (!mbar_wait(done, k & 1u)) {
        if (lane == 0) atomicExch(err, 3);
        break;
      }
      TC_STAMP(lane == 0 && q == 0, 1, 2 * k + s, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      u64 va = pack2(0.f, 0.f), vb = va;
      // 32 hidden units of this row at a time.  The accumulator holds y = -z log2(e) (scaled weights), a sigmoid is 1 / (1 + 2^y).  MUFU (ex2
      // and rcp share one 4-lane-per-scheduler pipe) and issue slots bound this loop, so FOUR sigmoids share one reciprocal -- with
      // (a, b, c, d) = 1 + 2^y, p = ab, q = cd, r = 1 / (pq):  1/a = b q r, 1/b = a q r, 1/c = d p r, 1/d = c p r -- and the adds /
      // multiplies / FMAs run as packed pairs.  y is clamped at 28.85 (z >= -20: sigmoid < 2.1e-9, far below the 1e-5 contract) so that
      // the product of four terms stays below 5.5e34.
      auto consume = [&](const uint32_t (&z)[32], int c0) {
#if BG_TC_EXP == 1
#pragma unroll
        for (int c = 0; c < 32; c += 2) va = add2(va, pack2(__uint_as_float(z[c]), __uint_as_float(z[c + 1])));
        return;
#endif
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float e0 = ex2_approx(fminf(__uint_as_float(z[c]), 28.853901f));
          const float e1 = ex2_approx(fminf(__uint_as_float(z[c + 1]), 28.853901f));
          const float e2 = ex2_approx(fminf(__uint_as_float(z[c + 2]), 28.853901f));
          const float e3 = ex2_approx(fminf(__uint_as_float(z[c + 3]), 28.853901f));
          const u64 ac = add2(pack2(e0, e2), one2), bd = add2(pack2(e1, e3), one2);
          const u64 pq = mul2(ac, bd);
          float p, qq;
          unpack2(pq, p, qq);
          const float r = rcp_approx(p * qq);
          const u64 inv = pack2(r * qq, r * p);  // (1/p, 1/q)
          const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(&sW2[c0 + c]);  // (w0, w2), (w1, w3): one broadcast 16-byte read
          va = fma2(mul2(bd, inv), w4.x, va);  // (1/a, 1/c): units c, c + 2
          vb = fma2(mul2(ac, inv), w4.y, vb);  // (1/b, 1/d): units c + 1, c + 3
        }
      };
      // the TMEM read of the next 32 columns is in flight while the current 32 are consumed
      uint32_t za[32], zb[32];
      tmem_ld32(tD, za);
      tmem_wait_ld32(za);
      tmem_ld32(tD + 32, zb);
      consume(za, 0);
      tmem_wait_ld32(zb);
      tmem_ld32(tD + 64, za);
      consume(zb, 32);
      tmem_wait_ld32(za);
      tmem_ld32(tD + 96, zb);
      consume(za, 64);
      tmem_wait_ld32(zb);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(dfree);  // the accumulator sits in registers: D[s] may be overwritten
      TC_STAMP(lane == 0 && q == 0, 1, 2 * k + s, 2);
      consume(zb, 96);
      float v0, v1, v2, v3;
      unpack2(va, v0, v1);
      unpack2(vb, v2, v3);
      const float v = (v0 + v1) + (v2 + v3);
      if (i < N) out_v[i] = (accumulate ? out_v[i] : 0.f) + (v + b2);  // accumulate: second half of the units of a wider net
      if (dyn && q == 0 && lane == 0) progress[s] = 2 * k + s + 1;  // flow control of the tile scheduler (the other quarters are within one tile of this one)
      TC_STAMP(lane == 0 && q == 0, 1, 2 * k + s, 3);
    }
  } else if (warp == MMA_WARP && lane == 0) {
    // ================= MMA issuer (one thread) =================
    const uint32_t sB_addr = smem_u32(sB);
    // K-major, no swizzle: LBO = distance between the two 8-wide K chunks, SBO = distance between 8-row groups
    const uint64_t bdesc0 = (uint64_t)((sB_addr >> 4) & 0x3fffu) | ((uint64_t)(KCHUNK_BYTES >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    for (uint32_t n = 0; tile_of(n) < n_tiles; ++n) {
      const uint32_t s = n & 1u, k = n >> 1;
      if (!mbar_wait(smem_u32(&bars[s]), k & 1u)) {  // A[s] written
        atomicExch(err, 2);
        break;
      }
      TC_STAMP(true, 2, n, 0);
      if (k >= 1 && !mbar_wait(smem_u32(&bars[4 + s]), (k - 1) & 1u)) {  // D[s] drained
        atomicExch(err, 4);
        break;
      }
      TC_STAMP(true, 2, n, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tA = tmem_base + s * (uint32_t)(A_COLS + D_COLS), tD = tA + A_COLS;
      // 26 dispatches, fully unrolled: every descriptor is bdesc0 plus a compile-time offset in its low word (no carry out of the 14-bit
      // address field: shared memory ends below 0x3900 << 4)
#pragma unroll
      for (int ks = 0; ks < (BG_TC_EXP == 2 ? 1 : KSTEPS); ++ks) {
#pragma unroll
        for (int j = 0; j < NSPLIT; ++j) {
          const uint64_t bdesc = bdesc0 + (uint64_t)((j * SPLIT_BYTES + ks * 2 * KCHUNK_BYTES) >> 4);
          if (ks == 0 && j == 0)  // the first dispatch overwrites the accumulator
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t}"
                ::"r"(tD), "r"(tA + ks * 8), "l"(bdesc), "r"(IDESC), "r"(0u)
                : "memory");
          else
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t}"
                ::"r"(tD), "r"(tA + ks * 8), "l"(bdesc), "r"(IDESC), "r"(0u)
                : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[2 + s])) : "memory");
      TC_STAMP(true, 2, n, 2);
    }
  }
  if (dyn && warp == SCHED_WARP && lane == 0) {
    // ================= tile scheduler (one thread) =================
    // publishes local tile n only while n < (tiles finished by the slower epilogue slot) + RUNAHEAD: the ring entry it overwrites, n - RING,
    // is then at least RING - RUNAHEAD tiles behind every reader.  After the end of the grid it publishes END_PAD END entries (the builders
    // look 3 NB tiles ahead) and stops.
    uint32_t n = 0, ends = 0;
    bool ended = false, ok = true;
    while (ok && ends < END_PAD) {
      uint32_t pair = 0;
      if (!ended) pair = atomicAdd(tile_ctr, 1u);
      for (uint32_t h = 0; h < 2 && ok; ++h) {
        const int64_t t = (int64_t)pair * 2 + h;
        const uint32_t tile = (!ended && t < n_tiles) ? (uint32_t)t : TILE_END;
        uint32_t spin = 0;
        while (true) {
          const uint32_t p0 = progress[0], p1 = progress[1];
          if (n < (p0 < p1 ? p0 : p1) + RUNAHEAD) break;
          if (++spin > (1u << 24)) {
            atomicExch(err, 6);
            ok = false;
            break;
          }
          __nanosleep(100);
        }
        if (!ok) break;
        ring[n % RING] = ((uint64_t)n << 32) | tile;
        if (tile == TILE_END) {
          ended = true;
          ++ends;
        }
        ++n;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace

#ifdef BG_TC_STAMPS
extern "C" int32_t bg_dev_tc_stamps(long long* host_out) { return (int32_t)cudaMemcpyFromSymbol(host_out, g_tc_stamps, sizeof(g_tc_stamps)); }
#endif

int64_t eval_tc_image_bytes() { return ((int64_t)B_BYTES + (H + 1) * 4 + 255) / 256 * 256; }

int32_t eval_tc_prepare(const float* packed, int32_t H_src, int32_t unit0, uint8_t* img, cudaStream_t stream) {
  k_prepare_tc<<<104, 256, 0, stream>>>(packed, H_src, unit0, img);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_prepare_tc launch");
  return BG_OK;
}

// grid-wide tile counters: one per launch in flight, handed out round robin from a per-device pool and zeroed on the launch's stream
constexpr int N_TILE_CTRS = 1024;
static unsigned int* g_tile_ctrs[64] = {};
static std::atomic<uint32_t> g_tile_ctr_next{0};
static std::atomic<int32_t> g_tile_mode{-1};

int32_t eval_tc_tile_schedule(int32_t mode) { return g_tile_mode.exchange(mode < 0 ? -1 : (mode ? 1 : 0)); }

int32_t eval_tc_launch(const EvalArgs& a, const uint8_t* img, int32_t* err_flag, cudaStream_t stream, int accumulate) {
  static DeviceOnce once;
  constexpr size_t smem = SMEM_BYTES;
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
    if (e == cudaSuccess) e = opt_in_shared(k_eval_tc<true>, smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(k_eval_tc)");
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e == cudaSuccess && (dev < 0 || dev >= 64)) {
      set_error("k_eval_tc: device index %d out of range", dev);
      return BG_ERR_ARG;
    }
    if (e == cudaSuccess) e = cudaMalloc(&g_tile_ctrs[dev], N_TILE_CTRS * sizeof(unsigned int));
    return check_cuda(e, "cudaMalloc(tile counters)");
  });
  if (rc0 != BG_OK) return rc0;
  int dev = 0;
  cudaError_t e0 = cudaGetDevice(&dev);
  if (e0 != cudaSuccess) return check_cuda(e0, "cudaGetDevice");
  // dynamic tile schedule for launches that share the GPU with other kernels (self-play plies, the training loop), static for bulk passes
  const int32_t mode = a.dynamic_tiles >= 0 ? a.dynamic_tiles : g_tile_mode.load();
  const bool dynamic = mode < 0 ? (a.N_dev ? a.max_N : a.N) <= EVAL_TC_DYNAMIC_MAX_ROWS : mode != 0;
  unsigned int* ctr = nullptr;
  if (dynamic) {
    ctr = g_tile_ctrs[dev] + (g_tile_ctr_next.fetch_add(1) % N_TILE_CTRS);
    e0 = cudaMemsetAsync(ctr, 0, sizeof(unsigned int), stream);
    if (e0 != cudaSuccess) return check_cuda(e0, "cudaMemsetAsync(tile counter)");
  }
  const int64_t bound = a.N_dev ? a.max_N : a.N;
  int64_t want = (bound + 255) / 256;
  if (want < 1) want = 1;
  const int grid = (int)(want < NUM_SMS ? want : NUM_SMS);
  if (a.codes)
    k_eval_tc<true><<<grid, THREADS, smem, stream>>>(a.boards, a.flags, a.N, a.N_dev, a.max_N, img, a.out_v, err_flag, a.start_dev, accumulate,
                                           a.codes, ctr);
  else
    k_eval_tc<false><<<grid, THREADS, smem, stream>>>(a.boards, a.flags, a.N, a.N_dev, a.max_N, img, a.out_v, err_flag, a.start_dev, accumulate, nullptr, ctr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "k_eval_tc launch");
  return BG_OK;
}

}  // namespace bg
