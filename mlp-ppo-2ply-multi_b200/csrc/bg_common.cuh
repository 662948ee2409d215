// bg_common.cuh -- shared device helpers for libbgarena (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/bgarena.h"

#define BG_WARP 32
#define BG_FULL 0xffffffffu

namespace bg {

// thread-local error string (host side)
void set_error(const char* fmt, ...);
int32_t check_cuda(cudaError_t e, const char* what);

// One-time initialisation of PER-DEVICE state (cudaFuncSetAttribute opt-ins, __constant__ tables): runs `init` once for every device the
// library is used on, thread safe.  (A process-global flag would leave a second device without its shared-memory opt-in / constant table.)
struct DeviceOnce {
  std::mutex mu;
  uint64_t done = 0;  // bit d: initialised on device d
  template <class F>
  int32_t run(F&& init) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice");
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && ((done >> dev) & 1ull)) return BG_OK;
    const int32_t rc = init();
    if (rc == BG_OK && dev >= 0 && dev < 64) done |= 1ull << dev;
    return rc;
  }
};

// Opt a kernel in to `bytes` of dynamic shared memory and ask for the maximum shared-memory carve-out.  Every large-shared-memory kernel of
// the library uses the same carve-out: kernels that prefer different L1 / shared splits cannot be resident on an SM at the same time, so a
// tail-tier CTA (2 x 82 KB) used to wait for the whole persistent evaluator (1 x 107 KB per SM, on the side stream) to drain.
template <class K>
inline cudaError_t opt_in_shared(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  return e;
}

// Makes `dev` current for the lifetime of the guard and restores the caller's device afterwards (handle-based entry points must not
// leave the calling thread on another device).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// expand 4 nibbles (low 16 bits) to 4 bytes
__device__ __forceinline__ uint32_t nib4_to_bytes(uint32_t x) {
  return (x & 0xfu) | ((x & 0xf0u) << 4) | ((x & 0xf00u) << 8) | ((x & 0xf000u) << 12);
}
// pack 4 bytes (each < 16) to 4 nibbles
__device__ __forceinline__ uint32_t bytes_to_nib4(uint32_t x) {
  return (x & 0xfu) | ((x >> 4) & 0xf0u) | ((x >> 8) & 0xf00u) | ((x >> 12) & 0xf000u);
}
// spread 4 bits to the LSB of 4 bytes
__device__ __forceinline__ uint32_t bits4_to_bytes(uint32_t h) {
  return (h & 1u) | ((h & 2u) << 7) | ((h & 4u) << 14) | ((h & 8u) << 21);
}

// word `lane` (0..12) of the start position, reference immutable_board.py:26-70: P1 {0:2, 11:5, 16:3, 18:5}, P2 {23:2, 12:5, 7:3, 5:5}
__device__ __forceinline__ uint32_t initial_board_word(int lane) {
  uint32_t v = 0;
  if (lane == 0) v = 2u;                          // p0[0] = 2
  if (lane == 2) v = 5u << 24;                    // p0[11] = 5
  if (lane == 4) v = 3u | (5u << 16);             // p0[16] = 3, p0[18] = 5
  if (lane == 6 + 1) v = (5u << 8) | (3u << 24);  // p1[5] = 5, p1[7] = 3
  if (lane == 6 + 3) v = 5u;                      // p1[12] = 5
  if (lane == 6 + 5) v = 2u << 24;                // p1[23] = 2
  return v;
}

__device__ __forceinline__ uint32_t mix32(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t h = a * 0x9E3779B1u;
  h = (h ^ (h >> 15)) + b * 0x85EBCA77u;
  h = (h ^ (h >> 13)) + c * 0xC2B2AE3Du;
  h = (h ^ (h >> 16)) + d * 0x27D4EB2Fu;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based
struct Philox {
  static __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
  }
  static __device__ __forceinline__ void gen(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    out[0] = (uint32_t)ctr_lo;
    out[1] = (uint32_t)(ctr_lo >> 32);
    out[2] = (uint32_t)ctr_hi;
    out[3] = (uint32_t)(ctr_hi >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(out, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
  }
};

}  // namespace bg
