// Shared helpers for libosb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/osb200.h"

namespace osb {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define OSB_CHECK(cond, ...)                                                        \
  do {                                                                              \
    if (!(cond)) { ::osb::set_error(__VA_ARGS__); return 1; }                       \
  } while (0)

#define OSB_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::osb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

#define OSB_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ::osb::count_launch();                                                          \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      ::osb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Coordinate keys.  Fields: x,y,z biased by 2^17 into 18 bits each, batch in the top 10 bits.
//   pack   = b<<54 | z<<36 | y<<18 | x        (hash key; neighbour = one 64-bit add)
//   morton = b<<54 | interleave(x,y,z)        (sort key; x is the least significant of each triple)
// ---------------------------------------------------------------------------------------------
constexpr int      kCoordBias  = 1 << 17;
constexpr int      kCoordLimit = (1 << 17) - 256;
constexpr uint64_t kEmptyKey   = 0xFFFFFFFFFFFFFFFFull;

struct __align__(16) HashSlot {
  unsigned long long key;
  int32_t            row;
  int32_t            pad;
};

__host__ __device__ inline uint64_t pack_key(int b, int x, int y, int z) {
  return ((uint64_t)(uint32_t)b << 54) | ((uint64_t)(uint32_t)(z + kCoordBias) << 36) |
         ((uint64_t)(uint32_t)(y + kCoordBias) << 18) | (uint64_t)(uint32_t)(x + kCoordBias);
}
// signed per-axis delta as one 64-bit addend (fields never borrow inside the valid range)
__host__ __device__ inline uint64_t pack_delta(int dx, int dy, int dz) {
  return (uint64_t)((int64_t)dx + ((int64_t)dy << 18) + ((int64_t)dz << 36));
}
__host__ __device__ inline uint64_t spread3(uint32_t v) {  // 18 bits -> every third bit
  uint64_t x = v & 0x3FFFFu;
  x = (x | (x << 32)) & 0x001F00000000FFFFull;
  x = (x | (x << 16)) & 0x001F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}
__host__ __device__ inline uint64_t morton_key(int b, int x, int y, int z) {
  return ((uint64_t)(uint32_t)b << 54) | spread3((uint32_t)(x + kCoordBias)) |
         (spread3((uint32_t)(y + kCoordBias)) << 1) | (spread3((uint32_t)(z + kCoordBias)) << 2);
}
__host__ __device__ inline uint64_t hash_u64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}

__device__ inline int hash_lookup(const HashSlot *__restrict__ slots, uint64_t mask, uint64_t key) {
  uint64_t s = hash_u64(key) & mask;
  while (true) {
    // 16-byte slot read in one transaction
    const int4 v = __ldg(reinterpret_cast<const int4 *>(slots + s));
    const uint64_t k = ((uint64_t)(uint32_t)v.y << 32) | (uint32_t)v.x;
    if (k == key) return v.z;
    if (k == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
}

// ---------------------------------------------------------------------------------------------
// split-fp32: v ~= hi + lo, both bf16 (round-to-nearest-even).  |v - hi - lo| <= 2^-17 |v|.
// ---------------------------------------------------------------------------------------------
__device__ inline void split_bf16(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ inline float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}
// byte offset of channel c (hi part) inside a split row; the lo part is +64 bytes
__host__ __device__ inline int split_off_hi(int c) { return (c >> 5) * 128 + (c & 31) * 2; }

}  // namespace osb
