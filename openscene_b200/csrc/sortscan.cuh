// Hand-written device sort / scan for the coordinate path (Morton sort of 64-bit keys with an int32 payload,
// run-head compaction, voxeliser unique).  Integer work on a few MB, L2-resident: launch count matters more
// than peak bandwidth.
//
//   radix_sort_pairs : stable LSD radix sort, 8-bit digits, over key bits [bit_lo, bit_hi) only (the callers know
//                      which bits vary), 3 kernels per pass: per-tile digit histogram -> exclusive scan of the
//                      digit-major [256][tiles] table -> stable scatter.  Inside a tile every warp owns a contiguous
//                      256-item chunk and ranks its items with __match_any_sync, so the order is (warp, round, lane)
//                      = input order; cross-warp offsets come from a 256-thread prefix over the 8 warp counters.
//   inclusive_scan_i32: reduce-then-scan over 2048-item tiles (3 kernels).
#pragma once
#include "common.cuh"

namespace osb {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                          // items per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // 2048 items per block
constexpr int RS_CHUNK = 32 * RS_ITEMS;              // 256 contiguous items per warp

static __global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint64_t *__restrict__ keys, int64_t n, int shift, int32_t *__restrict__ hist, int n_tiles) {
  __shared__ int32_t s_cnt[256];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&s_cnt[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = s_cnt[threadIdx.x];     // digit-major
}

// exclusive scan of `m` int32 values by one block (m = 256 * n_tiles, a few 10^4 .. 10^5).  A round covers 32 values per thread
// (eight 16-byte loads in flight per thread, 32768 values per round): the radix-sort histograms of the bench scenes
// (<= 128 tiles) take ONE round -- one load phase, one block scan, one store phase -- instead of four dependent ones.
static __global__ void __launch_bounds__(1024)
k_scan_single_block(int32_t *__restrict__ data, int64_t m) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
  for (int64_t base = 0; base < m; base += 32768) {
    const int64_t i0 = base + (int64_t)threadIdx.x * 32;
    int32_t v[32];
    if (vec && i0 + 32 <= m) {
      const int4 *p = reinterpret_cast<const int4 *>(data + i0);
#pragma unroll
      for (int q = 0; q < 8; ++q) { const int4 t = p[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (i0 + j < m) ? data[i0 + j] : 0;
    }
    int32_t tot = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) tot += v[j];
    int32_t x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;                                  // inclusive over warps
    }
    __syncthreads();
    const int32_t carry = s_carry;
    int32_t run = carry + (warp ? s_warp[warp - 1] : 0) + x - tot;     // exclusive prefix of this thread's first value
#pragma unroll
    for (int j = 0; j < 32; ++j) { const int32_t t = v[j]; v[j] = run; run += t; }
    if (vec && i0 + 32 <= m) {
      int4 *p = reinterpret_cast<int4 *>(data + i0);
#pragma unroll
      for (int q = 0; q < 8; ++q) p[q] = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (i0 + j < m) data[i0 + j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
}

static __global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint64_t *__restrict__ keys_in, const int32_t *__restrict__ vals_in, int64_t n, int shift,
             const int32_t *__restrict__ offs, int n_tiles, uint64_t *__restrict__ keys_out, int32_t *__restrict__ vals_out) {
  __shared__ int32_t s_wcnt[RS_WARPS][256];          // running per-warp digit counters, then exclusive over warps
  __shared__ int32_t s_base[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < RS_WARPS * 256; d += RS_THREADS) (&s_wcnt[0][0])[d] = 0;
  s_base[threadIdx.x] = offs[(int64_t)threadIdx.x * n_tiles + blockIdx.x];
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE + warp * RS_CHUNK;
  uint64_t key[RS_ITEMS];
  int32_t val[RS_ITEMS], rank[RS_ITEMS];
  const unsigned lt = (1u << lane) - 1;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = base + r * 32 + lane;
    const bool live = i < n;
    key[r] = live ? keys_in[i] : ~0ull;
    val[r] = live ? (vals_in ? vals_in[i] : (int32_t)i) : 0;
    const int d = live ? (int)((key[r] >> shift) & 255) : 256;     // dead lanes form their own group
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    rank[r] = live ? s_wcnt[warp][d & 255] + __popc(peers & lt) : -1;
    __syncwarp();                                                  // every lane has read the counter
    if (live && (peers & lt) == 0) s_wcnt[warp][d] += __popc(peers);   // group leader advances the warp's counter
    __syncwarp();
  }
  __syncthreads();
  {   // exclusive prefix over the warps of every digit (thread = digit)
    const int d = threadIdx.x;
    int32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) { const int32_t c = s_wcnt[w][d]; s_wcnt[w][d] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    if (rank[r] >= 0) {
      const int d = (int)((key[r] >> shift) & 255);
      const int64_t dst = (int64_t)s_base[d] + s_wcnt[warp][d] + rank[r];
      keys_out[dst] = key[r];
      vals_out[dst] = val[r];
    }
  }
}

// workspace: hist table + (nothing else); the caller provides ping-pong key/value buffers
static inline size_t radix_sort_ws_bytes(int64_t n) {
  const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
  return (size_t)(256 * n_tiles) * sizeof(int32_t) + 256;
}

// Sorts (keys, vals) by key bits [bit_lo, bit_hi).  vals_in == nullptr means payload = index.  The result ends up in
// (keys_b, vals_b) or (keys_a, vals_a); the function returns which: 0 -> a, 1 -> b, <0 on error.
// keys_a holds the input and is used as a ping-pong buffer (clobbered).
static inline int radix_sort_pairs(uint64_t *keys_a, int32_t *vals_a, uint64_t *keys_b, int32_t *vals_b,
                                   const int32_t *vals_in, int64_t n, int bit_lo, int bit_hi, void *ws, cudaStream_t stream) {
  const int n_tiles = (int)((n + RS_TILE - 1) / RS_TILE);
  int32_t *hist = reinterpret_cast<int32_t *>(ws);
  uint64_t *kin = keys_a, *kout = keys_b;
  int32_t *vin = vals_a, *vout = vals_b;
  const int32_t *vsrc = vals_in;
  int where = 0;
  bool first = true;
  for (int shift = bit_lo; shift < bit_hi || first; shift += 8) {
    k_rs_hist<<<n_tiles, RS_THREADS, 0, stream>>>(kin, n, shift, hist, n_tiles);
    k_scan_single_block<<<1, 1024, 0, stream>>>(hist, (int64_t)256 * n_tiles);
    k_rs_scatter<<<n_tiles, RS_THREADS, 0, stream>>>(kin, first ? vsrc : vin, n, shift, hist, n_tiles, kout, vout);
    count_launch(3);
    if (cudaGetLastError() != cudaSuccess) return -1;
    uint64_t *tk = kin; kin = kout; kout = tk;
    int32_t *tv = vin; vin = vout; vout = tv;
    where ^= 1;
    first = false;
  }
  return where;
}

// ------------------------------------------------------------------------------------------ scan
constexpr int SC_TILE = 2048;

static __global__ void __launch_bounds__(256)
k_sc_reduce(const int32_t *__restrict__ in, int64_t n, int32_t *__restrict__ tile_sums) {
  __shared__ int32_t s_w[8];
  const int64_t base = (int64_t)blockIdx.x * SC_TILE;
  int32_t acc = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    if (i < n) acc += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    tile_sums[blockIdx.x] = t;
  }
}

// inclusive scan inside a tile (thread owns 8 consecutive items) + exclusive tile offset
static __global__ void __launch_bounds__(256)
k_sc_scan(const int32_t *__restrict__ in, int64_t n, const int32_t *__restrict__ tile_offs, int32_t *__restrict__ out) {
  __shared__ int32_t s_w[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * SC_TILE + threadIdx.x * 8;
  int32_t v[8];
  int32_t run = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int64_t i = base + j; run += (i < n) ? in[i] : 0; v[j] = run; }
  int32_t x = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  int32_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_w[w];
  const int32_t off = tile_offs[blockIdx.x] + woff + x - run;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int64_t i = base + j; if (i < n) out[i] = off + v[j]; }
}

static inline size_t scan_ws_bytes(int64_t n) { return (size_t)((n + SC_TILE - 1) / SC_TILE) * sizeof(int32_t) + 256; }

static inline int inclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, void *ws, cudaStream_t stream) {
  const int n_tiles = (int)((n + SC_TILE - 1) / SC_TILE);
  int32_t *sums = reinterpret_cast<int32_t *>(ws);
  k_sc_reduce<<<n_tiles, 256, 0, stream>>>(in, n, sums);
  k_scan_single_block<<<1, 1024, 0, stream>>>(sums, n_tiles);
  k_sc_scan<<<n_tiles, 256, 0, stream>>>(in, n, sums, out);
  count_launch(3);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace osb
