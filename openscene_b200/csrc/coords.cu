// Coordinate sets, hash tables and kernel maps (integer work, HBM/L2 bound).
//
// Replaces the coordinate manager inside MinkowskiEngine that the reference drives through
// ME.SparseTensor(...) (run/evaluate.py:284) and the strided / 3x3x3 / 5x5x5 convolutions of
// models/mink_unet.py:47-113.  Row order inside the library is Morton order so that a tile of
// consecutive rows is a compact surface patch whose 27-neighbourhoods overlap (gather locality).
#include "common.cuh"

#include "sortscan.cuh"

#include <stdarg.h>
#include <string.h>
#include <algorithm>

namespace osb {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__device__ __forceinline__ int floordiv(int a, int s) { return a >= 0 ? a / s : -((-a + s - 1) / s); }

// ------------------------------------------------------------------------------------ kernels
__global__ void k_morton_from_coords(const int4 *__restrict__ coords, int64_t n, int32_t new_ts,
                                     uint64_t *__restrict__ morton, int32_t *__restrict__ idx,
                                     int32_t *__restrict__ status, unsigned long long *__restrict__ bits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long k_or = 0ull, k_and = ~0ull;
  if (i < n) {
  int4 c = coords[i];  // (b, x, y, z)
  bool bad = c.x < 0 || c.x >= 1024 || abs(c.y) >= kCoordLimit || abs(c.z) >= kCoordLimit || abs(c.w) >= kCoordLimit;
  if (bad) { atomicOr(status, 1); c = make_int4(0, 0, 0, 0); }
  if (new_ts > 1) {
    c.y = floordiv(c.y, new_ts) * new_ts;
    c.z = floordiv(c.z, new_ts) * new_ts;
    c.w = floordiv(c.w, new_ts) * new_ts;
  }
  const uint64_t key = morton_key(c.x, c.y, c.z, c.w);
  morton[i] = key;
  idx[i] = (int32_t)i;
  k_or = key; k_and = key;
  }
  // bits that differ between keys = OR & ~AND: the radix sort only needs those digit positions
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    k_or |= __shfl_xor_sync(0xffffffffu, k_or, o);
    k_and &= __shfl_xor_sync(0xffffffffu, k_and, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicOr(bits, k_or); atomicAnd(bits + 1, k_and); }
}

// level 0: permute coordinates into Morton order, build inverse permutation, flag duplicates
__global__ void k_permute_coords(const int4 *__restrict__ coords, const int32_t *__restrict__ perm,
                                 const uint64_t *__restrict__ morton_sorted, int64_t n,
                                 int4 *__restrict__ coords_int, int32_t *__restrict__ inv_perm,
                                 int32_t *__restrict__ status) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int32_t p = perm[r];
  coords_int[r] = coords[p];
  inv_perm[p] = (int32_t)r;
  if (r > 0 && morton_sorted[r] == morton_sorted[r - 1]) atomicOr(status, 2);
}

__global__ void k_hash_clear(HashSlot *__restrict__ slots, int64_t cap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  reinterpret_cast<int4 *>(slots)[i] = make_int4(-1, -1, -1, 0);
}

__global__ void k_hash_insert(const int4 *__restrict__ coords_int, int64_t n, HashSlot *__restrict__ slots,
                              uint64_t mask) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int4 c = coords_int[r];
  const uint64_t key = pack_key(c.x, c.y, c.z, c.w);
  uint64_t s = hash_u64(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&slots[s].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == kEmptyKey || prev == key) { slots[s].row = (int32_t)r; return; }
    s = (s + 1) & mask;
  }
}

// stride: heads of runs of equal parent key in the sorted order
__global__ void k_run_heads(const uint64_t *__restrict__ key_sorted, int64_t n, int32_t *__restrict__ heads) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  heads[j] = (j == 0 || key_sorted[j] != key_sorted[j - 1]) ? 1 : 0;
}

__global__ void k_emit_coarse(const int4 *__restrict__ coords_fine, const int32_t *__restrict__ order,
                              const int32_t *__restrict__ heads, const int32_t *__restrict__ ids, int64_t n,
                              int32_t new_ts, int4 *__restrict__ coords_coarse, int32_t *__restrict__ parent_of) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t child = order[j];
  const int32_t id = ids[j] - 1;  // inclusive scan of heads
  parent_of[child] = id;
  if (heads[j]) {
    int4 c = coords_fine[child];
    c.y = floordiv(c.y, new_ts) * new_ts;
    c.z = floordiv(c.z, new_ts) * new_ts;
    c.w = floordiv(c.w, new_ts) * new_ts;
    coords_coarse[id] = c;
  }
}

// ---- sort-free stride for power-of-two tensor strides (children are Morton sorted => parents are too) ----
// n_fine lives on the device (n_dev); launches cover the upper bound n_upper.
__global__ void k_pyr_heads(const int4 *__restrict__ coords_fine, const int32_t *__restrict__ n_dev, int64_t n_upper,
                            int32_t new_ts, int32_t *__restrict__ heads) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_upper) return;
  const int32_t n = *n_dev;
  int h = 0;
  if (i < n) {
    const int m = ~(new_ts - 1);
    const int4 c = coords_fine[i];
    if (i == 0) h = 1;
    else {
      const int4 d = coords_fine[i - 1];
      h = (c.x != d.x) || ((c.y & m) != (d.y & m)) || ((c.z & m) != (d.z & m)) || ((c.w & m) != (d.w & m));
    }
  }
  heads[i] = h;
}

__global__ void k_pyr_emit(const int4 *__restrict__ coords_fine, const int32_t *__restrict__ n_dev, int64_t n_upper,
                           const int32_t *__restrict__ heads, const int32_t *__restrict__ ids, int32_t new_ts,
                           int4 *__restrict__ coords_coarse, int32_t *__restrict__ parent_of, int32_t *__restrict__ n_coarse_dev) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_upper) return;
  const int32_t n = *n_dev;
  if (i >= n) return;
  const int32_t id = ids[i] - 1;
  parent_of[i] = id;
  if (heads[i]) {
    const int m = ~(new_ts - 1);               // two's complement: clearing low bits == floor to a multiple of new_ts
    int4 c = coords_fine[i];
    c.y &= m; c.z &= m; c.w &= m;
    coords_coarse[id] = c;
  }
  if (i == n - 1) *n_coarse_dev = id + 1;
}

// kernel map: grid.y = k; one thread per output row
__global__ void k_kernel_map(const int4 *__restrict__ coords_out, int64_t n_out, const HashSlot *__restrict__ slots,
                             uint64_t mask, int ksx, int ksy, int ksz, int step, int32_t *__restrict__ nbr,
                             int32_t *__restrict__ pairs_per_k) {
  const int k = blockIdx.y;
  int ix = k % ksx, iy = (k / ksx) % ksy, iz = k / (ksx * ksy);
  const int dx = ((ksx & 1) ? ix - ksx / 2 : ix) * step;
  const int dy = ((ksy & 1) ? iy - ksy / 2 : iy) * step;
  const int dz = ((ksz & 1) ? iz - ksz / 2 : iz) * step;
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  if (o < n_out) {
    int4 c = coords_out[o];
    const uint64_t key = pack_key(c.x, c.y + dx, c.z + dy, c.w + dz);
    const int row = hash_lookup(slots, mask, key);
    nbr[(int64_t)k * n_out + o] = row;
    hit = row >= 0;
  }
  if (pairs_per_k != nullptr) {
    int cnt = __syncthreads_count(hit);
    if (threadIdx.x == 0 && cnt) atomicAdd(pairs_per_k + k, cnt);
  }
}

// ---- occupancy grid (common.cuh): build + kernel map through it ----
__global__ void k_occgrid_build(const int4 *__restrict__ coords, int64_t n, OccGridView g, unsigned long long *__restrict__ bitmap,
                                int32_t *__restrict__ first_row, int32_t *__restrict__ status) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int4 c = coords[r];
  const uint32_t lim = 1u << g.nbits;
  if ((uint32_t)(c.y >> g.log2_ts) >= lim || (uint32_t)(c.z >> g.log2_ts) >= lim || (uint32_t)(c.w >> g.log2_ts) >= lim ||
      (uint32_t)c.x >= (uint32_t)g.n_batch) {
    atomicOr(status, 1);                                  // the caller sized the grid from the key bits: cannot happen
    return;
  }
  int64_t word; int bit;
  occgrid_cell(g, c.x, c.y, c.z, c.w, word, bit);
  atomicOr(bitmap + word, 1ull << bit);
  bool first = r == 0;
  if (!first) {                                           // rows are Morton sorted: a word's rows are contiguous
    const int4 d = coords[r - 1];
    int64_t pw; int pb;
    occgrid_cell(g, d.x, d.y, d.z, d.w, pw, pb);
    first = pw != word;
  }
  if (first) first_row[word] = (int32_t)r;
}

__global__ void k_kernel_map_grid(const int4 *__restrict__ coords_out, int64_t n_out, OccGridView g, int ksx, int ksy, int ksz,
                                  int step, int32_t *__restrict__ nbr, int32_t *__restrict__ pairs_per_k) {
  const int k = blockIdx.y;
  int ix = k % ksx, iy = (k / ksx) % ksy, iz = k / (ksx * ksy);
  const int dx = ((ksx & 1) ? ix - ksx / 2 : ix) * step;
  const int dy = ((ksy & 1) ? iy - ksy / 2 : iy) * step;
  const int dz = ((ksz & 1) ? iz - ksz / 2 : iz) * step;
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  if (o < n_out) {
    const int4 c = coords_out[o];
    const int row = occgrid_lookup(g, c.x, c.y + dx, c.z + dy, c.w + dz);
    nbr[(int64_t)k * n_out + o] = row;
    hit = row >= 0;
  }
  if (pairs_per_k != nullptr) {
    int cnt = __syncthreads_count(hit);
    if (threadIdx.x == 0 && cnt) atomicAdd(pairs_per_k + k, cnt);
  }
}

__global__ void k_fill_i32(int32_t *__restrict__ p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_kernel_map_transpose(const int32_t *__restrict__ nbr, int64_t n_out, int K,
                                       int32_t *__restrict__ nbr_t, int64_t n_in) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out * K) return;
  const int k = (int)(t / n_out);
  const int64_t o = t - (int64_t)k * n_out;
  const int32_t i = nbr[t];
  if (i >= 0) nbr_t[(int64_t)k * n_in + i] = (int32_t)o;
}

// ------------------------------------------------------------------------------ workspace carve
struct Carver {
  char *p;
  size_t left;
  bool ok = true;
  template <typename T>
  T *take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    if (bytes > left) { ok = false; return nullptr; }
    T *r = reinterpret_cast<T *>(p);
    p += bytes;
    left -= bytes;
    return r;
  }
};

static size_t cub_sort_bytes(int64_t n) { return radix_sort_ws_bytes(n); }

// digit range [lo, hi) of the bits in which the keys differ (OR & ~AND), rounded to the 8-bit radix passes
static void varying_bits(unsigned long long k_or, unsigned long long k_and, int *lo, int *hi) {
  const unsigned long long diff = k_or & ~k_and;
  if (diff == 0) { *lo = 0; *hi = 8; return; }
  int l = 0, h = 64;
  while (!((diff >> l) & 1ull)) ++l;
  while (!((diff >> (h - 1)) & 1ull)) --h;
  *lo = l; *hi = h;
}

static size_t cub_scan_bytes(int64_t n) { return scan_ws_bytes(n); }

// status_host[2..5] = OR (lo, hi 32 bits) and AND (lo, hi) of the Morton keys: the caller derives from them whether all
// coordinates are non-negative, how many bits the largest one has and the largest batch index (occupancy-grid sizing)
static void key_bits_to_status(const unsigned long long hb[2], int32_t *status_host) {
  status_host[2] = (int32_t)(uint32_t)(hb[0] & 0xffffffffull);
  status_host[3] = (int32_t)(uint32_t)(hb[0] >> 32);
  status_host[4] = (int32_t)(uint32_t)(hb[1] & 0xffffffffull);
  status_host[5] = (int32_t)(uint32_t)(hb[1] >> 32);
}

// Sort (key, index) pairs over bits [lo, hi) such that the sorted keys land in *keys_sorted and the permutation in `perm`.
// `k0` holds the keys (clobbered), `k1` / `v_tmp` are scratch.
static int sort_to(uint64_t *k0, uint64_t *k1, int32_t *v_tmp, int32_t *perm, int64_t n, int lo, int hi, void *ws,
                   cudaStream_t stream, uint64_t **keys_sorted) {
  const int passes = std::max(1, (hi - lo + 7) / 8);
  // the result of an odd number of passes lands in the 'b' buffers
  int32_t *va = (passes & 1) ? v_tmp : perm, *vb = (passes & 1) ? perm : v_tmp;
  const int where = radix_sort_pairs(k0, va, k1, vb, nullptr, n, lo, hi, ws, stream);
  if (where < 0) return 1;
  *keys_sorted = where ? k1 : k0;
  return 0;
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_version(void) { return OSB_VERSION; }
const char *osb_last_error(void) { return osb::g_err; }
int64_t osb_launch_count(void) { return osb::g_launches.load(); }

int osb_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0;
  OSB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  OSB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

size_t osb_coordset_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  size_t per = (size_t)n * (8 + 8 + 4 + 4 + 4 + 4) + 8 * 256;
  return per + cub_sort_bytes(n) + cub_scan_bytes(n) + 1024;
}

int osb_hash_build(const int32_t *coords_int, int64_t n, void *slots, int64_t cap, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(cap >= 2 && (cap & (cap - 1)) == 0 && cap >= 2 * n, "osb_hash_build: cap=%lld must be a power of two >= 2n (n=%lld)",
            (long long)cap, (long long)n);
  k_hash_clear<<<(unsigned)ceil_div(cap, 256), 256, 0, stream>>>((HashSlot *)slots, cap);
  OSB_LAUNCH_CHECK();
  if (n > 0) {
    k_hash_insert<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>((const int4 *)coords_int, n, (HashSlot *)slots,
                                                                  (uint64_t)cap - 1);
    OSB_LAUNCH_CHECK();
  }
  return 0;
}

int osb_coordset_build(const int32_t *coords, int64_t n, int32_t *coords_int, int32_t *perm, int32_t *inv_perm,
                       void *slots, int64_t cap, int32_t *status_host, void *ws, size_t ws_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(n > 0 && n < (1ll << 31) - 1024, "osb_coordset_build: n=%lld out of range", (long long)n);
  Carver cv{(char *)ws, ws_bytes};
  uint64_t *morton = cv.take<uint64_t>(n);
  uint64_t *morton_s = cv.take<uint64_t>(n);
  int32_t *idx = cv.take<int32_t>(n);
  int32_t *status = cv.take<int32_t>(64);
  size_t sort_bytes = cub_sort_bytes(n);
  void *cub_tmp = cv.take<char>(sort_bytes);
  OSB_CHECK(cv.ok, "osb_coordset_build: workspace too small (%zu bytes)", ws_bytes);
  const unsigned long long bits_init[2] = {0ull, ~0ull};
  unsigned long long *bits = reinterpret_cast<unsigned long long *>(status + 8);
  OSB_CUDA(cudaMemsetAsync(status, 0, 8, stream));
  OSB_CUDA(cudaMemcpyAsync(bits, bits_init, 16, cudaMemcpyHostToDevice, stream));
  const unsigned nb = (unsigned)ceil_div(n, 256);
  k_morton_from_coords<<<nb, 256, 0, stream>>>((const int4 *)coords, n, 1, morton, idx, status, bits);
  OSB_LAUNCH_CHECK();
  unsigned long long hb[2];
  OSB_CUDA(cudaMemcpyAsync(hb, bits, 16, cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaStreamSynchronize(stream));
  int lo, hi;
  varying_bits(hb[0], hb[1], &lo, &hi);
  uint64_t *sorted_keys = nullptr;
  OSB_CHECK(sort_to(morton, morton_s, idx, perm, n, lo, hi, cub_tmp, stream, &sorted_keys) == 0, "osb_coordset_build: sort failed");
  k_permute_coords<<<nb, 256, 0, stream>>>((const int4 *)coords, perm, sorted_keys, n, (int4 *)coords_int, inv_perm, status);
  OSB_LAUNCH_CHECK();
  if (slots != nullptr && osb_hash_build(coords_int, n, slots, cap, stream_)) return 1;
  OSB_CUDA(cudaMemcpyAsync(status_host, status, 8, cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaStreamSynchronize(stream));
  key_bits_to_status(hb, status_host);
  return 0;
}

int osb_coordset_stride(const int32_t *coords_fine, int64_t n, int32_t new_ts, int32_t *coords_coarse,
                        int32_t *parent_of, int64_t *n_coarse_host, void *ws, size_t ws_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(n > 0 && new_ts >= 1, "osb_coordset_stride: bad arguments n=%lld new_ts=%d", (long long)n, new_ts);
  Carver cv{(char *)ws, ws_bytes};
  uint64_t *key = cv.take<uint64_t>(n);
  uint64_t *key_s = cv.take<uint64_t>(n);
  int32_t *idx = cv.take<int32_t>(n);
  int32_t *order = cv.take<int32_t>(n);
  int32_t *heads = cv.take<int32_t>(n);
  int32_t *ids = cv.take<int32_t>(n);
  int32_t *status = cv.take<int32_t>(64);
  size_t sort_bytes = cub_sort_bytes(n), scan_bytes = cub_scan_bytes(n);
  void *cub_tmp = cv.take<char>(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
  OSB_CHECK(cv.ok, "osb_coordset_stride: workspace too small (%zu bytes)", ws_bytes);
  const unsigned long long bits_init[2] = {0ull, ~0ull};
  unsigned long long *bits = reinterpret_cast<unsigned long long *>(status + 8);
  OSB_CUDA(cudaMemsetAsync(status, 0, 8, stream));
  OSB_CUDA(cudaMemcpyAsync(bits, bits_init, 16, cudaMemcpyHostToDevice, stream));
  const unsigned nb = (unsigned)ceil_div(n, 256);
  k_morton_from_coords<<<nb, 256, 0, stream>>>((const int4 *)coords_fine, n, new_ts, key, idx, status, bits);
  OSB_LAUNCH_CHECK();
  unsigned long long hb[2];
  OSB_CUDA(cudaMemcpyAsync(hb, bits, 16, cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaStreamSynchronize(stream));
  int lo, hi;
  varying_bits(hb[0], hb[1], &lo, &hi);
  uint64_t *sorted_keys = nullptr;
  OSB_CHECK(sort_to(key, key_s, idx, order, n, lo, hi, cub_tmp, stream, &sorted_keys) == 0, "osb_coordset_stride: sort failed");
  k_run_heads<<<nb, 256, 0, stream>>>(sorted_keys, n, heads);
  OSB_LAUNCH_CHECK();
  OSB_CHECK(inclusive_scan_i32(heads, ids, n, cub_tmp, stream) == 0, "osb_coordset_stride: scan failed");
  k_emit_coarse<<<nb, 256, 0, stream>>>((const int4 *)coords_fine, order, heads, ids, n, new_ts, (int4 *)coords_coarse,
                                        parent_of);
  OSB_LAUNCH_CHECK();
  int32_t last = 0;
  OSB_CUDA(cudaMemcpyAsync(&last, ids + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaStreamSynchronize(stream));
  *n_coarse_host = last;
  return 0;
}

// Whole stride-2 pyramid in one call: tensor stride 1 set (Morton sort, permutation, hash table) and `n_levels`
// coarser sets with tensor strides 2, 4, ... 2^n_levels, without host round trips between levels.
//   coords_lvl   out  int32 [n_levels][n,4]  coarse coordinate sets (upper-bound sized), internal order
//   parent_lvl   out  int32 [n_levels][n]    parent_lvl[l][r] = row of fine row r (level l) in level l+1
//   n_host       out  int64 [n_levels+1] HOST: rows per level
// SYNC twice (sort bit range; counts + status).
int osb_coordset_pyramid(const int32_t *coords, int64_t n, int32_t n_levels, int32_t *coords_int, int32_t *perm,
                         int32_t *inv_perm, void *slots, int64_t cap, int32_t *coords_lvl, int32_t *parent_lvl,
                         int64_t *n_host, int32_t *status_host, void *ws, size_t ws_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(n > 0 && n < (1ll << 31) - 1024 && n_levels >= 0 && n_levels <= 8, "osb_coordset_pyramid: bad arguments");
  Carver cv{(char *)ws, ws_bytes};
  uint64_t *morton = cv.take<uint64_t>(n);
  uint64_t *morton_s = cv.take<uint64_t>(n);
  int32_t *idx = cv.take<int32_t>(n);
  int32_t *heads = cv.take<int32_t>(n);
  int32_t *ids = cv.take<int32_t>(n);
  int32_t *status = cv.take<int32_t>(64);          // [0..1] status, [8..11] bits, [16..] level counts
  size_t sort_bytes = cub_sort_bytes(n), scan_bytes = cub_scan_bytes(n);
  void *cub_tmp = cv.take<char>(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
  OSB_CHECK(cv.ok, "osb_coordset_pyramid: workspace too small (%zu bytes)", ws_bytes);
  const unsigned long long bits_init[2] = {0ull, ~0ull};
  unsigned long long *bits = reinterpret_cast<unsigned long long *>(status + 8);
  int32_t *counts = status + 16;
  OSB_CUDA(cudaMemsetAsync(status, 0, 256, stream));
  OSB_CUDA(cudaMemcpyAsync(bits, bits_init, 16, cudaMemcpyHostToDevice, stream));
  const int32_t n32 = (int32_t)n;
  OSB_CUDA(cudaMemcpyAsync(counts, &n32, 4, cudaMemcpyHostToDevice, stream));
  const unsigned nb = (unsigned)ceil_div(n, 256);
  k_morton_from_coords<<<nb, 256, 0, stream>>>((const int4 *)coords, n, 1, morton, idx, status, bits);
  OSB_LAUNCH_CHECK();
  unsigned long long hb[2];
  OSB_CUDA(cudaMemcpyAsync(hb, bits, 16, cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaStreamSynchronize(stream));
  int lo, hi;
  varying_bits(hb[0], hb[1], &lo, &hi);
  uint64_t *sorted_keys = nullptr;
  OSB_CHECK(sort_to(morton, morton_s, idx, perm, n, lo, hi, cub_tmp, stream, &sorted_keys) == 0, "osb_coordset_pyramid: sort failed");
  k_permute_coords<<<nb, 256, 0, stream>>>((const int4 *)coords, perm, sorted_keys, n, (int4 *)coords_int, inv_perm, status);
  OSB_LAUNCH_CHECK();
  if (slots != nullptr && osb_hash_build(coords_int, n, slots, cap, stream_)) return 1;
  const int32_t *fine = coords_int;
  for (int l = 0; l < n_levels; ++l) {
    int32_t *coarse = coords_lvl + (int64_t)l * n * 4;
    int32_t *parent = parent_lvl + (int64_t)l * n;
    const int32_t new_ts = 2 << l;
    k_pyr_heads<<<nb, 256, 0, stream>>>((const int4 *)fine, counts + l, n, new_ts, heads);
    OSB_LAUNCH_CHECK();
    OSB_CHECK(inclusive_scan_i32(heads, ids, n, cub_tmp, stream) == 0, "osb_coordset_pyramid: scan failed");
    k_pyr_emit<<<nb, 256, 0, stream>>>((const int4 *)fine, counts + l, n, heads, ids, new_ts, (int4 *)coarse, parent, counts + l + 1);
    OSB_LAUNCH_CHECK();
    fine = coarse;
  }
  int32_t hc[16];
  int32_t hs[2];
  OSB_CUDA(cudaMemcpyAsync(hc, counts, sizeof(int32_t) * (n_levels + 1), cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaMemcpyAsync(hs, status, 8, cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaStreamSynchronize(stream));
  for (int l = 0; l <= n_levels; ++l) n_host[l] = hc[l];
  status_host[0] = hs[0]; status_host[1] = hs[1];
  key_bits_to_status(hb, status_host);
  return 0;
}

int osb_kernel_map_build(const int32_t *coords_out, int64_t n_out, const void *slots_in, int64_t cap_in, int32_t ks_x,
                         int32_t ks_y, int32_t ks_z, int32_t step, int32_t *nbr, int32_t *pairs_per_k, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int K = ks_x * ks_y * ks_z;
  OSB_CHECK(n_out > 0 && K >= 1 && K <= 1024 && step >= 1, "osb_kernel_map_build: bad arguments");
  OSB_CHECK((cap_in & (cap_in - 1)) == 0, "osb_kernel_map_build: cap must be a power of two");
  if (pairs_per_k) OSB_CUDA(cudaMemsetAsync(pairs_per_k, 0, sizeof(int32_t) * K, stream));
  dim3 grid((unsigned)ceil_div(n_out, 256), K);
  k_kernel_map<<<grid, 256, 0, stream>>>((const int4 *)coords_out, n_out, (const HashSlot *)slots_in,
                                         (uint64_t)cap_in - 1, ks_x, ks_y, ks_z, step, nbr, pairs_per_k);
  OSB_LAUNCH_CHECK();
  return 0;
}

static int make_grid_view(const void *grid, int32_t log2_ts, int32_t nbits, int32_t n_batch, OccGridView *g, const char *who) {
  OSB_CHECK(grid != nullptr && nbits >= 2 && nbits <= 9 && log2_ts >= 0 && log2_ts <= 16 && n_batch >= 1 && n_batch <= 1024,
            "%s: bad occupancy grid (nbits %d, log2_ts %d, n_batch %d)", who, nbits, log2_ts, n_batch);
  const int64_t words = (int64_t)n_batch * occgrid_words_per_batch(nbits);
  OSB_CHECK(words <= ((int64_t)1 << 21), "%s: occupancy grid of %lld words is too large", who, (long long)words);
  g->bitmap = reinterpret_cast<const unsigned long long *>(grid);
  g->first_row = reinterpret_cast<const int32_t *>(reinterpret_cast<const unsigned long long *>(grid) + words);
  g->nbits = nbits; g->log2_ts = log2_ts; g->n_batch = n_batch;
  return 0;
}

size_t osb_occgrid_bytes(int32_t nbits, int32_t n_batch) {
  if (nbits < 2 || nbits > 9 || n_batch < 1) return 0;
  const int64_t words = (int64_t)n_batch * occgrid_words_per_batch(nbits);
  return words > ((int64_t)1 << 21) ? 0 : (size_t)words * 12;
}

int osb_occgrid_build(const int32_t *coords_int, int64_t n, int32_t log2_ts, int32_t nbits, int32_t n_batch, void *grid,
                      int32_t *status_dev, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OccGridView g;
  if (make_grid_view(grid, log2_ts, nbits, n_batch, &g, "osb_occgrid_build")) return 1;
  OSB_CHECK(n > 0 && status_dev != nullptr, "osb_occgrid_build: bad arguments");
  const int64_t words = (int64_t)n_batch * occgrid_words_per_batch(nbits);
  OSB_CUDA(cudaMemsetAsync(grid, 0, (size_t)words * 8, stream));          // bitmap only; first_row is read where a bit is set
  k_occgrid_build<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>((const int4 *)coords_int, n, g, (unsigned long long *)grid,
                                                                  const_cast<int32_t *>(g.first_row), status_dev);
  OSB_LAUNCH_CHECK();
  return 0;
}

int osb_kernel_map_build_grid(const int32_t *coords_out, int64_t n_out, const void *grid, int32_t log2_ts, int32_t nbits,
                              int32_t n_batch, int32_t ks_x, int32_t ks_y, int32_t ks_z, int32_t step, int32_t *nbr,
                              int32_t *pairs_per_k, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int K = ks_x * ks_y * ks_z;
  OSB_CHECK(n_out > 0 && K >= 1 && K <= 1024 && step >= 1, "osb_kernel_map_build_grid: bad arguments");
  OccGridView g;
  if (make_grid_view(grid, log2_ts, nbits, n_batch, &g, "osb_kernel_map_build_grid")) return 1;
  if (pairs_per_k) OSB_CUDA(cudaMemsetAsync(pairs_per_k, 0, sizeof(int32_t) * K, stream));
  dim3 grid_dim((unsigned)ceil_div(n_out, 256), K);
  k_kernel_map_grid<<<grid_dim, 256, 0, stream>>>((const int4 *)coords_out, n_out, g, ks_x, ks_y, ks_z, step, nbr, pairs_per_k);
  OSB_LAUNCH_CHECK();
  return 0;
}

int osb_kernel_map_transpose(const int32_t *nbr, int64_t n_out, int32_t K, int32_t *nbr_t, int64_t n_in, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(n_out > 0 && n_in > 0 && K >= 1, "osb_kernel_map_transpose: bad arguments");
  k_fill_i32<<<(unsigned)ceil_div(n_in * K, 256), 256, 0, stream>>>(nbr_t, n_in * K, -1);
  OSB_LAUNCH_CHECK();
  k_kernel_map_transpose<<<(unsigned)ceil_div(n_out * K, 256), 256, 0, stream>>>(nbr, n_out, K, nbr_t, n_in);
  OSB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
