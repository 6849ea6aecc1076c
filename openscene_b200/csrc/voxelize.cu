// GPU voxeliser: affine transform + floor (fp64), shift to origin, FNV-64 key, unique with
// first-occurrence representative in ascending-key order, inverse map.  Bit-compatible with
// dataset/voxelizer.py:116-130 + dataset/voxelization_utils.py:9-22,107-131 (np.unique semantics).
#include "common.cuh"

#include "sortscan.cuh"
#include <algorithm>

namespace osb {

struct Mat34 { double m[3][4]; };

template <typename T>
__global__ void k_vox_transform(const T *__restrict__ pts, int64_t n, Mat34 M, long long *__restrict__ cmin,
                                long long *__restrict__ cint) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  long long lmin[3] = {LLONG_MAX, LLONG_MAX, LLONG_MAX};
  if (i < n) {
    const double x = (double)pts[3 * i], y = (double)pts[3 * i + 1], z = (double)pts[3 * i + 2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      // row j of [p,1] @ M^T[:, :3]; accumulation order of a 4-term dot product with FMA
      double t = x * M.m[j][0];
      t = fma(y, M.m[j][1], t);
      t = fma(z, M.m[j][2], t);
      t = t + M.m[j][3];
      const long long c = (long long)floor(t);
      cint[3 * i + j] = c;
      lmin[j] = c;
    }
  }
  // block reduce min, one atomic per block and axis
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    long long v = lmin[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      long long w = __shfl_xor_sync(0xffffffffu, v, o);
      v = w < v ? w : v;
    }
    if ((threadIdx.x & 31) == 0 && v != LLONG_MAX) atomicMin(cmin + j, v);
  }
}

__global__ void k_vox_keys(const long long *__restrict__ cint, const long long *__restrict__ cmin, int64_t n,
                           uint64_t *__restrict__ key, int32_t *__restrict__ idx, int32_t *__restrict__ c32) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t h = 14695981039346656037ull;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const long long c = cint[3 * i + j] - cmin[j];
    c32[3 * i + j] = (int32_t)c;
    h *= 1099511628211ull;      // multiply THEN xor, whole uint64 words (voxelization_utils.py:19-21)
    h ^= (uint64_t)c;
  }
  key[i] = h;
  idx[i] = (int32_t)i;
}

__global__ void k_vox_heads(const uint64_t *__restrict__ key_s, int64_t n, int32_t *__restrict__ heads) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  heads[j] = (j == 0 || key_s[j] != key_s[j - 1]) ? 1 : 0;
}

__global__ void k_vox_emit(const int32_t *__restrict__ idx_s, const int32_t *__restrict__ heads,
                           const int32_t *__restrict__ ids, const int32_t *__restrict__ c32, int64_t n,
                           int32_t *__restrict__ coords_vox, int64_t *__restrict__ inds, int64_t *__restrict__ inds_reverse) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t p = idx_s[j];
  const int32_t v = ids[j] - 1;
  inds_reverse[p] = v;
  if (heads[j]) {   // stable sort + ascending payload => the run head is the first occurrence
    inds[v] = p;
    coords_vox[3 * v] = c32[3 * (int64_t)p];
    coords_vox[3 * v + 1] = c32[3 * (int64_t)p + 1];
    coords_vox[3 * v + 2] = c32[3 * (int64_t)p + 2];
  }
}

struct VCarver {
  char *p; size_t left; bool ok = true;
  template <typename T> T *take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    if (bytes > left) { ok = false; return nullptr; }
    T *r = reinterpret_cast<T *>(p); p += bytes; left -= bytes; return r;
  }
};

static size_t vox_sort_bytes(int64_t n) { return std::max(radix_sort_ws_bytes(n), scan_ws_bytes(n)); }

}  // namespace osb

using namespace osb;

extern "C" {

size_t osb_voxelize_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return (size_t)n * (24 + 8 + 8 + 4 + 4 + 12 + 4 + 4) + 12 * 256 + vox_sort_bytes(n) + 1024;
}

int osb_voxelize(const void *coords, int32_t coords_is_f64, int64_t n, const double *matrix_host, int32_t *coords_vox,
                 int64_t *inds, int64_t *inds_reverse, int64_t *n_vox_host, double *min_host, void *ws, size_t ws_bytes,
                 void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(n > 0 && n < (1ll << 31) - 1024, "osb_voxelize: n=%lld out of range", (long long)n);
  VCarver cv{(char *)ws, ws_bytes};
  long long *cint = cv.take<long long>(3 * n);
  uint64_t *key = cv.take<uint64_t>(n);
  uint64_t *key_s = cv.take<uint64_t>(n);
  int32_t *idx = cv.take<int32_t>(n);
  int32_t *idx_s = cv.take<int32_t>(n);
  int32_t *c32 = cv.take<int32_t>(3 * n);
  int32_t *heads = cv.take<int32_t>(n);
  int32_t *ids = cv.take<int32_t>(n);
  long long *cmin = cv.take<long long>(4);
  const size_t tmp_bytes = vox_sort_bytes(n);
  void *tmp = cv.take<char>(tmp_bytes);
  OSB_CHECK(cv.ok, "osb_voxelize: workspace too small (%zu bytes)", ws_bytes);

  Mat34 M;
  for (int j = 0; j < 3; ++j)
    for (int m = 0; m < 4; ++m) M.m[j][m] = matrix_host[4 * j + m];
  const long long init[4] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, 0};
  OSB_CUDA(cudaMemcpyAsync(cmin, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  const unsigned nb = (unsigned)ceil_div(n, 256);
  if (coords_is_f64) k_vox_transform<double><<<nb, 256, 0, stream>>>((const double *)coords, n, M, cmin, cint);
  else               k_vox_transform<float><<<nb, 256, 0, stream>>>((const float *)coords, n, M, cmin, cint);
  OSB_LAUNCH_CHECK();
  k_vox_keys<<<nb, 256, 0, stream>>>(cint, cmin, n, key, idx, c32);
  OSB_LAUNCH_CHECK();
  // stable LSD radix sort over all 64 key bits (8 passes: an even count, so the result is back in the 'a' buffers)
  const int where = radix_sort_pairs(key, idx, key_s, idx_s, nullptr, n, 0, 64, tmp, stream);
  OSB_CHECK(where >= 0, "osb_voxelize: sort failed");
  const uint64_t *ks = where ? key_s : key;
  const int32_t *is = where ? idx_s : idx;
  k_vox_heads<<<nb, 256, 0, stream>>>(ks, n, heads);
  OSB_LAUNCH_CHECK();
  OSB_CHECK(inclusive_scan_i32(heads, ids, n, tmp, stream) == 0, "osb_voxelize: scan failed");
  k_vox_emit<<<nb, 256, 0, stream>>>(is, heads, ids, c32, n, coords_vox, inds, inds_reverse);
  OSB_LAUNCH_CHECK();
  int32_t last = 0;
  long long hmin[4];
  OSB_CUDA(cudaMemcpyAsync(&last, ids + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaMemcpyAsync(hmin, cmin, sizeof(hmin), cudaMemcpyDeviceToHost, stream));
  OSB_CUDA(cudaStreamSynchronize(stream));
  *n_vox_host = last;
  if (min_host) for (int j = 0; j < 3; ++j) min_host[j] = (double)hmin[j];
  return 0;
}

}  // extern "C"
