// Stem convolution (conv0p1s1: 5x5x5, 3 -> 32; models/mink_unet.py:47-48) with the kernel-map probe
// fused in: one warp per output voxel probes the 125 offsets in the coordinate hash (4 rounds of 32
// lanes), compacts the hits with a ballot, then lane n accumulates output channel n over the hits.
// The 5^3 map (500 B / voxel) is never written to HBM.  fp32 FMA; BatchNorm(eval)+ReLU folded.
#include "common.cuh"
#include <algorithm>

namespace osb {

constexpr int STEM_WARPS = 8;
constexpr int STEM_MAXHIT = 128;

__global__ void __launch_bounds__(STEM_WARPS * 32)
k_conv_stem(const float *__restrict__ in, int cin, const int4 *__restrict__ coords, int64_t n,
            const HashSlot *__restrict__ slots, uint64_t mask, int ks, int step, const float *__restrict__ w, int cout,
            const float *__restrict__ scale, const float *__restrict__ shift, int relu, uint8_t *__restrict__ out_split,
            float *__restrict__ out_f32) {
  extern __shared__ float s_w[];                       // [K][cin][cout]
  __shared__ int32_t s_hit_row[STEM_WARPS][STEM_MAXHIT];
  __shared__ int16_t s_hit_k[STEM_WARPS][STEM_MAXHIT];
  const int K = ks * ks * ks;
  for (int e = threadIdx.x; e < K * cin * cout; e += blockDim.x) s_w[e] = __ldg(w + e);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = ks / 2;
  const int nch = (cout + 31) / 32;                    // channels per lane (1 or 2)
  const int64_t warps_total = (int64_t)gridDim.x * STEM_WARPS;
  for (int64_t o = (int64_t)blockIdx.x * STEM_WARPS + warp; o < n; o += warps_total) {
    const int4 c = __ldg(coords + o);
    int nhit = 0;
    for (int kb = 0; kb < K; kb += 32) {
      const int k = kb + lane;
      int row = -1;
      if (k < K) {
        const int ix = k % ks, iy = (k / ks) % ks, iz = k / (ks * ks);
        const int dx = ((ks & 1) ? ix - half : ix) * step, dy = ((ks & 1) ? iy - half : iy) * step,
                  dz = ((ks & 1) ? iz - half : iz) * step;
        row = hash_lookup(slots, mask, pack_key(c.x, c.y + dx, c.z + dy, c.w + dz));
      }
      const unsigned bal = __ballot_sync(0xffffffffu, row >= 0);
      if (row >= 0) {
        const int pos = nhit + __popc(bal & ((1u << lane) - 1));
        s_hit_row[warp][pos] = row;
        s_hit_k[warp][pos] = (int16_t)k;
      }
      nhit += __popc(bal);
    }
    __syncwarp();
    float acc[2] = {0.f, 0.f};
    for (int h = 0; h < nhit; ++h) {
      const int row = s_hit_row[warp][h], k = s_hit_k[warp][h];
      const float *wk = s_w + k * cin * cout;
      for (int ci = 0; ci < cin; ++ci) {
        const float x = __ldg(in + (int64_t)row * cin + ci);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (j < nch && lane + 32 * j < cout) acc[j] = fmaf(x, wk[ci * cout + lane + 32 * j], acc[j]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ch = lane + 32 * j;
      if (j >= nch || ch >= cout) continue;
      float y = acc[j];
      if (scale) y = fmaf(y, __ldg(scale + ch), __ldg(shift + ch));
      if (relu) y = fmaxf(y, 0.f);
      if (out_split) {
        __nv_bfloat16 hi, lo;
        split_bf16(y, hi, lo);
        uint8_t *p = out_split + o * (int64_t)cout * 4 + split_off_hi(ch);
        *reinterpret_cast<__nv_bfloat16 *>(p) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(p + 64) = lo;
      }
      if (out_f32) out_f32[o * cout + ch] = y;
    }
  }
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_conv_stem_fused(const float *in, int32_t cin, const int32_t *coords, int64_t n, const void *slots, int64_t cap,
                        int32_t ks, int32_t step, const float *w, int32_t cout, const float *scale, const float *shift,
                        int32_t relu, void *out_split, float *out_f32, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int K = ks * ks * ks;
  OSB_CHECK(cin >= 1 && cin <= 8 && cout >= 1 && cout <= 64, "osb_conv_stem_fused: needs cin <= 8, cout <= 64 (got %d, %d)", cin, cout);
  OSB_CHECK(K >= 1 && K <= STEM_MAXHIT, "osb_conv_stem_fused: kernel volume %d not supported", K);
  OSB_CHECK(out_split == nullptr || cout % 32 == 0, "osb_conv_stem_fused: split output needs cout %% 32 == 0");
  OSB_CHECK((scale == nullptr) == (shift == nullptr), "osb_conv_stem_fused: scale and shift go together");
  OSB_CHECK(n > 0 && (cap & (cap - 1)) == 0, "osb_conv_stem_fused: bad n / cap");
  const size_t smem = (size_t)K * cin * cout * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    OSB_CUDA(cudaFuncSetAttribute(k_conv_stem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  OSB_CHECK(smem <= 200 * 1024, "osb_conv_stem_fused: weights do not fit in shared memory");
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, STEM_WARPS), 148 * 4);
  k_conv_stem<<<grid, STEM_WARPS * 32, smem, stream>>>(in, cin, (const int4 *)coords, n, (const HashSlot *)slots,
                                                       (uint64_t)cap - 1, ks, step, w, cout, scale, shift, relu,
                                                       (uint8_t *)out_split, out_f32);
  OSB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
