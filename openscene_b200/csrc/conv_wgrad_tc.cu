// Weight gradient of the sparse convolution on 5th-gen tensor cores (run/distill.py:333, `loss.backward()` through
// every MinkowskiConvolution / MinkowskiConvolutionTranspose of models/mink_unet.py):
//
//   gW[k][ci][co] = sum_{o : nbr[k][o] >= 0}  x[nbr[k][o], ci] * gout[o, co]
//
// The reduction runs over ROWS, the dimension along which the split-bf16 activations are NOT contiguous: both operands
// are therefore MN-major UMMA operands.  A 128-byte line of a split row, [hi x32 | lo x32] of one 32-channel block, is 64
// consecutive "MN" elements; 8 consecutive rows form the 1024-byte 128B-swizzle atom (physically the same shared-memory
// image the forward kernel gathers, only read with the transpose bits of the instruction descriptor set).  One MMA
// (M128 x N<=256 x K16) multiplies 16 rows of two input-channel blocks with 16 rows of up to four output-channel blocks; its
// fp32 result tile holds, per (block, block) pair, the four products hi*hi, hi*lo, lo*hi, lo*lo, which the reduce kernel
// adds up (the full (hi+lo)(hi+lo) product: slightly MORE accurate than the three-term forward).
//
// Work unit = (offset k, pair of input blocks, group of <= 4 output blocks, range of output rows); one CTA per unit writes
// its raw 128 x N accumulator to a partial buffer, a second kernel sums quadrants and row ranges in a fixed order: no
// atomics, bit-reproducible.  Missing neighbours are zero-filled rows of the gathered operand (they add nothing).
#include "tc_ptx.cuh"
#include <algorithm>

namespace osb {

constexpr int WG_THREADS = 192;
constexpr int WG_ROWS = 128;                 // rows (the MMA K dimension) per pipeline stage
constexpr int WG_TILE = WG_ROWS * 128;       // one (128 rows x one 32-channel block) tile: 16 KB
constexpr int WG_STAGES = 2;
constexpr int WG_STAGE_BYTES = 6 * WG_TILE;  // 2 input-block tiles + 4 output-block tiles

struct WgradParams {
  const uint8_t *x;            // split rows [n_in, cin]
  const int32_t *nbr;          // [K][n_out] or NULL (identity, K == 1)
  int64_t n_out;
  int K, nbi, nbo;             // input / output channel blocks (cin / 32, cout / 32)
  int n_mt, n_nt, n_rs;        // unit grid: pairs of input blocks, groups of 4 output blocks, row ranges
  int64_t rows_per_rs;         // multiple of WG_ROWS
  float *partial;              // [units][128][256] raw accumulators
};

// MN-major, 128-byte swizzle: 8 K-rows = one 1024-byte atom (SBO), 64-element MN chunks LBO bytes apart
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(WG_THREADS)
k_conv_wgrad_tc(const __grid_constant__ CUtensorMap tmG, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + WG_STAGES * WG_STAGE_BYTES);    // fullA[2], fullB[2], empty[2], accum
  uint32_t *s_misc = reinterpret_cast<uint32_t *>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t fullA = smem_u32(bars), fullB = smem_u32(bars + 2), empty0 = smem_u32(bars + 4), accum_bar = smem_u32(bars + 6);

  // unit -> (k, mt, nt, rs)
  int u = blockIdx.x;
  const int rs = u % p.n_rs; u /= p.n_rs;
  const int ntile = u % p.n_nt; u /= p.n_nt;
  const int mt = u % p.n_mt;
  const int k = u / p.n_mt;
  const int ib0 = mt * 2, n_ib = min(2, p.nbi - ib0);          // input blocks of this unit
  const int ob0 = ntile * 4, n_ob = min(4, p.nbo - ob0);       // output blocks
  const int64_t r_begin = (int64_t)rs * p.rows_per_rs, r_end = min(r_begin + p.rows_per_rs, p.n_out);
  const int n_stage = (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS);

  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(fullA + 8 * s, 128); mbar_init(fullB + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 64) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_misc[0])), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_misc[0];

  if (warp == 0) {
    // ============ gout tiles by TMA: rows [r, r+128) x one 32-channel block each (rows past n_out: zero fill) ============
    int s = 0; uint32_t phase = 0;
    for (int t = 0; t < n_stage; ++t) {
      mbar_wait(empty0 + 8 * s, phase ^ 1);
      if (elect_one()) {
        const uint32_t fb = fullB + 8 * s;
        mbar_expect_tx(fb, (uint32_t)(n_ob * WG_TILE));
        const int row = (int)(r_begin + (int64_t)t * WG_ROWS);
        for (int b = 0; b < n_ob; ++b)
          tma_load_2d(smem_u32(smem + s * WG_STAGE_BYTES + (2 + b) * WG_TILE), &tmG, fb, (ob0 + b) * 64, row);
      }
      __syncwarp();
      if (++s == WG_STAGES) { s = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    // D = f32, A = B = bf16, both operands MN-major (bits 15, 16), N = 64 per output block, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((n_ob * 64) >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    int s = 0; uint32_t phase = 0;
    for (int t = 0; t < n_stage; ++t) {
      mbar_wait(fullB + 8 * s, phase);
      mbar_wait(fullA + 8 * s, phase);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t base = smem_u32(smem + s * WG_STAGE_BYTES);
        const uint64_t da = umma_desc_mn(base, WG_TILE), db = umma_desc_mn(base + 2 * WG_TILE, WG_TILE);
#pragma unroll
        for (int ks = 0; ks < WG_ROWS / 16; ++ks)          // 16 rows = two 1024-byte atoms per MMA
          umma_bf16(tmem_base, da + (uint64_t)(ks * 128), db + (uint64_t)(ks * 128), idesc, (t == 0 && ks == 0) ? 0u : 1u);
        umma_commit(empty0 + 8 * s);
      }
      __syncwarp();
      if (++s == WG_STAGES) { s = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(accum_bar);
    __syncwarp();
  } else {
    // ============ gathered x rows (32 rows per warp and stage, as in the forward kernels), then the epilogue ============
    const int w = warp - 2, j = lane & 7, q = lane >> 3;
    const int64_t row_bytes = (int64_t)p.nbi * 128;
    int s = 0; uint32_t phase = 0;
    for (int t = 0; t < n_stage; ++t) {
      int32_t ridx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t o = r_begin + (int64_t)t * WG_ROWS + w * 32 + 4 * i + q;
        ridx[i] = (o < r_end) ? (p.nbr ? __ldg(p.nbr + (int64_t)k * p.n_out + o) : (int32_t)o) : -1;
      }
      mbar_wait(empty0 + 8 * s, phase ^ 1);
      for (int b = 0; b < 2; ++b) {
        const uint32_t a_dst = smem_u32(smem + s * WG_STAGE_BYTES + b * WG_TILE) + (w * 32 + q) * 128;
        const bool have = b < n_ib;                      // a missing second block is multiplied as zeros (rows 64..127 of D unused)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m7 = (4 * i + q) & 7;
          const bool valid = have && ridx[i] >= 0;
          const uint8_t *sp = valid ? p.x + (int64_t)ridx[i] * row_bytes + (ib0 + b) * 128 + j * 16 : p.x;
          cp_async16(a_dst + i * 512 + ((j ^ m7) << 4), sp, valid ? 16u : 0u);
        }
      }
      cp_async_arrive_noinc(fullA + 8 * s);
      if (++s == WG_STAGES) { s = 0; phase ^= 1; }
    }
    // ---- epilogue: raw accumulator (128 x 64*n_ob fp32) -> partial[unit], full-line coalesced through a staging tile
    const int qq = warp & 3;                             // TMEM lane quarter
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t stg = smem_u32(smem) + (uint32_t)(warp - 2) * 4096u;
    const int rsub = lane >> 3, chunk = lane & 7, sw = lane & 7;
    const uint32_t my_line = stg + lane * 128;
    float *dst = p.partial + (int64_t)blockIdx.x * 128 * 256;
    for (int cb = 0; cb < n_ob * 2; ++cb) {
      uint32_t v0[16], v1[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(qq * 32) << 16) + cb * 32;
      tmem_ld16(taddr, v0);
      tmem_ld16(taddr + 16, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int g = 0; g < 4; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my_line + ((g ^ sw) << 4)), "r"(v0[4 * g]), "r"(v0[4 * g + 1]),
                     "r"(v0[4 * g + 2]), "r"(v0[4 * g + 3]) : "memory");
#pragma unroll
      for (int g = 0; g < 4; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my_line + (((4 + g) ^ sw) << 4)), "r"(v1[4 * g]), "r"(v1[4 * g + 1]),
                     "r"(v1[4 * g + 2]), "r"(v1[4 * g + 3]) : "memory");
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rsub;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + r * 128 + ((chunk ^ (r & 7)) << 4)));
        *reinterpret_cast<uint4 *>(dst + (int64_t)(qq * 32 + r) * 256 + cb * 32 + chunk * 4) = v;
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
}

// gw[k][ci][co] = sum over row ranges and over the four (hi|lo) x (hi|lo) quadrants of the unit's accumulator
__global__ void k_conv_wgrad_reduce(const float *__restrict__ partial, int K, int cin, int cout, int n_mt, int n_nt, int n_rs,
                                    float *__restrict__ gw) {
  const int64_t total = (int64_t)K * cin * cout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(e % cout);
    const int64_t kc = e / cout;
    const int ci = (int)(kc % cin), k = (int)(kc / cin);
    const int cb = ci >> 5, mt = cb >> 1, row_hi = (cb & 1) * 64 + (ci & 31);
    const int ob = co >> 5, ntile = ob >> 2, col_hi = (ob & 3) * 64 + (co & 31);
    float acc = 0.f;
    for (int rs = 0; rs < n_rs; ++rs) {
      const int64_t unit = (((int64_t)k * n_mt + mt) * n_nt + ntile) * n_rs + rs;
      const float *t = partial + unit * 128 * 256;
      acc += (t[row_hi * 256 + col_hi] + t[row_hi * 256 + col_hi + 32]) + (t[(row_hi + 32) * 256 + col_hi] + t[(row_hi + 32) * 256 + col_hi + 32]);
    }
    gw[e] = acc;
  }
}

}  // namespace osb

using namespace osb;

static void wgrad_plan(int64_t n_out, int K, int cin, int cout, int *n_mt, int *n_nt, int *n_rs, int64_t *rows_per_rs) {
  *n_mt = (cin / 32 + 1) / 2;
  *n_nt = (cout / 32 + 3) / 4;
  const int64_t base = (int64_t)K * *n_mt * *n_nt;
  const int64_t chunks = ceil_div(n_out, WG_ROWS);
  int64_t rs = std::max<int64_t>(1, (2 * 148 + base - 1) / base);           // about two CTAs' worth of units per SM ...
  rs = std::min(rs, std::max<int64_t>(1, chunks / 4));                      // ... but at least 4 stages per unit
  *rows_per_rs = ceil_div(chunks, rs) * WG_ROWS;
  *n_rs = (int)ceil_div(n_out, *rows_per_rs);
}

extern "C" {

size_t osb_conv_wgrad_tc_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout) {
  if (n_out <= 0 || K < 1 || cin < 32 || cout < 32) return 0;     // shapes osb_conv_wgrad_tc rejects: nothing to reserve
  int n_mt, n_nt, n_rs; int64_t rpr;
  wgrad_plan(n_out, K, cin, cout, &n_mt, &n_nt, &n_rs, &rpr);
  return (size_t)K * n_mt * n_nt * n_rs * 128 * 256 * sizeof(float);
}

int osb_conv_wgrad_tc(const void *x_split, int32_t cin, int64_t n_in, const int32_t *nbr, int64_t n_out, int32_t K,
                      const void *gout_split, int32_t cout, float *gw, void *ws, size_t ws_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(x_split && gout_split && gw, "osb_conv_wgrad_tc: null argument");
  OSB_CHECK(cin > 0 && cin % 32 == 0 && cout > 0 && cout % 32 == 0, "osb_conv_wgrad_tc: channel counts must be multiples of 32 (%d, %d)", cin, cout);
  OSB_CHECK(K >= 1 && (nbr != nullptr || K == 1), "osb_conv_wgrad_tc: identity map needs K == 1");
  OSB_CHECK(n_out > 0 && n_out < (1ll << 31) && n_in > 0, "osb_conv_wgrad_tc: bad row counts");
  WgradParams p{};
  wgrad_plan(n_out, K, cin, cout, &p.n_mt, &p.n_nt, &p.n_rs, &p.rows_per_rs);
  const size_t need = osb_conv_wgrad_tc_workspace_bytes(n_out, K, cin, cout);
  OSB_CHECK(ws != nullptr && ws_bytes >= need, "osb_conv_wgrad_tc: workspace of %zu bytes required (got %zu)", need, ws_bytes);
  p.x = (const uint8_t *)x_split; p.nbr = nbr; p.n_out = n_out; p.K = K; p.nbi = cin / 32; p.nbo = cout / 32;
  p.partial = (float *)ws;
  CUtensorMap tmG;
  if (make_tmap_2b(&tmG, gout_split, 2ull * cout, (uint64_t)n_out, WG_ROWS, 0)) return 1;
  const size_t smem_bytes = (size_t)WG_STAGES * WG_STAGE_BYTES + 128 + 1024;
  OSB_SMEM_ATTR_ONCE(k_conv_wgrad_tc, 227 * 1024);
  const int64_t units = (int64_t)K * p.n_mt * p.n_nt * p.n_rs;
  k_conv_wgrad_tc<<<(unsigned)units, WG_THREADS, smem_bytes, stream>>>(tmG, p);
  OSB_LAUNCH_CHECK();
  const int64_t total = (int64_t)K * cin * cout;
  k_conv_wgrad_reduce<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 8), 256, 0, stream>>>(p.partial, K, cin, cout, p.n_mt, p.n_nt,
                                                                                                     p.n_rs, gw);
  OSB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
