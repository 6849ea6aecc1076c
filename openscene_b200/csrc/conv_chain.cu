// Persistent, plan-driven sparse convolution on 5th-gen tensor cores (second generation of conv_tc.cu).
//
//   out[o,:] = epilogue( sum_k  in[nbr[k][o], :] @ W[k] )         (output-stationary, no atomics)
//
// One launch executes a LIST of convolution layers ("chain") described by ConvDesc records in device memory.  The grid
// is one CTA per SM; every CTA owns a contiguous range of each layer's work units (unit = 128 output rows x one N tile x
// one split of the (offset, channel-block) stage sequence) and walks it with warp-specialised roles that never leave
// their loops between units:
//
//   warps 0-3   epilogue of both TMEM accumulator buffers: TMEM -> registers -> BN affine / residual / ReLU -> swizzled
//               staging tile -> full-line coalesced stores (split rows, fp32 rows, or raw split-K partials)
//   warp 4      weight tiles: one cp.async.bulk per stage from the tile-major, pre-swizzled packing (no tensor map)
//   warps 5-6   tcgen05.mma issuers (one thread each): warp 5 owns sub-tile 0 of an item, warp 6 sub-tile 1; warp 5 owns TMEM
//   warps 8-15  gathered A rows: cp.async 16 B x 8 lanes per 128-byte row line, hand-applied 128B swizzle, the kernel
//               map read per offset straight from global memory, one offset ahead
//
// What changed against conv_tc.cu, and why (profiles/r01_ncu_full_conv_tc_96x96_k3_final.md, DESIGN.md "slot model"):
//   * separate rings for gathered rows and weight tiles; the whole SM's shared memory belongs to one CTA: 9-10 row
//     slots of 16 KB in flight per SM instead of 6 (the stage rate was latency x row bytes in flight);
//   * two 128-row sub-tiles share every weight tile (256 output rows per item): half the L2->SM weight stream;
//   * the accumulator is double buffered in TMEM (2 x 256 columns) and drained by dedicated epilogue warps while the
//     next item's MMAs run; barrier / TMEM set-up is paid once per CTA, not once per tile; work is split evenly over
//     the SMs (no 6-vs-5.2 wave tail);
//   * split-K partials are reduced INSIDE the kernel after a grid barrier, and consecutive small layers (levels 2-4 of
//     the U-Net) run in one launch with grid barriers between dependent layers: no launch / finish-kernel boundaries.
//
// Numerics are those of conv_tc.cu: split-bf16 operands (v = hi + lo), hi*Whi + hi*Wlo + lo*Whi on kind::f16 MMAs,
// fp32 accumulation in TMEM, deterministic (fixed-order) split-K reduction.
#include "tc_ptx.cuh"
#include <algorithm>
#include <cstring>
#include <string>

namespace osb {

constexpr int CH_THREADS = 384;                  // 12 warps (3 per scheduler -> 168 registers each): 4 epilogue, 1 weights, 2 MMA issuers, 5 gather
constexpr int CH_M = 128;                        // rows per sub-tile (UMMA M)
constexpr int CH_A_BYTES = CH_M * 128;           // one row slot: 128 rows x one 32-channel block
constexpr int CH_STG_BYTES = 4 * 4096;           // epilogue staging: 4 warps x (32 rows x 128 B)
constexpr int CH_SS_FLOATS = 768;                // folded BN constants kept in shared memory per layer (scale | shift)
constexpr int CH_MAX_SA = 12, CH_MAX_SB = 4;
// Warp roles.  Measured with per-role cycle counters (profiles/r02_chain_roles.md): every role is ONE warp walking a
// dependent instruction chain, so its fixed cost per row slot (barrier wait, address set-up, arrival: 300-500 cycles) is
// latency, not throughput.  Gather producers therefore own whole slots (ring slot s is always filled by warp s mod CH_A_WARPS, 32 copy
// instructions behind one wait / one arrival), the epilogue (idle 90 % of the time) gets four warps for both TMEM buffers.
constexpr int CH_W_EPI = 0;                       // warps 0-3: epilogue (warp % 4 = TMEM lane quarter), both accumulator buffers
constexpr int CH_W_B = 4;                         // weight tiles
constexpr int CH_W_MMA = 5;                       // warps 5, 6: MMA issuers; warp 5 owns the TMEM allocation
constexpr int CH_W_A = 7;                         // warps 7-11: gathered rows, one whole 128-row slot at a time each
constexpr int CH_A_WARPS = 5;
constexpr int CH_DESC_WORDS = 48;                // sizeof(ConvDesc) / 4
constexpr int CH_MAX_LAYERS = 16;                // layers per launch: the descriptors travel as kernel parameters (3 KB)

struct __align__(16) ConvDesc {
  const uint8_t *src0, *src1;      // split rows of the (up to) two sources ([src0 | src1] = ME.cat)
  const int32_t *nbr;              // [K][n_out] input row per (offset, output row), -1 = none; NULL = identity (K == 1)
  const uint8_t *wtiles;           // tile-major pre-swizzled weights (osb_conv_pack_weight_tiles)
  const float *scale, *shift;      // folded BatchNorm, or NULL
  const uint8_t *res;              // residual split rows [n_out, cout], or NULL
  uint8_t *out_split;              // split rows out, or NULL
  float *out_f32;                  // fp32 rows out, or NULL
  const int32_t *out_row_map;      // fp32 rows scattered: row o -> out_row_map[o]
  const int32_t *cmap;             // dense transposed conv: column block kch of row o -> fine row cmap[kch*n_out + o]
  float *partial;                  // [nsplit][n_out][cout_pad] raw accumulators (nsplit > 1)
  int64_t n_out;
  int K, nb0, nb1;
  int cout, cout_pad, nt, n_ntiles;
  int relu, cmap_cout, nsplit, m_tiles;
  int nsub_max;                    // sub-tiles per item that may share a weight tile: 2 when nt <= 128, else 1
  int barrier_before;              // grid barrier before this layer (it reads what an earlier layer of the launch wrote)
  int stages_per_split;            // ceil(K * (nb0 + nb1) / nsplit)
  int pad[8];
};
static_assert(sizeof(ConvDesc) == CH_DESC_WORDS * 4, "ConvDesc layout");
// The layer list lives in the kernel's parameter (constant) space: every field is a warp-uniform value to the compiler, so
// the single-thread roles (MMA issuers, weight producer) keep their slot / descriptor arithmetic on the uniform datapath
// instead of paying vector->uniform register moves in front of every tcgen05.mma (profiles/r02_chain_roles.md).
struct ChainArgs { ConvDesc d[CH_MAX_LAYERS]; };

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one non-blocking-ish probe of an mbarrier phase (true = complete).  Several probes issued back to back overlap their
// ~190-cycle round trips; a chain of mbar_wait calls pays them one after the other.
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}

// A wait that never completes must not hang the GPU: after ~2^24 probes the thread reports where it is stuck into the
// tuning buffer (if one is set; use pinned host memory so that the report survives the trap) and traps.
__device__ long long *g_chain_report = nullptr;
__device__ __noinline__ void chain_stuck(uint32_t bar, uint32_t parity, int tag, uint32_t it) {
  long long *r = g_chain_report;
  if (r && it == (1u << 20) + 1) {               // report once, keep waiting so that the other stuck roles can report too
    long long *o = r + 1 + 4 * ((blockIdx.x * 16 + (threadIdx.x >> 5)) % 1024);
    o[0] = ((long long)blockIdx.x << 32) | (threadIdx.x >> 5); o[1] = bar; o[2] = parity; o[3] = tag;
    r[0] = 1;
    __threadfence_system();
  }
  if (it > (1u << 23) || !r) __trap();
}
__device__ __forceinline__ void chain_wait(uint32_t bar, uint32_t parity, int tag) {
  for (uint32_t it = 0;; ++it) {
    if (mbar_try(bar, parity)) return;
    if (it > (1u << 20)) chain_stuck(bar, parity, tag, it);
  }
}

// wait of a role with slack (epilogue, weight producer): back off between polls so that the polling does not take issue slots
// from the roles on the critical path
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned sleep_ns) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) {
      __nanosleep(sleep_ns);
      if (it > (1u << 24)) __trap();
    }
  }
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Sense-reversing grid barrier over {count, generation} in global memory (both zero before the first use ever; the
// barrier leaves count == 0 behind, so no host-side reset between launches).  Every CTA of the grid must be resident:
// the launch uses at most one CTA per SM.
__device__ __forceinline__ void grid_barrier(unsigned *gbar, unsigned &gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned old = atomicAdd(gbar, 1u);
    if (old == gridDim.x - 1) {
      gbar[0] = 0;
      __threadfence();
      atomicAdd(gbar + 1, 1u);
    } else {
      unsigned it = 0;
      while (ld_acquire_u32(gbar + 1) == gen) {
        if (++it > (1u << 24)) __trap();          // a CTA that never arrives must not hang the GPU
      }
    }
    __threadfence();
  }
  gen += 1;
  __syncthreads();
}

// per-role cycle accounting (tuning; active only when a clock buffer is given): dbg_clock[cta*32 + slot]
#define CH_PROF_BEGIN() const long long _t0 = prof ? clock64() : 0
#define CH_PROF_END(var) do { if (prof) var += clock64() - _t0; } while (0)

// keep a value in its register: stops the compiler from re-deriving shared-window addresses (S2R + shifts) in hot loops
#define CH_KEEP(x) asm volatile("" : "+r"(x))

// 16-byte global -> shared copy; `ignore` != 0 writes zeros instead (a missing neighbour row)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t ignore) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "cp.async.cg.shared.global [%0], [%1], 16, p;\n\t}"
      ::"r"(dst), "l"(src), "r"(ignore)
      : "memory");
}

// ------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(CH_THREADS, 1)
k_conv_chain(const __grid_constant__ ChainArgs args, int n_layers, unsigned *gbar, int sa, int sb, int bslot, int flags,
             long long *dbg_clock) {
  extern __shared__ uint8_t smem_raw[];
  // All hot-loop addressing is done on 32-bit shared-window addresses computed once; the few generic accesses (descriptor,
  // BN constants) use pointers derived from smem_raw by an offset, so that the compiler keeps them in the shared space.
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base_u32 = (raw_u32 + 1023u) & ~1023u;                      // 1024-byte aligned: 128B-swizzle atoms
  uint8_t *smem = smem_raw + (base_u32 - raw_u32);
  const uint32_t a_ring = base_u32;                                          // row slots
  const uint32_t b_ring = a_ring + (uint32_t)sa * CH_A_BYTES;                // weight slots (bslot is a multiple of 1024)
  const uint32_t stg_u32 = b_ring + (uint32_t)sb * (uint32_t)bslot;          // epilogue staging: 4 warps x 4 KB
  uint32_t a_ring_k = a_ring, b_ring_k = b_ring;
  CH_KEEP(a_ring_k); CH_KEEP(b_ring_k);
  uint8_t *aux = smem + (stg_u32 - base_u32) + CH_STG_BYTES;
  float *s_ss = reinterpret_cast<float *>(aux);                              // [scale x CH_SS_FLOATS | shift x CH_SS_FLOATS]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_ss + 2 * CH_SS_FLOATS);
  uint32_t *s_misc = reinterpret_cast<uint32_t *>(bars + 40);                // [0] TMEM base

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (flags & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool prof = dbg_clock != nullptr;
  long long pw0 = 0, pw1 = 0, pw2 = 0, pt = 0;     // cycles in this role's waits (up to three kinds) and in its loop
  if (dbg_clock && tid == 0) dbg_clock[blockIdx.x * 32 + 0] = clock64();
  uint32_t fullA = smem_u32(bars), emptyA = fullA + 12 * 8, fullB = fullA + 24 * 8, emptyB = fullA + 28 * 8;
  uint32_t accFull = fullA + 32 * 8, accEmpty = fullA + 34 * 8;
  CH_KEEP(fullA); CH_KEEP(emptyA); CH_KEEP(fullB); CH_KEEP(emptyB); CH_KEEP(accFull); CH_KEEP(accEmpty);

  if (tid == 0) {
    for (int s = 0; s < sa; ++s) { mbar_init(fullA + 8 * s, 32); mbar_init(emptyA + 8 * s, 1); }   // fullA: the 32 lanes of the slot's warp
    for (int s = 0; s < sb; ++s) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 2); }   // emptyB: one arrival per issuer
    for (int b = 0; b < 2; ++b) { mbar_init(accFull + 8 * b, 2); mbar_init(accEmpty + 8 * b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CH_W_MMA) {   // all 512 TMEM columns: two accumulator buffers of 256 columns (one CTA per SM, no contention)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_misc[0])), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  unsigned bar_gen = 0;
  if (tid == 0 && gbar) bar_gen = ld_acquire_u32(gbar + 1);    // before this launch's first barrier can complete
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // Everything above touched no data of an earlier kernel in the stream; from here on we read activations.
  if (flags & 1) asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = s_misc[0];
  if (dbg_clock && tid == 0) dbg_clock[blockIdx.x * 32 + 1] = clock64();

  // pipeline state of this thread's role; persists over items and layers
  uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0, n_item = 0;
  uint32_t g_slot = 0;                             // gather producers: row slots the CTA has gone through before the current item
  uint32_t p_sl = warp >= CH_W_A ? (uint32_t)(warp - CH_W_A) : 0u, p_lapb = 0, p_par = 0;   // gather producers: my next ring slot, global index of slot 0 of its lap, lap parity

  for (int L = 0; L < n_layers; ++L) {
    __syncthreads();                                   // every role is done with the previous layer (and with s_ss)
    const ConvDesc *s_desc = &args.d[L];               // parameter space: uniform loads
    const int d_K = s_desc->K, d_nb0 = s_desc->nb0, d_nb1 = s_desc->nb1, d_nt = s_desc->nt, d_n_ntiles = s_desc->n_ntiles;
    const int d_m_tiles = s_desc->m_tiles, d_nsplit = s_desc->nsplit, d_nsub_max = s_desc->nsub_max, d_sps = s_desc->stages_per_split;
    const int64_t d_n_out = s_desc->n_out;
    {                                                  // folded BN constants of the layer -> shared memory
      const int nss = s_desc->cmap ? s_desc->cmap_cout : s_desc->cout;
      const float *sc = s_desc->scale, *sh = s_desc->shift;
      for (int c = tid; c < nss && c < CH_SS_FLOATS; c += CH_THREADS) {
        s_ss[c] = sc ? __ldg(sc + c) : 1.f;
        s_ss[CH_SS_FLOATS + c] = sh ? __ldg(sh + c) : 0.f;
      }
    }
    if (s_desc->barrier_before) {
      grid_barrier(gbar, bar_gen);                     // (only thread 0's copy of bar_gen is meaningful)
    } else {
      __syncthreads();
    }

    const int nb = d_nb0 + d_nb1;
    const int T = d_K * nb;                                            // stages of one full (offset, channel block) sweep
    const int64_t U = (int64_t)d_m_tiles * d_n_ntiles * d_nsplit;      // work units of the layer
    const int64_t u_begin = U * blockIdx.x / gridDim.x, u_end = U * (blockIdx.x + 1) / gridDim.x;
    const int per_z = d_m_tiles * d_n_ntiles;
    const uint32_t b_bytes = (uint32_t)d_nt * 128u;

    // item = 1 or 2 consecutive units (same split, same N tile, adjacent row tiles) sharing every weight tile
#define CH_FOR_ITEMS()                                                                                         \
    for (int64_t u = u_begin, _n; u < u_end; u += _n)                                                          \
      if (const int z = (int)(u / per_z), r_ = (int)(u - (int64_t)z * per_z), nti = r_ / d_m_tiles,            \
          m = r_ - nti * d_m_tiles, nsub = (d_nsub_max == 2 && u + 1 < u_end && m + 1 < d_m_tiles) ? 2 : 1,    \
          t_begin = min(z * d_sps, T), t_end = min(t_begin + d_sps, T);                                        \
          (_n = nsub, true))

    const long long _role_t0 = prof ? clock64() : 0;
    if (warp == CH_W_B) {
      // ============================ weight tiles ====================================
      const uint8_t *wtiles = s_desc->wtiles;
      CH_FOR_ITEMS() {
        (void)m;
        for (int t = t_begin; t < t_end; ++t) {
          { CH_PROF_BEGIN(); mbar_wait_relaxed(emptyB + 8 * b_slot, b_phase ^ 1, 64); CH_PROF_END(pw0); }
          if (elect_one()) {
            const uint32_t fb = fullB + 8 * b_slot;
            if (flags & 0x200) {                      // tuning: no weight loads
              mbar_expect_tx(fb, 0u);
            } else {
              mbar_expect_tx(fb, b_bytes);
              bulk_g2s(b_ring_k + b_slot * (uint32_t)bslot, wtiles + ((int64_t)t * d_n_ntiles + nti) * b_bytes, b_bytes, fb);
            }
          }
          __syncwarp();
          if (++b_slot == (uint32_t)sb) { b_slot = 0; b_phase ^= 1; }
        }
      }
    } else if (warp == CH_W_MMA || warp == CH_W_MMA + 1) {
      // ============ MMA issuers: warp CH_W_MMA owns sub-tile 0 of every item, the next warp sub-tile 1 ===============
      // One issuing thread spends ~64 cycles per tcgen05.mma plus ~400 cycles of barrier-wait / fence / commit per row
      // slot, more than the 288 cycles of tensor work a 96-channel slot carries; two issuers on disjoint accumulator
      // columns restore the slack two co-resident CTAs used to give.  Each sub-tile's MMAs are issued by one thread, in
      // stage order (bit-reproducible accumulation).
      const int mi = warp - CH_W_MMA;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d_nt >> 3) << 17) | ((uint32_t)(CH_M >> 4) << 24);
      CH_FOR_ITEMS() {
        (void)m; (void)nti;
        const uint32_t buf = n_item & 1u;
        const bool mine = mi < nsub;
        // Both issuers follow the full protocol of every item, also the one without a sub-tile of its own (single-sub-tile
        // items): its arrivals on emptyB / accFull may only happen in the phase they belong to, i.e. after the same waits.
        { CH_PROF_BEGIN(); mbar_wait(accEmpty + 8 * buf, ((n_item >> 1) & 1u) ^ 1u); CH_PROF_END(pw2); }   // the epilogue drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dcol = tmem_base + buf * 256u + (uint32_t)mi * 128u;
        // two stages per iteration: the fixed cost of an iteration (waits, proxy fence, election, descriptor set-up) is a
        // dependent chain of a few hundred cycles; 12 MMAs behind it instead of 6
        for (int t = t_begin; t < t_end;) {
          const int nst = min(2, t_end - t);
          uint32_t sl[2], sph[2], bs[2], bph[2];
          {
            uint32_t as_ = a_slot + (uint32_t)mi, ap_ = a_phase, bs_ = b_slot, bp_ = b_phase;
            if (as_ >= (uint32_t)sa) { as_ -= (uint32_t)sa; ap_ ^= 1u; }
#pragma unroll
            for (int jx = 0; jx < 2; ++jx) {
              sl[jx] = as_; sph[jx] = ap_; bs[jx] = bs_; bph[jx] = bp_;
              as_ += (uint32_t)nsub; if (as_ >= (uint32_t)sa) { as_ -= (uint32_t)sa; ap_ ^= 1u; }
              if (++bs_ == (uint32_t)sb) { bs_ = 0; bp_ ^= 1u; }
            }
          }
          if (mine) {
            {                                         // all barriers of the batch probed together (overlapping round trips)
              CH_PROF_BEGIN();
              const bool two = nst == 2;
              const uint32_t b1 = two ? bs[1] : bs[0], bp1 = two ? bph[1] : bph[0], a1 = two ? sl[1] : sl[0], ap1 = two ? sph[1] : sph[0];
              for (uint32_t it = 0;; ++it) {
                const uint32_t ok = mbar_try(fullB + 8 * bs[0], bph[0]) & mbar_try(fullA + 8 * sl[0], sph[0]) &
                                    mbar_try(fullB + 8 * b1, bp1) & mbar_try(fullA + 8 * a1, ap1);
                if (ok) break;
                if (it > (1u << 26)) __trap();
              }
              CH_PROF_END(pw1);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) writes -> UMMA reads
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (flags & 0x4000) {                     // tuning: plain arrivals instead of tcgen05.commit (only without MMAs)
              if (lane == 0) {
#pragma unroll
                for (int jx = 0; jx < 2; ++jx)
                  if (jx < nst) { mbar_arrive(emptyA + 8 * sl[jx]); mbar_arrive(emptyB + 8 * bs[jx]); }
              }
            } else if (elect_one()) {
#pragma unroll
              for (int jx = 0; jx < 2; ++jx) {
                if (jx < nst) {
                  const uint64_t db = umma_desc(b_ring_k + bs[jx] * (uint32_t)bslot);
                  const uint64_t da = umma_desc(a_ring_k + sl[jx] * (uint32_t)CH_A_BYTES);
                  // 128-byte line = [hi ch0-15 | hi ch16-31 | lo ch0-15 | lo ch16-31]; +2 per 32-byte K slice
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    if (flags & 0x400) break;         // tuning: no MMAs
                    umma_bf16(dcol, da + 2 * h, db + 2 * h, idesc, (h == 0 && jx == 0 && t == t_begin) ? 0u : 1u);   // hi * Whi
                    umma_bf16(dcol, da + 2 * h, db + 2 * h + 4, idesc, 1u);                                          // hi * Wlo
                    umma_bf16(dcol, da + 2 * h + 4, db + 2 * h, idesc, 1u);                                          // lo * Whi
                  }
                  umma_commit(emptyA + 8 * sl[jx]);                         // row slot free when these MMAs retire
                  umma_commit(emptyB + 8 * bs[jx]);                         // weight slot: one arrival per issuer
                }
              }
            }
          } else {
#pragma unroll
            for (int jx = 0; jx < 2; ++jx) {
              if (jx < nst) {
                mbar_wait(fullB + 8 * bs[jx], bph[jx]);                     // stay in step with the slot's phase ...
                if (lane == 0) mbar_arrive(emptyB + 8 * bs[jx]);            // ... nothing of mine reads this weight tile
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int jx = 0; jx < 2; ++jx) {
            if (jx < nst) {
              a_slot += (uint32_t)nsub;
              if (a_slot >= (uint32_t)sa) { a_slot -= (uint32_t)sa; a_phase ^= 1u; }
              if (++b_slot == (uint32_t)sb) { b_slot = 0; b_phase ^= 1; }
            }
          }
          t += nst;
        }
        if (mine) { if (elect_one()) umma_commit(accFull + 8 * buf); }
        else if (lane == 0) mbar_arrive(accFull + 8 * buf);
        __syncwarp();
        ++n_item;
      }
    } else if (warp >= CH_W_A) {
      // ================= gathered A rows: warp w fills every CH_A_WARPS-th row slot, all 128 rows of it ====================
      // 8 lanes cover one 128-byte row line (one L2 line), 4 rows per copy instruction, 32 instructions per slot behind ONE
      // barrier wait and ONE (self-tracking) arrival; the destination carries the 128B swizzle (chunk ^ (row & 7)); a
      // missing neighbour is a zero-fill copy.  The slot's 128 row indices are four coalesced loads (lane l: rows l, l+32, ..)
      // fetched one slot of this warp ahead and handed round by shuffles.
      const int w = warp - CH_W_A, j = lane & 7, q = lane >> 3;
      const int32_t *nbr = s_desc->nbr;
      const uint8_t *src0 = s_desc->src0 + j * 16, *src1 = s_desc->src1 + j * 16;
      const uint32_t rb0 = (uint32_t)d_nb0 * 128u, rb1 = (uint32_t)d_nb1 * 128u;
      const uint32_t off_even = (uint32_t)q * 128u + (uint32_t)((j ^ q) << 4);                 // rows 8n + q
      const uint32_t off_odd = (uint32_t)(4 + q) * 128u + (uint32_t)((j ^ (4 + q)) << 4);      // rows 8n + 4 + q
      // Ring slot s is always filled by warp s % CH_A_WARPS, lap after lap: a warp cannot run a lap ahead of itself, so the
      // parity of a slot's empty barrier is unambiguous (with slots dealt round-robin over the warps, a fast warp one lap
      // ahead of a slow one passed the parity test early and overwrote rows that had not been multiplied yet).
      CH_FOR_ITEMS() {
        (void)nti;
        const uint32_t n_slots = (uint32_t)((t_end - t_begin) * nsub), g_end = g_slot + n_slots;
        auto decode = [&](uint32_t jl_, int &k_, int &cb_, int &s_) {
          const int st = (int)jl_ / nsub;             // stage-major, sub-tile-minor: the order the issuers consume slots in
          s_ = (int)jl_ - st * nsub;
          const int tt = t_begin + st;
          k_ = tt / nb; cb_ = tt - k_ * nb;
        };
        auto fetch = [&](int k_, int s_, int32_t (&r)[4]) {
          const int32_t *nk = nbr ? nbr + (int64_t)k_ * d_n_out : nullptr;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int64_t o = (int64_t)(m + s_) * CH_M + 32 * i + lane;
            r[i] = (o < d_n_out) ? (nk ? __ldg(nk + o) : (int32_t)o) : -1;
          }
        };
        int32_t nxt[4];
        int k_n = 0, cb_n = 0, s_n = 0;
        const bool owner = (uint32_t)w < (uint32_t)sa;
        if (owner && p_lapb + p_sl < g_end) { decode(p_lapb + p_sl - g_slot, k_n, cb_n, s_n); fetch(k_n, s_n, nxt); }
        while (owner && p_lapb + p_sl < g_end) {
          int32_t cur[4];
          {
            CH_PROF_BEGIN();
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
            if (prof) { if (cur[0] + cur[1] + cur[2] + cur[3] == 0x7fffffff) pw2 += 1; }   // force the loads to land here
            CH_PROF_END(pw1);
          }
          const int cb = cb_n;
          const uint32_t sl = p_sl, par = p_par, jl = p_lapb + p_sl - g_slot;
          // my next slot (same warp, CH_A_WARPS slots on or the next lap)
          p_sl += CH_A_WARPS;
          if (p_sl >= (uint32_t)sa) { p_sl = (uint32_t)w; p_lapb += (uint32_t)sa; p_par ^= 1u; }
          if (p_lapb + p_sl < g_end) { decode(p_lapb + p_sl - g_slot, k_n, cb_n, s_n); fetch(k_n, s_n, nxt); }   // its row indices, ahead of time
          const bool first = cb < d_nb0;
          const uint8_t *src = (first ? src0 : src1) + (first ? cb : cb - d_nb0) * 128;
          const uint32_t rb = first ? rb0 : rb1;
          { CH_PROF_BEGIN(); chain_wait(emptyA + 8 * sl, par ^ 1u, 4 | ((int)jl << 8)); CH_PROF_END(pw0); }
          const uint32_t a_dst = a_ring_k + sl * (uint32_t)CH_A_BYTES;
          if (!(flags & 0x100)) {                     // tuning: bit 8 = no row copies
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {          // 8 copy instructions at a time: shuffles first, then the copies
              int32_t r8[8];
#pragma unroll
              for (int ii = 0; ii < 8; ++ii) r8[ii] = __shfl_sync(0xffffffffu, cur[g8], (4 * ii + q) & 31);   // rows 32 g8 + 4 ii + q
#pragma unroll
              for (int ii = 0; ii < 8; ++ii) {
                const int i = g8 * 8 + ii;
                const uint32_t rr = r8[ii] < 0 ? 0u : (uint32_t)r8[ii];
                // src-size form (16 or 0 bytes): the copy engine itself writes the zeros of a missing neighbour, so they are
                // covered by the self-tracking arrival below.  (The ignore-src predicate form produced intermittently stale
                // rows in the slot that is consumed right after it is filled: profiles/r02_chain_roles.md.)
                cp_async16(a_dst + (uint32_t)(i >> 1) * 1024u + ((i & 1) ? off_odd : off_even), src + (uint64_t)rr * rb,
                           r8[ii] < 0 ? 0u : 16u);
              }
            }
          }
          cp_async_arrive_noinc(fullA + 8 * sl);      // 32 self-tracking arrivals, fired by the copy engine
        }
        g_slot = g_end;
      }
    } else if (warp < 4) {
      // ================= epilogue: four warps drain both accumulator buffers in turn ====================
      // (their own work is ~5% of a layer; the other four warps of the former second group gather rows now)
      const int q = warp & 3;                         // TMEM lane quarter this warp may access
      const uint32_t stgw = stg_u32 + (uint32_t)warp * 4096u;
      const int rsub = lane >> 3, chunk = lane & 7, sw = lane & 7;
      const uint32_t my_line = stgw + lane * 128;
      const int d_cout = s_desc->cout, d_cout_pad = s_desc->cout_pad, d_relu = s_desc->relu, d_cmap_cout = s_desc->cmap_cout;
      const bool has_scale = s_desc->scale != nullptr;
      const uint8_t *d_res = s_desc->res;
      uint8_t *d_out_split = s_desc->out_split;
      float *d_out_f32 = s_desc->out_f32, *d_partial = s_desc->partial;
      const int32_t *d_row_map = s_desc->out_row_map, *d_cmap = s_desc->cmap;
      auto lds128 = [](uint32_t a) { uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; };
      auto sts128 = [](uint32_t a, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); };
      CH_FOR_ITEMS() {
        const uint32_t buf = n_item & 1u;
        { CH_PROF_BEGIN(); mbar_wait_relaxed(accFull + 8 * buf, (n_item >> 1) & 1u, 128); CH_PROF_END(pw0); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int s = 0; s < nsub; ++s) {
          const int64_t wrow0 = (int64_t)(m + s) * CH_M + q * 32;        // first global row of this warp
          const int64_t o = wrow0 + lane;
          int32_t my_orow = (int32_t)min(o, d_n_out - 1);
          if (d_row_map && o < d_n_out) my_orow = __ldg(d_row_map + o);
          // staged tile (32 rows x 128 B, swizzled) -> global, 4 full lines per instruction
          auto flush_tile = [&](uint8_t *base, int64_t row_bytes, int64_t col_byte, bool mapped) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + rsub;
              v[i] = lds128(stgw + r * 128 + ((chunk ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + rsub;
              const int32_t mo = __shfl_sync(0xffffffffu, my_orow, r);
              const int64_t grow = mapped ? (int64_t)mo : wrow0 + r;
              if (wrow0 + r < d_n_out && grow >= 0 && !(flags & 0x800))
                *reinterpret_cast<uint4 *>(base + grow * row_bytes + col_byte + chunk * 16) = v[i];
            }
          };
          for (int cbo = 0; cbo < d_nt / 32; ++cbo) {
            float y[32];
            {
              uint32_t v0[16], v1[16];
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)s * 128u + cbo * 32;
              tmem_ld16(taddr, v0);
              tmem_ld16(taddr + 16, v1);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) { y[jj] = __uint_as_float(v0[jj]); y[16 + jj] = __uint_as_float(v1[jj]); }
            }
            const int c0 = nti * d_nt + cbo * 32;             // first output channel of this 32-block
            if (d_nsplit > 1) {                               // raw partial sums; the reduce phase applies the epilogue
#pragma unroll
              for (int g = 0; g < 8; ++g)
                sts128(my_line + ((g ^ sw) << 4), make_uint4(__float_as_uint(y[4 * g]), __float_as_uint(y[4 * g + 1]),
                                                             __float_as_uint(y[4 * g + 2]), __float_as_uint(y[4 * g + 3])));
              __syncwarp();
              flush_tile(reinterpret_cast<uint8_t *>(d_partial + (int64_t)z * d_n_out * d_cout_pad), (int64_t)d_cout_pad * 4,
                         (int64_t)c0 * 4, false);
              __syncwarp();
              continue;
            }
            if (c0 >= d_cout) continue;                       // warp-uniform (padding columns)
            int oc0 = c0, out_c = d_cout, kch = 0;
            if (d_cmap) {                                     // dense transposed conv: this column block belongs to child kch
              kch = c0 / d_cmap_cout;
              oc0 = c0 - kch * d_cmap_cout;
              out_c = d_cmap_cout;
              my_orow = (o < d_n_out) ? __ldg(d_cmap + (int64_t)kch * d_n_out + o) : -1;
            }
            if (has_scale) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) y[jj] = fmaf(y[jj], s_ss[oc0 + jj], s_ss[CH_SS_FLOATS + oc0 + jj]);
            }
            if (d_res) {                                      // residual tile: coalesced load -> staging -> own row
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + rsub;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (wrow0 + r < d_n_out)
                  v = __ldcg(reinterpret_cast<const uint4 *>(d_res + (wrow0 + r) * (int64_t)d_cout * 4 + (c0 >> 5) * 128 + chunk * 16));
                sts128(stgw + r * 128 + ((chunk ^ (r & 7)) << 4), v);
              }
              __syncwarp();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 hq = lds128(my_line + ((g ^ sw) << 4)), lq = lds128(my_line + (((4 + g) ^ sw) << 4));
                const __nv_bfloat16 *hh = reinterpret_cast<const __nv_bfloat16 *>(&hq);
                const __nv_bfloat16 *ll = reinterpret_cast<const __nv_bfloat16 *>(&lq);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) y[g * 8 + jj] += join_bf16(hh[jj], ll[jj]);
              }
              __syncwarp();
            }
            if (d_relu) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) y[jj] = fmaxf(y[jj], 0.f);
            }
            if (d_out_split) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                __align__(16) __nv_bfloat16 hh[8], ll[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) split_bf16(y[g * 8 + jj], hh[jj], ll[jj]);
                sts128(my_line + ((g ^ sw) << 4), *reinterpret_cast<const uint4 *>(hh));
                sts128(my_line + (((4 + g) ^ sw) << 4), *reinterpret_cast<const uint4 *>(ll));
              }
              __syncwarp();
              flush_tile(d_out_split, (int64_t)out_c * 4, (int64_t)(oc0 >> 5) * 128, d_cmap != nullptr);
              __syncwarp();
            }
            if (d_out_f32) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                sts128(my_line + ((g ^ sw) << 4), make_uint4(__float_as_uint(y[4 * g]), __float_as_uint(y[4 * g + 1]),
                                                             __float_as_uint(y[4 * g + 2]), __float_as_uint(y[4 * g + 3])));
              __syncwarp();
              flush_tile(reinterpret_cast<uint8_t *>(d_out_f32), (int64_t)out_c * 4, (int64_t)oc0 * 4,
                         d_row_map != nullptr || d_cmap != nullptr);
              __syncwarp();
            }
          }
        }
        // this warp's TMEM reads of the buffer are complete (tcgen05.wait::ld above): hand it back to the MMA warps
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(accEmpty + 8 * buf);
        ++n_item;
      }
    }
#undef CH_FOR_ITEMS
    if (prof) pt += clock64() - _role_t0;

    if (d_nsplit > 1) {
      // ---- split-K: every partial is in global memory after this barrier; reduce + epilogue by all threads of the grid
      grid_barrier(gbar, bar_gen);
      const ConvDesc &d = args.d[L];
      const int groups = d.cout / 8;
      const int64_t total = d.n_out * groups;
      for (int64_t e = (int64_t)blockIdx.x * CH_THREADS + tid; e < total; e += (int64_t)gridDim.x * CH_THREADS) {
        const int64_t o = e / groups;
        const int c0 = (int)(e - o * groups) * 8;
        float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int z = 0; z < d.nsplit; ++z) {
          const float4 *pp = reinterpret_cast<const float4 *>(d.partial + ((int64_t)z * d.n_out + o) * d.cout_pad + c0);
          const float4 a = __ldcg(pp), b = __ldcg(pp + 1);
          y[0] += a.x; y[1] += a.y; y[2] += a.z; y[3] += a.w; y[4] += b.x; y[5] += b.y; y[6] += b.z; y[7] += b.w;
        }
        if (d.scale) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) y[jj] = fmaf(y[jj], s_ss[c0 + jj], s_ss[CH_SS_FLOATS + c0 + jj]);
        }
        const int64_t off = o * (int64_t)d.cout * 4 + split_off_hi(c0);
        if (d.res) {
          const uint4 hq = __ldcg(reinterpret_cast<const uint4 *>(d.res + off)), lq = __ldcg(reinterpret_cast<const uint4 *>(d.res + off + 64));
          const __nv_bfloat16 *hh = reinterpret_cast<const __nv_bfloat16 *>(&hq);
          const __nv_bfloat16 *ll = reinterpret_cast<const __nv_bfloat16 *>(&lq);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) y[jj] += join_bf16(hh[jj], ll[jj]);
        }
        if (d.relu) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) y[jj] = fmaxf(y[jj], 0.f);
        }
        if (d.out_split) {
          __align__(16) __nv_bfloat16 hh[8], ll[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) split_bf16(y[jj], hh[jj], ll[jj]);
          *reinterpret_cast<uint4 *>(d.out_split + off) = *reinterpret_cast<const uint4 *>(hh);
          *reinterpret_cast<uint4 *>(d.out_split + off + 64) = *reinterpret_cast<const uint4 *>(ll);
        }
        if (d.out_f32) {
          const int64_t orow = d.out_row_map ? (int64_t)__ldg(d.out_row_map + o) : o;
          float4 *op = reinterpret_cast<float4 *>(d.out_f32 + orow * d.cout + c0);
          op[0] = make_float4(y[0], y[1], y[2], y[3]);
          op[1] = make_float4(y[4], y[5], y[6], y[7]);
        }
      }
    }
  }

  if (dbg_clock && tid == 0) dbg_clock[blockIdx.x * 32 + 2] = clock64();
  if (dbg_clock && lane == 0 && (warp == CH_W_B || warp == CH_W_MMA || warp == CH_W_MMA + 1 || warp == CH_W_A || warp == 0)) {
    // rows of 4: [wait kind 0, wait kind 1, wait kind 2, role loop total]; B producer 4.., issuer0 8.., issuer1 12.., A producer 16.., epilogue 20..
    const int base = warp == CH_W_B ? 4 : warp == CH_W_MMA ? 8 : warp == CH_W_MMA + 1 ? 12 : warp == CH_W_A ? 16 : 20;
    long long *o = dbg_clock + blockIdx.x * 32 + base;
    o[0] = pw0; o[1] = pw1; o[2] = pw2; o[3] = pt;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == CH_W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// ------------------------------------------------------------ tile-major, pre-swizzled weight packing
// wtiles[((k*nb + cb) * n_ntiles + nti)] = nt rows (output channels) x 128 B, each row the split line
// [hi ch0-15 | hi ch16-31 | lo ch0-15 | lo ch16-31] of input channels [32cb, 32cb+32), with the 16-byte chunks of row n
// stored at chunk ^ (n & 7): exactly what a 128B-swizzled TMA tile load would leave in shared memory, so that one
// linear cp.async.bulk per stage fetches a ready-to-multiply B operand (no tensor map, nothing to encode per launch).
__global__ void k_pack_weight_tiles(const float *__restrict__ w, int K, int cin, int cout, int cout_pad, int nt, int transpose_w,
                                    uint8_t *__restrict__ wt) {
  const int nb = cin / 32, n_ntiles = cout_pad / nt;
  const int64_t total = (int64_t)K * cout_pad * cin;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % cin);
    const int64_t kn = e / cin;
    const int ng = (int)(kn % cout_pad), k = (int)(kn / cout_pad);
    float v = 0.f;
    if (ng < cout) v = transpose_w ? w[((int64_t)k * cout + ng) * cin + c] : w[((int64_t)k * cin + c) * cout + ng];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const int cb = c >> 5, ci = c & 31, nti = ng / nt, n = ng - nti * nt;
    uint8_t *tile = wt + ((((int64_t)k * nb + cb) * n_ntiles) + nti) * (int64_t)nt * 128;
    const int jh = ci >> 3, el = ci & 7;
    *reinterpret_cast<__nv_bfloat16 *>(tile + n * 128 + (((jh) ^ (n & 7)) << 4) + el * 2) = hi;
    *reinterpret_cast<__nv_bfloat16 *>(tile + n * 128 + (((4 + jh) ^ (n & 7)) << 4) + el * 2) = lo;
  }
}

}  // namespace osb

namespace osb { bool conv_tc_tuning(const char *name, int64_t v); }
using namespace osb;

static inline int chain_cout_pad(int cout) { return cout <= 256 ? (cout + 15) / 16 * 16 : (cout + 255) / 256 * 256; }
static inline int chain_nt(int cout) { const int cp = chain_cout_pad(cout); return cp <= 256 ? cp : 256; }

extern "C" {

size_t osb_conv_desc_bytes(void) { return sizeof(ConvDesc); }

size_t osb_conv_weight_tiles_bytes(int32_t K, int32_t cin, int32_t cout) { return (size_t)K * chain_cout_pad(cout) * cin * 4; }

int osb_conv_pack_weight_tiles(const float *w, int32_t K, int32_t cin, int32_t cout, int32_t transpose_w, void *wtiles, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(K >= 1 && cin % 32 == 0 && cin > 0 && cout > 0, "osb_conv_pack_weight_tiles: cin (%d) must be a multiple of 32", cin);
  const int cp = chain_cout_pad(cout), nt = chain_nt(cout);
  const int64_t total = (int64_t)K * cp * cin;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 32);
  k_pack_weight_tiles<<<grid, 256, 0, stream>>>(w, K, cin, cout, cp, nt, transpose_w, (uint8_t *)wtiles);
  OSB_LAUNCH_CHECK();
  return 0;
}

// Split factor of one layer when it runs on a grid of `grid_ctas` CTAs.  Base rule: as many splits as it takes to give every
// CTA a unit (grid / tiles).  That rule leaves half of the SMs idle on 76-tile levels (9.7 k rows: one unit of 108 stages per
// busy CTA), so a small cost model may RAISE the factor when it predicts a clear gain (it never lowers it: for the smallest
// levels the measured optimum is the base rule):
//   main loop   per CTA (pairs of units x ~0.75 us + single units x ~0.6 us) x stages per split; an item is two row-adjacent
//               units when the N tile is <= 128 wide (two issuers run them side by side); 1.6x for 256-wide N tiles
//   split cost  ~10 us (grid barrier, partial tiles out, reduce pass) + the partials written and read once at ~5 TB/s (L2)
// Constants fitted on the per-layer times of the bench scene (scripts/layer_times.py with OSB_CHAIN_MAX_TILES=0).
static int chain_nsplit(int64_t n_out, int K, int cin, int cout, int grid_ctas, int force, int nsub_knob) {
  const int cp = chain_cout_pad(cout), nt = chain_nt(cout);
  const int64_t tiles = ceil_div(n_out, CH_M) * (cp / nt);
  const int T = K * (cin / 32);
  const int cap = std::max(1, std::min(32, T));
  if (force > 0) return std::min(force, cap);
  const bool pairs = nt <= 128 && nsub_knob >= 2;
  const int base = (int)std::max<int64_t>(1, std::min<int64_t>(grid_ctas / tiles, cap));
  auto cost = [&](int ns, bool &ok) {
    const int sps = (T + ns - 1) / ns;
    ok = (T + sps - 1) / sps == ns;                                // else this many splits would leave empty ones
    const int64_t upc = ceil_div(tiles * ns, (int64_t)grid_ctas);
    const double per = pairs ? (double)(upc / 2) * 0.75 + (double)(upc % 2) * 0.6 : (double)upc * 0.6 * 1.6;
    double us = per * sps + 4.0;
    if (ns > 1) {
      const double partial_bytes = (double)ns * (double)n_out * cp * 4.0;
      if (partial_bytes > 64e6) ok = false;                        // scratch stays small (the engine provides 96 MB per layer)
      us += 10.0 + 2.0 * partial_bytes / 5e6;
    }
    return us;
  };
  bool ok = true;
  double best = cost(base, ok);
  int best_ns = base;
  for (int ns = base + 1; ns <= cap; ++ns) {
    const double us = cost(ns, ok);
    if (ok && us < best - std::max(0.5, 0.08 * best)) { best = us; best_ns = ns; }
  }
  return best_ns;
}

static int g_chain_force_split = 0;      // tuning: > 0 forces the split factor of every layer (1 disables splitting)
static int g_chain_nsub = 2;             // tuning: 1 = never pair sub-tiles
static int g_chain_grid = 0;             // tuning: CTAs per launch (0 = one per SM)
static int g_chain_sa = 0, g_chain_sb = 0;   // tuning: ring depths (0 = as many row slots as fit / 3 or 2 weight slots)
static long long *g_chain_dbg_clock = nullptr;
static int g_chain_dbg_skip = 0;         // tuning: bit0 no row copies, bit1 no weight loads, bit2 no MMAs, bit3 no stores,
                                         // bit5 no tcgen05 fence, bit6 plain arrivals for commits, bit8 legacy (consumer-side) completion

int osb_tuning_set(const char *name, int64_t value) {
  const std::string n(name ? name : "");
  if (n == "chain_force_split") g_chain_force_split = (int)value;
  else if (n == "chain_nsub") g_chain_nsub = (int)value;
  else if (n == "chain_grid") g_chain_grid = (int)value;
  else if (n == "chain_sa") g_chain_sa = (int)value;
  else if (n == "chain_sb") g_chain_sb = (int)value;
  else if (n == "chain_dbg_clock") g_chain_dbg_clock = (long long *)(intptr_t)value;
  else if (n == "chain_dbg_skip") g_chain_dbg_skip = (int)value;
  else if (n == "chain_report") {               // host-mapped int64 buffer [1 + 4*1024]: where a stuck wait was (tuning)
    long long *ptr = (long long *)(intptr_t)value;
    OSB_CUDA(cudaMemcpyToSymbol(g_chain_report, &ptr, sizeof(ptr)));
  }
  else { OSB_CHECK(conv_tc_tuning(n.c_str(), value), "osb_tuning_set: unknown knob '%s'", n.c_str()); }
  return 0;
}

int osb_conv_chain_grid(void) {
  if (g_chain_grid > 0) return g_chain_grid;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

size_t osb_conv_chain_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout) {
  if (n_out <= 0 || K < 1 || cin < 32 || cout <= 0) return 0;     // shapes osb_conv_desc_fill rejects: nothing to reserve
  const int ns = chain_nsplit(n_out, K, cin, cout, osb_conv_chain_grid(), g_chain_force_split, g_chain_nsub);
  return ns > 1 ? (size_t)ns * n_out * chain_cout_pad(cout) * sizeof(float) : 0;
}

int osb_conv_desc_fill(void *desc_host, const void *src0, int32_t c0, const void *src1, int32_t c1, const int32_t *nbr, int64_t n_out,
                       int32_t K, const void *wtiles, int32_t cout, const float *scale, const float *shift, const void *res,
                       int32_t relu, void *out_split, float *out_f32, const int32_t *out_row_map, const int32_t *cmap,
                       int32_t cmap_cout, void *ws, size_t ws_bytes, int32_t barrier_before) {
  OSB_CHECK(desc_host != nullptr, "osb_conv_desc_fill: no descriptor");
  OSB_CHECK(src0 && c0 > 0 && c0 % 32 == 0 && c1 >= 0 && c1 % 32 == 0, "osb_conv_desc_fill: channel counts must be multiples of 32 (c0=%d c1=%d)", c0, c1);
  OSB_CHECK((c1 == 0) == (src1 == nullptr), "osb_conv_desc_fill: src1 / c1 mismatch");
  OSB_CHECK(K >= 1 && K <= 32, "osb_conv_desc_fill: K=%d not supported (<= 32)", K);
  OSB_CHECK(nbr != nullptr || K == 1, "osb_conv_desc_fill: identity map needs K == 1");
  OSB_CHECK(n_out > 0 && n_out < (1ll << 31), "osb_conv_desc_fill: bad row count");
  OSB_CHECK(cout % 32 == 0 && cout > 0, "osb_conv_desc_fill: cout (%d) must be a multiple of 32", cout);
  OSB_CHECK(out_split || out_f32, "osb_conv_desc_fill: no output given");
  OSB_CHECK((scale == nullptr) == (shift == nullptr), "osb_conv_desc_fill: scale and shift go together");
  OSB_CHECK(cmap == nullptr || (cmap_cout > 0 && cmap_cout % 32 == 0 && cout % cmap_cout == 0 && res == nullptr),
            "osb_conv_desc_fill: bad dense-transpose arguments");
  OSB_CHECK((cmap ? cmap_cout : cout) <= CH_SS_FLOATS, "osb_conv_desc_fill: more than %d output channels per row", CH_SS_FLOATS);
  ConvDesc d{};
  const int cin = c0 + c1;
  d.src0 = (const uint8_t *)src0; d.src1 = (const uint8_t *)src1; d.nbr = nbr; d.wtiles = (const uint8_t *)wtiles;
  d.scale = scale; d.shift = shift; d.res = (const uint8_t *)res; d.out_split = (uint8_t *)out_split; d.out_f32 = out_f32;
  d.out_row_map = out_row_map; d.cmap = cmap; d.n_out = n_out; d.K = K; d.nb0 = c0 / 32; d.nb1 = c1 / 32;
  d.cout = cout; d.cout_pad = chain_cout_pad(cout); d.nt = chain_nt(cout); d.n_ntiles = d.cout_pad / d.nt;
  d.relu = relu; d.cmap_cout = cmap_cout; d.m_tiles = (int)ceil_div(n_out, CH_M);
  d.nsub_max = (d.nt <= 128 && g_chain_nsub >= 2) ? 2 : 1;
  d.barrier_before = barrier_before ? 1 : 0;
  d.nsplit = cmap ? 1 : chain_nsplit(n_out, K, cin, cout, osb_conv_chain_grid(), g_chain_force_split, g_chain_nsub);
  const int T = K * (cin / 32);
  d.stages_per_split = (T + d.nsplit - 1) / d.nsplit;
  d.nsplit = (T + d.stages_per_split - 1) / d.stages_per_split;            // no empty splits
  d.partial = nullptr;
  if (d.nsplit > 1) {
    const size_t need = (size_t)d.nsplit * n_out * d.cout_pad * sizeof(float);
    OSB_CHECK(ws != nullptr && ws_bytes >= need, "osb_conv_desc_fill: workspace of %zu bytes required (got %zu)", need, ws_bytes);
    d.partial = (float *)ws;
  }
  memcpy(desc_host, &d, sizeof(d));
  return 0;
}

int osb_conv_chain_launch(const void *descs_host, int32_t n_layers, void *grid_barrier_dev, int32_t flags, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OSB_CHECK(descs_host && n_layers >= 1, "osb_conv_chain_launch: bad arguments");
  const ConvDesc *h = (const ConvDesc *)descs_host;
  for (int i = 0; i < n_layers; ++i) {
    OSB_CHECK(i > 0 || !h[i].barrier_before, "osb_conv_chain_launch: the first layer of a launch cannot ask for a barrier");
    // a split layer's reduce phase is not followed by a barrier: the next layer may only reuse the partial buffer behind one
    OSB_CHECK(i == 0 || h[i].barrier_before || h[i].nsplit == 1 || h[i - 1].nsplit == 1 || h[i].partial != h[i - 1].partial,
              "osb_conv_chain_launch: layers %d and %d share a split workspace without a barrier between them", i - 1, i);
  }
  OSB_SMEM_ATTR_ONCE(k_conv_chain, 227 * 1024);
  const int grid = osb_conv_chain_grid();
  // the layer descriptors travel as kernel parameters: at most CH_MAX_LAYERS per launch, longer lists in several launches
  // (a launch boundary orders everything, so the first layer of a later launch needs no grid barrier)
  for (int l0 = 0; l0 < n_layers; l0 += CH_MAX_LAYERS) {
    const int cnt = std::min(CH_MAX_LAYERS, n_layers - l0);
    ChainArgs args;
    memcpy(args.d, h + l0, sizeof(ConvDesc) * cnt);
    args.d[0].barrier_before = 0;
    int nt_max = 0, need_bar = 0;
    for (int i = 0; i < cnt; ++i) {
      nt_max = std::max(nt_max, args.d[i].nt);
      need_bar |= (args.d[i].barrier_before || args.d[i].nsplit > 1);
    }
    OSB_CHECK(!need_bar || grid_barrier_dev != nullptr, "osb_conv_chain_launch: this chain needs the grid-barrier words");
    const int bslot = nt_max * 128;
    const int fixed = 1024 + CH_STG_BYTES + 2 * CH_SS_FLOATS * 4 + 40 * 8 + 64;   // alignment slack, staging, BN constants, barriers
    int sb = g_chain_sb > 0 ? g_chain_sb : (bslot >= 32768 ? 2 : 3);
    sb = std::min(sb, CH_MAX_SB);
    int sa = (227 * 1024 - fixed - sb * bslot) / CH_A_BYTES;
    if (g_chain_sa > 0) sa = std::min(sa, g_chain_sa);
    sa = std::min(sa, CH_MAX_SA);
    // An EVEN ring: with two sub-tiles per stage every row slot (and its two barriers) then belongs to one issuer for good.
    // With an odd ring the slots alternate between the issuers from lap to lap; a two-pipeline variant of this kernel produced
    // intermittently stale rows on hardware exactly then (and never with even rings): profiles/r02_chain_roles.md.
    sa &= ~1;
    OSB_CHECK(sa >= 4, "osb_conv_chain_launch: shared memory does not hold four row slots");
    const size_t smem_bytes = (size_t)sa * CH_A_BYTES + (size_t)sb * bslot + fixed;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = (flags & 1) ? 1 : 0;
    OSB_CUDA(cudaLaunchKernelEx(&cfg, k_conv_chain, args, (int)cnt, (unsigned *)grid_barrier_dev, sa, sb, bslot,
                                (int)((flags & 1) | (g_chain_dbg_skip << 8)), g_chain_dbg_clock));
    OSB_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
