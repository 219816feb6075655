// Implicit-GEMM convolution / linear layer for sm_100a.
//
//   D[128 x BLOCK_N] (fp32, TMEM) = sum over taps, 64-channel chunks of  A_tile[128 x 64] * W_tile[BLOCK_N x 64]^T
//
// * A tiles are fetched by TMA straight out of the NHWC bf16 activation tensor through a 5-D tiled tensor map
//   (c, w, phase, h, n): a 3x3 tap is just a (dw, dh) shift of the box origin and the conv zero padding is TMA's
//   out-of-bounds zero fill, so no im2col buffer ever exists.  Stride-2 convs use the (2C, W/2, 2, H/2, N) view of
//   the same memory (phase = input row parity, channel offset = input column parity).
// * W tiles come from the repacked weights [tap][Cout][Cin] (bf16) through a 3-D map.
// * Both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly.
// * Persistent kernel, one CTA per SM (CTA pairs / cta_group::2 for long-K GEMMs): warp 0 = TMA producer, warp 1 = MMA
//   issuer + TMEM owner (two accumulator buffers), warps 2..9 = epilogue (tcgen05.ld -> registers -> shared-memory
//   slot -> TMA store, with bias / per-sample bias / TMA-prefetched residual, fused GroupNorm statistics, the GEGLU
//   gate, or the Q|K / V^T split store).  See the comment on conv_gemm_kernel and DESIGN.md 3.1.
//
// Replaces (reference, PyTorch library calls): conv2d in ResBlock/Downsample/Upsample
// (ldm/modules/diffusionmodules/openaimodel.py:107-119,150-160,201-241), 1x1 proj_in/proj_out and all Linear layers of
// BasicTransformerBlock (ldm/modules/attention.py:38-65,198-230,270-297).
#include "internal.h"
#include "ptx.cuh"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

namespace pbe {

// Debug counters of CTA 0 (tools/gemm_probe.py): [0] MMA-warp cycles total, [1] waiting for smem stages (TMA),
// [2] waiting for a free TMEM buffer (epilogue), [3] producer cycles waiting for free stages, [4] k iterations.
__device__ long long g_gemm_dbg[16];   // [8..11]: epilogue warp per-chunk stages: ld+bias+add, residual+pack+stats, fence, chunks

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int NUM_SLOTS = 4;
constexpr int SLOT_BYTES = 128 * 32 * 4;  // one 32-column fp32 chunk of a 128-row tile
constexpr int EPI_STAGING_BYTES = NUM_SLOTS * SLOT_BYTES;
constexpr int BIAS_SCR_FLOATS = 128;                                  // per epilogue warp: this chunk's fused bias columns
constexpr int BIAS_SCR_BYTES = NUM_EPI_WARPS * BIAS_SCR_FLOATS * 4;   // 4 KB behind the barriers

__host__ __device__ constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// exact-erf GELU (torch.nn.functional.gelu default) for the GEGLU epilogue, division free:
//   erf(|g| / sqrt2) = 1 - 2^u(t),  t = min(|g|, 4 sqrt2),  u = t (c1 + t (c2 + t (c3 + t (c4 + t c5))))
// u is a weighted-minimax fit of log2(erfc(t / sqrt2)) (fit script: tools/fit_erf.py; max |erf error| 6.7e-7, max GELU
// error 1.3e-6 absolute -- bf16 stores the result with 2^-9 relative precision).  One MUFU.EX2 per output instead of the
// RCP + EX2 of Abramowitz-Stegun 7.1.26: the epilogue of the short-K GEGLU GEMMs was bound by the MUFU pipe.

// ---- packed fp32x2 arithmetic (FFMA2 on sm_100): two GEGLU outputs per instruction stream ----
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// NP output pairs at once, stage by stage: a[i] <- a[i] * gelu_erf(g[i])
template <int NP>
__device__ __forceinline__ void geglu_block(f32x2* a, const f32x2* g) {
  f32x2 t[NP], u[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float g0, g1;
    upk2(g[i], g0, g1);
    t[i] = pk2(fminf(fabsf(g0), 5.65685425f), fminf(fabsf(g1), 5.65685425f));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) u[i] = fma2(t[i], pk2(-5.204507615e-04f, -5.204507615e-04f), pk2(7.397474721e-03f, 7.397474721e-03f));
#pragma unroll
  for (int i = 0; i < NP; ++i) u[i] = fma2(u[i], t[i], pk2(-5.256118253e-02f, -5.256118253e-02f));
#pragma unroll
  for (int i = 0; i < NP; ++i) u[i] = fma2(u[i], t[i], pk2(-4.592547119e-01f, -4.592547119e-01f));
#pragma unroll
  for (int i = 0; i < NP; ++i) u[i] = fma2(u[i], t[i], pk2(-1.151091337e+00f, -1.151091337e+00f));
#pragma unroll
  for (int i = 0; i < NP; ++i) u[i] = mul2(u[i], t[i]);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float u0, u1;
    upk2(u[i], u0, u1);
    u[i] = pk2(ex2_approx(u0), ex2_approx(u1));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float e0, e1, g0, g1;
    upk2(u[i], e0, e1);
    upk2(g[i], g0, g1);
    const f32x2 hs = pk2(copysignf(fmaf(e0, -0.5f, 0.5f), g0), copysignf(fmaf(e1, -0.5f, 0.5f), g1));
    u[i] = add2(hs, pk2(0.5f, 0.5f));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) a[i] = mul2(a[i], mul2(g[i], u[i]));
}

// LayerNorm fold, writing side: add eight rounded 16-bit outputs of a row to its packed (sum, sum of squares) accumulators
__device__ __forceinline__ void ln_accumulate(const uint4& pk, f32x2& s2, f32x2& q2, int f16) {
  const float2 a = unpack_op2(pk.x, f16), b = unpack_op2(pk.y, f16), c = unpack_op2(pk.z, f16), d = unpack_op2(pk.w, f16);
  const f32x2 va = pk2(a.x, a.y), vb = pk2(b.x, b.y), vc = pk2(c.x, c.y), vd = pk2(d.x, d.y);
  s2 = add2(s2, va); q2 = fma2(va, va, q2);
  s2 = add2(s2, vb); q2 = fma2(vb, vb, q2);
  s2 = add2(s2, vc); q2 = fma2(vc, vc, q2);
  s2 = add2(s2, vd); q2 = fma2(vd, vd, q2);
}

// Work-unit coordinates (split slice, n tile, w / h / batch tile) advanced incrementally: unit += gridDim.x is a
// mixed-radix addition with carries (a dozen integer ops) instead of five runtime divisions per tile and role
// (ncu on the short-K GEMMs: index arithmetic was ~half of the epilogue's instruction stream).
struct TileCoord {
  int s, nt, wi, hi, ni;
};
struct TileStep {
  int S, NT, TW, TH;           // radices
  int ds, dnt, dwi, dhi, dni;  // digits of the stride gridDim.x
};
__device__ __forceinline__ TileCoord tile_coord_from_unit(int unit, const TileStep& r) {
  TileCoord c;
  c.s = unit % r.S; unit /= r.S;
  c.nt = unit % r.NT; unit /= r.NT;
  c.wi = unit % r.TW; unit /= r.TW;
  c.hi = unit % r.TH;
  c.ni = unit / r.TH;
  return c;
}
__device__ __forceinline__ void tile_coord_advance(TileCoord& c, const TileStep& r) {
  int carry;
  c.s += r.ds; carry = c.s >= r.S; c.s -= carry ? r.S : 0;
  c.nt += r.dnt + carry; carry = c.nt >= r.NT; c.nt -= carry ? r.NT : 0;
  c.wi += r.dwi + carry; carry = c.wi >= r.TW; c.wi -= carry ? r.TW : 0;
  c.hi += r.dhi + carry; carry = c.hi >= r.TH; c.hi -= carry ? r.TH : 0;
  c.ni += r.dni + carry;
}

// Persistent kernel: one CTA per SM walks work units (n-tile fastest, so co-running CTAs share A tiles and all weights
// in L2).  Two TMEM accumulator buffers: the MMA warp fills buffer (i+1)&1 while the 8 epilogue warps drain buffer i&1.
//
// CL == 2: the kernel runs as CTA pairs (2-CTA clusters, tcgen05 cta_group::2).  Both CTAs of a pair work on the same
// (split slice, n tile) and on two neighbouring m tiles (p.pair_dim picks the tile dimension that is paired): one
// M = 256 MMA, issued by the leader (rank 0), reads 128 activation rows and HALF of the weight tile from each CTA's
// shared memory and accumulates into both CTAs' TMEM.  Why: a B200 SM ingests ~64 B/clk from L2; a 128 x 160 tile needs
// (128 + 160) * 128 B per 64-deep k step = 576 clk of ingest for 320 clk of MMA (measured 555 clk per step, the MMA warp
// waiting on TMA 60 % of the time, tools/gemm_probe.py).  With the weight tile split over the pair each SM ingests
// (128 + 80) * 128 B = 416 clk per step, and the smaller stages allow a deeper ring.  (Multicasting the weight tile to
// both CTAs instead was measured slower than no cluster at all: it does not reduce what each SM has to ingest.)
//   full[stage]   lives in the leader: both producers' TMA loads complete_tx on it
//   empty[stage], tmem_full[buf]  exist in both CTAs: the leader's commits are multicast to the pair
//   tmem_empty[buf] lives in the leader: the epilogue warps of both CTAs arrive on it
template <int BLOCK_N, int STAGES, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO32,
                 const __grid_constant__ CUtensorMap tmO16, const __grid_constant__ ConvGemmParams p) {
  constexpr int B_STAGE_BYTES = (BLOCK_N / CL) * BLOCK_K * 2;  // a CTA pair splits the weight tile
  constexpr uint32_t TMEM_COLS = tmem_cols_for(2 * BLOCK_N);
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");

  griddep_launch_dependents();   // PDL: the next kernel may be scheduled behind this one; see griddep_wait below
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
  constexpr int PIPE_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  // after the pipeline stages: NUM_SLOTS epilogue slots of SLOT_BYTES (1024-B aligned), then the barriers
  const uint32_t sBar = smem_base + PIPE_BYTES + EPI_STAGING_BYTES;
  uint8_t* bar_gen = smem_gen + PIPE_BYTES + EPI_STAGING_BYTES;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int b) { return sBar + 8u * (2 * STAGES + b); };
  auto tmem_empty_bar = [&](int b) { return sBar + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_ptr_addr = sBar + 8u * (2 * STAGES + 4 + NUM_SLOTS);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(bar_gen + 8 * (2 * STAGES + 4 + NUM_SLOTS));

  const int warp = static_cast<int>(uniform_u32(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const int num_k_iters = p.num_taps * p.k_chunks;
  const int n_tiles = p.n_tiles;
  const int split_k = p.split_k;                                  // >1: partial sums go to a workspace slice
  // cluster-level work units: CTA `crank` of cluster `vblock` takes tile 2 * digit + crank along the paired dimension
  const int crank = CL == 2 ? static_cast<int>(blockIdx.x & 1u) : 0;
  const int vblock = static_cast<int>(blockIdx.x) / CL;
  const int vgrid = static_cast<int>(gridDim.x) / CL;
  const int mw = (CL == 2 && p.pair_dim == 0) ? 2 : 1, mh = (CL == 2 && p.pair_dim == 1) ? 2 : 1,
            mn = (CL == 2 && p.pair_dim == 2) ? 2 : 1;
  const int cw = mw == 2 ? crank : 0, ch = mh == 2 ? crank : 0, cn = mn == 2 ? crank : 0;
  const int total_tiles = (p.tiles_w * p.tiles_h * p.tiles_n / CL) * n_tiles * split_k;  // work units (tile x split)
  TileStep tstep;
  tstep.S = split_k; tstep.NT = n_tiles; tstep.TW = p.tiles_w / mw; tstep.TH = p.tiles_h / mh;
  {
    const TileCoord d = tile_coord_from_unit(vgrid, tstep);
    tstep.ds = d.s; tstep.dnt = d.nt; tstep.dwi = d.wi; tstep.dhi = d.hi; tstep.dni = d.ni;
  }
  auto k_range = [&](int sidx, int& k_begin, int& k_end) {
    k_begin = (num_k_iters * sidx) / split_k;
    k_end = (num_k_iters * (sidx + 1)) / split_k;
  };

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), CL * NUM_EPI_WARPS);
    }
    for (int sl = 0; sl < NUM_SLOTS; ++sl) mbar_init(sBar + 8u * (2 * STAGES + 4 + sl), 1);
    if (p.has_res) tma_prefetch_desc(&tmR);
    if (p.has_o32) tma_prefetch_desc(&tmO32);
    if (p.has_o16) tma_prefetch_desc(&tmO16);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CL == 2) {
      tmem_alloc_pair(tmem_ptr_addr, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr_addr, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_ptr_gen);
  // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail;
  // nothing below may touch global memory before the previous kernel has completed and flushed
  griddep_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    // The whole warp runs this loop with warp-uniform values and one elected lane issues: no divisions, no table
    // look-ups and no divergent single-lane region per k step (that version cost ~550 clk of dependent latency per
    // step -- more than the 320 clk of MMA it feeds -- and was THE bound of every long-K GEMM here).
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t lead_full0 = CL == 2 ? map_to_cta(full_bar(0), 0) : full_bar(0);  // barriers are 8 B apart
    const bool dbg = p.debug != 0 && blockIdx.x == 0;
    long long t_empty = 0;
    TileCoord tc = tile_coord_from_unit(vblock, tstep);
    for (int unit = vblock; unit < total_tiles; unit += vgrid, tile_coord_advance(tc, tstep)) {
      int k_begin, k_end;
      k_range(tc.s, k_begin, k_end);
      const int n_base = tc.nt * BLOCK_N + crank * (BLOCK_N / CL);  // a pair: my half of the weight tile
      const int w0 = (tc.wi * mw + cw) * p.tw;
      const int h0 = (tc.hi * mh + ch) * p.th;
      const int n0 = (tc.ni * mn + cn) * p.tn;
      int tap = k_begin / p.k_chunks;
      int kc = k_begin - tap * p.k_chunks;
      int c_off = p.tap_coff[tap] + kc * BLOCK_K, cw0 = w0 + p.tap_dw[tap], cph = p.tap_ph[tap], ch0 = h0 + p.tap_dh[tap];
      for (int it = k_begin; it < k_end; ++it) {
        const long long tw = dbg ? clock64() : 0;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (dbg) t_empty += clock64() - tw;
        // a pair accounts both CTAs' boxes on the leader's barrier
        if (CL == 1 || crank == 0) mbar_expect_tx_elect(full_bar(stage), CL * (A_STAGE_BYTES + B_STAGE_BYTES));
        const uint32_t fb = lead_full0 + 8u * stage;
        tma_load_5d_elect<CL == 2>(sA + stage * A_STAGE_BYTES, &tmA, fb, c_off, cw0, cph, ch0, n0);
        tma_load_3d_elect<CL == 2>(sB + stage * B_STAGE_BYTES, &tmB, fb, kc * BLOCK_K, n_base, tap);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        c_off += BLOCK_K;
        if (++kc == p.k_chunks && it + 1 < k_end) {
          kc = 0;
          ++tap;
          c_off = p.tap_coff[tap]; cw0 = w0 + p.tap_dw[tap]; cph = p.tap_ph[tap]; ch0 = h0 + p.tap_dh[tap];
        }
      }
    }
    if (dbg && lane == 0) g_gemm_dbg[3] = t_empty;
  } else if (warp == 1 && (CL == 1 || crank == 0)) {
    // ================= MMA issuer (pair: the leader CTA only) =================
    const uint32_t idesc = umma_idesc_f16(BLOCK_M * CL, BLOCK_N) | p.idesc_fmt;   // operand format bits from the plan
    int stage = 0;
    uint32_t phase = 0;
    int ti = 0;
    int s_idx = vblock % split_k;
    const int s_step = vgrid % split_k;
    const bool dbg = p.debug != 0 && blockIdx.x == 0;
    long long t_all = dbg ? clock64() : 0, t_full = 0, t_tmem = 0, n_it = 0;
    for (int unit = vblock; unit < total_tiles; unit += vgrid, ++ti) {
      const int buf = ti & 1;
      int k_begin, k_end;
      k_range(s_idx, k_begin, k_end);
      s_idx += s_step;
      if (s_idx >= split_k) s_idx -= split_k;
      long long tw = dbg ? clock64() : 0;
      mbar_wait(tmem_empty_bar(buf), ((ti >> 1) & 1) ^ 1u);  // epilogue has drained this accumulator
      if (dbg) t_tmem += clock64() - tw;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BLOCK_N;
      for (int it = k_begin; it < k_end; ++it) {
        if (dbg) { tw = clock64(); ++n_it; }
        mbar_wait(full_bar(stage), phase);
        if (dbg) t_full += clock64() - tw;
        tc_fence_after();
        {
          const uint64_t adesc = umma_desc_sw128(sA + stage * A_STAGE_BYTES);
          const uint64_t bdesc = umma_desc_sw128(sB + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the 16-B-unit start address field
            if (CL == 2) umma_bf16_ss_pair_elect(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (it > k_begin || k > 0) ? 1u : 0u);
            else umma_bf16_ss_elect(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (it > k_begin || k > 0) ? 1u : 0u);
          }
          // smem slot free once these MMAs retire; accumulator complete after the last k step (both CTAs of a pair)
          if (CL == 2) {
            umma_commit_pair_elect(empty_bar(stage), static_cast<uint16_t>(3));
            if (it == k_end - 1) umma_commit_pair_elect(tmem_full_bar(buf), static_cast<uint16_t>(3));
          } else {
            umma_commit_elect(empty_bar(stage));
            if (it == k_end - 1) umma_commit_elect(tmem_full_bar(buf));
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (dbg && lane == 0) {
      g_gemm_dbg[0] = clock64() - t_all;
      g_gemm_dbg[1] = t_full;
      g_gemm_dbg[2] = t_tmem;
      g_gemm_dbg[4] = n_it;
    }
  } else if (warp >= 2) {
    // ================= Epilogue (warps 2..9) =================
    // Two warp-sets of 4 warps (one warp per TMEM lane quarter). A warp-set handles every other 32-column chunk of a
    // tile (alternating per tile for balance) and owns two 16 KB shared-memory slots.  Per chunk:
    //   TMA has prefetched the fp32 residual box [128 rows x 32 cols] into the slot (128B swizzle)  ->  each thread
    //   (= one accumulator row) adds bias / per-sample bias / residual to its 32 TMEM columns  ->  writes the result
    //   back into the slot  ->  one elected thread issues the TMA store (rows / columns outside the tensor are clipped
    //   by the hardware) and, once the store has read the slot, the residual prefetch two chunks ahead.
    // No global load or store instruction is on this path except the rare bf16 side copy and the V^T scatter.
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int ws = ew >> 2;   // warp-set
    const int bar_id = 1 + ws;
    const bool elected = ((ew & 3) == 0) && lane == 0;
    const int row = q * 32 + lane;
    constexpr bool geglu = BLOCK_N == 256;   // the 256-wide tile exists for the GEGLU epilogue only (build_gemm_plan enforces it)
    // epilogue features as locals that fold to constants in the GEGLU kernel: its epilogue (16-bit output, bias, optional
    // LayerNorm fold -- nothing else) then compiles without the residual / fp32 / statistics / V^T paths and their registers
    const bool e_has_res = !geglu && p.has_res, e_has_o32 = !geglu && p.has_o32, e_has_o16 = geglu || p.has_o16;
    const bool e_qkv = !geglu && p.mode == EPI_QKV;
    float* const e_stats_out = geglu ? nullptr : p.stats_out;
    float2* const e_ln_stats_out = geglu ? nullptr : p.ln_stats_out;
    const int e_act = geglu ? 0 : p.act;
    const int chunk_cols = 32;
    const int nchunks = geglu ? (BLOCK_N / 2) / 32 : BLOCK_N / 32;
    const int out_cols_total = geglu ? p.n_total / 2 : (e_qkv ? p.qk_cols : p.n_total);
    const int tile_out_cols = geglu ? BLOCK_N / 2 : BLOCK_N;
    const uint32_t slot_base = smem_base + PIPE_BYTES;
    uint8_t* slot_gen_base = smem_gen + PIPE_BYTES;
    auto res_full_bar = [&](int s) { return sBar + 8u * (2 * STAGES + 4 + s); };
    // "accumulator drained" goes to the CTA that issues the MMAs
    const uint32_t lead_tmem_empty0 = CL == 2 ? map_to_cta(tmem_empty_bar(0), 0) : tmem_empty_bar(0);
    auto tmem_empty_arrive = [&](int b) {
      if (CL == 2) mbar_arrive_cluster(lead_tmem_empty0 + 8u * b);
      else mbar_arrive(tmem_empty_bar(b));
    };
    const float* __restrict__ rowbias = geglu ? nullptr : p.rowbias;
    bf16* __restrict__ out_bf16 = p.out_bf16;

    const int nb_pad = p.tiles_n * p.tn;
    auto tile_is_vt = [&](int n_tile) { return e_qkv && n_tile * BLOCK_N >= p.qk_cols; };
    auto chunk_valid = [&](int n_tile, int c) { return c < nchunks && n_tile * tile_out_cols + c * chunk_cols < out_cols_total; };

    // ---- prefetch iterator (elected thread): walks the chunks this warp-set will consume, in order ----
    int pf_ti = 0, pf_tile = vblock, pf_c = ws & 1, pf_seq = 0;
    TileCoord pf_tc = tile_coord_from_unit(vblock, tstep);
    auto pf_issue_next = [&]() {
      // find the next valid chunk at or after (pf_ti, pf_c)
      while (pf_tile < total_tiles) {
        const int n_tile = pf_tc.nt, w0 = (pf_tc.wi * mw + cw) * p.tw, h0 = (pf_tc.hi * mh + ch) * p.th,
                  n0 = (pf_tc.ni * mn + cn) * p.tn;
        if (!tile_is_vt(n_tile) && chunk_valid(n_tile, pf_c)) {
          const int slot = ws * 2 + (pf_seq & 1);
          if (e_has_res) {
            mbar_expect_tx(res_full_bar(slot), p.res16 ? SLOT_BYTES / 2 : SLOT_BYTES);   // 16-bit residual: 128 rows x 64 B
            tma_load_4d(slot_base + slot * SLOT_BYTES, &tmR, res_full_bar(slot), n_tile * tile_out_cols + pf_c * chunk_cols,
                        w0, h0, n0);
          } else {
            mbar_arrive(res_full_bar(slot));
          }
          ++pf_seq;
          pf_c += 2;
          return;
        }
        // advance to the next tile
        ++pf_ti;
        pf_tile += vgrid;
        tile_coord_advance(pf_tc, tstep);
        pf_c = (ws + pf_ti) & 1;
      }
    };
    if (elected && e_has_res) {
      pf_issue_next();
      pf_issue_next();
    }
    const bool store_warp = (ew & 3) == 0;

    // ---- fused bias columns of a chunk: lane = column, fetched one chunk ahead, broadcast through a per-warp scratch ----
    // (Every thread used to load the chunk's 32 bias + 32 per-sample bias values itself: 16 LDG.128 per chunk whose results the
    // adds consumed pair by pair -- a chain of global-load round trips in the middle of every chunk, and the reason the
    // short-K GEMMs were epilogue-bound at 2-4 k cycles per chunk, tools/unet_gemm_dbg.py.)  Needs the sample index to be
    // warp-uniform: a warp's 32 rows are 32 pixels of one sample whenever tw*th % 32 == 0 (or the tile is one sample).
    // BLOCK_N = 256 (the GEGLU tile; 5 pair stages leave no room) keeps its scratch in the upper half of the slot, which its
    // 16-bit-only epilogue never stages into; a plain epilogue on a 256-wide tile (op-level tests only) takes the per-thread loads
    constexpr bool DEDICATED_SCR = BLOCK_N != 256;
    float* wscr = reinterpret_cast<float*>(bar_gen + 8 * (2 * STAGES + 6 + NUM_SLOTS)) + ew * BIAS_SCR_FLOATS;
    const bool bias_uniform = DEDICATED_SCR && ((rowbias == nullptr) || p.tn == 1 || ((p.tw * p.th) & 31) == 0);
    const bool bias_any = (p.bias != nullptr) || (rowbias != nullptr);
    int bpf_key = -1;
    float bpf0 = 0.0f, bpf1 = 0.0f;
    const bool ln_on = p.ln_in != nullptr;
    // columns [col0, col0 + 32) (+ the gate columns col0 + BLOCK_N/2 for GEGLU) of sample on_w
    auto bias_fetch = [&](int key, int col0, int on_w) {
      bpf_key = key;
      bpf0 = 0.0f; bpf1 = 0.0f;
      const int col = col0 + lane;
      if (geglu) {
        bpf0 = __ldg(p.bias + col);
        bpf1 = __ldg(p.bias + col + BLOCK_N / 2);
      } else if (col < p.n_total) {
        // two registers, summed at the point of use one chunk later: an add here would stall the warp on both loads
        if (p.bias != nullptr) bpf0 = __ldg(p.bias + col);
        if (rowbias != nullptr && on_w < p.Nb) bpf1 = __ldg(rowbias + static_cast<long long>(on_w) * p.rowbias_ld + col);
      }
    };

    int ti = 0;
    int seq = 0;  // chunks consumed by this warp-set
    // debug counters of CTA 0, epilogue warp 3 (not the electing warp): total, waiting for an accumulator, slot, barrier
    const bool edbg = p.debug != 0 && blockIdx.x == 0 && warp == 3 && lane == 0;
    // debug counters accumulate straight into g_gemm_dbg (zeroed by the host): no registers held across the chunk loop
    const unsigned e_all = edbg ? static_cast<unsigned>(clock()) : 0u;
    unsigned tc0 = 0u, te = 0u;
    auto dbg_add = [&](int i, unsigned d) { g_gemm_dbg[i] += d; };
    // position of my accumulator row inside a tile: constant over all tiles
    const int wl = row % p.tw;
    const int hl = (row / p.tw) % p.th;
    const int nl = row / (p.tw * p.th);
    float2 lnp0 = make_float2(0.f, 0.f), lnp1 = lnp0, lnp2 = lnp0, lnp3 = lnp0;   // LayerNorm-fold row partials of the next tile
    int lnp_tile = -1;
    TileCoord tc = tile_coord_from_unit(vblock, tstep);
    for (int tile = vblock; tile < total_tiles; tile += vgrid, ++ti, tile_coord_advance(tc, tstep)) {
      const int buf = ti & 1;
      const int n_tile = tc.nt, w0 = (tc.wi * mw + cw) * p.tw, h0 = (tc.hi * mh + ch) * p.th,
                n0 = (tc.ni * mn + cn) * p.tn;
      const int n_base = n_tile * BLOCK_N;
      const int ow = w0 + wl, oh = h0 + hl, on = n0 + nl;
      const int on_w = __shfl_sync(0xffffffffu, on, 0);
      // the chunk this warp-set handles first in its next tile (bias prefetch across the tile boundary)
      TileCoord tcn = tc;
      tile_coord_advance(tcn, tstep);
      const bool next_tile_ok = tile + vgrid < total_tiles;
      const int next_c_first = (ws + ti + 1) & 1;
      const int next_on_w = __shfl_sync(0xffffffffu, (tcn.ni * mn + cn) * p.tn + nl, 0);
      const bool my_valid = (ow < p.Wo) && (oh < p.Ho) && (on < p.Nb);
      const long long my_m = (static_cast<long long>(on) * p.Ho + oh) * p.Wo + ow;

      // LayerNorm fold, reading side: this row's (sum, sum of squares) partials -> rstd and -mean * rstd.  The loads are
      // issued before the accumulator wait and summed in a fixed order (part 0, 1, ...).
      float ln_r = 1.0f;
      if (ln_on) {
        float s_sum = 0.0f, s_sq = 0.0f;
        if (lnp_tile == tile) {   // prefetched while the previous tile was processed (rows with at most four partials)
          s_sum = ((lnp0.x + lnp1.x) + lnp2.x) + lnp3.x;
          s_sq = ((lnp0.y + lnp1.y) + lnp2.y) + lnp3.y;
        } else if (my_valid) {
          const float2* lp = p.ln_in + my_m;
          float2 t0 = __ldg(lp), t1 = __ldg(lp + p.ln_stride), t2 = make_float2(0.f, 0.f), t3 = make_float2(0.f, 0.f);
          if (p.ln_parts > 2) { t2 = __ldg(lp + 2 * p.ln_stride); t3 = __ldg(lp + 3 * p.ln_stride); }
          s_sum = ((t0.x + t1.x) + t2.x) + t3.x;
          s_sq = ((t0.y + t1.y) + t2.y) + t3.y;
          for (int part = 4; part < p.ln_parts; ++part) {
            const float2 t = __ldg(lp + static_cast<long long>(part) * p.ln_stride);
            s_sum += t.x; s_sq += t.y;
          }
        }
        const float mean = s_sum * p.ln_inv_c;
        const float var = fmaxf(fmaf(s_sq, p.ln_inv_c, -mean * mean), 0.0f);
        ln_r = rsqrtf(var + p.ln_eps);
        // the next tile's partials: loads in flight during this whole tile, summed (in the same order) at its start
        if (p.ln_parts <= 4 && next_tile_ok) {
          const int ow2 = (tcn.wi * mw + cw) * p.tw + wl, oh2 = (tcn.hi * mh + ch) * p.th + hl, on2 = (tcn.ni * mn + cn) * p.tn + nl;
          lnp0 = lnp1 = lnp2 = lnp3 = make_float2(0.f, 0.f);
          if (ow2 < p.Wo && oh2 < p.Ho && on2 < p.Nb) {
            const float2* lp = p.ln_in + (static_cast<long long>(on2) * p.Ho + oh2) * p.Wo + ow2;
            lnp0 = __ldg(lp); lnp1 = __ldg(lp + p.ln_stride);
            if (p.ln_parts > 2) { lnp2 = __ldg(lp + 2 * p.ln_stride); lnp3 = __ldg(lp + 3 * p.ln_stride); }
          }
          lnp_tile = tile + vgrid;
        }
      }
      // writing side: this thread's row statistics over the chunks it handles in this tile
      f32x2 ln_s2 = pk2(0.0f, 0.0f), ln_q2 = pk2(0.0f, 0.0f);
      if (edbg) te = static_cast<unsigned>(clock());
      mbar_wait(tmem_full_bar(buf), (ti >> 1) & 1);
      if (edbg) dbg_add(6, static_cast<unsigned>(clock()) - te);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
      const int c_first = (ws + ti) & 1;

      if (tile_is_vt(n_tile)) {
        // V^T scatter: out_vt[b][c][token]; lanes are consecutive tokens -> 64-B contiguous per column
        int last_c = -1;
        for (int c = c_first; c < nchunks; c += 2) last_c = c;
        if (last_c < 0 && lane == 0) tmem_empty_arrive(buf);
#pragma unroll 1
        for (int c = c_first; c <= last_c; c += 2) {
          uint32_t v[32];
          tmem_ld_x32(taddr + c * 32, v);
          const bool vt_bias = p.bias != nullptr && n_base < p.bias_cols;
          if (vt_bias) {   // v projection with a bias (VAE attention; none in the U-Net): through the per-warp scratch, as below
            if (bpf_key != tile * 8 + c) bias_fetch(tile * 8 + c, n_base + c * 32, on_w);
            const float cb = bpf0 + bpf1;
            if (c + 2 <= last_c) bias_fetch(tile * 8 + c + 2, n_base + (c + 2) * 32, on_w);
            else if (next_tile_ok) bias_fetch((tile + vgrid) * 8 + next_c_first, tcn.nt * BLOCK_N + next_c_first * 32, next_on_w);
            __syncwarp();
            wscr[lane] = cb;
            __syncwarp();
          }
          tmem_ld_wait();
          if (c == last_c) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tmem_empty_arrive(buf);
          }
          if (my_valid) {
            // sample / token of my row: (n, h*W + w), or -- flat [1,1,M,C] activations -- (m / vt_tokens, m % vt_tokens)
            const long long ntok = p.vt_tokens ? p.vt_tokens : static_cast<long long>(p.Ho) * p.Wo;
            const long long tokens = (ntok + 7) & ~7ll;   // row pitch: vt_pitch(), internal.h
            const long long smp = p.vt_tokens ? my_m / ntok : on;
            const long long tok = p.vt_tokens ? my_m - smp * ntok : static_cast<long long>(oh) * p.Wo + ow;
            const int vC = p.n_total - p.qk_cols;
            bf16* dst = p.out_vt + (smp * vC + (n_base + c * 32 - p.qk_cols)) * tokens + tok;
            // rstd * acc (LayerNorm fold; rstd = 1 without it) + bias
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 a = vt_bias ? *reinterpret_cast<const float4*>(wscr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[j + 0] = __float_as_uint(fmaf(ln_r, __uint_as_float(v[j + 0]), a.x));
              v[j + 1] = __float_as_uint(fmaf(ln_r, __uint_as_float(v[j + 1]), a.y));
              v[j + 2] = __float_as_uint(fmaf(ln_r, __uint_as_float(v[j + 2]), a.z));
              v[j + 3] = __float_as_uint(fmaf(ln_r, __uint_as_float(v[j + 3]), a.w));
            }
#pragma unroll
            for (int j = 0; j < 32; ++j)
              reinterpret_cast<unsigned short*>(dst)[j * tokens] = to_op16(__uint_as_float(v[j]), p.out16_f16);
          }
        }
        continue;
      }

      int last_c = -1;
      for (int c = c_first; chunk_valid(n_tile, c); c += 2) last_c = c;
      if (last_c < 0 && lane == 0) tmem_empty_arrive(buf);
#pragma unroll 1
      for (int c = c_first; c <= last_c; c += 2, ++seq) {
        if (edbg) { tc0 = static_cast<unsigned>(clock()); dbg_add(11, 1u); }
        const int slot = ws * 2 + (seq & 1);
        const uint32_t slot_addr = slot_base + slot * SLOT_BYTES;
        uint8_t* slot_gen = slot_gen_base + slot * SLOT_BYTES;
        float o[32];
        if constexpr (geglu) {
          constexpr int HALF = BLOCK_N / 2;
          uint32_t va[32], vg[32];
          tmem_ld_x32(taddr + c * 32, va);
          tmem_ld_x32(taddr + HALF + c * 32, vg);
          // this chunk's 32 value + 32 gate biases: one coalesced load per lane, broadcast through a per-warp scratch in
          // the unused upper half of the slot (bf16-only epilogues stage 8 KB of the 16 KB).  Sixteen float4 __ldg's
          // inside the GELU math stalled every group of four outputs on a global-load round trip.
          float* bscr = DEDICATED_SCR ? wscr : reinterpret_cast<float*>(slot_gen + 8192 + (ew & 3) * 512);
          {
            if (bpf_key != tile * 8 + c) bias_fetch(tile * 8 + c, n_base + c * 32, on_w);   // warp-uniform; normally prefetched
            const float b_val = bpf0, b_gate = bpf1;
            if (c + 2 <= last_c) bias_fetch(tile * 8 + c + 2, n_base + (c + 2) * 32, on_w);
            else if (next_tile_ok) bias_fetch((tile + vgrid) * 8 + next_c_first, tcn.nt * BLOCK_N + next_c_first * 32, next_on_w);
            __syncwarp();   // previous chunk's reads of the scratch are done
            bscr[lane] = b_val;
            bscr[32 + lane] = b_gate;
            __syncwarp();
          }
          tmem_ld_wait();
          if (c == last_c) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tmem_empty_arrive(buf);
          }
          // eight output pairs per block, evaluated stage by stage: every dependent step of the GELU polynomial has
          // seven independent neighbours to hide its latency behind (two warps per scheduler cannot)
#pragma unroll
          const f32x2 r2 = pk2(ln_r, ln_r);
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 16) {
            f32x2 av[8], gv[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 ba = *reinterpret_cast<const float4*>(bscr + j0 + 4 * i);
              const float4 bg = *reinterpret_cast<const float4*>(bscr + 32 + j0 + 4 * i);
              const int j = j0 + 4 * i;
              // LayerNorm fold: rstd * acc + bias' (rstd = 1 without it: the fma is then exactly the add)
              av[2 * i] = fma2(r2, pk2(__uint_as_float(va[j + 0]), __uint_as_float(va[j + 1])), pk2(ba.x, ba.y));
              av[2 * i + 1] = fma2(r2, pk2(__uint_as_float(va[j + 2]), __uint_as_float(va[j + 3])), pk2(ba.z, ba.w));
              gv[2 * i] = fma2(r2, pk2(__uint_as_float(vg[j + 0]), __uint_as_float(vg[j + 1])), pk2(bg.x, bg.y));
              gv[2 * i + 1] = fma2(r2, pk2(__uint_as_float(vg[j + 2]), __uint_as_float(vg[j + 3])), pk2(bg.z, bg.w));
            }
            geglu_block<8>(av, gv);
#pragma unroll
            for (int i = 0; i < 8; ++i) upk2(av[i], o[j0 + 2 * i], o[j0 + 2 * i + 1]);
          }
        } else {
          uint32_t v[32];
          tmem_ld_x32(taddr + c * 32, v);
          const int col = n_base + c * 32;
          const bool tile_bias = bias_uniform && bias_any && n_base < p.bias_cols;   // (QKV with a folded LayerNorm: only Q has a bias)
          if (tile_bias) {
            if (bpf_key != tile * 8 + c) bias_fetch(tile * 8 + c, col, on_w);   // warp-uniform; normally prefetched a chunk ago
            const float cb = bpf0 + bpf1;
            if (c + 2 <= last_c) bias_fetch(tile * 8 + c + 2, n_base + (c + 2) * 32, on_w);
            else if (next_tile_ok) bias_fetch((tile + vgrid) * 8 + next_c_first, tcn.nt * BLOCK_N + next_c_first * 32, next_on_w);
            __syncwarp();   // previous chunk's reads of the scratch are done
            wscr[lane] = cb;
            __syncwarp();
          }
          tmem_ld_wait();
          if (c == last_c) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tmem_empty_arrive(buf);
          }
          if (bias_uniform) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              if (tile_bias) a = *reinterpret_cast<const float4*>(wscr + j);
              // rstd * acc + bias (LayerNorm fold); rstd = 1 without it and the fma is exactly the add
              o[j + 0] = fmaf(ln_r, __uint_as_float(v[j + 0]), a.x);
              o[j + 1] = fmaf(ln_r, __uint_as_float(v[j + 1]), a.y);
              o[j + 2] = fmaf(ln_r, __uint_as_float(v[j + 2]), a.z);
              o[j + 3] = fmaf(ln_r, __uint_as_float(v[j + 3]), a.w);
            }
          } else {   // tiles whose warps straddle samples (tiny feature maps): per-thread loads
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias != nullptr && col + j < p.n_total) a = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
              if (rowbias != nullptr && my_valid && col + j < p.n_total) {
                const float4 rb = __ldg(reinterpret_cast<const float4*>(rowbias + static_cast<long long>(on) * p.rowbias_ld + col + j));
                a.x += rb.x; a.y += rb.y; a.z += rb.z; a.w += rb.w;
              }
              o[j + 0] = __uint_as_float(v[j + 0]) + a.x;
              o[j + 1] = __uint_as_float(v[j + 1]) + a.y;
              o[j + 2] = __uint_as_float(v[j + 2]) + a.z;
              o[j + 3] = __uint_as_float(v[j + 3]) + a.w;
            }
          }
          if (e_act == 1) {   // quick_gelu (CLIP MLP): x * sigmoid(1.702 x)
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __fdividef(o[j], 1.0f + __expf(-1.702f * o[j]));
          }
        }
        // the slot is ours once the residual prefetch (or the plain arrive that stands in for it) has landed
        if (edbg) { te = static_cast<unsigned>(clock()); dbg_add(8, te - tc0); }
        // Without a residual nothing is prefetched into the slot: it is free once the store issued from it two
        // chunks ago has read it, which the storing warp checks before it joins the previous chunk's barrier (below).
        if (e_has_res) mbar_wait(res_full_bar(slot), (seq >> 1) & 1);
        if (edbg) { tc0 = static_cast<unsigned>(clock()); dbg_add(12, tc0 - te); }
        uint8_t* my_row128 = slot_gen + row * 128;
        if (e_has_res && p.res16) {
          // 16-bit residual stream: the box landed in the compact [128 rows x 64 B] layout (64B swizzle) the 16-bit output
          // uses -- a thread reads its own row here and overwrites exactly those bytes below: no barrier in between
          const uint8_t* rrow = slot_gen + row * 64;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint4 rr = *reinterpret_cast<const uint4*>(rrow + ((u ^ ((row >> 1) & 3)) << 4));
            const float2 r0 = unpack_op2(rr.x, p.out16_f16), r1 = unpack_op2(rr.y, p.out16_f16),
                         r2 = unpack_op2(rr.z, p.out16_f16), r3 = unpack_op2(rr.w, p.out16_f16);
            o[8 * u + 0] += r0.x; o[8 * u + 1] += r0.y; o[8 * u + 2] += r1.x; o[8 * u + 3] += r1.y;
            o[8 * u + 4] += r2.x; o[8 * u + 5] += r2.y; o[8 * u + 6] += r3.x; o[8 * u + 7] += r3.y;
          }
        } else if (e_has_res) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 rr = *reinterpret_cast<const float4*>(my_row128 + ((u ^ (row & 7)) << 4));
            o[4 * u + 0] += rr.x; o[4 * u + 1] += rr.y; o[4 * u + 2] += rr.z; o[4 * u + 3] += rr.w;
          }
        }
        if (e_has_o32) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<float4*>(my_row128 + ((u ^ (row & 7)) << 4)) =
                make_float4(o[4 * u + 0], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
          if (e_stats_out != nullptr) {
            // GroupNorm statistics of the tensor being written: lane = column, walk my warp's 32 rows of the slot
            // (conflict-free: one 128-B row per step).  Fixed order, no atomics -> bit-reproducible.
            __syncwarp();
            const unsigned vmask = __ballot_sync(0xffffffffu, my_valid);
            const long long m_first = __shfl_sync(0xffffffffu, my_m, 0);
            const long long on_first = __shfl_sync(0xffffffffu, on, 0);
            const int col = n_base + c * 32 + lane;
            if (vmask != 0u && col < p.n_total) {
              // shared-space loads and branch-free accumulation into four independent chains (the first version --
              // generic loads plus a divergent `if` per row -- cost an S2UR, a BSSY/BSYNC pair and a branch per
              // element and doubled the epilogue time of every GEMM that carried statistics)
              float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
              const uint32_t wrow = slot_addr + (q * 32) * 128 + ((lane & 3) << 2);
              const int l2 = lane >> 2;
#pragma unroll
              for (int r = 0; r < 32; r += 4) {
                float x0 = lds_f32(wrow + (r + 0) * 128 + ((l2 ^ ((r + 0) & 7)) << 4));
                float x1 = lds_f32(wrow + (r + 1) * 128 + ((l2 ^ ((r + 1) & 7)) << 4));
                float x2 = lds_f32(wrow + (r + 2) * 128 + ((l2 ^ ((r + 2) & 7)) << 4));
                float x3 = lds_f32(wrow + (r + 3) * 128 + ((l2 ^ ((r + 3) & 7)) << 4));
                if (vmask != 0xffffffffu) {   // warp-uniform: only ragged tiles pay for the selects
                  x0 = ((vmask >> (r + 0)) & 1u) ? x0 : 0.0f; x1 = ((vmask >> (r + 1)) & 1u) ? x1 : 0.0f;
                  x2 = ((vmask >> (r + 2)) & 1u) ? x2 : 0.0f; x3 = ((vmask >> (r + 3)) & 1u) ? x3 : 0.0f;
                }
                s0 += x0; q0 = fmaf(x0, x0, q0); s1 += x1; q1 = fmaf(x1, x1, q1);
                s2 += x2; q2 = fmaf(x2, x2, q2); s3 += x3; q3 = fmaf(x3, x3, q3);
              }
              *reinterpret_cast<float2*>(e_stats_out + (((m_first >> 5) + on_first * p.stats_sample_extra + p.stats_block_off) * p.n_total + col) * 2) =
                  make_float2((s0 + s1) + (s2 + s3), (q0 + q1) + (q2 + q3));
            }
          }
          if (e_has_o16 && my_valid) {  // rare side copy (feeds a stride-2 conv): direct 64-B row store
            const int col = n_base + c * 32;
            bf16* dst = out_bf16 + my_m * p.ld_out + col;
            if (((p.n_total | p.ld_out) & 7) != 0) {  // narrow / unaligned rows: scalar stores
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < p.n_total) reinterpret_cast<unsigned short*>(dst)[j] = to_op16(o[j], p.out16_f16);
            } else
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (col + 8 * u < p.n_total) {
                uint4 pk;
                pk.x = pack_op2(o[8 * u + 0], o[8 * u + 1], p.out16_f16);
                pk.y = pack_op2(o[8 * u + 2], o[8 * u + 3], p.out16_f16);
                pk.z = pack_op2(o[8 * u + 4], o[8 * u + 5], p.out16_f16);
                pk.w = pack_op2(o[8 * u + 6], o[8 * u + 7], p.out16_f16);
                *reinterpret_cast<uint4*>(dst + 8 * u) = pk;
              }
            }
          }
        } else {
          // 16-bit-only output: compact [128 rows x 64 B] layout (64B swizzle) overlaps other rows' fp32 residual
          if (e_has_res && !p.res16) named_bar_sync(bar_id, 128);
          uint8_t* my_row64 = slot_gen + row * 64;
          if (p.out16_f16) {   // warp-uniform: one conversion per pair on either path
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 pk;
              pk.x = pack_f16x2(o[8 * u + 0], o[8 * u + 1]);
              pk.y = pack_f16x2(o[8 * u + 2], o[8 * u + 3]);
              pk.z = pack_f16x2(o[8 * u + 4], o[8 * u + 5]);
              pk.w = pack_f16x2(o[8 * u + 6], o[8 * u + 7]);
              *reinterpret_cast<uint4*>(my_row64 + ((u ^ ((row >> 1) & 3)) << 4)) = pk;
              if (e_ln_stats_out != nullptr) ln_accumulate(pk, ln_s2, ln_q2, 1);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 pk;
              pk.x = pack_bf16x2(o[8 * u + 0], o[8 * u + 1]);
              pk.y = pack_bf16x2(o[8 * u + 2], o[8 * u + 3]);
              pk.z = pack_bf16x2(o[8 * u + 4], o[8 * u + 5]);
              pk.w = pack_bf16x2(o[8 * u + 6], o[8 * u + 7]);
              *reinterpret_cast<uint4*>(my_row64 + ((u ^ ((row >> 1) & 3)) << 4)) = pk;
              if (e_ln_stats_out != nullptr) ln_accumulate(pk, ln_s2, ln_q2, 0);
            }
          }
          if (e_stats_out != nullptr) {
            // GroupNorm statistics of the 16-bit tensor being written, taken from the ROUNDED values in the slot (they are
            // the statistics of what the next GroupNorm will read): lane = column, walk my warp's 32 rows.  Fixed order.
            __syncwarp();
            const unsigned vmask = __ballot_sync(0xffffffffu, my_valid);
            const long long m_first = __shfl_sync(0xffffffffu, my_m, 0);
            const long long on_first = __shfl_sync(0xffffffffu, on, 0);
            const int col = n_base + c * 32 + lane;
            if (vmask != 0u && col < p.n_total) {
              float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
              const uint32_t wbase = slot_addr + (q * 32) * 64 + ((lane & 7) << 1);
              const int unit = lane >> 3;
#pragma unroll
              for (int r = 0; r < 32; r += 2) {
                const int ra = q * 32 + r, rb2 = ra + 1;
                unsigned short ha, hb;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(ha) : "r"(wbase + r * 64 + ((unit ^ ((ra >> 1) & 3)) << 4)));
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hb) : "r"(wbase + (r + 1) * 64 + ((unit ^ ((rb2 >> 1) & 3)) << 4)));
                float xa = op16_to_float(ha, p.out16_f16), xb = op16_to_float(hb, p.out16_f16);
                if (vmask != 0xffffffffu) {
                  xa = ((vmask >> r) & 1u) ? xa : 0.0f;
                  xb = ((vmask >> (r + 1)) & 1u) ? xb : 0.0f;
                }
                s0 += xa; q0 = fmaf(xa, xa, q0);
                s1 += xb; q1 = fmaf(xb, xb, q1);
              }
              *reinterpret_cast<float2*>(e_stats_out + (((m_first >> 5) + on_first * p.stats_sample_extra + p.stats_block_off) * p.n_total + col) * 2) = make_float2(s0 + s1, q0 + q1);
            }
          }
        }
        if (edbg) { te = static_cast<unsigned>(clock()); dbg_add(9, te - tc0); }
        fence_async_smem();
        const int ocol = n_tile * tile_out_cols + c * chunk_cols;
        if (edbg) { tc0 = static_cast<unsigned>(clock()); dbg_add(10, tc0 - te); te = tc0; }
        if (e_has_res) {
          named_bar_sync(bar_id, 128);
          if (edbg) dbg_add(13, static_cast<unsigned>(clock()) - te);
          if (elected) {
            if (e_has_o32) tma_store_4d(&tmO32, slot_addr, ocol, w0, h0, n0 + tc.s * nb_pad);
            else tma_store_4d(&tmO16, slot_addr, ocol, w0, h0, n0);
            tma_store_commit();
            tma_store_wait_read0();  // slot may be overwritten again
            pf_issue_next();         // residual prefetch for the chunk two ahead, into this slot
          }
        } else {
          // The storing warp runs warp-uniform code (elected lane issues).  Before the barrier it makes sure the store
          // of the PREVIOUS chunk (issued a whole chunk ago, from the other slot) has read its slot: passing the
          // barrier then tells all four warps that the other slot may be overwritten -- no mbarrier round trip, no
          // stall on the store just issued, no single-lane code between two barriers.
          if (store_warp) tma_store_wait_read0();
          named_bar_sync(bar_id, 128);
          if (edbg) dbg_add(13, static_cast<unsigned>(clock()) - te);
          if (store_warp) {
            if (e_has_o32) tma_store_4d_commit_elect(&tmO32, slot_addr, ocol, w0, h0, n0 + tc.s * nb_pad);
            else tma_store_4d_commit_elect(&tmO16, slot_addr, ocol, w0, h0, n0);
          }
        }
      }
      if (e_ln_stats_out != nullptr && my_valid) {   // LayerNorm fold, writing side: partial (n tile, chunk parity) of my row
        float s0, s1, q0, q1;
        upk2(ln_s2, s0, s1);
        upk2(ln_q2, q0, q1);
        e_ln_stats_out[static_cast<long long>(n_tile * 2 + c_first) * p.ln_stats_stride + my_m] = make_float2(s0 + s1, q0 + q1);
      }
    }
    if (edbg) {
      g_gemm_dbg[5] = static_cast<unsigned>(clock()) - e_all;
      g_gemm_dbg[7] = (g_gemm_dbg[12] << 32) | (g_gemm_dbg[13] & 0xffffffffll);
    }
    if (store_warp) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();  // no CTA leaves (or frees TMEM) while the pair's MMAs / signals may still touch it
  if (warp == 1) {
    if (CL == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int STAGES, int CL>
constexpr size_t smem_bytes_for() {
  return 1024 + STAGES * (A_STAGE_BYTES + (BLOCK_N / CL) * BLOCK_K * 2) + EPI_STAGING_BYTES + 8 * (2 * STAGES + 6 + NUM_SLOTS) +
         (BLOCK_N != 256 ? BIAS_SCR_BYTES : 0);
}

template <int BLOCK_N, int STAGES, int CL>
int launch_tc(const GemmPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  constexpr size_t smem = smem_bytes_for<BLOCK_N, STAGES, CL>();
  if (!attr_set) {
    PBE_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, STAGES, CL>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  static_assert(smem <= 227 * 1024, "shared memory budget");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = plan.grid;
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (CL > 1) {
    // persistent clusters are statically scheduled: never launch more than can be co-resident (GPCs with an odd number
    // of free SMs cannot host a last 2-CTA cluster)
    static int max_clusters = 0;
    if (max_clusters == 0) {
      PBE_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, conv_gemm_kernel<BLOCK_N, STAGES, CL>, &cfg));
      if (getenv("PBE_GEMM_DEBUG")) fprintf(stderr, "[pbe] conv_gemm<%d,%d> max active %d-CTA clusters: %d\n", BLOCK_N, STAGES, CL, max_clusters);
      PBE_REQUIRE(max_clusters > 0, "no co-resident cluster fits");
    }
    cfg.gridDim.x = std::min<unsigned>(cfg.gridDim.x, static_cast<unsigned>(max_clusters * CL));
  }
  PBE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BLOCK_N, STAGES, CL>, plan.tmA, plan.tmB, plan.tmR,
                                    plan.tmO32, plan.tmO16, plan.p));
  return 0;
}
template <int BLOCK_N, int STAGES, int STAGES_PAIR>
int launch_t(const GemmPlan& plan, cudaStream_t stream) {
  return plan.cluster == 2 ? launch_tc<BLOCK_N, STAGES_PAIR, 2>(plan, stream) : launch_tc<BLOCK_N, STAGES, 1>(plan, stream);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// 128-row tile = tn samples x th rows x tw pixels (powers of two): the shape that covers the output with the fewest tiles;
// ties go to the widest, then the tallest tile (longest contiguous runs for TMA).  (The first version minimised the
// padding of W and H one after the other and ignored Nb: a flat [1, 1, 257, C] activation got 1-pixel x 128-sample tiles,
// 257 of them with one valid row each.)
void pick_tile(int Wo, int Ho, int Nb, int* tw, int* th, int* tn) {
  long best_cost = -1;
  for (int w = 128; w >= 1; w >>= 1)
    for (int h = 128 / w; h >= 1; h >>= 1) {
      const int n = 128 / (w * h);
      const long cost = static_cast<long>((Wo + w - 1) / w) * ((Ho + h - 1) / h) * ((Nb + n - 1) / n);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        *tw = w; *th = h; *tn = n;
      }
    }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(SplitKReduce r) {
  griddep_enter();
  const long long n4 = r.M * (r.N / 4);
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  const long long m = idx / (r.N / 4);
  const int col = static_cast<int>(idx - m * (r.N / 4)) * 4;
  float4 acc = *reinterpret_cast<const float4*>(r.ws + m * r.N + col);
  for (int s = 1; s < r.S; ++s) {  // fixed order -> bit-reproducible
    const float4 v = *reinterpret_cast<const float4*>(r.ws + s * r.slice_stride + m * r.N + col);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (r.bias) {
    const float4 b = *reinterpret_cast<const float4*>(r.bias + col);
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (r.rowbias) {
    const long long smp = m / r.HW;
    const float4 b = *reinterpret_cast<const float4*>(r.rowbias + smp * r.rowbias_ld + col);
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  const long long o = m * r.ld_out + col;
  if (r.residual) {
    const float4 b = *reinterpret_cast<const float4*>(r.residual + o);
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (r.residual16) {
    const uint2 u = *reinterpret_cast<const uint2*>(r.residual16 + o);
    const float2 b0 = unpack_op2(u.x, r.out16_f16), b1 = unpack_op2(u.y, r.out16_f16);
    acc.x += b0.x; acc.y += b0.y; acc.z += b1.x; acc.w += b1.y;
  }
  if (r.out_f32) *reinterpret_cast<float4*>(r.out_f32 + o) = acc;
  if (r.out_bf16) {
    uint2 pk;
    pk.x = pack_op2(acc.x, acc.y, r.out16_f16);
    pk.y = pack_op2(acc.z, acc.w, r.out16_f16);
    *reinterpret_cast<uint2*>(r.out_bf16 + o) = pk;
  }
}

int auto_block_n(const ConvGemmDesc& d) {
  if (d.block_n) return d.block_n;
  if (d.mode == EPI_GEGLU) return 256;
  if (d.Cout % 160 == 0) return 160;
  if (d.Cout % 128 == 0) return 128;
  if (d.Cout <= 32) return 32;
  if (d.Cout <= 64) return 64;
  return 128;
}

}  // namespace

int gemm_read_debug_counters(long long* out8) {
  PBE_CHECK_CUDA(cudaMemcpyFromSymbol(out8, g_gemm_dbg, sizeof(long long) * 8));
  return 0;
}

int gemm_read_debug_counters16(long long* out16) {
  PBE_CHECK_CUDA(cudaMemcpyFromSymbol(out16, g_gemm_dbg, sizeof(long long) * 16));
  return 0;
}

int gemm_reset_debug_counters() {
  const long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  PBE_CHECK_CUDA(cudaMemcpyToSymbol(g_gemm_dbg, z, sizeof(z)));
  return 0;
}

int gemm_ln_parts(const ConvGemmDesc& d) {
  const int bn = auto_block_n(d);
  if (d.mode != EPI_STD || d.out_f32 != nullptr || d.epi_act != 0 || bn == 256 || bn < 64 || d.Cout % bn != 0 || d.up_phase) return 0;
  return 2 * (d.Cout / bn);
}

int gemm_split_k(const ConvGemmDesc& d) {
  if (d.mode != EPI_STD || d.Cout % 4 != 0 || d.epi_act != 0) return 1;   // the reduce kernel applies no activation
  if (d.ln_stats_out != nullptr) return 1;   // LayerNorm-fold row statistics come out of the GEMM's own epilogue
  int tw, th, tn;
  const int Wo = d.W / d.stride, Ho = d.H / d.stride;
  const int nb = d.split_batch ? d.split_batch : d.Nb;
  pick_tile(Wo, Ho, nb, &tw, &th, &tn);
  const int bn = auto_block_n(d);
  const int tiles = ((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((nb + tn - 1) / tn) * ((d.Cout + bn - 1) / bn);
  const int k_iters = d.ksize * d.ksize * (d.C / 64);
  const int sms = num_sms();
  if (tiles * 2 > sms) return 1;
  int S = sms / tiles;
  S = std::min(S, std::max(1, k_iters / 8));
  S = std::min(S, 16);
  return std::max(S, 1);
}

bool gemm_can_fuse_stats(const ConvGemmDesc& d) {
  if (d.mode != EPI_STD || gemm_split_k(d) > 1) return false;
  int tw, th, tn;
  const int Wo = d.W / d.stride, Ho = d.H / d.stride;
  pick_tile(Wo, Ho, d.Nb, &tw, &th, &tn);
  if (Wo % tw != 0 || Ho % th != 0) return false;          // no partial tiles inside a sample
  if (!(tw % 32 == 0 || (tw == Wo && 32 % tw == 0))) return false;  // a warp's 32 rows are 32 consecutive pixels
  if ((tw * th) % 32 != 0 || (Ho * Wo) % 32 != 0) return false;     // ... of one sample
  return true;
}

size_t gemm_splitk_ws_bytes(const ConvGemmDesc& d) {
  const int S = gemm_split_k(d);
  if (S <= 1) return 0;
  int tw, th, tn;
  const int Wo = d.W / d.stride, Ho = d.H / d.stride;
  pick_tile(Wo, Ho, d.Nb, &tw, &th, &tn);
  const size_t nb_pad = static_cast<size_t>((d.Nb + tn - 1) / tn) * tn;
  return static_cast<size_t>(S) * nb_pad * Ho * Wo * d.Cout * sizeof(float);
}

int build_gemm_plan(const ConvGemmDesc& d, GemmPlan* plan) {
  PBE_REQUIRE(d.C % 64 == 0, "activation channels must be a multiple of 64");
  PBE_REQUIRE(d.ksize == 1 || d.ksize == 3 || (d.ksize == 2 && d.up_phase >= 1 && d.up_phase <= 4), "kernel size 1 or 3 (2: sub-pixel phase)");
  PBE_REQUIRE(d.up_phase == 0 || (d.ksize == 2 && d.stride == 1 && d.mode == EPI_STD && d.residual == nullptr && d.residual16 == nullptr &&
                                  d.splitk_ws == nullptr && d.act_ld == 0),
              "sub-pixel phase conv: ksize 2, stride 1, plain epilogue, no residual, no split-K");
  PBE_REQUIRE(d.Cout % 4 == 0 && d.ld_out % 4 == 0 && d.rowbias_ld % 4 == 0, "output columns / leading dims % 4");
  PBE_REQUIRE(d.stride == 1 || d.stride == 2, "stride 1 or 2");
  PBE_REQUIRE(d.stride == 1 || (d.H % 2 == 0 && d.W % 2 == 0 && d.ksize == 3), "stride-2 conv needs even H, W, k=3");
  ConvGemmParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.Wo = d.W / d.stride;
  p.Ho = d.H / d.stride;
  p.Nb = d.Nb;
  pick_tile(p.Wo, p.Ho, p.Nb, &p.tw, &p.th, &p.tn);
  p.tiles_w = (p.Wo + p.tw - 1) / p.tw;
  p.tiles_h = (p.Ho + p.th - 1) / p.th;
  p.tiles_n = (p.Nb + p.tn - 1) / p.tn;
  p.num_taps = d.ksize * d.ksize;
  p.k_chunks = d.C / 64;
  p.n_total = d.Cout;
  for (int kh = 0; kh < d.ksize; ++kh)
    for (int kw = 0; kw < d.ksize; ++kw) {
      const int t = kh * d.ksize + kw;
      if (d.ksize == 1) {
        p.tap_dw[t] = p.tap_dh[t] = p.tap_ph[t] = 0;
        p.tap_coff[t] = 0;
      } else if (d.ksize == 2) {
        // phase (a, b) of upsample + 3x3: output row 2i+a reads low-resolution rows {i-1, i} (a = 0) or {i, i+1} (a = 1)
        const int a = (d.up_phase - 1) >> 1, b = (d.up_phase - 1) & 1;
        p.tap_dh[t] = static_cast<int8_t>(a == 0 ? kh - 1 : kh);
        p.tap_dw[t] = static_cast<int8_t>(b == 0 ? kw - 1 : kw);
        p.tap_ph[t] = 0;
        p.tap_coff[t] = 0;
      } else if (d.stride == 1) {
        p.tap_dw[t] = static_cast<int8_t>(kw - 1);
        p.tap_dh[t] = static_cast<int8_t>(kh - 1);
        p.tap_ph[t] = 0;
        p.tap_coff[t] = 0;
      } else if (d.pad_end) {
        // padding (0,1,0,1): input col = 2*ow + kw: kw=0 -> parity 0 ; kw=1 -> parity 1 ; kw=2 -> parity 0, shift +1
        p.tap_dw[t] = static_cast<int8_t>(kw == 2 ? 1 : 0);
        p.tap_coff[t] = (kw == 1) ? d.C : 0;
        p.tap_dh[t] = static_cast<int8_t>(kh == 2 ? 1 : 0);
        p.tap_ph[t] = static_cast<int8_t>(kh == 1 ? 1 : 0);
      } else {
        // input col = 2*ow + kw - 1: kw=0 -> parity 1, shift -1 ; kw=1 -> parity 0 ; kw=2 -> parity 1, shift 0
        p.tap_dw[t] = static_cast<int8_t>(kw == 0 ? -1 : 0);
        p.tap_coff[t] = (kw == 1) ? 0 : d.C;
        p.tap_dh[t] = static_cast<int8_t>(kh == 0 ? -1 : 0);
        p.tap_ph[t] = static_cast<int8_t>(kh == 1 ? 0 : 1);
      }
    }
  p.mode = d.mode;
  p.bias = d.bias;
  p.rowbias = d.rowbias;
  p.rowbias_ld = d.rowbias_ld ? d.rowbias_ld : d.Cout;
  p.bias_cols = d.bias_cols > 0 ? d.bias_cols : d.Cout;
  p.residual = d.residual16 ? reinterpret_cast<const float*>(d.residual16) : d.residual;
  p.res16 = d.residual16 != nullptr;
  PBE_REQUIRE(!(d.residual16 && (d.residual || d.out_f32)), "a 16-bit residual goes with a 16-bit-only output");
  p.out_f32 = d.out_f32;
  p.out_bf16 = d.out_bf16;
  p.out_vt = d.out_vt;
  p.qk_cols = d.qk_cols;
  p.vt_tokens = d.vt_tokens;
  p.act = d.epi_act;
  {
    const int f16 = operand_f16();
    p.idesc_fmt = (f16 && !d.operands_bf16) ? 0u : ((1u << 7) | (1u << 10));
    p.out16_f16 = (f16 && !d.out16_bf16) ? 1 : 0;
  }
  PBE_REQUIRE(d.epi_act == 0 || d.mode == EPI_STD, "activation epilogue needs mode STD");
  if (d.ln_stats_out != nullptr) {
    PBE_REQUIRE(gemm_ln_parts(d) > 0 && d.out_bf16 != nullptr && d.ln_stats_stride >= static_cast<long long>(d.Nb) * (d.H / d.stride) * (d.W / d.stride),
                "LayerNorm-fold row statistics: plain epilogue, 16-bit-only output, full n tiles");
    p.ln_stats_out = d.ln_stats_out;
    p.ln_stats_stride = d.ln_stats_stride;
  }
  if (d.ln_in != nullptr) {
    PBE_REQUIRE(d.bias != nullptr && d.ln_parts >= 2 && d.ln_parts % 2 == 0 && d.rowbias == nullptr && d.residual == nullptr &&
                    d.residual16 == nullptr && d.out_f32 == nullptr && d.epi_act == 0 && d.ksize == 1 && d.splitk_ws == nullptr,
                "LayerNorm-fold consumer: 1x1 GEMM with a (folded) bias, 16-bit output, no residual / split-K");
    p.ln_in = d.ln_in;
    p.ln_stride = d.ln_stride;
    p.ln_parts = d.ln_parts;
    p.ln_inv_c = 1.0f / static_cast<float>(d.C);
    p.ln_eps = d.ln_eps;
  }
  p.ld_out = d.ld_out ? d.ld_out : (d.mode == EPI_GEGLU ? d.Cout / 2 : d.Cout);

  const int bn = auto_block_n(d);
  p.split_k = (d.splitk_ws != nullptr) ? gemm_split_k(d) : 1;
  PBE_REQUIRE(d.res_ld == 0 || p.split_k == 1, "res_ld excludes split-K (the reduce kernel reads residual and output at one pitch)");
  p.debug = getenv("PBE_GEMM_DEBUG") ? 1 : 0;
  p.stats_out = nullptr;
  p.stats_sample_extra = 0; p.stats_block_off = 0;
  if (d.stats_out != nullptr && d.up_phase) {
    const int blocks_low = p.Ho * p.Wo / 32;
    p.stats_sample_extra = 3 * blocks_low;                       // a sample owns 4 * blocks_low row blocks of the full-res tensor
    p.stats_block_off = (d.up_phase - 1) * blocks_low;
  }
  if (d.stats_out != nullptr) {
    PBE_REQUIRE(gemm_can_fuse_stats(d) && p.split_k == 1 && (d.out_f32 != nullptr || d.out_bf16 != nullptr),
                "fused GroupNorm statistics not available for this GEMM");
    p.stats_out = d.stats_out;
  }
  plan->red = SplitKReduce{};
  plan->red.S = p.split_k;
  plan->red.out16_f16 = p.out16_f16;
  PBE_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 160 || bn == 256, "unsupported BLOCK_N");
  if (d.mode == EPI_GEGLU) PBE_REQUIRE(bn == 256 && d.Cout % 256 == 0, "GEGLU needs BLOCK_N=256 | Cout");
  PBE_REQUIRE((bn == 256) == (d.mode == EPI_GEGLU), "the 256-wide tile is the GEGLU tile (its kernel compiles only that epilogue)");
  if (d.mode == EPI_QKV) PBE_REQUIRE(d.qk_cols % bn == 0 && d.Cout % bn == 0, "QKV split must align with BLOCK_N");
  plan->block_n = bn;
  p.n_tiles = (d.Cout + bn - 1) / bn;
  {
    // pair neighbouring m tiles into 2-CTA clusters that share the weight tile (first tile dimension with an even count)
    static const bool cluster_on = [] { const char* e = getenv("PBE_GEMM_CLUSTER"); return e == nullptr || atoi(e) != 0; }();
    p.pair_dim = -1;
    // pairs pay off when the main loop dominates (long K: 3x3 convs, ff.out); the short-K GEMMs are epilogue-bound and
    // measured ~15 % slower with the two epilogues coupled through one MMA issuer
    const int pair_min_k = [] { const char* e = getenv("PBE_GEMM_PAIR_MIN_K"); return e ? atoi(e) : 12; }();
    if (cluster_on && p.num_taps * p.k_chunks >= pair_min_k * p.split_k) {
      if (p.tiles_w % 2 == 0) p.pair_dim = 0;
      else if (p.tiles_h % 2 == 0) p.pair_dim = 1;
      else if (p.tiles_n % 2 == 0) p.pair_dim = 2;
    }
    plan->cluster = p.pair_dim >= 0 ? 2 : 1;
    const int total = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles * p.split_k;
    int grid = std::min(total, num_sms());
    if (plan->cluster == 2) grid &= ~1;
    plan->grid = dim3(grid, 1, 1);
  }

  // A: 5-D view (c, w, phase, h, n)
  {
    const uint64_t C = d.C, W = d.W, H = d.H, N = d.Nb;
    const uint64_t L = d.act_ld ? d.act_ld : d.C;   // pixel pitch (a column slice of a wider tensor)
    PBE_REQUIRE(L >= C && L % 8 == 0 && (d.act_ld == 0 || d.stride == 1), "act_ld: >= C, multiple of 8, stride-1 only");
    uint64_t dims[5], strides[4];
    if (d.stride == 1) {
      dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
      strides[0] = L * 2; strides[1] = W * L * 2; strides[2] = W * L * 2; strides[3] = H * W * L * 2;
    } else {
      dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
      strides[0] = 2 * C * 2; strides[1] = W * C * 2; strides[2] = 2 * W * C * 2; strides[3] = H * W * C * 2;
    }
    const uint32_t box[5] = {64u, static_cast<uint32_t>(p.tw), 1u, static_cast<uint32_t>(p.th),
                             static_cast<uint32_t>(p.tn)};
    int rc = make_tmap_bf16(&plan->tmA, d.act, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  // Epilogue maps: residual (fp32 load), fp32 store, bf16 store; geometry = output pixels, box = 32 columns x tile
  {
    const int out_cols = (d.mode == EPI_GEGLU) ? d.Cout / 2 : (d.mode == EPI_QKV ? d.qk_cols : d.Cout);
    const uint64_t ld = static_cast<uint64_t>(p.ld_out);
    const uint64_t Wo = p.Wo, Ho = p.Ho, Nb = p.Nb;
    const uint64_t dims[4] = {static_cast<uint64_t>(out_cols), Wo, Ho, Nb};
    const uint32_t box[4] = {32u, static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), static_cast<uint32_t>(p.tn)};
    p.has_res = d.residual != nullptr || d.residual16 != nullptr;
    p.has_o32 = d.out_f32 != nullptr;
    p.has_o16 = d.out_bf16 != nullptr;
    PBE_REQUIRE(p.has_o32 || p.has_o16, "GEMM needs an output");
    if (p.split_k > 1) {
      // main kernel: raw fp32 partial tiles into the workspace, slices stacked along the batch coordinate
      const uint64_t nb_pad = static_cast<uint64_t>(p.tiles_n) * p.tn;
      SplitKReduce& r = plan->red;
      r.ws = d.splitk_ws;
      r.slice_stride = static_cast<long long>(nb_pad) * p.Ho * p.Wo * d.Cout;
      r.M = static_cast<long long>(p.Nb) * p.Ho * p.Wo;
      r.N = d.Cout;
      r.HW = p.Ho * p.Wo;
      r.bias = d.bias; r.rowbias = d.rowbias; r.rowbias_ld = p.rowbias_ld; r.residual = d.residual;
      r.residual16 = d.residual16;
      r.out_f32 = d.out_f32; r.out_bf16 = d.out_bf16; r.ld_out = p.ld_out;
      p.bias = nullptr; p.rowbias = nullptr; p.residual = nullptr; p.res16 = 0; p.out_bf16 = nullptr;
      p.out_f32 = d.splitk_ws;
      p.has_res = 0; p.has_o32 = 1; p.has_o16 = 0;
      const uint64_t wdims[4] = {static_cast<uint64_t>(d.Cout), Wo, Ho, nb_pad * p.split_k};
      const uint64_t wld = static_cast<uint64_t>(d.Cout);
      const uint64_t ws32[3] = {wld * 4, Wo * wld * 4, Ho * Wo * wld * 4};
      int rcw = make_tmap(&plan->tmO32, d.splitk_ws, true, 4, wdims, ws32, box, 128);
      if (rcw) return rcw;
      plan->tmR = plan->tmO32;
      plan->tmO16 = plan->tmO32;
    } else {
    PBE_REQUIRE(!(d.mode != EPI_STD && (p.has_res || p.has_o32)), "GEGLU / QKV epilogues write bf16 only");
    // sub-pixel phase (a, b): the same pixel grid, written to every second pixel / row of the full-resolution tensor
    const uint64_t us = d.up_phase ? 2 : 1;
    const uint64_t s32[3] = {us * ld * 4, us * (us * Wo) * ld * 4, (us * Ho) * (us * Wo) * ld * 4};
    const uint64_t s16[3] = {us * ld * 2, us * (us * Wo) * ld * 2, (us * Ho) * (us * Wo) * ld * 2};
    const size_t ph_off = d.up_phase ? (static_cast<size_t>((d.up_phase - 1) >> 1) * 2 * Wo + ((d.up_phase - 1) & 1)) * ld : 0;
    PBE_REQUIRE(!(d.up_phase && p.has_o16 && p.has_o32), "sub-pixel phase conv writes one output tensor");
    int rc = 0;
    const uint64_t rl = d.res_ld ? static_cast<uint64_t>(d.res_ld) : ld;   // the residual's own row pitch
    PBE_REQUIRE(d.res_ld == 0 || (d.res_ld % 8 == 0 && d.res_ld >= out_cols && !d.up_phase), "res_ld: >= output columns, multiple of 8");
    const uint64_t r32[3] = {rl * 4, Wo * rl * 4, Ho * Wo * rl * 4};
    const uint64_t r16[3] = {rl * 2, Wo * rl * 2, Ho * Wo * rl * 2};
    if (p.has_res && p.res16) rc = make_tmap(&plan->tmR, d.residual16, false, 4, dims, d.res_ld ? r16 : s16, box, 64);
    else if (p.has_res) rc = make_tmap(&plan->tmR, d.residual, true, 4, dims, d.res_ld ? r32 : s32, box, 128);
    if (rc) return rc;
    if (p.has_o32) rc = make_tmap(&plan->tmO32, d.out_f32 + ph_off, true, 4, dims, s32, box, 128);
    if (rc) return rc;
    const bool o16_tma = p.has_o16 && !p.has_o32;  // with both outputs the bf16 copy is a direct row store
    if (o16_tma) rc = make_tmap(&plan->tmO16, d.out_bf16 + ph_off, false, 4, dims, s16, box, 64);
    if (rc) return rc;
    if (!p.has_res) plan->tmR = p.has_o32 ? plan->tmO32 : plan->tmO16;
    if (!p.has_o32) plan->tmO32 = plan->tmO16;
    if (!o16_tma) plan->tmO16 = plan->tmO32;
    }
  }
  // B: 3-D (cin, cout, tap)
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(d.C), static_cast<uint64_t>(d.Cout),
                              static_cast<uint64_t>(p.num_taps)};
    const uint64_t wl = d.wt_ld ? d.wt_ld : d.C;
    PBE_REQUIRE(wl >= static_cast<uint64_t>(d.C) && wl % 8 == 0 && (d.wt_ld == 0 || d.ksize == 1), "wt_ld: >= C, multiple of 8, 1x1 only");
    const uint64_t strides[2] = {wl * 2, wl * d.Cout * 2};
    // CTA pair: each CTA loads its half of the n tile
    const uint32_t box[3] = {64u, static_cast<uint32_t>(bn / plan->cluster), 1u};
    int rc = make_tmap_bf16(&plan->tmB, d.wt, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  return 0;
}

static int launch_main(const GemmPlan& plan, cudaStream_t stream);

int launch_gemm_plan(const GemmPlan& plan, cudaStream_t stream) {
  int rc = launch_main(plan, stream);
  if (rc) return rc;
  if (plan.red.S > 1) {
    const long long n4 = plan.red.M * (plan.red.N / 4);
    PBE_CHECK_CUDA(launch_k(splitk_reduce_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, stream, plan.red));
    PBE_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

static int launch_main(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.block_n) {
    // stages: single CTA / CTA pair (half weight tile per CTA)
    case 32: return launch_t<32, 6, 8>(plan, stream);
    case 64: return launch_t<64, 6, 7>(plan, stream);
    case 128: return launch_t<128, 4, 6>(plan, stream);
    case 160: return launch_t<160, 4, 6>(plan, stream);
    case 256: return launch_t<256, 3, 5>(plan, stream);
    default: set_error("launch_gemm_plan: bad block_n"); return -1;
  }
}

}  // namespace pbe
