// Flash-style self-attention for the SpatialTransformer blocks, on tcgen05 / TMEM (sm_100a).
//
//   O = softmax(Q K^T * d^-1/2) V        per (sample, head), N tokens, no mask, never materialising the N x N matrix.
//
// Replaces CrossAttention.forward with context=None (ldm/modules/attention.py:207-230), where the reference builds
// sim[(B*8), N, N] in fp32 (1 GB per layer at N=4096, B=2).
//
// flash_attn_kernel (head dims > 64, short sequences): one CTA = one (sample, head, 128-query tile).  warp 0: TMA
// producer; warp 1: MMA issuer; warps 2..5: softmax (one query row per thread).  S = Q K^T lands in a double-buffered
// 128x128 fp32 TMEM tile, the softmax warps read it with tcgen05.ld, keep the reference maximum / row sum in registers,
// write P (bf16 pairs) into a fourth TMEM region, and the MMA warp accumulates O += P V (A operand from tensor memory)
// into a third.  TMEM: S0 [0,128) S1 [128,256) O [256, 256+dv <= 416) P [448,512).
// flash_attn2_kernel (head dims <= 64, the 64x64-latent level): 256 queries per CTA, 20 warps on five warpgroups -- see its comment.
// Q and K come from the fused projection output [B, N, 2C];
// V arrives transposed ([B, C, N], written by the projection GEMM's epilogue) so both MMAs take K-major operands.
// Head dims that are not a multiple of 64 (40, 80, 160) rely on TMA out-of-bounds zero fill.
#include "internal.h"
#include "ptx.cuh"

#include <stdlib.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>

namespace pbe {

namespace {

constexpr int ATT_THREADS = 320;   // flash_attn_kernel: TMA warp, MMA warp, EIGHT softmax warps (two threads per query row)
constexpr int QT = 128;   // queries per CTA
constexpr int KT = 128;   // keys per tile
constexpr int CHUNK_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr float ATT2_BIAS = 64.0f;      // single-pass tiles: exponentials are kept 2^-64 below the reference maximum

struct AttnParams {
  int N, heads, d, dv, ksteps, C;
  float scale_log2;
  bf16* out;
  int two_pass;   // 1 = exact running maximum in every tile (PBE_ATTN_TWO_PASS=1, and the overflow re-run), 0 = single-pass tiles
  // overflow handling of the single-pass tiles: a work item (one CTA's query block) whose row sum left the safe range is
  // recomputed by the same CTA, in the same launch, with the exact running maximum (flags: per-item marks of the persistent
  // kernel, null = no second pass)
  int* flags;
  int out_f16;               // the output feeds a GEMM: written in the library's operand format (fp16 unless PBE_OPERANDS=bf16)
  int qtiles, total_items;   // flash_attn3_kernel (persistent): 256-query blocks per (sample, head), work items in total
};

// Row sums are kept ATT2_BIAS powers of two below 1; a row sum above this bound means some exponential ran far above the
// reference maximum (the single-pass tiles' failure mode) and P / O may have overflowed: the item is recomputed exactly.
constexpr float ATT_L_SAFE = 1.1529215e18f;   // 2^60

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// O[lane, 0:dv) *= alpha in tensor memory.  Deliberately NOT inlined: it runs in the rare "reference moved" branch of the
// softmax loops, and as a call its register needs (and the spills around it) stay inside that branch.
__device__ __noinline__ void rescale_o_rows(uint32_t taddr, int dv, float alpha) {
  for (int c = 0; c < dv; c += 16) {
    uint32_t o[16];
    tmem_ld_x16(taddr + c, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st_x16(taddr + c, o);
  }
  tmem_st_wait();
}

// VROWS: rows (value channels, dv <= VROWS) a V^T stage is sized for -- at d = 80 a 96-row stage lets THREE K / V stages fit
// next to Q.  With two, tile j + 2 could only be requested once P.V(j) had released its stage, i.e. one tile of compute
// ahead of its use: the loop ran at the TMA round trip (~2 us per 128-key tile, 3.4 x the MUFU floor of its exponentials)
// whatever the softmax warps did -- doubling them changed nothing (tools/attn_probe.py 16 1024 8 80: 115 us either way).
template <int DK_CHUNKS, int KV_STAGES, int VROWS>
__global__ void __launch_bounds__(ATT_THREADS, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                  const __grid_constant__ AttnParams p) {
  griddep_launch_dependents();   // PDL (ptx.cuh): successor may be scheduled; griddep_wait() after the prologue
  // One work item per CTA (query tile, head, sample).  If its single-pass tiles overflow (row sum out of the safe range) the
  // CTA runs the item a second time with the exact running maximum (pass 1) -- same launch, Q still in shared memory; the
  // barrier phases simply continue (tile index jb = pass * T + j).
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int v_chunk_bytes = p.dv * 128;  // one 64-key chunk of V^T: dv rows x 128 B
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + DK_CHUNKS * CHUNK_BYTES;                     // KV_STAGES x DK_CHUNKS chunks
  const uint32_t sV = sK + KV_STAGES * DK_CHUNKS * CHUNK_BYTES;         // KV_STAGES x 2 chunks of (dv x 64)
  const uint32_t sBar = sV + KV_STAGES * 2 * VROWS * 128;               // V region sized for dv <= VROWS
  uint8_t* bar_gen = smem_gen + (sBar - smem_base);

  const uint32_t q_full = sBar;
  auto k_full = [&](int s) { return sBar + 8u * (1 + s); };
  auto v_full = [&](int s) { return sBar + 8u * (1 + KV_STAGES + s); };
  auto kv_empty = [&](int s) { return sBar + 8u * (1 + 2 * KV_STAGES + s); };
  auto s_full = [&](int b) { return sBar + 8u * (1 + 3 * KV_STAGES + b); };
  auto s_free = [&](int b) { return sBar + 8u * (3 + 3 * KV_STAGES + b); };
  const uint32_t p_full = sBar + 8u * (5 + 3 * KV_STAGES);
  const uint32_t pv_done = sBar + 8u * (6 + 3 * KV_STAGES);
  // K stages are released by QK^T(j) itself -- a whole tile before P.V(j) releases the V stage -- so K runs a tile further ahead
  auto k_empty = [&](int s) { return sBar + 8u * (8 + 3 * KV_STAGES + s); };
  const uint32_t tmem_ptr_addr = sBar + 8u * (7 + 3 * KV_STAGES);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(bar_gen + 8 * (7 + 3 * KV_STAGES));
  volatile int* cta_flagged = reinterpret_cast<volatile int*>(bar_gen + 8 * (7 + 3 * KV_STAGES) + 4);
  // exchange between the two threads of a query row: [tile parity][0 maxima | 1 row sums][key half][row]
  float* xch = reinterpret_cast<float*>(bar_gen + 8 * (8 + 4 * KV_STAGES));

  const int warp = static_cast<int>(uniform_u32(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int T = (p.N + KT - 1) / KT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(kv_empty(s), 1);
      mbar_init(k_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(s_free(i), 8);
    }
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    *cta_flagged = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_ptr_gen);
  griddep_wait();   // PDL: the prologue above overlapped the previous kernel; no dependent global access before this
  const uint32_t tmem_O = tmem_base + 256;

  for (int pass = 0; pass < 2; ++pass) {
  const int base = pass * T;                       // tiles before this pass (barrier phases / ring stages continue)
  const bool exact = (p.two_pass != 0) || pass == 1;
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      if (pass == 0) {
        mbar_expect_tx(q_full, DK_CHUNKS * CHUNK_BYTES);
        for (int kc = 0; kc < DK_CHUNKS; ++kc)
          tma_load_4d(sQ + kc * CHUNK_BYTES, &tmQK, q_full, kc * 64, head, q0, b);
      }
      for (int step = 0; step <= T; ++step) {   // K of tile `step`, then V of tile `step - 1`: K is needed (and freed) a tile earlier
        if (step < T) {
          const int j = step, st = (base + j) % KV_STAGES;
          mbar_wait(k_empty(st), (((base + j) / KV_STAGES) & 1) ^ 1u);
          mbar_expect_tx(k_full(st), DK_CHUNKS * CHUNK_BYTES);
          for (int kc = 0; kc < DK_CHUNKS; ++kc)
            tma_load_4d(sK + (st * DK_CHUNKS + kc) * CHUNK_BYTES, &tmQK, k_full(st), kc * 64, p.heads + head, j * KT, b);
        }
        if (step >= 1) {
          const int j = step - 1, st = (base + j) % KV_STAGES;
          mbar_wait(kv_empty(st), (((base + j) / KV_STAGES) & 1) ^ 1u);
          mbar_expect_tx(v_full(st), 2 * v_chunk_bytes);
          for (int jj = 0; jj < 2; ++jj)
            tma_load_3d(sV + st * 2 * VROWS * 128 + jj * v_chunk_bytes, &tmV, v_full(st), j * KT + jj * 64, head * p.d, b);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc_qk = umma_idesc_bf16(128, KT);
    const uint32_t idesc_pv = umma_idesc_bf16(128, p.dv);
    auto issue_qk = [&](int j) {
      const int jb = base + j;
      const int st = jb % KV_STAGES;
      const int sb = jb & 1;
      mbar_wait(k_full(st), (jb / KV_STAGES) & 1);
      mbar_wait(s_free(sb), ((jb >> 1) & 1) ^ 1u);
      tc_fence_after();
      {
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint64_t adesc = umma_desc_sw128(sQ + (ks >> 2) * CHUNK_BYTES) + 2u * (ks & 3);
          const uint64_t bdesc = umma_desc_sw128(sK + (st * DK_CHUNKS + (ks >> 2)) * CHUNK_BYTES) + 2u * (ks & 3);
          umma_bf16_ss_elect(tmem_base + sb * 128, adesc, bdesc, idesc_qk, ks > 0 ? 1u : 0u);
        }
        umma_commit_elect(s_full(sb));
        umma_commit_elect(k_empty(st));   // the K stage is free once these MMAs have read it
      }
      __syncwarp();
    };
    if (pass == 0) mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < T; ++j) {
      if (KV_STAGES >= 2 && j + 1 < T) issue_qk(j + 1);
      const int jb = base + j;
      const int st = jb % KV_STAGES;
      mbar_wait(v_full(st), (jb / KV_STAGES) & 1);
      mbar_wait(p_full, jb & 1);
      tc_fence_after();
      {
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) {
          const uint64_t bdesc = umma_desc_sw128(sV + st * 2 * VROWS * 128 + (ks >> 2) * v_chunk_bytes) + 2u * (ks & 3);
          umma_bf16_ts_elect(tmem_O, tmem_base + 448 + 8u * ks, bdesc, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit_elect(kv_empty(st));   // the V stage
        umma_commit_elect(pv_done);
      }
      __syncwarp();
      if (KV_STAGES < 2 && j + 1 < T) issue_qk(j + 1);
    }
  } else {
    // ================= softmax / correction / output (warps 2..9) =================
    // TWO threads per query row, 64 keys of the tile each (warps 2..5: keys 0..63, warps 6..9: keys 64..127; warp w and
    // w + 4 hold the same 32 rows).  One row per thread left each scheduler a single softmax warp: ncu of the 32 x 32 level
    // showed the kernel latency-bound (XU pipe 31 %, issue slots 18 %, profiles/r01_ncu_attn1_v6_summary.txt).  With two
    // warps per scheduler the exponentials of one hide the dependent chains of the other, and the 64 live logits leave
    // room in the 168 registers a 10-warp CTA gets.  The halves meet twice per tile through 4 KB of shared memory and a
    // 64-thread named barrier per row quarter: the exact maximum (tile 0 / exact pass) and the tile's row sum, which both
    // threads add in the same order, so both carry the same reference, the same running sum and take the same decisions.
    // As in flash_attn2_kernel: the exact row maximum is taken for tile 0 only; later tiles use a reference that follows
    // ref + ATT2_BIAS + log2(row sum) of the previous tile and only moves when that exceeds it by 2^8; FFMA2 / FADD2 packed
    // math; bf16 P goes to tensor memory and is the A operand of P.V.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tmem_P = tmem_base + 448;
    constexpr int HK = KT / 2;   // keys per thread and tile
    float mrs = 0.0f;         // reference maximum, log2 units
    float l_run = 0.0f;
    float sm1 = 0.0f, rf1 = 0.0f;   // row sum of the previous tile and the reference it was computed against
    const float sl2 = p.scale_log2;
    const f32x2 sl2_2 = pk2(sl2, sl2);
    // O columns of this thread in the rare rescale and in the output: 16-column chunks, alternating between the halves
    auto rescale_mine = [&](float alpha) {
      for (int c = half * 16; c < p.dv; c += 32) rescale_o_rows(tmem_O + lane_off + c, 16, alpha);
    };

    // My half of S(j) is requested from tensor memory at the END of tile j - 1, before P(j - 1) is handed over: tcgen05.ld
    // executes in order with the MMAs, so a load issued after p_full would sit behind P.V(j - 1) and QK^T(j + 1) (~700 clk)
    // while the softmax warps idle -- and P.V(j) cannot start before they finish.  (Measured at N = 1024, d = 80: the tile
    // time was the SUM of the tensor work and the exponentials, 2.1 k cycles, not their maximum.)
    float s[HK];   // the tile's logits, then its exponentials: loaded into, transformed in and packed from the same registers
    auto request_s = [&](int jb2) {
      const int sb2 = jb2 & 1;
      mbar_wait(s_full(sb2), (jb2 >> 1) & 1);
      tc_fence_after();
      tmem_ld_x32_f(tmem_base + lane_off + sb2 * 128 + half * HK, s);
      tmem_ld_x32_f(tmem_base + lane_off + sb2 * 128 + half * HK + 32, s + 32);
    };
    request_s(base);
    for (int j = 0; j < T; ++j) {
      const int jb = base + j;
      const int sb = jb & 1;
      float* xmax = xch + (jb & 1) * 512;
      float* xsum = xmax + 256;
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(sb));

      const int kvalid = p.N - j * KT - half * HK;  // keys of my half that are valid in this tile
      if (kvalid < HK) {
#pragma unroll
        for (int i = 0; i < HK; ++i)
          if (i >= kvalid) s[i] = -INFINITY;
      }
      float alpha = 1.0f;
      if (j == 0 || exact) {
        // exact row maximum: tile 0 always; every tile in the exact re-run of a flagged item (classic online softmax)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < HK; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1]));
          mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
        }
        xmax[half * 128 + row] = fmaxf(mx0, mx1);
        named_bar_sync(1 + q, 64);
        const float mt = fmaxf(xmax[row], xmax[128 + row]) * sl2;
        if (j == 0) {
          mrs = mt;
        } else if (__any_sync(0xffffffffu, mt > mrs)) {
          const float m_new = fmaxf(mrs, mt);
          alpha = ex2(mrs - m_new);
          mrs = m_new;
          mbar_wait(pv_done, (jb - 1) & 1);
          tc_fence_after();
          rescale_mine(alpha);
        }
      } else {
        const float est = rf1 + ATT2_BIAS + __log2f(sm1);
        if (__any_sync(0xffffffffu, est - mrs > 8.0f)) {
          const float m_new = fmaxf(mrs, est);
          alpha = ex2(mrs - m_new);
          mrs = m_new;
          mbar_wait(pv_done, (jb - 1) & 1);
          tc_fence_after();
          rescale_mine(alpha);
        }
      }
      const float mneg = -mrs - ATT2_BIAS;
      const f32x2 mneg_2 = pk2(mneg, mneg);
      f32x2 acc0 = pk2(0.0f, 0.0f), acc1 = acc0;
#pragma unroll
      for (int i = 0; i < HK; i += 4) {
        float x0, x1, x2, x3;
        upk2(fma2(pk2(s[i], s[i + 1]), sl2_2, mneg_2), x0, x1);
        upk2(fma2(pk2(s[i + 2], s[i + 3]), sl2_2, mneg_2), x2, x3);
        s[i] = ex2(x0); s[i + 1] = ex2(x1); s[i + 2] = ex2(x2); s[i + 3] = ex2(x3);
        acc0 = add2(acc0, pk2(s[i], s[i + 1]));
        acc1 = add2(acc1, pk2(s[i + 2], s[i + 3]));
      }
      float t0, t1, t2, t3;
      upk2(acc0, t0, t1);
      upk2(acc1, t2, t3);
      xsum[half * 128 + row] = (t0 + t1) + (t2 + t3);

      if (j > 0) {   // P is single-buffered: P.V(j-1) must have read it (pass 1, tile 0: waited for in pass 0's epilogue)
        mbar_wait(pv_done, (jb - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < HK; c += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(s[c + 2 * i], s[c + 2 * i + 1]);
        tmem_st_x16(tmem_P + lane_off + ((half * HK + c) >> 1), pk);
      }
      tmem_st_wait();
      // ahead of P.V(j) in the tensor core's queue (see above) -- when QK^T(j + 1) has been issued ahead of P.V(j) at all:
      // with a single K / V stage it follows P.V(j), and waiting for it here would wait for ourselves
      if (KV_STAGES >= 2 && j + 1 < T) request_s(jb + 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (KV_STAGES < 2 && j + 1 < T) request_s(jb + 1);

      // the tile's row sum: both halves, added in the same order by both threads (the P.V MMA is already running)
      named_bar_sync(1 + q, 64);
      const float tsum = xsum[row] + xsum[128 + row];
      l_run = l_run * alpha + tsum;
      sm1 = tsum;
      rf1 = mrs;
    }

    // ---- final: O / l -> out[b, q0+row, head*d + :] ----
    mbar_wait(pv_done, (base + T - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int tok = q0 + row;
    if (!exact && p.flags != nullptr) {   // single-pass tiles left their safe range: recompute the item exactly (pass 1)
      const bool bad = (tok < p.N) && !(l_run < ATT_L_SAFE);
      if (__any_sync(0xffffffffu, bad) && lane == 0) *cta_flagged = 1;
    }
    bf16* orow = p.out + (static_cast<long long>(b) * p.N + tok) * p.C + head * p.d;
    for (int c = half * 16; c < p.dv; c += 32) {
      uint32_t o[16];
      tmem_ld_x16(tmem_O + lane_off + c, o);
      tmem_ld_wait();
      if (tok < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          if (c + i < p.d) {  // d % 8 == 0
            uint4 pk;
            pk.x = pack_op2(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l, p.out_f16);
            pk.y = pack_op2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l, p.out_f16);
            pk.z = pack_op2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l, p.out_f16);
            pk.w = pack_op2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l, p.out_f16);
            *reinterpret_cast<uint4*>(orow + c + i) = pk;
          }
        }
      }
    }
  }

  if (pass == 1 || p.flags == nullptr || p.two_pass) break;
  tc_fence_before();
  __syncthreads();
  if (*cta_flagged == 0) break;
  tc_fence_after();
  }   // pass

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// Two-warpgroup variant for head dims <= 64 (the 64x64-latent level: N = 4096, d = 40 — 92 % of all attention time).
// One CTA = one (sample, head, 256-query pair of tiles).  Softmax is MUFU(ex2)-bound at d = 40, so the point is to keep
// two softmax warps resident per scheduler: warpgroup A works on S_A(j) while the tensor core produces S_B(j) and the
// P·V products; K/V tiles are loaded once for both query tiles.
// TMEM: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512) -- P (bf16 pairs) is the A
// operand of P.V read straight from tensor memory: tcgen05.mma costs exactly its shared-memory operand bytes / 128 B/clk
// (micro-benchmark, profiles/README.md), and with P staged in shared memory the tensor core's operand reads (1088 clk per
// 256 x 128 tile pair), the softmax warps' P stores (512 clk) and the K/V TMA writes (224 clk) kept the shared-memory /
// MIO path busier than the MUFU pipe they were queued behind (ncu: MUFU.EX2 stalled on mio_throttle, XU 61 %).
// ------------------------------------------------------------------------------------------------------------------
// warps 0..15 softmax (two warpgroups of eight), 16 / 17 MMA issue, 18 Q / K / V TMA, 19 idle: the service warps form a
// hardware warpgroup of their own, give most of their registers back (setmaxnreg.dec) and the softmax warps take them
// (setmaxnreg.inc): at the 96 registers the launch grants 20 warps, the softmax loop (64 logits of S live per thread) spilled
// its loop state and recomputed its addresses every tile -- 40 % of its instructions (ncu, profiles/r01_ncu_attn2_v5).
constexpr int ATT2_THREADS = 640;
constexpr int ATT2_SOFTMAX_REGS = 104, ATT2_SERVICE_REGS = 32;
constexpr int ATT2_KV_STAGES = 3;

// 128-key tiles, 256 queries per CTA, SIXTEEN softmax warps: every query row is shared by two threads (64 keys each),
// so each scheduler always has four softmax warps to pick from.  Findings that shaped this (clock64 traces, ncu):
//  * the softmax is issue/latency bound, not MUFU bound: with two softmax warps per scheduler IPC was ~0.4;
//  * eager rescaling of O waited for the tensor core in ~half of all tiles on random data -> lazy rescaling (the
//    reference maximum only moves when exceeded by 2^8);
//  * one MMA-issuing warp per warpgroup (mbarrier round trips serialise otherwise);
//  * exponentials are computed in place first, sums / bf16 packing / stores afterwards.
//  * S(j+1) = Q K^T is issued as soon as the softmax warps have READ S(j) for the last time (s_free), not when P(j) is
//    published: the tensor-core round trip overlaps the exponentials of tile j instead of idling the warpgroup;
//  * P lives in tensor memory, one buffer per warpgroup: the first P store of tile j waits for P.V(j-1), which was
//    issued a whole half-tile of exponentials earlier;
//  * the K/V ring has its own warp, so neither MMA warp ever blocks on the other warpgroup's P.V.
// warps: 0..7 softmax wg 0 | 8..15 softmax wg 1 | 16, 17 MMA issue for wg 0, 1 | 18 Q / K / V TMA | 19 idle
__global__ void __launch_bounds__(ATT2_THREADS, 1)
flash_attn2_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ AttnParams p) {
  griddep_launch_dependents();   // PDL (ptx.cuh): successor may be scheduled; griddep_wait() after the prologue
  extern __shared__ __align__(1024) uint8_t smem_att2[];   // 1024-B aligned base (128B-swizzle atoms), no padding budget
  uint8_t* smem_raw = smem_att2;
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) __trap();
  uint8_t* smem_gen = smem_raw;
  constexpr int V_STAGE_BYTES = 2 * 64 * 128;  // two 64-key chunks of up to 64 rows
  const int v_chunk_bytes = p.dv * 128;
  const uint32_t sQ = smem_base;                          // 2 query tiles
  const uint32_t sK = sQ + 2 * CHUNK_BYTES;               // ATT2_KV_STAGES stages
  const uint32_t sV = sK + ATT2_KV_STAGES * CHUNK_BYTES;  // ATT2_KV_STAGES stages
  const uint32_t sX = sV + ATT2_KV_STAGES * V_STAGE_BYTES;  // row-max / row-sum exchange: [2 wg][2 halves][128] floats
  const uint32_t sBar = sX + 2 * 2 * 128 * 4;
  uint8_t* bar_gen = smem_gen + (sBar - smem_base);
  float* x_gen = reinterpret_cast<float*>(smem_gen + (sX - smem_base));
  constexpr int ST = ATT2_KV_STAGES;
  const uint32_t q_full = sBar;
  auto kv_full = [&](int s) { return sBar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return sBar + 8u * (1 + ST + s); };
  auto s_full = [&](int g) { return sBar + 8u * (1 + 2 * ST + g); };
  auto p_full = [&](int g) { return sBar + 8u * (3 + 2 * ST + g); };
  auto pv_done = [&](int g, int par) { return sBar + 8u * (5 + 2 * ST + g * 2 + par); };  // P.V of the even / odd tiles
  auto s_free = [&](int g) { return sBar + 8u * (9 + 2 * ST + g); };
  const uint32_t tmem_ptr_addr = sBar + 8u * (11 + 2 * ST);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(bar_gen + 8 * (11 + 2 * ST));

  const int warp = static_cast<int>(uniform_u32(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * QT;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int T = (p.N + KT - 1) / KT;

  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT2_KV_STAGES; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 2);   // one tcgen05.commit arrival per MMA warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 8);
      mbar_init(pv_done(s, 0), 1);
      mbar_init(pv_done(s, 1), 1);
      mbar_init(s_free(s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 16) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_ptr_gen);
  griddep_wait();   // PDL: the prologue above overlapped the previous kernel; no dependent global access before this

  // setmaxnreg sits at the top of each role's branch (ptxas bounds the code it dominates); all four warps of the service
  // warpgroup execute the .dec
  if (warp == 16 || warp == 17) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_SERVICE_REGS));
    // ================= MMA issue for warpgroup g (uniform code, elect.sync issues) =================
    const int g = warp - 16;
    const uint32_t idesc_qk = umma_idesc_bf16(128, KT);
    const uint32_t idesc_pv = umma_idesc_bf16(128, p.dv);
    const uint64_t qdesc = umma_desc_sw128(sQ + g * CHUNK_BYTES);
    const uint32_t tmem_S = tmem_base + g * 128;
    const uint32_t tmem_O = tmem_base + 256 + g * 64;
    const uint32_t tmem_P = tmem_base + 384 + g * 64;
    auto issue_qk = [&](int j) {
      const uint64_t kdesc = umma_desc_sw128(sK + (j % ATT2_KV_STAGES) * CHUNK_BYTES);
      for (int ks = 0; ks < p.ksteps; ++ks) umma_bf16_ss_elect(tmem_S, qdesc + 2u * ks, kdesc + 2u * ks, idesc_qk, ks > 0 ? 1u : 0u);
      umma_commit_elect(s_full(g));
    };
    mbar_wait(q_full, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    issue_qk(0);
    int st = 0, st_next = 1;            // stage of tile j / j + 1
    uint32_t ph_next = 0;               // kv_full parity of tile j + 1
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) {
        mbar_wait(s_free(g), j & 1);    // every softmax warp has read S(j) for the last time
        mbar_wait(kv_full(st_next), ph_next);
        tc_fence_after();
        issue_qk(j + 1);
      }
      mbar_wait(p_full(g), j & 1);      // P(j) is in tensor memory
      tc_fence_after();
      {
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) {
          const uint64_t vdesc = umma_desc_sw128(sV + st * V_STAGE_BYTES + (ks >> 2) * v_chunk_bytes) + 2u * (ks & 3);
          umma_bf16_ts_elect(tmem_O, tmem_P + 8u * ks, vdesc, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit_elect(pv_done(g, j & 1));
        umma_commit_elect(kv_empty(st));
      }
      st = st_next;
      if (++st_next == ATT2_KV_STAGES) { st_next = 0; ph_next ^= 1u; }
    }
  } else if (warp == 19) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_SERVICE_REGS));
  } else if (warp == 18) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_SERVICE_REGS));
    // ================= Q / K / V loads (uniform code, elected issue) =================
    mbar_expect_tx_elect(q_full, 2 * CHUNK_BYTES);
    tma_load_4d_elect(sQ, &tmQK, q_full, 0, head, q0, b);
    tma_load_4d_elect(sQ + CHUNK_BYTES, &tmQK, q_full, 0, head, q0 + QT, b);
    int st = 0;
    uint32_t ph = 0;
    for (int t = 0; t < T; ++t) {
      if (t >= ATT2_KV_STAGES) mbar_wait(kv_empty(st), ph ^ 1u);   // P.V of the tile that used this stage has retired
      mbar_expect_tx_elect(kv_full(st), CHUNK_BYTES + 2 * v_chunk_bytes);
      tma_load_4d_elect(sK + st * CHUNK_BYTES, &tmQK, kv_full(st), 0, p.heads + head, t * KT, b);
      tma_load_3d_elect<false>(sV + st * V_STAGE_BYTES, &tmV, kv_full(st), t * KT, head * p.d, b);
      tma_load_3d_elect<false>(sV + st * V_STAGE_BYTES + v_chunk_bytes, &tmV, kv_full(st), t * KT + 64, head * p.d, b);
      if (++st == ATT2_KV_STAGES) { st = 0; ph ^= 1u; }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ATT2_SOFTMAX_REGS));
    const int sw = warp;                // 0..15
    const int g = sw >> 3;              // warpgroup = query tile
    const int sub = (sw >> 2) & 1;      // which 64-key half of the row this thread owns
    const int q = warp & 3;             // TMEM lane quarter
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tmem_S = tmem_base + g * 128 + sub * 64;
    const uint32_t tmem_O = tmem_base + 256 + g * 64;
    const uint32_t tmem_P = tmem_base + 384 + g * 64 + sub * 32;   // my 64 keys of the row = 32 columns of bf16 pairs
    // bf16 P -> tensor memory (A operand of P.V): 32 exponentials -> 16 columns
    auto store_p = [&](const uint32_t (&v)[32], int h) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      tmem_st_x16(tmem_P + lane_off + h * 16, pk);
    };
    float* xmine = x_gen + (g * 2 + sub) * 128 + row;
    float* xpeer = x_gen + (g * 2 + (sub ^ 1)) * 128 + row;
    const int bar_id = 1 + g;
    float m_run = -INFINITY;   // reference maximum used by the exponentials (lazy)
    float l_run = 0.0f;        // partial row sum over this thread's columns
    const float sl2 = p.scale_log2;
    const float bias = p.two_pass ? 0.0f : ATT2_BIAS;

    auto tile = [&](int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      mbar_wait(s_full(g), j & 1);
      tc_fence_after();
      // pass 1: partial row maximum over my 64 columns, 32 at a time (values are re-read from TMEM in pass 2:
      // 32 live registers instead of 64 keeps this rarely used path small)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_x32(tmem_S + lane_off + h * 32, v);
        tmem_ld_wait();
        if (MASKED) {
          const int kvalid = p.N - j * KT - sub * 64 - h * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= kvalid) v[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[i + 0]), __uint_as_float(v[i + 4])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[i + 1]), __uint_as_float(v[i + 5])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 6])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[i + 3]), __uint_as_float(v[i + 7])));
        }
      }
      const float mpart = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      *xmine = mpart;
      named_bar_sync(bar_id, 256);
      const float mx = fmaxf(mpart, *xpeer);   // identical in both halves of the row
      // lazy rescaling: both halves of a row take the same decision
      float alpha = 1.0f;
      if (__any_sync(0xffffffffu, (mx - m_run) * sl2 > 8.0f)) {
        const float m_new = fmaxf(m_run, mx);
        alpha = ex2((m_run - m_new) * sl2);  // first tile: ex2(-inf) = 0
        m_run = m_new;
        if (j > 0 && sub == 0) {
          mbar_wait(pv_done(g, (j - 1) & 1), ((j - 1) >> 1) & 1);
          tc_fence_after();
          rescale_o_rows(tmem_O + lane_off, p.dv, alpha);
        }
      }
      const float mneg = fmaf(-m_run, sl2, -bias);
      // pass 2: exponentials, partial row sum, bf16 P -> tensor memory
      float sum0 = 0.0f, sum1 = 0.0f, sum2 = 0.0f, sum3 = 0.0f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_x32(tmem_S + lane_off + h * 32, v);
        tmem_ld_wait();
        if (h == 1) {  // last read of S(j): the tensor core may start S(j+1) while the exponentials run
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free(g));
        }
        if (MASKED) {
          const int kvalid = p.N - j * KT - sub * 64 - h * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= kvalid) v[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(ex2(fmaf(__uint_as_float(v[i]), sl2, mneg)));
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          sum0 += __uint_as_float(v[c + 0]) + __uint_as_float(v[c + 4]);
          sum1 += __uint_as_float(v[c + 1]) + __uint_as_float(v[c + 5]);
          sum2 += __uint_as_float(v[c + 2]) + __uint_as_float(v[c + 6]);
          sum3 += __uint_as_float(v[c + 3]) + __uint_as_float(v[c + 7]);
        }
        if (h == 0 && j > 0) {   // P is single-buffered: P.V(j-1) must have read it
          mbar_wait(pv_done(g, (j - 1) & 1), ((j - 1) >> 1) & 1);
          tc_fence_after();
        }
        store_p(v, h);
      }
      l_run = l_run * alpha + ((sum0 + sum1) + (sum2 + sum3));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
    };
    // ---- single-pass tiles (every tile but the first) ----
    // The exponentials of tile j do not need the maximum of tile j: any reference works as long as nothing overflows, and
    // the softmax warps are bound by their instruction count (ncu: 6.5 issued instructions per exponential in the two-pass
    // loop -- at the MUFU rate of one warp-instruction per 8 clk and four warps per scheduler that is 82 % of all issue
    // slots).  Tile 0 takes the exact row maximum (two passes, above).  Later tiles read S ONCE and spend ~2.5 math
    // instructions per exponential: FFMA2 (two logits), MUFU.EX2, FADD2 (two sums), F2FP (two bf16) -- and no maximum at all:
    //  * the row sum of a tile bounds its maximum: max <= log2(sum) <= max + 7.  The sum is needed anyway, so the reference
    //    follows  ref(tile) + ATT2_BIAS + log2(row sum of the tile)  and only moves when that exceeds it by 2^8.
    //  * exchange: a row is shared by two threads (64 keys each).  At the start of tile j a thread publishes the bf16-rounded
    //    sum over ITS columns of tile j-1 and reads its partner's sum of tile j-2; both add the same two rounded numbers, so
    //    both take the same decision.  Ordering comes from the S hand-off that already exists: a partner's slot of tile j-2
    //    was written before it arrived on s_free(j-1), which precedes S(j) and so my s_full(j) wait; the slot is rewritten
    //    (tile j) only after s_full(j+1), i.e. after my s_free(j) arrive, which follows my read.
    //  * range: values are kept ATT2_BIAS powers of two below 1 (the common factor cancels in O / l): a logit may exceed every
    //    key of the tiles up to j-2 by (127 + ATT2_BIAS) / log2(e) = 132 nats before fp32 overflows (the output row is then
    //    NaN, not silently wrong; fp32 softmax is a one-hot long before that), and terms 2^-55 below the maximum still count.
    //    PBE_ATTN_TWO_PASS=1 selects the exact running maximum for every tile.
    //  * both halves of S(j) are loaded before the first exponential, so the tensor core gets s_free(j) -- and produces
    //    S(j+1) -- a whole tile of exponentials ahead of its use.
    float mrs = 0.0f;                    // reference maximum in log2 units
    float sm1 = 0.0f, sm2 = 0.0f;        // bf16-rounded sums over my columns of tiles j-1 and j-2 ...
    float rf1 = 0.0f, rf2 = 0.0f;        // ... and the references those tiles were computed against
    // [2 wg][2 slots][2 halves][128] bf16: a warpgroup's slots alias ITS OWN 1 KB of xmine / xpeer (tile 0, final row sums),
    // whose uses are ordered against the slots by the warpgroup's own barriers -- the two warpgroups run unsynchronised
    const uint32_t xs_mine = sX + (((g * 2 + 0) * 2 + sub) * 128 + row) * 2;        // slot s: + s * 512 bytes
    const uint32_t xs_peer = sX + (((g * 2 + 0) * 2 + (sub ^ 1)) * 128 + row) * 2;
    const f32x2 sl2_2 = pk2(sl2, sl2);
    auto tile1p = [&](int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      mbar_wait(s_full(g), j & 1);
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld_x32(tmem_S + lane_off, va);        // in flight during the exchange below
      tmem_ld_x32(tmem_S + lane_off + 32, vb);
      float alpha = 1.0f;
      if (j >= 2) {
        // sm1 is already bf16-rounded: its upper half is the bf16 pattern (shared-space accesses: no generic address math)
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(xs_mine + (((j - 1) & 1) << 9)),
                     "h"(static_cast<unsigned short>(__float_as_uint(sm1) >> 16)) : "memory");
        if (j >= 3) {
          unsigned short peer_bits;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(peer_bits) : "r"(xs_peer + ((j & 1) << 9)) : "memory");
          const float peer = __uint_as_float(static_cast<uint32_t>(peer_bits) << 16);
          const float est = rf2 + ATT2_BIAS + __log2f(sm2 + peer);   // >= the row maximum of tile j-2, by at most 7
          if (__any_sync(0xffffffffu, est - mrs > 8.0f)) {
            const float m_new = fmaxf(mrs, est);
            alpha = ex2(mrs - m_new);
            mrs = m_new;
            if (sub == 0) {
              mbar_wait(pv_done(g, (j - 1) & 1), ((j - 1) >> 1) & 1);
              tc_fence_after();
              rescale_o_rows(tmem_O + lane_off, p.dv, alpha);
            }
          }
        }
      }
      const float mneg = -mrs - ATT2_BIAS;
      const f32x2 mneg_2 = pk2(mneg, mneg);
      f32x2 acc0 = pk2(0.0f, 0.0f), acc1 = acc0;
      auto half = [&](uint32_t (&v)[32], int h) {
        if (MASKED) {
          const int kvalid = p.N - j * KT - sub * 64 - h * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= kvalid) v[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1;
          upk2(fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sl2_2, mneg_2), x0, x1);
          v[i] = __float_as_uint(ex2(x0));
          v[i + 1] = __float_as_uint(ex2(x1));
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          acc0 = add2(acc0, pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          acc1 = add2(acc1, pk2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
        }
        if (h == 0) {   // P is single-buffered: P.V(j-1) must have read it (j >= 1 here)
          mbar_wait(pv_done(g, (j - 1) & 1), ((j - 1) >> 1) & 1);
          tc_fence_after();
        }
        store_p(v, h);
      };
      tmem_ld_wait_x32x2(va, vb);
      tc_fence_before();                         // last read of S(j): the tensor core may start S(j+1)
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(g));
      half(va, 0);
      half(vb, 1);
      float t0, t1, t2, t3;
      upk2(acc0, t0, t1);
      upk2(acc1, t2, t3);
      const float tsum = (t0 + t1) + (t2 + t3);
      l_run = l_run * alpha + tsum;
      sm2 = sm1; rf2 = rf1;
      sm1 = __bfloat162float(__float2bfloat16_rn(tsum)); rf1 = mrs;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
    };
    const int T_full = p.N / KT;
    if (p.two_pass) {
      for (int j = 0; j < T_full; ++j) tile(j, std::false_type{});
      if (T_full < T) tile(T_full, std::true_type{});
    } else {
      tile(0, std::false_type{});          // N > 128 here: the first tile is always full
      mrs = m_run * sl2;
      for (int j = 1; j < T_full; ++j) tile1p(j, std::false_type{});
      if (T_full < T) tile1p(T_full, std::true_type{});
    }

    // combine the two partial row sums, then the `sub == 0` thread of each row writes O / l
    named_bar_sync(bar_id, 256);   // all reads of the max-exchange slots are done
    *xmine = l_run;
    named_bar_sync(bar_id, 256);
    const float l_tot = l_run + *xpeer;
    mbar_wait(pv_done(g, (T - 1) & 1), ((T - 1) >> 1) & 1);
    tc_fence_after();
    if (sub == 0) {
      const float inv_l = 1.0f / l_tot;
      const int tok = q0 + g * QT + row;
      bf16* orow = p.out + (static_cast<long long>(b) * p.N + tok) * p.C + head * p.d;
      for (int c = 0; c < p.dv; c += 16) {
        uint32_t o[16];
        tmem_ld_x16(tmem_O + lane_off + c, o);
        tmem_ld_wait();
        if (tok < p.N) {
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            if (c + i < p.d) {
              uint4 pk;
              pk.x = pack_op2(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l, p.out_f16);
              pk.y = pack_op2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l, p.out_f16);
              pk.z = pack_op2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l, p.out_f16);
              pk.w = pack_op2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l, p.out_f16);
              *reinterpret_cast<uint4*>(orow + c + i) = pk;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// flash_attn3_kernel: flash_attn2_kernel (same roles, same TMEM map, same single-pass tiles) with the two changes its ncu
// profile asked for (profiles/r01_ncu_attn2_v6_summary.txt: XU pipe 76.5 %, issue slots 47 %, tensor pipe 28 %):
//  * PERSISTENT.  One CTA per SM walks work items (256-query block, head, sample), block index fastest so that the CTAs
//    running at the same time share K / V in L2.  Barrier phases run on a global tile counter, Q is double-buffered, so the
//    service warps run ahead across items: the next item's Q / K / V loads and its first Q.K^T overlap the current item's
//    last exponentials and its O / l epilogue.  Launch, TMEM allocation, barrier set-up and the first load round trip
//    (about 7 % of a 48 us CTA before) are paid once per SM instead of once per item.
//  * PART OF THE EXPONENTIALS ON THE FMA PIPE.  At d = 40 the MUFU pipe (16 ex2 / clk / SM) is the floor: 2048 clk per
//    128-key tile pair against ~800 clk of tensor work.  POLY16 of every 16 exponentials are evaluated as
//    2^x = 2^n * p(f), n = round(x), f = x - n in [-1/2, 1/2], p a degree-3 minimax polynomial (max relative error 1.0e-4
//    in fp32 -- P is rounded to bf16, 3.9e-3, right after): magic-number rounding, three FFMA2 (two logits per
//    instruction) and one integer shift-add into the exponent field; 5 issue slots per exponential instead of 1, on pipes
//    that were half empty.  (The same idea as FlashAttention-4's software exp2; here the share is a template parameter.)
// Overflow of the single-pass tiles (a logit more than ~132 nats above every earlier key of the row) is DETECTED: the
// row sum leaves its safe range, the item's flag is set and launch_attn3 runs the exact (two-pass) variant over the
// flagged items right behind -- the row is recomputed instead of coming out as NaN.
// ------------------------------------------------------------------------------------------------------------------
constexpr float EXP2_C0 = 0.9999281168f, EXP2_C1 = 0.6932609677f, EXP2_C2 = 0.2426108569f, EXP2_C3 = 0.05517145991f;

// 2^x for two logits on the FMA / ALU pipes.  t = x + 1.5 * 2^23 holds round(x) in its low mantissa bits; t - 1.5 * 2^23
// and x - round(x) are exact; (bits(t) << 23) is n << 23 (the constant's bits leave the word), added to p's bits it scales
// p(f) in [0.707, 1.414] by 2^n.  x is clamped at -125 so the exponent field cannot wrap (2^-125 is 0 next to any row sum).
__device__ __forceinline__ void exp2_fma2(f32x2 x, uint32_t& o0, uint32_t& o1) {
  float x0, x1;
  upk2(x, x0, x1);
  const f32x2 xc = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const f32x2 t = add2(xc, pk2(12582912.0f, 12582912.0f));
  const f32x2 r = add2(t, pk2(-12582912.0f, -12582912.0f));
  const f32x2 f = fma2(r, pk2(-1.0f, -1.0f), xc);
  f32x2 pl = fma2(pk2(EXP2_C3, EXP2_C3), f, pk2(EXP2_C2, EXP2_C2));
  pl = fma2(pl, f, pk2(EXP2_C1, EXP2_C1));
  pl = fma2(pl, f, pk2(EXP2_C0, EXP2_C0));
  float p0, p1, t0, t1;
  upk2(pl, p0, p1);
  upk2(t, t0, t1);
  o0 = __float_as_uint(p0) + (__float_as_uint(t0) << 23);
  o1 = __float_as_uint(p1) + (__float_as_uint(t1) << 23);
}

// which of every 8 logit pairs take the FMA-pipe path, spread so the two kinds interleave in program order
template <int POLY16> struct PolyMask;
template <> struct PolyMask<0> { static constexpr unsigned value = 0x00u; };
template <> struct PolyMask<2> { static constexpr unsigned value = 0x08u; };
template <> struct PolyMask<4> { static constexpr unsigned value = 0x22u; };
template <> struct PolyMask<6> { static constexpr unsigned value = 0x4Au; };
template <> struct PolyMask<8> { static constexpr unsigned value = 0xAAu; };

constexpr int ATT3_ST = 3;   // K / V ring depth
// shared memory: 2 Q buffers x 2 query tiles | K ring | V ring | exchange (tile-0 maxima + 16-bit sums, final sums) | barriers
constexpr int ATT3_V_STAGE_BYTES = 2 * 64 * 128;
constexpr int ATT3_NBAR = 4 + 2 * ATT3_ST + 12;   // q_full[2] q_free[2] kv_full[ST] kv_empty[ST] s_full[2] p_full[2] pv_done[4] s_free[2] + tmem ptr + pad
constexpr size_t ATT3_SMEM = 4 * CHUNK_BYTES + ATT3_ST * CHUNK_BYTES + ATT3_ST * ATT3_V_STAGE_BYTES + 2048 + 2048 + 8 * ATT3_NBAR;

template <int POLY16>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
flash_attn3_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ AttnParams p) {
  griddep_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_att3[];
  uint8_t* smem_gen = smem_att3;
  const uint32_t smem_base = smem_u32(smem_gen);
  if ((smem_base & 1023u) != 0u) __trap();
  const int v_chunk_bytes = p.dv * 128;
  const uint32_t sQ = smem_base;                              // [2 buffers][2 query tiles]
  const uint32_t sK = sQ + 4 * CHUNK_BYTES;
  const uint32_t sV = sK + ATT3_ST * CHUNK_BYTES;
  const uint32_t sX = sV + ATT3_ST * ATT3_V_STAGE_BYTES;      // tile-0 maxima (fp32) / running 16-bit sums, as in flash_attn2_kernel
  const uint32_t sXF = sX + 2048;                             // final row sums: [2 wg][2 halves][128] floats (own region: the
                                                              // next item's tile-0 exchange must not overwrite them mid-read)
  const uint32_t sBar = sXF + 2048;
  float* x_gen = reinterpret_cast<float*>(smem_gen + (sX - smem_base));
  float* xf_gen = reinterpret_cast<float*>(smem_gen + (sXF - smem_base));
  constexpr int ST = ATT3_ST;
  auto q_full = [&](int qb) { return sBar + 8u * qb; };
  auto q_free = [&](int qb) { return sBar + 8u * (2 + qb); };
  auto kv_full = [&](int s) { return sBar + 8u * (4 + s); };
  auto kv_empty = [&](int s) { return sBar + 8u * (4 + ST + s); };
  auto s_full = [&](int g) { return sBar + 8u * (4 + 2 * ST + g); };
  auto p_full = [&](int g) { return sBar + 8u * (6 + 2 * ST + g); };
  auto pv_done = [&](int g, uint32_t par) { return sBar + 8u * (8 + 2 * ST + g * 2 + par); };
  auto s_free = [&](int g) { return sBar + 8u * (12 + 2 * ST + g); };
  const uint32_t tmem_ptr_addr = sBar + 8u * (14 + 2 * ST);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_addr - smem_base));

  const int warp = static_cast<int>(uniform_u32(threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  const int T = (p.N + KT - 1) / KT;
  const int total = p.total_items;
  const int step = static_cast<int>(gridDim.x);
  // Two passes over this CTA's items inside ONE launch: pass 0 = every item with single-pass tiles; pass 1 = the items
  // whose row sums left the safe range in pass 0 (normally none: the CTA then leaves after one barrier), recomputed with
  // the exact running maximum.  The flags are written by this CTA's softmax warps in pass 0, read by every role in pass 1
  // (after a CTA barrier) and cleared at the end.
  volatile int* cta_flagged = reinterpret_cast<volatile int*>(smem_gen + (tmem_ptr_addr + 4u - smem_base));
  auto next_live = [&](int item, int pass) {
    if (pass != 0)
      while (item < total && *reinterpret_cast<volatile int*>(p.flags + item) == 0) item += step;
    return item;
  };
  // between the passes: every thread of the CTA meets here; true = some item of this CTA needs the exact pass
  auto second_pass = [&]() {
    if (p.flags == nullptr || p.two_pass) return false;
    named_bar_sync(3, ATT2_THREADS);
    return *cta_flagged != 0;
  };
  auto coords = [&](int item, int& q0, int& head, int& b) {
    const int qt = item % p.qtiles;
    const int r = item / p.qtiles;
    head = r % p.heads;
    b = r / p.heads;
    q0 = qt * 2 * QT;
  };

  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full(i), 1);
      mbar_init(q_free(i), 2);       // one tcgen05.commit per MMA warp
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 8);
      mbar_init(pv_done(i, 0), 1);
      mbar_init(pv_done(i, 1), 1);
      mbar_init(s_free(i), 8);
    }
    for (int s = 0; s < ST; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 2);
    }
    *cta_flagged = 0;
    fence_barrier_init();
  }
  if (warp == 16) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_ptr_gen);
  griddep_wait();

  const int first = static_cast<int>(blockIdx.x);

  if (warp == 16 || warp == 17) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_SERVICE_REGS));
    // ================= MMA issue for warpgroup g: one flat sequence of tiles over all items of this CTA =================
    const int g = warp - 16;
    const uint32_t idesc_qk = umma_idesc_bf16(128, KT);
    const uint32_t idesc_pv = umma_idesc_bf16(128, p.dv);
    const uint32_t tmem_S = tmem_base + g * 128;
    const uint32_t tmem_O = tmem_base + 256 + g * 64;
    const uint32_t tmem_P = tmem_base + 384 + g * 64;
    auto issue_qk = [&](int qb, int st) {
      const uint64_t qdesc = umma_desc_sw128(sQ + (qb * 2 + g) * CHUNK_BYTES);
      const uint64_t kdesc = umma_desc_sw128(sK + st * CHUNK_BYTES);
      for (int ks = 0; ks < p.ksteps; ++ks) umma_bf16_ss_elect(tmem_S, qdesc + 2u * ks, kdesc + 2u * ks, idesc_qk, ks > 0 ? 1u : 0u);
      umma_commit_elect(s_full(g));
    };
    uint32_t k = 0;                      // items so far (Q buffer / phase)
    uint32_t gt = 0;                     // tiles so far
    int st = 0;                          // K / V stage of tile gt ...
    uint32_t ph = 0;                     // ... and its kv_full parity
    for (int pass = 0; pass < 2; ++pass) {
      int item = next_live(first, pass);
      if (item < total) {
        // first tile of the pass: nothing is in flight
        if (gt > 0) mbar_wait(s_free(g), (gt - 1u) & 1u);
        mbar_wait(q_full(k & 1u), (k >> 1) & 1u);
        mbar_wait(kv_full(st), ph);
        tc_fence_after();
        issue_qk(static_cast<int>(k & 1u), st);
        if (T == 1) umma_commit_elect(q_free(k & 1u));
        while (true) {
          const int nxt = next_live(item + step, pass);
          for (int j = 0; j < T; ++j, ++gt) {
            const bool last = (j + 1 == T);
            const int st_next = (st + 1 == ST) ? 0 : st + 1;
            const uint32_t ph_next = (st + 1 == ST) ? (ph ^ 1u) : ph;
            if (!last || nxt < total) {
              // S(gt + 1): as soon as every softmax warp has read S(gt) for the last time
              mbar_wait(s_free(g), gt & 1u);
              const uint32_t kk = last ? k + 1 : k;          // the item tile gt + 1 belongs to
              if (last) mbar_wait(q_full(kk & 1u), (kk >> 1) & 1u);
              mbar_wait(kv_full(st_next), ph_next);
              tc_fence_after();
              issue_qk(static_cast<int>(kk & 1u), st_next);
              // the last Q.K^T of an item: its Q buffer may be refilled once these MMAs have retired
              if (last ? (T == 1) : (j + 2 == T)) umma_commit_elect(q_free(kk & 1u));
            }
            mbar_wait(p_full(g), gt & 1u);      // P(gt) is in tensor memory
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < KT / 16; ++ks) {
              const uint64_t vdesc = umma_desc_sw128(sV + st * ATT3_V_STAGE_BYTES + (ks >> 2) * v_chunk_bytes) + 2u * (ks & 3);
              umma_bf16_ts_elect(tmem_O, tmem_P + 8u * ks, vdesc, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit_elect(pv_done(g, gt & 1u));
            umma_commit_elect(kv_empty(st));
            st = st_next;
            ph = ph_next;
          }
          ++k;
          if (nxt >= total) break;
          item = nxt;
        }
      }
      if (pass == 1 || !second_pass()) break;
    }
  } else if (warp == 19) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_SERVICE_REGS));
    second_pass();
  } else if (warp == 18) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_SERVICE_REGS));
    // ================= Q / K / V loads, running ahead of the other roles across items =================
    uint32_t k = 0, gt = 0;
    int st = 0;
    uint32_t ph = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int item = next_live(first, pass); item < total; item = next_live(item + step, pass), ++k) {
        int q0, head, b;
        coords(item, q0, head, b);
        const int qb = static_cast<int>(k & 1u);
        if (k >= 2) mbar_wait(q_free(qb), ((k >> 1) - 1u) & 1u);   // the Q.K^T MMAs of item k - 2 have retired
        mbar_expect_tx_elect(q_full(qb), 2 * CHUNK_BYTES);
        tma_load_4d_elect(sQ + (qb * 2) * CHUNK_BYTES, &tmQK, q_full(qb), 0, head, q0, b);
        tma_load_4d_elect(sQ + (qb * 2 + 1) * CHUNK_BYTES, &tmQK, q_full(qb), 0, head, q0 + QT, b);
        for (int t = 0; t < T; ++t, ++gt) {
          if (gt >= static_cast<uint32_t>(ST)) mbar_wait(kv_empty(st), ph ^ 1u);   // both P.V of the tile that used this stage retired
          mbar_expect_tx_elect(kv_full(st), CHUNK_BYTES + 2 * v_chunk_bytes);
          tma_load_4d_elect(sK + st * CHUNK_BYTES, &tmQK, kv_full(st), 0, p.heads + head, t * KT, b);
          tma_load_3d_elect<false>(sV + st * ATT3_V_STAGE_BYTES, &tmV, kv_full(st), t * KT, head * p.d, b);
          tma_load_3d_elect<false>(sV + st * ATT3_V_STAGE_BYTES + v_chunk_bytes, &tmV, kv_full(st), t * KT + 64, head * p.d, b);
          if (++st == ST) { st = 0; ph ^= 1u; }
        }
      }
      if (pass == 1 || !second_pass()) break;
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ATT2_SOFTMAX_REGS));
    // ================= softmax (16 warps: warpgroup g = query tile, two threads per row) =================
    const int g = warp >> 3;
    const int sub = (warp >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tmem_S = tmem_base + g * 128 + sub * 64;
    const uint32_t tmem_O = tmem_base + 256 + g * 64;
    const uint32_t tmem_P = tmem_base + 384 + g * 64 + sub * 32;
    auto store_p = [&](const uint32_t (&v)[32], int h) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      tmem_st_x16(tmem_P + lane_off + h * 16, pk);
    };
    float* xmine = x_gen + (g * 2 + sub) * 128 + row;
    float* xpeer = x_gen + (g * 2 + (sub ^ 1)) * 128 + row;
    float* xfmine = xf_gen + (g * 2 + sub) * 128 + row;
    float* xfpeer = xf_gen + (g * 2 + (sub ^ 1)) * 128 + row;
    const int bar_id = 1 + g;
    const float sl2 = p.scale_log2;
    const f32x2 sl2_2 = pk2(sl2, sl2);
    const uint32_t xs_mine = sX + (((g * 2 + 0) * 2 + sub) * 128 + row) * 2;
    const uint32_t xs_peer = sX + (((g * 2 + 0) * 2 + (sub ^ 1)) * 128 + row) * 2;
    const int T_full = p.N / KT;
    uint32_t gt = 0;   // tiles so far (barrier phases)

    for (int pass = 0; pass < 2; ++pass) {
    const bool exact = (p.two_pass != 0) || pass == 1;
    const float bias = exact ? 0.0f : ATT2_BIAS;
    for (int item = next_live(first, pass); item < total; item = next_live(item + step, pass)) {
      int q0, head, b;
      coords(item, q0, head, b);
      float m_run = -INFINITY;   // reference maximum of the two-pass tiles
      float l_run = 0.0f;        // partial row sum over this thread's columns
      float mrs = 0.0f;          // reference maximum in log2 units (single-pass tiles)
      float sm1 = 0.0f, sm2 = 0.0f, rf1 = 0.0f, rf2 = 0.0f;

      // exact tile (tile 0 of every item; every tile of the exact re-run): see flash_attn2_kernel
      auto tile = [&](int j, auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        mbar_wait(s_full(g), gt & 1u);
        tc_fence_after();
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_x32(tmem_S + lane_off + h * 32, v);
          tmem_ld_wait();
          if (MASKED) {
            const int kvalid = p.N - j * KT - sub * 64 - h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kvalid) v[i] = 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[i + 0]), __uint_as_float(v[i + 4])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[i + 1]), __uint_as_float(v[i + 5])));
            mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 6])));
            mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[i + 3]), __uint_as_float(v[i + 7])));
          }
        }
        const float mpart = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        *xmine = mpart;
        named_bar_sync(bar_id, 256);
        const float mx = fmaxf(mpart, *xpeer);
        float alpha = 1.0f;
        if (__any_sync(0xffffffffu, (mx - m_run) * sl2 > 8.0f)) {
          const float m_new = fmaxf(m_run, mx);
          alpha = ex2((m_run - m_new) * sl2);
          m_run = m_new;
          if (j > 0 && sub == 0) {
            mbar_wait(pv_done(g, (gt - 1u) & 1u), ((gt - 1u) >> 1) & 1u);
            tc_fence_after();
            rescale_o_rows(tmem_O + lane_off, p.dv, alpha);
          }
        }
        const float mneg = fmaf(-m_run, sl2, -bias);
        float sum0 = 0.0f, sum1 = 0.0f, sum2 = 0.0f, sum3 = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_x32(tmem_S + lane_off + h * 32, v);
          tmem_ld_wait();
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(g));
          }
          if (MASKED) {
            const int kvalid = p.N - j * KT - sub * 64 - h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kvalid) v[i] = 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(ex2(fmaf(__uint_as_float(v[i]), sl2, mneg)));
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            sum0 += __uint_as_float(v[c + 0]) + __uint_as_float(v[c + 4]);
            sum1 += __uint_as_float(v[c + 1]) + __uint_as_float(v[c + 5]);
            sum2 += __uint_as_float(v[c + 2]) + __uint_as_float(v[c + 6]);
            sum3 += __uint_as_float(v[c + 3]) + __uint_as_float(v[c + 7]);
          }
          if (h == 0 && j > 0) {   // P is single-buffered: P.V of the previous tile must have read it.  (Tile 0 of an item:
                                   // every softmax thread waited for the previous item's last P.V in its epilogue.)
            mbar_wait(pv_done(g, (gt - 1u) & 1u), ((gt - 1u) >> 1) & 1u);
            tc_fence_after();
          }
          store_p(v, h);
        }
        l_run = l_run * alpha + ((sum0 + sum1) + (sum2 + sum3));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
        ++gt;
      };

      // single-pass tile (flash_attn2_kernel's tile1p) with PolyMask<POLY16> of the exponentials on the FMA pipe
      auto tile1p = [&](int j, auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        mbar_wait(s_full(g), gt & 1u);
        tc_fence_after();
        uint32_t va[32], vb[32];
        tmem_ld_x32(tmem_S + lane_off, va);
        tmem_ld_x32(tmem_S + lane_off + 32, vb);
        float alpha = 1.0f;
        if (j >= 2) {
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(xs_mine + (((j - 1) & 1) << 9)),
                       "h"(static_cast<unsigned short>(__float_as_uint(sm1) >> 16)) : "memory");
          if (j >= 3) {
            unsigned short peer_bits;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(peer_bits) : "r"(xs_peer + ((j & 1) << 9)) : "memory");
            const float peer = __uint_as_float(static_cast<uint32_t>(peer_bits) << 16);
            const float est = rf2 + ATT2_BIAS + __log2f(sm2 + peer);
            if (__any_sync(0xffffffffu, est - mrs > 8.0f)) {
              const float m_new = fmaxf(mrs, est);
              alpha = ex2(mrs - m_new);
              mrs = m_new;
              if (sub == 0) {
                mbar_wait(pv_done(g, (gt - 1u) & 1u), ((gt - 1u) >> 1) & 1u);
                tc_fence_after();
                rescale_o_rows(tmem_O + lane_off, p.dv, alpha);
              }
            }
          }
        }
        const float mneg = -mrs - ATT2_BIAS;
        const f32x2 mneg_2 = pk2(mneg, mneg);
        f32x2 acc0 = pk2(0.0f, 0.0f), acc1 = acc0;
        auto half = [&](uint32_t (&v)[32], int h) {
          if (MASKED) {
            const int kvalid = p.N - j * KT - sub * 64 - h * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kvalid) v[i] = 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const f32x2 x = fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sl2_2, mneg_2);
            if ((PolyMask<POLY16>::value >> ((i >> 1) & 7)) & 1u) {
              exp2_fma2(x, v[i], v[i + 1]);
            } else {
              float x0, x1;
              upk2(x, x0, x1);
              v[i] = __float_as_uint(ex2(x0));
              v[i + 1] = __float_as_uint(ex2(x1));
            }
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            acc0 = add2(acc0, pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            acc1 = add2(acc1, pk2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
          }
          store_p(v, h);
        };
        // P is single-buffered: P.V of the previous tile (issued a whole tile of exponentials ago) must have read it.  Waited
        // for HERE, in front of both halves, so that the two halves form one straight-line block and the scheduler can run
        // the FMA-pipe work of one under the MUFU work of the other.
        mbar_wait(pv_done(g, (gt - 1u) & 1u), ((gt - 1u) >> 1) & 1u);
        tmem_ld_wait_x32x2(va, vb);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(g));
        half(va, 0);
        half(vb, 1);
        float t0, t1, t2, t3;
        upk2(acc0, t0, t1);
        upk2(acc1, t2, t3);
        const float tsum = (t0 + t1) + (t2 + t3);
        l_run = l_run * alpha + tsum;
        sm2 = sm1; rf2 = rf1;
        sm1 = __bfloat162float(__float2bfloat16_rn(tsum)); rf1 = mrs;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
        ++gt;
      };

      if (exact) {
        for (int j = 0; j < T_full; ++j) tile(j, std::false_type{});
        if (T_full < T) tile(T_full, std::true_type{});
      } else {
        tile(0, std::false_type{});          // N > 128 here: the first tile is always full
        mrs = m_run * sl2;
        for (int j = 1; j < T_full; ++j) tile1p(j, std::false_type{});
        if (T_full < T) tile1p(T_full, std::true_type{});
      }

      // ---- row sums of the two halves, overflow check, O / l -> out ----
      *xfmine = l_run;
      named_bar_sync(bar_id, 256);
      const float l_tot = l_run + *xfpeer;
      mbar_wait(pv_done(g, (gt - 1u) & 1u), ((gt - 1u) >> 1) & 1u);
      tc_fence_after();
      const int tok = q0 + g * QT + row;
      if (!exact && p.flags != nullptr) {
        const bool bad = (tok < p.N) && !(l_tot < ATT_L_SAFE);       // also true for NaN
        if (__any_sync(0xffffffffu, bad) && lane == 0) {
          *reinterpret_cast<volatile int*>(p.flags + item) = 1;
          *cta_flagged = 1;
        }
      }
      if (sub == 0) {
        const float inv_l = 1.0f / l_tot;
        bf16* orow = p.out + (static_cast<long long>(b) * p.N + tok) * p.C + head * p.d;
        for (int c = 0; c < p.dv; c += 16) {
          uint32_t o[16];
          tmem_ld_x16(tmem_O + lane_off + c, o);
          tmem_ld_wait();
          if (tok < p.N) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              if (c + i < p.d) {
                uint4 pk;
                pk.x = pack_op2(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l, p.out_f16);
                pk.y = pack_op2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l, p.out_f16);
                pk.z = pack_op2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l, p.out_f16);
                pk.w = pack_op2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l, p.out_f16);
                *reinterpret_cast<uint4*>(orow + c + i) = pk;
              }
            }
          }
        }
        tc_fence_before();   // these reads of O precede the next item's first P.V (ordered through p_full)
      }
      // the final sums are read; the next item's tile 0 reuses the max-exchange slots only after its own barrier
      named_bar_sync(bar_id, 256);
    }
    if (pass == 1 || !second_pass()) break;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (*cta_flagged != 0) {   // the flags this CTA raised are consumed: clean for the next launch
    for (int item = static_cast<int>(blockIdx.x) + static_cast<int>(threadIdx.x) * step; item < total; item += step * ATT2_THREADS)
      p.flags[item] = 0;
  }
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

constexpr size_t ATT2_SMEM = 2 * CHUNK_BYTES + ATT2_KV_STAGES * CHUNK_BYTES + ATT2_KV_STAGES * (2 * 64 * 128) +
                             2048 + 8 * (12 + 2 * ATT2_KV_STAGES);

void fill_params(const AttnPlan& plan, AttnParams* p) {
  p->N = plan.N; p->heads = plan.heads; p->d = plan.d;
  p->dv = (plan.d + 15) / 16 * 16;
  p->ksteps = (plan.d + 15) / 16;
  p->C = plan.heads * plan.d;
  p->scale_log2 = plan.scale_log2;
  p->out = plan.out;
  p->flags = plan.flags;
  p->out_f16 = plan.out_f16;
  p->two_pass = 0;
  p->qtiles = (plan.N + 2 * QT - 1) / (2 * QT);
  p->total_items = p->qtiles * plan.heads * plan.B;
}

// PBE_ATTN_TWO_PASS=1: exact running maximum in every tile (no single-pass tiles, no re-run); PBE_ATTN_KERNEL=2: the
// round-1 non-persistent kernel for head dims <= 64 (A/B comparisons); PBE_ATTN_POLY=0|2|4|6|8: exponentials per 16 on the
// FMA pipe (default 6); PBE_ATTN_RERUN=0: skip the exact re-run launch (overflowing rows then stay NaN, as in round 1).
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
bool attn_two_pass() { static const int v = env_int("PBE_ATTN_TWO_PASS", 0); return v != 0; }
bool attn_rerun() { static const int v = env_int("PBE_ATTN_RERUN", 1); return v != 0; }

int launch_attn2(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  static_assert(ATT2_SMEM <= 227 * 1024, "attention smem");
  if (!attr_set) {
    PBE_CHECK_CUDA(cudaFuncSetAttribute(flash_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(ATT2_SMEM)));
    attr_set = true;
  }
  AttnParams p;
  fill_params(plan, &p);
  p.flags = nullptr;
  p.two_pass = attn_two_pass() ? 1 : 0;
  dim3 grid((plan.N + 2 * QT - 1) / (2 * QT), plan.heads, plan.B);
  PBE_CHECK_CUDA(launch_k(flash_attn2_kernel, dim3(grid), dim3(ATT2_THREADS), ATT2_SMEM, stream, plan.tmQ, plan.tmV, p));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attn_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int POLY16>
int launch_attn3_t(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  static_assert(ATT3_SMEM <= 227 * 1024, "attention smem");
  if (!attr_set) {
    PBE_CHECK_CUDA(cudaFuncSetAttribute(flash_attn3_kernel<POLY16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(ATT3_SMEM)));
    attr_set = true;
  }
  AttnParams p;
  fill_params(plan, &p);
  const dim3 grid(static_cast<unsigned>(std::min(p.total_items, attn_num_sms())));
  if (attn_two_pass()) {   // exact everywhere: one launch, nothing to re-run
    p.two_pass = 1;
    p.flags = nullptr;
    PBE_CHECK_CUDA(launch_k(flash_attn3_kernel<POLY16>, grid, dim3(ATT2_THREADS), ATT3_SMEM, stream, plan.tmQ, plan.tmV, p));
    return 0;
  }
  if (!attn_rerun()) p.flags = nullptr;   // no exact second pass: an overflowing row stays NaN (round-1 behaviour)
  PBE_CHECK_CUDA(launch_k(flash_attn3_kernel<POLY16>, grid, dim3(ATT2_THREADS), ATT3_SMEM, stream, plan.tmQ, plan.tmV, p));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_attn3(const AttnPlan& plan, cudaStream_t stream) {
  static const int poly = env_int("PBE_ATTN_POLY", 6);
  switch (poly) {
    case 0: return launch_attn3_t<0>(plan, stream);
    case 2: return launch_attn3_t<2>(plan, stream);
    case 4: return launch_attn3_t<4>(plan, stream);
    case 8: return launch_attn3_t<8>(plan, stream);
    default: return launch_attn3_t<6>(plan, stream);
  }
}

template <int DK_CHUNKS, int KV_STAGES, int VROWS>
constexpr size_t attn_smem_bytes() {
  return 1024 + DK_CHUNKS * CHUNK_BYTES + KV_STAGES * DK_CHUNKS * CHUNK_BYTES + KV_STAGES * 2 * VROWS * 128 +
         8 * (8 + 4 * KV_STAGES) + 4096;
}

template <int DK_CHUNKS, int KV_STAGES, int VROWS>
int launch_attn_t(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  constexpr size_t smem = attn_smem_bytes<DK_CHUNKS, KV_STAGES, VROWS>();
  static_assert(smem <= 227 * 1024, "attention smem");
  if (!attr_set) {
    PBE_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<DK_CHUNKS, KV_STAGES, VROWS>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  AttnParams p;
  fill_params(plan, &p);
  PBE_REQUIRE(p.dv <= VROWS, "value rows exceed the V stage of this instantiation");
  if (attn_two_pass()) { p.two_pass = 1; p.flags = nullptr; }
  else if (!attn_rerun()) p.flags = nullptr;
  PBE_CHECK_CUDA(launch_k(flash_attn_kernel<DK_CHUNKS, KV_STAGES, VROWS>, dim3(plan.grid), dim3(ATT_THREADS), smem, stream, plan.tmQ, plan.tmV, p));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Overflow flags: one int per work item, zero between launches (the exact re-run clears what it consumed).  Plans built by
// an engine get their own slice of a per-device pool; one-off plans (pbe_op_self_attention) share the first slice, which is
// safe for launches on one stream.
constexpr int FLAG_POOL_INTS = 1 << 20, FLAG_SHARED_INTS = 1 << 16;
int* attn_flags(int n, bool own) {
  static std::mutex mu;
  static std::map<int, std::pair<int*, int>> pools;   // device -> (base, next free)
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto it = pools.find(dev);
  if (it == pools.end()) {
    int* base = nullptr;
    if (cudaMalloc(&base, sizeof(int) * FLAG_POOL_INTS) != cudaSuccess) return nullptr;
    cudaMemset(base, 0, sizeof(int) * FLAG_POOL_INTS);
    it = pools.emplace(dev, std::make_pair(base, FLAG_SHARED_INTS)).first;
  }
  if (n > FLAG_SHARED_INTS) return nullptr;           // no flags: such a plan runs without the re-run
  if (!own || it->second.second + n > FLAG_POOL_INTS) return it->second.first;
  int* r = it->second.first + it->second.second;
  it->second.second += n;
  return r;
}

}  // namespace

int build_attn_plan(const bf16* qk, const bf16* vt, bf16* out, int B, int N, int heads, int d, AttnPlan* plan, bool own_flags) {
  PBE_REQUIRE(d % 8 == 0 && d <= 160, "head dim must be a multiple of 8, <= 160");
  const int C = heads * d;
  plan->B = B; plan->N = N; plan->heads = heads; plan->d = d;
  plan->flags = attn_flags(((N + QT - 1) / QT) * heads * B, own_flags);
  plan->out_f16 = operand_f16();
  plan->scale_log2 = static_cast<float>(1.4426950408889634 / sqrt(static_cast<double>(d)));
  plan->out = out;
  plan->grid = dim3((N + QT - 1) / QT, heads, B);
  const int dv = (d + 15) / 16 * 16;
  {
    // Q|K: (d, 2*heads, N, B) over [B, N, 2C]
    const uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(2 * heads), static_cast<uint64_t>(N),
                              static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(2 * C) * 2,
                                 static_cast<uint64_t>(N) * 2 * C * 2};
    const uint32_t box[4] = {64u, 1u, 128u, 1u};
    int rc = make_tmap_bf16(&plan->tmQ, qk, 4, dims, strides, box, true);
    if (rc) return rc;
    const uint32_t box64[4] = {64u, 1u, 64u, 1u};
    rc = make_tmap_bf16(&plan->tmK, qk, 4, dims, strides, box64, true);
    if (rc) return rc;
  }
  {
    // V^T: (N, C, B) over [B, C, vt_pitch(N)]
    const uint64_t Np = static_cast<uint64_t>(vt_pitch(N));
    const uint64_t dims[3] = {static_cast<uint64_t>(N), static_cast<uint64_t>(C), static_cast<uint64_t>(B)};
    const uint64_t strides[2] = {Np * 2, static_cast<uint64_t>(C) * Np * 2};
    const uint32_t box[3] = {64u, static_cast<uint32_t>(dv), 1u};
    int rc = make_tmap_bf16(&plan->tmV, vt, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  plan->smem = 0;
  return 0;
}

int attn_num_launches(const AttnPlan& plan) {
  (void)plan;
  return 1;   // the exact pass over overflowed items runs inside the same launch
}

int launch_attn_plan(const AttnPlan& plan, cudaStream_t stream) {
  const int chunks = (plan.d + 63) / 64;
  static const int which = env_int("PBE_ATTN_KERNEL", 3);
  if (chunks == 1 && plan.N > QT) return which == 2 ? launch_attn2(plan, stream) : launch_attn3(plan, stream);
  switch (chunks) {
    case 1: return launch_attn_t<1, 3, 64>(plan, stream);
    case 2: return plan.d <= 96 ? launch_attn_t<2, 3, 96>(plan, stream) : launch_attn_t<2, 2, 128>(plan, stream);
    case 3: return launch_attn_t<3, 1, 160>(plan, stream);
    default: set_error("launch_attn_plan: head dim too large"); return -1;
  }
}

}  // namespace pbe
