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
// * warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM owner, warps 2..5 = epilogue
//   (tcgen05.ld -> registers -> smem transpose -> coalesced global I/O with bias / per-sample bias / residual,
//   GEGLU gate, or the Q|K / V^T split store).
//
// Replaces (reference, PyTorch library calls): conv2d in ResBlock/Downsample/Upsample
// (ldm/modules/diffusionmodules/openaimodel.py:107-119,150-160,201-241), 1x1 proj_in/proj_out and all Linear layers of
// BasicTransformerBlock (ldm/modules/attention.py:38-65,198-230,270-297).
#include "internal.h"
#include "ptx.cuh"

namespace pbe {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_THREADS = 192;
constexpr int STAGE_LD = 36;  // epilogue transpose buffer row pitch (floats); 16-B aligned rows
constexpr int STAGE_LD4 = STAGE_LD / 4;

__host__ __device__ constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// exact-erf GELU (torch.nn.functional.gelu default). erf via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below
// the bf16 resolution of the stored result): 1 MUFU.RCP + 1 MUFU.EX2 + ~10 FMA instead of libdevice erff.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = exp2f(-z * z * 1.4426950408889634f);
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float erfv = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erfv);
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ ConvGemmParams p) {
  constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(BLOCK_N);
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");
  static_assert(4 * 32 * STAGE_LD * 4 <= STAGES * A_STAGE_BYTES, "epilogue staging aliases the A stages");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t sBar = sB + STAGES * B_STAGE_BYTES;  // full[STAGES], empty[STAGES], tmem_full, tmem_ptr
  uint8_t* bar_gen = smem_gen + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = sBar + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_addr = sBar + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(bar_gen + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates ----
  int mt = blockIdx.x;
  const int tile_w = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int tile_h = mt % p.tiles_h;
  const int tile_n = mt / p.tiles_h;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = tile_n * p.tn;
  const int n_base = blockIdx.y * BLOCK_N;
  const int num_k_iters = p.num_taps * p.k_chunks;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_k_iters; ++it) {
        const int tap = it / p.k_chunks;
        const int kc = it - tap * p.k_chunks;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + B_STAGE_BYTES);
        tma_load_5d(sA + stage * A_STAGE_BYTES, &tmA, full_bar(stage), p.tap_coff[tap] + kc * BLOCK_K,
                    w0 + p.tap_dw[tap], p.tap_ph[tap], h0 + p.tap_dh[tap], n0);
        tma_load_3d(sB + stage * B_STAGE_BYTES, &tmB, full_bar(stage), kc * BLOCK_K, n_base, tap);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < num_k_iters; ++it) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = umma_desc_sw128(sA + stage * A_STAGE_BYTES);
        const uint64_t bdesc = umma_desc_sw128(sB + stage * B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the 16-B-unit start address field
          umma_bf16_ss(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));  // smem slot free once these MMAs retire
        if (it == num_k_iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ================= Epilogue (warps 2..5) =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float* stg = reinterpret_cast<float*>(smem_gen) + (warp - 2) * 32 * STAGE_LD;

    // Geometry of my row (row = q*32 + lane), exchanged by shuffle while writing coalesced rows.
    const int row = q * 32 + lane;
    const int wl = row % p.tw;
    const int hl = (row / p.tw) % p.th;
    const int nl = row / (p.tw * p.th);
    const int ow = w0 + wl, oh = h0 + hl, on = n0 + nl;
    const bool my_valid = (ow < p.Wo) && (oh < p.Ho) && (on < p.Nb);
    const long long my_m = (static_cast<long long>(on) * p.Ho + oh) * p.Wo + ow;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    // Coalesced output pass shared by all modes: the 32x32 fp32 block staged in shared memory (thread = row) is
    // re-read with lane = (row quad, 4 consecutive columns): every instruction touches 4 rows x 128 contiguous bytes.
    const int rq = lane >> 3;           // row within a group of 4
    const int cq = (lane & 7) * 4;      // first of my 4 columns inside the 32-column chunk
    float4* stg4 = reinterpret_cast<float4*>(stg);

    if (p.mode == EPI_GEGLU) {
      // tile-local columns [0, BLOCK_N/2) = value half, [BLOCK_N/2, BLOCK_N) = gate half (weights pre-interleaved)
      constexpr int HALF = BLOCK_N / 2;
      const int out_col0 = blockIdx.y * HALF;
#pragma unroll 1
      for (int c0 = 0; c0 < HALF; c0 += 32) {
        uint32_t va[32], vg[32];
        tmem_ld_x32(taddr + c0, va);
        tmem_ld_x32(taddr + HALF + c0, vg);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + c0 + j));
          const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + HALF + c0 + j));
          float4 o;
          o.x = (__uint_as_float(va[j + 0]) + ba.x) * gelu_erf(__uint_as_float(vg[j + 0]) + bg.x);
          o.y = (__uint_as_float(va[j + 1]) + ba.y) * gelu_erf(__uint_as_float(vg[j + 1]) + bg.y);
          o.z = (__uint_as_float(va[j + 2]) + ba.z) * gelu_erf(__uint_as_float(vg[j + 2]) + bg.z);
          o.w = (__uint_as_float(va[j + 3]) + ba.w) * gelu_erf(__uint_as_float(vg[j + 3]) + bg.w);
          stg4[lane * STAGE_LD4 + (j >> 2)] = o;
        }
        __syncwarp();
        const int col = out_col0 + c0 + cq;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rq;
          const long long m = __shfl_sync(0xffffffffu, my_m, r);
          const int vb = __shfl_sync(0xffffffffu, my_valid ? 1 : 0, r);
          if (vb) {
            const float4 x = stg4[r * STAGE_LD4 + (cq >> 2)];
            uint2 pk;
            pk.x = pack_bf16x2(x.x, x.y);
            pk.y = pack_bf16x2(x.z, x.w);
            *reinterpret_cast<uint2*>(p.out_bf16 + m * p.ld_out + col) = pk;
          }
        }
        __syncwarp();
      }
    } else {
      const float* __restrict__ residual = p.residual;
      const float* __restrict__ rowbias = p.rowbias;
      float* __restrict__ out_f32 = p.out_f32;
      bf16* __restrict__ out_bf16 = p.out_bf16;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        if (n_base + c0 >= p.n_total) break;
        uint32_t v[32];
        tmem_ld_x32(taddr + c0, v);
        tmem_ld_wait();
        if (p.mode == EPI_QKV && n_base + c0 >= p.qk_cols) {
          // V^T store: out_vt[b][c][token]; lanes are consecutive tokens -> 64-B contiguous per column
          if (my_valid) {
            const long long tokens = static_cast<long long>(p.Ho) * p.Wo;
            const long long tok = static_cast<long long>(oh) * p.Wo + ow;
            const int vC = p.n_total - p.qk_cols;
            bf16* dst = p.out_vt + (static_cast<long long>(on) * vC + (n_base + c0 - p.qk_cols)) * tokens + tok;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * tokens] = __float2bfloat16(__uint_as_float(v[j]));
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          stg4[lane * STAGE_LD4 + (j >> 2)] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                          __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        __syncwarp();
        const int col = n_base + c0 + cq;
        const bool colv = col < p.n_total;  // n_total % 4 == 0
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr && colv) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        // issue every global read of this chunk before the first dependent use (8 independent 16-B loads per lane)
        long long moff[8];
        float4 add4[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rq;
          const long long m = __shfl_sync(0xffffffffu, my_m, r);
          const int vb = __shfl_sync(0xffffffffu, my_valid ? on : -1, r);
          moff[it] = (vb >= 0 && colv) ? m * p.ld_out + col : -1;
          float4 a = bias4;
          if (moff[it] >= 0) {
            if (residual != nullptr) {
              const float4 rr = __ldg(reinterpret_cast<const float4*>(residual + moff[it]));
              a.x += rr.x; a.y += rr.y; a.z += rr.z; a.w += rr.w;
            }
            if (rowbias != nullptr) {
              const float4 rb = __ldg(reinterpret_cast<const float4*>(rowbias + static_cast<long long>(vb) * p.rowbias_ld + col));
              a.x += rb.x; a.y += rb.y; a.z += rb.z; a.w += rb.w;
            }
          }
          add4[it] = a;
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (moff[it] >= 0) {
            const int r = it * 4 + rq;
            float4 x = stg4[r * STAGE_LD4 + (cq >> 2)];
            x.x += add4[it].x; x.y += add4[it].y; x.z += add4[it].z; x.w += add4[it].w;
            if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + moff[it]) = x;
            if (out_bf16 != nullptr) {
              uint2 pk;
              pk.x = pack_bf16x2(x.x, x.y);
              pk.y = pack_bf16x2(x.z, x.w);
              *reinterpret_cast<uint2*>(out_bf16 + moff[it]) = pk;
            }
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BLOCK_N, int STAGES>
constexpr size_t smem_bytes_for() {
  return 1024 + STAGES * (A_STAGE_BYTES + BLOCK_N * BLOCK_K * 2) + 8 * (2 * STAGES + 2);
}

template <int BLOCK_N, int STAGES>
int launch_t(const GemmPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  constexpr size_t smem = smem_bytes_for<BLOCK_N, STAGES>();
  if (!attr_set) {
    PBE_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    attr_set = true;
  }
  conv_gemm_kernel<BLOCK_N, STAGES><<<plan.grid, NUM_THREADS, smem, stream>>>(plan.tmA, plan.tmB, plan.p);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

void pick_tile(int Wo, int Ho, int Nb, int* tw, int* th, int* tn) {
  auto best = [](int extent, int budget) {
    int best_t = 1;
    long best_pad = -1;
    for (int t = 1; t <= budget; t *= 2) {
      long padded = static_cast<long>((extent + t - 1) / t) * t;
      if (best_pad < 0 || padded < best_pad || (padded == best_pad && t > best_t)) {
        best_pad = padded;
        best_t = t;
      }
    }
    return best_t;
  };
  *tw = best(Wo, 128);
  *th = best(Ho, 128 / *tw);
  *tn = 128 / (*tw * *th);
  (void)Nb;
}

}  // namespace

int build_gemm_plan(const ConvGemmDesc& d, GemmPlan* plan) {
  PBE_REQUIRE(d.C % 64 == 0, "activation channels must be a multiple of 64");
  PBE_REQUIRE(d.ksize == 1 || d.ksize == 3, "kernel size 1 or 3");
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
      } else if (d.stride == 1) {
        p.tap_dw[t] = static_cast<int8_t>(kw - 1);
        p.tap_dh[t] = static_cast<int8_t>(kh - 1);
        p.tap_ph[t] = 0;
        p.tap_coff[t] = 0;
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
  p.residual = d.residual;
  p.out_f32 = d.out_f32;
  p.out_bf16 = d.out_bf16;
  p.out_vt = d.out_vt;
  p.qk_cols = d.qk_cols;
  p.ld_out = d.ld_out ? d.ld_out : (d.mode == EPI_GEGLU ? d.Cout / 2 : d.Cout);

  int bn = d.block_n;
  if (bn == 0) {
    if (d.mode == EPI_GEGLU) bn = 128;
    else if (d.Cout % 160 == 0) bn = 160;
    else if (d.Cout % 128 == 0) bn = 128;
    else if (d.Cout <= 32) bn = 32;
    else if (d.Cout <= 64) bn = 64;
    else bn = 128;
  }
  PBE_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 160 || bn == 256, "unsupported BLOCK_N");
  if (d.mode == EPI_GEGLU) PBE_REQUIRE(bn == 128 && d.Cout % 128 == 0, "GEGLU needs BLOCK_N=128 | Cout");
  if (d.mode == EPI_QKV) PBE_REQUIRE(d.qk_cols % bn == 0 && d.Cout % bn == 0, "QKV split must align with BLOCK_N");
  plan->block_n = bn;
  plan->grid = dim3(p.tiles_w * p.tiles_h * p.tiles_n, (d.Cout + bn - 1) / bn, 1);

  // A: 5-D view (c, w, phase, h, n)
  {
    const uint64_t C = d.C, W = d.W, H = d.H, N = d.Nb;
    uint64_t dims[5], strides[4];
    if (d.stride == 1) {
      dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
      strides[0] = C * 2; strides[1] = W * C * 2; strides[2] = W * C * 2; strides[3] = H * W * C * 2;
    } else {
      dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
      strides[0] = 2 * C * 2; strides[1] = W * C * 2; strides[2] = 2 * W * C * 2; strides[3] = H * W * C * 2;
    }
    const uint32_t box[5] = {64u, static_cast<uint32_t>(p.tw), 1u, static_cast<uint32_t>(p.th),
                             static_cast<uint32_t>(p.tn)};
    int rc = make_tmap_bf16(&plan->tmA, d.act, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  // B: 3-D (cin, cout, tap)
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(d.C), static_cast<uint64_t>(d.Cout),
                              static_cast<uint64_t>(p.num_taps)};
    const uint64_t strides[2] = {static_cast<uint64_t>(d.C) * 2, static_cast<uint64_t>(d.C) * d.Cout * 2};
    const uint32_t box[3] = {64u, static_cast<uint32_t>(bn), 1u};
    int rc = make_tmap_bf16(&plan->tmB, d.wt, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  return 0;
}

int launch_gemm_plan(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.block_n) {
    case 32: return launch_t<32, 4>(plan, stream);
    case 64: return launch_t<64, 4>(plan, stream);
    case 128: return launch_t<128, 3>(plan, stream);
    case 160: return launch_t<160, 3>(plan, stream);
    case 256: return launch_t<256, 4>(plan, stream);
    default: set_error("launch_gemm_plan: bad block_n"); return -1;
  }
}

}  // namespace pbe
