// Layout / elementwise helpers and the small CUDA-core dense layers of the timestep-embedding path.
// All are HBM- or latency-bound; none is GEMM-shaped enough for the tensor cores (M = batch rows only).
#include "internal.h"
#include "ptx.cuh"

namespace pbe {

namespace {

__global__ void upsample2x_kernel(const float* __restrict__ x, bf16* __restrict__ y, int Nb, int H, int W, int C, int f16) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  // one thread per input float4
  const long long total = static_cast<long long>(Nb) * H * W * (C / 4);
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = static_cast<int>(idx % (C / 4));
  long long pix = idx / (C / 4);
  const int w = static_cast<int>(pix % W);
  pix /= W;
  const int h = static_cast<int>(pix % H);
  const int n = static_cast<int>(pix / H);
  const float4 v = *reinterpret_cast<const float4*>(x + idx * 4);
  uint2 pk;
  pk.x = pack_op2(v.x, v.y, f16);
  pk.y = pack_op2(v.z, v.w, f16);
  const int W2 = 2 * W, H2 = 2 * H;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long o = ((static_cast<long long>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dx) * C + cv * 4;
      *reinterpret_cast<uint2*>(y + o) = pk;
    }
}

// nearest 2x from a 16-bit source: a pure copy, one thread per 8 channels (16 bytes) of an input pixel
__global__ void upsample2x_16_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int Nb, int H, int W, int C) {
  griddep_enter();
  const long long total = static_cast<long long>(Nb) * H * W * (C / 8);
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = static_cast<int>(idx % (C / 8));
  long long pix = idx / (C / 8);
  const int w = static_cast<int>(pix % W);
  pix /= W;
  const int h = static_cast<int>(pix % H);
  const int n = static_cast<int>(pix / H);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + idx);
  const int W2 = 2 * W, H2 = 2 * H;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long o = ((static_cast<long long>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dx) * C + cv * 8;
      *reinterpret_cast<uint4*>(y + o) = v;
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n4, int f16) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  uint2 pk;
  pk.x = pack_op2(v.x, v.y, f16);
  pk.y = pack_op2(v.z, v.w, f16);
  reinterpret_cast<uint2*>(y)[i] = pk;
}

// thread per (n, pixel): gathers Cin channel planes (coalesced across threads), writes Cpad bf16 contiguous
__global__ void pack_input_kernel(const float* __restrict__ x, bf16* __restrict__ y, int Nb, int Cin, int HW,
                                  int Cpad, int f16) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
  bf16* dst = y + idx * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      v[i] = (c < Cin) ? x[(static_cast<long long>(n) * Cin + c) * HW + pix] : 0.0f;
    }
    uint4 pk;
    pk.x = pack_op2(v[0], v[1], f16);
    pk.y = pack_op2(v[2], v[3], f16);
    pk.z = pack_op2(v[4], v[5], f16);
    pk.w = pack_op2(v[6], v[7], f16);
    *reinterpret_cast<uint4*>(dst + c0) = pk;
  }
}

__global__ void unpack_output_kernel(const float* __restrict__ y, float* __restrict__ out, int Nb, int Cout, int HW,
                                     int ld) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
  for (int c = 0; c < Cout; ++c) out[(static_cast<long long>(n) * Cout + c) * HW + pix] = y[idx * ld + c];
}

__device__ __forceinline__ float silu_f(float t) { return t / (1.0f + expf(-t)); }

// y[b][o] = act(bias[o] + sum_k W[o][k] x[b][k]) for a handful of batch rows (timestep / context embeddings, the CLIP
// mapper: fp32).  One CTA of SL_WARPS warps owns SL_OW outputs and SL_RB batch rows at a time; the warps take interleaved
// 128-float slices of K, lanes the float4s of a slice, so a CTA has all of its weight loads in flight at once and there
// are O / SL_OW CTAs (the previous version gave a WARP the whole K loop -- 10 dependent round trips to HBM at K = 1280 --
// and filled 40 SMs: 43 us for the 6.5 MB of time_embed.2).  Every x float4 is loaded once per warp and reused by the
// SL_OW weight rows.  The SL_OW * SL_RB partial sums of a warp are combined by a halving butterfly (62 shuffles instead of
// 5 per value), the warps' results through shared memory in warp order: a fixed summation order, no atomics.
constexpr int SL_OW = 4, SL_RB = 16, SL_WARPS = 8;
__global__ void __launch_bounds__(SL_WARPS * 32) small_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           float* __restrict__ y_silu, int B, int K, int O,
                                                           int pre_silu, int post_act, const float* __restrict__ residual) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  __shared__ float s_part[SL_WARPS][SL_OW * SL_RB];
  const int o0 = blockIdx.x * SL_OW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < B; b0 += SL_RB) {
    float acc[SL_OW * SL_RB];
#pragma unroll
    for (int n = 0; n < SL_OW * SL_RB; ++n) acc[n] = 0.0f;
    for (int k = (warp * 32 + lane) * 4; k < K; k += SL_WARPS * 128) {
      float4 xv[SL_RB];
#pragma unroll
      for (int i = 0; i < SL_RB; ++i) {
        xv[i] = (b0 + i < B) ? *reinterpret_cast<const float4*>(x + static_cast<long long>(b0 + i) * K + k)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        if (pre_silu) { xv[i].x = silu_f(xv[i].x); xv[i].y = silu_f(xv[i].y); xv[i].z = silu_f(xv[i].z); xv[i].w = silu_f(xv[i].w); }
      }
#pragma unroll
      for (int j = 0; j < SL_OW; ++j) {
        const float4 w = (o0 + j < O) ? __ldg(reinterpret_cast<const float4*>(W + static_cast<long long>(o0 + j) * K + k))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < SL_RB; ++i)
          acc[j * SL_RB + i] += (w.x * xv[i].x + w.y * xv[i].y) + (w.z * xv[i].z + w.w * xv[i].w);
      }
    }
    // halving butterfly: after the step with offset `off` a lane keeps the half of the values selected by that lane bit
#pragma unroll
    for (int off = 16, cnt = SL_OW * SL_RB; off > 0; off >>= 1, cnt >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int n = 0; n < cnt / 2; ++n) {
        const float send = up ? acc[n] : acc[n + cnt / 2];
        const float keep = up ? acc[n + cnt / 2] : acc[n];
        acc[n] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    // lane bits 4..0 picked halves of 64, 32, ..., 4 values: this lane holds n = 2 * lane + {0, 1}, n = j * SL_RB + i
    if (b0 > 0) __syncthreads();   // the previous pass has been read
    s_part[warp][2 * lane] = acc[0];
    s_part[warp][2 * lane + 1] = acc[1];
    __syncthreads();
    if (threadIdx.x < SL_OW * SL_RB) {
      const int n = threadIdx.x;
      const int j = n / SL_RB, i = n % SL_RB;
      if (o0 + j < O && b0 + i < B) {
        float v = s_part[0][n];
#pragma unroll
        for (int w = 1; w < SL_WARPS; ++w) v += s_part[w][n];
        v += bias ? bias[o0 + j] : 0.0f;
        if (post_act == 1) v = silu_f(v);
        else if (post_act == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));   // nn.GELU() (erf)
        if (residual != nullptr) v += residual[static_cast<long long>(b0 + i) * O + o0 + j];
        y[static_cast<long long>(b0 + i) * O + o0 + j] = v;
        if (y_silu) y_silu[static_cast<long long>(b0 + i) * O + o0 + j] = silu_f(v);
      }
    }
  }
}

// reference: timestep_embedding, ldm/modules/diffusionmodules/util.py:151-171  (cos || sin, fp32)
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ out, int B, int dim) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= B * half) return;
  const int b = idx / half, k = idx % half;
  // torch: exp(-log(1e4) * arange(half, fp32) / half) evaluated in fp32
  const float neg_log = -9.210340371976184f;
  const float f = expf((neg_log * static_cast<float>(k)) / static_cast<float>(half));
  const float a = static_cast<float>(t[b]) * f;
  out[b * dim + k] = cosf(a);
  out[b * dim + half + k] = sinf(a);
  if ((dim & 1) && k == 0) out[b * dim + dim - 1] = 0.0f;
}

__global__ void add_rowvec_kernel(const float* __restrict__ a, const float* __restrict__ v, float* __restrict__ y,
                                  int B, int N) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  y[idx] = a[idx % N] + v[idx];
}

// ---- VAE decode helpers (AutoencoderKL.decode, ldm/models/autoencoder.py:66-69; AttnBlock model.py:152-182) ----
// post_quant_conv (1x1, embed_dim -> z_channels, fp32) fused with the NCHW -> NHWC bf16 pack (channels padded with zeros)
__global__ void vae_pack_input_kernel(const float* __restrict__ z, const float* __restrict__ Wp, const float* __restrict__ bp,
                                      bf16* __restrict__ y, int Nb, int Cin, int Cz, int HW, int Cpad, int f16) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
  float zin[8];
  for (int i = 0; i < Cin; ++i) zin[i] = z[(static_cast<long long>(n) * Cin + i) * HW + pix];
  bf16* dst = y + idx * Cpad;
  for (int o = 0; o < Cpad; ++o) {
    float v = 0.0f;
    if (o < Cz) {
      v = bp[o];
      for (int i = 0; i < Cin; ++i) v += Wp[o * Cin + i] * zin[i];
    }
    reinterpret_cast<unsigned short*>(dst)[o] = to_op16(v, f16);
  }
}

// quant_conv (1x1, Cin -> Cout, fp32; autoencoder.py:36,60) fused with the NHWC -> NCHW unpack of the encoder output
__global__ void vae_unpack_moments_kernel(const float* __restrict__ y, const float* __restrict__ Wq,
                                          const float* __restrict__ bq, float* __restrict__ out, int Nb, int Cin, int Cout,
                                          int HW, int ld) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
  float hin[16];
  for (int i = 0; i < Cin; ++i) hin[i] = y[idx * ld + i];
  for (int o = 0; o < Cout; ++o) {
    float v = bq[o];
    for (int i = 0; i < Cin; ++i) v += Wq[o * Cin + i] * hin[i];
    out[(static_cast<long long>(n) * Cout + o) * HW + pix] = v;
  }
}

// Post-processing of decoded images (scripts/inference.py:346-348,379-380): u8[b,h,w,c] = trunc(255 * clamp((x+1)/2, 0, 1))
// from fp32 NCHW in one pass (the reference does clamp on the device, permute + scale + astype(uint8) in numpy).
__global__ void postprocess_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int Nb, int C, int HW) {
  griddep_enter();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
  for (int c = 0; c < C; ++c) {
    float v = __fdiv_rn(__fadd_rn(x[(static_cast<long long>(n) * C + c) * HW + pix], 1.0f), 2.0f);
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    out[idx * C + c] = static_cast<uint8_t>(__fmul_rn(255.0f, v));
  }
}

// ---- pre-processing of an edit request (third "next" row, device part) ----
// ToTensor + Normalize (scripts/inference.py:106-124 get_tensor / get_tensor_clip; ldm/data/test_bench_dataset.py:37-61):
// out[b,c,h,w] = (u8[b,h,w,c] / 255 - mean[c]) / std[c], the reference's fp32 op sequence (true divisions, no FMA).
struct Norm3 { float mean[3], std[3]; };
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int Nb, int HW, Norm3 nm) {
  griddep_enter();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn(static_cast<float>(in[idx * 3 + c]), 255.0f);
    out[(static_cast<long long>(n) * 3 + c) * HW + pix] = __fdiv_rn(__fsub_rn(v, nm.mean[c]), nm.std[c]);
  }
}

// Mask and masked image (scripts/inference.py:311-318; ldm/data/test_bench_dataset.py:89-98):
// mask = 1 - m/255 (scripts: then 0 / 1 at the 0.5 threshold), image = (u8/255 - 0.5)/0.5, inpaint = image * mask.
__global__ void prepare_inpaint_u8_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                          float* __restrict__ image_out, float* __restrict__ mask_out,
                                          float* __restrict__ inpaint_out, int Nb, int HW, int binarize) {
  griddep_enter();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
  float m = __fsub_rn(1.0f, __fdiv_rn(static_cast<float>(mask[idx]), 255.0f));
  if (binarize) m = (m < 0.5f) ? 0.0f : 1.0f;
  if (mask_out != nullptr) mask_out[idx] = m;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(img[idx * 3 + c]), 255.0f), 0.5f), 0.5f);
    const long long o = (static_cast<long long>(n) * 3 + c) * HW + pix;
    if (image_out != nullptr) image_out[o] = v;
    inpaint_out[o] = __fmul_rn(v, m);
  }
}

// torchvision Resize([h, w]) of a float tensor = F.interpolate(mode="bilinear", align_corners=False) (scripts/inference.py:332,
// the latent-resolution mask).  ATen upsample_bilinear2d: src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out.
__global__ void resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int NC, int H, int W, int h, int w) {
  griddep_enter();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(NC) * h * w) return;
  const int ox = static_cast<int>(idx % w), oy = static_cast<int>((idx / w) % h);
  const long long nc = idx / (static_cast<long long>(w) * h);
  const float sy = static_cast<float>(H) / static_cast<float>(h), sx = static_cast<float>(W) / static_cast<float>(w);
  float fy = __fsub_rn(__fmul_rn(sy, __fadd_rn(static_cast<float>(oy), 0.5f)), 0.5f);
  float fx = __fsub_rn(__fmul_rn(sx, __fadd_rn(static_cast<float>(ox), 0.5f)), 0.5f);
  fy = fy < 0.0f ? 0.0f : fy;
  fx = fx < 0.0f ? 0.0f : fx;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly1 = __fsub_rn(fy, static_cast<float>(y0)), lx1 = __fsub_rn(fx, static_cast<float>(x0));
  const float ly0 = __fsub_rn(1.0f, ly1), lx0 = __fsub_rn(1.0f, lx1);
  const float* src = in + nc * H * W;
  const float top = __fadd_rn(__fmul_rn(lx0, src[static_cast<long long>(y0) * W + x0]), __fmul_rn(lx1, src[static_cast<long long>(y0) * W + x1]));
  const float bot = __fadd_rn(__fmul_rn(lx0, src[static_cast<long long>(y1) * W + x0]), __fmul_rn(lx1, src[static_cast<long long>(y1) * W + x1]));
  out[idx] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// The same with antialias=True (the default of torchvision >= 0.17 for tensors): ATen _upsample_bilinear2d_aa, a separable
// triangle filter of support max(scale, 1) whose weights are normalised per output index; width pass, then height pass.
__device__ __forceinline__ void aa_window(int i, int in_size, float scale, int& lo, int& size, float& center, float& invscale,
                                          float& total) {
  const float support = scale >= 1.0f ? scale : 1.0f;
  center = static_cast<float>(static_cast<double>(scale) * (static_cast<double>(i) + 0.5));
  invscale = scale >= 1.0f ? static_cast<float>(1.0 / static_cast<double>(scale)) : 1.0f;
  lo = max(static_cast<int>(static_cast<long long>(static_cast<double>(__fsub_rn(center, support)) + 0.5)), 0);
  size = min(static_cast<int>(static_cast<long long>(static_cast<double>(__fadd_rn(center, support)) + 0.5)), in_size) - lo;
  total = 0.0f;
  for (int j = 0; j < size; ++j) {
    float x = static_cast<float>((static_cast<double>(__fsub_rn(static_cast<float>(j + lo), center)) + 0.5) * static_cast<double>(invscale));
    x = fabsf(x);
    total = __fadd_rn(total, x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f);
  }
}
__device__ __forceinline__ float aa_weight(int j, int lo, float center, float invscale, float total) {
  float x = static_cast<float>((static_cast<double>(__fsub_rn(static_cast<float>(j + lo), center)) + 0.5) * static_cast<double>(invscale));
  x = fabsf(x);
  const float wgt = x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f;
  return total != 0.0f ? __fdiv_rn(wgt, total) : wgt;
}
__global__ void resize_bilinear_aa_kernel(const float* __restrict__ in, float* __restrict__ out, int NC, int H, int W, int h, int w) {
  griddep_enter();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(NC) * h * w) return;
  const int ox = static_cast<int>(idx % w), oy = static_cast<int>((idx / w) % h);
  const long long nc = idx / (static_cast<long long>(w) * h);
  const float sy = static_cast<float>(H) / static_cast<float>(h), sx = static_cast<float>(W) / static_cast<float>(w);
  int xlo, xn, ylo, yn;
  float xc, xi, xt, yc, yi, yt;
  aa_window(ox, W, sx, xlo, xn, xc, xi, xt);
  aa_window(oy, H, sy, ylo, yn, yc, yi, yt);
  const float* src = in + nc * H * W;
  float acc = 0.0f;
  for (int jy = 0; jy < yn; ++jy) {
    const float* row = src + static_cast<long long>(ylo + jy) * W + xlo;
    float t = __fmul_rn(row[0], aa_weight(0, xlo, xc, xi, xt));                       // width pass of this source row
    for (int jx = 1; jx < xn; ++jx) t = __fadd_rn(t, __fmul_rn(row[jx], aa_weight(jx, xlo, xc, xi, xt)));
    const float v = __fmul_rn(t, aa_weight(jy, ylo, yc, yi, yt));                     // height pass
    acc = jy == 0 ? v : __fadd_rn(acc, v);
  }
  out[idx] = acc;
}

// CLIP ViT patch embedding input (transformers CLIPVisionEmbeddings.patch_embedding: Conv2d(3, C, patch, stride=patch,
// bias=False)): one row per patch, columns in the conv weight's (c, kh, kw) order, zero padded to Kpad
__global__ void clip_pack_patches_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int H, int W, int patch,
                                         int Kpad, int f16) {
  griddep_enter();
  const int pw = W / patch, ph = H / patch;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * ph * pw * Kpad;
  if (idx >= total) return;
  const int k = static_cast<int>(idx % Kpad);
  const long long row = idx / Kpad;
  const int px = static_cast<int>(row % pw);
  const int py = static_cast<int>((row / pw) % ph);
  const int b = static_cast<int>(row / (static_cast<long long>(pw) * ph));
  float v = 0.0f;
  if (k < 3 * patch * patch) {
    const int c = k / (patch * patch), r = k % (patch * patch), kh = r / patch, kw = r % patch;
    v = img[((static_cast<long long>(b) * 3 + c) * H + py * patch + kh) * W + px * patch + kw];
  }
  reinterpret_cast<unsigned short*>(out)[idx] = to_op16(v, f16);
}

// tokens[b, 0] = class_embedding + pos[0]; tokens[b, 1 + p] = patch_emb[b, p] + pos[1 + p]  (CLIPVisionEmbeddings.forward)
__global__ void clip_embed_kernel(const float* __restrict__ patch_emb, const float* __restrict__ cls,
                                  const float* __restrict__ pos, float* __restrict__ tokens, int B, int P, int C) {
  griddep_enter();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * (P + 1) * C;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const int t = static_cast<int>((idx / C) % (P + 1));
  const int b = static_cast<int>(idx / (static_cast<long long>(C) * (P + 1)));
  const float base = t == 0 ? cls[c] : patch_emb[(static_cast<long long>(b) * P + (t - 1)) * C + c];
  tokens[idx] = base + pos[static_cast<long long>(t) * C + c];
}

// P[r][:] = softmax(S[r][:] * scale), fp32 in, bf16 out; one 256-thread block per row, the row lives in registers
constexpr int SM_MAXV = 16;   // float4 per thread -> rows of up to 16384
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, bf16* __restrict__ P, int N,
                                                           float scale_log2, int f16) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  __shared__ float s_red[8];
  const long long row = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(S + row * N);
  const int nv = N / 4;
  float4 v[SM_MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int k = threadIdx.x + i * 256;
    if (k < nv) {
      v[i] = src[k];
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
  auto block_reduce = [&](float val, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, val, o);
      val = is_max ? fmaxf(val, other) : val + other;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = val;
    __syncthreads();
    float r = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, s_red[i]) : r + s_red[i];
    return r;
  };
  mx = block_reduce(mx, true);
  const float off = -mx * scale_log2;
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int k = threadIdx.x + i * 256;
    if (k < nv) {
      v[i].x = exp2f(fmaf(v[i].x, scale_log2, off)); v[i].y = exp2f(fmaf(v[i].y, scale_log2, off));
      v[i].z = exp2f(fmaf(v[i].z, scale_log2, off)); v[i].w = exp2f(fmaf(v[i].w, scale_log2, off));
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  sum = block_reduce(sum, false);
  const float inv = 1.0f / sum;
  uint2* dst = reinterpret_cast<uint2*>(P + row * N);
#pragma unroll
  for (int i = 0; i < SM_MAXV; ++i) {
    const int k = threadIdx.x + i * 256;
    if (k < nv) {
      dst[k] = make_uint2(pack_op2(v[i].x * inv, v[i].y * inv, f16), pack_op2(v[i].z * inv, v[i].w * inv, f16));
    }
  }
}

}  // namespace

int launch_upsample2x_bf16(const float* x, bf16* y, int Nb, int H, int W, int C, cudaStream_t stream) {
  PBE_REQUIRE(C % 4 == 0, "upsample channels % 4");
  const long long total = static_cast<long long>(Nb) * H * W * (C / 4);
  PBE_CHECK_CUDA(launch_k(upsample2x_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, x, y, Nb, H, W, C, operand_f16()));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_upsample2x_16(const bf16* x, bf16* y, int Nb, int H, int W, int C, cudaStream_t stream) {
  PBE_REQUIRE(C % 8 == 0, "upsample channels % 8");
  const long long total = static_cast<long long>(Nb) * H * W * (C / 8);
  PBE_CHECK_CUDA(launch_k(upsample2x_16_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, x, y, Nb, H, W, C));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_cast_bf16(const float* x, bf16* y, size_t n, cudaStream_t stream) {
  PBE_REQUIRE(n % 4 == 0, "cast length % 4");
  const size_t n4 = n / 4;
  PBE_CHECK_CUDA(launch_k(cast_bf16_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, stream, x, y, n4, operand_f16()));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// C[M, N] = A[M, K] . B[K, N], fp32 in / out, fp64 accumulation: products of two weight matrices at load time (engine.cu:
// proj_out . ff.net.2 -- the composed matrix is rounded ONCE to the 16-bit operand format afterwards)
__global__ void __launch_bounds__(256) matmul_f64acc_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ C, int M, int K, int N) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    sa[ty][tx] = (row < M && k0 + tx < K) ? A[static_cast<size_t>(row) * K + k0 + tx] : 0.0f;
    sb[ty][tx] = (k0 + ty < K && col < N) ? B[static_cast<size_t>(k0 + ty) * N + col] : 0.0f;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) acc = fma(static_cast<double>(sa[ty][kk]), static_cast<double>(sb[kk][tx]), acc);
    __syncthreads();
  }
  if (row < M && col < N) C[static_cast<size_t>(row) * N + col] = static_cast<float>(acc);
}

int launch_matmul_f64acc(const float* A, const float* B, float* C, int M, int K, int N, cudaStream_t stream) {
  PBE_REQUIRE(M > 0 && K > 0 && N > 0, "matmul sizes");
  matmul_f64acc_kernel<<<dim3((N + 15) / 16, (M + 15) / 16), dim3(256), 0, stream>>>(A, B, C, M, K, N);   // load time: never under PDL
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_input(const float* x, bf16* y, int Nb, int Cin, int H, int W, int Cpad, cudaStream_t stream) {
  PBE_REQUIRE(Cpad % 8 == 0 && Cpad >= Cin, "padded channel count");
  const long long total = static_cast<long long>(Nb) * H * W;
  PBE_CHECK_CUDA(launch_k(pack_input_kernel, dim3(static_cast<unsigned>((total + 127) / 128)), dim3(128), 0, stream, x, y, Nb, Cin, H * W, Cpad, operand_f16()));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack_output(const float* y, float* out, int Nb, int Cout, int H, int W, int ld, cudaStream_t stream) {
  const long long total = static_cast<long long>(Nb) * H * W;
  PBE_CHECK_CUDA(launch_k(unpack_output_kernel, dim3(static_cast<unsigned>((total + 127) / 128)), dim3(128), 0, stream, y, out, Nb, Cout, H * W, ld));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_small_linear(const float* x, const float* W, const float* bias, float* y, int B, int K, int O, int pre_silu,
                        int post_act, cudaStream_t stream, float* y_silu, const float* residual) {
  PBE_REQUIRE(K % 4 == 0, "small_linear K % 4");
  PBE_CHECK_CUDA(launch_k(small_linear_kernel, dim3((O + SL_OW - 1) / SL_OW), dim3(SL_WARPS * 32), 0, stream, x, W, bias, y, y_silu, B, K, O, pre_silu, post_act, residual));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_vae_pack_input(const float* z, const float* Wp, const float* bp, bf16* y, int Nb, int Cin, int Cz, int H, int W,
                          int Cpad, cudaStream_t stream) {
  PBE_REQUIRE(Cin <= 8 && Cz <= Cpad, "vae_pack_input: embed_dim <= 8");
  const long long total = static_cast<long long>(Nb) * H * W;
  PBE_CHECK_CUDA(launch_k(vae_pack_input_kernel, dim3(static_cast<unsigned>((total + 127) / 128)), dim3(128), 0, stream, z, Wp, bp, y, Nb, Cin, Cz, H * W, Cpad, operand_f16()));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_vae_unpack_moments(const float* y, const float* Wq, const float* bq, float* out, int Nb, int Cin, int Cout,
                              int H, int W, int ld, cudaStream_t stream) {
  PBE_REQUIRE(Cin <= 16, "vae_unpack_moments: at most 16 moment channels");
  const long long total = static_cast<long long>(Nb) * H * W;
  PBE_CHECK_CUDA(launch_k(vae_unpack_moments_kernel, dim3(static_cast<unsigned>((total + 127) / 128)), dim3(128), 0, stream, y, Wq, bq, out, Nb, Cin, Cout,
                                                                                         H * W, ld));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_postprocess_u8(const float* x, uint8_t* out, int Nb, int C, int H, int W, cudaStream_t stream) {
  const long long total = static_cast<long long>(Nb) * H * W;
  PBE_CHECK_CUDA(launch_k(postprocess_u8_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, x,
                          out, Nb, C, H * W));
  return 0;
}

int launch_normalize_u8(const uint8_t* in, float* out, int Nb, int H, int W, const float* mean, const float* stdv,
                        cudaStream_t stream) {
  Norm3 nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean[c]; nm.std[c] = stdv[c]; }
  const long long total = static_cast<long long>(Nb) * H * W;
  PBE_CHECK_CUDA(launch_k(normalize_u8_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, in, out,
                          Nb, H * W, nm));
  return 0;
}

int launch_prepare_inpaint_u8(const uint8_t* img, const uint8_t* mask, float* image_out, float* mask_out, float* inpaint_out,
                              int Nb, int H, int W, int binarize, cudaStream_t stream) {
  const long long total = static_cast<long long>(Nb) * H * W;
  PBE_CHECK_CUDA(launch_k(prepare_inpaint_u8_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, img,
                          mask, image_out, mask_out, inpaint_out, Nb, H * W, binarize));
  return 0;
}

int launch_resize_bilinear(const float* in, float* out, int NC, int H, int W, int h, int w, int antialias, cudaStream_t stream) {
  PBE_REQUIRE(NC > 0 && H > 0 && W > 0 && h > 0 && w > 0, "resize_bilinear: empty geometry");
  const long long total = static_cast<long long>(NC) * h * w;
  const dim3 grid(static_cast<unsigned>((total + 127) / 128));
  if (antialias) PBE_CHECK_CUDA(launch_k(resize_bilinear_aa_kernel, grid, dim3(128), 0, stream, in, out, NC, H, W, h, w));
  else PBE_CHECK_CUDA(launch_k(resize_bilinear_kernel, grid, dim3(128), 0, stream, in, out, NC, H, W, h, w));
  return 0;
}

int launch_clip_pack_patches(const float* img, bf16* out, int B, int H, int W, int patch, int Kpad, cudaStream_t stream) {
  PBE_REQUIRE(H % patch == 0 && W % patch == 0 && Kpad >= 3 * patch * patch, "clip_pack_patches: image / patch geometry");
  const long long total = static_cast<long long>(B) * (H / patch) * (W / patch) * Kpad;
  PBE_CHECK_CUDA(launch_k(clip_pack_patches_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, img,
                          out, B, H, W, patch, Kpad, operand_f16()));
  return 0;
}

int launch_clip_embed(const float* patch_emb, const float* cls, const float* pos, float* tokens, int B, int P, int C,
                      cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * (P + 1) * C;
  PBE_CHECK_CUDA(launch_k(clip_embed_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, patch_emb,
                          cls, pos, tokens, B, P, C));
  return 0;
}

int launch_softmax_rows(const float* S, bf16* P, long long rows, int N, float scale, cudaStream_t stream) {
  PBE_REQUIRE(N % 4 == 0 && N <= 4 * 256 * SM_MAXV, "softmax_rows: row length % 4 == 0, <= 16384");
  PBE_CHECK_CUDA(launch_k(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(256), 0, stream, S, P, N, scale * 1.4426950408889634f, operand_f16()));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_timestep_embedding(const int64_t* t, float* out, int B, int dim, cudaStream_t stream) {
  const int n = B * (dim / 2);
  PBE_CHECK_CUDA(launch_k(timestep_embedding_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, t, out, B, dim));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_add_rowvec(const float* a, const float* v, float* y, int B, int N, cudaStream_t stream) {
  PBE_CHECK_CUDA(launch_k(add_rowvec_kernel, dim3((B * N + 255) / 256), dim3(256), 0, stream, a, v, y, B, N));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pbe
