// Layout / elementwise helpers and the small CUDA-core dense layers of the timestep-embedding path.
// All are HBM- or latency-bound; none is GEMM-shaped enough for the tensor cores (M = batch rows only).
#include "internal.h"

namespace pbe {

namespace {

__global__ void upsample2x_kernel(const float* __restrict__ x, bf16* __restrict__ y, int Nb, int H, int W, int C) {
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
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  const int W2 = 2 * W, H2 = 2 * H;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const long long o = ((static_cast<long long>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dx) * C + cv * 4;
      *reinterpret_cast<uint2*>(y + o) = pk;
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  reinterpret_cast<uint2*>(y)[i] = pk;
}

// thread per (n, pixel): gathers Cin channel planes (coalesced across threads), writes Cpad bf16 contiguous
__global__ void pack_input_kernel(const float* __restrict__ x, bf16* __restrict__ y, int Nb, int Cin, int HW,
                                  int Cpad) {
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
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]),
                   c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    pk.z = *reinterpret_cast<uint32_t*>(&c);
    pk.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(dst + c0) = pk;
  }
}

__global__ void unpack_output_kernel(const float* __restrict__ y, float* __restrict__ out, int Nb, int Cout, int HW,
                                     int ld) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(Nb) * HW) return;
  const int n = static_cast<int>(idx / HW);
  const int pix = static_cast<int>(idx % HW);
  for (int c = 0; c < Cout; ++c) out[(static_cast<long long>(n) * Cout + c) * HW + pix] = y[idx * ld + c];
}

__device__ __forceinline__ float silu_f(float t) { return t / (1.0f + expf(-t)); }

// one warp per output feature; batch rows processed 8 at a time
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           float* __restrict__ y_silu, int B, int K, int O,
                                                           int pre_silu, int post_silu) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= O) return;
  const float* wr = W + static_cast<long long>(o) * K;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 w = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (b0 + i < B) {
          float4 xv = *reinterpret_cast<const float4*>(x + static_cast<long long>(b0 + i) * K + k);
          if (pre_silu) {
            xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w);
          }
          acc[i] += (w.x * xv.x + w.y * xv.y) + (w.z * xv.z + w.w * xv.w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], s);
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (b0 + i < B) {
          float v = acc[i] + (bias ? bias[o] : 0.0f);
          if (post_silu) v = silu_f(v);
          y[static_cast<long long>(b0 + i) * O + o] = v;
          if (y_silu) y_silu[static_cast<long long>(b0 + i) * O + o] = silu_f(v);
        }
      }
    }
  }
}

// reference: timestep_embedding, ldm/modules/diffusionmodules/util.py:151-171  (cos || sin, fp32)
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, float* __restrict__ out, int B, int dim) {
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
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  y[idx] = a[idx % N] + v[idx];
}

}  // namespace

int launch_upsample2x_bf16(const float* x, bf16* y, int Nb, int H, int W, int C, cudaStream_t stream) {
  PBE_REQUIRE(C % 4 == 0, "upsample channels % 4");
  const long long total = static_cast<long long>(Nb) * H * W * (C / 4);
  upsample2x_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, y, Nb, H, W, C);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_cast_bf16(const float* x, bf16* y, size_t n, cudaStream_t stream) {
  PBE_REQUIRE(n % 4 == 0, "cast length % 4");
  const size_t n4 = n / 4;
  cast_bf16_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(x, y, n4);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_input(const float* x, bf16* y, int Nb, int Cin, int H, int W, int Cpad, cudaStream_t stream) {
  PBE_REQUIRE(Cpad % 8 == 0 && Cpad >= Cin, "padded channel count");
  const long long total = static_cast<long long>(Nb) * H * W;
  pack_input_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, stream>>>(x, y, Nb, Cin, H * W, Cpad);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_unpack_output(const float* y, float* out, int Nb, int Cout, int H, int W, int ld, cudaStream_t stream) {
  const long long total = static_cast<long long>(Nb) * H * W;
  unpack_output_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, stream>>>(y, out, Nb, Cout, H * W, ld);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_small_linear(const float* x, const float* W, const float* bias, float* y, int B, int K, int O, int pre_silu,
                        int post_silu, cudaStream_t stream, float* y_silu) {
  PBE_REQUIRE(K % 4 == 0, "small_linear K % 4");
  small_linear_kernel<<<(O + 7) / 8, 256, 0, stream>>>(x, W, bias, y, y_silu, B, K, O, pre_silu, post_silu);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_timestep_embedding(const int64_t* t, float* out, int B, int dim, cudaStream_t stream) {
  const int n = B * (dim / 2);
  timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, stream>>>(t, out, B, dim);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_add_rowvec(const float* a, const float* v, float* y, int B, int N, cudaStream_t stream) {
  add_rowvec_kernel<<<(B * N + 255) / 256, 256, 0, stream>>>(a, v, y, B, N);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pbe
