// GroupNorm(32) (+SiLU) and LayerNorm producers of the bf16 GEMM operands.  HBM-bound passes:
//   GroupNorm: 2 reads of fp32 x (stats + apply) + 1 bf16 write (+1 optional raw bf16 write)
//   LayerNorm: 1 read of fp32 x (row kept in registers) + 1 bf16 write
// Statistics stay in fp32/double and the reduction order is fixed (no atomics) -> bit-reproducible.
//
// Replaces GroupNorm32 + SiLU (ldm/modules/diffusionmodules/util.py:199-216, openaimodel.py:201-204,225-228,824-826),
// attention.Normalize (ldm/modules/attention.py:77-78) and LayerNorm norm1/norm3 (attention.py:240-242).
// The GroupNorm input may be the channel concat of two tensors (th.cat([h, hs.pop()], 1), openaimodel.py:883):
// the concat is never materialised in fp32, only the normalised bf16 operand is.
#include "internal.h"

namespace pbe {

namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_COLS = 5;  // float2 columns per thread: C/2 <= 5*256 -> C <= 2560

__device__ __forceinline__ float2 ld2(const float* x0, int C0, const float* x1, int C1, long long pix, int c) {
  // c even; C0 even
  if (c < C0) return *reinterpret_cast<const float2*>(x0 + pix * C0 + c);
  return *reinterpret_cast<const float2*>(x1 + pix * C1 + (c - C0));
}

// grid (slabs, Nb). partial[b][slab][g] = (sum, sumsq) over the slab's pixels and the group's channels.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const float* __restrict__ x0, int C0,
                                                              const float* __restrict__ x1, int C1, int HW, int slabs,
                                                              float* __restrict__ partial) {
  __shared__ float s_sum[GN_THREADS * GN_MAX_COLS];
  __shared__ float s_sq[GN_THREADS * GN_MAX_COLS];
  const int C = C0 + C1;
  const int ncols = C / 2;
  const int b = blockIdx.y;
  const int slab = blockIdx.x;
  const int pix_per = (HW + slabs - 1) / slabs;
  const int p0 = slab * pix_per;
  const int p1 = min(HW, p0 + pix_per);
  float sum[GN_MAX_COLS], sq[GN_MAX_COLS];
#pragma unroll
  for (int i = 0; i < GN_MAX_COLS; ++i) sum[i] = sq[i] = 0.0f;
  for (int pix = p0; pix < p1; ++pix) {
    const long long gp = static_cast<long long>(b) * HW + pix;
#pragma unroll
    for (int i = 0; i < GN_MAX_COLS; ++i) {
      const int col = threadIdx.x + i * GN_THREADS;
      if (col < ncols) {
        const float2 v = ld2(x0, C0, x1, C1, gp, col * 2);
        sum[i] += v.x + v.y;
        sq[i] += v.x * v.x + v.y * v.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < GN_MAX_COLS; ++i) {
    s_sum[threadIdx.x + i * GN_THREADS] = sum[i];
    s_sq[threadIdx.x + i * GN_THREADS] = sq[i];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    const int cols_per_group = ncols / 32;
    float a = 0.0f, q = 0.0f;
    for (int i = 0; i < cols_per_group; ++i) {
      a += s_sum[g * cols_per_group + i];
      q += s_sq[g * cols_per_group + i];
    }
    float* dst = partial + ((static_cast<long long>(b) * slabs + slab) * 32 + g) * 2;
    dst[0] = a;
    dst[1] = q;
  }
}

// grid (pixel blocks, Nb)
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const float* __restrict__ x0, int C0,
                                                              const float* __restrict__ x1, int C1, int HW, int slabs,
                                                              const float* __restrict__ partial,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int silu,
                                                              bf16* __restrict__ y, bf16* __restrict__ raw,
                                                              int pix_per_block) {
  extern __shared__ float s_ab[];  // a[C], b[C]
  __shared__ float s_mean[32], s_rstd[32];
  const int C = C0 + C1;
  const int b = blockIdx.y;
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    double a = 0.0, q = 0.0;
    for (int s = 0; s < slabs; ++s) {
      const float* src = partial + ((static_cast<long long>(b) * slabs + s) * 32 + g) * 2;
      a += static_cast<double>(src[0]);
      q += static_cast<double>(src[1]);
    }
    const double cnt = static_cast<double>(HW) * (C / 32);
    const double mean = a / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const int cpg = C / 32;
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const int g = c / cpg;
    const float a = gamma[c] * s_rstd[g];
    s_ab[c] = a;
    s_ab[C + c] = beta[c] - s_mean[g] * a;
  }
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const int vec_per_pix = C / 4;
  const int total = (p1 - p0) * vec_per_pix;
  for (int idx = threadIdx.x; idx < total; idx += GN_THREADS) {
    const int pix = p0 + idx / vec_per_pix;
    const int c = (idx % vec_per_pix) * 4;
    const long long gp = static_cast<long long>(b) * HW + pix;
    float4 v;
    if (c < C0) v = *reinterpret_cast<const float4*>(x0 + gp * C0 + c);
    else v = *reinterpret_cast<const float4*>(x1 + gp * C1 + (c - C0));
    float r[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = r[i] * s_ab[c + i] + s_ab[C + c + i];
      if (silu) t = t / (1.0f + __expf(-t));
      o[i] = t;
    }
    const long long off = gp * C + c;
    __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]);
    __nv_bfloat162 hi = __floats2bfloat162_rn(o[2], o[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y + off) = pk;
    if (raw != nullptr) {
      __nv_bfloat162 rlo = __floats2bfloat162_rn(r[0], r[1]);
      __nv_bfloat162 rhi = __floats2bfloat162_rn(r[2], r[3]);
      uint2 rk;
      rk.x = *reinterpret_cast<uint32_t*>(&rlo);
      rk.y = *reinterpret_cast<uint32_t*>(&rhi);
      *reinterpret_cast<uint2*>(raw + off) = rk;
    }
  }
}

// One warp per row; row held in registers (C <= 1280).
constexpr int LN_MAX_VEC = 10;
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ y, int M,
                                                        int C, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C / 4;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  float4 v[LN_MAX_VEC];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      v[i] = xr[k];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, d2 = v[i].w - mean;
      sq += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  const float4* gr = reinterpret_cast<const float4*>(gamma);
  const float4* br = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long long>(row) * C);
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      const float4 g = gr[k], bb = br[k];
      __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      yr[k] = pk;
    }
  }
}

}  // namespace

int gn_num_slabs(int HW) {
  int s = HW / 64;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

int launch_groupnorm(const GroupNormArgs& a, cudaStream_t stream) {
  const int C = a.C0 + a.C1;
  PBE_REQUIRE(C % 64 == 0 && C <= 2 * GN_THREADS * GN_MAX_COLS, "GroupNorm channels must be a multiple of 64, <= 2560");
  PBE_REQUIRE(a.C0 % 4 == 0 && a.C1 % 4 == 0, "GroupNorm concat halves must be multiples of 4 channels");
  const int slabs = gn_num_slabs(a.HW);
  gn_stats_kernel<<<dim3(slabs, a.Nb), GN_THREADS, 0, stream>>>(a.x0, a.C0, a.x1, a.C1, a.HW, slabs, a.partial);
  PBE_CHECK_CUDA(cudaGetLastError());
  int pix_per_block = (64 * 1024) / (C * 4);  // ~64 KB of input per block
  if (pix_per_block < 1) pix_per_block = 1;
  const int blocks = (a.HW + pix_per_block - 1) / pix_per_block;
  gn_apply_kernel<<<dim3(blocks, a.Nb), GN_THREADS, 2 * C * sizeof(float), stream>>>(
      a.x0, a.C0, a.x1, a.C1, a.HW, slabs, a.partial, a.gamma, a.beta, a.eps, a.silu, a.y, a.raw, pix_per_block);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, bf16* y, int M, int C, float eps,
                     cudaStream_t stream) {
  PBE_REQUIRE(C % 4 == 0 && C / 4 <= 32 * LN_MAX_VEC, "LayerNorm width must be a multiple of 4, <= 1280");
  layernorm_kernel<<<(M + 7) / 8, 256, 0, stream>>>(x, gamma, beta, y, M, C, eps);
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pbe
