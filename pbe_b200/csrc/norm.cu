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

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_C = 2560;

__device__ __forceinline__ float4 ld4(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                                      long long pix, int c) {
  // c % 4 == 0, C0 % 4 == 0: a float4 never straddles the concat seam
  if (c < C0) return __ldg(reinterpret_cast<const float4*>(x0 + pix * C0 + c));
  return __ldg(reinterpret_cast<const float4*>(x1 + pix * C1 + (c - C0)));
}

// grid (slabs, Nb). partial[b][slab][g] = (sum, sumsq) over the slab's pixels and the group's channels.
// Threads own float4 channel columns (per-CHANNEL accumulators, so group boundaries that are not multiples of 4
// channels - 10, 30 channels per group - need no special casing); narrow tensors split the block over several pixel
// rows; four pixels are in flight per thread.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const float* __restrict__ x0, int C0,
                                                              const float* __restrict__ x1, int C1, int HW, int slabs,
                                                              float* __restrict__ partial) {
  __shared__ float s_sum[GN_MAX_C];
  __shared__ float s_sq[GN_MAX_C];
  const int C = C0 + C1;
  const int ncol4 = C / 4;
  const int b = blockIdx.y;
  const int slab = blockIdx.x;
  const int pix_per = (HW + slabs - 1) / slabs;
  const int p0 = slab * pix_per;
  const int p1 = min(HW, p0 + pix_per);
  const int rows_par = ncol4 <= GN_THREADS ? GN_THREADS / ncol4 : 1;
  const int ncols_thr = ncol4 <= GN_THREADS ? 1 : (ncol4 + GN_THREADS - 1) / GN_THREADS;  // <= 3
  const int prow = ncol4 <= GN_THREADS ? threadIdx.x / ncol4 : 0;
  const int col0 = ncol4 <= GN_THREADS ? threadIdx.x % ncol4 : threadIdx.x;
  const bool active = prow < rows_par;
  const long long base = static_cast<long long>(b) * HW;
  for (int ci = 0; ci < ncols_thr; ++ci) {
    const int col = col0 + ci * GN_THREADS;
    float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f, qx = 0.f, qy = 0.f, qz = 0.f, qw = 0.f;
    if (active && col < ncol4) {
      const int c = col * 4;
      int pix = p0 + prow;
      for (; pix + 3 * rows_par < p1; pix += 4 * rows_par) {
        const float4 a0 = ld4(x0, C0, x1, C1, base + pix, c);
        const float4 a1 = ld4(x0, C0, x1, C1, base + pix + rows_par, c);
        const float4 a2 = ld4(x0, C0, x1, C1, base + pix + 2 * rows_par, c);
        const float4 a3 = ld4(x0, C0, x1, C1, base + pix + 3 * rows_par, c);
        sx += (a0.x + a1.x) + (a2.x + a3.x); qx += (a0.x * a0.x + a1.x * a1.x) + (a2.x * a2.x + a3.x * a3.x);
        sy += (a0.y + a1.y) + (a2.y + a3.y); qy += (a0.y * a0.y + a1.y * a1.y) + (a2.y * a2.y + a3.y * a3.y);
        sz += (a0.z + a1.z) + (a2.z + a3.z); qz += (a0.z * a0.z + a1.z * a1.z) + (a2.z * a2.z + a3.z * a3.z);
        sw += (a0.w + a1.w) + (a2.w + a3.w); qw += (a0.w * a0.w + a1.w * a1.w) + (a2.w * a2.w + a3.w * a3.w);
      }
      for (; pix < p1; pix += rows_par) {
        const float4 a0 = ld4(x0, C0, x1, C1, base + pix, c);
        sx += a0.x; qx += a0.x * a0.x;
        sy += a0.y; qy += a0.y * a0.y;
        sz += a0.z; qz += a0.z * a0.z;
        sw += a0.w; qw += a0.w * a0.w;
      }
    }
    // cross-row reduction in a fixed order: row r adds into the channel slot after rows < r
    for (int r = 0; r < rows_par; ++r) {
      if (active && prow == r && col < ncol4) {
        const int c = col * 4;
        if (r == 0) {
          s_sum[c] = sx; s_sum[c + 1] = sy; s_sum[c + 2] = sz; s_sum[c + 3] = sw;
          s_sq[c] = qx; s_sq[c + 1] = qy; s_sq[c + 2] = qz; s_sq[c + 3] = qw;
        } else {
          s_sum[c] += sx; s_sum[c + 1] += sy; s_sum[c + 2] += sz; s_sum[c + 3] += sw;
          s_sq[c] += qx; s_sq[c + 1] += qy; s_sq[c + 2] += qz; s_sq[c + 3] += qw;
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    const int cpg = C / 32;
    float a = 0.0f, q = 0.0f;
    for (int i = 0; i < cpg; ++i) {
      a += s_sum[g * cpg + i];
      q += s_sq[g * cpg + i];
    }
    float* dst = partial + ((static_cast<long long>(b) * slabs + slab) * 32 + g) * 2;
    dst[0] = a;
    dst[1] = q;
  }
}

// grid (Nb): per-(sample, channel) affine y = x * a + b with a = gamma * rstd, b = beta - mean * a.
__global__ void __launch_bounds__(GN_THREADS) gn_finalize_kernel(const float* __restrict__ partial, int slabs, int HW,
                                                                 int C, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps,
                                                                 float2* __restrict__ ab) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.x;
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
    ab[static_cast<long long>(b) * C + c] = make_float2(a, beta[c] - s_mean[g] * a);
  }
}

// grid (32 groups, Nb): same result as gn_finalize_kernel, but from the per-32-row column statistics the producing
// GEMM epilogues wrote ([Nb*HW/32][C_i][2] per source): no pass over x at all.  One small CTA per (sample, group):
// threads = (channel of the group) x (interleaved subsets of the row blocks); fixed-order double reduction.
constexpr int GNF_THREADS = 128;
__global__ void __launch_bounds__(GNF_THREADS) gn_finalize_fused_kernel(const float* __restrict__ st0, int C0,
                                                                        const float* __restrict__ st1, int C1, int HW,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, float eps,
                                                                        float2* __restrict__ ab) {
  __shared__ double s_a[GNF_THREADS], s_q[GNF_THREADS];
  __shared__ float s_mean, s_rstd;
  const int C = C0 + C1;
  const int cpg = C / 32;
  const int g = blockIdx.x, b = blockIdx.y;
  const int blocks = HW / 32;
  const int nsub = GNF_THREADS / cpg;          // >= 1 (cpg <= 80)
  const int t = threadIdx.x;
  const int ci = t % cpg, sub = t / cpg;
  double a = 0.0, q = 0.0;
  if (sub < nsub) {
    const int c = g * cpg + ci;
    const float* src = (c < C0) ? st0 + (static_cast<long long>(b) * blocks * C0 + c) * 2
                                : st1 + (static_cast<long long>(b) * blocks * C1 + (c - C0)) * 2;
    const long long ld = static_cast<long long>((c < C0) ? C0 : C1) * 2;
    int k = sub;
    for (; k + 3 * nsub < blocks; k += 4 * nsub) {   // four independent loads in flight
      const float2 v0 = *reinterpret_cast<const float2*>(src + k * ld);
      const float2 v1 = *reinterpret_cast<const float2*>(src + (k + nsub) * ld);
      const float2 v2 = *reinterpret_cast<const float2*>(src + (k + 2 * nsub) * ld);
      const float2 v3 = *reinterpret_cast<const float2*>(src + (k + 3 * nsub) * ld);
      a += (static_cast<double>(v0.x) + static_cast<double>(v1.x)) + (static_cast<double>(v2.x) + static_cast<double>(v3.x));
      q += (static_cast<double>(v0.y) + static_cast<double>(v1.y)) + (static_cast<double>(v2.y) + static_cast<double>(v3.y));
    }
    for (; k < blocks; k += nsub) {
      const float2 v = *reinterpret_cast<const float2*>(src + k * ld);
      a += static_cast<double>(v.x);
      q += static_cast<double>(v.y);
    }
  }
  s_a[t] = a;
  s_q[t] = q;
  __syncthreads();
  if (t == 0) {
    double sa = 0.0, sq = 0.0;
    for (int i = 0; i < cpg * nsub; ++i) {
      sa += s_a[i];
      sq += s_q[i];
    }
    const double cnt = static_cast<double>(HW) * cpg;
    const double mean = sa / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean = static_cast<float>(mean);
    s_rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  if (t < cpg) {
    const int c = g * cpg + t;
    const float sc = gamma[c] * s_rstd;
    ab[static_cast<long long>(b) * C + c] = make_float2(sc, beta[c] - s_mean * sc);
  }
}

// pure streaming pass: four float4 of x per thread, all loads issued before the first use
__device__ __forceinline__ void gn_apply_one(const float4 v, const float4 ab01, const float4 ab23, int silu,
                                             bf16* __restrict__ y, bf16* __restrict__ raw, long long off) {
  float o0 = fmaf(v.x, ab01.x, ab01.y), o1 = fmaf(v.y, ab01.z, ab01.w);
  float o2 = fmaf(v.z, ab23.x, ab23.y), o3 = fmaf(v.w, ab23.z, ab23.w);
  if (silu) {
    o0 = o0 / (1.0f + __expf(-o0)); o1 = o1 / (1.0f + __expf(-o1));
    o2 = o2 / (1.0f + __expf(-o2)); o3 = o3 / (1.0f + __expf(-o3));
  }
  uint2 pk;
  pk.x = pack2(o0, o1);
  pk.y = pack2(o2, o3);
  *reinterpret_cast<uint2*>(y + off) = pk;
  if (raw != nullptr) {
    uint2 rk;
    rk.x = pack2(v.x, v.y);
    rk.y = pack2(v.z, v.w);
    *reinterpret_cast<uint2*>(raw + off) = rk;
  }
}

__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const float* __restrict__ x0, int C0,
                                                              const float* __restrict__ x1, int C1, int HW,
                                                              long long total_vec, const float2* __restrict__ ab,
                                                              int silu, bf16* __restrict__ y, bf16* __restrict__ raw) {
  const int C = C0 + C1;
  const int vec_per_pix = C / 4;
  const long long base = static_cast<long long>(blockIdx.x) * (GN_THREADS * 4) + threadIdx.x;
  float4 v[4], a01[4], a23[4];
  long long off[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long idx = base + u * GN_THREADS;
    off[u] = -1;
    if (idx < total_vec) {
      const long long gp = idx / vec_per_pix;
      const int c = static_cast<int>(idx - gp * vec_per_pix) * 4;
      const int b = static_cast<int>(gp / HW);
      v[u] = ld4(x0, C0, x1, C1, gp, c);
      a01[u] = __ldg(reinterpret_cast<const float4*>(ab + static_cast<long long>(b) * C + c));
      a23[u] = __ldg(reinterpret_cast<const float4*>(ab + static_cast<long long>(b) * C + c + 2));
      off[u] = gp * C + c;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (off[u] >= 0) gn_apply_one(v[u], a01[u], a23[u], silu, y, raw, off[u]);
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

int gn_workspace_floats(int Nb, int HW, int C) {
  return Nb * gn_num_slabs(HW) * 64 + Nb * C * 2 + 64;
}

int launch_groupnorm(const GroupNormArgs& a, cudaStream_t stream) {
  const int C = a.C0 + a.C1;
  PBE_REQUIRE(C % 64 == 0 && C <= GN_MAX_C, "GroupNorm channels must be a multiple of 64, <= 2560");
  PBE_REQUIRE(a.C0 % 4 == 0 && a.C1 % 4 == 0, "GroupNorm concat halves must be multiples of 4 channels");
  const int slabs = gn_num_slabs(a.HW);
  float* partial = a.partial;
  float2* ab = reinterpret_cast<float2*>(a.partial + ((static_cast<size_t>(a.Nb) * slabs * 64 + 3) & ~static_cast<size_t>(3)));
  const bool fused = a.stats0 != nullptr && (a.C1 == 0 || a.stats1 != nullptr) && a.HW % 32 == 0;
  if (fused) {
    gn_finalize_fused_kernel<<<dim3(32, a.Nb), GNF_THREADS, 0, stream>>>(a.stats0, a.C0, a.stats1, a.C1, a.HW, a.gamma,
                                                                        a.beta, a.eps, ab);
    PBE_CHECK_CUDA(cudaGetLastError());
  } else {
    gn_stats_kernel<<<dim3(slabs, a.Nb), GN_THREADS, 0, stream>>>(a.x0, a.C0, a.x1, a.C1, a.HW, slabs, partial);
    PBE_CHECK_CUDA(cudaGetLastError());
    gn_finalize_kernel<<<a.Nb, GN_THREADS, 0, stream>>>(partial, slabs, a.HW, C, a.gamma, a.beta, a.eps, ab);
    PBE_CHECK_CUDA(cudaGetLastError());
  }
  const long long total_vec = static_cast<long long>(a.Nb) * a.HW * (C / 4);
  long long blocks = (total_vec + GN_THREADS * 4 - 1) / (GN_THREADS * 4);
  if (blocks < 1) blocks = 1;
  gn_apply_kernel<<<static_cast<unsigned>(blocks), GN_THREADS, 0, stream>>>(a.x0, a.C0, a.x1, a.C1, a.HW, total_vec, ab,
                                                                         a.silu, a.y, a.raw);
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
