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
#include "ptx.cuh"

#include <stdlib.h>

#include <algorithm>

namespace pbe {

namespace {


constexpr int GN_THREADS = 256;
constexpr int GN_MAX_C = 2560;

// four consecutive channels starting at element index `idx` of a tensor that holds fp32 (in16 = 0) or 16-bit operands
__device__ __forceinline__ float4 ld4e(const float* __restrict__ x, long long idx, int in16, int f16) {
  if (!in16) return __ldg(reinterpret_cast<const float4*>(x + idx));
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(x) + idx));
  const float2 a = unpack_op2(u.x, f16), b = unpack_op2(u.y, f16);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 cvt4(uint2 u, int f16) {
  const float2 a = unpack_op2(u.x, f16), b = unpack_op2(u.y, f16);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 ldraw4(const float* __restrict__ x, long long idx) {   // four 16-bit operands, unconverted
  return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(x) + idx));
}
__device__ __forceinline__ float2 ld2e(const float* __restrict__ x, long long idx, int in16, int f16) {
  if (!in16) return __ldg(reinterpret_cast<const float2*>(x + idx));
  return unpack_op2(__ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned short*>(x) + idx)), f16);
}
__device__ __forceinline__ float4 ld4(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                                      long long pix, int c, int in16, int f16) {
  // c % 4 == 0, C0 % 4 == 0: a float4 never straddles the concat seam
  if (c < C0) return ld4e(x0, pix * C0 + c, in16, f16);
  return ld4e(x1, pix * C1 + (c - C0), in16, f16);
}

// grid (slabs, Nb). partial[b][slab][g] = (sum, sumsq) over the slab's pixels and the group's channels.
// Threads own float4 channel columns (per-CHANNEL accumulators, so group boundaries that are not multiples of 4
// channels - 10, 30 channels per group - need no special casing); narrow tensors split the block over several pixel
// rows; four pixels are in flight per thread.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const float* __restrict__ x0, int C0,
                                                              const float* __restrict__ x1, int C1, int HW, int slabs,
                                                              float* __restrict__ partial, int in16, int f16) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
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
        float4 a0, a1, a2, a3;
        if (in16) {   // all four loads in flight before the first conversion
          const float* sp = (c < C0) ? x0 : x1;
          const int cc = (c < C0) ? c : c - C0, cl = (c < C0) ? C0 : C1;
          const uint2 r0 = ldraw4(sp, (base + pix) * cl + cc), r1 = ldraw4(sp, (base + pix + rows_par) * cl + cc),
                      r2 = ldraw4(sp, (base + pix + 2 * rows_par) * cl + cc), r3 = ldraw4(sp, (base + pix + 3 * rows_par) * cl + cc);
          a0 = cvt4(r0, f16); a1 = cvt4(r1, f16); a2 = cvt4(r2, f16); a3 = cvt4(r3, f16);
        } else {
          a0 = ld4(x0, C0, x1, C1, base + pix, c, 0, f16);
          a1 = ld4(x0, C0, x1, C1, base + pix + rows_par, c, 0, f16);
          a2 = ld4(x0, C0, x1, C1, base + pix + 2 * rows_par, c, 0, f16);
          a3 = ld4(x0, C0, x1, C1, base + pix + 3 * rows_par, c, 0, f16);
        }
        sx += (a0.x + a1.x) + (a2.x + a3.x); qx += (a0.x * a0.x + a1.x * a1.x) + (a2.x * a2.x + a3.x * a3.x);
        sy += (a0.y + a1.y) + (a2.y + a3.y); qy += (a0.y * a0.y + a1.y * a1.y) + (a2.y * a2.y + a3.y * a3.y);
        sz += (a0.z + a1.z) + (a2.z + a3.z); qz += (a0.z * a0.z + a1.z * a1.z) + (a2.z * a2.z + a3.z * a3.z);
        sw += (a0.w + a1.w) + (a2.w + a3.w); qw += (a0.w * a0.w + a1.w * a1.w) + (a2.w * a2.w + a3.w * a3.w);
      }
      for (; pix < p1; pix += rows_par) {
        const float4 a0 = ld4(x0, C0, x1, C1, base + pix, c, in16, f16);
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
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
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
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
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
    // every load of this thread is issued before the first add (the kernel is one wave of tiny CTAs: pure latency);
    // the order of the additions is fixed (k ascending), so the result does not depend on timing
    constexpr int MAXL = 16;
    int k = sub;
    while (k < blocks) {
      float2 v[MAXL];
      int n = 0;
#pragma unroll
      for (int i = 0; i < MAXL; ++i) {
        const int kk = k + i * nsub;
        if (kk < blocks) { v[i] = *reinterpret_cast<const float2*>(src + kk * ld); n = i + 1; }
      }
#pragma unroll
      for (int i = 0; i < MAXL; ++i) {
        if (i < n) { a += static_cast<double>(v[i].x); q += static_cast<double>(v[i].y); }
      }
      k += MAXL * nsub;
    }
  }
  // fixed-order tree: butterfly inside each warp, then the four warp totals in order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((t & 31) == 0) { s_a[t >> 5] = a; s_q[t >> 5] = q; }
  __syncthreads();
  if (t == 0) {
    double sa = 0.0, sq = 0.0;
    for (int i = 0; i < GNF_THREADS / 32; ++i) {
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

// Streaming pass y = silu(x * a + b) (+ raw bf16 copy).  grid (pixel blocks, Nb); a thread owns ONE float4 channel
// column of one sample (its scale / shift are loaded once, no index arithmetic per element) and walks GN_UNROLL pixels
// with every load issued before the first use.  (The first version recomputed pixel / sample / channel with two 64-bit
// divisions per float4 and re-read the scale / shift table per element: 2.2-3.4 TB/s.)
constexpr int GN_UNROLL = 8;
// x * sigmoid(x) = x * rcp(1 + 2^(-x log2 e)): FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL.  (__fdividef carries a range fix-up for
// denominators above 2^126 -- an FSETP and two more FMULs per element; here the denominator overflowing to +inf gives
// rcp = 0 and the right limit.  ncu, profiles/r02_ncu_gnapply_v4_summary.txt: the pass was issue-bound at 24 instructions
// per element, 75 % issue-active, not HBM-bound.)
__device__ __forceinline__ float silu_fast(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

template <int IN16>
__global__ void __launch_bounds__(1024) gn_apply_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1,
                                                        int C1, int HW, int rows_par, const float2* __restrict__ ab,
                                                        int silu, bf16* __restrict__ y, bf16* __restrict__ raw, int f16) {
  constexpr int in16 = IN16;
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const int C = C0 + C1;
  const int vpp = C / 4;
  const int t = threadIdx.x;
  const int prow = t / vpp;          // blockDim.x == vpp * rows_par
  const int c = (t - prow * vpp) * 4;
  const int b = blockIdx.y;
  const int p_begin = blockIdx.x * (rows_par * GN_UNROLL) + prow;
  const float* __restrict__ src;     // element-indexed through ld4e (fp32 or 16-bit source)
  long long sbase;
  int ld;
  if (c < C0) { src = x0; sbase = static_cast<long long>(b) * HW * C0 + c; ld = C0; }
  else { src = x1; sbase = static_cast<long long>(b) * HW * C1 + (c - C0); ld = C1; }
  const float4 s01 = __ldg(reinterpret_cast<const float4*>(ab + static_cast<long long>(b) * C + c));
  const float4 s23 = __ldg(reinterpret_cast<const float4*>(ab + static_cast<long long>(b) * C + c + 2));
  const long long obase = static_cast<long long>(b) * HW * C + c;
  float4 v[GN_UNROLL];
  if (in16) {
    // every load is issued before the first conversion (a load whose result is converted on the spot reuses its
    // destination registers and serialises on the previous one: 2 TB/s instead of 4)
    uint2 r16[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int pix = p_begin + u * rows_par;
      r16[u] = (pix < HW) ? ldraw4(src, sbase + static_cast<long long>(pix) * ld) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) v[u] = cvt4(r16[u], f16);
  } else {
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int pix = p_begin + u * rows_par;
      if (pix < HW) v[u] = __ldg(reinterpret_cast<const float4*>(src + sbase + static_cast<long long>(pix) * ld));
    }
  }
#pragma unroll
  for (int u = 0; u < GN_UNROLL; ++u) {
    const int pix = p_begin + u * rows_par;
    if (pix < HW) {
      float o0, o1, o2, o3;
      upk2(fma2(pk2(v[u].x, v[u].y), pk2(s01.x, s01.z), pk2(s01.y, s01.w)), o0, o1);
      upk2(fma2(pk2(v[u].z, v[u].w), pk2(s23.x, s23.z), pk2(s23.y, s23.w)), o2, o3);
      if (silu) { o0 = silu_fast(o0); o1 = silu_fast(o1); o2 = silu_fast(o2); o3 = silu_fast(o3); }
      const long long off = obase + static_cast<long long>(pix) * C;
      *reinterpret_cast<uint2*>(y + off) = make_uint2(pack_op2(o0, o1, f16), pack_op2(o2, o3, f16));
      if (raw != nullptr)
        *reinterpret_cast<uint2*>(raw + off) = make_uint2(pack_op2(v[u].x, v[u].y, f16), pack_op2(v[u].z, v[u].w, f16));
    }
  }
}

// The same pass for the 16-bit stream, eight channels per thread: one 16-byte load and one 16-byte store per pixel and
// thread, byte pointers advanced by a constant per pixel, SiLU as h + h * tanh(h) with h = x / 2 (packed FMUL2, two
// MUFU.TANH, packed FFMA2: 2 issue slots per element instead of 5; tanh.approx is 2^-11 relative, the size of the 16-bit
// rounding the result gets next -- the v1 eps error against the fp32 reference is unchanged at 1.34e-3).  The 4-channel
// kernel above spent ~20 issue slots per element (64-bit index arithmetic per pixel, five-slot SiLU, one conversion
// per element) and ran at 4 TB/s, issue-bound (profiles/r02_ncu_gnapply_v4_summary.txt); the raw concat copy needs no
// conversion at all here (same bits).
constexpr int GN8_UNROLL = 4;
__device__ __forceinline__ void silu2_tanh(float& a, float& b) {
  float h0, h1, t0, t1;
  upk2(mul2(pk2(a, b), pk2(0.5f, 0.5f)), h0, h1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
  upk2(fma2(pk2(h0, h1), pk2(t0, t1), pk2(h0, h1)), a, b);
}
__global__ void __launch_bounds__(1024) gn_apply8_kernel(const unsigned short* __restrict__ x0, int C0,
                                                         const unsigned short* __restrict__ x1, int C1, int HW, int rows_par,
                                                         const float2* __restrict__ ab, int silu, bf16* __restrict__ y,
                                                         bf16* __restrict__ raw, int f16) {
  griddep_enter();
  const int C = C0 + C1;
  const int vpp = C / 8;
  const int t = threadIdx.x;
  const int prow = t / vpp;
  const int c = (t - prow * vpp) * 8;
  const int b = blockIdx.y;
  const int p_begin = blockIdx.x * (rows_par * GN8_UNROLL) + prow;
  const char* sp;
  size_t sstep;
  if (c < C0) {
    sp = reinterpret_cast<const char*>(x0 + (static_cast<size_t>(b) * HW + p_begin) * C0 + c);
    sstep = static_cast<size_t>(rows_par) * C0 * 2;
  } else {
    sp = reinterpret_cast<const char*>(x1 + (static_cast<size_t>(b) * HW + p_begin) * C1 + (c - C0));
    sstep = static_cast<size_t>(rows_par) * C1 * 2;
  }
  const size_t ooff = ((static_cast<size_t>(b) * HW + p_begin) * C + c) * 2, ostep = static_cast<size_t>(rows_par) * C * 2;
  // every load first: the pixels' data, then the eight (scale, shift) pairs of my channels
  uint4 r[GN8_UNROLL];
#pragma unroll
  for (int u = 0; u < GN8_UNROLL; ++u)
    r[u] = (p_begin + u * rows_par < HW) ? __ldg(reinterpret_cast<const uint4*>(sp + u * sstep)) : make_uint4(0u, 0u, 0u, 0u);
  const float4* abp = reinterpret_cast<const float4*>(ab + static_cast<size_t>(b) * C + c);
  const float4 s0 = __ldg(abp), s1 = __ldg(abp + 1), s2 = __ldg(abp + 2), s3 = __ldg(abp + 3);   // (a, b) of channels c .. c + 7
  const f32x2 a01 = pk2(s0.x, s0.z), b01 = pk2(s0.y, s0.w), a23 = pk2(s1.x, s1.z), b23 = pk2(s1.y, s1.w);
  const f32x2 a45 = pk2(s2.x, s2.z), b45 = pk2(s2.y, s2.w), a67 = pk2(s3.x, s3.z), b67 = pk2(s3.y, s3.w);
  char* yp = reinterpret_cast<char*>(y) + ooff;
  char* rp = raw != nullptr ? reinterpret_cast<char*>(raw) + ooff : nullptr;
#pragma unroll
  for (int u = 0; u < GN8_UNROLL; ++u) {
    if (p_begin + u * rows_par < HW) {
      const float2 v01 = unpack_op2(r[u].x, f16), v23 = unpack_op2(r[u].y, f16), v45 = unpack_op2(r[u].z, f16), v67 = unpack_op2(r[u].w, f16);
      float o0, o1, o2, o3, o4, o5, o6, o7;
      upk2(fma2(pk2(v01.x, v01.y), a01, b01), o0, o1);
      upk2(fma2(pk2(v23.x, v23.y), a23, b23), o2, o3);
      upk2(fma2(pk2(v45.x, v45.y), a45, b45), o4, o5);
      upk2(fma2(pk2(v67.x, v67.y), a67, b67), o6, o7);
      if (silu) { silu2_tanh(o0, o1); silu2_tanh(o2, o3); silu2_tanh(o4, o5); silu2_tanh(o6, o7); }
      *reinterpret_cast<uint4*>(yp + u * ostep) =
          make_uint4(pack_op2(o0, o1, f16), pack_op2(o2, o3, f16), pack_op2(o4, o5, f16), pack_op2(o6, o7, f16));
      if (rp != nullptr) *reinterpret_cast<uint4*>(rp + u * ostep) = r[u];
    }
  }
}

// Low-resolution GroupNorm in ONE kernel: grid (32 groups, Nb); the (sample, group) slice (HW pixels x C/32 channels,
// <= 80 KB of fp32) is staged in shared memory, so x is read once from HBM, statistics are an exact two-pass
// (mean, then centred sum of squares; fixed-order reductions) and the normalised bf16 operand is written straight
// from shared memory.  Replaces the stats -> finalize -> apply chain for the 8x8 / 16x16 levels, where three launches
// over a few MB were pure latency.
constexpr int GNS_THREADS = 512;
__device__ __forceinline__ double block_sum_double(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int i = 0; i < GNS_THREADS / 32; ++i) tot += s_red[i];
  return tot;
}
__global__ void __launch_bounds__(GNS_THREADS) gn_small_kernel(const float* __restrict__ x0, int C0,
                                                               const float* __restrict__ x1, int C1, int HW, int rows_par,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, int silu,
                                                               bf16* __restrict__ y, bf16* __restrict__ raw, int f16, int in16) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  extern __shared__ float2 s_x[];  // [HW][cpg / 2]
  __shared__ double s_red[GNS_THREADS / 32];
  const int C = C0 + C1;
  const int cpg = C / 32;
  const int hv = cpg / 2;  // float2 per pixel of this group (cpg is even; a float2 never straddles the concat seam)
  const int g = blockIdx.x, b = blockIdx.y;
  // a thread owns one channel pair and walks pixels prow, prow + rows_par, ... (threads >= hv * rows_par idle)
  const int prow = threadIdx.x / hv;
  const int cv = threadIdx.x - prow * hv;
  const bool active = prow < rows_par;
  const int c = g * cpg + cv * 2;
  const float* __restrict__ src;     // element-indexed through ld2e (fp32 or 16-bit source)
  long long sbase;
  int ld;
  if (c < C0) { src = x0; sbase = static_cast<long long>(b) * HW * C0 + c; ld = C0; }
  else { src = x1; sbase = static_cast<long long>(b) * HW * C1 + (c - C0); ld = C1; }
  float ps = 0.0f;
  if (active) {
    int pix = prow;
    for (; pix + 3 * rows_par < HW; pix += 4 * rows_par) {
      float2 v0, v1, v2, v3;
      if (in16) {   // all four loads in flight before the first conversion
        const unsigned short* s16p = reinterpret_cast<const unsigned short*>(src) + sbase;
        const uint32_t r0 = __ldg(reinterpret_cast<const uint32_t*>(s16p + static_cast<long long>(pix) * ld));
        const uint32_t r1 = __ldg(reinterpret_cast<const uint32_t*>(s16p + static_cast<long long>(pix + rows_par) * ld));
        const uint32_t r2 = __ldg(reinterpret_cast<const uint32_t*>(s16p + static_cast<long long>(pix + 2 * rows_par) * ld));
        const uint32_t r3 = __ldg(reinterpret_cast<const uint32_t*>(s16p + static_cast<long long>(pix + 3 * rows_par) * ld));
        v0 = unpack_op2(r0, f16); v1 = unpack_op2(r1, f16); v2 = unpack_op2(r2, f16); v3 = unpack_op2(r3, f16);
      } else {
        v0 = __ldg(reinterpret_cast<const float2*>(src + sbase + static_cast<long long>(pix) * ld));
        v1 = __ldg(reinterpret_cast<const float2*>(src + sbase + static_cast<long long>(pix + rows_par) * ld));
        v2 = __ldg(reinterpret_cast<const float2*>(src + sbase + static_cast<long long>(pix + 2 * rows_par) * ld));
        v3 = __ldg(reinterpret_cast<const float2*>(src + sbase + static_cast<long long>(pix + 3 * rows_par) * ld));
      }
      s_x[pix * hv + cv] = v0;
      s_x[(pix + rows_par) * hv + cv] = v1;
      s_x[(pix + 2 * rows_par) * hv + cv] = v2;
      s_x[(pix + 3 * rows_par) * hv + cv] = v3;
      ps += ((v0.x + v0.y) + (v1.x + v1.y)) + ((v2.x + v2.y) + (v3.x + v3.y));
    }
    for (; pix < HW; pix += rows_par) {
      const float2 v = ld2e(src, sbase + static_cast<long long>(pix) * ld, in16, f16);
      s_x[pix * hv + cv] = v;
      ps += v.x + v.y;
    }
  }
  const double cnt = static_cast<double>(HW) * cpg;
  const float mean = static_cast<float>(block_sum_double(static_cast<double>(ps), s_red) / cnt);
  // each thread re-reads exactly the elements it wrote: no barrier needed between the passes
  float pq = 0.0f;
  if (active)
    for (int pix = prow; pix < HW; pix += rows_par) {
      const float2 v = s_x[pix * hv + cv];
      const float dx = v.x - mean, dy = v.y - mean;
      pq = fmaf(dx, dx, fmaf(dy, dy, pq));
    }
  const double var = block_sum_double(static_cast<double>(pq), s_red) / cnt;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (active) {
    const float2 ga = __ldg(reinterpret_cast<const float2*>(gamma + c));
    const float2 be = __ldg(reinterpret_cast<const float2*>(beta + c));
    const float a0 = ga.x * rstd, a1 = ga.y * rstd;
    const float b0 = be.x - mean * a0, b1 = be.y - mean * a1;
    bf16* __restrict__ yo = y + static_cast<long long>(b) * HW * C + c;
    bf16* __restrict__ ro = raw != nullptr ? raw + static_cast<long long>(b) * HW * C + c : nullptr;
#pragma unroll 4
    for (int pix = prow; pix < HW; pix += rows_par) {
      const float2 v = s_x[pix * hv + cv];
      float o0 = fmaf(v.x, a0, b0), o1 = fmaf(v.y, a1, b1);
      if (silu) { o0 = silu_fast(o0); o1 = silu_fast(o1); }
      *reinterpret_cast<uint32_t*>(yo + static_cast<long long>(pix) * C) = pack_op2(o0, o1, f16);
      if (ro != nullptr) *reinterpret_cast<uint32_t*>(ro + static_cast<long long>(pix) * C) = pack_op2(v.x, v.y, f16);
    }
  }
}

constexpr int LN_MAX_VEC = 10;
// One warp per row; row held in registers (C <= 1280).  NV = float4 per lane: instantiated for the widths the U-Net
// uses so the 320-wide rows of the 64x64 level do not carry the 1280-wide register footprint (occupancy = bytes in
// flight for this pure streaming pass).
template <int NV, int IN16>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ y, int M,
                                                        int C, float eps, float* __restrict__ y32, long long ld_x, int f16) {
  constexpr int in16 = IN16;
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C / 4;
  const long long xbase = static_cast<long long>(row) * ld_x;
  float4 v[NV];
  if (in16) {
    uint2 r16[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = lane + i * 32;
      r16[i] = (k < nvec) ? ldraw4(x, xbase + 4 * k) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = cvt4(r16[i], f16);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = lane + i * 32;
      if (k < nvec) v[i] = __ldg(reinterpret_cast<const float4*>(x + xbase) + k);
    }
  }
  // packed fp32x2 arithmetic (FADD2 / FFMA2 / FMUL2): the pass is issue-bound (ncu: 85 % issue-active, 29 instructions per
  // element before), not HBM-bound, so instructions per element are what counts.  Lanes beyond the row hold zeros.
  f32x2 acc = pk2(0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int k = lane + i * 32;
    if (k >= nvec) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc = add2(acc, add2(pk2(v[i].x, v[i].y), pk2(v[i].z, v[i].w)));
  }
  float s0, s1;
  upk2(acc, s0, s1);
  float sum = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  const f32x2 nmean = pk2(-mean, -mean);
  f32x2 d[NV][2];
  f32x2 sq2 = pk2(0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int k = lane + i * 32;
    d[i][0] = add2(pk2(v[i].x, v[i].y), nmean);
    d[i][1] = add2(pk2(v[i].z, v[i].w), nmean);
    if (k < nvec) {   // (padding lanes would contribute mean^2)
      sq2 = fma2(d[i][0], d[i][0], sq2);
      sq2 = fma2(d[i][1], d[i][1], sq2);
    }
  }
  float q0, q1;
  upk2(sq2, q0, q1);
  float sq = q0 + q1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  const f32x2 rstd2 = pk2(rstd, rstd);
  const float4* gr = reinterpret_cast<const float4*>(gamma);
  const float4* br = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long long>(row) * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      const float4 g = __ldg(gr + k), bb = __ldg(br + k);
      float4 o;
      upk2(fma2(d[i][0], mul2(pk2(g.x, g.y), rstd2), pk2(bb.x, bb.y)), o.x, o.y);
      upk2(fma2(d[i][1], mul2(pk2(g.z, g.w), rstd2), pk2(bb.z, bb.w)), o.z, o.w);
      if (y != nullptr) yr[k] = make_uint2(pack_op2(o.x, o.y, f16), pack_op2(o.z, o.w, f16));
      if (y32 != nullptr) reinterpret_cast<float4*>(y32 + static_cast<long long>(row) * C)[k] = o;
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

int gn_num_launches(const GroupNormArgs& a) {
  const int C = a.C0 + a.C1;
  const bool fused = a.stats0 != nullptr && (a.C1 == 0 || a.stats1 != nullptr) && a.HW % 32 == 0;
  if (fused) return 2;
  return static_cast<size_t>(a.HW) * (C / 32) * sizeof(float) <= 96 * 1024 ? 1 : 3;
}

int launch_groupnorm(const GroupNormArgs& a, cudaStream_t stream) {
  const int C = a.C0 + a.C1;
  PBE_REQUIRE(C % 64 == 0 && C <= GN_MAX_C, "GroupNorm channels must be a multiple of 64, <= 2560");
  PBE_REQUIRE(a.C0 % 4 == 0 && a.C1 % 4 == 0, "GroupNorm concat halves must be multiples of 4 channels");
  const bool fused = a.stats0 != nullptr && (a.C1 == 0 || a.stats1 != nullptr) && a.HW % 32 == 0;
  const size_t small_smem = static_cast<size_t>(a.HW) * (C / 32) * sizeof(float);
  if (!fused && small_smem <= 96 * 1024) {
    // low-resolution levels: one kernel, the (sample, group) slice staged in shared memory
    static bool attr_set = false;
    if (!attr_set) {
      PBE_CHECK_CUDA(cudaFuncSetAttribute(gn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set = true;
    }
    const int hv = C / 64;
    const int rows_par = std::min(GNS_THREADS / hv, a.HW);
    PBE_CHECK_CUDA(launch_k(gn_small_kernel, dim3(dim3(32, a.Nb)), dim3(GNS_THREADS), small_smem, stream, a.x0, a.C0, a.x1, a.C1, a.HW, rows_par, a.gamma,
                                                                        a.beta, a.eps, a.silu, a.y, a.raw, operand_f16(), a.in16));
    PBE_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int slabs = gn_num_slabs(a.HW);
  float* partial = a.partial;
  float2* ab = reinterpret_cast<float2*>(a.partial + ((static_cast<size_t>(a.Nb) * slabs * 64 + 3) & ~static_cast<size_t>(3)));
  if (fused) {
    PBE_CHECK_CUDA(launch_k(gn_finalize_fused_kernel, dim3(dim3(32, a.Nb)), dim3(GNF_THREADS), 0, stream, a.stats0, a.C0, a.stats1, a.C1, a.HW, a.gamma,
                                                                        a.beta, a.eps, ab));
    PBE_CHECK_CUDA(cudaGetLastError());
  } else {
    PBE_CHECK_CUDA(launch_k(gn_stats_kernel, dim3(dim3(slabs, a.Nb)), dim3(GN_THREADS), 0, stream, a.x0, a.C0, a.x1, a.C1, a.HW, slabs, partial, a.in16, operand_f16()));
    PBE_CHECK_CUDA(cudaGetLastError());
    PBE_CHECK_CUDA(launch_k(gn_finalize_kernel, dim3(a.Nb), dim3(GN_THREADS), 0, stream, partial, slabs, a.HW, C, a.gamma, a.beta, a.eps, ab));
    PBE_CHECK_CUDA(cudaGetLastError());
  }
  static const bool apply8 = [] { const char* e = getenv("PBE_GN_APPLY8"); return e == nullptr || atoi(e) != 0; }();
  if (a.in16 && apply8 && a.C0 % 8 == 0 && a.C1 % 8 == 0) {
    const int vpp8 = C / 8;
    const int rp8 = vpp8 >= 256 ? 1 : 256 / vpp8;
    const int ppb = rp8 * GN8_UNROLL;
    PBE_CHECK_CUDA(launch_k(gn_apply8_kernel, dim3(dim3((a.HW + ppb - 1) / ppb, a.Nb)), dim3(vpp8 * rp8), 0, stream,
        reinterpret_cast<const unsigned short*>(a.x0), a.C0, reinterpret_cast<const unsigned short*>(a.x1), a.C1, a.HW, rp8, ab, a.silu, a.y,
        a.raw, operand_f16()));
    PBE_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int vpp = C / 4;
  const int rows_par = vpp >= 256 ? 1 : 256 / vpp;
  const int pix_per_block = rows_par * GN_UNROLL;
  if (a.in16)
    PBE_CHECK_CUDA(launch_k(gn_apply_kernel<1>, dim3(dim3((a.HW + pix_per_block - 1) / pix_per_block, a.Nb)), dim3(vpp * rows_par), 0, stream,
        a.x0, a.C0, a.x1, a.C1, a.HW, rows_par, ab, a.silu, a.y, a.raw, operand_f16()));
  else
    PBE_CHECK_CUDA(launch_k(gn_apply_kernel<0>, dim3(dim3((a.HW + pix_per_block - 1) / pix_per_block, a.Nb)), dim3(vpp * rows_par), 0, stream,
        a.x0, a.C0, a.x1, a.C1, a.HW, rows_par, ab, a.silu, a.y, a.raw, operand_f16()));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, bf16* y, int M, int C, float eps,
                     cudaStream_t stream, float* y32, long long ld_x, int in16) {
  if (ld_x == 0) ld_x = C;
  PBE_REQUIRE(ld_x % 4 == 0 && (y != nullptr || y32 != nullptr), "LayerNorm: row stride % 4, at least one output");
  PBE_REQUIRE(C % 4 == 0 && C / 4 <= 32 * LN_MAX_VEC, "LayerNorm width must be a multiple of 4, <= 1280");
  const int nv = (C / 4 + 31) / 32;
  const int f16 = operand_f16();
  const dim3 grid((M + 7) / 8), block(256);
#define PBE_LN_LAUNCH(NVV, I16) PBE_CHECK_CUDA(launch_k(layernorm_kernel<NVV, I16>, grid, block, 0, stream, x, gamma, beta, y, M, C, eps, y32, ld_x, f16))
  if (in16) {
    if (nv <= 3) PBE_LN_LAUNCH(3, 1); else if (nv <= 5) PBE_LN_LAUNCH(5, 1); else PBE_LN_LAUNCH(LN_MAX_VEC, 1);
  } else {
    if (nv <= 3) PBE_LN_LAUNCH(3, 0); else if (nv <= 5) PBE_LN_LAUNCH(5, 0); else PBE_LN_LAUNCH(LN_MAX_VEC, 0);
  }
#undef PBE_LN_LAUNCH
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pbe
