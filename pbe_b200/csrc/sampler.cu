// Fused sampler update (K11): classifier-free-guidance combine + PLMS Adams-Bashforth extrapolation (or DDIM) +
// pred_x0 / x_prev, one vectorised pass over [B,4,H,W] fp32.  HBM-bound: <= 6 reads + 3 writes of the latent.
//
// Arithmetic restates, operation by operation and in the same order, PLMSSampler.p_sample_plms
// (ldm/models/diffusion/plms.py:185-189 CFG, :230-246 multistep, :202-219 x_prev) and DDIMSampler.p_sample_ddim
// (ldm/models/diffusion/ddim.py:209-242).  Every op uses the round-to-nearest intrinsics so nvcc cannot contract
// a*b+c into an FMA: PyTorch eager rounds after each op, and the parity test is bit-exact.
#include "internal.h"
#include "ptx.cuh"

namespace pbe {

namespace {

struct Coef {
  float scale;
  float sqrt_one_minus_at;
  float sqrt_at;       // a_t.sqrt()
  float sqrt_aprev;    // a_prev.sqrt()
  float dir_coef;      // (1 - a_prev - sigma_t**2).sqrt()
  float sigma_t;
  float temperature;   // noise = (sigma_t * randn) * temperature, in the reference's order (plms.py:214, ddim.py:238)
  int cfg, order;
};

__device__ __forceinline__ float step_one(float eu, float ec, float h1, float h2, float h3, float x, float nz,
                                          const Coef& k, float* e_out, float* x0_out) {
  float e = eu;
  if (k.cfg) e = __fadd_rn(eu, __fmul_rn(k.scale, __fsub_rn(ec, eu)));  // e_u + s * (e_c - e_u)
  float ep;
  switch (k.order) {
    case 0: ep = e; break;
    case 1: ep = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, e), h1), 2.0f); break;
    case 2:
      ep = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, e), __fmul_rn(16.0f, h1)), __fmul_rn(5.0f, h2)), 12.0f);
      break;
    case 3:
      ep = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.0f, e), __fmul_rn(59.0f, h1)), __fmul_rn(37.0f, h2)),
                               __fmul_rn(9.0f, h3)),
                     24.0f);
      break;
    default: ep = __fdiv_rn(__fadd_rn(h1, e), 2.0f); break;  // (e_t + e_t_next) / 2, h1 = e_t of the first eval
  }
  const float pred_x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sqrt_one_minus_at, ep)), k.sqrt_at);
  const float dir_xt = __fmul_rn(k.dir_coef, ep);
  float xp = __fadd_rn(__fmul_rn(k.sqrt_aprev, pred_x0), dir_xt);
  if (k.sigma_t != 0.0f) xp = __fadd_rn(xp, __fmul_rn(__fmul_rn(k.sigma_t, nz), k.temperature));
  *e_out = e;
  *x0_out = pred_x0;
  return xp;
}

__global__ void __launch_bounds__(256) sampler_step_kernel(SamplerStepArgs a, Coef k) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t n4 = a.n / 4;
  if (i >= n4) return;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 eu = reinterpret_cast<const float4*>(a.eps_uc)[i];
  const float4 ec = k.cfg ? reinterpret_cast<const float4*>(a.eps_c)[i] : z4;
  const float4 h1 = (k.order >= 1) ? reinterpret_cast<const float4*>(a.h1)[i] : z4;
  const float4 h2 = (k.order == 2 || k.order == 3) ? reinterpret_cast<const float4*>(a.h2)[i] : z4;
  const float4 h3 = (k.order == 3) ? reinterpret_cast<const float4*>(a.h3)[i] : z4;
  const float4 x = reinterpret_cast<const float4*>(a.x)[i];
  const float4 nz = (k.sigma_t != 0.0f && a.noise) ? reinterpret_cast<const float4*>(a.noise)[i] : z4;
  float4 e, x0, xp;
  xp.x = step_one(eu.x, ec.x, h1.x, h2.x, h3.x, x.x, nz.x, k, &e.x, &x0.x);
  xp.y = step_one(eu.y, ec.y, h1.y, h2.y, h3.y, x.y, nz.y, k, &e.y, &x0.y);
  xp.z = step_one(eu.z, ec.z, h1.z, h2.z, h3.z, x.z, nz.z, k, &e.z, &x0.z);
  xp.w = step_one(eu.w, ec.w, h1.w, h2.w, h3.w, x.w, nz.w, k, &e.w, &x0.w);
  if (a.e_out) reinterpret_cast<float4*>(a.e_out)[i] = e;
  if (a.pred_x0) reinterpret_cast<float4*>(a.pred_x0)[i] = x0;
  reinterpret_cast<float4*>(a.x_prev)[i] = xp;
}

__global__ void build_unet_input_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                        const float* __restrict__ mask, float* __restrict__ out, int B, int Cx,
                                        int Cz, int Cm, int HW, int dup) {
  griddep_enter();   // PDL: let the successor start, wait for the predecessor (ptx.cuh)
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per = static_cast<long long>(Cx + Cz + Cm) * HW;
  if (idx >= static_cast<long long>(B) * per) return;
  const int b = static_cast<int>(idx / per);
  const int r = static_cast<int>(idx % per);
  const int c = r / HW, pix = r % HW;
  float v;
  if (c < Cx) v = x[(static_cast<long long>(b) * Cx + c) * HW + pix];
  else if (c < Cx + Cz) v = z[(static_cast<long long>(b) * Cz + (c - Cx)) * HW + pix];
  else v = mask[(static_cast<long long>(b) * Cm + (c - Cx - Cz)) * HW + pix];
  out[idx] = v;
  if (dup == 2) out[static_cast<long long>(B) * per + idx] = v;
}

}  // namespace

int launch_sampler_step(const SamplerStepArgs& a, cudaStream_t stream) {
  PBE_REQUIRE(a.n % 4 == 0, "latent element count % 4");
  PBE_REQUIRE(a.order >= 0 && a.order <= 4, "order");
  Coef k;
  k.scale = a.scale;
  k.cfg = a.cfg;
  k.order = a.order;
  k.sqrt_one_minus_at = a.sqrt_one_minus_at;
  k.sigma_t = a.sigma_t;
  k.temperature = a.temperature;
  // fp32 IEEE ops, same as torch.full(..., fp32).sqrt() / (1. - a_prev - sigma_t**2).sqrt()
  k.sqrt_at = sqrtf(a.a_t);
  k.sqrt_aprev = sqrtf(a.a_prev);
  volatile float one_minus = 1.0f - a.a_prev;
  volatile float sig2 = a.sigma_t * a.sigma_t;
  volatile float inner = one_minus - sig2;
  k.dir_coef = sqrtf(inner);
  const size_t n4 = a.n / 4;
  PBE_CHECK_CUDA(launch_k(sampler_step_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, stream, a, k));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_build_unet_input(const float* x, const float* z, const float* mask, float* out, int B, int Cx, int Cz, int Cm,
                            int HW, int dup, cudaStream_t stream) {
  PBE_REQUIRE(B > 0 && Cx > 0 && Cz >= 0 && Cm >= 0 && HW > 0, "build_unet_input: empty tensor");
  const long long total = static_cast<long long>(B) * (Cx + Cz + Cm) * HW;
  PBE_CHECK_CUDA(launch_k(build_unet_input_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, x, z, mask, out, B, Cx, Cz, Cm, HW, dup));
  PBE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pbe
