// C-ABI entry points for single operators (used by the parity tests and micro-benchmarks).
// Declared in include/pbe_b200.h.  All pointers are device pointers; calls are asynchronous on `stream`.
#include "../../include/pbe_b200.h"
#include "internal.h"

using namespace pbe;

extern "C" {

const char* pbe_last_error(void) { return get_error(); }

static float* g_stats_out = nullptr;

int pbe_op_conv_gemm(const void* act_bf16, int Nb, int H, int W, int C, int ksize, int stride, const void* wt_bf16,
                     int Cout, int mode, const float* bias, const float* rowbias, const float* residual,
                     float* out_f32, void* out_bf16, void* out_vt, int qk_cols, int block_n, void* stream) {
  ConvGemmDesc d{};
  d.act = static_cast<const bf16*>(act_bf16);
  d.Nb = Nb; d.H = H; d.W = W; d.C = C;
  d.ksize = ksize; d.stride = stride;
  d.wt = static_cast<const bf16*>(wt_bf16);
  d.Cout = Cout;
  d.mode = mode;
  d.bias = bias; d.rowbias = rowbias; d.residual = residual;
  d.out_f32 = out_f32;
  d.out_bf16 = static_cast<bf16*>(out_bf16);
  d.out_vt = static_cast<bf16*>(out_vt);
  d.qk_cols = qk_cols;
  d.ld_out = (mode == EPI_QKV) ? qk_cols : 0;
  d.block_n = block_n;
  {  // split-K workspace for small grids (grown on demand, owned by the library; op-level API only)
    static float* ws = nullptr;
    static size_t ws_bytes = 0;
    const size_t need = gemm_splitk_ws_bytes(d);
    if (need > ws_bytes) {
      if (ws) cudaFree(ws);
      ws = nullptr;
      ws_bytes = 0;
      if (cudaMalloc(&ws, need) != cudaSuccess) { set_error("split-K workspace allocation failed"); return -2; }
      ws_bytes = need;
    }
    d.splitk_ws = need ? ws : nullptr;
  }
  if (g_stats_out != nullptr) {   // pbe_debug_set_gemm_stats_out: fused GroupNorm statistics of out_f32
    d.splitk_ws = nullptr;
    if (!gemm_can_fuse_stats(d) || out_f32 == nullptr) { set_error("fused statistics not available for this GEMM"); return -3; }
    d.stats_out = g_stats_out;
  }
  GemmPlan plan;
  int rc = build_gemm_plan(d, &plan);
  if (rc) return rc;
  return launch_gemm_plan(plan, static_cast<cudaStream_t>(stream));
}

void pbe_debug_set_gemm_stats_out(float* dev_buffer) { g_stats_out = dev_buffer; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int pbe_debug_gemm_counters(long long* out8) { return gemm_read_debug_counters(out8); }

int pbe_set_operand_format(int f16) {
  set_operand_f16(f16);
  return 0;
}
int pbe_get_operand_format(void) { return operand_f16(); }

int pbe_op_self_attention(const void* qk_bf16, const void* vt_bf16, void* out_bf16, int B, int N, int heads, int d,
                          void* stream) {
  AttnPlan plan;
  int rc = build_attn_plan(static_cast<const bf16*>(qk_bf16), static_cast<const bf16*>(vt_bf16),
                           static_cast<bf16*>(out_bf16), B, N, heads, d, &plan);
  if (rc) return rc;
  return launch_attn_plan(plan, static_cast<cudaStream_t>(stream));
}


int64_t pbe_op_groupnorm_workspace_bytes(int Nb, int HW) {
  return static_cast<int64_t>(gn_workspace_floats(Nb, HW, 2560)) * sizeof(float);
}

int pbe_op_groupnorm(const float* x0, int C0, const float* x1, int C1, int Nb, int HW, const float* gamma,
                     const float* beta, float eps, int silu, void* y_bf16, void* raw_bf16, void* workspace,
                     void* stream) {
  GroupNormArgs a{};
  a.x0 = x0; a.C0 = C0; a.x1 = x1; a.C1 = C1; a.Nb = Nb; a.HW = HW; a.gamma = gamma; a.beta = beta; a.eps = eps;
  a.silu = silu; a.y = static_cast<bf16*>(y_bf16); a.raw = static_cast<bf16*>(raw_bf16);
  a.partial = static_cast<float*>(workspace);
  return launch_groupnorm(a, static_cast<cudaStream_t>(stream));
}

int pbe_op_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16, int M, int C, float eps,
                     void* stream) {
  return launch_layernorm(x, gamma, beta, static_cast<bf16*>(y_bf16), M, C, eps, static_cast<cudaStream_t>(stream));
}

int pbe_op_small_linear(const float* x, const float* W, const float* bias, float* y, int B, int K, int O, int pre_silu,
                        int post_act, const float* residual, float* y_silu, void* stream) {
  if (x == nullptr || W == nullptr || y == nullptr) { set_error("pbe_op_small_linear: null argument"); return -1; }
  return launch_small_linear(x, W, bias, y, B, K, O, pre_silu, post_act, static_cast<cudaStream_t>(stream), y_silu, residual);
}

int pbe_op_upsample2x(const float* x, void* y_bf16, int Nb, int H, int W, int C, void* stream) {
  return launch_upsample2x_bf16(x, static_cast<bf16*>(y_bf16), Nb, H, W, C, static_cast<cudaStream_t>(stream));
}

int pbe_sampler_step(const float* eps_uc, const float* eps_c, float scale, int cfg, int order, const float* h1,
                     const float* h2, const float* h3, const float* x, float a_t, float a_prev, float sigma_t,
                     float sqrt_one_minus_at, const float* noise, float temperature, float* e_out, float* x_prev,
                     float* pred_x0, int64_t n, void* stream) {
  SamplerStepArgs a{};
  a.eps_uc = eps_uc; a.eps_c = eps_c; a.scale = scale; a.cfg = cfg; a.order = order;
  a.h1 = h1; a.h2 = h2; a.h3 = h3; a.x = x; a.a_t = a_t; a.a_prev = a_prev; a.sigma_t = sigma_t;
  a.sqrt_one_minus_at = sqrt_one_minus_at; a.noise = noise; a.temperature = temperature; a.e_out = e_out; a.x_prev = x_prev; a.pred_x0 = pred_x0;
  a.n = static_cast<size_t>(n);
  if (cfg && eps_c == nullptr) { set_error("pbe_sampler_step: cfg set but eps_c is NULL"); return -1; }
  if (order >= 1 && h1 == nullptr) { set_error("pbe_sampler_step: history missing"); return -1; }
  if ((order == 2 || order == 3) && h2 == nullptr) { set_error("pbe_sampler_step: history missing"); return -1; }
  if (order == 3 && h3 == nullptr) { set_error("pbe_sampler_step: history missing"); return -1; }
  if (sigma_t != 0.0f && noise == nullptr) { set_error("pbe_sampler_step: sigma_t != 0 needs noise"); return -1; }
  return launch_sampler_step(a, static_cast<cudaStream_t>(stream));
}

int pbe_build_unet_input(const float* x, const float* z_inpaint, const float* mask, float* out, int B, int Cx, int Cz,
                         int Cm, int HW, int dup, void* stream) {
  if (dup != 1 && dup != 2) { set_error("pbe_build_unet_input: dup must be 1 or 2"); return -1; }
  if (x == nullptr || out == nullptr || (Cz > 0 && z_inpaint == nullptr) || (Cm > 0 && mask == nullptr)) {
    set_error("pbe_build_unet_input: NULL tensor");
    return -1;
  }
  if (B <= 0 || Cx <= 0 || Cz < 0 || Cm < 0 || HW <= 0) { set_error("pbe_build_unet_input: bad extents"); return -1; }
  return launch_build_unet_input(x, z_inpaint, mask, out, B, Cx, Cz, Cm, HW, dup, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
