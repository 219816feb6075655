// C-ABI entry points for single operators (used by the parity tests and micro-benchmarks).
// Declared in include/pbe_b200.h.  All pointers are device pointers; calls are asynchronous on `stream`.
#include "../../include/pbe_b200.h"
#include "internal.h"

using namespace pbe;

extern "C" {

const char* pbe_last_error(void) { return get_error(); }

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
  d.block_n = block_n;
  GemmPlan plan;
  int rc = build_gemm_plan(d, &plan);
  if (rc) return rc;
  return launch_gemm_plan(plan, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
