/* pbe_b200 — C ABI of the B200-native Paint-by-Example denoising hot path.
 *
 * The reference (zhanwenchen/pbe) is pure Python/PyTorch and has no FFI of its own; this header DEFINES the boundary
 * a maintainer binds (ctypes stub in INTEGRATION.md).  Each entry point names the reference interface it replaces.
 * Conventions: plain pointers and sizes, no torch types; device pointers unless stated; asynchronous on `stream`
 * (a cudaStream_t passed as void*); return 0 on success, negative on error (message via pbe_last_error()); nothing
 * throws across the ABI.
 */
#ifndef PBE_B200_H_
#define PBE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Last error message of the calling thread (never NULL). */
const char* pbe_last_error(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Operator-level entry points (parity tests, micro-benchmarks)
 * ------------------------------------------------------------------------------------------------------------- */

/* Implicit-GEMM conv / linear on tcgen05.  Replaces torch conv2d / Linear calls of
 * ldm/modules/diffusionmodules/openaimodel.py:107-119,150-160,201-241 and ldm/modules/attention.py:38-65,198-230,270-297.
 *   act_bf16 : NHWC bf16 [Nb,H,W,C], C % 64 == 0        wt_bf16 : [ksize*ksize][Cout][C] bf16
 *   mode 0 (STD)  : out = acc + bias[n] + rowbias[b,n] + residual[m,n]  -> out_f32 and/or out_bf16 (row-major [M,Cout])
 *   mode 1 (GEGLU): out_bf16[m, j] = (acc_a+b_a) * gelu_erf(acc_g+b_g); weight rows interleaved per 128-col tile
 *   mode 2 (QKV)  : cols < qk_cols -> out_bf16 [M, qk_cols]; cols >= qk_cols -> out_vt [Nb][Cout-qk_cols][H*W]
 */
int pbe_op_conv_gemm(const void* act_bf16, int Nb, int H, int W, int C, int ksize, int stride, const void* wt_bf16,
                     int Cout, int mode, const float* bias, const float* rowbias, const float* residual,
                     float* out_f32, void* out_bf16, void* out_vt, int qk_cols, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PBE_B200_H_ */
