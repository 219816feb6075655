// Internal host-side API shared by the kernels, the engine and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

namespace pbe {

typedef __nv_bfloat16 bf16;   // "a 16-bit GEMM operand": the bits are fp16 or bf16 depending on operand_f16() below

// Format of the 16-bit tensor-core operands (activations written by the norm kernels / GEMM epilogues, repacked weights):
// fp16 (default) or bf16 (PBE_OPERANDS=bf16 or pbe_set_operand_format(0)).  Both run kind::f16 MMAs at the same rate with
// fp32 accumulation; fp16 carries three more mantissa bits (operand rounding 2^-12 instead of 2^-9: the U-Net's eps error
// against the fp32 reference drops from 9.8e-3 to 1.8e-3, tools/parity_attribution.py) and is the precision the reference
// itself runs at under torch.autocast (scripts/inference.py:301-303); conversions saturate at +-65504 instead of
// overflowing.  The flash-attention kernels keep bf16 Q / K / V / P (P is kept 2^-64 below 1, outside the fp16 range).
// Process-wide; engines record the format they were built with and refuse to run under another.
int operand_f16();
void set_operand_f16(int f16);

// Thread-local error string (also copied into the handle by the C-ABI layer).
void set_error(const std::string& msg);
const char* get_error();

#define PBE_CHECK_CUDA(expr)                                                                          \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      ::pbe::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + \
                       ":" + std::to_string(__LINE__));                                               \
      return -2;                                                                                      \
    }                                                                                                 \
  } while (0)

#define PBE_REQUIRE(cond, msg)                                                              \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ::pbe::set_error(std::string("requirement failed: ") + #cond + " — " + (msg) + " at " + \
                       __FILE__ + ":" + std::to_string(__LINE__));                          \
      return -1;                                                                            \
    }                                                                                       \
  } while (0)

// ------------------------------------------------------------------------------------------------
// TMA tensor maps (cuTensorMapEncodeTiled resolved through the runtime; no -lcuda link dependency)
// ------------------------------------------------------------------------------------------------
// dims[0] is the innermost (contiguous) dimension. strides_bytes has rank-1 entries (for dims 1..rank-1).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128);
int make_tmap(CUtensorMap* out, const void* base, bool is_f32, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// ------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution / linear on tcgen05 (gemm_tc.cu)
//   out[m, n] = epilogue( sum_{tap, c} act[pixel(m) shifted by tap, c] * wt[tap][n][c] )
// ------------------------------------------------------------------------------------------------
enum EpiMode : int { EPI_STD = 0, EPI_GEGLU = 1, EPI_QKV = 2 };

struct ConvGemmParams {
  // M-tile geometry: 128 rows = tn x th x tw output pixels (w fastest)
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_n;
  int Wo, Ho, Nb;          // output extents
  int num_taps, k_chunks;  // K loop = num_taps * k_chunks chunks of 64 channels
  int n_total;             // number of GEMM columns (Cout)
  int n_tiles;             // ceil(n_total / BLOCK_N)
  int has_res, has_o32, has_o16;  // which epilogue tensor maps are live
  int debug;                      // 1: CTA 0 records wait-time counters (PBE_GEMM_DEBUG)
  int pair_dim;                   // 2-CTA clusters: tile dimension (0 w, 1 h, 2 n) whose neighbours form a cluster; -1 none
  int split_k;                    // >1: K range split over work units, partial sums to a workspace
  float* stats_out;               // optional [M/32][n_total][2] per-32-row (sum, sumsq) of the fp32 output (GroupNorm)
  int stats_sample_extra, stats_block_off;   // sub-pixel phases: row-block index += sample * extra + off (full-resolution layout)
  int8_t tap_dw[9], tap_dh[9], tap_ph[9];
  int tap_coff[9];
  // epilogue
  int mode;
  const float* bias;      // [n_total] or null
  const float* rowbias;   // [Nb, rowbias_ld] or null (timestep-embedding / folded cross-attention term)
  int rowbias_ld;
  const float* residual;  // [M, ld_out] fp32 or null
  int res16;              // 1: `residual` points to 16-bit data (operand format) -- the 16-bit residual stream of the U-Net
  float* out_f32;         // [M, ld_out] or null
  bf16* out_bf16;         // [M, ld_out] or null
  int ld_out;
  // EPI_QKV: columns [0, qk_cols) go to out_bf16 (ld_out), columns >= qk_cols go transposed to out_vt[b][c][token]
  bf16* out_vt;
  int qk_cols;
  int vt_tokens;          // EPI_QKV: tokens per sample of the V^T store when the activation is one flat row range (0: H*W)
  int act;                // EPI_STD: 0 none, 1 quick_gelu x * sigmoid(1.702 x) (CLIP MLP), applied after the biases
  // LayerNorm folded around the GEMMs of a transformer block (engine.cu): the GEMM that WRITES a row-normalised tensor emits
  // per-row partial (sum, sum of squares) of its rounded 16-bit output, one float2 per (n tile, chunk parity) -> ln_stats_out
  // [2 * n_tiles][ln_stats_stride]; the GEMM that READS it multiplies the raw tensor by gamma-scaled, row-CENTRED weights
  // (sum over k of W'[n][k] = 0, so x W'^T = (x - mean) W'^T: the mean needs no term of its own) and finishes LayerNorm per
  // row in its epilogue: out = rstd * acc + bias'[n]  (attention.py:270-276 norm1 / norm3)
  float2* ln_stats_out; long long ln_stats_stride;
  const float2* ln_in; long long ln_stride; int ln_parts; float ln_inv_c, ln_eps;
  int bias_cols;          // `bias` applies to columns [0, bias_cols) only (n tiles beyond it skip the bias machinery)
  uint32_t idesc_fmt;     // operand-format bits of the MMA instruction descriptor (bf16: A and B format 1; fp16: 0)
  int out16_f16;          // 16-bit outputs (out_bf16 / out_vt) are written as fp16 (else bf16)
};

// Deterministic split-K: slice s of the K range writes its partial tile to ws[s]; a second small kernel sums the
// slices in fixed order and applies the fused epilogue terms.
struct SplitKReduce {
  int out16_f16;                                   // format of out_bf16
  const float* ws; int S; long long slice_stride;  // floats between slices
  long long M; int N; int HW;                      // rows, columns, rows per sample
  const float* bias; const float* rowbias; int rowbias_ld; const float* residual;
  const bf16* residual16;                          // 16-bit residual (operand format, same ld_out) instead of `residual`
  float* out_f32; bf16* out_bf16; int ld_out;
};

struct GemmPlan {
  SplitKReduce red;
  CUtensorMap tmA, tmB;
  CUtensorMap tmR, tmO32, tmO16;  // residual (fp32 load), fp32 output store, bf16 output store (epilogue TMA)
  ConvGemmParams p;
  int block_n;
  int cluster;  // CTAs per cluster (1 or 2)
  dim3 grid;
  size_t smem;
};

// Kernel launch, optionally (small-batch U-Net plans, or PBE_PDL=1) with programmatic dependent launch: consecutive kernels of a stream (and of
// the captured CUDA graph) then overlap the successor's prologue with the predecessor's tail; every kernel of this
// library calls griddep_wait() (ptx.cuh) before it touches memory its predecessor may still be writing.
bool pdl_enabled();
void pdl_suppress(bool on);   // true: launches on this thread go without PDL whatever the scope / environment says (run_op_list)
void pdl_set_scope(int v);   // 1 / 0: the launches (and graph captures) that follow on this thread use / do not use PDL; -1: default
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct ConvGemmDesc {
  const bf16* act;  // NHWC activations [Nb, H, W, C]
  int Nb, H, W, C;  // input geometry (C % 64 == 0)
  int c_real;       // channels that carry data (0 -> C); only used for FLOP accounting
  int ksize;        // 1 or 3 (pad = ksize/2)
  int stride;       // 1 or 2
  const bf16* wt;   // [ksize*ksize][Cout][C]
  int Cout;
  int mode;
  const float* bias;
  const float* rowbias;
  int rowbias_ld;  // 0 -> Cout
  const float* residual;
  const bf16* residual16;  // 16-bit residual in the operand format (the U-Net's 16-bit residual stream); excludes `residual` / out_f32
  float* out_f32;
  bf16* out_bf16;
  int ld_out;  // 0 -> Cout (or Cout/2 for GEGLU)
  int res_ld;  // elements between consecutive rows of the residual (0 -> the output's leading dimension): an output that is a
               // column slice of a wider tensor next to a contiguous residual (no split-K)
  bf16* out_vt;
  int qk_cols;
  int block_n;  // 0 -> auto
  float* splitk_ws;        // optional workspace enabling split-K (size from gemm_splitk_ws_bytes)
  float* stats_out;        // optional fused GroupNorm statistics of out_f32 (see gemm_can_fuse_stats)
  int act_ld;              // elements between consecutive pixels of `act` (0 -> C): an operand that is a column slice
  int wt_ld;               // elements between consecutive output rows of `wt` (0 -> C; ksize 1 only)
  int vt_tokens;           // EPI_QKV over a flat [1,1,M,C] activation: tokens per sample (0: H*W and sample = n)
  int epi_act;             // EPI_STD: 0 none, 1 quick_gelu (applied after bias / rowbias, before the residual)
  int split_batch;         // batch the split-K heuristic should assume (0 -> Nb): a CFG-pair prefix GEMM over B samples
                           // must accumulate in the same order as the same layer over the full 2B batch
  int pad_end;             // stride-2 3x3 only: 1 = zero padding (0,1,0,1) as the VAE Downsample (model.py:74-76)
                           // instead of the symmetric padding 1 of the U-Net Downsample
  int up_phase;            // 0: ordinary conv.  1..4: sub-pixel phase (a, b) = ((up_phase-1) >> 1, (up_phase-1) & 1) of
                           // "nearest-2x upsample, then 3x3 conv" (openaimodel.py:109-119): ksize = 2, `act` is the LOW-resolution
                           // tensor [Nb,H,W,C], `wt` the phase's four combined taps [4][Cout][C], and the outputs / statistics
                           // address the FULL-resolution tensor [Nb,2H,2W,Cout] at pixels (2i+a, 2j+b).  4/9 of the FLOPs of
                           // the literal form, no upsampled copy of the input.
  // LayerNorm fold (see ConvGemmParams): producer side -- per-row partial statistics of the 16-bit output; forces split_k = 1
  float2* ln_stats_out; long long ln_stats_stride;   // [2 * n_tiles][ln_stats_stride] (stride in rows; >= rows of this GEMM)
  // consumer side -- `act` is the raw (un-normalised) tensor, `wt` carries gamma, `bias` carries beta (and the layer's own bias)
  const float2* ln_in; long long ln_stride; int ln_parts; float ln_eps;
  int bias_cols;           // 0 -> Cout: `bias` applies to columns [0, bias_cols) only (must be a multiple of the n tile)
  int out16_bf16;          // 1: the 16-bit outputs are bf16 whatever the operand format (Q | K | V^T read by the flash kernels)
  int operands_bf16;       // 1: act / wt are bf16 whatever the operand format
};
// Split factor build_gemm_plan will use for this problem when a workspace is supplied (1 = no split), and its size.
int gemm_read_debug_counters(long long* out8);
int gemm_reset_debug_counters();
int gemm_read_debug_counters16(long long* out16);
int gemm_split_k(const ConvGemmDesc& d);
// True when the epilogue can emit per-32-row column statistics for this geometry (no split-K, aligned tiles).
bool gemm_can_fuse_stats(const ConvGemmDesc& d);
// Number of per-row partials (2 * n tiles) a GEMM with ln_stats_out writes, or 0 when this GEMM cannot emit them
// (needs the plain epilogue, a 16-bit-only output, full n tiles of at least two 32-column chunks).
int gemm_ln_parts(const ConvGemmDesc& d);
size_t gemm_splitk_ws_bytes(const ConvGemmDesc& d);

int build_gemm_plan(const ConvGemmDesc& d, GemmPlan* plan);
int launch_gemm_plan(const GemmPlan& plan, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Flash self-attention on tcgen05 (attn_tc.cu)
//   qk: [B, N, 2C] bf16 (Q in cols [0,C), K in cols [C,2C)), vt: [B, C, N] bf16, out: [B, N, C] bf16
// ------------------------------------------------------------------------------------------------
struct AttnPlan {
  CUtensorMap tmQ, tmK, tmV;
  int B, N, heads, d;
  float scale_log2;
  bf16* out;
  dim3 grid;
  size_t smem;
  int out_f16;  // output format: the library's operand format at plan-build time
  int* flags;   // per-item overflow flags of the single-pass softmax (attn_tc.cu); null = no exact re-run
};
// Row pitch (elements) of the transposed V buffer [B][C][vt_pitch(N)]: token counts that are not multiples of 8 are padded
// so that every row starts 16-byte aligned (TMA global stride requirement); the pad columns are never read.
inline int vt_pitch(int N) { return (N + 7) & ~7; }
// own_flags: give the plan its own overflow-flag slice (long-lived engine plans); false = the shared slice (one-off plans)
int build_attn_plan(const bf16* qk, const bf16* vt, bf16* out, int B, int N, int heads, int d, AttnPlan* plan,
                    bool own_flags = false);
int launch_attn_plan(const AttnPlan& plan, cudaStream_t stream);
int attn_num_launches(const AttnPlan& plan);   // 2 with the exact re-run of overflowed items behind the single-pass kernel

// ------------------------------------------------------------------------------------------------
// Normalisation / elementwise kernels (norm.cu, misc.cu)
// ------------------------------------------------------------------------------------------------
// GroupNorm(32 groups) over NHWC fp32 input that may be the channel-concat of two tensors (x0: C0 ch, x1: C1 ch).
// Writes y = [silu](gn(x)) as bf16 [M, C0+C1]; optionally also the raw concat as bf16 (for 1x1 skip convs).
struct GroupNormArgs {
  const float* x0; int C0;
  const float* x1; int C1;   // x1 may be null (C1 = 0)
  int in16;                  // 1: x0 / x1 point to 16-bit data in the operand format (the U-Net's 16-bit residual stream)
  int Nb, HW;
  const float* gamma; const float* beta;
  float eps; int silu;
  bf16* y; bf16* raw;        // raw may be null; 16-bit operands in the library's operand format (operand_f16())
  float* partial;            // workspace of gn_workspace_floats(Nb, HW, C0+C1) floats (16-B aligned)
  // optional statistics fused into the producers' GEMM epilogues: [Nb*HW/32][C_i][2]; when every live source has
  // them the stats pass over x is skipped
  const float* stats0; const float* stats1;
};
int gn_num_slabs(int HW);
int gn_workspace_floats(int Nb, int HW, int C);
int gn_num_launches(const GroupNormArgs& a);  // kernels launch_groupnorm will launch for these arguments
int launch_groupnorm(const GroupNormArgs& a, cudaStream_t stream);

// LayerNorm over the last dim of fp32 [M, C] -> bf16 [M, C]
// y (bf16) and / or y32 (fp32), row r of x at x + r * ld_x (ld_x = 0: contiguous rows of C)
// in16: x points to 16-bit rows in the operand format
int launch_layernorm(const float* x, const float* gamma, const float* beta, bf16* y, int M, int C, float eps,
                     cudaStream_t stream, float* y32 = nullptr, long long ld_x = 0, int in16 = 0);

// nearest 2x upsample, fp32 NHWC [Nb,H,W,C] -> bf16 NHWC [Nb,2H,2W,C]
int launch_upsample2x_bf16(const float* x, bf16* y, int Nb, int H, int W, int C, cudaStream_t stream);
// the same from a 16-bit source (a pure copy: [Nb,H,W,C] 16-bit -> [Nb,2H,2W,C] 16-bit)
int launch_upsample2x_16(const bf16* x, bf16* y, int Nb, int H, int W, int C, cudaStream_t stream);
// images fp32 NCHW in [-1, 1] -> uint8 NHWC (clamp((x+1)/2) * 255, truncated)
int launch_postprocess_u8(const float* x, uint8_t* out, int Nb, int C, int H, int W, cudaStream_t stream);
int launch_normalize_u8(const uint8_t* in, float* out, int Nb, int H, int W, const float* mean, const float* stdv,
                        cudaStream_t stream);
int launch_prepare_inpaint_u8(const uint8_t* img, const uint8_t* mask, float* image_out, float* mask_out, float* inpaint_out,
                              int Nb, int H, int W, int binarize, cudaStream_t stream);
int launch_resize_bilinear(const float* in, float* out, int NC, int H, int W, int h, int w, int antialias, cudaStream_t stream);
// VAE encode tail: quant_conv (1x1, fp32) fused with the NHWC -> NCHW unpack of the moments
int launch_vae_unpack_moments(const float* y, const float* Wq, const float* bq, float* out, int Nb, int Cin, int Cout,
                              int H, int W, int ld, cudaStream_t stream);
// VAE decode helpers: post_quant_conv fused with the NCHW -> NHWC bf16 pack; row softmax fp32 -> bf16
int launch_vae_pack_input(const float* z, const float* Wp, const float* bp, bf16* y, int Nb, int Cin, int Cz, int H, int W,
                          int Cpad, cudaStream_t stream);
int launch_softmax_rows(const float* S, bf16* P, long long rows, int N, float scale, cudaStream_t stream);
// fp32 -> bf16 cast of a flat buffer
int launch_cast_bf16(const float* x, bf16* y, size_t n, cudaStream_t stream);

// C[M, N] = A[M, K] . B[K, N] (row-major fp32, fp64 accumulation): weight products composed at load time
int launch_matmul_f64acc(const float* A, const float* B, float* C, int M, int K, int N, cudaStream_t stream);

// x NCHW fp32 [Nb, Cin, H, W] -> NHWC bf16 [Nb, H, W, Cpad] (channels >= Cin zero)
int launch_pack_input(const float* x, bf16* y, int Nb, int Cin, int H, int W, int Cpad, cudaStream_t stream);
// y NHWC fp32 [Nb, H, W, ld] (first Cout channels) -> NCHW fp32 [Nb, Cout, H, W]
int launch_unpack_output(const float* y, float* out, int Nb, int Cout, int H, int W, int ld, cudaStream_t stream);

// Small dense layers on CUDA cores (M = batch rows only): y[b, o] = act_in(x[b, :]) . W[o, :] + bias[o]
// pre_silu applies SiLU to x on load. W is fp32 [O, K] row-major.
// y_silu (optional) additionally receives silu(y) so consumers that all start with SiLU apply it once.
// post_act: 0 none, 1 SiLU, 2 GELU (erf); residual (optional, [B, O]) is added after the activation
int launch_small_linear(const float* x, const float* W, const float* bias, float* y, int B, int K, int O, int pre_silu,
                        int post_act, cudaStream_t stream, float* y_silu = nullptr, const float* residual = nullptr);
// CLIP ViT front-end helpers (transformers CLIPVisionEmbeddings): patches [B,3,H,W] fp32 -> [B*P, Kpad] bf16 rows in
// (c, kh, kw) order; tokens[b,0] = class + pos[0], tokens[b,1+p] = patch_embed[b,p] + pos[1+p]
int launch_clip_pack_patches(const float* img, bf16* out, int B, int H, int W, int patch, int Kpad, cudaStream_t stream);
int launch_clip_embed(const float* patch_emb, const float* cls, const float* pos, float* tokens, int B, int P, int C,
                      cudaStream_t stream);
// sinusoidal timestep embedding [B, dim] (cos || sin), reference util.py:151-171
int launch_timestep_embedding(const int64_t* t, float* out, int B, int dim, cudaStream_t stream);
// y[b, n] = a[n] + v[b, n]
int launch_add_rowvec(const float* a, const float* v, float* y, int B, int N, cudaStream_t stream);

// Fused sampler update (sampler.cu): CFG combine + PLMS/DDIM multistep + x_prev / pred_x0.
struct SamplerStepArgs {
  const float* eps_uc;   // [B, 4, H, W] NCHW (unconditional half) — or the only eps when cfg == 0
  const float* eps_c;    // conditional half (null when cfg == 0)
  float scale;           // guidance scale
  int cfg;
  int order;             // 0: e' = e ; 1: (3e - h1)/2 ; 2: (23e-16h1+5h2)/12 ; 3: (55e-59h1+37h2-9h3)/24 ;
                         // 4: e' = (e_prev_first + e)/2  (PLMS first-step second evaluation, h1 = first eps)
  const float* h1; const float* h2; const float* h3;
  const float* x;        // current latent [B,4,H,W]
  float a_t, a_prev, sigma_t, sqrt_one_minus_at;
  const float* noise;    // optional [B,4,H,W], multiplied by sigma_t (null when sigma_t == 0)
  float temperature;     // the noise term is (sigma_t * noise) * temperature
  float* e_out;          // CFG-combined eps (history entry); may be null
  float* x_prev;         // may alias nothing else
  float* pred_x0;        // may be null
  size_t n;              // elements
};
int launch_sampler_step(const SamplerStepArgs& a, cudaStream_t stream);
// out[b] = cat(x[b] (Cx ch), z[b] (Cz ch), mask[b] (Cm ch)) duplicated for both CFG halves: out [2B or B, Cx+Cz+Cm, H, W]
int launch_build_unet_input(const float* x, const float* z, const float* mask, float* out, int B, int Cx, int Cz, int Cm,
                            int HW, int dup, cudaStream_t stream);

}  // namespace pbe
