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
 *   mode 1 (GEGLU): out_bf16[m, j] = (acc_a+b_a) * gelu_erf(acc_g+b_g); weight rows interleaved per 256-col tile
 *   mode 2 (QKV)  : cols < qk_cols -> out_bf16 [M, qk_cols]; cols >= qk_cols -> out_vt [Nb][Cout-qk_cols][pitch],
 *                   pitch = H*W rounded up to a multiple of 8 (rows start 16-byte aligned; pad columns are not written)
 */
int pbe_op_conv_gemm(const void* act_bf16, int Nb, int H, int W, int C, int ksize, int stride, const void* wt_bf16,
                     int Cout, int mode, const float* bias, const float* rowbias, const float* residual,
                     float* out_f32, void* out_bf16, void* out_vt, int qk_cols, int block_n, void* stream);

/* Test / measurement aid: while dev_buffer != NULL, pbe_op_conv_gemm (mode 0, fp32 output) also writes the GroupNorm
 * statistics the engine fuses into producing GEMMs: [M/32][Cout][2] = (sum, sum of squares) of out_f32 over each run of
 * 32 consecutive output rows. Errors if the geometry cannot fuse them (split-K grids, ragged tiles). */
void pbe_debug_set_gemm_stats_out(float* dev_buffer);

/* Debug aid (env PBE_GEMM_DEBUG=1): wait-time counters of CTA 0 of the last conv_gemm launch: [0] MMA-warp cycles,
 * [1] cycles waiting for TMA stages, [2] cycles waiting for a free TMEM buffer, [4] K iterations. Host array of 8. */
int pbe_debug_gemm_counters(long long* out8);

/* Flash self-attention on tcgen05. Replaces CrossAttention.forward with context=None, ldm/modules/attention.py:207-230.
 *   qk_bf16 [B,N,2C] (Q | K), vt_bf16 [B,C,pitch] with pitch = N rounded up to a multiple of 8, out_bf16 [B,N,C];
 *   C = heads*d, scale = d^-1/2. */
int pbe_op_self_attention(const void* qk_bf16, const void* vt_bf16, void* out_bf16, int B, int N, int heads, int d,
                          void* stream);

/* GroupNorm(32 groups)(+SiLU) of the channel-concat of x0 [Nb,HW,C0] and x1 [Nb,HW,C1] (fp32 NHWC; x1 may be NULL)
 * -> y_bf16 [Nb,HW,C0+C1] (+ optional raw bf16 copy). Replaces GroupNorm32/Normalize (+SiLU, + th.cat):
 * ldm/modules/diffusionmodules/util.py:199-216, ldm/modules/attention.py:77-78, openaimodel.py:883.
 * workspace: at least pbe_op_groupnorm_workspace_bytes(Nb, HW) bytes. */
int pbe_op_groupnorm(const float* x0, int C0, const float* x1, int C1, int Nb, int HW, const float* gamma,
                     const float* beta, float eps, int silu, void* y_bf16, void* raw_bf16, void* workspace,
                     void* stream);
int64_t pbe_op_groupnorm_workspace_bytes(int Nb, int HW);

/* LayerNorm over the last dim of fp32 [M,C] -> bf16. Replaces norm1/norm3, ldm/modules/attention.py:240-242. */
int pbe_op_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16, int M, int C, float eps,
                     void* stream);

/* fp32 Linear on a handful of rows: y[B,O] = act(bias + W[O,K] . f(x[B,K])) (+ residual[B,O]), f = SiLU when pre_silu,
 * act = 0 none / 1 SiLU / 2 erf-GELU; y_silu (optional) also receives SiLU(y).  The CUDA-core GEMV behind time_embed
 * (openaimodel.py:623-628), the folded single-key cross-attention to_out(to_v(c)) (attention.py:207-230) and the one-token
 * mapper of the conditioning front-end (encoders/xf.py:31-130).  K % 4 == 0; bias / residual / y_silu may be NULL. */
int pbe_op_small_linear(const float* x, const float* W, const float* bias, float* y, int B, int K, int O, int pre_silu,
                        int post_act, const float* residual, float* y_silu, void* stream);

/* Format of the 16-bit tensor-core operands this library writes and reads (activations produced by the normalisation
 * kernels and GEMM epilogues, repacked weights, and the 16-bit tensors of the pbe_op_* entry points): 1 = fp16 (default),
 * 0 = bf16 (also: environment PBE_OPERANDS=bf16).  Same MMA rate, fp32 accumulation either way; fp16 has three more mantissa
 * bits and is what the reference runs at under torch.autocast (scripts/inference.py:301-303); conversions saturate at
 * +-65504.  Process-wide: set it before creating engines (an engine refuses to run under a format other than the one it was
 * built with).  The self-attention entry point always takes bf16 Q | K | V^T and writes its output in this format. */
int pbe_set_operand_format(int f16);
int pbe_get_operand_format(void);

/* nearest-2x upsample fp32 NHWC -> bf16 NHWC. Replaces F.interpolate in Upsample.forward, openaimodel.py:109-119. */
int pbe_op_upsample2x(const float* x, void* y_bf16, int Nb, int H, int W, int C, void* stream);

/* Fused CFG combine + PLMS/DDIM multistep + x_prev/pred_x0 update over n fp32 elements (K11).
 * Replaces plms.py:185-189,202-219,230-246 and ddim.py:209-242.  order: 0 DDIM/plain, 1..3 Adams-Bashforth with
 * h1 (most recent) .. h3, 4 = PLMS first-step average (h1 = first eps, eps_* = second evaluation).
 * eps_c may be NULL when cfg == 0; noise may be NULL when sigma_t == 0; e_out / pred_x0 may be NULL.
 * The noise term is (sigma_t * noise) * temperature, rounded in the reference's order (plms.py:214, ddim.py:238). */
int pbe_sampler_step(const float* eps_uc, const float* eps_c, float scale, int cfg, int order, const float* h1,
                     const float* h2, const float* h3, const float* x, float a_t, float a_prev, float sigma_t,
                     float sqrt_one_minus_at, const float* noise, float temperature, float* e_out, float* x_prev,
                     float* pred_x0, int64_t n, void* stream);

/* out[dup*B, Cx+Cz+Cm, H*W] = cat(x[B,Cx,HW], z_inpaint[B,Cz,HW], mask[B,Cm,HW]) repeated `dup` (1 or 2) times along the
 * batch (Paint-by-Example: 4 + 4 + 1 channels).  All three inputs must share B and HW -- the caller checks, as torch.cat would.
 * Replaces torch_cat((x, images_inpaint, images_mask), 1) and torch_cat([x]*2): plms.py:185-186,225; ddim.py:200,209. */
int pbe_build_unet_input(const float* x, const float* z_inpaint, const float* mask, float* out, int B, int Cx, int Cz,
                         int Cm, int HW, int dup, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Engine: the whole UNetModel.forward (ldm/modules/diffusionmodules/openaimodel.py:852-889) behind
 * LatentDiffusion.apply_model (ldm/models/diffusion/latent_diffusion.py:646-743)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct pbe_config {
  int32_t in_channels;     /* 9  (configs/v1.yaml:34) */
  int32_t out_channels;    /* 4 */
  int32_t model_channels;  /* 320, multiple of 64 */
  int32_t num_res_blocks;  /* 2 */
  int32_t num_levels;      /* len(channel_mult) */
  int32_t channel_mult[8]; /* 1,2,4,4 */
  int32_t num_attention_resolutions;
  int32_t attention_resolutions[8]; /* 4,2,1 */
  int32_t num_heads;                /* 8 */
  int32_t context_dim;              /* 768 */
} pbe_config;

typedef struct pbe_engine* pbe_handle;

int pbe_create(const pbe_config* cfg, pbe_handle* out);
void pbe_destroy(pbe_handle h);
/* name = reference state-dict key relative to the U-Net (e.g. "input_blocks.0.0.weight", i.e. the part after
 * "model.diffusion_model."); host fp32 data; the library repacks and owns its copy. */
int pbe_load_weight(pbe_handle h, const char* name, const float* host_data, const int64_t* shape, int rank);
/* Repack + upload everything loaded so far; errors name the first missing / mis-shaped tensor. */
int pbe_finalize_weights(pbe_handle h);
/* Fold the single-key cross-attention for context ctx [Bc, context_dim] fp32 (device): v = to_out(to_v(ctx)) per
 * SpatialTransformer (attention.py:207-230 with a 1-token context). Call whenever the context changes. */
int pbe_set_context(pbe_handle h, const float* ctx, int Bc, void* stream);
/* eps[Bc,out_ch,H,W] = UNet(x[Bc,in_ch,H,W] fp32 NCHW, t[Bc] int64, context set by pbe_set_context). */
int pbe_unet_forward(pbe_handle h, const float* x, const int64_t* t, float* eps, int Bc, int H, int W, void* stream);
/* The classifier-free-guidance evaluation of p_sample_plms / p_sample_ddim (plms.py:185-188, ddim.py:200-212):
 *   e_t_uncond, e_t = apply_model(cat([x] * 2), cat([t] * 2), cat([uc, c])).chunk(2)
 * with x [B,in_ch,H,W] and t [B] given ONCE: eps[2B,out_ch,H,W] holds the unconditional half then the conditional half;
 * the context set by pbe_set_context must have 2B rows (uc rows first).  The two halves differ only where the context
 * enters, so the layers before the first cross-attention are evaluated once for the B shared samples; the result is
 * bit-identical to pbe_unet_forward on the duplicated batch. */
int pbe_unet_forward_cfg_pair(pbe_handle h, const float* x, const int64_t* t, float* eps, int B, int H, int W, void* stream);
/* 1 = replay a captured CUDA graph per forward (default), 0 = launch kernels one by one. */
int pbe_set_use_graph(pbe_handle h, int enable);
/* Measurement aid: one eager forward with a CUDA-event pair around every op of the launch plan. Fills ms_out[i]
 * (device time of op i, ms) and returns the number of ops (<0 on error). pbe_op_info describes op i of the current
 * shape: name, kernel family, algorithmic FLOPs (2*MAC, unpadded) and algorithmic HBM bytes. */
int pbe_profile_forward(pbe_handle h, const float* x, const int64_t* t, float* eps, int Bc, int H, int W, void* stream,
                        float* ms_out, int max_ops);
int pbe_op_info(pbe_handle h, int i, const char** name, const char** family, double* flops, double* bytes);
/* number of kernels one pbe_unet_forward launches for the current shape (0 before the first forward). */
int pbe_launches_per_forward(pbe_handle h);

/* ---------------------------------------------------------------------------------------------------------------
 * VAE decode (first "next" row of SURVEY.md 8f): AutoencoderKL.decode = Decoder(post_quant_conv(z)),
 * ldm/models/autoencoder.py:66-69 + ldm/modules/diffusionmodules/model.py:474-580, reached from
 * LatentDiffusion.decode_first_stage (ldm/models/diffusion/latent_diffusion.py:444-508) after the sampling loop.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct pbe_vae_config {
  int32_t embed_dim;      /* 4  (configs/v1.yaml:51) */
  int32_t z_channels;     /* 4 */
  int32_t ch;             /* 128, multiple of 64 */
  int32_t out_ch;         /* 3 */
  int32_t num_levels;     /* len(ch_mult) */
  int32_t ch_mult[8];     /* 1,2,4,4 */
  int32_t num_res_blocks; /* 2 (the decoder runs num_res_blocks + 1 blocks per level) */
  int32_t in_channels;    /* 3  (encoder input; double_z = true: the encoder emits 2 * z_channels moment channels) */
} pbe_vae_config;

typedef struct pbe_vae* pbe_vae_handle;

int pbe_vae_create(const pbe_vae_config* cfg, pbe_vae_handle* out);
void pbe_vae_destroy(pbe_vae_handle h);
/* name = reference state-dict key relative to the autoencoder ("decoder.conv_in.weight", "post_quant_conv.bias",
 * "encoder.down.0.block.0.norm1.weight", "quant_conv.weight", ..., i.e. the part after "first_stage_model."); host fp32
 * data. Either half may be left out: pbe_vae_finalize_weights prepares the halves that are complete, and the entry
 * point of a missing half fails. */
int pbe_vae_load_weight(pbe_vae_handle h, const char* name, const float* host_data, const int64_t* shape, int rank);
int pbe_vae_finalize_weights(pbe_vae_handle h);
/* out[B,out_ch,fH,fW] = decode(z[B,embed_dim,H,W]) (fp32 NCHW, device), f = 2^(num_levels-1); H*W % 64 == 0. */
int pbe_vae_decode(pbe_vae_handle h, const float* z, float* out, int B, int H, int W, void* stream);
/* moments[B, 2*embed_dim, H/f, W/f] = quant_conv(Encoder(x[B,in_channels,H,W])) (fp32 NCHW, device):
 * AutoencoderKL.encode, autoencoder.py:56-64, before DiagonalGaussianDistribution (mean | logvar along dim 1);
 * reached from LatentDiffusion.encode_first_stage (latent_diffusion.py:571-610). H, W multiples of 8f. */
int pbe_vae_encode(pbe_vae_handle h, const float* x, float* moments, int B, int H, int W, void* stream);
/* Measurement aid, as pbe_profile_forward / pbe_op_info (encode = 0: decode, 1: encode). */
int pbe_vae_profile(pbe_vae_handle h, int encode, const float* in, float* out, int B, int H, int W, void* stream,
                    float* ms_out, int max_ops);
/* Post-processing of decoded images, scripts/inference.py:346-348,379-380 (second "next" row, device part):
 * out_u8[b,h,w,c] = (uint8) (255 * clamp((img[b,c,h,w] + 1) / 2, 0, 1)), fp32 NCHW -> uint8 NHWC, same fp32 rounding
 * and truncation as the reference's clamp -> numpy -> astype(uint8) sequence. */
int pbe_postprocess_u8(const float* img, uint8_t* out_u8, int B, int C, int H, int W, void* stream);
/* Pre-processing of an edit request (third "next" row of SURVEY.md 8f, device part); all buffers on the device, uint8 images
 * in the HWC layout PIL / numpy hand over, fp32 outputs NCHW.  Bit-identical to the reference's fp32 op sequences.
 *  - pbe_normalize_u8: torchvision ToTensor + Normalize(mean, std) as built by get_tensor() / get_tensor_clip()
 *    (scripts/inference.py:106-124, ldm/data/test_bench_dataset.py:37-61): out[b,c,h,w] = (u8[b,h,w,c]/255 - mean[c]) / std[c];
 *    mean3 / std3 are HOST arrays of three floats.
 *  - pbe_prepare_inpaint_u8: mask = 1 - m/255, thresholded to {0, 1} at 0.5 when binarize != 0 (scripts/inference.py:311-317;
 *    binarize = 0 is ldm/data/test_bench_dataset.py:89-92), image = get_tensor()(img), inpaint = image * mask (:318, :98).
 *    image_out [B,3,H,W] and mask_out [B,1,H,W] may be NULL; inpaint_out [B,3,H,W] is required.
 *  - pbe_resize_bilinear: torchvision Resize([h, w]) of a float tensor [NC,H,W] -> [NC,h,w] (scripts/inference.py:332, the
 *    latent-resolution mask) = F.interpolate(bilinear, align_corners=False); antialias = 0 is the reference's pinned
 *    torchvision 0.12 behaviour for tensors, antialias = 1 the default of torchvision >= 0.17. */
int pbe_normalize_u8(const uint8_t* img_u8, float* out, int B, int H, int W, const float* mean3, const float* std3, void* stream);
int pbe_prepare_inpaint_u8(const uint8_t* img_u8, const uint8_t* mask_u8, int B, int H, int W, int binarize, float* image_out,
                           float* mask_out, float* inpaint_out, void* stream);
int pbe_resize_bilinear(const float* in, float* out, int NC, int H, int W, int h, int w, int antialias, void* stream);
int pbe_vae_op_info(pbe_vae_handle h, int i, const char** name, const char** family, double* flops);
int pbe_vae_launches_per_decode(pbe_vae_handle h);

/* ---------------------------------------------------------------------------------------------------------------
 * Conditioning front-end (second "next" row of SURVEY.md 8f): FrozenCLIPImageEmbedder.forward,
 * ldm/modules/encoders/modules.py:138-171 = CLIP ViT vision tower (transformers CLIPVisionModel, pooler_output)
 * -> unsqueeze(1) -> mapper (ldm/modules/encoders/xf.py Transformer(1, width, 5, 1)) -> final_ln.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct pbe_clip_config {
  int32_t image_size;     /* 224 */
  int32_t patch_size;     /* 14 */
  int32_t width;          /* 1024, multiple of 64 */
  int32_t layers;         /* 24 */
  int32_t heads;          /* 16 (head dim = width / heads <= 64) */
  int32_t mlp_dim;        /* 4096 */
  int32_t mapper_layers;  /* 5 */
} pbe_clip_config;

typedef struct pbe_clip* pbe_clip_handle;

int pbe_clip_create(const pbe_clip_config* cfg, pbe_clip_handle* out);
void pbe_clip_destroy(pbe_clip_handle h);
/* name = state-dict key relative to the embedder ("transformer.vision_model.encoder.layers.0.mlp.fc1.weight",
 * "mapper.resblocks.0.attn.c_qkv.weight", "final_ln.bias", ..., i.e. the part after "cond_stage_model."). */
int pbe_clip_load_weight(pbe_clip_handle h, const char* name, const float* host_data, const int64_t* shape, int rank);
int pbe_clip_finalize_weights(pbe_clip_handle h);
/* z[B, 1, width] = final_ln(mapper(pooler_output(image[B, 3, image_size, image_size]))) (fp32, device). */
int pbe_clip_encode(pbe_clip_handle h, const float* image, float* z, int B, void* stream);
int pbe_clip_launches_per_encode(pbe_clip_handle h);

#ifdef __cplusplus
}
#endif
#endif /* PBE_B200_H_ */
