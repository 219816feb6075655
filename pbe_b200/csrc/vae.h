// VAE engine: AutoencoderKL.decode = Decoder(post_quant_conv(z)) and AutoencoderKL.encode = quant_conv(Encoder(x))
// (ldm/models/autoencoder.py:56-69, ldm/modules/diffusionmodules/model.py:370-580) as fixed lists of sm_100a launches
// built from the same kernels as the U-Net (implicit-GEMM convs, GroupNorm + swish, nearest 2x upsample, stride-2 conv).
// The single-head d = C mid-block attention (model.py:152-182) runs as two GEMMs per image around a row softmax.
#pragma once
#include "engine.h"

namespace pbe {

struct VaeResW {
  int cin = 0, cout = 0;
  NormW n1, n2;
  ConvW conv1, conv2, nin;
  bool has_nin = false;
};
struct VaeAttnW {
  int c = 0;
  NormW norm;
  ConvW qkv, proj_out;  // q | k | v fused: [3c, c] + bias[3c]
};

struct VaePrepared {
  int B = 0, H = 0, W = 0;
  Arena persist, scratch;
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<std::string> op_names;
  std::vector<std::string> op_family;
  std::vector<double> op_flops;
  float* z_stage = nullptr;    // input staging: latents [B, embed_dim, H, W] (decode) or images [B, in_ch, H, W] (encode)
  float* out_stage = nullptr;  // output staging: images [B, out_ch, fH, fW] (decode) or moments [B, 2*embed, H/f, W/f]
  int launches = 0;
  cudaGraphExec_t graph = nullptr;
  ~VaePrepared();
};

class VaeModel : public WeightLoader {
 public:
  explicit VaeModel(const pbe_vae_config& cfg) : cfg_(cfg) {}
  ~VaeModel() {
    prepared_.clear();
    if (cap_stream_) cudaStreamDestroy(cap_stream_);
  }
  // Repacks whichever halves have been loaded (decoder.* + post_quant_conv.*, encoder.* + quant_conv.*).
  int finalize();
  // z [B, embed_dim, H, W] fp32 NCHW (device) -> out [B, out_ch, f*H, f*W] fp32 NCHW (device), f = 2^(levels-1)
  int decode(const float* z, float* out, int B, int H, int W, cudaStream_t stream);
  // x [B, in_channels, H, W] fp32 NCHW (device) -> moments [B, 2*embed_dim, H/f, W/f] fp32 NCHW (mean | logvar)
  int encode(const float* x, float* moments, int B, int H, int W, cudaStream_t stream);
  // eager run with an event pair around every op (enc = 0: decode, 1: encode); returns the number of ops
  int profile(int enc, const float* in, float* out, int B, int H, int W, cudaStream_t stream, float* ms, int max_ops);
  const VaePrepared* current() const { return cur_; }
  std::string last_error;

 private:
  int finalize_decoder();
  int finalize_encoder();
  int make_res(const std::string& pfx, int cin, int cout, VaeResW* r);
  int make_attn(const std::string& pfx, int C, VaeAttnW* a);
  int prepare(int enc, int B, int H, int W);
  int build(VaePrepared& P, bool dry, int enc);
  int run(VaePrepared& P, const float* in, size_t in_bytes, float* out, size_t out_bytes, cudaStream_t stream, float* ms,
          int max_ops);

  pbe_vae_config cfg_;
  bool has_dec_ = false, has_enc_ = false;
  float *pq_w_ = nullptr, *pq_b_ = nullptr;   // post_quant_conv, fp32 [z_channels][embed_dim]
  ConvW conv_in_, conv_out_;
  VaeResW mid1_, mid2_;
  VaeAttnW attn_;
  std::vector<std::vector<VaeResW>> up_blocks_;   // [level][block]
  std::vector<ConvW> up_convs_;                   // [level] (unused for level 0)
  NormW norm_out_;
  // encoder half
  float *q_w_ = nullptr, *q_b_ = nullptr;         // quant_conv, fp32 [2*embed][2*z_channels]
  ConvW e_conv_in_, e_conv_out_;
  std::vector<std::vector<VaeResW>> down_blocks_;  // [level][block]
  std::vector<ConvW> down_convs_;                  // [level] (unused for the last level)
  VaeResW e_mid1_, e_mid2_;
  VaeAttnW e_attn_;
  NormW e_norm_out_;
  std::map<std::tuple<int, int, int, int>, std::unique_ptr<VaePrepared>> prepared_;
  VaePrepared* cur_ = nullptr;
  cudaStream_t cap_stream_ = nullptr;
};

}  // namespace pbe
