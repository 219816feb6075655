// VAE decoder engine: AutoencoderKL.decode = Decoder(post_quant_conv(z))
// (ldm/models/autoencoder.py:66-69, ldm/modules/diffusionmodules/model.py:474-580) as a fixed list of sm_100a launches
// built from the same kernels as the U-Net (implicit-GEMM convs, GroupNorm + swish, nearest 2x upsample).
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
  float* z_stage = nullptr;    // [B, embed_dim, H, W]
  float* out_stage = nullptr;  // [B, out_ch, 8H, 8W]
  int launches = 0;
  ~VaePrepared();
};

class VaeDecoder : public WeightLoader {
 public:
  explicit VaeDecoder(const pbe_vae_config& cfg) : cfg_(cfg) {}
  ~VaeDecoder() { prepared_.clear(); }
  int finalize();
  // z [B, embed_dim, H, W] fp32 NCHW (device) -> out [B, out_ch, f*H, f*W] fp32 NCHW (device), f = 2^(levels-1)
  int decode(const float* z, float* out, int B, int H, int W, cudaStream_t stream);
  int profile_decode(const float* z, float* out, int B, int H, int W, cudaStream_t stream, float* ms, int max_ops);
  const VaePrepared* current() const { return cur_; }
  std::string last_error;

 private:
  int prepare(int B, int H, int W);
  int build(VaePrepared& P, bool dry);

  pbe_vae_config cfg_;
  float *pq_w_ = nullptr, *pq_b_ = nullptr;   // post_quant_conv, fp32 [z_channels][embed_dim]
  ConvW conv_in_, conv_out_;
  VaeResW mid1_, mid2_;
  VaeAttnW attn_;
  std::vector<std::vector<VaeResW>> up_blocks_;   // [level][block]
  std::vector<ConvW> up_convs_;                   // [level] (unused for level 0)
  NormW norm_out_;
  std::map<std::tuple<int, int, int>, std::unique_ptr<VaePrepared>> prepared_;
  VaePrepared* cur_ = nullptr;
};

}  // namespace pbe
