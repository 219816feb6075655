// Conditioning front-end engine: FrozenCLIPImageEmbedder.forward (ldm/modules/encoders/modules.py:138-171):
//   pooled = CLIPVisionModel(image).pooler_output        (transformers, pinned 4.19.2 in the reference's environment.yaml:27;
//                                                          CLIPVisionTransformer: patch conv + class / position embeddings,
//                                                          pre_layrnorm, N x [LN, MHA, LN, quick_gelu MLP], post_layernorm of CLS)
//   z = final_ln(mapper(pooled.unsqueeze(1)))            (ldm/modules/encoders/xf.py:107-130: 5 residual blocks on ONE token,
//                                                          so attention(x) = c_proj(v(x)): softmax over a single key is 1)
// The ViT tower runs on the U-Net's kernels (tcgen05 GEMMs with bias / residual / quick_gelu / Q|K,V^T epilogues, flash
// attention, LayerNorm); the single-token mapper runs as fp32 GEMVs.
#pragma once
#include "engine.h"

namespace pbe {

struct ClipLayerW {
  NormW ln1, ln2;
  ConvW qkv, out_proj, fc1, fc2;   // qkv: [3C, C] (q | k | v) + bias[3C]
};
struct ClipMapW {
  NormW ln1, ln2;
  float *wv = nullptr, *bv = nullptr;        // rows [2C, 3C) of attn.c_qkv (the value projection), fp32
  float *wproj = nullptr, *bproj = nullptr;  // attn.c_proj
  float *wfc = nullptr, *bfc = nullptr;      // mlp.c_fc  [4C, C]
  float *wfc2 = nullptr, *bfc2 = nullptr;    // mlp.c_proj [C, 4C]
};

struct ClipPrepared {
  int B = 0;
  Arena persist;
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<std::string> op_names;
  float* img_stage = nullptr;  // [B, 3, S, S]
  float* z_stage = nullptr;    // [B, C]
  int launches = 0;
  cudaGraphExec_t graph = nullptr;
  ~ClipPrepared() {
    if (graph) cudaGraphExecDestroy(graph);
    if (persist.base_) cudaFree(persist.base_);
  }
};

class ClipEncoder : public WeightLoader {
 public:
  explicit ClipEncoder(const pbe_clip_config& cfg) : cfg_(cfg) {}
  ~ClipEncoder() {
    prepared_.clear();
    if (cap_stream_) cudaStreamDestroy(cap_stream_);
  }
  int finalize();
  int encode(const float* image, float* z, int B, cudaStream_t stream);
  int launches() const { return cur_ ? cur_->launches : 0; }
  std::string last_error;

 private:
  int prepare(int B);
  int build(ClipPrepared& P, bool dry);
  int make_linear(const std::string& prefix, int cin, int cout, ConvW* w) { return make_conv(prefix, 1, cin, cout, w); }
  int upload_named(const std::string& name, size_t expect, float** dst, size_t row0 = 0, size_t rows = 0, size_t cols = 0);

  pbe_clip_config cfg_;
  bf16* patch_w_ = nullptr;   // [C][Kpad]
  int kpad_ = 0;
  float *cls_ = nullptr, *pos_ = nullptr;
  NormW pre_ln_, post_ln_, final_ln_;
  std::vector<ClipLayerW> layers_;
  std::vector<ClipMapW> mapper_;
  std::map<int, std::unique_ptr<ClipPrepared>> prepared_;
  ClipPrepared* cur_ = nullptr;
  cudaStream_t cap_stream_ = nullptr;
};

}  // namespace pbe
