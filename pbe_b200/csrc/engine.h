// U-Net engine: owns repacked weights, per-shape launch plans and workspaces; runs UNetModel.forward
// (ldm/modules/diffusionmodules/openaimodel.py:852-889) as a fixed list of sm_100a kernel launches (CUDA-graph replayed).
#pragma once
#include <stdlib.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pbe_b200.h"
#include "internal.h"

namespace pbe {

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

struct ConvW {
  bf16* wup = nullptr; // Upsample convs: the four sub-pixel phases' combined 2x2 taps [phase][4][cout][cin] (ConvGemmDesc::up_phase)
  bf16* w = nullptr;   // [k*k][cout][cin_pad]
  float* b = nullptr;  // [cout]
  int cin = 0, cin_pad = 0, cout = 0, k = 1;
};
struct NormW {
  float* g = nullptr;
  float* b = nullptr;
  int c = 0;
};
struct ResW {
  int cin = 0, cout = 0, emb_off = 0;
  NormW gn1, gn2;
  ConvW conv1, conv2, skip;
  bool has_skip = false;
};
struct STW {
  int c = 0, heads = 0, d = 0;
  NormW gn, ln1, ln3;
  ConvW proj_in, qkv, to_out, ff1, ff2, proj_out;
  ConvW ffproj;          // proj_out o ff.net.2 as ONE GEMM over [ GEGLU output | x2 ]: [c][5c] = [ W_po W_ff2 | W_po ] (engine.cu)
  float* wv2 = nullptr;  // attn2.to_v  [c, ctx]
  float* wo2 = nullptr;  // attn2.to_out.0.weight [c, c]
  float* bo2 = nullptr;  // attn2.to_out.0.bias
  int ctx_vec_off = 0;   // offset of this block's folded cross-attention vector inside ctx_vecs rows
};

struct Module {
  enum Kind { CONV_IN, RES, ST, DOWN, UP, OUT } kind;
  int idx = 0;             // index into the kind's weight table
  bool pop_skip = false;   // RES in output blocks: input = cat(h, hs.pop())
  bool push_skip = false;  // push the result onto hs after this module
};

class Arena {
 public:
  void reset(bool dry) { dry_ = dry; off_ = 0; if (dry) high_ = 0; }
  void* alloc(size_t bytes) {
    off_ = (off_ + 255) & ~static_cast<size_t>(255);
    void* p = dry_ ? nullptr : static_cast<void*>(base_ + off_);
    off_ += bytes;
    if (off_ > high_) high_ = off_;
    return p;
  }
  size_t mark() const { return off_; }
  void rewind(size_t m) { off_ = m; }
  size_t high() const { return high_; }
  char* base_ = nullptr;
  size_t cap_ = 0;

 private:
  bool dry_ = true;
  size_t off_ = 0, high_ = 0;
};

struct Prepared {
  int Bc = 0, H = 0, W = 0;
  bool pair = false;   // CFG pair plan: samples [0, Bc/2) and [Bc/2, Bc) share x and t (see Engine::forward_pair)
  Arena persist, scratch;
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<uint8_t> op_flags;       // graph lanes (OP_SIDE / OP_FORK / OP_JOIN, run_op_list)
  std::vector<std::string> op_names;
  std::vector<std::string> op_family;  // kernel family: conv_gemm, attention, groupnorm, layernorm, ...
  std::vector<double> op_flops;        // algorithmic FLOPs (2*MAC) of the op, 0 for non-GEMM ops
  std::vector<double> op_bytes;        // algorithmic HBM bytes of the op
  float* x_stage = nullptr;    // [Bc, in_ch, H, W]
  int64_t* t_stage = nullptr;  // [Bc]
  float* eps_stage = nullptr;  // [Bc, out_ch, H, W]
  cudaGraphExec_t graph = nullptr;
  int launches = 0;  // kernels per forward
  unsigned long long last_use = 0;   // Engine::prepare's LRU clock
  ~Prepared();
};

// Runs a launch plan: the first call runs every op eagerly once (function attributes are set outside capture), captures
// the list on a private stream and instantiates a CUDA graph; every call then replays the graph on the caller's stream.
// With use_graph = false the ops are simply launched one by one.
// `flags` (one per op, optional): lanes of the captured graph.  OP_SIDE ops are captured on a second stream, i.e. as a branch
// beside the main line; OP_FORK (with OP_SIDE) makes the branch wait for everything on the main line so far, OP_JOIN (a
// main-line op) makes the main line wait for the branch.  An op on one lane must not depend on ops of the other lane that
// lie between its fork and its join.  The eager paths run the list in order.
enum : uint8_t { OP_SIDE = 1, OP_FORK = 2, OP_JOIN = 4 };
int run_op_list(const std::vector<std::function<int(cudaStream_t)>>& ops, const std::vector<std::string>& names,
                cudaStream_t stream, bool use_graph, cudaGraphExec_t* graph, cudaStream_t* cap_stream,
                const std::vector<uint8_t>* flags = nullptr, cudaStream_t* side_stream = nullptr);

// Host-side weight staging shared by the U-Net engine and the VAE decoder: tensors arrive by reference state-dict name,
// are repacked at finalize() and uploaded once.
class WeightLoader {
 public:
  int load_weight(const char* name, const float* host, const int64_t* shape, int rank);

 protected:
  ~WeightLoader();
  const HostTensor* find(const std::string& name);
  int get(const std::string& name, const HostTensor** out);
  int make_conv(const std::string& prefix, int k, int cin, int cout, ConvW* w, int cin_pad = 0, int cout_pad = 0);
  int make_upconv_phases(const std::string& prefix, int cin, int cout, ConvW* w);   // fills w->wup
  int make_norm(const std::string& prefix, int c, NormW* n);
  int upload_f32(const std::vector<float>& v, float** dst);
  int upload_bf16(const std::vector<float>& v, bf16** dst);
  float round_operand(float v) const;   // v rounded to the 16-bit operand format and back
  // LayerNorm fold (norm1 -> qkv, norm3 -> GEGLU projection): gamma into the weight columns, rows centred, beta into the bias
  void fold_layernorm(std::vector<float>& w, std::vector<float>& bias, const std::vector<float>& gamma,
                      const std::vector<float>& beta, int rows, int c) const;
  std::unordered_map<std::string, HostTensor> host_;
  int fmt_f16_ = operand_f16();   // operand format the weights are repacked in (fixed at construction)
  bool finalized_ = false;
  std::vector<void*> dev_allocs_;
};

class Engine : public WeightLoader {
 public:
  explicit Engine(const pbe_config& cfg) : cfg_(cfg) {}
  ~Engine();
  int finalize();
  int set_context(const float* ctx_dev, int Bc, cudaStream_t stream);
  int forward(const float* x, const int64_t* t, float* eps, int Bc, int H, int W, cudaStream_t stream);
  // Classifier-free-guidance pair (plms.py:185-188: x_in = cat([x] * 2), t_in = cat([t] * 2), c_in = cat([uc, c])): the two
  // halves of the batch differ only in the context, and the context first enters at the first SpatialTransformer's
  // attn1.to_out (+ folded cross-attention).  Everything before that point -- conv_in, the first ResBlock, the first
  // transformer's norm / proj_in / LayerNorm / qkv / self-attention -- is computed ONCE for the B shared samples;
  // results are bit-identical to forward() on the duplicated batch.  x [B, in_ch, H, W], t [B] -> eps [2B, out_ch, H, W];
  // set_context must have been called with 2B rows (unconditional rows first).
  int forward_pair(const float* x, const int64_t* t, float* eps, int B, int H, int W, cudaStream_t stream);
  int launches_per_forward() const { return cur_ ? cur_->launches : 0; }
  // Eager forward with a CUDA event pair around every op; fills ms[i] for op i (returns number of ops, <0 on error).
  int profile_forward(const float* x, const int64_t* t, float* eps, int Bc, int H, int W, cudaStream_t stream,
                      float* ms, int max_ops);
  const Prepared* current() const { return cur_; }
  bool use_graph = true;
  // 16-bit residual stream (operand format) between ops; PBE_STREAM=fp32 keeps the round-1 fp32 stream (A/B, debugging)
  // (bf16 operands keep the fp32 stream by default: bf16 rounding of operands AND stream together exceeds the 1e-2 parity bar)
  bool stream16_ = [] {
    const char* e = getenv("PBE_STREAM");
    if (e != nullptr) return !(e[0] == 'f' || e[0] == 'F');
    return operand_f16() != 0;
  }();
  // nearest-2x upsample + 3x3 conv as four sub-pixel phase convs (16-bit stream only); PBE_SUBPIXEL_UP=0: the literal form
  // LayerNorm statistics from the producing GEMM's epilogue, applied in the consuming GEMM's epilogue (16-bit stream only);
  // PBE_LN_FOLD=0: a standalone normalise-only LayerNorm pass in front of the same (gamma / beta folded) GEMMs
  int ln_fold_ = [] { const char* e = getenv("PBE_LN_FOLD"); return e == nullptr ? 1 : atoi(e); }();
  // ff.net.2 -> (+x2) -> proj_out is linear end to end (attention.py:276, 335-336): one GEMM with K = 5C over the buffer that
  // holds the GEGLU output and x2 side by side, weights composed at load time (16-bit stream + LayerNorm fold only);
  // PBE_FF_PROJ_MERGE=0: the two GEMMs of the literal form
  int ff_proj_merge_ = [] { const char* e = getenv("PBE_FF_PROJ_MERGE"); return e == nullptr ? 1 : atoi(e); }();
  int subpixel_up_ = [] { const char* e = getenv("PBE_SUBPIXEL_UP"); return e == nullptr ? 1 : atoi(e); }();   // 2: at every size (tests)
  std::string last_error;

 private:
  int prepare(int Bc, int H, int W, bool pair = false);
  bool pair_plan_possible() const;
  int build(Prepared& P, bool dry);

  pbe_config cfg_;

  // network description
  std::vector<Module> modules_;
  std::vector<ResW> res_;
  std::vector<STW> st_;
  std::vector<ConvW> convs_;  // conv_in, downsample, upsample, out convs
  NormW out_norm_;
  float *te_w0_ = nullptr, *te_b0_ = nullptr, *te_w1_ = nullptr, *te_b1_ = nullptr;  // time_embed
  bf16* emb_w_ = nullptr;   // all ResBlock emb_layers concatenated [emb_total, 4*mc] (one GEMM per U-Net call)
  float* emb_b_ = nullptr;
  int emb_total_ = 0;
  int ctx_total_ = 0;  // sum of C over SpatialTransformers

  // context-dependent state (folded single-key cross-attention, K4)
  float* ctx_vecs_ = nullptr;  // [ctx_Bc, ctx_total_]
  float* ctx_tmp_ = nullptr;
  float *ln_ones_ = nullptr, *ln_zeros_ = nullptr;   // unit affine for the normalise-only LayerNorm pass
  int ctx_Bc_ = 0;

  std::map<std::tuple<int, int, int, int>, std::unique_ptr<Prepared>> prepared_;
  Prepared* cur_ = nullptr;
  static constexpr size_t kMaxPlans = 6;   // launch plans kept per engine (least recently used evicted)
  unsigned long long use_clock_ = 0;
  cudaStream_t cap_stream_ = nullptr;
  cudaStream_t side_stream_ = nullptr;   // capture-time second lane (run_op_list)
  // Branches of the captured graph: the timestep-embedding ops (they depend on t only) run beside conv_in and the first
  // GroupNorm and join at the first ResBlock's conv1; in small-batch plans a ResBlock's 1x1 skip convolution runs beside
  // its conv1 / GroupNorm.  PBE_GRAPH_LANES=0: one line.
  int graph_lanes_ = [] { const char* e = getenv("PBE_GRAPH_LANES"); return e == nullptr ? 1 : atoi(e); }();
};

}  // namespace pbe
