// VAE engine implementation (decoder and encoder launch plans).  See vae.h.
//
// Same data conventions as the U-Net engine: NHWC activations, fp32 block outputs (GroupNorm statistics fused into the
// producing GEMM's epilogue), bf16 GEMM operands written by the GroupNorm(+swish) kernels, fp32 accumulation in TMEM.
// Block outputs ping-pong between two buffers sized for the largest tensor (nothing is skip-connected in the decoder).
#include "vae.h"

#include <math.h>

#include <algorithm>

namespace pbe {

VaePrepared::~VaePrepared() {
  if (graph) cudaGraphExecDestroy(graph);
  if (persist.base_) cudaFree(persist.base_);
  if (scratch.base_) cudaFree(scratch.base_);
}

int VaeModel::make_res(const std::string& pfx, int cin, int cout, VaeResW* r) {
  int e;
  r->cin = cin; r->cout = cout;
  if ((e = make_norm(pfx + ".norm1", cin, &r->n1))) return e;
  if ((e = make_conv(pfx + ".conv1", 3, cin, cout, &r->conv1))) return e;
  if ((e = make_norm(pfx + ".norm2", cout, &r->n2))) return e;
  if ((e = make_conv(pfx + ".conv2", 3, cout, cout, &r->conv2))) return e;
  r->has_nin = cin != cout;
  if (r->has_nin) {
    PBE_REQUIRE(find(pfx + ".conv_shortcut.weight") == nullptr, "conv_shortcut=True ResnetBlocks are not supported");
    if ((e = make_conv(pfx + ".nin_shortcut", 1, cin, cout, &r->nin))) return e;
  }
  return 0;
}

// AttnBlock: q, k, v 1x1 convs fused into one [3C, C] GEMM (Q | K row-major, V transposed per sample)
int VaeModel::make_attn(const std::string& pfx, int C, VaeAttnW* a) {
  int rc;
  a->c = C;
  if ((rc = make_norm(pfx + ".norm", C, &a->norm))) return rc;
  std::vector<float> w(static_cast<size_t>(3) * C * C), b(static_cast<size_t>(3) * C);
  const char* names[3] = {"q", "k", "v"};
  for (int j = 0; j < 3; ++j) {
    const HostTensor *Wj, *Bj;
    if ((rc = get(pfx + "." + names[j] + ".weight", &Wj))) return rc;
    if ((rc = get(pfx + "." + names[j] + ".bias", &Bj))) return rc;
    PBE_REQUIRE(Wj->data.size() == static_cast<size_t>(C) * C && Bj->data.size() == static_cast<size_t>(C),
                "attn_1 q/k/v have the wrong shape");
    std::copy(Wj->data.begin(), Wj->data.end(), w.begin() + static_cast<size_t>(j) * C * C);
    std::copy(Bj->data.begin(), Bj->data.end(), b.begin() + static_cast<size_t>(j) * C);
  }
  if ((rc = upload_bf16(w, &a->qkv.w))) return rc;
  if ((rc = upload_f32(b, &a->qkv.b))) return rc;
  a->qkv.cin = a->qkv.cin_pad = C; a->qkv.cout = 3 * C; a->qkv.k = 1;
  return make_conv(pfx + ".proj_out", 1, C, C, &a->proj_out);
}

int VaeModel::finalize() {
  PBE_REQUIRE(!finalized_, "weights already finalized");
  const int L = cfg_.num_levels;
  PBE_REQUIRE(L >= 1 && L <= 8 && cfg_.ch % 64 == 0 && cfg_.z_channels <= 8 && cfg_.embed_dim <= 8 && cfg_.out_ch <= 4 &&
                  cfg_.in_channels <= 64,
              "unsupported VAE configuration");
  int rc;
  if (find("decoder.conv_in.weight") != nullptr) {
    if ((rc = finalize_decoder())) return rc;
    has_dec_ = true;
  }
  if (find("encoder.conv_in.weight") != nullptr) {
    if ((rc = finalize_encoder())) return rc;
    has_enc_ = true;
  }
  PBE_REQUIRE(has_dec_ || has_enc_, "neither decoder.* nor encoder.* weights were loaded");
  finalized_ = true;
  host_.clear();
  return 0;
}

int VaeModel::finalize_decoder() {
  const int L = cfg_.num_levels;
  int rc;
  // post_quant_conv: [z_channels, embed_dim, 1, 1] fp32 (applied inside the input pack kernel)
  {
    const HostTensor *W, *Bv;
    if ((rc = get("post_quant_conv.weight", &W))) return rc;
    if ((rc = get("post_quant_conv.bias", &Bv))) return rc;
    PBE_REQUIRE(static_cast<int>(W->data.size()) == cfg_.z_channels * cfg_.embed_dim &&
                    static_cast<int>(Bv->data.size()) == cfg_.z_channels,
                "post_quant_conv has the wrong shape");
    if ((rc = upload_f32(W->data, &pq_w_))) return rc;
    if ((rc = upload_f32(Bv->data, &pq_b_))) return rc;
  }
  int block_in = cfg_.ch * cfg_.ch_mult[L - 1];
  if ((rc = make_conv("decoder.conv_in", 3, cfg_.z_channels, block_in, &conv_in_, 64))) return rc;
  if ((rc = make_res("decoder.mid.block_1", block_in, block_in, &mid1_))) return rc;
  if ((rc = make_attn("decoder.mid.attn_1", block_in, &attn_))) return rc;
  if ((rc = make_res("decoder.mid.block_2", block_in, block_in, &mid2_))) return rc;
  up_blocks_.assign(L, {});
  up_convs_.assign(L, ConvW{});
  for (int lvl = L - 1; lvl >= 0; --lvl) {
    const int block_out = cfg_.ch * cfg_.ch_mult[lvl];
    PBE_REQUIRE(block_out % 64 == 0, "decoder channels must be multiples of 64");
    for (int i = 0; i < cfg_.num_res_blocks + 1; ++i) {
      VaeResW r;
      const std::string pfx = "decoder.up." + std::to_string(lvl) + ".block." + std::to_string(i);
      if ((rc = make_res(pfx, block_in, block_out, &r))) return rc;
      up_blocks_[lvl].push_back(r);
      block_in = block_out;
      PBE_REQUIRE(find("decoder.up." + std::to_string(lvl) + ".attn." + std::to_string(i) + ".norm.weight") == nullptr,
                  "decoder attn_resolutions other than [] are not supported");
    }
    if (lvl != 0)
      if ((rc = make_conv("decoder.up." + std::to_string(lvl) + ".upsample.conv", 3, block_in, block_in, &up_convs_[lvl])))
        return rc;
  }
  if ((rc = make_norm("decoder.norm_out", block_in, &norm_out_))) return rc;
  return make_conv("decoder.conv_out", 3, block_in, cfg_.out_ch, &conv_out_, 0, 4);
}

// Encoder.__init__ (model.py:370-438) + quant_conv (autoencoder.py:36)
int VaeModel::finalize_encoder() {
  const int L = cfg_.num_levels;
  const int zc2 = 2 * cfg_.z_channels, e2 = 2 * cfg_.embed_dim;
  int rc;
  {
    const HostTensor *W, *Bv;
    if ((rc = get("quant_conv.weight", &W))) return rc;
    if ((rc = get("quant_conv.bias", &Bv))) return rc;
    PBE_REQUIRE(static_cast<int>(W->data.size()) == e2 * zc2 && static_cast<int>(Bv->data.size()) == e2,
                "quant_conv has the wrong shape");
    if ((rc = upload_f32(W->data, &q_w_))) return rc;
    if ((rc = upload_f32(Bv->data, &q_b_))) return rc;
  }
  if ((rc = make_conv("encoder.conv_in", 3, cfg_.in_channels, cfg_.ch, &e_conv_in_, 64))) return rc;
  down_blocks_.assign(L, {});
  down_convs_.assign(L, ConvW{});
  int block_in = cfg_.ch;
  for (int lvl = 0; lvl < L; ++lvl) {
    const int block_out = cfg_.ch * cfg_.ch_mult[lvl];
    PBE_REQUIRE(block_out % 64 == 0, "encoder channels must be multiples of 64");
    for (int i = 0; i < cfg_.num_res_blocks; ++i) {
      VaeResW r;
      if ((rc = make_res("encoder.down." + std::to_string(lvl) + ".block." + std::to_string(i), block_in, block_out, &r)))
        return rc;
      down_blocks_[lvl].push_back(r);
      block_in = block_out;
    }
    if (lvl != L - 1)
      if ((rc = make_conv("encoder.down." + std::to_string(lvl) + ".downsample.conv", 3, block_in, block_in, &down_convs_[lvl])))
        return rc;
  }
  if ((rc = make_res("encoder.mid.block_1", block_in, block_in, &e_mid1_))) return rc;
  if ((rc = make_attn("encoder.mid.attn_1", block_in, &e_attn_))) return rc;
  if ((rc = make_res("encoder.mid.block_2", block_in, block_in, &e_mid2_))) return rc;
  if ((rc = make_norm("encoder.norm_out", block_in, &e_norm_out_))) return rc;
  PBE_REQUIRE(zc2 % 4 == 0, "2 * z_channels must be a multiple of 4");
  return make_conv("encoder.conv_out", 3, block_in, zc2, &e_conv_out_);
}

int VaeModel::prepare(int enc, int B, int H, int W) {
  const auto key = std::make_tuple(enc, B, H, W);
  auto it = prepared_.find(key);
  if (it != prepared_.end()) { cur_ = it->second.get(); return 0; }
  PBE_REQUIRE(finalized_, "pbe_vae_finalize_weights has not been called");
  PBE_REQUIRE(enc ? has_enc_ : has_dec_, enc ? "the encoder half (encoder.*, quant_conv.*) was not loaded"
                                              : "the decoder half (decoder.*, post_quant_conv.*) was not loaded");
  const int f = 1 << (cfg_.num_levels - 1);
  if (enc) {
    PBE_REQUIRE(B >= 1 && H % f == 0 && W % f == 0 && ((H / f) * (W / f)) % 64 == 0,
                "image H, W must be multiples of the downsampling factor with (H/f)*(W/f) a multiple of 64");
  } else {
    PBE_REQUIRE(B >= 1 && H >= 1 && W >= 1 && (H * W) % 64 == 0, "latent H*W must be a multiple of 64");
  }
  auto P = std::make_unique<VaePrepared>();
  P->B = B; P->H = H; P->W = W;
  int rc = build(*P, true, enc);
  if (rc) return rc;
  for (Arena* a : {&P->persist, &P->scratch}) {
    a->cap_ = a->high() + 4096;
    void* p = nullptr;
    PBE_CHECK_CUDA(cudaMalloc(&p, a->cap_));
    a->base_ = static_cast<char*>(p);
    PBE_CHECK_CUDA(cudaMemset(p, 0, a->cap_));
  }
  rc = build(*P, false, enc);
  if (rc) return rc;
  cur_ = P.get();
  prepared_[key] = std::move(P);
  return 0;
}

int VaeModel::build(VaePrepared& P, bool dry, int enc) {
  const int B = P.B, H0 = P.H, W0 = P.W, L = cfg_.num_levels;
  P.persist.reset(dry);
  P.scratch.reset(dry);
  auto PA = [&](size_t bytes) { return P.persist.alloc(bytes); };
  auto SA = [&](size_t bytes) { return P.scratch.alloc(bytes); };
  int launches = 0, err = 0;
  P.ops.clear(); P.op_names.clear(); P.op_family.clear(); P.op_flops.clear();
  auto add_op = [&](const std::string& name, int nlaunch, std::function<int(cudaStream_t)> fn, const std::string& family,
                    double flops) {
    if (!dry) {
      P.ops.push_back(std::move(fn));
      P.op_names.push_back(name);
      P.op_family.push_back(family);
      P.op_flops.push_back(flops);
    }
    launches += nlaunch;
  };
  auto add_gemm = [&](const std::string& name, ConvGemmDesc d) {
    const size_t ws_bytes = d.stats_out ? 0 : gemm_splitk_ws_bytes(d);
    d.splitk_ws = ws_bytes ? static_cast<float*>(SA(ws_bytes)) : nullptr;
    if (dry) { launches += ws_bytes ? 2 : 1; return; }
    auto plan = std::make_shared<GemmPlan>();
    int rc = build_gemm_plan(d, plan.get());
    if (rc && !err) { err = rc; last_error = std::string(get_error()) + " [" + name + "]"; }
    const double Mrows = static_cast<double>(d.Nb) * d.H * d.W;
    const double flops = 2.0 * Mrows * d.Cout * static_cast<double>(d.c_real ? d.c_real : d.C) * d.ksize * d.ksize;
    add_op(name, ws_bytes ? 2 : 1, [plan](cudaStream_t s) { return launch_gemm_plan(*plan, s); }, "conv_gemm", flops);
  };
  auto add_gn = [&](const std::string& name, GroupNormArgs a) {
    a.partial = static_cast<float*>(SA(static_cast<size_t>(gn_workspace_floats(a.Nb, a.HW, a.C0 + a.C1)) * sizeof(float)));
    add_op(name, gn_num_launches(a), [a](cudaStream_t s) { return launch_groupnorm(a, s); }, "groupnorm", 0.0);
  };

  // largest block output (elements) and largest statistics buffer over the whole network
  size_t max_elems = 0, max_stats = 0;
  {
    auto visit = [&](int ch, int h2, int w2) {
      const size_t rows = static_cast<size_t>(B) * h2 * w2;
      max_elems = std::max(max_elems, rows * ch);
      max_stats = std::max(max_stats, (rows / 32 + 1) * ch * 2);
    };
    int hh = H0, ww = W0;
    if (enc) {
      visit(cfg_.ch, hh, ww);
      for (int lvl = 0; lvl < L; ++lvl) {
        visit(cfg_.ch * cfg_.ch_mult[lvl], hh, ww);
        if (lvl != L - 1) { hh /= 2; ww /= 2; visit(cfg_.ch * cfg_.ch_mult[lvl], hh, ww); }
      }
    } else {
      int c = cfg_.ch * cfg_.ch_mult[L - 1];
      visit(c, hh, ww);
      for (int lvl = L - 1; lvl >= 0; --lvl) {
        c = cfg_.ch * cfg_.ch_mult[lvl];
        visit(c, hh, ww);
        if (lvl != 0) { hh *= 2; ww *= 2; visit(c, hh, ww); }
      }
    }
  }
  float* chain[2] = {static_cast<float*>(PA(max_elems * sizeof(float))), static_cast<float*>(PA(max_elems * sizeof(float)))};
  float* chain_stats[2] = {static_cast<float*>(PA(max_stats * sizeof(float))), static_cast<float*>(PA(max_stats * sizeof(float)))};
  int cur = 0;   // chain buffer that holds the current activation

  struct Act {
    float* f32;
    int C, H, W;
    float* stats;
    bool has_stats;
  };
  // Output activation in the other chain buffer; fused GroupNorm statistics when the GEMM geometry allows
  auto next_act = [&](int C, int H, int W) {
    cur ^= 1;
    return Act{chain[cur], C, H, W, nullptr, false};
  };
  auto want_stats = [&](ConvGemmDesc& d, Act& o, float* stats_buf) {
    d.splitk_ws = nullptr;
    if (!gemm_can_fuse_stats(d)) return;
    o.stats = stats_buf;
    o.has_stats = true;
    d.stats_out = stats_buf;
  };
  const int fdown = 1 << (L - 1);
  if (enc) {
    P.z_stage = static_cast<float*>(PA(static_cast<size_t>(B) * cfg_.in_channels * H0 * W0 * sizeof(float)));
    P.out_stage = static_cast<float*>(PA(static_cast<size_t>(B) * 2 * cfg_.embed_dim * (H0 / fdown) * (W0 / fdown) * sizeof(float)));
  } else {
    P.z_stage = static_cast<float*>(PA(static_cast<size_t>(B) * cfg_.embed_dim * H0 * W0 * sizeof(float)));
    P.out_stage = static_cast<float*>(PA(static_cast<size_t>(B) * cfg_.out_ch * H0 * fdown * W0 * fdown * sizeof(float)));
  }
  // raw bf16 copy of a block output that feeds a stride-2 Downsample conv (encoder only)
  bf16* down_in16 = enc ? static_cast<bf16*>(PA(max_elems * sizeof(bf16))) : nullptr;

  auto res_block = [&](const std::string& tag, const VaeResW& r, Act& h, bf16* out16 = nullptr) {
    const size_t smark = P.scratch.mark();
    const size_t M = static_cast<size_t>(B) * h.H * h.W;
    bf16* a1 = static_cast<bf16*>(SA(M * r.cin * sizeof(bf16)));
    bf16* raw = r.has_nin ? static_cast<bf16*>(SA(M * r.cin * sizeof(bf16))) : nullptr;
    GroupNormArgs g1{};
    g1.x0 = h.f32; g1.C0 = r.cin; g1.Nb = B; g1.HW = h.H * h.W; g1.gamma = r.n1.g; g1.beta = r.n1.b; g1.eps = 1e-6f;
    g1.silu = 1; g1.y = a1; g1.raw = raw; g1.stats0 = h.has_stats ? h.stats : nullptr;
    add_gn(tag + ".norm1", g1);
    float* h1 = static_cast<float*>(SA(M * r.cout * sizeof(float)));
    Act h1act{h1, r.cout, h.H, h.W, nullptr, false};
    {
      ConvGemmDesc d{};
      d.act = a1; d.Nb = B; d.H = h.H; d.W = h.W; d.C = r.cin; d.ksize = 3; d.stride = 1;
      d.wt = r.conv1.w; d.Cout = r.cout; d.mode = EPI_STD; d.bias = r.conv1.b; d.out_f32 = h1;
      want_stats(d, h1act, static_cast<float*>(SA((M / 32 + 1) * r.cout * 2 * sizeof(float))));
      add_gemm(tag + ".conv1", d);
    }
    bf16* a2 = static_cast<bf16*>(SA(M * r.cout * sizeof(bf16)));
    GroupNormArgs g2{};
    g2.x0 = h1; g2.C0 = r.cout; g2.Nb = B; g2.HW = h.H * h.W; g2.gamma = r.n2.g; g2.beta = r.n2.b; g2.eps = 1e-6f;
    g2.silu = 1; g2.y = a2; g2.stats0 = h1act.has_stats ? h1act.stats : nullptr;
    add_gn(tag + ".norm2", g2);
    const float* resid = h.f32;
    if (r.has_nin) {
      float* sc = static_cast<float*>(SA(M * r.cout * sizeof(float)));
      ConvGemmDesc d{};
      d.act = raw; d.Nb = B; d.H = h.H; d.W = h.W; d.C = r.cin; d.ksize = 1; d.stride = 1;
      d.wt = r.nin.w; d.Cout = r.cout; d.mode = EPI_STD; d.bias = r.nin.b; d.out_f32 = sc;
      add_gemm(tag + ".nin_shortcut", d);
      resid = sc;
    }
    Act o = next_act(r.cout, h.H, h.W);
    {
      ConvGemmDesc d{};
      d.act = a2; d.Nb = B; d.H = h.H; d.W = h.W; d.C = r.cout; d.ksize = 3; d.stride = 1;
      d.wt = r.conv2.w; d.Cout = r.cout; d.mode = EPI_STD; d.bias = r.conv2.b; d.residual = resid; d.out_f32 = o.f32;
      d.out_bf16 = out16;
      want_stats(d, o, chain_stats[cur]);
      add_gemm(tag + ".conv2", d);
    }
    h = o;
    P.scratch.rewind(smark);
  };

  // AttnBlock (model.py:152-182): h + proj_out(softmax(q k^T / sqrt(C)) v), single head of width C
  auto attn_block = [&](const std::string& tag, const VaeAttnW& aw, Act& h) {
    const size_t smark = P.scratch.mark();
    const int C = aw.c, N = h.H * h.W;
    const size_t M = static_cast<size_t>(B) * N;
    bf16* a = static_cast<bf16*>(SA(M * C * sizeof(bf16)));
    GroupNormArgs g{};
    g.x0 = h.f32; g.C0 = C; g.Nb = B; g.HW = N; g.gamma = aw.norm.g; g.beta = aw.norm.b; g.eps = 1e-6f; g.silu = 0;
    g.y = a; g.stats0 = h.has_stats ? h.stats : nullptr;
    add_gn(tag + ".norm", g);
    bf16* qk = static_cast<bf16*>(SA(M * 2 * C * sizeof(bf16)));
    bf16* vt = static_cast<bf16*>(SA(M * C * sizeof(bf16)));   // [B][C][N]: N % 64 == 0 here, so vt_pitch(N) == N
    {
      ConvGemmDesc d{};
      d.act = a; d.Nb = B; d.H = h.H; d.W = h.W; d.C = C; d.ksize = 1; d.stride = 1;
      d.wt = aw.qkv.w; d.Cout = 3 * C; d.mode = EPI_QKV; d.bias = aw.qkv.b; d.out_bf16 = qk; d.ld_out = 2 * C;
      d.out_vt = vt; d.qk_cols = 2 * C;
      add_gemm(tag + ".qkv", d);
    }
    float* S = static_cast<float*>(SA(static_cast<size_t>(N) * N * sizeof(float)));
    bf16* Pm = static_cast<bf16*>(SA(static_cast<size_t>(N) * N * sizeof(bf16)));
    bf16* o16 = static_cast<bf16*>(SA(M * C * sizeof(bf16)));
    const float scale = 1.0f / sqrtf(static_cast<float>(C));
    for (int b = 0; b < B; ++b) {
      const std::string tb = tag + "[" + std::to_string(b) + "]";
      {
        ConvGemmDesc d{};   // S = Q K^T: Q is the activation, K the "weight" (both column slices of the Q|K buffer)
        d.act = qk + static_cast<size_t>(b) * N * 2 * C; d.act_ld = 2 * C; d.Nb = 1; d.H = h.H; d.W = h.W; d.C = C;
        d.ksize = 1; d.stride = 1; d.wt = qk + static_cast<size_t>(b) * N * 2 * C + C; d.wt_ld = 2 * C; d.Cout = N;
        d.mode = EPI_STD; d.out_f32 = S;
        add_gemm(tb + ".qk", d);
      }
      add_op(tb + ".softmax", 1, [=](cudaStream_t s) { return launch_softmax_rows(S, Pm, N, N, scale, s); }, "softmax", 0.0);
      {
        ConvGemmDesc d{};   // O = P V: P [N, N] is the activation, V^T [C, N] the weight
        d.act = Pm; d.Nb = 1; d.H = h.H; d.W = h.W; d.C = N; d.ksize = 1; d.stride = 1;
        d.wt = vt + static_cast<size_t>(b) * C * N; d.Cout = C; d.mode = EPI_STD;
        d.out_bf16 = o16 + static_cast<size_t>(b) * N * C;
        add_gemm(tb + ".pv", d);
      }
    }
    Act o = next_act(C, h.H, h.W);
    {
      ConvGemmDesc d{};
      d.act = o16; d.Nb = B; d.H = h.H; d.W = h.W; d.C = C; d.ksize = 1; d.stride = 1;
      d.wt = aw.proj_out.w; d.Cout = C; d.mode = EPI_STD; d.bias = aw.proj_out.b; d.residual = h.f32; d.out_f32 = o.f32;
      want_stats(d, o, chain_stats[cur]);
      add_gemm(tag + ".proj_out", d);
    }
    h = o;
    P.scratch.rewind(smark);
  };

  Act h{nullptr, 0, H0, W0, nullptr, false};
  if (!enc) {
    // ================= decoder: Decoder.forward after post_quant_conv (model.py:542-580) =================
    {
      const size_t smark = P.scratch.mark();
      const size_t M = static_cast<size_t>(B) * H0 * W0;
      bf16* xin = static_cast<bf16*>(SA(M * 64 * sizeof(bf16)));
      const float* zs = P.z_stage;
      const float *pw = pq_w_, *pb = pq_b_;
      const int ed = cfg_.embed_dim, zc = cfg_.z_channels;
      add_op("post_quant_conv+pack", 1,
             [=](cudaStream_t s) { return launch_vae_pack_input(zs, pw, pb, xin, B, ed, zc, H0, W0, 64, s); }, "misc", 0.0);
      Act o{chain[cur], conv_in_.cout, H0, W0, nullptr, false};
      ConvGemmDesc d{};
      d.act = xin; d.Nb = B; d.H = H0; d.W = W0; d.C = 64; d.c_real = zc; d.ksize = 3; d.stride = 1;
      d.wt = conv_in_.w; d.Cout = conv_in_.cout; d.mode = EPI_STD; d.bias = conv_in_.b; d.out_f32 = o.f32;
      want_stats(d, o, chain_stats[cur]);
      add_gemm("conv_in", d);
      h = o;
      P.scratch.rewind(smark);
    }
    res_block("mid.block_1", mid1_, h);
    attn_block("mid.attn_1", attn_, h);
    res_block("mid.block_2", mid2_, h);
    for (int lvl = L - 1; lvl >= 0; --lvl) {
      for (size_t i = 0; i < up_blocks_[lvl].size(); ++i)
        res_block("up." + std::to_string(lvl) + ".block." + std::to_string(i), up_blocks_[lvl][i], h);
      if (lvl != 0) {
        const size_t smark = P.scratch.mark();
        const size_t M = static_cast<size_t>(B) * (2 * h.H) * (2 * h.W);
        bf16* up = static_cast<bf16*>(SA(M * h.C * sizeof(bf16)));
        const float* src = h.f32;
        const int hh = h.H, ww = h.W, cc = h.C;
        add_op("up." + std::to_string(lvl) + ".upsample2x", 1,
               [=](cudaStream_t s) { return launch_upsample2x_bf16(src, up, B, hh, ww, cc, s); }, "upsample", 0.0);
        Act o = next_act(h.C, 2 * h.H, 2 * h.W);
        ConvGemmDesc d{};
        d.act = up; d.Nb = B; d.H = 2 * hh; d.W = 2 * ww; d.C = cc; d.ksize = 3; d.stride = 1;
        d.wt = up_convs_[lvl].w; d.Cout = cc; d.mode = EPI_STD; d.bias = up_convs_[lvl].b; d.out_f32 = o.f32;
        want_stats(d, o, chain_stats[cur]);
        add_gemm("up." + std::to_string(lvl) + ".upsample.conv", d);
        h = o;
        P.scratch.rewind(smark);
      }
    }
    {
      const size_t smark = P.scratch.mark();
      const size_t M = static_cast<size_t>(B) * h.H * h.W;
      bf16* a = static_cast<bf16*>(SA(M * h.C * sizeof(bf16)));
      GroupNormArgs g{};
      g.x0 = h.f32; g.C0 = h.C; g.Nb = B; g.HW = h.H * h.W; g.gamma = norm_out_.g; g.beta = norm_out_.b; g.eps = 1e-6f;
      g.silu = 1; g.y = a; g.stats0 = h.has_stats ? h.stats : nullptr;
      add_gn("norm_out", g);
      float* y = static_cast<float*>(SA(M * 4 * sizeof(float)));
      ConvGemmDesc d{};
      d.act = a; d.Nb = B; d.H = h.H; d.W = h.W; d.C = h.C; d.ksize = 3; d.stride = 1;
      d.wt = conv_out_.w; d.Cout = 4; d.mode = EPI_STD; d.bias = conv_out_.b; d.out_f32 = y;
      add_gemm("conv_out", d);
      float* outp = P.out_stage;
      const int oc = cfg_.out_ch, hh = h.H, ww = h.W;
      add_op("unpack_output", 1, [=](cudaStream_t s) { return launch_unpack_output(y, outp, B, oc, hh, ww, 4, s); }, "misc", 0.0);
      P.scratch.rewind(smark);
    }
  } else {
    // ================= encoder: quant_conv(Encoder.forward(x)) (model.py:440-471, autoencoder.py:56-61) =================
    {
      const size_t smark = P.scratch.mark();
      const size_t M = static_cast<size_t>(B) * H0 * W0;
      bf16* xin = static_cast<bf16*>(SA(M * 64 * sizeof(bf16)));
      const float* xs = P.z_stage;
      const int cin = cfg_.in_channels;
      add_op("pack_input", 1, [=](cudaStream_t s) { return launch_pack_input(xs, xin, B, cin, H0, W0, 64, s); }, "misc", 0.0);
      Act o{chain[cur], e_conv_in_.cout, H0, W0, nullptr, false};
      ConvGemmDesc d{};
      d.act = xin; d.Nb = B; d.H = H0; d.W = W0; d.C = 64; d.c_real = cin; d.ksize = 3; d.stride = 1;
      d.wt = e_conv_in_.w; d.Cout = e_conv_in_.cout; d.mode = EPI_STD; d.bias = e_conv_in_.b; d.out_f32 = o.f32;
      want_stats(d, o, chain_stats[cur]);
      add_gemm("conv_in", d);
      h = o;
      P.scratch.rewind(smark);
    }
    for (int lvl = 0; lvl < L; ++lvl) {
      const size_t nb = down_blocks_[lvl].size();
      for (size_t i = 0; i < nb; ++i) {
        const bool feeds_down = (lvl != L - 1) && (i + 1 == nb);   // a stride-2 conv reads raw bf16 activations
        res_block("down." + std::to_string(lvl) + ".block." + std::to_string(i), down_blocks_[lvl][i], h,
                  feeds_down ? down_in16 : nullptr);
      }
      if (lvl != L - 1) {
        if (nb == 0 || (h.H & 1) || (h.W & 1)) { err = -5; last_error = "Downsample needs a preceding block and even H, W"; return err; }
        Act o = next_act(h.C, h.H / 2, h.W / 2);
        ConvGemmDesc d{};   // Downsample: zero padding (0,1,0,1) then a stride-2 3x3 conv (model.py:62-77)
        d.act = down_in16; d.Nb = B; d.H = h.H; d.W = h.W; d.C = h.C; d.ksize = 3; d.stride = 2; d.pad_end = 1;
        d.wt = down_convs_[lvl].w; d.Cout = h.C; d.mode = EPI_STD; d.bias = down_convs_[lvl].b; d.out_f32 = o.f32;
        want_stats(d, o, chain_stats[cur]);
        add_gemm("down." + std::to_string(lvl) + ".downsample", d);
        h = o;
      }
    }
    res_block("mid.block_1", e_mid1_, h);
    attn_block("mid.attn_1", e_attn_, h);
    res_block("mid.block_2", e_mid2_, h);
    {
      const size_t smark = P.scratch.mark();
      const size_t M = static_cast<size_t>(B) * h.H * h.W;
      bf16* a = static_cast<bf16*>(SA(M * h.C * sizeof(bf16)));
      GroupNormArgs g{};
      g.x0 = h.f32; g.C0 = h.C; g.Nb = B; g.HW = h.H * h.W; g.gamma = e_norm_out_.g; g.beta = e_norm_out_.b; g.eps = 1e-6f;
      g.silu = 1; g.y = a; g.stats0 = h.has_stats ? h.stats : nullptr;
      add_gn("norm_out", g);
      const int zc2 = 2 * cfg_.z_channels, e2 = 2 * cfg_.embed_dim;
      float* y = static_cast<float*>(SA(M * zc2 * sizeof(float)));
      ConvGemmDesc d{};
      d.act = a; d.Nb = B; d.H = h.H; d.W = h.W; d.C = h.C; d.ksize = 3; d.stride = 1;
      d.wt = e_conv_out_.w; d.Cout = zc2; d.mode = EPI_STD; d.bias = e_conv_out_.b; d.out_f32 = y;
      add_gemm("conv_out", d);
      float* outp = P.out_stage;
      const float *qw = q_w_, *qb = q_b_;
      const int hh = h.H, ww = h.W;
      add_op("quant_conv+unpack", 1,
             [=](cudaStream_t s) { return launch_vae_unpack_moments(y, qw, qb, outp, B, zc2, e2, hh, ww, zc2, s); }, "misc", 0.0);
      P.scratch.rewind(smark);
    }
  }
  P.launches = launches;
  return err;
}

int VaeModel::run(VaePrepared& P, const float* in, size_t in_bytes, float* out, size_t out_bytes, cudaStream_t stream,
                  float* ms, int max_ops) {
  const int n = static_cast<int>(P.ops.size());
  std::vector<cudaEvent_t> ev;
  if (ms != nullptr) {
    PBE_REQUIRE(n <= max_ops, "profile buffer too small");
    ev.resize(n + 1);
    for (auto& e : ev) PBE_CHECK_CUDA(cudaEventCreate(&e));
  }
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.z_stage, in, in_bytes, cudaMemcpyDeviceToDevice, stream));
  if (ms == nullptr) {   // product path: CUDA-graph replay of the launch plan
    int rc = run_op_list(P.ops, P.op_names, stream, true, &P.graph, &cap_stream_);
    if (rc) { last_error = get_error(); return rc; }
    PBE_CHECK_CUDA(cudaMemcpyAsync(out, P.out_stage, out_bytes, cudaMemcpyDeviceToDevice, stream));
    return 0;
  }
  PBE_CHECK_CUDA(cudaEventRecord(ev[0], stream));
  for (int i = 0; i < n; ++i) {
    int rc = P.ops[i](stream);
    if (rc) { last_error = std::string(get_error()) + " [" + P.op_names[i] + "]"; set_error(last_error); return rc; }
    if (ms != nullptr) PBE_CHECK_CUDA(cudaEventRecord(ev[i + 1], stream));
  }
  PBE_CHECK_CUDA(cudaMemcpyAsync(out, P.out_stage, out_bytes, cudaMemcpyDeviceToDevice, stream));
  if (ms != nullptr) {
    PBE_CHECK_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < n; ++i) PBE_CHECK_CUDA(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    for (auto& e : ev) cudaEventDestroy(e);
    return n;
  }
  return 0;
}

int VaeModel::decode(const float* z, float* out, int B, int H, int W, cudaStream_t stream) {
  int rc = prepare(0, B, H, W);
  if (rc) return rc;
  const size_t f = static_cast<size_t>(1) << (cfg_.num_levels - 1);
  return run(*cur_, z, static_cast<size_t>(B) * cfg_.embed_dim * H * W * sizeof(float), out,
             static_cast<size_t>(B) * cfg_.out_ch * H * f * W * f * sizeof(float), stream, nullptr, 0);
}

int VaeModel::encode(const float* x, float* moments, int B, int H, int W, cudaStream_t stream) {
  int rc = prepare(1, B, H, W);
  if (rc) return rc;
  const size_t f = static_cast<size_t>(1) << (cfg_.num_levels - 1);
  return run(*cur_, x, static_cast<size_t>(B) * cfg_.in_channels * H * W * sizeof(float), moments,
             static_cast<size_t>(B) * 2 * cfg_.embed_dim * (H / f) * (W / f) * sizeof(float), stream, nullptr, 0);
}

int VaeModel::profile(int enc, const float* in, float* out, int B, int H, int W, cudaStream_t stream, float* ms, int max_ops) {
  int rc = prepare(enc ? 1 : 0, B, H, W);
  if (rc) return rc;
  const size_t f = static_cast<size_t>(1) << (cfg_.num_levels - 1);
  const size_t in_bytes = enc ? static_cast<size_t>(B) * cfg_.in_channels * H * W * sizeof(float)
                              : static_cast<size_t>(B) * cfg_.embed_dim * H * W * sizeof(float);
  const size_t out_bytes = enc ? static_cast<size_t>(B) * 2 * cfg_.embed_dim * (H / f) * (W / f) * sizeof(float)
                               : static_cast<size_t>(B) * cfg_.out_ch * H * f * W * f * sizeof(float);
  return run(*cur_, in, in_bytes, out, out_bytes, stream, ms, max_ops);
}

}  // namespace pbe
