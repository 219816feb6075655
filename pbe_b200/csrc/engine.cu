// U-Net engine implementation.  See engine.h.
//
// Data layout in HBM: activations are NHWC.  Every GEMM operand and the residual stream between ops are 16-bit in the
// library's operand format (fp16 by default; PBE_STREAM=fp32 keeps an fp32 stream); statistics for GroupNorm / LayerNorm are
// fp32, taken from the rounded values in the producing GEMM's epilogue; accumulation is fp32 in TMEM.  Weights are repacked
// once at load: conv OIHW fp32 -> [tap][O][I] 16-bit, Linear [O,I] -> 16-bit, q/k/v fused to one [3C,C] matrix with norm1
// folded in, the GEGLU projection (value/gate rows interleaved per 256-column tile) with norm3 folded in, the Upsample convs
// additionally as four sub-pixel phase kernels, all 22 ResBlock emb_layers concatenated into one matrix.
#include "engine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace pbe {

namespace {

constexpr int MAX_BC = 128;

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return static_cast<uint16_t>(u >> 16);
}

}  // namespace

Prepared::~Prepared() {
  if (graph) cudaGraphExecDestroy(graph);
  if (persist.base_) cudaFree(persist.base_);
  if (scratch.base_) cudaFree(scratch.base_);
}

Engine::~Engine() {
  prepared_.clear();
  if (cap_stream_) cudaStreamDestroy(cap_stream_);
  if (side_stream_) cudaStreamDestroy(side_stream_);
}

WeightLoader::~WeightLoader() {
  for (void* p : dev_allocs_) cudaFree(p);
}

int WeightLoader::load_weight(const char* name, const float* host, const int64_t* shape, int rank) {
  PBE_REQUIRE(!finalized_, "weights already finalized");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < rank; ++i) {
    t.shape.push_back(shape[i]);
    n *= static_cast<size_t>(shape[i]);
  }
  t.data.assign(host, host + n);
  host_[name] = std::move(t);
  return 0;
}

const HostTensor* WeightLoader::find(const std::string& name) {
  auto it = host_.find(name);
  return it == host_.end() ? nullptr : &it->second;
}

int WeightLoader::get(const std::string& name, const HostTensor** out) {
  *out = find(name);
  if (*out == nullptr) {
    set_error("missing weight: " + name);
    return -4;
  }
  return 0;
}

int WeightLoader::upload_f32(const std::vector<float>& v, float** dst) {
  void* p = nullptr;
  PBE_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(float)));
  dev_allocs_.push_back(p);
  PBE_CHECK_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  *dst = static_cast<float*>(p);
  return 0;
}

// fp32 -> fp16, round to nearest even, saturating at +-65504 (the same rule as the device conversions)
static uint16_t f32_to_f16_rn_sat(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t a = u & 0x7fffffffu;
  if (a > 0x7f800000u) return static_cast<uint16_t>(sign | 0x7e00u);              // NaN
  if (a >= 0x477ff000u) return static_cast<uint16_t>(sign | 0x7bffu);             // >= 65520 rounds past the largest finite
  if (a < 0x33000001u) return static_cast<uint16_t>(sign);                        // < 2^-25: rounds to zero
  const int e = static_cast<int>(a >> 23) - 127;
  uint32_t m = (a & 0x7fffffu) | 0x800000u;
  if (e < -14) {                                                                  // subnormal result
    const int sh = 13 + (-14 - e);
    const uint32_t q = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
    return static_cast<uint16_t>(sign | (q + ((rem > half || (rem == half && (q & 1u))) ? 1u : 0u)));
  }
  const uint32_t q = (static_cast<uint32_t>(e + 15) << 10) | ((m >> 13) & 0x3ffu), rem = m & 0x1fffu;
  return static_cast<uint16_t>(sign | (q + ((rem > 0x1000u || (rem == 0x1000u && (q & 1u))) ? 1u : 0u)));
}

static float f16_bits_to_f32(uint16_t h) {
  const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  uint32_t u;
  if (e == 0) {
    if (m == 0) u = sign;
    else {   // subnormal: renormalise
      int sh = 0;
      uint32_t mm = m;
      while (!(mm & 0x400u)) { mm <<= 1; ++sh; }
      u = sign | (static_cast<uint32_t>(127 - 15 - sh + 1) << 23) | ((mm & 0x3ffu) << 13);
    }
  } else if (e == 31) u = sign | 0x7f800000u | (m << 13);
  else u = sign | ((e + 112u) << 23) | (m << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

float WeightLoader::round_operand(float v) const {
  if (fmt_f16_) return f16_bits_to_f32(f32_to_f16_rn_sat(v));
  const uint32_t u = static_cast<uint32_t>(f32_to_bf16_rn(v)) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// LayerNorm folded into the linear layer behind it: y = W LN(x) + b with LN(x) = (x - mean) rstd gamma + beta becomes
//   y = rstd * (W'' x) + b',   W''[o][i] = W[o][i] gamma[i] - mean_i(W[o][:] gamma[:]),   b'[o] = b[o] + sum_i W[o][i] beta[i]
// (rows of W'' sum to zero, so W'' x = W'' (x - mean): the GEMM runs on the raw tensor and only rstd is left for its epilogue).
void WeightLoader::fold_layernorm(std::vector<float>& w, std::vector<float>& bias, const std::vector<float>& gamma,
                                  const std::vector<float>& beta, int rows, int c) const {
  bias.resize(rows, 0.0f);
  for (int o = 0; o < rows; ++o) {
    float* wr = &w[static_cast<size_t>(o) * c];
    double b = bias[o], cs = 0.0;
    for (int i = 0; i < c; ++i) {
      b += static_cast<double>(beta[i]) * wr[i];
      wr[i] *= gamma[i];
      cs += wr[i];
    }
    const float centre = static_cast<float>(cs / c);
    for (int i = 0; i < c; ++i) wr[i] -= centre;
    bias[o] = static_cast<float>(b);
  }
}

// GEMM weights in the library's operand format (fp16 by default, internal.h: operand_f16)
int WeightLoader::upload_bf16(const std::vector<float>& v, bf16** dst) {
  std::vector<uint16_t> h(v.size());
  if (fmt_f16_) for (size_t i = 0; i < v.size(); ++i) h[i] = f32_to_f16_rn_sat(v[i]);
  else for (size_t i = 0; i < v.size(); ++i) h[i] = f32_to_bf16_rn(v[i]);
  void* p = nullptr;
  PBE_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(h.size(), 1) * sizeof(uint16_t)));
  dev_allocs_.push_back(p);
  PBE_CHECK_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  *dst = static_cast<bf16*>(p);
  return 0;
}

// conv / linear weight `prefix.weight` ([O,I,k,k] or [O,I]) (+ `prefix.bias`) -> [k*k][O][I_pad] bf16
int WeightLoader::make_conv(const std::string& prefix, int k, int cin, int cout, ConvW* w, int cin_pad, int cout_pad) {
  const HostTensor* W;
  int rc = get(prefix + ".weight", &W);
  if (rc) return rc;
  if (cin_pad == 0) cin_pad = cin;
  const size_t expect = static_cast<size_t>(cout) * cin * k * k;
  if (W->data.size() != expect) {
    set_error("weight " + prefix + ".weight has " + std::to_string(W->data.size()) + " elements, expected " +
              std::to_string(expect));
    return -4;
  }
  if (cout_pad == 0) cout_pad = cout;   // zero output rows (e.g. a 3-channel conv_out run as 4 columns)
  std::vector<float> packed(static_cast<size_t>(k) * k * cout_pad * cin_pad, 0.0f);
  for (int o = 0; o < cout; ++o)
    for (int i = 0; i < cin; ++i)
      for (int t = 0; t < k * k; ++t)
        packed[(static_cast<size_t>(t) * cout_pad + o) * cin_pad + i] = W->data[(static_cast<size_t>(o) * cin + i) * k * k + t];
  rc = upload_bf16(packed, &w->w);
  if (rc) return rc;
  w->cin = cin; w->cin_pad = cin_pad; w->cout = cout_pad; w->k = k;
  const HostTensor* B = find(prefix + ".bias");
  if (B) {
    if (static_cast<int>(B->data.size()) != cout) {
      set_error("bias " + prefix + ".bias has wrong size");
      return -4;
    }
    std::vector<float> bp(B->data);
    bp.resize(cout_pad, 0.0f);
    rc = upload_f32(bp, &w->b);
    if (rc) return rc;
  }
  return 0;
}

// "nearest-2x upsample, then 3x3 conv with padding 1" (Upsample.forward, openaimodel.py:109-119) == four 2x2 convs on the
// LOW-resolution tensor, one per output-pixel parity (a, b): output row 2i+a sees upsampled rows 2i+a-1 .. 2i+a+1, which are
// the low-resolution rows {i-1, i, i} (a = 0) or {i, i, i+1} (a = 1) -- taps that read the same source row are summed
// (in fp32, before the one rounding to the operand format).  Same for columns.  [phase = 2a+b][tap = 2*th+tw][cout][cin].
int WeightLoader::make_upconv_phases(const std::string& prefix, int cin, int cout, ConvW* w) {
  const HostTensor* W;
  int rc = get(prefix + ".weight", &W);
  if (rc) return rc;
  if (W->data.size() != static_cast<size_t>(cout) * cin * 9) { set_error("weight " + prefix + ".weight has the wrong size"); return -4; }
  std::vector<float> packed(static_cast<size_t>(16) * cout * cin, 0.0f);
  auto src_taps = [](int parity, int t, int* lo, int* hi) {   // which of the 3 original taps fold into effective tap t
    if (parity == 0) { if (t == 0) { *lo = 0; *hi = 0; } else { *lo = 1; *hi = 2; } }
    else             { if (t == 0) { *lo = 0; *hi = 1; } else { *lo = 2; *hi = 2; } }
  };
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int th = 0; th < 2; ++th)
        for (int tw = 0; tw < 2; ++tw) {
          int h0, h1, w0, w1;
          src_taps(a, th, &h0, &h1);
          src_taps(b, tw, &w0, &w1);
          float* dst = packed.data() + (static_cast<size_t>((a * 2 + b) * 4 + th * 2 + tw) * cout) * cin;
          for (int o = 0; o < cout; ++o)
            for (int i = 0; i < cin; ++i) {
              float acc = 0.0f;
              for (int kh = h0; kh <= h1; ++kh)
                for (int kw = w0; kw <= w1; ++kw) acc += W->data[(static_cast<size_t>(o) * cin + i) * 9 + kh * 3 + kw];
              dst[static_cast<size_t>(o) * cin + i] = acc;
            }
        }
  return upload_bf16(packed, &w->wup);
}

int WeightLoader::make_norm(const std::string& prefix, int c, NormW* n) {
  const HostTensor *G, *B;
  int rc = get(prefix + ".weight", &G);
  if (rc) return rc;
  rc = get(prefix + ".bias", &B);
  if (rc) return rc;
  if (static_cast<int>(G->data.size()) != c || static_cast<int>(B->data.size()) != c) {
    set_error("norm " + prefix + " has wrong size");
    return -4;
  }
  rc = upload_f32(G->data, &n->g);
  if (rc) return rc;
  rc = upload_f32(B->data, &n->b);
  if (rc) return rc;
  n->c = c;
  return 0;
}

// Mirrors UNetModel.__init__ (openaimodel.py:558-834) for use_spatial_transformer=True, transformer_depth=1,
// resblock_updown=False, conv_resample=True, legacy=False.
int Engine::finalize() {
  PBE_REQUIRE(!finalized_, "already finalized");
  const int mc = cfg_.model_channels;
  const int ted = 4 * mc;
  PBE_REQUIRE(mc % 64 == 0, "model_channels must be a multiple of 64");
  PBE_REQUIRE(cfg_.num_levels >= 1 && cfg_.num_levels <= 8, "num_levels");
  PBE_REQUIRE(cfg_.in_channels <= 64, "in_channels <= 64");
  int rc;

  std::vector<float> emb_w, emb_b;  // concatenated emb_layers
  auto add_res = [&](const std::string& pfx, int cin, int cout) -> int {
    ResW r;
    r.cin = cin; r.cout = cout;
    int e;
    if ((e = make_norm(pfx + ".in_layers.0", cin, &r.gn1))) return e;
    if ((e = make_conv(pfx + ".in_layers.2", 3, cin, cout, &r.conv1))) return e;
    if ((e = make_norm(pfx + ".out_layers.0", cout, &r.gn2))) return e;
    if ((e = make_conv(pfx + ".out_layers.3", 3, cout, cout, &r.conv2))) return e;
    r.has_skip = (cin != cout);
    if (r.has_skip && (e = make_conv(pfx + ".skip_connection", 1, cin, cout, &r.skip))) return e;
    const HostTensor *EW, *EB;
    if ((e = get(pfx + ".emb_layers.1.weight", &EW))) return e;
    if ((e = get(pfx + ".emb_layers.1.bias", &EB))) return e;
    if (EW->data.size() != static_cast<size_t>(cout) * ted || EB->data.size() != static_cast<size_t>(cout)) {
      set_error("emb_layers of " + pfx + " have wrong size");
      return -4;
    }
    r.emb_off = static_cast<int>(emb_b.size());
    emb_w.insert(emb_w.end(), EW->data.begin(), EW->data.end());
    emb_b.insert(emb_b.end(), EB->data.begin(), EB->data.end());
    res_.push_back(r);
    return 0;
  };
  auto add_st = [&](const std::string& pfx, int c) -> int {
    STW s;
    s.c = c; s.heads = cfg_.num_heads; s.d = c / cfg_.num_heads;
    PBE_REQUIRE(c % cfg_.num_heads == 0 && s.d % 8 == 0, "head dim must be a multiple of 8");
    const std::string tb = pfx + ".transformer_blocks.0";
    int e;
    if ((e = make_norm(pfx + ".norm", c, &s.gn))) return e;
    if ((e = make_conv(pfx + ".proj_in", 1, c, c, &s.proj_in))) return e;
    if ((e = make_norm(tb + ".norm1", c, &s.ln1))) return e;
    if ((e = make_norm(tb + ".norm3", c, &s.ln3))) return e;
    {  // fused q|k|v projection [3C, C]
      const HostTensor *Q, *K, *V;
      if ((e = get(tb + ".attn1.to_q.weight", &Q))) return e;
      if ((e = get(tb + ".attn1.to_k.weight", &K))) return e;
      if ((e = get(tb + ".attn1.to_v.weight", &V))) return e;
      const size_t cc = static_cast<size_t>(c) * c;
      if (Q->data.size() != cc || K->data.size() != cc || V->data.size() != cc) {
        set_error("attn1 q/k/v of " + pfx + " have wrong size");
        return -4;
      }
      std::vector<float> w;
      w.reserve(3 * cc);
      w.insert(w.end(), Q->data.begin(), Q->data.end());
      w.insert(w.end(), K->data.begin(), K->data.end());
      w.insert(w.end(), V->data.begin(), V->data.end());
      {   // norm1 folded into the projection.  to_q / to_k / to_v have no bias of their own; beta's image under them is:
          //   q: a real bias (kept: columns [0, c) of the GEMM);
          //   k: q . b_k is constant over the keys of a query row and cancels in the softmax -> dropped;
          //   v: sum_j p_j (v_j + b_v) = (sum_j p_j v_j) + b_v -> leaves the attention untouched and becomes W_out b_v in
          //      attn1.to_out's bias (added to the host copy here, before to_out is repacked below).
        const HostTensor *G, *Bt;
        if ((e = get(tb + ".norm1.weight", &G))) return e;
        if ((e = get(tb + ".norm1.bias", &Bt))) return e;
        std::vector<float> bias;
        fold_layernorm(w, bias, G->data, Bt->data, 3 * c, c);
        std::vector<float> bq(bias.begin(), bias.begin() + c);
        bq.resize(3 * c, 0.0f);
        if ((e = upload_f32(bq, &s.qkv.b))) return e;
        const HostTensor* WO;
        if ((e = get(tb + ".attn1.to_out.0.weight", &WO))) return e;
        auto ob = host_.find(tb + ".attn1.to_out.0.bias");
        if (ob == host_.end() || ob->second.data.size() != static_cast<size_t>(c) || WO->data.size() != cc) {
          set_error("attn1.to_out of " + pfx + " has wrong size");
          return -4;
        }
        for (int o = 0; o < c; ++o) {
          double acc = ob->second.data[o];
          for (int i = 0; i < c; ++i) acc += static_cast<double>(WO->data[static_cast<size_t>(o) * c + i]) * bias[2 * c + i];
          ob->second.data[o] = static_cast<float>(acc);
        }
      }
      if ((e = upload_bf16(w, &s.qkv.w))) return e;
      s.qkv.cin = s.qkv.cin_pad = c; s.qkv.cout = 3 * c; s.qkv.k = 1;
    }
    if ((e = make_conv(tb + ".attn1.to_out.0", 1, c, c, &s.to_out))) return e;
    {  // GEGLU projection [8C, C]: rows [0,4C) value, [4C,8C) gate -> interleave per 128-column tile (64 value + 64 gate)
      const HostTensor *W, *B;
      if ((e = get(tb + ".ff.net.0.proj.weight", &W))) return e;
      if ((e = get(tb + ".ff.net.0.proj.bias", &B))) return e;
      const int inner = 4 * c;
      if (W->data.size() != static_cast<size_t>(2 * inner) * c || B->data.size() != static_cast<size_t>(2 * inner)) {
        set_error("GEGLU proj of " + pfx + " has wrong size");
        return -4;
      }
      PBE_REQUIRE(inner % 128 == 0, "GEGLU inner dim % 128");
      std::vector<float> w(W->data.size()), b(B->data.size());
      for (int t = 0; t < inner / 128; ++t)
        for (int j = 0; j < 128; ++j) {
          const int src_v = t * 128 + j, src_g = inner + t * 128 + j;
          const int dst_v = t * 256 + j, dst_g = t * 256 + 128 + j;
          memcpy(&w[static_cast<size_t>(dst_v) * c], &W->data[static_cast<size_t>(src_v) * c], c * sizeof(float));
          memcpy(&w[static_cast<size_t>(dst_g) * c], &W->data[static_cast<size_t>(src_g) * c], c * sizeof(float));
          b[dst_v] = B->data[src_v];
          b[dst_g] = B->data[src_g];
        }
      {   // norm3 folded into the GEGLU projection
        const HostTensor *G, *Bt;
        if ((e = get(tb + ".norm3.weight", &G))) return e;
        if ((e = get(tb + ".norm3.bias", &Bt))) return e;
        fold_layernorm(w, b, G->data, Bt->data, 2 * inner, c);
      }
      if ((e = upload_bf16(w, &s.ff1.w))) return e;
      if ((e = upload_f32(b, &s.ff1.b))) return e;
      s.ff1.cin = s.ff1.cin_pad = c; s.ff1.cout = 2 * inner; s.ff1.k = 1;
    }
    if ((e = make_conv(tb + ".ff.net.2", 1, 4 * c, c, &s.ff2))) return e;
    if ((e = make_conv(pfx + ".proj_out", 1, c, c, &s.proj_out))) return e;
    if (ff_proj_merge_) {
      // x3 = W_ff2 g + b_ff2 + x2 has one reader, proj_out, and nothing non-linear in between (attention.py:276, 335-336):
      //   proj_out(x3) = (W_po W_ff2) g + W_po x2 + (W_po b_ff2 + b_po)
      // = one GEMM with K = 4c + c over the buffer [ g | x2 ] -- the same FLOPs as the two, one epilogue, one launch and
      // one [M, c] round trip fewer, and x3 is never rounded to 16 bits.  The product is taken in fp64 on the device and
      // rounded once to the operand format.
      const HostTensor *WF, *BF, *WP, *BP;
      if ((e = get(tb + ".ff.net.2.weight", &WF))) return e;
      if ((e = get(tb + ".ff.net.2.bias", &BF))) return e;
      if ((e = get(pfx + ".proj_out.weight", &WP))) return e;
      if ((e = get(pfx + ".proj_out.bias", &BP))) return e;
      const size_t inner = static_cast<size_t>(4) * c;
      if (WF->data.size() != inner * c || WP->data.size() != static_cast<size_t>(c) * c || BF->data.size() != static_cast<size_t>(c) ||
          BP->data.size() != static_cast<size_t>(c)) {
        set_error("ff.net.2 / proj_out of " + pfx + " have wrong size");
        return -4;
      }
      float *dA = nullptr, *dB = nullptr, *dC = nullptr;
      PBE_CHECK_CUDA(cudaMalloc(&dA, WP->data.size() * sizeof(float)));
      PBE_CHECK_CUDA(cudaMalloc(&dB, WF->data.size() * sizeof(float)));
      PBE_CHECK_CUDA(cudaMalloc(&dC, WF->data.size() * sizeof(float)));
      PBE_CHECK_CUDA(cudaMemcpy(dA, WP->data.data(), WP->data.size() * sizeof(float), cudaMemcpyHostToDevice));
      PBE_CHECK_CUDA(cudaMemcpy(dB, WF->data.data(), WF->data.size() * sizeof(float), cudaMemcpyHostToDevice));
      if ((e = launch_matmul_f64acc(dA, dB, dC, c, c, static_cast<int>(inner), nullptr))) return e;
      std::vector<float> prod(WF->data.size());
      PBE_CHECK_CUDA(cudaMemcpy(prod.data(), dC, prod.size() * sizeof(float), cudaMemcpyDeviceToHost));
      PBE_CHECK_CUDA(cudaFree(dA));
      PBE_CHECK_CUDA(cudaFree(dB));
      PBE_CHECK_CUDA(cudaFree(dC));
      const size_t kk = inner + c;
      std::vector<float> w(static_cast<size_t>(c) * kk), b(c);
      for (int o = 0; o < c; ++o) {
        memcpy(&w[o * kk], &prod[o * inner], inner * sizeof(float));
        memcpy(&w[o * kk + inner], &WP->data[static_cast<size_t>(o) * c], c * sizeof(float));
        double acc = BP->data[o];
        for (int i = 0; i < c; ++i) acc += static_cast<double>(WP->data[static_cast<size_t>(o) * c + i]) * BF->data[i];
        b[o] = static_cast<float>(acc);
      }
      if ((e = upload_bf16(w, &s.ffproj.w))) return e;
      if ((e = upload_f32(b, &s.ffproj.b))) return e;
      s.ffproj.cin = s.ffproj.cin_pad = static_cast<int>(kk); s.ffproj.cout = c; s.ffproj.k = 1;
    }
    {  // single-key cross-attention folds to to_out(to_v(ctx)); to_q / to_k / norm2 are dead (attention.py:207-230)
      const HostTensor *V2, *O2, *B2;
      if ((e = get(tb + ".attn2.to_v.weight", &V2))) return e;
      if ((e = get(tb + ".attn2.to_out.0.weight", &O2))) return e;
      if ((e = get(tb + ".attn2.to_out.0.bias", &B2))) return e;
      if (V2->data.size() != static_cast<size_t>(c) * cfg_.context_dim || O2->data.size() != static_cast<size_t>(c) * c) {
        set_error("attn2 weights of " + pfx + " have wrong size");
        return -4;
      }
      if ((e = upload_f32(V2->data, &s.wv2))) return e;
      if ((e = upload_f32(O2->data, &s.wo2))) return e;
      if ((e = upload_f32(B2->data, &s.bo2))) return e;
    }
    s.ctx_vec_off = ctx_total_;
    ctx_total_ += c;
    st_.push_back(s);
    return 0;
  };
  auto is_attn = [&](int ds) {
    for (int i = 0; i < cfg_.num_attention_resolutions; ++i)
      if (cfg_.attention_resolutions[i] == ds) return true;
    return false;
  };

  // time_embed
  {
    const HostTensor *W0, *B0, *W1, *B1;
    if ((rc = get("time_embed.0.weight", &W0))) return rc;
    if ((rc = get("time_embed.0.bias", &B0))) return rc;
    if ((rc = get("time_embed.2.weight", &W1))) return rc;
    if ((rc = get("time_embed.2.bias", &B1))) return rc;
    PBE_REQUIRE(W0->data.size() == static_cast<size_t>(ted) * mc && W1->data.size() == static_cast<size_t>(ted) * ted,
                "time_embed sizes");
    if ((rc = upload_f32(W0->data, &te_w0_))) return rc;
    if ((rc = upload_f32(B0->data, &te_b0_))) return rc;
    if ((rc = upload_f32(W1->data, &te_w1_))) return rc;
    if ((rc = upload_f32(B1->data, &te_b1_))) return rc;
  }

  // input blocks
  {
    ConvW c;
    if ((rc = make_conv("input_blocks.0.0", 3, cfg_.in_channels, mc, &c, 64))) return rc;
    convs_.push_back(c);
    modules_.push_back({Module::CONV_IN, static_cast<int>(convs_.size()) - 1, false, true});
  }
  std::vector<int> chans = {mc};
  int ch = mc, ds = 1, ib = 1;
  for (int level = 0; level < cfg_.num_levels; ++level) {
    const int mult = cfg_.channel_mult[level];
    for (int r = 0; r < cfg_.num_res_blocks; ++r) {
      const std::string pfx = "input_blocks." + std::to_string(ib);
      if ((rc = add_res(pfx + ".0", ch, mult * mc))) return rc;
      ch = mult * mc;
      const bool attn = is_attn(ds);
      modules_.push_back({Module::RES, static_cast<int>(res_.size()) - 1, false, !attn});
      if (attn) {
        if ((rc = add_st(pfx + ".1", ch))) return rc;
        modules_.push_back({Module::ST, static_cast<int>(st_.size()) - 1, false, true});
      }
      chans.push_back(ch);
      ++ib;
    }
    if (level != cfg_.num_levels - 1) {
      ConvW c;
      if ((rc = make_conv("input_blocks." + std::to_string(ib) + ".0.op", 3, ch, ch, &c))) return rc;
      convs_.push_back(c);
      modules_.push_back({Module::DOWN, static_cast<int>(convs_.size()) - 1, false, true});
      chans.push_back(ch);
      ++ib;
      ds *= 2;
    }
  }
  // middle
  if ((rc = add_res("middle_block.0", ch, ch))) return rc;
  modules_.push_back({Module::RES, static_cast<int>(res_.size()) - 1, false, false});
  if ((rc = add_st("middle_block.1", ch))) return rc;
  modules_.push_back({Module::ST, static_cast<int>(st_.size()) - 1, false, false});
  if ((rc = add_res("middle_block.2", ch, ch))) return rc;
  modules_.push_back({Module::RES, static_cast<int>(res_.size()) - 1, false, false});
  // output blocks
  int ob = 0;
  for (int level = cfg_.num_levels - 1; level >= 0; --level) {
    const int mult = cfg_.channel_mult[level];
    for (int i = 0; i <= cfg_.num_res_blocks; ++i) {
      const int ich = chans.back();
      chans.pop_back();
      const std::string pfx = "output_blocks." + std::to_string(ob);
      if ((rc = add_res(pfx + ".0", ch + ich, mc * mult))) return rc;
      ch = mc * mult;
      modules_.push_back({Module::RES, static_cast<int>(res_.size()) - 1, true, false});
      int sub = 1;
      if (is_attn(ds)) {
        if ((rc = add_st(pfx + ".1", ch))) return rc;
        modules_.push_back({Module::ST, static_cast<int>(st_.size()) - 1, false, false});
        sub = 2;
      }
      if (level && i == cfg_.num_res_blocks) {
        ConvW c;
        if ((rc = make_conv(pfx + "." + std::to_string(sub) + ".conv", 3, ch, ch, &c))) return rc;
        if ((rc = make_upconv_phases(pfx + "." + std::to_string(sub) + ".conv", ch, ch, &c))) return rc;
        convs_.push_back(c);
        modules_.push_back({Module::UP, static_cast<int>(convs_.size()) - 1, false, false});
        ds /= 2;
      }
      ++ob;
    }
  }
  // out
  if ((rc = make_norm("out.0", ch, &out_norm_))) return rc;
  {
    ConvW c;
    if ((rc = make_conv("out.2", 3, mc, cfg_.out_channels, &c))) return rc;
    convs_.push_back(c);
    modules_.push_back({Module::OUT, static_cast<int>(convs_.size()) - 1, false, false});
  }
  emb_total_ = static_cast<int>(emb_b.size());
  if ((rc = upload_bf16(emb_w, &emb_w_))) return rc;
  if ((rc = upload_f32(emb_b, &emb_b_))) return rc;

  {
    void* p = nullptr;
    PBE_CHECK_CUDA(cudaMalloc(&p, static_cast<size_t>(MAX_BC) * ctx_total_ * sizeof(float)));
    dev_allocs_.push_back(p);
    ctx_vecs_ = static_cast<float*>(p);
    PBE_CHECK_CUDA(cudaMemset(p, 0, static_cast<size_t>(MAX_BC) * ctx_total_ * sizeof(float)));
    int cmax = 0;
    for (const auto& s : st_) cmax = std::max(cmax, s.c);
    PBE_CHECK_CUDA(cudaMalloc(&p, static_cast<size_t>(MAX_BC) * cmax * sizeof(float)));
    dev_allocs_.push_back(p);
    ctx_tmp_ = static_cast<float*>(p);
    std::vector<float> ones(cmax, 1.0f), zeros(cmax, 0.0f);
    if ((rc = upload_f32(ones, &ln_ones_))) return rc;
    if ((rc = upload_f32(zeros, &ln_zeros_))) return rc;
  }
  host_.clear();
  finalized_ = true;
  return 0;
}

// K4: v[b] = to_out(to_v(ctx[b])) + bias for every SpatialTransformer; constant over all timesteps of a sample() call.
int Engine::set_context(const float* ctx_dev, int Bc, cudaStream_t stream) {
  PBE_REQUIRE(finalized_, "weights not finalized");
  PBE_REQUIRE(Bc >= 1 && Bc <= MAX_BC, "context batch out of range");
  for (const auto& s : st_) {
    int rc = launch_small_linear(ctx_dev, s.wv2, nullptr, ctx_tmp_, Bc, cfg_.context_dim, s.c, 0, 0, stream);
    if (rc) return rc;
    // rows of ctx_vecs_ have stride ctx_total_: write through a strided view by launching per-ST into a packed tmp
    // then scattering is avoided by letting small_linear write with O = s.c into a [Bc, s.c] block and the GEMM
    // epilogue reading it with rowbias_ld = s.c; blocks are laid out back to back: offset = MAX_BC * ctx_vec_off.
    rc = launch_small_linear(ctx_tmp_, s.wo2, s.bo2, ctx_vecs_ + static_cast<size_t>(MAX_BC) * s.ctx_vec_off, Bc, s.c,
                             s.c, 0, 0, stream);
    if (rc) return rc;
  }
  ctx_Bc_ = Bc;
  return 0;
}

int Engine::build(Prepared& P, bool dry) {
  const int BcFull = P.Bc, H0 = P.H, W0 = P.W;
  // CFG pair plan: ops are emitted for the B = Bc/2 shared samples until the context enters (first SpatialTransformer's
  // attn1.to_out), then for all Bc samples; `Bc` is the batch of the ops currently being emitted.
  const bool pair = P.pair;
  bool diverged = !pair;
  int Bc = pair ? BcFull / 2 : BcFull;
  const int mc = cfg_.model_channels, ted = 4 * mc;
  P.persist.reset(dry);
  P.scratch.reset(dry);
  P.ops.clear();
  P.op_flags.clear();
  P.op_names.clear();
  int launches = 0;
  auto PA = [&](size_t bytes) { return P.persist.alloc(bytes); };
  auto SA = [&](size_t bytes) { return P.scratch.alloc(bytes); };
  P.op_family.clear();
  P.op_flops.clear();
  P.op_bytes.clear();
  auto add_op_meta = [&](const std::string& name, int nlaunch, std::function<int(cudaStream_t)> fn,
                         const std::string& family, double flops, double bytes) {
    if (!dry) {
      P.ops.push_back(std::move(fn));
      P.op_flags.push_back(0);
      P.op_names.push_back(name);
      P.op_family.push_back(family);
      P.op_flops.push_back(flops);
      P.op_bytes.push_back(bytes);
    }
    launches += nlaunch;
  };
  auto add_op = [&](const std::string& name, int nlaunch, std::function<int(cudaStream_t)> fn) {
    add_op_meta(name, nlaunch, std::move(fn), "misc", 0.0, 0.0);
  };
  int err = 0;
  auto add_gemm = [&](const std::string& name, ConvGemmDesc d) {
    if (!diverged) d.split_batch = 2 * d.Nb;   // CFG-pair prefix: same split-K decision (summation order) as the full batch
    const size_t ws_bytes = (d.stats_out || d.up_phase) ? 0 : gemm_splitk_ws_bytes(d);
    d.splitk_ws = ws_bytes ? static_cast<float*>(SA(ws_bytes)) : nullptr;
    if (dry) { launches += ws_bytes ? 2 : 1; return; }
    auto plan = std::make_shared<GemmPlan>();
    int rc = build_gemm_plan(d, plan.get());
    if (rc && !err) { err = rc; last_error = std::string(get_error()) + " [" + name + "]"; }
    const double Mrows = static_cast<double>(d.Nb) * (d.H / d.stride) * (d.W / d.stride);
    const double kreal = static_cast<double>(d.c_real ? d.c_real : d.C) * d.ksize * d.ksize;
    const double flops = 2.0 * Mrows * d.Cout * kreal;
    const double out_cols = (d.mode == EPI_GEGLU) ? d.Cout / 2.0 : d.Cout;
    double bytes = static_cast<double>(d.Nb) * d.H * d.W * d.C * 2.0 + static_cast<double>(d.Cout) * d.C * d.ksize * d.ksize * 2.0;
    if (d.out_f32) bytes += Mrows * out_cols * 4.0;
    if (d.out_bf16 || d.out_vt) bytes += Mrows * out_cols * 2.0;
    if (d.residual) bytes += Mrows * out_cols * 4.0;
    if (d.residual16) bytes += Mrows * out_cols * 2.0;
    add_op_meta(name, ws_bytes ? 2 : 1, [plan](cudaStream_t s) { return launch_gemm_plan(*plan, s); }, "conv_gemm", flops,
                bytes);
  };
  auto add_gn = [&](const std::string& name, GroupNormArgs a) {
    a.partial = static_cast<float*>(SA(static_cast<size_t>(gn_workspace_floats(a.Nb, a.HW, a.C0 + a.C1)) * sizeof(float)));
    const double n = static_cast<double>(a.Nb) * a.HW * (a.C0 + a.C1);
    const int launches = gn_num_launches(a);
    const double in_b = a.in16 ? 2.0 : 4.0;
    add_op_meta(name, launches, [a](cudaStream_t s) { return launch_groupnorm(a, s); }, "groupnorm", 0.0,
                n * ((launches == 3 ? 2.0 * in_b : in_b) + 2.0 + (a.raw ? 2.0 : 0.0)));
  };

  P.x_stage = static_cast<float*>(PA(static_cast<size_t>(BcFull) * cfg_.in_channels * H0 * W0 * sizeof(float)));
  P.t_stage = static_cast<int64_t*>(PA(static_cast<size_t>(BcFull) * sizeof(int64_t)));
  P.eps_stage = static_cast<float*>(PA(static_cast<size_t>(BcFull) * cfg_.out_channels * H0 * W0 * sizeof(float)));

  // ---- timestep embedding path (K8) ----
  float* t_emb = static_cast<float*>(PA(static_cast<size_t>(BcFull) * mc * sizeof(float)));
  float* t_hid = static_cast<float*>(PA(static_cast<size_t>(BcFull) * ted * sizeof(float)));
  float* emb = static_cast<float*>(PA(static_cast<size_t>(BcFull) * ted * sizeof(float)));
  float* emb_silu = static_cast<float*>(PA(static_cast<size_t>(BcFull) * ted * sizeof(float)));
  float* emb_all = static_cast<float*>(PA(static_cast<size_t>(BcFull) * emb_total_ * sizeof(float)));
  {
    const int64_t* tp = P.t_stage;
    float *w0 = te_w0_, *b0 = te_b0_, *w1 = te_w1_, *b1 = te_b1_;
    bf16* emb_silu16 = static_cast<bf16*>(PA(static_cast<size_t>(BcFull) * ted * sizeof(bf16)));
    add_op("timestep_embedding", 1, [=](cudaStream_t s) { return launch_timestep_embedding(tp, t_emb, BcFull, mc, s); });
    add_op("time_embed.0+silu", 1,
           [=](cudaStream_t s) { return launch_small_linear(t_emb, w0, b0, t_hid, BcFull, mc, ted, 0, 1, s); });
    add_op("time_embed.2", 1,
           [=](cudaStream_t s) { return launch_small_linear(t_hid, w1, b1, emb, BcFull, ted, ted, 0, 0, s, emb_silu); });
    // every ResBlock's emb_layers = Linear(SiLU(emb)) (openaimodel.py:218-224): SiLU once, all 22 Linears as ONE
    // [BcFull, 1280] x [1280, 20160] GEMM on the tensor cores (a 52 MB bf16 weight stream; as batched fp32 GEMVs it was
    // latency bound at 185 us per call)
    add_op("emb_silu.bf16", 1,
           [=](cudaStream_t s) { return launch_cast_bf16(emb_silu, emb_silu16, static_cast<size_t>(BcFull) * ted, s); });
    {
      ConvGemmDesc d{};
      d.act = emb_silu16; d.Nb = BcFull; d.H = 1; d.W = 1; d.C = ted; d.ksize = 1; d.stride = 1;
      d.wt = emb_w_; d.Cout = emb_total_; d.mode = EPI_STD; d.bias = emb_b_; d.out_f32 = emb_all;
      add_gemm("emb_layers(all)", d);
    }
  }

  // the ops so far depend on t only: in the captured graph they form a branch beside conv_in / the first GroupNorm and join
  // the main line at the first consumer of emb_all, the first ResBlock's conv1 (run_op_list)
  if (!dry) {
    for (uint8_t& f : P.op_flags) f = OP_SIDE;
    P.op_flags[0] |= OP_FORK;
  }
  bool emb_joined = false, dup_on_side = false;
  // latency-bound plans (CFG batch <= 4) also run a ResBlock's 1x1 skip convolution beside its conv1 / GroupNorm (it reads
  // the block's input only); a full GPU gains nothing from the concurrency
  const bool skip_on_side = BcFull <= 4;
  auto flag_last = [&](uint8_t f) { if (!dry && !P.op_flags.empty()) P.op_flags.back() |= f; };

  // The residual stream between ops (block outputs, skip tensors, ResBlock h, transformer t0 / t1) is 16-bit in the
  // operand format (fp16 by default) -- what the reference itself carries under torch.autocast -- unless PBE_STREAM=fp32:
  // every GEMM then stores 16-bit only (rounded once from its fp32 accumulator + bias + residual), GroupNorm / LayerNorm
  // read 2 bytes per element instead of 4, residuals enter the epilogues as 64-byte rows, and no raw 16-bit side copies
  // are needed for the stride-2 and 1x1 skip convs (the stream IS the operand).  Statistics stay fp32.
  const bool s16 = stream16_;
  const size_t esz = s16 ? sizeof(bf16) : sizeof(float);
  struct Act {
    float* f32;
    bf16* b16;  // fp32 stream: optional 16-bit copy; 16-bit stream: the tensor itself
    int C, H, W;
    bool has16 = false;
    float* stats = nullptr;   // fused GroupNorm statistics written by the producing GEMM (or null)
    bool has_stats = false;
  };
  auto new_act = [&](bool persistent, size_t elems, int C, int H, int W) {
    Act a{nullptr, nullptr, C, H, W};
    void* ptr = persistent ? PA(elems * esz) : SA(elems * esz);
    if (s16) { a.b16 = static_cast<bf16*>(ptr); a.has16 = true; }
    else a.f32 = static_cast<float*>(ptr);
    return a;
  };
  auto as_f32ptr = [&](const Act& a) { return s16 ? reinterpret_cast<const float*>(a.b16) : a.f32; };   // GroupNormArgs / LayerNorm source
  auto set_out = [&](ConvGemmDesc& d, const Act& a) {
    if (s16) d.out_bf16 = a.b16;
    else { d.out_f32 = a.f32; d.out_bf16 = a.b16; }
  };
  auto set_res = [&](ConvGemmDesc& d, const Act& a) {
    if (s16) d.residual16 = a.b16;
    else d.residual = a.f32;
  };
  // Attach fused GroupNorm statistics to a GEMM whose output `o` is normalised by the next op.
  auto want_stats = [&](ConvGemmDesc& d, Act& o, bool persistent) {
    d.splitk_ws = nullptr;
    if (!diverged) d.split_batch = 2 * d.Nb;   // as in add_gemm: decide like the full CFG batch would
    if (!gemm_can_fuse_stats(d)) return;
    const size_t rows = static_cast<size_t>(d.Nb) * (d.H / d.stride) * (d.W / d.stride);
    const size_t bytes = rows / 32 * d.Cout * 2 * sizeof(float) * (diverged ? 1 : 2);   // prefix tensors are duplicated later
    o.stats = static_cast<float*>(persistent ? PA(bytes) : SA(bytes));
    o.has_stats = true;
    d.stats_out = o.stats;
  };
  // LayerNorm fold: let GEMM `d` (writing a [rows_total, Cout] 16-bit tensor, of which it may cover a slice) emit per-row
  // statistics partials; returns the number of partials (0: not available -> normalise-only pass)
  auto want_ln_stats = [&](ConvGemmDesc& d, size_t rows_total, float2** stats) -> int {
    if (!s16 || !ln_fold_) return 0;
    const int parts = gemm_ln_parts(d);
    if (parts <= 0 || parts > 16) return 0;
    *stats = static_cast<float2*>(SA(static_cast<size_t>(parts) * rows_total * sizeof(float2)));
    d.ln_stats_out = *stats;
    d.ln_stats_stride = static_cast<long long>(rows_total);
    return parts;
  };
  std::vector<Act> hs;
  Act h{nullptr, nullptr, 0, H0, W0};

  for (size_t mi = 0; mi < modules_.size(); ++mi) {
    const Module& m = modules_[mi];
    const size_t smark = P.scratch.mark();
    const std::string tag = "m" + std::to_string(mi);
    // does the next module consume a raw 16-bit copy of this module's output? (fp32 stream only: a stride-2 conv reads raw activations)
    const bool next_is_down = !s16 && (mi + 1 < modules_.size() && modules_[mi + 1].kind == Module::DOWN);
    switch (m.kind) {
      case Module::CONV_IN: {
        const ConvW& c = convs_[m.idx];
        const size_t M = static_cast<size_t>(Bc) * H0 * W0;
        bf16* xin = static_cast<bf16*>(SA(M * 64 * sizeof(bf16)));
        const float* xs = P.x_stage;
        const int cin = cfg_.in_channels;
        add_op(tag + ".pack_input", 1, [=](cudaStream_t s) { return launch_pack_input(xs, xin, Bc, cin, H0, W0, 64, s); });
        Act o = new_act(true, M * (diverged ? 1 : 2) * c.cout, c.cout, H0, W0);
        ConvGemmDesc d{};
        d.act = xin; d.Nb = Bc; d.H = H0; d.W = W0; d.C = 64; d.c_real = cin; d.ksize = 3; d.stride = 1;
        d.wt = c.w; d.Cout = c.cout; d.mode = EPI_STD; d.bias = c.b;
        set_out(d, o);
        want_stats(d, o, true);
        add_gemm(tag + ".conv_in", d);
        h = o;
        break;
      }
      case Module::RES: {
        const ResW& r = res_[m.idx];
        Act skip{nullptr, nullptr, 0, 0, 0};
        if (m.pop_skip) {
          skip = hs.back();
          hs.pop_back();
        }
        const size_t M = static_cast<size_t>(Bc) * h.H * h.W;
        const int cin = h.C + skip.C;
        if (cin != r.cin) {
          err = -5;
          last_error = "channel mismatch at " + tag;
          return err;
        }
        bf16* a1 = static_cast<bf16*>(SA(M * cin * sizeof(bf16)));
        // the 1x1 skip conv reads the raw (un-normalised) input: with a 16-bit stream and no concat that is h itself
        const bool raw_needed = r.has_skip && !(s16 && skip.C == 0);
        bf16* raw = raw_needed ? static_cast<bf16*>(SA(M * cin * sizeof(bf16))) : nullptr;
        GroupNormArgs g1{};
        g1.x0 = as_f32ptr(h); g1.C0 = h.C; g1.x1 = skip.C ? as_f32ptr(skip) : nullptr; g1.C1 = skip.C; g1.Nb = Bc; g1.HW = h.H * h.W;
        g1.in16 = s16 ? 1 : 0;
        g1.gamma = r.gn1.g; g1.beta = r.gn1.b; g1.eps = 1e-5f; g1.silu = 1; g1.y = a1; g1.raw = raw;
        g1.stats0 = h.has_stats ? h.stats : nullptr;
        g1.stats1 = (skip.C > 0 && skip.has_stats) ? skip.stats : nullptr;
        if (skip.C > 0 && !(h.has_stats && skip.has_stats)) g1.stats0 = g1.stats1 = nullptr;
        add_gn(tag + ".gn1", g1);
        Act resid = h;
        auto emit_skip = [&](bool side) {
          Act sk = new_act(false, M * r.cout, r.cout, h.H, h.W);
          ConvGemmDesc d{};
          d.act = raw_needed ? raw : h.b16; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = cin; d.ksize = 1; d.stride = 1;
          d.wt = r.skip.w; d.Cout = r.cout; d.mode = EPI_STD; d.bias = r.skip.b;
          set_out(d, sk);
          add_gemm(tag + ".skip", d);
          if (side) flag_last(OP_SIDE | OP_FORK);
          resid = sk;
        };
        const bool skip_side = r.has_skip && skip_on_side && graph_lanes_;
        if (skip_side) emit_skip(true);   // a branch of the graph beside conv1 / gn2, joined in front of conv2
        Act h1act = new_act(false, M * r.cout, r.cout, h.H, h.W);
        {
          ConvGemmDesc d{};
          d.act = a1; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = cin; d.ksize = 3; d.stride = 1;
          d.wt = r.conv1.w; d.Cout = r.cout; d.mode = EPI_STD; d.bias = r.conv1.b;
          d.rowbias = emb_all + r.emb_off; d.rowbias_ld = emb_total_;
          set_out(d, h1act);
          want_stats(d, h1act, false);
          add_gemm(tag + ".conv1", d);
          if (!emb_joined) { flag_last(OP_JOIN); emb_joined = true; }   // first reader of emb_all
        }
        bf16* a2 = static_cast<bf16*>(SA(M * r.cout * sizeof(bf16)));
        GroupNormArgs g2{};
        g2.x0 = as_f32ptr(h1act); g2.C0 = r.cout; g2.x1 = nullptr; g2.C1 = 0; g2.Nb = Bc; g2.HW = h.H * h.W;
        g2.in16 = s16 ? 1 : 0;
        g2.gamma = r.gn2.g; g2.beta = r.gn2.b; g2.eps = 1e-5f; g2.silu = 1; g2.y = a2; g2.raw = nullptr;
        g2.stats0 = h1act.has_stats ? h1act.stats : nullptr;
        add_gn(tag + ".gn2", g2);
        if (r.has_skip && !skip_side) emit_skip(false);
        Act o = new_act(true, M * (diverged ? 1 : 2) * r.cout, r.cout, h.H, h.W);
        if (next_is_down) { o.b16 = static_cast<bf16*>(PA(M * (diverged ? 1 : 2) * r.cout * sizeof(bf16))); o.has16 = true; }
        {
          ConvGemmDesc d{};
          d.act = a2; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = r.cout; d.ksize = 3; d.stride = 1;
          d.wt = r.conv2.w; d.Cout = r.cout; d.mode = EPI_STD; d.bias = r.conv2.b;
          set_res(d, resid);
          set_out(d, o);
          want_stats(d, o, true);
          add_gemm(tag + ".conv2", d);
          if (skip_side) flag_last(OP_JOIN);
        }
        h = o;
        break;
      }
      case Module::ST: {
        const STW& s = st_[m.idx];
        const int C = s.c, N = h.H * h.W;
        size_t M = static_cast<size_t>(Bc) * N;
        if (C != h.C) { err = -5; last_error = "channel mismatch at " + tag; return err; }
        float2 *ln1_stats = nullptr, *ln3_stats = nullptr;
        int ln1_parts = 0, ln3_parts = 0;
        bf16* a = static_cast<bf16*>(SA(M * C * sizeof(bf16)));
        GroupNormArgs g{};
        g.x0 = as_f32ptr(h); g.C0 = C; g.x1 = nullptr; g.C1 = 0; g.Nb = Bc; g.HW = N;
        g.in16 = s16 ? 1 : 0;
        g.gamma = s.gn.g; g.beta = s.gn.b; g.eps = 1e-6f; g.silu = 0; g.y = a; g.raw = nullptr;
        g.stats0 = h.has_stats ? h.stats : nullptr;
        add_gn(tag + ".norm", g);
        Act t0 = new_act(false, M * C, C, h.H, h.W);
        {
          ConvGemmDesc d{};
          d.act = a; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = C; d.ksize = 1; d.stride = 1;
          d.wt = s.proj_in.w; d.Cout = C; d.mode = EPI_STD; d.bias = s.proj_in.b;
          set_out(d, t0);
          ln1_parts = want_ln_stats(d, M, &ln1_stats);
          add_gemm(tag + ".proj_in", d);
        }
        // norm1 (attention.py:271): gamma / beta live in the qkv weights; either the row statistics come out of proj_in's
        // epilogue and qkv finishes the normalisation in its own (no pass over the tensor at all), or a normalise-only pass
        bf16* n1 = ln1_parts ? t0.b16 : static_cast<bf16*>(SA(M * C * sizeof(bf16)));
        if (!ln1_parts) {
          const float *gg = ln_ones_, *bb = ln_zeros_;
          const int Mi = static_cast<int>(M);
          const float* src = as_f32ptr(t0);
          const int in16 = s16 ? 1 : 0;
          add_op_meta(tag + ".ln1", 1, [=](cudaStream_t st) { return launch_layernorm(src, gg, bb, n1, Mi, C, 1e-5f, st, nullptr, 0, in16); },
                      "layernorm", 0.0, static_cast<double>(M) * C * (esz + 2.0));
        }
        bf16* qk = static_cast<bf16*>(SA(M * 2 * C * sizeof(bf16)));
        bf16* vt = static_cast<bf16*>(SA(static_cast<size_t>(Bc) * C * vt_pitch(N) * sizeof(bf16)));   // [Bc][C][pitch]
        {
          ConvGemmDesc d{};
          d.act = n1; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = C; d.ksize = 1; d.stride = 1;
          d.wt = s.qkv.w; d.Cout = 3 * C; d.mode = EPI_QKV; d.out_bf16 = qk; d.ld_out = 2 * C; d.out_vt = vt;
          d.qk_cols = 2 * C;
          d.bias = s.qkv.b; d.bias_cols = C;   // only q keeps the image of norm1's beta (finalize)
          if (ln1_parts) { d.ln_in = ln1_stats; d.ln_stride = static_cast<long long>(M); d.ln_parts = ln1_parts; d.ln_eps = 1e-5f; }
          d.out16_bf16 = 1;   // Q | K | V^T feed the flash-attention kernels, which work in bf16 (P lives far below the fp16 range)
          d.block_n = (C % 160 == 0) ? 160 : ((C % 128 == 0) ? 128 : 64);
          add_gemm(tag + ".qkv", d);
        }
        bf16* ao = static_cast<bf16*>(SA(M * C * sizeof(bf16)));
        if (!dry) {
          auto plan = std::make_shared<AttnPlan>();
          int rc = build_attn_plan(qk, vt, ao, Bc, N, s.heads, s.d, plan.get(), true);
          if (rc && !err) { err = rc; last_error = std::string(get_error()) + " [" + tag + ".attn]"; }
          add_op_meta(tag + ".attn1", attn_num_launches(*plan), [plan](cudaStream_t st) { return launch_attn_plan(*plan, st); }, "attention",
                      4.0 * Bc * static_cast<double>(N) * N * C, static_cast<double>(M) * C * 2.0 * 4.0);
        } else {
          launches += 1;
        }
        const size_t M_t1 = M * (diverged ? 1 : 2);   // rows of t1 (the CFG-pair prefix writes it in two halves)
        // ff.net.2 + proj_out as one GEMM (engine.h: ff_proj_merge_): t1 (= x2) lives in columns [4C, 5C) of the buffer whose
        // columns [0, 4C) receive the GEGLU output, so [ g | x2 ] is ONE K = 5C operand.  Needs the LayerNorm fold on t1
        // (its only other reader is then the GEGLU GEMM, through a pitched tensor map).
        bool merge = false;
        if (s16 && ln_fold_ && ff_proj_merge_ && s.ffproj.w != nullptr) {
          ConvGemmDesc probe{};
          probe.Nb = Bc; probe.H = h.H; probe.W = h.W; probe.C = C; probe.ksize = 1; probe.stride = 1; probe.Cout = C; probe.mode = EPI_STD;
          const int parts = gemm_ln_parts(probe);
          merge = parts > 0 && parts <= 16;
        }
        const int catld = 5 * C;
        bf16* cat = merge ? static_cast<bf16*>(SA(M_t1 * catld * sizeof(bf16))) : nullptr;
        Act t1{nullptr, nullptr, C, h.H, h.W};
        if (merge) { t1.b16 = cat + 4 * C; t1.has16 = true; }
        else t1 = new_act(false, M_t1 * C, C, h.H, h.W);
        const size_t t1_ld = merge ? static_cast<size_t>(catld) : static_cast<size_t>(C);
        for (int half = 0; half < (diverged ? 1 : 2); ++half) {
          // x1 = to_out(attn) + b + x ; x2 = x1 + to_out2(to_v2(ctx))  (single-key cross-attention, folded).
          // CFG pair plan: this is where the context enters -- the shared activations feed one GEMM per half, each
          // with its own rows of the per-sample context bias, writing its half of the full batch.
          ConvGemmDesc d{};
          d.act = ao; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = C; d.ksize = 1; d.stride = 1;
          d.wt = s.to_out.w; d.Cout = C; d.mode = EPI_STD; d.bias = s.to_out.b;
          d.rowbias = ctx_vecs_ + static_cast<size_t>(MAX_BC) * s.ctx_vec_off + static_cast<size_t>(half) * Bc * C;
          d.rowbias_ld = C;
          set_res(d, t0);
          Act dst = t1;
          if (s16) dst.b16 = t1.b16 + static_cast<size_t>(half) * M * t1_ld;
          else dst.f32 = t1.f32 + static_cast<size_t>(half) * M * C;
          set_out(d, dst);
          if (merge) { d.ld_out = catld; d.res_ld = C; }
          if (half == 0) ln3_parts = want_ln_stats(d, M_t1, &ln3_stats);
          else if (ln3_parts) { d.ln_stats_out = ln3_stats; d.ln_stats_stride = static_cast<long long>(M_t1); }
          if (ln3_parts) d.ln_stats_out = ln3_stats + static_cast<size_t>(half) * M;
          add_gemm(tag + (half ? ".attn1.to_out+attn2[cond]" : ".attn1.to_out+attn2"), d);
        }
        if (!diverged) {
          // from here on the two halves differ: duplicate the shared block outputs that later ops read per sample
          // (the skip tensors pushed so far and this transformer's input, the residual of proj_out)
          std::vector<Act*> dup;
          for (Act& a0 : hs) dup.push_back(&a0);
          dup.push_back(&h);
          // (copies of tensors nobody reads before this transformer's last GEMM: a branch of the graph beside the GEMMs in between)
          bool first_dup = true;
          auto dup_lane = [&]() {
            if (!graph_lanes_) return;
            flag_last(first_dup ? (OP_SIDE | OP_FORK) : OP_SIDE);
            first_dup = false;
            dup_on_side = true;
          };
          for (Act* a0 : dup) {
            const size_t elems = static_cast<size_t>(Bc) * a0->H * a0->W * a0->C;
            if (a0->f32) {
              float* f = a0->f32;
              add_op(tag + ".dup_f32", 0, [=](cudaStream_t st) {
                PBE_CHECK_CUDA(cudaMemcpyAsync(f + elems, f, elems * sizeof(float), cudaMemcpyDeviceToDevice, st));
                return 0;
              });
              dup_lane();
            }
            if (a0->has_stats) {
              const size_t sel = static_cast<size_t>(Bc) * a0->H * a0->W / 32 * a0->C * 2;
              float* sp = a0->stats;
              add_op(tag + ".dup_stats", 0, [=](cudaStream_t st) {
                PBE_CHECK_CUDA(cudaMemcpyAsync(sp + sel, sp, sel * sizeof(float), cudaMemcpyDeviceToDevice, st));
                return 0;
              });
              dup_lane();
            }
            if (a0->has16) {
              bf16* bp = a0->b16;
              add_op(tag + ".dup_b16", 0, [=](cudaStream_t st) {
                PBE_CHECK_CUDA(cudaMemcpyAsync(bp + elems, bp, elems * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
                return 0;
              });
              dup_lane();
            }
          }
          diverged = true;
          Bc = BcFull;
          M = static_cast<size_t>(Bc) * N;
        }
        bf16* n3 = ln3_parts ? t1.b16 : static_cast<bf16*>(SA(M * C * sizeof(bf16)));
        if (!ln3_parts) {
          const float *gg = ln_ones_, *bb = ln_zeros_;
          const int Mi = static_cast<int>(M);
          const float* src = as_f32ptr(t1);
          const int in16 = s16 ? 1 : 0;
          add_op_meta(tag + ".ln3", 1, [=](cudaStream_t st) { return launch_layernorm(src, gg, bb, n3, Mi, C, 1e-5f, st, nullptr, 0, in16); },
                      "layernorm", 0.0, static_cast<double>(M) * C * (esz + 2.0));
        }
        if (merge && !ln3_parts) { err = -5; last_error = "ff / proj_out merge without LayerNorm-fold statistics at " + tag; return err; }
        bf16* gg = merge ? cat : static_cast<bf16*>(SA(M * 4 * C * sizeof(bf16)));
        {
          ConvGemmDesc d{};
          d.act = n3; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = C; d.ksize = 1; d.stride = 1;
          d.wt = s.ff1.w; d.Cout = 8 * C; d.mode = EPI_GEGLU; d.bias = s.ff1.b; d.out_bf16 = gg; d.ld_out = 4 * C;
          if (merge) { d.act_ld = catld; d.ld_out = catld; }
          if (ln3_parts) { d.ln_in = ln3_stats; d.ln_stride = static_cast<long long>(M); d.ln_parts = ln3_parts; d.ln_eps = 1e-5f; }
          add_gemm(tag + ".ff.geglu", d);
        }
        bf16* t2 = merge ? nullptr : static_cast<bf16*>(SA(M * C * sizeof(bf16)));
        if (!merge) {
          ConvGemmDesc d{};
          d.act = gg; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = 4 * C; d.ksize = 1; d.stride = 1;
          d.wt = s.ff2.w; d.Cout = C; d.mode = EPI_STD; d.bias = s.ff2.b; d.out_bf16 = t2;
          set_res(d, t1);
          add_gemm(tag + ".ff.out", d);
        }
        Act o = new_act(true, M * C, C, h.H, h.W);
        if (next_is_down) { o.b16 = static_cast<bf16*>(PA(M * C * sizeof(bf16))); o.has16 = true; }
        {
          ConvGemmDesc d{};
          d.act = merge ? cat : t2; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = merge ? catld : C; d.ksize = 1; d.stride = 1;
          d.wt = merge ? s.ffproj.w : s.proj_out.w; d.Cout = C; d.mode = EPI_STD; d.bias = merge ? s.ffproj.b : s.proj_out.b;
          set_res(d, h);
          set_out(d, o);
          want_stats(d, o, true);
          add_gemm(tag + (merge ? ".ff.out+proj_out" : ".proj_out"), d);
          if (dup_on_side) { flag_last(OP_JOIN); dup_on_side = false; }   // reads the duplicated h
        }
        h = o;
        break;
      }
      case Module::DOWN: {
        const ConvW& c = convs_[m.idx];
        if (!h.has16 || (h.H & 1) || (h.W & 1)) {
          err = -5;
          last_error = "downsample needs a 16-bit copy and even H, W at " + tag;
          return err;
        }
        const size_t M = static_cast<size_t>(Bc) * (h.H / 2) * (h.W / 2);
        Act o = new_act(true, M * c.cout, c.cout, h.H / 2, h.W / 2);
        ConvGemmDesc d{};
        d.act = h.b16; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = h.C; d.ksize = 3; d.stride = 2;
        d.wt = c.w; d.Cout = c.cout; d.mode = EPI_STD; d.bias = c.b;
        set_out(d, o);
        want_stats(d, o, true);
        add_gemm(tag + ".downsample", d);
        h = o;
        break;
      }
      case Module::UP: {
        const ConvW& c = convs_[m.idx];
        const size_t M = static_cast<size_t>(Bc) * (2 * h.H) * (2 * h.W);
        Act o = new_act(true, M * c.cout, c.cout, 2 * h.H, 2 * h.W);
        // Each phase is its own launch over the low-resolution pixel grid: worth it when one such launch still fills the GPU
        // (CFG batch 16: the 16->32 and 32->64 levels); small batches keep the literal form, whose single launch has 4x the tiles.
        const long up_tiles = ((static_cast<long>(Bc) * h.H * h.W + 127) / 128) * ((c.cout + 159) / 160);
        if (s16 && subpixel_up_ && (up_tiles >= 100 || subpixel_up_ >= 2)) {
          // four sub-pixel phase GEMMs on the low-resolution stream tensor itself (no upsampled copy, 4/9 of the FLOPs)
          ConvGemmDesc d{};
          d.act = h.b16; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = h.C; d.ksize = 2; d.stride = 1;
          d.Cout = c.cout; d.mode = EPI_STD; d.bias = c.b;
          set_out(d, o);
          d.up_phase = 1;
          if (gemm_can_fuse_stats(d)) {
            o.stats = static_cast<float*>(PA(M / 32 * c.cout * 2 * sizeof(float)));
            o.has_stats = true;
          }
          for (int ph = 0; ph < 4; ++ph) {
            d.up_phase = ph + 1;
            d.wt = c.wup + static_cast<size_t>(ph) * 4 * c.cout * h.C;
            d.stats_out = o.has_stats ? o.stats : nullptr;
            add_gemm(tag + ".upsample.conv[phase " + std::to_string(ph) + "]", d);
          }
          h = o;
          break;
        }
        bf16* up = static_cast<bf16*>(SA(M * h.C * sizeof(bf16)));
        {
          const float* src = h.f32;
          const bf16* src16 = h.b16;
          const int hh = h.H, ww = h.W, cc = h.C;
          add_op_meta(tag + ".upsample2x", 1,
                      [=](cudaStream_t st) {
                        return s16 ? launch_upsample2x_16(src16, up, Bc, hh, ww, cc, st) : launch_upsample2x_bf16(src, up, Bc, hh, ww, cc, st);
                      },
                      "upsample", 0.0, static_cast<double>(Bc) * hh * ww * cc * (esz + 8.0));
        }
        ConvGemmDesc d{};
        d.act = up; d.Nb = Bc; d.H = 2 * h.H; d.W = 2 * h.W; d.C = h.C; d.ksize = 3; d.stride = 1;
        d.wt = c.w; d.Cout = c.cout; d.mode = EPI_STD; d.bias = c.b;
        set_out(d, o);
        want_stats(d, o, true);
        add_gemm(tag + ".upsample.conv", d);
        h = o;
        break;
      }
      case Module::OUT: {
        const ConvW& c = convs_[m.idx];
        const size_t M = static_cast<size_t>(Bc) * h.H * h.W;
        bf16* a = static_cast<bf16*>(SA(M * h.C * sizeof(bf16)));
        GroupNormArgs g{};
        g.x0 = as_f32ptr(h); g.C0 = h.C; g.x1 = nullptr; g.C1 = 0; g.Nb = Bc; g.HW = h.H * h.W;
        g.in16 = s16 ? 1 : 0;
        g.gamma = out_norm_.g; g.beta = out_norm_.b; g.eps = 1e-5f; g.silu = 1; g.y = a; g.raw = nullptr;
        g.stats0 = h.has_stats ? h.stats : nullptr;
        add_gn(tag + ".out.norm", g);
        float* y = static_cast<float*>(SA(M * c.cout * sizeof(float)));       // eps leaves the network in fp32
        ConvGemmDesc d{};
        d.act = a; d.Nb = Bc; d.H = h.H; d.W = h.W; d.C = h.C; d.ksize = 3; d.stride = 1;
        d.wt = c.w; d.Cout = c.cout; d.mode = EPI_STD; d.bias = c.b; d.out_f32 = y; d.block_n = 32;
        add_gemm(tag + ".out.conv", d);
        float* dst = P.eps_stage;
        const int co = c.cout, hh = h.H, ww = h.W;
        add_op(tag + ".unpack_output", 1,
               [=](cudaStream_t st) { return launch_unpack_output(y, dst, Bc, co, hh, ww, co, st); });
        break;
      }
    }
    if (m.push_skip) hs.push_back(h);
    P.scratch.rewind(smark);
    if (err) return err;
  }
  if (!hs.empty()) {
    last_error = "skip stack not empty after build";
    return -5;
  }
  P.launches = launches;
  return 0;
}

bool Engine::pair_plan_possible() const {
  return modules_.size() > 2 && modules_[0].kind == Module::CONV_IN && modules_[1].kind == Module::RES &&
         modules_[2].kind == Module::ST && !modules_[1].pop_skip;
}

int Engine::prepare(int Bc, int H, int W, bool pair) {
  PBE_REQUIRE(finalized_, "weights not finalized");
  PBE_REQUIRE(fmt_f16_ == operand_f16(), "the operand format (pbe_set_operand_format / PBE_OPERANDS) changed after this engine was built");
  PBE_REQUIRE(Bc >= 1 && Bc <= MAX_BC, "batch out of range");
  int down = 1;
  for (int i = 1; i < cfg_.num_levels; ++i) down *= 2;
  PBE_REQUIRE(H % down == 0 && W % down == 0, "latent size must be divisible by 2^(levels-1)");
  auto key = std::make_tuple(Bc, H, W, pair ? 1 : 0);
  auto it = prepared_.find(key);
  if (it != prepared_.end()) {
    cur_ = it->second.get();
    cur_->last_use = ++use_clock_;
    return 0;
  }
  // a plan holds its arenas (> 1 GB at CFG batch 16) and a CUDA graph: keep the most recently used few
  while (prepared_.size() >= kMaxPlans) {
    auto victim = prepared_.begin();
    for (auto jt = prepared_.begin(); jt != prepared_.end(); ++jt)
      if (jt->second->last_use < victim->second->last_use) victim = jt;
    if (cur_ == victim->second.get()) cur_ = nullptr;
    prepared_.erase(victim);
  }
  auto P = std::make_unique<Prepared>();
  P->Bc = Bc; P->H = H; P->W = W; P->pair = pair;
  int rc = build(*P, true);
  if (rc) { set_error(last_error); return rc; }
  const size_t pbytes = P->persist.high() + 4096, sbytes = P->scratch.high() + 4096;
  void* p = nullptr;
  PBE_CHECK_CUDA(cudaMalloc(&p, pbytes));
  P->persist.base_ = static_cast<char*>(p);
  P->persist.cap_ = pbytes;
  PBE_CHECK_CUDA(cudaMalloc(&p, sbytes));
  P->scratch.base_ = static_cast<char*>(p);
  P->scratch.cap_ = sbytes;
  rc = build(*P, false);
  if (rc) { set_error(last_error); return rc; }
  cur_ = P.get();
  cur_->last_use = ++use_clock_;
  prepared_[key] = std::move(P);
  return 0;
}

int Engine::profile_forward(const float* x, const int64_t* t, float* eps, int Bc, int H, int W, cudaStream_t stream,
                            float* ms, int max_ops) {
  int rc = prepare(Bc, H, W);
  if (rc) return rc;
  Prepared& P = *cur_;
  PBE_REQUIRE(ctx_Bc_ == Bc, "set_context must be called with the same batch before forward");
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.x_stage, x, static_cast<size_t>(Bc) * cfg_.in_channels * H * W * sizeof(float),
                                 cudaMemcpyDeviceToDevice, stream));
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.t_stage, t, static_cast<size_t>(Bc) * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                                 stream));
  const size_t n = P.ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) PBE_CHECK_CUDA(cudaEventCreate(&e));
  PBE_CHECK_CUDA(cudaEventRecord(ev[0], stream));
  // PBE_GEMM_DEBUG: in-situ wait counters of CTA 0 of every conv_gemm op (serialises the pass; times are then not comparable)
  const bool gdbg = getenv("PBE_GEMM_DEBUG") != nullptr;
  for (size_t i = 0; i < n; ++i) {
    if (gdbg && P.op_family[i] == "conv_gemm") gemm_reset_debug_counters();
    rc = P.ops[i](stream);
    if (rc) return rc;
    PBE_CHECK_CUDA(cudaEventRecord(ev[i + 1], stream));
    if (gdbg && P.op_family[i] == "conv_gemm") {
      long long c[16];
      PBE_CHECK_CUDA(cudaStreamSynchronize(stream));
      gemm_read_debug_counters16(c);
      fprintf(stderr, "[gemm-dbg] %-34s mma %8lld cyc (wait tma %3.0f%% tmem %3.0f%%) %4lld k-it %5.0f cyc/it | prod wait-empty %8lld | epi %8lld cyc (wait acc %3.0f%% slot %3.0f%% bar %3.0f%%) %3lld chunks: ld+bias %5.0f pack %5.0f fence %5.0f cyc/chunk\n",
              P.op_names[i].c_str(), c[0], c[0] ? 100.0 * c[1] / c[0] : 0.0, c[0] ? 100.0 * c[2] / c[0] : 0.0, c[4],
              c[4] ? static_cast<double>(c[0]) / c[4] : 0.0, c[3], c[5], c[5] ? 100.0 * c[6] / c[5] : 0.0,
              c[5] ? 100.0 * (c[7] >> 32) / c[5] : 0.0, c[5] ? 100.0 * (c[7] & 0xffffffffll) / c[5] : 0.0, c[11],
              c[11] ? static_cast<double>(c[8]) / c[11] : 0.0, c[11] ? static_cast<double>(c[9]) / c[11] : 0.0,
              c[11] ? static_cast<double>(c[10]) / c[11] : 0.0);
    }
  }
  PBE_CHECK_CUDA(cudaMemcpyAsync(eps, P.eps_stage, static_cast<size_t>(Bc) * cfg_.out_channels * H * W * sizeof(float),
                                 cudaMemcpyDeviceToDevice, stream));
  PBE_CHECK_CUDA(cudaStreamSynchronize(stream));
  for (size_t i = 0; i < n && static_cast<int>(i) < max_ops; ++i) PBE_CHECK_CUDA(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  return static_cast<int>(n);
}

int run_op_list(const std::vector<std::function<int(cudaStream_t)>>& ops, const std::vector<std::string>& names,
                cudaStream_t stream, bool use_graph, cudaGraphExec_t* graph, cudaStream_t* cap_stream,
                const std::vector<uint8_t>* flags, cudaStream_t* side_stream) {
  int rc;
  const bool lanes = flags != nullptr && flags->size() == ops.size() && side_stream != nullptr;
  if (!use_graph) {
    for (size_t i = 0; i < ops.size(); ++i) {
      rc = ops[i](stream);
      if (rc) { set_error(std::string(get_error()) + " [" + names[i] + "]"); return rc; }
    }
    return 0;
  }
  if (*graph == nullptr) {
    // warm every kernel once eagerly (sets function attributes outside capture), then capture
    for (size_t i = 0; i < ops.size(); ++i) {
      rc = ops[i](stream);
      if (rc) { set_error(std::string(get_error()) + " [" + names[i] + "]"); return rc; }
    }
    PBE_CHECK_CUDA(cudaStreamSynchronize(stream));
    // capture on a private stream (the caller's may be the legacy default stream, which cannot capture);
    // the instantiated graph is then launched on the caller's stream.
    if (*cap_stream == nullptr) PBE_CHECK_CUDA(cudaStreamCreateWithFlags(cap_stream, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    // Side lane (OP_SIDE ops): a second capture stream, i.e. a branch of the graph.  OP_FORK: the side lane first waits for
    // everything issued on the main lane so far; OP_JOIN (a main-lane op): the main lane first waits for the side lane.  The
    // captured events become plain graph edges; the eager paths above simply run the list in order, which satisfies both.
    // Ops next to a fork / join are launched without programmatic dependent launch (full edges only on those nodes).
    std::vector<cudaEvent_t> events;
    auto new_event = [&](cudaEvent_t* e) -> int {
      PBE_CHECK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
      events.push_back(*e);
      return 0;
    };
    if (lanes && *side_stream == nullptr) PBE_CHECK_CUDA(cudaStreamCreateWithFlags(side_stream, cudaStreamNonBlocking));
    PBE_CHECK_CUDA(cudaStreamBeginCapture(*cap_stream, cudaStreamCaptureModeThreadLocal));
    bool side_open = false;   // the side lane holds work the main lane has not joined yet
    auto fail = [&](int code) {
      cudaStreamEndCapture(*cap_stream, &g);
      if (g) cudaGraphDestroy(g);
      for (cudaEvent_t e : events) cudaEventDestroy(e);
      pdl_suppress(false);
      return code;
    };
    for (size_t i = 0; i < ops.size(); ++i) {
      const uint8_t f = lanes ? (*flags)[i] : 0;
      cudaStream_t st = *cap_stream;
      bool edge = false;
      if (f & OP_SIDE) {
        st = *side_stream;
        if ((f & OP_FORK) || !side_open) {
          cudaEvent_t e;
          if (new_event(&e) || cudaEventRecord(e, *cap_stream) != cudaSuccess || cudaStreamWaitEvent(*side_stream, e, 0) != cudaSuccess) {
            set_error("graph fork failed at " + names[i]);
            return fail(-2);
          }
          side_open = true;
          edge = true;
        }
      } else if ((f & OP_JOIN) && side_open) {
        cudaEvent_t e;
        if (new_event(&e) || cudaEventRecord(e, *side_stream) != cudaSuccess || cudaStreamWaitEvent(*cap_stream, e, 0) != cudaSuccess) {
          set_error("graph join failed at " + names[i]);
          return fail(-2);
        }
        side_open = false;
        edge = true;
      }
      pdl_suppress(edge);
      rc = ops[i](st);
      pdl_suppress(false);
      if (rc) return fail(rc);
    }
    if (side_open) {   // every branch ends in the main lane before the capture does
      cudaEvent_t e;
      if (new_event(&e) || cudaEventRecord(e, *side_stream) != cudaSuccess || cudaStreamWaitEvent(*cap_stream, e, 0) != cudaSuccess) {
        set_error("graph join failed at the end of the op list");
        return fail(-2);
      }
    }
    PBE_CHECK_CUDA(cudaStreamEndCapture(*cap_stream, &g));
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    PBE_CHECK_CUDA(cudaGraphInstantiate(graph, g, 0));
    cudaGraphDestroy(g);
  }
  PBE_CHECK_CUDA(cudaGraphLaunch(*graph, stream));
  return 0;
}

int Engine::forward_pair(const float* x, const int64_t* t, float* eps, int B, int H, int W, cudaStream_t stream) {
  const int Bc = 2 * B;
  const bool pair = pair_plan_possible();
  int rc = prepare(Bc, H, W, pair);
  if (rc) return rc;
  Prepared& P = *cur_;
  PBE_REQUIRE(ctx_Bc_ == Bc, "set_context must be called with 2 * B rows (unconditional, then conditional) before forward_pair");
  const size_t xb = static_cast<size_t>(B) * cfg_.in_channels * H * W * sizeof(float);
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.x_stage, x, xb, cudaMemcpyDeviceToDevice, stream));
  if (!pair)   // no shared-prefix plan for this architecture: run the ordinary plan on the duplicated batch
    PBE_CHECK_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(P.x_stage) + xb, x, xb, cudaMemcpyDeviceToDevice, stream));
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.t_stage, t, static_cast<size_t>(B) * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.t_stage + B, t, static_cast<size_t>(B) * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
  pdl_set_scope(Bc <= 4 ? 1 : 0);
  rc = run_op_list(P.ops, P.op_names, stream, use_graph, &P.graph, &cap_stream_, graph_lanes_ ? &P.op_flags : nullptr, &side_stream_);
  pdl_set_scope(-1);
  if (rc) return rc;
  PBE_CHECK_CUDA(cudaMemcpyAsync(eps, P.eps_stage, static_cast<size_t>(Bc) * cfg_.out_channels * H * W * sizeof(float),
                                 cudaMemcpyDeviceToDevice, stream));
  return 0;
}

int Engine::forward(const float* x, const int64_t* t, float* eps, int Bc, int H, int W, cudaStream_t stream) {
  int rc = prepare(Bc, H, W);
  if (rc) return rc;
  Prepared& P = *cur_;
  PBE_REQUIRE(ctx_Bc_ == Bc, "set_context must be called with the same batch before forward");
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.x_stage, x, static_cast<size_t>(Bc) * cfg_.in_channels * H * W * sizeof(float),
                                 cudaMemcpyDeviceToDevice, stream));
  PBE_CHECK_CUDA(cudaMemcpyAsync(P.t_stage, t, static_cast<size_t>(Bc) * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                                 stream));
  // latency-bound small batches gain from programmatic dependent launch, throughput batches lose (tmap.cu: pdl_enabled)
  pdl_set_scope(Bc <= 4 ? 1 : 0);
  rc = run_op_list(P.ops, P.op_names, stream, use_graph, &P.graph, &cap_stream_, graph_lanes_ ? &P.op_flags : nullptr, &side_stream_);
  pdl_set_scope(-1);
  if (rc) return rc;
  PBE_CHECK_CUDA(cudaMemcpyAsync(eps, P.eps_stage, static_cast<size_t>(Bc) * cfg_.out_channels * H * W * sizeof(float),
                                 cudaMemcpyDeviceToDevice, stream));
  return 0;
}

}  // namespace pbe
