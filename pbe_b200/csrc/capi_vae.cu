// C-ABI of the VAE decoder (include/pbe_b200.h).  Nothing throws across the boundary.
#include "vae.h"

#include <new>

using namespace pbe;

struct pbe_vae {
  VaeModel* d;
};

extern "C" {

int pbe_vae_create(const pbe_vae_config* cfg, pbe_vae_handle* out) {
  if (cfg == nullptr || out == nullptr) { set_error("pbe_vae_create: null argument"); return -1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("pbe_vae_create: no CUDA device (this library has no CPU fallback)");
    return -2;
  }
  try {
    pbe_vae* h = new pbe_vae;
    h->d = new VaeModel(*cfg);
    *out = h;
  } catch (const std::exception& ex) {
    set_error(std::string("pbe_vae_create: ") + ex.what());
    return -1;
  }
  return 0;
}

void pbe_vae_destroy(pbe_vae_handle h) {
  if (h == nullptr) return;
  delete h->d;
  delete h;
}

#define PBE_VAE_GUARD(stmt)                                      \
  if (h == nullptr) { set_error("null handle"); return -1; }     \
  try { return (stmt); }                                         \
  catch (const std::exception& ex) { set_error(ex.what()); return -1; }

int pbe_vae_load_weight(pbe_vae_handle h, const char* name, const float* host_data, const int64_t* shape, int rank) {
  PBE_VAE_GUARD(h->d->load_weight(name, host_data, shape, rank));
}
int pbe_vae_finalize_weights(pbe_vae_handle h) { PBE_VAE_GUARD(h->d->finalize()); }
int pbe_vae_decode(pbe_vae_handle h, const float* z, float* out, int B, int H, int W, void* stream) {
  PBE_VAE_GUARD(h->d->decode(z, out, B, H, W, static_cast<cudaStream_t>(stream)));
}
int pbe_vae_encode(pbe_vae_handle h, const float* x, float* moments, int B, int H, int W, void* stream) {
  PBE_VAE_GUARD(h->d->encode(x, moments, B, H, W, static_cast<cudaStream_t>(stream)));
}
int pbe_vae_profile(pbe_vae_handle h, int encode, const float* in, float* out, int B, int H, int W, void* stream,
                    float* ms_out, int max_ops) {
  PBE_VAE_GUARD(h->d->profile(encode, in, out, B, H, W, static_cast<cudaStream_t>(stream), ms_out, max_ops));
}
int pbe_vae_op_info(pbe_vae_handle h, int i, const char** name, const char** family, double* flops) {
  if (h == nullptr || h->d->current() == nullptr) { set_error("no prepared shape"); return -1; }
  const VaePrepared* P = h->d->current();
  if (i < 0 || i >= static_cast<int>(P->ops.size())) { set_error("op index out of range"); return -1; }
  if (name) *name = P->op_names[i].c_str();
  if (family) *family = P->op_family[i].c_str();
  if (flops) *flops = P->op_flops[i];
  return 0;
}
int pbe_postprocess_u8(const float* img, uint8_t* out, int B, int C, int H, int W, void* stream) {
  if (img == nullptr || out == nullptr) { set_error("pbe_postprocess_u8: null argument"); return -1; }
  return launch_postprocess_u8(img, out, B, C, H, W, static_cast<cudaStream_t>(stream));
}
int pbe_normalize_u8(const uint8_t* img_u8, float* out, int B, int H, int W, const float* mean3, const float* std3, void* stream) {
  if (img_u8 == nullptr || out == nullptr || mean3 == nullptr || std3 == nullptr) { set_error("pbe_normalize_u8: null argument"); return -1; }
  if (B <= 0 || H <= 0 || W <= 0) { set_error("pbe_normalize_u8: empty geometry"); return -1; }
  for (int c = 0; c < 3; ++c)
    if (std3[c] == 0.0f) { set_error("pbe_normalize_u8: std evaluated to zero, leading to division by zero"); return -1; }
  return launch_normalize_u8(img_u8, out, B, H, W, mean3, std3, static_cast<cudaStream_t>(stream));
}
int pbe_prepare_inpaint_u8(const uint8_t* img_u8, const uint8_t* mask_u8, int B, int H, int W, int binarize, float* image_out,
                           float* mask_out, float* inpaint_out, void* stream) {
  if (img_u8 == nullptr || mask_u8 == nullptr || inpaint_out == nullptr) { set_error("pbe_prepare_inpaint_u8: null argument"); return -1; }
  if (B <= 0 || H <= 0 || W <= 0) { set_error("pbe_prepare_inpaint_u8: empty geometry"); return -1; }
  return launch_prepare_inpaint_u8(img_u8, mask_u8, image_out, mask_out, inpaint_out, B, H, W, binarize,
                                   static_cast<cudaStream_t>(stream));
}
int pbe_resize_bilinear(const float* in, float* out, int NC, int H, int W, int h, int w, int antialias, void* stream) {
  if (in == nullptr || out == nullptr) { set_error("pbe_resize_bilinear: null argument"); return -1; }
  return launch_resize_bilinear(in, out, NC, H, W, h, w, antialias, static_cast<cudaStream_t>(stream));
}
int pbe_vae_launches_per_decode(pbe_vae_handle h) {
  if (h == nullptr || h->d->current() == nullptr) return 0;
  return h->d->current()->launches;
}

}  // extern "C"
