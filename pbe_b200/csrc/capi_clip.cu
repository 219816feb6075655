// C-ABI of the conditioning front-end (include/pbe_b200.h).  Nothing throws across the boundary.
#include "clip.h"

#include <new>

using namespace pbe;

struct pbe_clip {
  ClipEncoder* e;
};

extern "C" {

int pbe_clip_create(const pbe_clip_config* cfg, pbe_clip_handle* out) {
  if (cfg == nullptr || out == nullptr) { set_error("pbe_clip_create: null argument"); return -1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("pbe_clip_create: no CUDA device (this library has no CPU fallback)");
    return -2;
  }
  try {
    pbe_clip* h = new pbe_clip;
    h->e = new ClipEncoder(*cfg);
    *out = h;
  } catch (const std::exception& ex) {
    set_error(std::string("pbe_clip_create: ") + ex.what());
    return -1;
  }
  return 0;
}

void pbe_clip_destroy(pbe_clip_handle h) {
  if (h == nullptr) return;
  delete h->e;
  delete h;
}

#define PBE_CLIP_GUARD(stmt)                                     \
  if (h == nullptr) { set_error("null handle"); return -1; }     \
  try { return (stmt); }                                         \
  catch (const std::exception& ex) { set_error(ex.what()); return -1; }

int pbe_clip_load_weight(pbe_clip_handle h, const char* name, const float* host_data, const int64_t* shape, int rank) {
  PBE_CLIP_GUARD(h->e->load_weight(name, host_data, shape, rank));
}
int pbe_clip_finalize_weights(pbe_clip_handle h) { PBE_CLIP_GUARD(h->e->finalize()); }
int pbe_clip_encode(pbe_clip_handle h, const float* image, float* z, int B, void* stream) {
  PBE_CLIP_GUARD(h->e->encode(image, z, B, static_cast<cudaStream_t>(stream)));
}
int pbe_clip_launches_per_encode(pbe_clip_handle h) { return h == nullptr ? 0 : h->e->launches(); }

}  // extern "C"
