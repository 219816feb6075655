// C-ABI of the engine (include/pbe_b200.h).  Nothing throws across the boundary.
#include "engine.h"

#include <new>

using namespace pbe;

struct pbe_engine {
  Engine* e;
};

extern "C" {

int pbe_create(const pbe_config* cfg, pbe_handle* out) {
  if (cfg == nullptr || out == nullptr) { set_error("pbe_create: null argument"); return -1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("pbe_create: no CUDA device (this library has no CPU fallback)");
    return -2;
  }
  try {
    pbe_engine* h = new pbe_engine;
    h->e = new Engine(*cfg);
    *out = h;
  } catch (const std::exception& ex) {
    set_error(std::string("pbe_create: ") + ex.what());
    return -1;
  }
  return 0;
}

void pbe_destroy(pbe_handle h) {
  if (h == nullptr) return;
  delete h->e;
  delete h;
}

#define PBE_GUARD(stmt)                                          \
  if (h == nullptr) { set_error("null handle"); return -1; }     \
  try { return (stmt); }                                         \
  catch (const std::exception& ex) { set_error(ex.what()); return -1; }

int pbe_load_weight(pbe_handle h, const char* name, const float* host_data, const int64_t* shape, int rank) {
  PBE_GUARD(h->e->load_weight(name, host_data, shape, rank));
}
int pbe_finalize_weights(pbe_handle h) { PBE_GUARD(h->e->finalize()); }
int pbe_set_context(pbe_handle h, const float* ctx, int Bc, void* stream) {
  PBE_GUARD(h->e->set_context(ctx, Bc, static_cast<cudaStream_t>(stream)));
}
int pbe_unet_forward(pbe_handle h, const float* x, const int64_t* t, float* eps, int Bc, int H, int W, void* stream) {
  PBE_GUARD(h->e->forward(x, t, eps, Bc, H, W, static_cast<cudaStream_t>(stream)));
}
int pbe_unet_forward_cfg_pair(pbe_handle h, const float* x, const int64_t* t, float* eps, int B, int H, int W, void* stream) {
  PBE_GUARD(h->e->forward_pair(x, t, eps, B, H, W, static_cast<cudaStream_t>(stream)));
}
int pbe_set_use_graph(pbe_handle h, int enable) {
  if (h == nullptr) { set_error("null handle"); return -1; }
  h->e->use_graph = enable != 0;
  return 0;
}
int pbe_profile_forward(pbe_handle h, const float* x, const int64_t* t, float* eps, int Bc, int H, int W, void* stream,
                        float* ms_out, int max_ops) {
  PBE_GUARD(h->e->profile_forward(x, t, eps, Bc, H, W, static_cast<cudaStream_t>(stream), ms_out, max_ops));
}
int pbe_op_info(pbe_handle h, int i, const char** name, const char** family, double* flops, double* bytes) {
  if (h == nullptr || h->e->current() == nullptr) { set_error("no prepared shape"); return -1; }
  const Prepared* P = h->e->current();
  if (i < 0 || i >= static_cast<int>(P->ops.size())) { set_error("op index out of range"); return -1; }
  if (name) *name = P->op_names[i].c_str();
  if (family) *family = P->op_family[i].c_str();
  if (flops) *flops = P->op_flops[i];
  if (bytes) *bytes = P->op_bytes[i];
  return 0;
}
int pbe_launches_per_forward(pbe_handle h) {
  if (h == nullptr) return 0;
  return h->e->launches_per_forward();
}

}  // extern "C"
