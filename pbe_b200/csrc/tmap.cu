// Error string plumbing and TMA tensor-map construction.
#include "internal.h"

#include <stdlib.h>

#include <mutex>

namespace pbe {

static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128) {
  return make_tmap(out, base, /*is_f32=*/false, rank, dims, strides_bytes, box, swizzle128 ? 128 : 0);
}

int make_tmap(CUtensorMap* out, const void* base, bool is_f32, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = resolve_encode();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled could not be resolved from the CUDA driver");
    return -3;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstrides[i - 1] = strides_bytes[i - 1];
  }
  const CUtensorMapSwizzle swz = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstrides, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string m = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " rank " +
                    std::to_string(rank) + " dims[";
    for (int i = 0; i < rank; ++i) m += std::to_string(dims[i]) + (i + 1 < rank ? "," : "] box[");
    for (int i = 0; i < rank; ++i) m += std::to_string(box[i]) + (i + 1 < rank ? "," : "] strides[");
    for (int i = 0; i + 1 < rank; ++i) m += std::to_string(strides_bytes[i]) + (i + 2 < rank ? "," : "]");
    set_error(m);
    return -3;
  }
  return 0;
}

namespace {
int g_operand_f16 = -1;   // -1: not decided yet (PBE_OPERANDS)
}
int operand_f16() {
  if (g_operand_f16 < 0) {
    const char* e = getenv("PBE_OPERANDS");
    g_operand_f16 = (e != nullptr && (e[0] == 'b' || e[0] == 'B')) ? 0 : 1;
  }
  return g_operand_f16;
}
void set_operand_f16(int f16) { g_operand_f16 = f16 ? 1 : 0; }

namespace {
thread_local int g_pdl_scope = -1;   // -1: no preference from the engine that is building / capturing a launch plan
thread_local bool g_pdl_suppress = false;
}
void pdl_set_scope(int v) { g_pdl_scope = v; }
void pdl_suppress(bool on) { g_pdl_suppress = on; }

bool pdl_enabled() {
  // Measured on B200 inside the captured graph: PDL gives +1.7 % at CFG batch 2 (4.20 -> 4.13 ms per U-Net call) but
  // -1.5 % at CFG batch 16 (17.25 -> 17.60 ms).  So the U-Net engine asks for it (pdl_set_scope) only while it captures
  // small-batch plans; PBE_PDL=0 / 1 overrides everything.
  static const int env = [] { const char* e = getenv("PBE_PDL"); return e == nullptr ? -1 : (atoi(e) != 0 ? 1 : 0); }();
  if (g_pdl_suppress) return false;   // a graph node next to a fork / join of the capture lanes: full dependencies only
  if (env >= 0) return env == 1;
  return g_pdl_scope == 1;
}

}  // namespace pbe
