// Error string, version, device queries for the C ABI (include/gossipnet_b200.h).
#include <stdarg.h>

#include "gn_common.cuh"

namespace gn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  return n;
}

static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }

}  // namespace gn

extern "C" {

int gn_set_pdl(int enable) {
  const int old = gn::g_pdl;
  gn::g_pdl = enable != 0;
  return old;
}

const char* gn_last_error(void) { return gn::g_err; }
int gn_abi_version(void) { return 1; }
int gn_sm_count(void) { return gn::sm_count(); }

}  // extern "C"

// ---- tensor-map encoder (gn_tma.cuh) ------------------------------------------------
#include "gn_tma.cuh"

namespace gn {

tmap_encode_fn tmap_encoder() {
  // process constant (the driver's entry point), resolved once; C++11 makes this thread safe
  static const tmap_encode_fn fn = []() -> tmap_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<tmap_encode_fn>(p);
  }();
  return fn;
}

static int encode_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, const void* base,
                          uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_rows,
                          uint32_t box_cols) {
  const tmap_encode_fn enc = tmap_encoder();
  if (enc == nullptr) return -1;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, dtype, 2, const_cast<void*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols) {
  return encode_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rows, cols, pitch_bytes,
                        box_rows, box_cols);
}

int encode_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols) {
  return encode_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rows, cols, pitch_bytes,
                        box_rows, box_cols);
}

}  // namespace gn
