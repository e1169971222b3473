// Library-level plumbing of the C-ABI: version, last-error string, device info and
// TMA tensor-map creation (driver entry point fetched through the runtime).
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "../../include/cinema_b200.h"
#include "common.cuh"

#include <cstdlib>

static thread_local char g_err[1024] = "";

void cb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* cb_last_error(void) { return g_err; }
extern "C" int cb_version(void) { return CINEMA_B200_ABI_VERSION; }

int cb_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" int cb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  CB_CUDA(cudaGetDevice(&dev));
  if (sm_count) CB_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) CB_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) CB_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_tmap_nd_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int cb_make_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes) {
  return make_tmap_nd_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

int cb_make_tmap_nd_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes) {
  return make_tmap_nd_typed(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

static int make_tmap_nd_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  CB_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  CB_CHECK_ARG(rank >= 1 && rank <= 5, "tensor map rank %d out of range", rank);
  CB_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address must be 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      CB_CHECK_ARG((gstr[i - 1] & 15) == 0, "TMA stride %llu of dim %d is not a multiple of 16 bytes",
                   (unsigned long long)gstr[i - 1], i);
    }
    CB_CHECK_ARG(box[i] >= 1 && box[i] <= 256, "TMA box dim %u out of range", box[i]);
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cb_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u, sw %d)", (int)r,
                 rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0],
                 rank > 1 ? box[1] : 1, swizzle_bytes);
    return -2;
  }
  return 0;
}

int cb_make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                    uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {pitch_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return cb_make_tmap_nd(out, base, 2, dims, strides, box, swizzle_bytes);
}

static int g_pdl = -1;  // -1: read CB_PDL on first use

bool cb_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("CB_PDL");
    g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

extern "C" int cb_set_pdl(int enabled) {
  const int prev = cb_pdl_enabled() ? 1 : 0;
  g_pdl = enabled ? 1 : 0;
  return prev;
}
