// C-ABI plumbing: version, error string, device binding, TMA descriptor encoding.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ap_host.h"

namespace ap {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap_;
  va_start(ap_, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap_);
  va_end(ap_);
  return code;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static int g_num_sms = 0;

static int resolve_driver() {
  if (g_encode) return AP_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return fail(AP_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return AP_OK;
}

int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }

bool pdl_enabled() {
  // measured on one box (profiles/r02_summary.md): UNet3D call 48.16 ms with, 47.57 ms without; 12.32 vs 12.34 frames/s ->
  // no gain (the kernel-to-kernel gap inside a replayed graph is not launch latency), so it is opt-in
  static const bool on = getenv("AP_PDL") && atoi(getenv("AP_PDL")) != 0;
  return on;
}

int encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, bool swizzle128, int elem_bytes, int swizzle_bytes) {
  int rc = resolve_driver();
  if (rc) return rc;
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
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(AP_ERR_INVALID, "TMA base %p not 16B aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (gstrides[i] % 16 != 0) return fail(AP_ERR_INVALID, "TMA stride %d = %llu not a multiple of 16 B", i,
                                           (unsigned long long)gstrides[i]);
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                           : (elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                              : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = g_encode(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                            : (swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(AP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]",
                (int)r, rank, (unsigned long long)gdims[0], (unsigned long long)(rank > 1 ? gdims[1] : 0),
                (unsigned long long)(rank > 2 ? gdims[2] : 0), (unsigned long long)(rank > 3 ? gdims[3] : 0),
                (unsigned long long)(rank > 4 ? gdims[4] : 0), gbox[0], rank > 1 ? gbox[1] : 0, rank > 2 ? gbox[2] : 0,
                rank > 3 ? gbox[3] : 0, rank > 4 ? gbox[4] : 0);
  }
  return AP_OK;
}

}  // namespace ap

extern "C" int ap_version(void) { return AP_VERSION; }

extern "C" const char* ap_last_error(void) { return ap::g_err; }

extern "C" int ap_init(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return ap::fail(AP_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return ap::fail(AP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return ap::fail(AP_ERR_DEVICE, "device %d is sm_%d%d; this library only runs on sm_100 (B200)", device, prop.major,
                    prop.minor);
  ap::g_num_sms = prop.multiProcessorCount;
  return ap::resolve_driver();
}
