// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aniportrait_b200.h"

namespace ap {

// printf-style; stores the message in a thread-local buffer read by ap_last_error(). Returns `code`.
int fail(int code, const char* fmt, ...);

// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency, so the
// library also loads on a CPU-only box for symbol checks).
// dims/box are innermost-first; strides_bytes has rank-1 entries (stride of dim 1..rank-1).
int encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, bool swizzle128, int elem_bytes = 2, int swizzle_bytes = 0);

int num_sms();

#define AP_CHECK_CUDA(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return ap::fail(AP_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define AP_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return ap::fail(AP_ERR_INVALID, __VA_ARGS__); \
  } while (0)

}  // namespace ap
