// Host-side helpers shared by every entry point of libb200diff.so: error reporting,
// the cuTensorMapEncodeTiled entry point (resolved through the runtime so the library does
// not link libcuda), launch checking.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200_diffusion.h"

namespace b200 {

int set_error(int code, const char* fmt, ...);  // formats into the thread-local buffer, returns code
int num_sms();

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();  // nullptr (and last_error set) if unavailable

// 16-bit tensor map, SWIZZLE_128B, zero OOB fill.  dims/strides innermost-first; strides in BYTES for
// dims 1..rank-1.  Returns 0 or an error code.
int make_tensor_map_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const char* what, int swizzle_bytes = 128);

#define B200_CHECK_ARG(cond, ...) \
  do {                            \
    if (!(cond)) return b200::set_error(B200_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define B200_CHECK_CUDA(expr)                                                                      \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return b200::set_error(B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                                  \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

bool pdl_enabled();  // lib.cu: programmatic dependent launch on unless B200_NO_PDL=1

// <<<>>> replacement that tags the launch for programmatic stream serialization (see pdl_wait / pdl_trigger)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int rup(int v, int m) { return (v + m - 1) / m * m; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace b200
