// Library-level entry points: version, error string, device init, tensor-map encoding.
#include <stdlib.h>
#include <string.h>

#include "host_common.h"

namespace b200 {

static thread_local char g_err[512] = "";
static int g_num_sms = 0;
static EncodeTiledFn g_encode = nullptr;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static const bool on = !(getenv("B200_NO_PDL") && atoi(getenv("B200_NO_PDL")) != 0);
  return on;
}

int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

EncodeTiledFn encode_tiled_fn() {
  if (g_encode) return g_encode;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return g_encode;
}

int make_tensor_map_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const char* what, int swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return B200_ERR_CUDA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                      : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(B200_ERR_INVALID,
                     "cuTensorMapEncodeTiled(%s) failed with CUresult %d (base %p rank %d dims %llu,%llu,%llu,%llu "
                     "strides %llu,%llu,%llu box %u,%u,%u,%u)",
                     what, static_cast<int>(r), base, rank, (unsigned long long)gdim[0],
                     (unsigned long long)(rank > 1 ? gdim[1] : 0), (unsigned long long)(rank > 2 ? gdim[2] : 0),
                     (unsigned long long)(rank > 3 ? gdim[3] : 0), (unsigned long long)(rank > 1 ? gstr[0] : 0),
                     (unsigned long long)(rank > 2 ? gstr[1] : 0), (unsigned long long)(rank > 3 ? gstr[2] : 0), bdim[0],
                     rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0);
  }
  return 0;
}

int init_conv_gemm();  // conv_gemm.cu
int init_attention();  // attention.cu
int init_text_attention();  // text_attention.cu

}  // namespace b200

extern "C" {

int b200_version(void) { return 100; }

const char* b200_last_error(void) { return b200::g_err; }

int b200_num_sms(void) { return b200::num_sms(); }

static int init_device(int device);

int b200_init(int device) {
  // the process's current device is the caller's business: validate / configure `device`, then put back what was current
  int prev = -1;
  const bool have_prev = cudaGetDevice(&prev) == cudaSuccess;
  const int rc = init_device(device);
  if (have_prev && prev != device) cudaSetDevice(prev);
  return rc;
}

static int init_device(int device) {
  B200_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    return b200::set_error(B200_ERR_UNSUPPORTED, "device %d is sm_%d%d; libb200diff needs sm_100a (B200)", device,
                           prop.major, prop.minor);
  }
  b200::g_num_sms = prop.multiProcessorCount;
  if (!b200::encode_tiled_fn()) return B200_ERR_CUDA;
  int r = b200::init_conv_gemm();
  if (r) return r;
  r = b200::init_attention();
  if (r) return r;
  r = b200::init_text_attention();
  if (r) return r;
  return 0;
}

}  // extern "C"
