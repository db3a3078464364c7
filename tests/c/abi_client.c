/* A host with nothing but a C compiler and dlopen: loads libb200diff.so, resolves the entry points declared in
 * include/b200_diffusion.h and exercises the parts of the contract that need no GPU (version, host helpers, argument
 * validation with error strings).  Built and run by tests/test_abi.py; proves the boundary is plain C. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200_diffusion.h"

typedef int (*version_fn)(void);
typedef const char* (*last_error_fn)(void);
typedef int64_t (*packed_k_fn)(int32_t, int32_t, int32_t);
typedef int32_t (*pick_tile_fn)(int64_t, int32_t, int32_t);
typedef int (*conv_gemm_fn)(const b200_conv_gemm_args*, void*);

#define CHECK(cond, msg)                        \
  do {                                          \
    if (!(cond)) {                              \
      fprintf(stderr, "FAIL: %s\n", msg);       \
      return 1;                                 \
    }                                           \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "dlopen: %s\n", dlerror());
    return 3;
  }
  version_fn version = (version_fn)dlsym(h, "b200_version");
  last_error_fn last_error = (last_error_fn)dlsym(h, "b200_last_error");
  packed_k_fn packed_k = (packed_k_fn)dlsym(h, "b200_conv_gemm_packed_k");
  pick_tile_fn pick_tile = (pick_tile_fn)dlsym(h, "b200_conv_gemm_pick_tile_n");
  conv_gemm_fn conv_gemm = (conv_gemm_fn)dlsym(h, "b200_conv_gemm");
  CHECK(version && last_error && packed_k && pick_tile && conv_gemm, "missing symbol");
  CHECK(version() >= 100, "version");
  CHECK(packed_k(3, 320, 0) == 2880, "packed_k conv3x3 320");
  CHECK(packed_k(1, 1280, 640) == 1920, "packed_k two sources");
  CHECK(pick_tile(2048, 10240, 1) == 256, "geglu tile");
  /* argument validation happens before any CUDA call: a zeroed struct must be refused with a message */
  b200_conv_gemm_args a;
  memset(&a, 0, sizeof(a));
  CHECK(conv_gemm(&a, NULL) == B200_ERR_INVALID, "null operands accepted");
  CHECK(strstr(last_error(), "null") != NULL, "error string");
  CHECK(conv_gemm(NULL, NULL) == B200_ERR_INVALID, "null args accepted");
  static char buf[64];
  a.x[0] = buf;
  a.w = buf;
  a.y = buf;
  a.ksize = 5;
  CHECK(conv_gemm(&a, NULL) == B200_ERR_INVALID && strstr(last_error(), "ksize") != NULL, "ksize 5 accepted");
  printf("abi client ok: version %d, sizeof(b200_conv_gemm_args) = %zu\n", version(), sizeof(a));
  return 0;
}
