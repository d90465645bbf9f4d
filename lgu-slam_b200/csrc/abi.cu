// abi.cu -- library identification + error plumbing of the C ABI (include/lgu_corr.h).
#include <cstdarg>
#include <cstdio>
#define LGU_STR2(x) #x
#define LGU_STR(x) LGU_STR2(x)
#include "common.cuh"

namespace lgu {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  return LGU_OK;
}
}  // namespace lgu

extern "C" {
int lgu_abi_version(void) { return 1; }
const char* lgu_build_info(void) {
  return "lgu_corr sm_100a, nvcc " LGU_STR(__CUDACC_VER_MAJOR__) "." LGU_STR(__CUDACC_VER_MINOR__) ", built " __DATE__;
}
const char* lgu_last_error_string(void) { return lgu::g_err; }
}
