// C-ABI glue: error plumbing + device check.  Kernel entry points live next to their kernels.
#include "common.cuh"
#include "gemm_tf32.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace f2g {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const int on = [] {
    const char* v = getenv("F2G_PDL");
    return v ? atoi(v) : 1;
  }();
  return on != 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace f2g

extern "C" {

int f2g_abi_version(void) { return F2G_ABI_VERSION; }

const char* f2g_last_error(void) { return f2g::g_err; }

int f2g_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    f2g::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return (int)e;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    f2g::set_error("flow2gan_b200 kernels are built for sm_100a only; device is sm_%d%d", major,
                   minor);
    return f2g::F2G_EARCH;
  }
  return 0;
}

int f2g_gemm_tf32(const F2GGemm* problems, int n_problems, void* stream) {
  return f2g::gemm_tf32_group(problems, n_problems, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
