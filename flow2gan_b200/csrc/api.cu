// C-ABI glue: error plumbing + device check.  Kernel entry points live next to their kernels.
#include "common.cuh"
#include <string.h>
#include "gemm_tf32.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace f2g {

static thread_local char g_err[1024] = "";

// Watchdog report of the chained GEMM launches (gemm_pair.cu): four ints in mapped pinned host memory
// {flag, problem, row tile, counter value seen}.  A consumer tile whose producer counter does not
// arrive within the spin bound writes its coordinates here and traps; host memory stays readable
// after the context has been poisoned by the trap, so the next failing call can say what happened.
static int* g_watch_host = nullptr;
static int* g_watch_dev = nullptr;

int* chain_watchdog_dev() {
  if (!g_watch_dev) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    memset(h, 0, 4 * sizeof(int));
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaFreeHost(h);
      return nullptr;
    }
    g_watch_host = static_cast<int*>(h);
    g_watch_dev = static_cast<int*>(d);
  }
  return g_watch_dev;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (g_watch_host && ((volatile int*)g_watch_host)[0]) {     // every later error is a consequence of the trap
    const size_t n = strlen(g_err);
    snprintf(g_err + n, sizeof(g_err) - n,
             " -- a chained GEMM launch timed out waiting for its producer tiles (problem %d, row tile %d, counter %d) "
             "and trapped: chained launches of one device must not run concurrently (include/flow2gan_b200.h)",
             g_watch_host[1], g_watch_host[2], g_watch_host[3]);
  }
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

int f2g_chain_watchdog(int out[4]) {
  for (int i = 0; i < 4; ++i) out[i] = f2g::g_watch_host ? ((volatile int*)f2g::g_watch_host)[i] : 0;
  return out[0];
}

int f2g_gemm_tf32(const F2GGemm* problems, int n_problems, void* stream) {
  return f2g::gemm_tf32_group(problems, n_problems, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
