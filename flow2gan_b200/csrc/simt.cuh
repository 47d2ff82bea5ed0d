// Portable subset for the thread-independent SIMT kernels of the data path (datapath.cu).
//
// Under nvcc this is ordinary CUDA.  Under a host compiler with -DF2G_HOST_EMUL the same kernel
// bodies run as sequential loops over (block, thread): tests/ builds that variant with g++ so the
// index arithmetic and rounding of these kernels are checked on a machine without a GPU
// (tests/test_datapath_cpu.py).  The emulated build is test infrastructure only -- the product
// library never contains it and `_lib.py` never loads it.
//
// Rules for kernels written against this header: no __syncthreads, no shared memory, no warp
// intrinsics outside f2g::simt_block_{sum,max}; every thread's work may run in any order.
#pragma once

#ifndef F2G_HOST_EMUL
// ------------------------------------------------------------------------------------ CUDA
#include "common.cuh"

#include <string.h>

#define F2G_KERNEL __global__
#define F2G_SIMT_DEV __device__ __forceinline__
#define F2G_GRID_CONSTANT __grid_constant__
#define F2G_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)

namespace f2g {
// per-thread partials -> one atomic per warp (the kernels using these are HBM streams; a few
// thousand atomics per launch are noise)
F2G_SIMT_DEV void simt_block_sum(float v, float* dst) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) atomicAdd(dst, v);
}
// max of non-negative floats: their bit patterns order like unsigned integers
F2G_SIMT_DEV void simt_block_max_nonneg(float v, float* dst) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(dst), __float_as_uint(v));
}
F2G_SIMT_DEV float simt_fmul(float a, float b) { return __fmul_rn(a, b); }
F2G_SIMT_DEV float simt_fadd(float a, float b) { return __fadd_rn(a, b); }
F2G_SIMT_DEV double simt_dmul(double a, double b) { return __dmul_rn(a, b); }
F2G_SIMT_DEV double simt_dadd(double a, double b) { return __dadd_rn(a, b); }
F2G_SIMT_DEV int simt_rint(float v) { return __float2int_rn(v); }
inline int simt_memset_async(void* p, int v, size_t n, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(p, v, n, s);
  if (e != cudaSuccess) {
    set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
}  // namespace f2g

#else
// ------------------------------------------------------------------------------- host emulation
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define F2G_KERNEL static
#define F2G_SIMT_DEV static inline
#define F2G_GRID_CONSTANT
#define __restrict__ __restrict

// -DF2G_EMUL_REVERSE runs blocks and threads in descending order: results must not depend on it
#ifdef F2G_EMUL_REVERSE
#define F2G_EMUL_ORDER(i, n) ((n) - 1u - (i))
#else
#define F2G_EMUL_ORDER(i, n) (i)
#endif

typedef void* cudaStream_t;
struct f2g_dim3 {
  unsigned x, y, z;
};
static f2g_dim3 threadIdx, blockIdx, blockDim, gridDim;

#define F2G_LAUNCH(kernel, grid, block, stream, ...)                            \
  do {                                                                          \
    (void)(stream);                                                             \
    gridDim = {(unsigned)(grid), 1, 1};                                         \
    blockDim = {(unsigned)(block), 1, 1};                                       \
    for (unsigned b_ = 0; b_ < gridDim.x; ++b_)                                 \
      for (unsigned t_ = 0; t_ < blockDim.x; ++t_) {                            \
        blockIdx = {F2G_EMUL_ORDER(b_, gridDim.x), 0, 0};                       \
        threadIdx = {F2G_EMUL_ORDER(t_, blockDim.x), 0, 0};                     \
        kernel(__VA_ARGS__);                                                    \
      }                                                                         \
  } while (0)

namespace f2g {
enum { F2G_OK = 0, F2G_EINVAL = -1, F2G_EDRIVER = -2, F2G_EARCH = -3 };
inline char g_emul_err[512];       // one buffer for every translation unit of the emulated library
static inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_emul_err, sizeof(g_emul_err), fmt, ap);
  va_end(ap);
}
static inline int check_launch(const char*) { return 0; }
// the compiler flags of the emulated build forbid contraction (-ffp-contract=off)
F2G_SIMT_DEV void simt_block_sum(float v, float* dst) { *dst += v; }
F2G_SIMT_DEV void simt_block_max_nonneg(float v, float* dst) {
  if (v > *dst) *dst = v;
}
F2G_SIMT_DEV float simt_fmul(float a, float b) { return a * b; }
F2G_SIMT_DEV float simt_fadd(float a, float b) { return a + b; }
F2G_SIMT_DEV double simt_dmul(double a, double b) { return a * b; }
F2G_SIMT_DEV double simt_dadd(double a, double b) { return a + b; }
F2G_SIMT_DEV int simt_rint(float v) { return (int)lrintf(v); }   // default mode: nearest-even
static inline int simt_memset_async(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return 0;
}
}  // namespace f2g
static inline long long min(long long a, long long b) { return a < b ? a : b; }
extern "C" __attribute__((weak)) const char* f2g_emul_last_error(void) { return f2g::g_emul_err; }
#endif

// global thread id / stride of a 1-D launch (64-bit: waveforms of hours still index correctly)
#define F2G_GTID ((long long)blockIdx.x * blockDim.x + threadIdx.x)
#define F2G_GSTRIDE ((long long)gridDim.x * blockDim.x)
