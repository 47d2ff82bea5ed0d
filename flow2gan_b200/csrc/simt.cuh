// Portable layer for the SIMT kernels of this library (everything but the tcgen05 / TMA GEMMs).
//
// Under nvcc this is ordinary CUDA (it includes common.cuh and the launch macros expand to <<<>>>).
// Under a host compiler with -DF2G_HOST_EMUL the same kernel bodies run on the host --
// thread-independent kernels (F2G_LAUNCH) as sequential loops over (block, thread), cooperative ones
// (F2G_LAUNCH_COOP / launch_pdl: shared memory, barriers, shuffles, atomics) with one host thread per
// warp whose lanes are user-level contexts -- so that tests/ can check index arithmetic, rounding and
// the Python host layer on a machine without a GPU (tests/_emul.py, tests/test_*_cpu.py).  The
// emulated build is test infrastructure only: the product library never contains it and `_lib.py`
// never loads it.
//
// F2G_LAUNCH kernels must not use __syncthreads, shared memory or warp intrinsics other than
// f2g::simt_block_{sum,max}: their threads may run in any order.
#pragma once

#ifndef F2G_HOST_EMUL
// ------------------------------------------------------------------------------------ CUDA
#include "common.cuh"

#include <string.h>

#define F2G_KERNEL __global__
#define F2G_SIMT_DEV __device__ __forceinline__
#define F2G_GRID_CONSTANT __grid_constant__
#define F2G_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define F2G_LAUNCH_COOP(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define F2G_LAUNCH_COOP_SMEM(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// dynamic shared memory of the enclosing kernel, viewed as T[]
#define F2G_DYN_SMEM(T, name) extern __shared__ T name[]

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
F2G_SIMT_DEV float4 unpack_half4(uint2 u) {
  const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only): costs no registers while in flight
F2G_SIMT_DEV void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
F2G_SIMT_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
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
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <ucontext.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#define F2G_KERNEL static
#define F2G_SIMT_DEV static inline
#define F2G_DEVINL static inline
#define F2G_GRID_CONSTANT
#define __restrict__ __restrict
#define __global__ static
#define __shared__ static      /* blocks run one after the other: one copy per kernel is "per block" */
#define __launch_bounds__(...)
#define __grid_constant__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline

// -DF2G_EMUL_REVERSE runs blocks (and, in the sequential mode, threads) in descending order: results
// must not depend on it
#ifdef F2G_EMUL_REVERSE
#define F2G_EMUL_ORDER(i, n) ((n) - 1u - (i))
#else
#define F2G_EMUL_ORDER(i, n) (i)
#endif

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }
struct f2g_dim3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 {
  int x, y;
};
struct alignas(8) uint2 {
  unsigned x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
struct alignas(8) float2 {
  float x, y;
};
struct __half {
  unsigned short bits;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline thread_local f2g_dim3 threadIdx, blockIdx;
inline f2g_dim3 blockDim, gridDim;

// Cooperative kernels (shared memory, __syncthreads, warp shuffles, atomics) run block after block
// with one host thread per warp (see below); thread-independent kernels run as plain loops
// (F2G_LAUNCH).  Tests only, small problems.
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
static inline int bringup_int(const char*, int dflt) { return dflt; }   // no timing knobs on the host
alignas(128) inline unsigned char g_dyn_smem[228 * 1024];     // "dynamic shared memory" of the running block
inline std::mutex g_atomic_mu;

// Cooperative execution model: one host thread per WARP, its 32 lanes are user-level contexts
// (ucontext) that the warp thread runs round-robin and that yield at every warp- or block-level
// synchronisation point (shuffles, __syncwarp, __syncthreads).  A warp-level point completes when
// every live lane of the warp has reached it; a block-level point additionally waits for the other
// warp threads.  Exited lanes / warps count as arrived, like on the GPU.
enum { LANE_READY = 0, LANE_WAIT_WARP = 1, LANE_WAIT_BLOCK = 2, LANE_DONE = 3 };
constexpr size_t EMUL_STACK = 96 * 1024;
struct EmulLane {
  ucontext_t ctx;
  int state;
};
struct EmulWarp {
  ucontext_t sched;
  EmulLane lane[32];
  unsigned nlanes, base_tid, cur, phase;
  float shfl[2][32];
  void (*trampoline)(void*);
  void* body;
};
inline thread_local EmulWarp* g_warp = nullptr;

struct EmulBlockBarrier {       // barrier over the warps that have not finished the current block
  std::mutex m;
  std::condition_variable cv;
  unsigned expected = 0, arrived = 0, gen = 0;
  void reset(unsigned n) { expected = n; arrived = 0; }
  void arrive_and_wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived == expected) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
  void drop() {
    std::unique_lock<std::mutex> lk(m);
    --expected;
    if (expected > 0 && arrived == expected) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    }
  }
};
inline EmulBlockBarrier g_block_sync;
inline pthread_barrier_t g_block_end;

inline void emul_yield(int state) {
  EmulWarp* w = g_warp;
  w->lane[w->cur].state = state;
  swapcontext(&w->lane[w->cur].ctx, &w->sched);
}
inline void emul_lane_main() {
  EmulWarp* w = g_warp;
  w->trampoline(w->body);
  w->lane[w->cur].state = LANE_DONE;       // uc_link returns to the warp scheduler
}

// runs one block's worth of this warp's lanes to completion
inline void emul_run_warp_block(EmulWarp* w, char* stacks) {
  for (unsigned l = 0; l < w->nlanes; ++l) {
    getcontext(&w->lane[l].ctx);
    w->lane[l].ctx.uc_stack.ss_sp = stacks + (size_t)l * EMUL_STACK;
    w->lane[l].ctx.uc_stack.ss_size = EMUL_STACK;
    w->lane[l].ctx.uc_link = &w->sched;
    makecontext(&w->lane[l].ctx, emul_lane_main, 0);
    w->lane[l].state = LANE_READY;
  }
  w->phase = 0;
  for (;;) {
    for (unsigned l = 0; l < w->nlanes; ++l) {
      if (w->lane[l].state != LANE_READY) continue;
      w->cur = l;
      threadIdx = {w->base_tid + l, 0, 0};
      swapcontext(&w->sched, &w->lane[l].ctx);
    }
    unsigned n_warp = 0, n_block = 0, n_done = 0;
    for (unsigned l = 0; l < w->nlanes; ++l) {
      n_warp += w->lane[l].state == LANE_WAIT_WARP;
      n_block += w->lane[l].state == LANE_WAIT_BLOCK;
      n_done += w->lane[l].state == LANE_DONE;
    }
    if (n_done == w->nlanes) return;
    if (n_warp && n_block) {
      fprintf(stderr, "f2g emulation: lanes of one warp wait at different synchronisation points\n");
      abort();
    }
    if (n_block) g_block_sync.arrive_and_wait();
    else ++w->phase;
    for (unsigned l = 0; l < w->nlanes; ++l)
      if (w->lane[l].state != LANE_DONE) w->lane[l].state = LANE_READY;
  }
}

template <typename F>
inline void emul_run_grid(dim3 grid, dim3 block, F&& body) {
  gridDim = {grid.x, grid.y, grid.z};
  blockDim = {block.x, block.y, block.z};
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  pthread_barrier_init(&g_block_end, nullptr, nwarps);
  g_block_sync.reset(nwarps);
  using Body = typename std::remove_reference<F>::type;
  std::vector<std::thread> th;
  th.reserve(nwarps);
  for (unsigned wi = 0; wi < nwarps; ++wi)
    th.emplace_back([&, wi]() {
      EmulWarp* w = new EmulWarp();
      char* stacks = static_cast<char*>(malloc(32 * EMUL_STACK));
      w->nlanes = nthreads - 32 * wi < 32 ? nthreads - 32 * wi : 32;
      w->base_tid = 32 * wi;
      w->trampoline = [](void* b) { (*static_cast<Body*>(b))(); };
      w->body = const_cast<void*>(static_cast<const void*>(&body));
      g_warp = w;
      for (unsigned b = 0; b < nblocks; ++b) {
        blockIdx = {F2G_EMUL_ORDER(b % grid.x, grid.x), (b / grid.x) % grid.y, b / (grid.x * grid.y)};
        emul_run_warp_block(w, stacks);
        g_block_sync.drop();                                   // this warp is out of the block's barriers
        // the block is over for every warp before its successor reuses the "shared memory" statics
        if (pthread_barrier_wait(&g_block_end) == PTHREAD_BARRIER_SERIAL_THREAD) g_block_sync.reset(nwarps);
        pthread_barrier_wait(&g_block_end);
      }
      g_warp = nullptr;
      free(stacks);
      delete w;
    });
  for (auto& x : th) x.join();
  pthread_barrier_destroy(&g_block_end);
}

// same call shape as common.cuh's launch_pdl (programmatic dependent launch has no host meaning)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args&&... args) {
  emul_run_grid(grid, block, [&]() { kern(static_cast<KArgs>(args)...); });
  return cudaSuccess;
}
}  // namespace f2g

#define F2G_DYN_SMEM(T, name) T* const name = reinterpret_cast<T*>(f2g::g_dyn_smem)
static inline void __syncthreads() { f2g::emul_yield(f2g::LANE_WAIT_BLOCK); }
static inline void __syncwarp() { f2g::emul_yield(f2g::LANE_WAIT_WARP); }
static inline float __shfl_xor_sync(unsigned, float v, int o) {
  f2g::EmulWarp* w = f2g::g_warp;
  const unsigned l = w->cur, p = w->phase & 1u;      // double-buffered: nobody is more than one phase ahead
  w->shfl[p][l] = v;
  f2g::emul_yield(f2g::LANE_WAIT_WARP);
  return w->shfl[p][l ^ (unsigned)o];
}
static inline float atomicAdd(float* p, float v) {
  std::lock_guard<std::mutex> lk(f2g::g_atomic_mu);
  const float old = *p;
  *p = old + v;
  return old;
}
static inline int atomicOr(int* p, int v) {
  std::lock_guard<std::mutex> lk(f2g::g_atomic_mu);
  const int old = *p;
  *p = old | v;
  return old;
}
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline int __float2int_rd(float v) { return (int)floorf(v); }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline void sincospif(float x, float* s, float* c) {
  *s = (float)sin(3.14159265358979323846 * (double)x);
  *c = (float)cos(3.14159265358979323846 * (double)x);
}
static inline float cospif(float x) { return (float)cos(3.14159265358979323846 * (double)x); }
template <typename T>
static inline cudaError_t cudaMemcpyToSymbol(T& symbol, const void* src, size_t n) {
  memcpy(&symbol, src, n);
  return cudaSuccess;
}
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename K>
static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

#define F2G_LAUNCH(kernel, grid, block, stream, ...)                            \
  do {                                                                          \
    (void)(stream);                                                             \
    const dim3 g_ = dim3(grid);                                                 \
    gridDim = {g_.x, g_.y, g_.z};                                               \
    blockDim = {(unsigned)(block), 1, 1};                                       \
    for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                                   \
      for (unsigned by_ = 0; by_ < g_.y; ++by_)                                 \
        for (unsigned b_ = 0; b_ < g_.x; ++b_)                                  \
          for (unsigned t_ = 0; t_ < blockDim.x; ++t_) {                        \
            blockIdx = {F2G_EMUL_ORDER(b_, g_.x), by_, bz_};                    \
            threadIdx = {F2G_EMUL_ORDER(t_, blockDim.x), 0, 0};                 \
            kernel(__VA_ARGS__);                                                \
          }                                                                     \
  } while (0)
#define F2G_LAUNCH_COOP(kernel, grid, block, stream, ...)                                         \
  do {                                                                                            \
    (void)(stream);                                                                               \
    f2g::emul_run_grid(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); });                  \
  } while (0)
#define F2G_LAUNCH_COOP_SMEM(kernel, grid, block, smem, stream, ...)                              \
  do {                                                                                            \
    (void)(stream);                                                                               \
    if ((size_t)(smem) > sizeof(f2g::g_dyn_smem)) abort();                                        \
    f2g::emul_run_grid(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); });                  \
  } while (0)

namespace f2g {
F2G_SIMT_DEV float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
F2G_SIMT_DEV float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// cvt.rna.tf32.f32 for finite inputs: nearest, ties away from zero, 10 explicit significand bits
F2G_SIMT_DEV float tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}
// cvt.rn.satfinite.f16x2.f32: nearest-even, clamped to the finite fp16 range
F2G_SIMT_DEV unsigned short half_bits_sat(float v) {
  if (v != v) return 0x7fffu;                      // NaN stays NaN (as the hardware conversion does)
  v = fminf(fmaxf(v, -65504.0f), 65504.0f);
  const _Float16 h = (_Float16)v;
  unsigned short b;
  memcpy(&b, &h, 2);
  return b;
}
F2G_SIMT_DEV uint2 pack_half4(float4 v) {
  return make_uint2((unsigned)half_bits_sat(v.x) | ((unsigned)half_bits_sat(v.y) << 16),
                    (unsigned)half_bits_sat(v.z) | ((unsigned)half_bits_sat(v.w) << 16));
}
F2G_SIMT_DEV unsigned half2_track(unsigned acc, unsigned w) {
  w &= 0x7fff7fffu;
  const unsigned lo = (acc & 0xffffu) > (w & 0xffffu) ? (acc & 0xffffu) : (w & 0xffffu);
  const unsigned hi = (acc >> 16) > (w >> 16) ? (acc >> 16) : (w >> 16);
  return lo | (hi << 16);
}
F2G_SIMT_DEV bool half2_out_of_range(unsigned acc) { return (acc & 0xffffu) >= 0x7bffu || (acc >> 16) >= 0x7bffu; }
F2G_SIMT_DEV float half_bits_to_float(unsigned short b) {
  _Float16 h;
  memcpy(&h, &b, 2);
  return (float)h;
}
F2G_SIMT_DEV float4 unpack_half4(uint2 u) {
  return make_float4(half_bits_to_float((unsigned short)(u.x & 0xffffu)), half_bits_to_float((unsigned short)(u.x >> 16)),
                     half_bits_to_float((unsigned short)(u.y & 0xffffu)), half_bits_to_float((unsigned short)(u.y >> 16)));
}
F2G_SIMT_DEV void pdl_wait() {}
F2G_SIMT_DEV void pdl_launch() {}
F2G_SIMT_DEV void cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
F2G_SIMT_DEV void cp_async_wait_all() {}
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
extern "C" __attribute__((weak)) const char* f2g_emul_last_error(void) { return f2g::g_emul_err; }
#endif

// global thread id / stride of a 1-D launch (64-bit: waveforms of hours still index correctly)
#define F2G_GTID ((long long)blockIdx.x * blockDim.x + threadIdx.x)
#define F2G_GSTRIDE ((long long)gridDim.x * blockDim.x)
