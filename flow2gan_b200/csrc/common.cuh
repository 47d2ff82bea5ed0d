// Shared device helpers for the flow2gan_b200 kernels (sm_100a only).
// Thin inline-PTX wrappers for mbarrier / TMA / tcgen05 + small math helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#ifndef F2G_DEVINL
#define F2G_DEVINL __device__ __forceinline__
#endif

namespace f2g {

// ---------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns 0 on success, a cudaError_t (>0) or a negative
// F2G_E* code otherwise; the last message is kept for f2g_last_error().
// ---------------------------------------------------------------------------------------
enum { F2G_OK = 0, F2G_EINVAL = -1, F2G_EDRIVER = -2, F2G_EARCH = -3 };
void set_error(const char* fmt, ...);
int check_launch(const char* what);
// Bring-up / timing-experiment knobs (tile-schedule variants, epilogue stubs, descriptor experiments) exist
// only in a -DF2G_BRINGUP build (F2G_BRINGUP=1 python -m flow2gan_b200._build --force, used by tools/);
// the product library ignores the environment and always takes the default.
#ifdef F2G_BRINGUP
#include <stdlib.h>
static inline int bringup_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
#else
static inline int bringup_int(const char*, int dflt) { return dflt; }
#endif
int* chain_watchdog_dev();   // mapped pinned {flag, problem, row tile, counter} (api.cu), nullptr if unavailable

// ---------------------------------------------------------------------------------------
// numerics
// ---------------------------------------------------------------------------------------
F2G_DEVINL float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Same rounding (nearest, ties away from zero) for finite inputs in two integer instructions;
// cvt.rna.tf32.f32 expands to a 4-5 instruction sequence with an Inf/NaN guard.  Inf/NaN inputs
// are not preserved -- only used on GEMM epilogue outputs.
F2G_DEVINL float tf32_rna_fast(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// fp32 -> fp16 pair, round-to-nearest-even, clamped to the finite fp16 range (operands of the
// kind::f16 GEMMs: an overflow saturates instead of poisoning the contraction with Inf)
F2G_DEVINL uint32_t pack_half2_sat(float a, float b) {
  uint32_t r;     // {b, a} -> upper / lower half; .satfinite clamps to +-65504 in the conversion
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
F2G_DEVINL uint2 pack_half4(float4 v) {
  return make_uint2(pack_half2_sat(v.x, v.y), pack_half2_sat(v.z, v.w));
}
// fp16 range guard on the CONVERTED words: the saturating conversion maps everything beyond the finite
// range to +-65504 (0x7bff) and NaN to NaN (> 0x7c00), so the running per-half maximum of |bits| reaching
// 0x7bff means "a value was clamped, is not finite, or sat exactly on the limit" -- one LOP3 + one VMAX
// per two values instead of float compares per value.
F2G_DEVINL uint32_t half2_track(uint32_t acc, uint32_t w) { return __vmaxu2(acc, w & 0x7fff7fffu); }
F2G_DEVINL bool half2_out_of_range(uint32_t acc) { return (acc & 0xffffu) >= 0x7bffu || (acc >> 16) >= 0x7bffu; }

F2G_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
F2G_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------
// shared-memory addresses, mbarrier
// ---------------------------------------------------------------------------------------
F2G_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// explicit shared-space accesses (never let a rounded generic pointer degrade to LD.E/ST.E)
F2G_DEVINL void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
F2G_DEVINL float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

F2G_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
F2G_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
F2G_DEVINL void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
F2G_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
F2G_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
F2G_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a hung pipeline traps instead of wedging the GPU until the watchdog.
F2G_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("f2g: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x,
             (int)threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2D / 3D / 4D tile loads into shared memory
// ---------------------------------------------------------------------------------------
F2G_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
F2G_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
F2G_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
F2G_DEVINL void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ---------------------------------------------------------------------------------------
F2G_DEVINL void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
F2G_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
F2G_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
F2G_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
F2G_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate, one CTA.
F2G_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
F2G_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread i <-> lane base+i).
F2G_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
F2G_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1 on sm_100, [49,52) base offset = 0, [61,64) layout type (2 = SWIZZLE_128B).
//   layout type 1 = SWIZZLE_128B_BASE32B: the only layout legal for MN-major TF32 operands
//   (32-byte chunks swizzled over 4-row groups; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
F2G_DEVINL uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
F2G_DEVINL uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor layout).
__host__ __device__ inline uint32_t make_idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format = F32
  d |= 2u << 7;                        // a_format = TF32
  d |= 2u << 10;                       // b_format = TF32
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}
// kind::f16 with fp16 operands (a_format = b_format = F16 = 0), fp32 accumulate, K-major
__host__ __device__ inline uint32_t make_idesc_f16(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format = F32
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch.  A kernel launched through launch_pdl() may become resident
// while its predecessor in the stream is still draining: everything before pdl_wait() (barrier
// init, TMEM allocation, descriptor prefetch, index math) overlaps the predecessor's tail.
// pdl_wait() returns once the predecessor grid has fully completed and its writes are visible,
// so it must precede the first global-memory access; it is a no-op for ordinary launches.
// pdl_launch() lets the successor start becoming resident as SMs free up.
// ---------------------------------------------------------------------------------------
F2G_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
F2G_DEVINL void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

F2G_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// host: launch `kern` with the programmatic-stream-serialization attribute (F2G_PDL=0 disables)
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace f2g
