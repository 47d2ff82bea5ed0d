// How fast can ONE SM pull L2-resident data into shared memory with bulk copies, as a function of
// the copy size, the number of copies in flight and the number of issuing warps?  (Round-2 follow-up
// of tools/fabric_probe.cu, whose per-iteration time turned out to be independent of the bytes moved:
// 16 / 32 / 64 KB per iteration all took 0.5 us, i.e. it measured its own issue loop.)
//
// Every issuing warp (lane 0) owns `stages` slots of `chunk` bytes and keeps them all in flight:
// wait slot s, re-issue slot s (power-of-two address arithmetic only, no cross-CTA signalling).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate_probe tools/tma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void rate_kernel(const uint8_t* __restrict__ src, uint32_t ws_mask, int iters, int stages, uint32_t chunk, int spin) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nw * stages; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) {
    uint8_t* mine = smem + (size_t)warp * stages * chunk;
    uint64_t* mb = bar + warp * stages;
    uint32_t off = (blockIdx.x * 2654435761u + warp * 40503u * chunk) & ws_mask;
    for (int s = 0; s < stages; ++s) {
      mbar_expect_tx(&mb[s], chunk);
      bulk_load(mine + (size_t)s * chunk, src + (off & ~(chunk - 1)), chunk, &mb[s]);
      off = (off + 7 * chunk) & ws_mask;
    }
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (spin) { while (!mbar_test(&mb[s], ph)) {} } else { while (!mbar_try(&mb[s], ph)) {} }
      mbar_expect_tx(&mb[s], chunk);
      bulk_load(mine + (size_t)s * chunk, src + (off & ~(chunk - 1)), chunk, &mb[s]);
      off = (off + 7 * chunk) & ws_mask;
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    for (int k = 0; k < stages; ++k) {      // drain
      if (spin) { while (!mbar_test(&mb[s], ph)) {} } else { while (!mbar_try(&mb[s], ph)) {} }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CHECK(cudaGetDevice(&dev));
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const size_t ws = 64u << 20;                     // 64 MB working set: L2-resident (126 MB)
  uint8_t* src = nullptr;
  CHECK(cudaMalloc(&src, ws));
  CHECK(cudaMemset(src, 1, ws));
  CHECK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
  printf("# %d SMs, %d MHz nominal; bulk copies from a 64 MB L2-resident buffer; per-SM delivered rate\n", sms, khz / 1000);
  printf("# warps  chunkKB  stages  inflightKB  wait      ms     TB/s   B/clk/SM  us/copy/warp\n");
  const int chunks[] = {2, 4, 8, 16, 32, 64};
  const int warps[] = {1, 2, 4};
  for (int spin = 0; spin < 2; ++spin)
    for (int wi = 0; wi < 3; ++wi)
      for (int ci = 0; ci < 6; ++ci)
        for (int stages = 2; stages <= 16; stages *= 2) {
          const int w = warps[wi], ck = chunks[ci];
          const int inflight = w * stages * ck;
          if (inflight > 192 || inflight < 32 || w * stages > 64) continue;
          if (spin && !(w == 1 || ck == 16)) continue;         // the spin variant: a subset
          const int iters = 2048;
          cudaEvent_t e0, e1;
          CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
          rate_kernel<<<sms, 32 * w, (size_t)inflight * 1024>>>(src, (uint32_t)(ws - 1), 64, stages, ck * 1024, spin);
          CHECK(cudaDeviceSynchronize());
          CHECK(cudaEventRecord(e0));
          rate_kernel<<<sms, 32 * w, (size_t)inflight * 1024>>>(src, (uint32_t)(ws - 1), iters, stages, ck * 1024, spin);
          CHECK(cudaEventRecord(e1));
          CHECK(cudaDeviceSynchronize());
          float ms = 0.f;
          CHECK(cudaEventElapsedTime(&ms, e0, e1));
          const double bytes = (double)sms * w * (iters + stages) * ck * 1024.0;
          const double tbs = bytes / (ms * 1e-3) / 1e12;
          printf("  %d      %2d       %2d      %3d       %s  %7.3f  %6.2f   %6.1f     %6.3f\n", w, ck, stages, inflight,
                 spin ? "test_wait" : "try_wait ", ms, tbs, tbs * 1e12 / sms / (khz * 1e3), ms * 1e3 / (iters + stages));
        }
  CHECK(cudaFree(src));
  return 0;
}
