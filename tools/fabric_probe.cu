// L2 -> SM delivery probe for the GEMM operand-byte question:
// how many bytes per second reach the SMs' shared memory when the CTAs of a cluster pull
//   mode 0  distinct chunks                         (per-SM ingest / chip L2 ceiling),
//   mode 1  the SAME chunk, every CTA with its own bulk copy (does the L2 de-duplicate?),
//   mode 2  the same chunk as cluster-multicast slices (one L2 read per cluster),
// for cluster sizes 1 / 2 / 4 / 8.  The working set (32 MB) stays L2-resident, nothing reads the
// staged data, so the number is the fabric's, not a kernel's.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fabric_probe tools/fabric_probe.cu && ./fabric_probe
//
// Prints one line per (mode, cluster): delivered GB/s summed over SMs, and per SM in B/clk.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

#ifndef PROBE_STAGES
#define PROBE_STAGES 4
#endif
#ifndef PROBE_CHUNK_KB
#define PROBE_CHUNK_KB 32
#endif
constexpr int STAGES = PROBE_STAGES;
constexpr int CHUNK = PROBE_CHUNK_KB * 1024;          // bytes per stage (32 KB = one GEMM pipeline stage per CTA)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(b, parity)) {
    if (++spins > (1u << 26)) {
      printf("fabric_probe: mbarrier timeout (block %d)\n", (int)blockIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t rank) {
  uint32_t addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(b)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// One issuing thread per CTA.  full[s]: CHUNK bytes landed in this CTA's slot s.  empty[s]: every CTA of
// the cluster has seen its full[s] for the previous use (so nobody's slot is overwritten early and no
// barrier receives bytes of two phases at once) -- one remote arrive per CTA per use.
__global__ void probe_kernel(const uint8_t* __restrict__ src, size_t n_chunks, int iters, int mode) {
  extern __shared__ __align__(128) uint8_t stage[];
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ __align__(8) uint64_t empty[STAGES];
  const uint32_t rank = cluster_rank(), csz = cluster_size();
  const size_t cluster_id = blockIdx.x / csz;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], csz);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  if (threadIdx.x == 0) {
    const uint32_t slice = CHUNK / csz;
    const uint16_t mask = (uint16_t)((1u << csz) - 1u);
    int s = 0;
    // iteration i: (re)issue load i into slot i % STAGES once the whole cluster has consumed the
    // slot's previous use, then consume the oldest outstanding load (i - STAGES + 1) and tell the cluster
    for (int i = 0; i < iters + STAGES; ++i) {
      if (i < iters) {
        if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
        mbar_expect_tx(&full[s], CHUNK);
        size_t c;
        if (mode == 0) c = ((size_t)blockIdx.x * 977 + (size_t)i * 131) % n_chunks;      // own chunk
        else c = (cluster_id * 977 + (size_t)i * 131) % n_chunks;                         // cluster-shared chunk
        const uint8_t* g = src + c * CHUNK;
        if (mode == 2 && csz > 1)
          bulk_load_mc(stage + (size_t)s * CHUNK + rank * slice, g + rank * slice, slice, &full[s], mask);
        else
          bulk_load(stage + (size_t)s * CHUNK, g, CHUNK, &full[s]);
      }
      const int j = i - (STAGES - 1);          // the oldest outstanding load
      if (j >= 0 && j < iters) {
        const int sj = j % STAGES;
        mbar_wait(&full[sj], (j / STAGES) & 1);
        for (uint32_t r = 0; r < csz; ++r) mbar_arrive_remote(&empty[sj], r);
      }
      if (++s == STAGES) s = 0;
    }
  }
  __syncthreads();
  cluster_sync();          // no CTA exits while a peer may still signal its barriers
}

// Plain LDG ingest: every thread streams 16-byte loads of an L2-resident buffer (8 independent loads in
// flight per thread), results folded into a register so nothing is optimised away.
__global__ void __launch_bounds__(1024, 1) ldg_kernel(const uint4* __restrict__ src, size_t n_vec, int iters, unsigned* sink) {
  unsigned acc = 0;
  size_t i = ((size_t)blockIdx.x * 7919u * 1024u + threadIdx.x) % n_vec;
  for (int it = 0; it < iters; ++it) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + (i + (size_t)j * 1024) % n_vec);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    i = (i + 8 * 1024) % n_vec;
  }
  if (acc == 0x12345678u) *sink = acc;
}

static float run(const uint8_t* src, size_t n_chunks, int mode, int csz, int sms, int iters, int* ctas_out) {
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = STAGES * CHUNK;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = csz;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(csz);
  int max_clusters = 0;
  CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, probe_kernel, &cfg));
  int clusters = sms / csz;
  if (max_clusters > 0 && clusters > max_clusters) clusters = max_clusters;
  cfg.gridDim = dim3(clusters * csz);
  *ctas_out = clusters * csz;
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  CHECK(cudaLaunchKernelEx(&cfg, probe_kernel, src, n_chunks, 64, mode));      // warm-up (also pulls src into L2)
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaEventRecord(e0));
  CHECK(cudaLaunchKernelEx(&cfg, probe_kernel, src, n_chunks, iters, mode));
  CHECK(cudaEventRecord(e1));
  CHECK(cudaDeviceSynchronize());
  float ms = 0.f;
  CHECK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CHECK(cudaGetDevice(&dev));
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const size_t n_chunks = 1024;                                   // 32 MB working set
  uint8_t* src = nullptr;
  CHECK(cudaMalloc(&src, n_chunks * CHUNK));
  CHECK(cudaMemset(src, 1, n_chunks * CHUNK));
  CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK));
  CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int iters = 4096;
  const char* names[3] = {"distinct chunks      ", "same chunk, unicast  ", "same chunk, multicast"};
  printf("# %d SMs, %.0f MHz nominal; %d KB stages x %d, %d iterations per CTA\n", sms, khz / 1e3, CHUNK / 1024,
         STAGES, iters);
  for (int csz = 1; csz <= 8; csz *= 2)
    for (int mode = 0; mode < 3; ++mode) {
      if (csz == 1 && mode == 2) continue;
      int ctas = 0;
      const float ms = run(src, n_chunks, mode, csz, sms, iters, &ctas);
      const double bytes = (double)ctas * iters * CHUNK;          // delivered into shared memory
      const double gbs = bytes / (ms * 1e-3) / 1e9;
      printf("cluster %d  %s  %3d CTAs  %8.3f ms  delivered %8.1f GB/s  = %5.1f B/clk/SM\n", csz, names[mode], ctas,
             ms, gbs, gbs * 1e9 / ctas / (khz * 1e3));
    }
  {  // LDG path, L2-resident 32 MB
    unsigned* sink = nullptr;
    CHECK(cudaMalloc(&sink, 4));
    const int it2 = 512;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const size_t n_vec = n_chunks * CHUNK / 16;
    ldg_kernel<<<sms, 1024>>>(reinterpret_cast<const uint4*>(src), n_vec, 16, sink);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    ldg_kernel<<<sms, 1024>>>(reinterpret_cast<const uint4*>(src), n_vec, it2, sink);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0.f;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)sms * 1024 * it2 * 8 * 16;
    const double gbs = bytes / (ms * 1e-3) / 1e9;
    printf("LDG.128 x8 per thread, 1024 thr/SM  %3d CTAs  %8.3f ms  delivered %8.1f GB/s  = %5.1f B/clk/SM\n", sms, ms,
           gbs, gbs * 1e9 / sms / (khz * 1e3));
  }
  CHECK(cudaFree(src));
  return 0;
}
