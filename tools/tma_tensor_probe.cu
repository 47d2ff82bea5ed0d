// Per-SM rate of 2-D TENSOR-MAP loads as the GEMM issues them (box = 64 fp16 x 128 rows = 16 KB, 128-byte
// swizzle, rows `pitch` bytes apart) against plain bulk copies of the same size -- is the GEMM's operand
// stream slower than a contiguous stream because every 128-byte row is its own L2 request?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_tensor_probe tools/tma_tensor_probe.cu -lcuda
#include <cuda.h>
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
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
constexpr uint32_t CHUNK = 16384;
// mode 0: tensor loads (box 64 x 128); mode 1: bulk 16 KB.  `warps` issuing warps, `stages` slots each.
__global__ void probe(const __grid_constant__ CUtensorMap map, const uint8_t* __restrict__ src, int row_tiles, int k_blocks,
                      int iters, int stages, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nw * stages; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) {
    uint8_t* mine = smem + (size_t)warp * stages * CHUNK;
    uint64_t* mb = bar + warp * stages;
    uint32_t rt = (blockIdx.x * 37u + warp * 11u) % row_tiles, kb = 0;
    auto issue = [&](int slot) {
      mbar_expect_tx(&mb[slot], CHUNK);
      if (mode == 0) tma_2d(mine + (size_t)slot * CHUNK, &map, &mb[slot], (int)kb * 64, (int)rt * 128);
      else bulk_load(mine + (size_t)slot * CHUNK, src + ((size_t)rt * k_blocks + kb) * CHUNK, CHUNK, &mb[slot]);
      if (++kb == (uint32_t)k_blocks) { kb = 0; rt = (rt + 97u) % row_tiles; }   // walk along K like a GEMM main loop
    };
    for (int s0 = 0; s0 < stages; ++s0) issue(s0);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      while (!mbar_try(&mb[s], ph)) {}
      issue(s);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    for (int k = 0; k < stages; ++k) {
      while (!mbar_try(&mb[s], ph)) {}
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  int dev = 0, sms = 0, khz = 0;
  CHECK(cudaGetDevice(&dev));
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
  printf("# %d SMs; 16 KB copies; tensor = box {64 fp16, 128 rows} SWIZZLE_128B over a (rows, K) fp16 matrix, bulk = contiguous 16 KB\n", sms);
  printf("# K     rows   MB    warps stages  mode     ms    TB/s  B/clk/SM\n");
  const int Ks[] = {384, 768, 2304};
  for (int ki = 0; ki < 3; ++ki) {
    const int K = Ks[ki], k_blocks = K / 64;
    const int rows = (48 << 20) / (K * 2) / 128 * 128;                      // ~48 MB matrix: L2 resident
    const int row_tiles = rows / 128;
    uint8_t* src = nullptr;
    CHECK(cudaMalloc(&src, (size_t)rows * K * 2));
    CHECK(cudaMemset(src, 1, (size_t)rows * K * 2));
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int mode = 0; mode < 2; ++mode)
      for (int w = 1; w <= 2; ++w)
        for (int stages = 2; stages <= 8; stages *= 2) {
          if (w * stages * 16 > 192) continue;
          const int iters = 2048;
          cudaEvent_t e0, e1;
          CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
          probe<<<sms, 32 * w, (size_t)w * stages * CHUNK>>>(map, src, row_tiles, k_blocks, 64, stages, mode);
          CHECK(cudaDeviceSynchronize());
          CHECK(cudaEventRecord(e0));
          probe<<<sms, 32 * w, (size_t)w * stages * CHUNK>>>(map, src, row_tiles, k_blocks, iters, stages, mode);
          CHECK(cudaEventRecord(e1));
          CHECK(cudaDeviceSynchronize());
          float ms = 0.f;
          CHECK(cudaEventElapsedTime(&ms, e0, e1));
          const double tbs = (double)sms * w * iters * CHUNK / (ms * 1e-3) / 1e12;
          printf("  %4d  %6d  %3d   %d     %d     %s  %6.3f  %5.2f   %5.1f\n", K, rows, (int)((size_t)rows * K * 2 >> 20), w, stages,
                 mode ? "bulk  " : "tensor", ms, tbs, tbs * 1e12 / sms / (khz * 1e3));
        }
    CHECK(cudaFree(src));
  }
  return 0;
}
