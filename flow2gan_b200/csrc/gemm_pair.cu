// CTA-pair (cta_group::2) grouped, persistent TF32 GEMM: 256 x BN output tiles on two SMs.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )          fp32 in HBM, TF32 operands, fp32 accumulate
//
// Why a pair: with fp32-container operands a 128x128 single-CTA tile moves 128 B of operands per
// tensor cycle through the L2->SM fabric (32 FLOP/B); measured on B200 the chip sustains about
// 10 TB/s there, which capped the one-CTA kernel (gemm_tf32.cu) at ~300 TF/s.  A CTA pair shares
// one 256-row MMA: each CTA stages its own 128 rows of A and only HALF of the B tile, so a
// 256x256 tile needs 64 B per tensor cycle per SM (64 FLOP/B).
//
// Same call sites as gemm_tf32.cu (every 1x1 conv / Linear of flow2gan/models/modules.py:443-451,
// 563-593 and the Conv2d contractions of flow2gan/models/discriminators.py:65-76,171-184).
//
// Roles per CTA (320 threads): warp 0 = TMA producer (both CTAs; all complete_tx land on the
// LEADER's full barrier), warp 1 = MMA issuer (leader CTA only, tcgen05.mma.cta_group::2, commits
// multicast to both CTAs), warps 2-9 = epilogue (two warps per TMEM lane quarter, alternating
// 32-column chunks).  TMEM: two 256-column accumulator buffers per CTA, so the epilogue of tile i
// overlaps the main loop of tile i+1.  BN is chosen per problem (32..256, multiple of 32).
#include "common.cuh"
#include "gemm_tf32.h"


namespace f2g {

constexpr int PBM = 128;                       // rows per CTA (pair tile = 256 rows)
constexpr int PBK = 32;                        // 32 fp32 = one 128 B swizzle row
constexpr int P_A_BYTES = PBM * PBK * 4;       // 16 KB
constexpr int P_STAGE_BYTES = 2 * P_A_BYTES;   // A + up to 128 B-tile rows
// -DF2G_EPI_WARPS=16 (with F2G_EPI_WARPS=16 python -m flow2gan_b200._build --force) builds a 20-warp CTA with four
// epilogue warps per scheduler (96 registers, 4 stages).  Measured round 2: identical launch and per-tile times --
// the epilogue is not limited by per-warp latency either (profiles/r02_gemm_probes.md).  Default 8.
#ifndef F2G_EPI_WARPS
#define F2G_EPI_WARPS 8
#endif
constexpr int P_STAGES = F2G_EPI_WARPS > 8 ? 4 : 5;      // 16 epilogue warps need 80 KB of epilogue scratch
constexpr int P_EPI_WARPS = F2G_EPI_WARPS;               // multiple of 4: one warp per TMEM lane quarter and column group
constexpr int P_EPI_GROUPS = P_EPI_WARPS / 4;
// TMA producer warps: warp 0 plus P_PROD_WARPS - 1 warps behind the epilogue warps.  One elected thread
// needs ~270 cycles per (wait, expect_tx, cp.async.bulk) round (tools/tma_rate_probe.cu: 0.137 us per
// copy per issuing warp, independent of the copy size up to 16 KB, scaling linearly with the number of
// issuing warps) -- with a single producer the 2 (K-major) to 8 (MN-major: 32-row boxes) copies of a
// pipeline stage took as long as or longer than the stage's 512 tensor cycles.  The copies of a stage
// are dealt round-robin to the producer warps.
constexpr int P_PROD_WARPS = 3;       // 12 warps per CTA: register allocation is per 4 warps (13 warps cost the budget of 16)
constexpr int P_FIRST_EXTRA_PROD = 2 + P_EPI_WARPS;            // warp index of producer 1
constexpr int P_THREADS = 64 + 32 * P_EPI_WARPS + 32 * (P_PROD_WARPS - 1);
// per epilogue warp: a 32 x 36 fp32 transpose pad (4608 B) or one 32-row x 128 B swizzled fp16 box for a TMA store
// (4096 B, 1024 B aligned: SWIZZLE_128B keys on address bits 7..9)
constexpr int P_SCRATCH_WARP_BYTES = 5120;
constexpr int P_SCRATCH_BYTES = P_EPI_WARPS * P_SCRATCH_WARP_BYTES;
constexpr int P_PARAM_BYTES = 2 * 3 * 256 * 4;     // bias / slope / residual-scale of a tile's columns, two tiles (double buffer)
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + P_SCRATCH_BYTES + P_PARAM_BYTES;
constexpr int P_TMEM_COLS = 512;
constexpr int P_MAX_SCHED = 1024;
constexpr int P_MAX_PAIRS = 78;

struct alignas(64) PProblem {
  CUtensorMap map_a;
  CUtensorMap map_b;
  CUtensorMap map_c;        // fp16 destination as 64-column x 32-row store boxes (tma_c)
  float* c;
  float* c_pre;
  const float* bias;
  const float* slope;
  const float* res;
  const float* res_scale;
  const float* row_scale;
  const float* gate;
  int ldc, ld_res, ld_gate, ld_pre;
  int M, N, K, bn;
  int m_tiles, n_tiles, tile_begin;
  int split_k, kb_total, kb_per;
  int act, round_tf32, accumulate;
  float leaky, alpha;
  int seg_len, seg_shift;   // windowed A operand (F2GGemm::a_seg_len / a_seg_shift), 0 = plain
  int c_f16;                // C stored as fp16 (F2GGemm::c_f16)
  int* done;                // chaining (F2GGemm::done_counter / wait_counter)
  const int* wait;
  int wait_count;
  int* sat_flag;            // fp16 range guard of a c_f16 destination (F2GGemm::sat_flag)
  int tma_c;                // bias+activation -> fp16 tiles leave through TMA stores (epilogue below)
};

struct alignas(64) PGroup {
  PProblem p[F2G_GEMM_MAX_PROBLEMS];
  int n_problems;
  int total_tiles;
  // Static longest-processing-time schedule (host-computed; groups of <= P_MAX_SCHED tiles):
  // pair p runs tiles sched[pair_off[p] .. pair_off[p+1]).  Tiles of one launch differ up to 3x
  // in cost (K = 2304 vs 1152 ...), so round-robin left ~40 % of the SM-time of the pwconv2
  // group idle.  use_sched == 0: plain round-robin (tile = pair + i * npairs).
  int use_sched;
  uint16_t pair_off[P_MAX_PAIRS + 2];
  uint32_t sched[P_MAX_SCHED];   // host-decoded tiles: problem (3 bits) | K split (9) | row tile (10) | column tile (10)
  int dbg;   // bring-up (F2G_PAIR_DBG): bit0 = epilogue drains TMEM but stores nothing,
             // bit1 = epilogue skips TMEM loads too
  int* watchdog;   // mapped pinned host ints {flag, problem, row tile, counter seen} or nullptr
};

// Spin bound of a chained consumer tile: 2^22 polls x (>= 64 ns sleep + one L2 round trip) is seconds, five
// orders of magnitude beyond any launch of this library.  Reaching it means the producer CTAs are not
// running (two chained launches sharing the device, see the header): report and trap instead of hanging.
constexpr uint32_t P_CHAIN_SPIN_LIMIT = 1u << 22;

struct PTile {
  int prob, m0, n0, kb0, kb1;
};

F2G_DEVINL PTile pdecode(const PGroup& g, int tile) {
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < g.n_problems; ++i)
    if (tile >= g.p[i].tile_begin) pi = i;
  int local = tile - g.p[pi].tile_begin;
  const int nt = g.p[pi].n_tiles;
  const int mn = nt * g.p[pi].m_tiles;
  const int ks = local / mn;
  local -= ks * mn;
  PTile t;
  t.prob = pi;
  t.m0 = (local / nt) * (2 * PBM);
  t.n0 = (local % nt) * g.p[pi].bn;
  t.kb0 = ks * g.p[pi].kb_per;
  t.kb1 = min(t.kb0 + g.p[pi].kb_per, g.p[pi].kb_total);
  return t;
}

// A scheduled tile, decoded on the host (build_schedule): the in-kernel decode above -- a search over the
// problems and three integer divisions -- was ~600 cycles per tile on the epilogue warps' critical path.
F2G_DEVINL PTile pdecode_packed(const PGroup& g, uint32_t e) {
  PTile t;
  t.prob = (int)(e & 7u);
  const PProblem& p = g.p[t.prob];
  const int ks = (int)((e >> 3) & 511u);
  t.m0 = (int)((e >> 12) & 1023u) * (2 * PBM);
  t.n0 = (int)(e >> 22) * p.bn;
  t.kb0 = ks * p.kb_per;
  t.kb1 = min(t.kb0 + p.kb_per, p.kb_total);
  return t;
}
F2G_DEVINL PTile ptile(const PGroup& g, int s_off, int ti, int pair, int npairs) {
  return g.use_sched ? pdecode_packed(g, g.sched[s_off + ti]) : pdecode(g, pair + ti * npairs);
}

// ------------------------------- cluster / cta_group::2 PTX ---------------------------------
F2G_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
F2G_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\t"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
F2G_DEVINL uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Default (.release.cta) semantics on purpose: a cluster-scope release compiles to MEMBAR.ALL.GPU,
// which made every epilogue warp wait for all of its global stores once per tile (ncu: 7 % of
// the samples).  The only thing this arrive orders is the warp's tcgen05.ld traffic, which is
// fenced by tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
F2G_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load whose completion bytes are credited to an mbarrier of the pair's leader CTA
F2G_DEVINL void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
F2G_DEVINL void tmem_alloc_cg2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
F2G_DEVINL void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
F2G_DEVINL void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
F2G_DEVINL void umma_tf32_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
F2G_DEVINL void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs
F2G_DEVINL void umma_commit_cg2(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// PEPI_MLP (fp16 operands only): per problem either BIAS_ACT with an fp16 destination or BIAS_RES --
// the two halves of a chained pwconv1 -> pwconv2 launch, both on their fast paths.
#ifdef F2G_BRINGUP
#define F2G_PROF_DECL long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long prof_c = clock64()
#define F2G_PROF(slot) do { const long long n_ = clock64(); prof_t[slot] += n_ - prof_c; prof_c = n_; } while (0)
#else
#define F2G_PROF_DECL do { } while (0)
#define F2G_PROF(slot) do { } while (0)
#endif

// TMA store of one shared-memory box (bulk async-group completion), and the waits on this thread's groups:
// .read = the source buffers may be overwritten; plain = the global writes are complete.
F2G_DEVINL void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
F2G_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
F2G_DEVINL void tma_store_wait_done() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// non-blocking phase test (try_wait may suspend the thread for a system-dependent time before it says no)
F2G_DEVINL bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
F2G_DEVINL void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
F2G_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
F2G_DEVINL void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

enum { PEPI_GENERIC = 0, PEPI_BIAS_ACT = 1, PEPI_BIAS_RES = 2, PEPI_PLAIN = 3, PEPI_MLP = 4 };

F2G_DEVINL float4 lds_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
F2G_DEVINL int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// orders the async proxy (TMA reads issued after it) behind what this thread has observed
// through the generic proxy (the acquire above)
F2G_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Fast epilogue of one 32x32 chunk, after the transpose through `scratch`: this lane owns the
// column quad `col` of rows rsub, rsub+4, ..., rsub+28.  Straight-line on purpose (one epilogue
// warp or two per SM sub-partition: nothing hides a dependent chain or a branch): all eight
// LDS.128 / LDG.128 are issued up front, rounding is a mask instead of a branch, the optional
// pre-activation store is its own pass.  FULL = all 32 rows valid (no predicates at all).
//   x = alpha*acc + bias ; BIAS_ACT: [pre = x] x = prelu(x) ; BIAS_RES: x += rsc * res ;
//   PLAIN: x += C (if rp) ; x = tf32(x) if asked ; C = x
template <int EPI, bool FULL>
F2G_DEVINL void epi_fast_chunk(const float* __restrict__ sl, float* __restrict__ cp, size_t cstep,
                               const float* __restrict__ rp, size_t rstep, float* __restrict__ pp,
                               size_t pstep, int rows_left, float alpha, float4 bias4, float4 slope4,
                               float4 rsc4, uint32_t radd, uint32_t rmask) {
  float4 xv[8], rv[8];
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) xv[rr] = *reinterpret_cast<const float4*>(sl + rr * (4 * 36));
  if (EPI != PEPI_BIAS_ACT) {
    if (rp) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr)
        rv[rr] = (FULL || rr * 4 < rows_left) ? *reinterpret_cast<const float4*>(rp + rr * rstep)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) rv[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    xv[rr].x = fmaf(xv[rr].x, alpha, bias4.x); xv[rr].y = fmaf(xv[rr].y, alpha, bias4.y);
    xv[rr].z = fmaf(xv[rr].z, alpha, bias4.z); xv[rr].w = fmaf(xv[rr].w, alpha, bias4.w);
  }
  if (EPI == PEPI_BIAS_ACT) {
    if (pp) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr)
        if (FULL || rr * 4 < rows_left) *reinterpret_cast<float4*>(pp + rr * pstep) = xv[rr];
    }
  }
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    float4 x = xv[rr];
    if (EPI == PEPI_BIAS_ACT) {
      x.x = x.x > 0.f ? x.x : x.x * slope4.x; x.y = x.y > 0.f ? x.y : x.y * slope4.y;
      x.z = x.z > 0.f ? x.z : x.z * slope4.z; x.w = x.w > 0.f ? x.w : x.w * slope4.w;
    } else if (EPI == PEPI_BIAS_RES) {
      x.x = fmaf(rsc4.x, rv[rr].x, x.x); x.y = fmaf(rsc4.y, rv[rr].y, x.y);
      x.z = fmaf(rsc4.z, rv[rr].z, x.z); x.w = fmaf(rsc4.w, rv[rr].w, x.w);
    } else {
      x.x += rv[rr].x; x.y += rv[rr].y; x.z += rv[rr].z; x.w += rv[rr].w;
    }
    x.x = __uint_as_float((__float_as_uint(x.x) + radd) & rmask);
    x.y = __uint_as_float((__float_as_uint(x.y) + radd) & rmask);
    x.z = __uint_as_float((__float_as_uint(x.z) + radd) & rmask);
    x.w = __uint_as_float((__float_as_uint(x.w) + radd) & rmask);
    if (FULL || rr * 4 < rows_left) *reinterpret_cast<float4*>(cp + rr * cstep) = x;
  }
}

// F16 = 1: fp16 operands (kind::f16, 64 elements per 128 B swizzle row, K = 16 per instruction);
// the byte geometry of the pipeline (stages, descriptors, 32 B K-steps) is the TF32 one.
// BIAS_ACT chunk with an fp16 destination (the hidden activation of a ConvNeXt block, which only
// the next GEMM reads): x = prelu(alpha*acc + bias) -> RN fp16, one 8-byte store per row quad.
template <bool FULL>
F2G_DEVINL uint32_t epi_fast_chunk_h(const float* __restrict__ sl, __half* __restrict__ cp, size_t cstep,
                                     int rows_left, float alpha, float4 bias4, float4 slope4, uint32_t sat_acc) {
  float4 xv[8];
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) xv[rr] = *reinterpret_cast<const float4*>(sl + rr * (4 * 36));
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    float4 x;
    x.x = fmaf(xv[rr].x, alpha, bias4.x); x.y = fmaf(xv[rr].y, alpha, bias4.y);
    x.z = fmaf(xv[rr].z, alpha, bias4.z); x.w = fmaf(xv[rr].w, alpha, bias4.w);
    x.x = x.x > 0.f ? x.x : x.x * slope4.x; x.y = x.y > 0.f ? x.y : x.y * slope4.y;      // NaN propagates
    x.z = x.z > 0.f ? x.z : x.z * slope4.z; x.w = x.w > 0.f ? x.w : x.w * slope4.w;
    if (FULL || rr * 4 < rows_left) {
      const uint2 hw = pack_half4(x);
      *reinterpret_cast<uint2*>(cp + rr * cstep) = hw;
      sat_acc = half2_track(half2_track(sat_acc, hw.x), hw.y);     // fp16 range guard (common.cuh)
    }
  }
  return sat_acc;
}

template <int A_MN, int B_MN, int EPI, int F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ PGroup g) {
  static_assert(!F16 || (!A_MN && !B_MN), "fp16 operands are K-major only");
  constexpr int KELEM = F16 ? 64 : PBK;        // contraction elements per pipeline stage
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("f2g gemm_pair: dynamic smem base not 1024B aligned\n");
    __trap();
  }
  __shared__ __align__(8) uint64_t full_bar[P_STAGES];     // used in the leader CTA only
  __shared__ __align__(8) uint64_t empty_bar[P_STAGES];    // per CTA, multicast commit
  __shared__ __align__(8) uint64_t tmem_full_bar[2];       // per CTA, multicast commit
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];      // leader CTA only, 16 arrivals
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int s_off = g.use_sched ? g.pair_off[pair] : 0;
  const int my_tiles = g.use_sched ? (int)g.pair_off[pair + 1] - s_off
                                   : (g.total_tiles > pair ? (g.total_tiles - pair + npairs - 1) / npairs : 0);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.n_problems; ++i) {
      tma_prefetch_desc(&g.p[i].map_a);
      tma_prefetch_desc(&g.p[i].map_b);
      if (g.p[i].tma_c) tma_prefetch_desc(&g.p[i].map_c);
    }
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * P_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(&tmem_base_smem, P_TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();                  // everything above overlapped the previous kernel's tail
  pdl_launch();

  const int prod = warp == 0 ? 0 : (warp >= P_FIRST_EXTRA_PROD ? warp - P_FIRST_EXTRA_PROD + 1 : -1);
  if (prod >= 0) {
    // ------------------------------- TMA producers (both CTAs) -------------------------
    // Copy ops of one stage: A first (1 K-major box of 128 rows, or PBM/32 MN-major boxes), then B
    // (1 box of bn/2 rows, or bn/64 MN-major boxes); op i belongs to producer i % P_PROD_WARPS.
    // Producer 0 of the leader CTA posts the stage's expected byte count (both CTAs' copies).
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const PTile tc = ptile(g, s_off, ti, pair, npairs);
        const PProblem& pr = g.p[tc.prob];
        const int bhalf = pr.bn >> 1;
        const int m_cta = tc.m0 + (int)rank * PBM;
        const int n_cta = tc.n0 + (int)rank * bhalf;
        const uint32_t stage_tx = 2u * (uint32_t)(P_A_BYTES + bhalf * PBK * 4);
        constexpr int n_a = A_MN ? PBM / 32 : 1;
        const int n_b = B_MN ? (bhalf >> 5) : 1;
        const int n_kb = tc.kb1 - tc.kb0;
        if (prod >= n_a + n_b) {            // nothing to copy for this tile: just keep the ring position
          stage += n_kb;
          phase ^= (uint32_t)((stage / P_STAGES) & 1);
          stage %= P_STAGES;
          continue;
        }
        bool mine_a = false;
#pragma unroll
        for (int j = 0; j < n_a; ++j) mine_a |= (j % P_PROD_WARPS) == prod;
        if (pr.wait && mine_a) {   // chained consumer: the A rows of this 256-row tile come from a producer
                                   // problem of this launch; wait until all of its tiles over them are stored
          const int* wp = pr.wait + tc.m0 / (2 * PBM);
          uint32_t polls = 0;
          while (ld_acquire_gpu(wp) < pr.wait_count) {
            __nanosleep(64);
            if (++polls > P_CHAIN_SPIN_LIMIT) {
              if (g.watchdog) {
                volatile int* wd = g.watchdog;
                wd[1] = tc.prob; wd[2] = tc.m0 / (2 * PBM); wd[3] = ld_acquire_gpu(wp);
                __threadfence_system();
                wd[0] = 1;
                __threadfence_system();
              }
              __trap();
            }
          }
          fence_proxy_async_all();
        }
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * P_STAGE_BYTES;
          uint8_t* sb = sa + P_A_BYTES;
          if (rank == 0 && prod == 0) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          const int kc = kb * KELEM;                    // beyond K: TMA zero-fills
          // windowed A (implicit im2col): contraction segment s = one kernel row, which lives
          // seg_shift buffer rows further; within a segment consecutive rows overlap
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < PBM / 32; ++j) {
              if ((j % P_PROD_WARPS) != prod) continue;
              int mm = m_cta + 32 * j, kk = kc;
              if (pr.seg_len) {
                const int s = mm / pr.seg_len;
                mm -= s * pr.seg_len;
                kk += s * pr.seg_shift;
              }
              tma_load_2d_cg2(sa + j * 4096, &pr.map_a, lbar, mm, kk);
            }
          } else if (prod == 0) {
            int kca = kc, ma = m_cta;
            if (pr.seg_len) {
              const int s = kc / pr.seg_len;
              kca -= s * pr.seg_len;
              ma += s * pr.seg_shift;
            }
            tma_load_2d_cg2(sa, &pr.map_a, lbar, kca, ma);
          }
          if (B_MN) {
            for (int j = 0; j < (bhalf >> 5); ++j)
              if (((n_a + j) % P_PROD_WARPS) == prod) tma_load_2d_cg2(sb + j * 4096, &pr.map_b, lbar, n_cta + 32 * j, kc);
          } else if ((n_a % P_PROD_WARPS) == prod) {
            tma_load_2d_cg2(sb, &pr.map_b, lbar, kc, n_cta);
          }
          if (++stage == P_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) ----------------------
    if (rank == 0 && elect_one()) {
      const uint32_t a_addr0 = smem_u32(smem), b_addr0 = a_addr0 + P_A_BYTES;
      // MN-major TF32 operands: SWIZZLE_128B_BASE32B, 32-wide MN blocks 4096 B apart (LBO),
      // 4-row swizzle atoms 512 B apart (SBO), one k-step (8 rows) = 1024 B  (see gemm_tf32.cu)
      const uint64_t adesc0 = A_MN ? make_smem_desc(a_addr0, 4096, 512, 1) : make_smem_desc_sw128(a_addr0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(b_addr0, 4096, 512, 1) : make_smem_desc_sw128(b_addr0, 16, 1024);
      constexpr int a_kstep = A_MN ? 1024 : 32, b_kstep = B_MN ? 1024 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int ab = 0;
      uint32_t ab_phase = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const PTile tc = ptile(g, s_off, ti, pair, npairs);
        const uint32_t idesc = F16 ? make_idesc_f16(2 * PBM, g.p[tc.prob].bn)
                                   : make_idesc_tf32(2 * PBM, g.p[tc.prob].bn, A_MN, B_MN);
        mbar_wait(&tmem_empty_bar[ab], ab_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + ab * 256;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t soff = (uint32_t)(stage * P_STAGE_BYTES) >> 4;
#pragma unroll
          for (int k = 0; k < PBK / 8; ++k) {
            const uint64_t adesc = adesc0 + soff + (uint32_t)((k * a_kstep) >> 4);
            const uint64_t bdesc = bdesc0 + soff + (uint32_t)((k * b_kstep) >> 4);
            if (F16) umma_f16_cg2(tmem_d, adesc, bdesc, idesc, (kb > tc.kb0 || k) ? 1u : 0u);
            else umma_tf32_cg2(tmem_d, adesc, bdesc, idesc, (kb > tc.kb0 || k) ? 1u : 0u);
          }
          umma_commit_cg2(&empty_bar[stage]);     // frees this smem slot in both CTAs
          if (++stage == P_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_cg2(&tmem_full_bar[ab]);      // accumulator complete -> both epilogues
        ab ^= 1;
        if (ab == 0) ab_phase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue warps (both CTAs) ------------------------
    // TMEM lane = output row of this CTA's 128-row half.  Warps (q, half) drain lanes
    // [32q, 32q+32) in 32-column chunks c0 = 32*(2i + half): tcgen05.ld -> 8 x STS.128 into a
    // padded scratch -> re-read as 4 rows x 8 column-quads so every LDG/STG is 128-bit and one
    // warp instruction covers four full 128 B row segments.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;                    // column group: chunks half, half + P_EPI_GROUPS, ...
    float* const scratch = reinterpret_cast<float*>(smem + P_STAGES * P_STAGE_BYTES + ew * P_SCRATCH_WARP_BYTES);
    float* const sparam0 = reinterpret_cast<float*>(smem + P_STAGES * P_STAGE_BYTES + P_SCRATCH_BYTES);
    const int cg = lane & 7, rsub = lane >> 3;
    const int et = ew * 32 + lane;
    const uint32_t lead_empty0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t lead_empty1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    int ab = 0;
    uint32_t ab_phase = 0;
    F2G_PROF_DECL;
    // Per-tile bookkeeping runs ONE TILE AHEAD (measured with the bring-up cycle profile: tile decode 550
    // + parameter staging 640 cycles per tile sat between the end of one tile and the wait for the next
    // accumulator, 15 % of an epilogue-bound tile): tile i+1 is decoded and its bias / slope /
    // residual-scale columns are fetched into the other half of the parameter buffer while tile i's
    // accumulator is still being produced; the barrier that ends tile i publishes them.
    // the columns are fetched with 4-byte LDGSTS: the warp does not wait for them (measured: the plain
    // load -> STS version cost ~900 cycles per tile on the epilogue's critical path); complete before the
    // barrier that publishes the buffer (cp_async_wait_all below)
    auto stage_params = [&](const PTile& t, float* sp) {
      const PProblem& p = g.p[t.prob];
      const float* const bias_p = p.bias;
      const float* const slope_p = p.slope;
      const float* const rsc_p = p.res_scale;
      const float leaky = p.leaky;
      const int bn = p.bn, nn = p.N;
      for (int c = et; c < bn; c += 32 * P_EPI_WARPS) {
        const int colc = t.n0 + c;
        const bool okc = colc < nn;
        if (bias_p && okc) cp_async4(sp + c, bias_p + colc); else sp[c] = 0.f;
        if (slope_p && okc) cp_async4(sp + 256 + c, slope_p + colc); else sp[256 + c] = leaky;
        if (rsc_p && okc) cp_async4(sp + 512 + c, rsc_p + colc); else sp[512 + c] = 1.f;
      }
    };
    // TMA-store epilogue (PProblem::tma_c) state of this warp: lane 0 owns the bulk async-groups of the warp's
    // stores.  `stg_busy`: a store may still be reading the staging box; `pend_done`: counter of a finished
    // producer tile whose publish waits for the completion of its stores -- deferred so that nobody stalls on
    // the write latency, but never past a point where this CTA could block on its own consumers (below).
    int* pend_done = nullptr;
    bool stg_busy = false;
    // publish the pending tile once its stores are complete, optionally allowing the `keep` most recent groups
    // (a later tile's own boxes) to stay in flight.  The completion of a TMA store takes ~1.4 us (measured:
    // waiting for it right after the tile cost 2700 cycles per tile), so a producer tile is published at the
    // next tile's first box (store_sync) -- or at once when this CTA is about to idle or to depend on it, or
    // at the end of the next tile if that tile had no box for this warp
    auto publish = [&](int keep) {
      if (lane == 0) {
        if (keep >= 2) asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
        else if (keep == 1) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
        else tma_store_wait_done();
        F2G_PROF(6);
        // the completed stores are in L2, the point of coherence of every reader of this launch (TMA loads of
        // the consumer tiles, after ld.acquire.gpu + fence.proxy.async): a relaxed L2 atomic behind the
        // completion orders the publish after the data.  (A release fence here -- MEMBAR.ALL.GPU -- also
        // waited for this tile's in-flight stores: +1400 cycles per tile.)
        if (g.dbg & 32) { fence_proxy_async_all(); __threadfence(); }
        atomicAdd(pend_done, 1);
      }
      pend_done = nullptr;
      if (keep == 0) stg_busy = false;
      __syncwarp();
    };
    auto store_sync = [&]() {              // everything of this warp retired: nothing pending, staging box free
      if (pend_done) publish(0);
      else if (stg_busy) {
        if (lane == 0) tma_store_wait_read();
        stg_busy = false;
        __syncwarp();
      }
    };
    PTile tc = {0, 0, 0, 0, 0};
    if (my_tiles > 0) {
      tc = ptile(g, s_off, 0, pair, npairs);
      stage_params(tc, sparam0);
      cp_async_wait_all();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * P_EPI_WARPS) : "memory");
    }
    for (int ti = 0; ti < my_tiles; ++ti) {
      const PProblem& pr = g.p[tc.prob];
      const int BN = pr.bn;
      float* const sparam = sparam0 + (ti & 1) * 768;
      F2G_PROF(0);
      const int n0 = tc.n0;
      const int row_base = tc.m0 + (int)rank * PBM + q * 32;
      const int N = pr.N, ldc = pr.ldc, ld_res = pr.ld_res, ld_gate = pr.ld_gate, ld_pre = pr.ld_pre;
      const int rows = min(32, pr.M - row_base);
      float* const cbase = pr.c;
      float* const pre_p = pr.c_pre;
      const float* const res_p = pr.res;
      const float* const rowsc_p = pr.row_scale;
      const float* const gate_p = pr.gate;
      const int act = pr.act;
      const bool do_round = pr.round_tf32 != 0, do_acc = pr.accumulate != 0;
      const float alpha = pr.alpha;
      const int split_k = pr.split_k;
      const bool c16 = F16 && pr.c_f16 != 0;       // fp16 destination (never in the TF32 instantiations)
      const uint32_t radd = do_round ? 0x1000u : 0u, rmask = do_round ? 0xffffe000u : 0xffffffffu;
      const bool vec_c = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(cbase) & 15) == 0);
      const bool vec_res = !res_p || (((ld_res & 3) == 0) && ((reinterpret_cast<uintptr_t>(res_p) & 15) == 0));
      const bool vec_gate = !gate_p || (((ld_gate & 3) == 0) && ((reinterpret_cast<uintptr_t>(gate_p) & 15) == 0));
      const bool vec_pre = !pre_p || (((ld_pre & 3) == 0) && ((reinterpret_cast<uintptr_t>(pre_p) & 15) == 0));
      const bool vec_all = vec_c && vec_res && vec_gate && vec_pre;
      int* const done_p = pr.done;
      const int done_idx = tc.m0 / (2 * PBM);
      int* const sat_p = pr.sat_flag;

      // the next tile's decode + parameter columns, under this tile's main loop
      PTile tn = tc;
      if (ti + 1 < my_tiles) {
        tn = ptile(g, s_off, ti + 1, pair, npairs);
        stage_params(tn, sparam0 + ((ti + 1) & 1) * 768);
      }
      F2G_PROF(1);

      // TMA-store path: per problem -- except for the last such tile before this pair turns to tiles that may
      // wait for it (or runs out of tiles): a TMA store completes ~1.4 us after its issue, plain stores behind
      // a fence are visible in a third of that, and at the phase boundary of a chained launch that latency is
      // MMA idle time (measured: 34 us of chain waits per inference step against 18 us)
      const bool tma_c = F16 && (EPI == PEPI_BIAS_ACT || EPI == PEPI_MLP) && pr.tma_c != 0 &&
                         (done_p == nullptr || (ti + 1 < my_tiles && g.p[tn.prob].tma_c != 0));
      // a pending publish must not wait behind an accumulator that may itself be waiting for it (a consumer
      // tile of this launch whose A rows this CTA produced): if the accumulator is not ready yet, publish now
      if (pend_done && !mbar_test(&tmem_full_bar[ab], ab_phase)) store_sync();
      mbar_wait(&tmem_full_bar[ab], ab_phase);
      tc_fence_after();
      F2G_PROF(2);
      uint32_t sat_acc = 0;    // fp16 range guard of a c_f16 destination (F2GGemm::sat_flag; common.cuh)
      bool released = false;   // TMEM buffer already handed back (store path: after the last tcgen05.ld)
      int boxes = 0;           // store groups this warp commits in this tile (rows <= 0: none, the count is unused)
      if (!tma_c && (pend_done || stg_busy)) store_sync();
      if (tma_c) {
        // bias + PReLU -> fp16 tile through TMA stores.  The SM's shared-memory pipe is what short-K tiles are
        // bound by (the operand ring alone uses ~80 % of it, profiles/r02_gemm_probes.md): this path moves
        // 2 + 2 bytes per output through it (fp16 staging write + the store engine's read) instead of
        // 4 + 4 (fp32 transpose) + 2 (STG), and needs no transpose at all: lane = row, the thread converts
        // its 64 columns in registers (column parameters as broadcast LDS.128) and writes its 128 B row
        // of a SWIZZLE_128B box; lane 0 issues one 32-row x 64-column store per box.
        const uint32_t stg = smem_u32(scratch);
        const uint32_t row_addr = stg + (uint32_t)lane * 128u;
        const uint32_t swz = (uint32_t)(lane & 7);
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ab * 256;
        const int ncols = min(BN, N - n0);
        boxes = 0;
#pragma unroll 1
        for (int cb = half * 64; cb < ncols; cb += 64 * P_EPI_GROUPS) {
          if (g.dbg & 2) break;
          uint32_t hw[32];
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            uint32_t v[32];
            tmem_ld_32x32(trow + cb + 32 * hc, v);
            tmem_ld_wait();
            if (hc == 1 && cb + 64 * P_EPI_GROUPS >= ncols) {
              // that was this warp's last read of the accumulator: hand the TMEM buffer back to the MMA
              // issuer now, not after the math / staging / store of this box and the tile-end barrier
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(ab ? lead_empty1 : lead_empty0);
              released = true;
            }
            const float* sp = sparam + cb + 32 * hc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = lds_f4(sp + 4 * j);
              const float4 s4 = lds_f4(sp + 256 + 4 * j);
              float4 x;
              x.x = fmaf(__uint_as_float(v[4 * j]), alpha, b4.x); x.y = fmaf(__uint_as_float(v[4 * j + 1]), alpha, b4.y);
              x.z = fmaf(__uint_as_float(v[4 * j + 2]), alpha, b4.z); x.w = fmaf(__uint_as_float(v[4 * j + 3]), alpha, b4.w);
              x.x = x.x > 0.f ? x.x : x.x * s4.x; x.y = x.y > 0.f ? x.y : x.y * s4.y;      // NaN propagates
              x.z = x.z > 0.f ? x.z : x.z * s4.z; x.w = x.w > 0.f ? x.w : x.w * s4.w;
              const uint2 w2 = pack_half4(x);
              hw[16 * hc + 2 * j] = w2.x;
              hw[16 * hc + 2 * j + 1] = w2.y;
              sat_acc = half2_track(half2_track(sat_acc, w2.x), w2.y);
            }
          }
          F2G_PROF(3);
          if (g.dbg & 1) continue;
          // the previous tile's stores were issued a decode, an accumulator wait and a box of math ago
          // (>= 1.5 us): publish it here rather than at the end of this tile (a tile later is late enough to
          // hold up consumers on other pairs); otherwise only the engine's read of the previous box matters
          if (pend_done || stg_busy) store_sync();
          F2G_PROF(4);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            sts_u4(row_addr + ((((uint32_t)i) ^ swz) << 4), hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && rows > 0) tma_store_2d(&pr.map_c, stg, n0 + cb, row_base);
          stg_busy = true;
          ++boxes;
          F2G_PROF(5);
        }
      } else
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 32 * P_EPI_GROUPS) {
        if (n0 + c0 >= N || (g.dbg & 2)) break;
        // (issuing the next chunk's tcgen05.ld here, before this chunk's LDS / math / stores, was measured:
        // no change -- the TMEM read is not what the epilogue waits for; profiles/r02_gemm_probes.md)
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ab * 256 + c0, v);
        tmem_ld_wait();
        F2G_PROF(3);
        if (rows <= 0 || (g.dbg & 1)) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(scratch + lane * 36 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
        __syncwarp();
        const int col = n0 + c0 + 4 * cg;
        const int ncol = min(4, N - col);
        // opaque 128-bit loads: left to itself the compiler sinks the slope load into the PReLU selects as
        // 32 predicated LDS.32 per chunk -- a shared-memory round trip inside every
        // FFMA -> compare -> multiply -> convert chain
        const float4 bias4 = lds_f4(sparam + c0 + 4 * cg);
        const float4 slope4 = lds_f4(sparam + 256 + c0 + 4 * cg);
        const float4 rsc4 = lds_f4(sparam + 512 + c0 + 4 * cg);
        const bool chunk_full = vec_all && (n0 + c0 + 32 <= N) && split_k == 1;   // warp-uniform

        if (F16 && (EPI == PEPI_BIAS_ACT || EPI == PEPI_MLP) && chunk_full && c16) {
          const float* sl = scratch + rsub * 36 + 4 * cg;
          __half* hp = reinterpret_cast<__half*>(cbase) + (size_t)(row_base + rsub) * ldc + col;
          F2G_PROF(4);
          if (rows >= 32) sat_acc = epi_fast_chunk_h<true>(sl, hp, (size_t)4 * ldc, 32, alpha, bias4, slope4, sat_acc);
          else sat_acc = epi_fast_chunk_h<false>(sl, hp, (size_t)4 * ldc, rows - rsub, alpha, bias4, slope4, sat_acc);
          __syncwarp();
          F2G_PROF(5);
          continue;
        }
        constexpr int EF = EPI == PEPI_MLP ? PEPI_BIAS_RES : EPI;   // fp32-destination fast path
        if (EF != PEPI_GENERIC && chunk_full && !c16 && (EPI != PEPI_MLP || act == F2G_ACT_NONE)) {
          const size_t r0 = (size_t)(row_base + rsub);
          const float* sl = scratch + rsub * 36 + 4 * cg;
          float* cp = cbase + r0 * ldc + col;
          const float* rp = EF == PEPI_BIAS_RES ? (res_p ? res_p + r0 * ld_res + col : nullptr)
                                                : ((EF == PEPI_PLAIN && do_acc) ? cp : nullptr);
          float* pp = (EF == PEPI_BIAS_ACT && pre_p) ? pre_p + r0 * ld_pre + col : nullptr;
          const size_t rstep = EF == PEPI_BIAS_RES ? (size_t)4 * ld_res : (size_t)4 * ldc;
          if (rows >= 32)
            epi_fast_chunk<EF, true>(sl, cp, (size_t)4 * ldc, rp, rstep, pp, (size_t)4 * ld_pre, 32, alpha,
                                     bias4, slope4, rsc4, radd, rmask);
          else
            epi_fast_chunk<EF, false>(sl, cp, (size_t)4 * ldc, rp, rstep, pp, (size_t)4 * ld_pre,
                                      rows - rsub, alpha, bias4, slope4, rsc4, radd, rmask);
          __syncwarp();
          continue;
        }

        const float bias[4] = {bias4.x, bias4.y, bias4.z, bias4.w};
        const float slope[4] = {slope4.x, slope4.y, slope4.z, slope4.w};
        const float rsc[4] = {rsc4.x, rsc4.y, rsc4.z, rsc4.w};
        if (split_k > 1) {   // partial-K tile: atomically accumulate into the pre-zeroed C
#pragma unroll 1
          for (int rr = 0; rr < 8; ++rr) {
            const int i = rr * 4 + rsub;
            if (i >= rows) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < ncol)
                atomicAdd(cbase + (size_t)(row_base + i) * ldc + col + e, scratch[i * 36 + 4 * cg + e] * alpha);
          }
          __syncwarp();
          continue;
        }
        const bool quad = vec_all && ncol == 4;
        if (quad) {
#pragma unroll 2
          for (int rr = 0; rr < 8; ++rr) {
            const int i = rr * 4 + rsub;
            if (i >= rows) continue;
            const int row = row_base + i;
            const float4 xq = *reinterpret_cast<const float4*>(scratch + i * 36 + 4 * cg);
            float x[4] = {fmaf(xq.x, alpha, bias[0]), fmaf(xq.y, alpha, bias[1]), fmaf(xq.z, alpha, bias[2]),
                          fmaf(xq.w, alpha, bias[3])};
            if (pre_p)
              *reinterpret_cast<float4*>(pre_p + (size_t)row * ld_pre + col) = make_float4(x[0], x[1], x[2], x[3]);
            if (act == F2G_ACT_PRELU || act == F2G_ACT_LEAKY) {
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] = x[e] > 0.f ? x[e] : x[e] * slope[e];
            } else if (act == F2G_ACT_SILU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] = x[e] / (1.f + __expf(-x[e]));
            }
            if (gate_p) {  // multiply by d(act)/dz evaluated at a saved pre-activation
              const float4 t = __ldg(reinterpret_cast<const float4*>(gate_p + (size_t)row * ld_gate + col));
              x[0] *= (t.x > 0.f ? 1.f : slope[0]); x[1] *= (t.y > 0.f ? 1.f : slope[1]);
              x[2] *= (t.z > 0.f ? 1.f : slope[2]); x[3] *= (t.w > 0.f ? 1.f : slope[3]);
            }
            if (res_p) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(res_p + (size_t)row * ld_res + col));
              x[0] = fmaf(rsc[0], t.x, x[0]); x[1] = fmaf(rsc[1], t.y, x[1]);
              x[2] = fmaf(rsc[2], t.z, x[2]); x[3] = fmaf(rsc[3], t.w, x[3]);
            }
            if (rowsc_p) {
              const float rs = __ldg(rowsc_p + row);
              x[0] *= rs; x[1] *= rs; x[2] *= rs; x[3] *= rs;
            }
            if (c16) {
              const uint2 hw = pack_half4(make_float4(x[0], x[1], x[2], x[3]));
              *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(cbase) + (size_t)row * ldc + col) = hw;
              sat_acc = half2_track(half2_track(sat_acc, hw.x), hw.y);
              continue;
            }
            float* dst = cbase + (size_t)row * ldc + col;
            if (do_acc) {
              const float4 t = *reinterpret_cast<const float4*>(dst);
              x[0] += t.x; x[1] += t.y; x[2] += t.z; x[3] += t.w;
            }
            if (do_round) {
              x[0] = tf32_rna_fast(x[0]); x[1] = tf32_rna_fast(x[1]); x[2] = tf32_rna_fast(x[2]); x[3] = tf32_rna_fast(x[3]);
            }
            *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
          }
        } else if (ncol > 0) {   // unaligned leading dimension or N tail: scalar path
#pragma unroll 1
          for (int rr = 0; rr < 8; ++rr) {
            const int i = rr * 4 + rsub;
            if (i >= rows) continue;
            const int row = row_base + i;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (e >= ncol) continue;
              float x = fmaf(scratch[i * 36 + 4 * cg + e], alpha, bias[e]);
              if (pre_p) pre_p[(size_t)row * ld_pre + col + e] = x;
              if (act == F2G_ACT_PRELU || act == F2G_ACT_LEAKY) x = x > 0.f ? x : x * slope[e];
              else if (act == F2G_ACT_SILU) x = x / (1.f + __expf(-x));
              if (gate_p) x *= (__ldg(gate_p + (size_t)row * ld_gate + col + e) > 0.f ? 1.f : slope[e]);
              if (res_p) x = fmaf(rsc[e], __ldg(res_p + (size_t)row * ld_res + col + e), x);
              if (rowsc_p) x *= __ldg(rowsc_p + row);
              if (c16) {
                const uint32_t hw = pack_half2_sat(x, 0.f);
                reinterpret_cast<unsigned short*>(cbase)[(size_t)row * ldc + col + e] = (unsigned short)(hw & 0xffffu);
                sat_acc = half2_track(sat_acc, hw);
                continue;
              }
              float* dst = cbase + (size_t)row * ldc + col + e;
              if (do_acc) x += *dst;
              *dst = do_round ? tf32_rna_fast(x) : x;
            }
          }
        }
        __syncwarp();
      }
      F2G_PROF(6);
      if (F16 && sat_p && half2_out_of_range(sat_acc)) atomicOr(sat_p, 1);
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ab ? lead_empty1 : lead_empty0);
      }
      cp_async_wait_all();
      // one barrier per tile: the next tile's staged parameters become visible, this tile's buffer half
      // may be refilled (by the tile after next), and all 256 threads' stores precede the publish below
      asm volatile("bar.sync 1, %0;" ::"n"(32 * P_EPI_WARPS) : "memory");
      if (tma_c) {               // every warp publishes its own rows once its stores have completed
        if (pend_done) publish(rows > 0 ? boxes : 0);
        pend_done = done_p ? done_p + done_idx : nullptr;
      } else if (et == 0 && done_p) {
        __threadfence();
        atomicAdd(done_p + done_idx, pr.tma_c ? P_EPI_WARPS : 1);     // store-path problems count warps
      }
      ab ^= 1;
      if (ab == 0) ab_phase ^= 1;
      tc = tn;
      F2G_PROF(7);
    }
    if (pend_done || stg_busy) store_sync();   // shared memory must outlive the store engine's reads
#ifdef F2G_BRINGUP
    if ((g.dbg & 16) && blockIdx.x == 0 && lane == 0 && (ew == 0 || ew == 5))
      printf("epi prof warp %d tiles %d cycles: decode %lld params+bar %lld wait_mma %lld ldtm %lld sts+lds_params %lld "
             "chunk_math+stg %lld other_paths %lld tile_end %lld\n", ew, my_tiles, prof_t[0], prof_t[1], prof_t[2],
             prof_t[3], prof_t[4], prof_t[5], prof_t[6], prof_t[7]);
#endif
  }

  // Neither CTA may exit (or free TMEM) while its peer can still signal its barriers / read its
  // shared memory through the pair MMA.
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, P_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*PEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// f2g_gemm_plan (host logic only, no device needed): the group is planned -- tile geometry, chaining, the
// per-pair schedule -- exactly as for a launch, but tensor maps are not encoded and nothing is launched.
struct PlanOut {
  int pairs;
  PGroup* g;
  int used_pairs;
};
static thread_local PlanOut* g_plan_out = nullptr;

static PEncodeTiledFn pair_encode_fn() {
  static PEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%d)", (int)e);
      return nullptr;
    }
    fn = reinterpret_cast<PEncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map; dims/strides innermost first; box = {32, box_rows}.  The row pitch may be
// SMALLER than the row length (overlapping rows: the implicit im2col view of a strided conv).
static int pair_encode_2d(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer,
                          uint64_t outer_stride_elems, uint32_t box_rows, bool mn_major, bool f16 = false) {
  const uint64_t esize = f16 ? 2 : 4;
  if (g_plan_out == nullptr && pair_encode_fn() == nullptr) return F2G_EDRIVER;
  PEncodeTiledFn fn = g_plan_out ? nullptr : pair_encode_fn();
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (outer_stride_elems * esize) % 16 != 0) {
    set_error("gemm operand must be 16B aligned with a row pitch multiple of 16 bytes "
              "(ptr=%p ld=%llu %s)", (const void*)base, (unsigned long long)outer_stride_elems,
              f16 ? "fp16" : "fp32");
    return F2G_EINVAL;
  }
  if (!fn) return 0;                       // plan only
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {outer_stride_elems * esize};
  cuuint32_t box[2] = {f16 ? 64u : 32u, box_rows};      // 128 B = one swizzle row either way
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu box_rows=%u", (int)r,
              (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)outer_stride_elems, box_rows);
    return F2G_EDRIVER;
  }
  return 0;
}

// N tile of a problem: the operand bytes a pair streams per output column are ~ (256 + bn) / bn,
// so wide tiles win unless the last tile is mostly padding.  B_MN needs 32-column TMA boxes per
// CTA (bn multiple of 64).
static int pick_bn(int N, bool b_mn) {
  const int step = b_mn ? 64 : 32;
  int best = 256;
  long best_cost = -1;
  for (int bn = 256; bn >= step; bn -= step) {
    const long tiles = (N + bn - 1) / bn;
    const long cost = tiles * (256 + bn);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}


// Greedy LPT: tiles sorted by decreasing cost, each to the least-loaded pair (binary heap).
static void build_schedule(PGroup& g, int pairs) {
  g.use_sched = 0;
  const int T = g.total_tiles;
  if (T > P_MAX_SCHED || pairs > P_MAX_PAIRS) return;
  static thread_local int cost[P_MAX_SCHED];
  static thread_local uint16_t order[P_MAX_SCHED];
  // tiles are laid out problem by problem (problems already sorted by decreasing K), every tile
  // of a problem costs the same: the natural order IS the decreasing-cost order unless N tiles
  // differ, so a stable counting pass is enough
  int n = 0;
  bool sorted = true;
  for (int pi = 0; pi < g.n_problems; ++pi) {
    const PProblem& p = g.p[pi];
    const int cnt = p.m_tiles * p.n_tiles * p.split_k;
    // waiting (phase 1) tiles sort behind every phase-0 tile: LPT within a phase, phases in order
    const int c = p.kb_per * (256 + p.bn) + 8 * p.bn + (p.wait ? 0 : (1 << 24));
    for (int i = 0; i < cnt; ++i) {
      cost[n] = c;
      order[n] = (uint16_t)n;
      if (n && cost[n - 1] < c) sorted = false;
      ++n;
    }
  }
  const bool round_robin = (cost[0] == cost[n - 1] && sorted) || T <= pairs;   // homogeneous tiles / one wave
  if (!sorted) {
    for (int i = 1; i < n; ++i) {          // insertion sort by cost (few distinct values, mostly sorted)
      const uint16_t o = order[i];
      int j = i;
      while (j > 0 && cost[order[j - 1]] < cost[o]) { order[j] = order[j - 1]; --j; }
      order[j] = o;
    }
  }
  long load[P_MAX_PAIRS];
  int heap[P_MAX_PAIRS];
  uint16_t owner[P_MAX_SCHED];
  int count[P_MAX_PAIRS];
  for (int i = 0; i < pairs; ++i) { load[i] = 0; heap[i] = i; count[i] = 0; }
  for (int i = 0; i < n; ++i) {
    if (round_robin) {
      owner[i] = (uint16_t)(i % pairs);
      ++count[i % pairs];
      continue;
    }
    const int p = heap[0];                // least-loaded pair (ties: lowest index first)
    owner[i] = (uint16_t)p;
    load[p] += cost[order[i]] & ((1 << 24) - 1);
    ++count[p];
    int k = 0;                            // sift down
    for (;;) {
      int l = 2 * k + 1, r = l + 1, m = k;
      if (l < pairs && (load[heap[l]] < load[heap[m]] || (load[heap[l]] == load[heap[m]] && heap[l] < heap[m]))) m = l;
      if (r < pairs && (load[heap[r]] < load[heap[m]] || (load[heap[r]] == load[heap[m]] && heap[r] < heap[m]))) m = r;
      if (m == k) break;
      const int t = heap[k]; heap[k] = heap[m]; heap[m] = t;
      k = m;
    }
  }
  int off = 0;
  int start[P_MAX_PAIRS];
  for (int i = 0; i < pairs; ++i) { g.pair_off[i] = (uint16_t)off; start[i] = off; off += count[i]; }
  g.pair_off[pairs] = (uint16_t)off;
  for (int i = 0; i < n; ++i) {
    const int tile = order[i];
    int pi = 0;
    for (int j = 1; j < g.n_problems; ++j)
      if (tile >= g.p[j].tile_begin) pi = j;
    const PProblem& p = g.p[pi];
    int local = tile - p.tile_begin;
    const int mn = p.n_tiles * p.m_tiles;
    const int ks = local / mn;
    local -= ks * mn;
    const int mt = local / p.n_tiles, nt = local % p.n_tiles;
    if (ks > 511 || mt > 1023 || nt > 1023 || pi > 7) return;      // does not fit the packed entry: round-robin
    g.sched[start[owner[i]]++] = (uint32_t)pi | ((uint32_t)ks << 3) | ((uint32_t)mt << 12) | ((uint32_t)nt << 22);
  }
  g.use_sched = 1;
}

template <int A_MN, int B_MN, int EPI, int F16 = 0>
static int pair_launch(PGroup& g, cudaStream_t stream) {
  if (g_plan_out) {
    const int pairs = g.total_tiles < g_plan_out->pairs ? g.total_tiles : g_plan_out->pairs;
    build_schedule(g, pairs);
    *g_plan_out->g = g;
    g_plan_out->used_pairs = pairs;
    return 0;
  }
  static int max_pairs = 0;
  auto kern = gemm_pair_kernel<A_MN, B_MN, EPI, F16>;
  if (!max_pairs) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_pair smem=%d): %s", P_SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(sms & ~1);
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = P_SMEM_BYTES;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    if (e != cudaSuccess || nc < 1) {
      cudaGetLastError();
      nc = sms / 2;
    }
    nc = bringup_int("F2G_PAIRS", nc);
    max_pairs = nc < 1 ? 1 : nc;
  }
  const int pairs = g.total_tiles < max_pairs ? g.total_tiles : max_pairs;
  static const int no_sched = bringup_int("F2G_PAIR_RR", 0);
  if (!no_sched) build_schedule(g, pairs);
  cudaError_t le = launch_pdl(kern, dim3(2 * pairs), dim3(P_THREADS), (size_t)P_SMEM_BYTES, stream, g);
  if (le != cudaSuccess) {
    set_error("gemm_pair launch: %s", cudaGetErrorString(le));
    return (int)le;
  }
  return check_launch("gemm_pair");
}

int gemm_pair_group(const F2GGemm* descs, int n, cudaStream_t stream) {
  if (n < 1 || n > F2G_GEMM_MAX_PROBLEMS) {
    set_error("gemm group size %d out of range", n);
    return F2G_EINVAL;
  }
  PGroup g;
  memset(&g, 0, sizeof(g));
  const int a_mn = descs[0].a_mn, b_mn = descs[0].b_mn;
  const bool f16 = descs[0].ab_f16 != 0;
  if (f16 && (a_mn || b_mn)) {
    set_error("fp16 gemm operands must be K-major (a_mn = b_mn = 0)");
    return F2G_EINVAL;
  }
  const int kelem = f16 ? 64 : PBK;
  // heaviest tiles first: with a static round-robin schedule the long-K problems must not land
  // in the last (partial) wave
  // Chained groups: every problem that waits on a counter (phase 1) comes after all the others
  // (phase 0) -- in the tile numbering AND in every pair's schedule -- so no CTA ever waits for a
  // tile that is queued behind a waiting tile.
  int order[F2G_GEMM_MAX_PROBLEMS];
  bool chained = false;
  for (int i = 0; i < n; ++i) { order[i] = i; chained |= descs[i].wait_counter != nullptr; }
  auto before = [&](int a, int b) {   // a must precede b
    const int pa = descs[a].wait_counter ? 1 : 0, pb = descs[b].wait_counter ? 1 : 0;
    return pa != pb ? pa < pb : descs[a].K > descs[b].K;
  };
  for (int i = 1; i < n; ++i)
    for (int j = i; j > 0 && before(order[j], order[j - 1]); --j) {
      const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
    }
  // N tile per problem.  Few-row groups (CondEncoder: 6 row tiles) would leave most CTA pairs
  // idle with the byte-optimal wide tiles and are bound by one tile's latency chain: narrow the
  // tiles (>= 64 columns) until the group has at least ~one tile per pair.
  int bns[F2G_GEMM_MAX_PROBLEMS];
  {
    static const int fill = bringup_int("F2G_PAIR_FILL", 1);
    const int step = b_mn ? 64 : 32;
    // Chained consumer problems take the caller's N tile (F2GGemm::bn) instead of the byte-optimal
    // one: narrower tiles for the problems the LPT schedule places last shorten the tail of the
    // launch (tools/sched_sim.py: -5 % modelled; measured -3.5 % of the inference step).
    const int bn_hint = 1;
    for (int oi = 0; oi < n; ++oi) {
      const F2GGemm& d = descs[order[oi]];
      bns[oi] = pick_bn(d.N, b_mn != 0);
      if (bn_hint && d.wait_counter && d.bn >= 64 && d.bn <= 256 && d.bn % step == 0) bns[oi] = d.bn;
    }
    for (int round = 0; fill && round < 3; ++round) {
      long tl = 0;
      for (int oi = 0; oi < n; ++oi) {
        const F2GGemm& d = descs[order[oi]];
        const int sk = d.split_k < 1 ? 1 : d.split_k;
        tl += (long)((d.M + 2 * PBM - 1) / (2 * PBM)) * ((d.N + bns[oi] - 1) / bns[oi]) * sk;
      }
      if (tl >= 56) break;
      bool changed = false;
      for (int oi = 0; oi < n; ++oi) {
        const int nb = ((bns[oi] / 2 + step - 1) / step) * step;
        if (nb >= 64 && nb < bns[oi]) { bns[oi] = nb; changed = true; }
      }
      if (!changed) break;
    }
  }
  int tiles = 0;
  for (int oi = 0; oi < n; ++oi) {
    const F2GGemm& d = descs[order[oi]];
    if (d.a_mn != a_mn || d.b_mn != b_mn || (d.ab_f16 != 0) != f16) {
      set_error("all problems of a gemm group must share operand majors and operand type");
      return F2G_EINVAL;
    }
    if (d.c_f16 && (!f16 || d.res || d.gate || d.accumulate || d.c_pre || d.split_k > 1 || d.row_scale)) {
      set_error("c_f16 needs ab_f16 and a bias / bias+activation epilogue");
      return F2G_EINVAL;
    }
    if (f16 && d.a_seg_len) {
      set_error("windowed operands are fp32 only");
      return F2G_EINVAL;
    }
    if (d.M <= 0 || d.N <= 0 || d.K <= 0) {
      set_error("gemm problem %d has empty shape %dx%dx%d", order[oi], d.M, d.N, d.K);
      return F2G_EINVAL;
    }
    PProblem& p = g.p[oi];
    const int bn = bns[oi];
    int rc;
    if (d.a_seg_len && ((d.a_seg_len & 31) || d.a_rows <= 0)) {
      set_error("windowed gemm operand: a_seg_len=%d must be a multiple of 32 and a_rows=%d > 0",
                d.a_seg_len, d.a_rows);
      return F2G_EINVAL;
    }
    if (d.a_seg_len)
      rc = a_mn ? pair_encode_2d(&p.map_a, d.a, d.a_seg_len, d.a_rows, d.lda, 32, true)
                : pair_encode_2d(&p.map_a, d.a, d.a_seg_len, d.a_rows, d.lda, PBM, false);
    else
      rc = a_mn ? pair_encode_2d(&p.map_a, d.a, d.M, d.K, d.lda, 32, true)
                : pair_encode_2d(&p.map_a, d.a, d.K, d.M, d.lda, PBM, false, f16);
    if (rc) return rc;
    p.seg_len = d.a_seg_len; p.seg_shift = d.a_seg_shift;
    p.c_f16 = d.c_f16;
    p.sat_flag = d.c_f16 ? d.sat_flag : nullptr;
    rc = b_mn ? pair_encode_2d(&p.map_b, d.b, d.N, d.K, d.ldb, 32, true)
              : pair_encode_2d(&p.map_b, d.b, d.K, d.N, d.ldb, bn / 2, false, f16);
    if (rc) return rc;
    // fp16 bias+activation destinations leave through TMA stores when the tile splits into whole 64-column boxes
    static const int no_tma_c = bringup_int("F2G_PAIR_NO_TMA_STORE", 0);
    p.tma_c = 0;
    // (N % 8: measured -- a store box clipped at an extent that is not a multiple of 16 bytes zeroed the rest of
    // that 16-byte granule outside the tensor; tests/test_kernels_gpu.py keeps N = 250 on the plain-store path)
    if (!no_tma_c && f16 && d.c_f16 && d.bias && (d.act == F2G_ACT_PRELU || d.act == F2G_ACT_LEAKY) && bn % 64 == 0 &&
        (d.N & 7) == 0 && (d.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(d.c) & 15) == 0) {
      rc = pair_encode_2d(&p.map_c, d.c, d.N, d.M, d.ldc, 32, false, true);
      if (rc) return rc;
      p.tma_c = 1;
    }
    p.c = d.c; p.ldc = d.ldc; p.c_pre = d.c_pre; p.ld_pre = d.ld_pre;
    p.bias = d.bias; p.slope = d.slope; p.res = d.res; p.res_scale = d.res_scale;
    p.row_scale = d.row_scale; p.gate = d.gate;
    p.ld_res = d.ld_res; p.ld_gate = d.ld_gate;
    p.M = d.M; p.N = d.N; p.K = d.K; p.bn = bn;
    p.m_tiles = (d.M + 2 * PBM - 1) / (2 * PBM);
    p.n_tiles = (d.N + bn - 1) / bn;
    p.tile_begin = tiles;
    p.act = d.act; p.round_tf32 = d.round_tf32; p.accumulate = d.accumulate;
    p.leaky = d.leaky; p.alpha = d.alpha == 0.f ? 1.f : d.alpha;
    p.kb_total = (d.K + kelem - 1) / kelem;
    int sk = d.split_k < 1 ? 1 : d.split_k;
    if (sk > p.kb_total) sk = p.kb_total;
    p.kb_per = (p.kb_total + sk - 1) / sk;
    p.split_k = (p.kb_total + p.kb_per - 1) / p.kb_per;
    if (p.split_k > 1 && (d.bias || d.act || d.res || d.gate || d.row_scale || d.c_pre || d.round_tf32)) {
      set_error("split-K gemm supports the plain (alpha) epilogue only");
      return F2G_EINVAL;
    }
    p.done = d.done_counter; p.wait = d.wait_counter; p.wait_count = 0;
    if ((p.done || p.wait) && p.split_k > 1) {
      set_error("chained gemm problems cannot be split along K");
      return F2G_EINVAL;
    }
    tiles += p.m_tiles * p.n_tiles * p.split_k;
  }
  for (int oi = 0; oi < n; ++oi) {   // resolve consumers to their producers
    PProblem& c = g.p[oi];
    if (!c.wait) continue;
    const PProblem* prod = nullptr;
    for (int oj = 0; oj < n; ++oj)
      if (g.p[oj].done == c.wait && !g.p[oj].wait) prod = &g.p[oj];
    if (!prod || prod->M != c.M) {
      set_error("gemm wait_counter without a producer (done_counter of a non-waiting problem with the same M) in the group");
      return F2G_EINVAL;
    }
    c.wait_count = 2 * prod->n_tiles * (prod->tma_c ? P_EPI_WARPS : 1);   // per-warp publishes on the store path
    static const int no_wait = bringup_int("F2G_PAIR_NO_CHAIN_WAIT", 0);   // timing experiment: what the chain waits cost
    if (no_wait) c.wait_count = 0;
  }
  g.n_problems = n;
  g.total_tiles = tiles;
  g.watchdog = (chained && !g_plan_out) ? chain_watchdog_dev() : nullptr;
  static const int dbg = bringup_int("F2G_PAIR_DBG", 0);
  g.dbg = dbg;

  int epi = -1;
  bool mlp_ok = true;
  for (int i = 0; i < n; ++i) {
    const F2GGemm& d = descs[i];
    int e = PEPI_GENERIC;
    const bool odd = d.gate || d.row_scale || d.act == F2G_ACT_SILU;
    if (!odd && d.bias && (d.act == F2G_ACT_PRELU || d.act == F2G_ACT_LEAKY) && !d.res && !d.accumulate)
      e = PEPI_BIAS_ACT;
    else if (!odd && d.act == F2G_ACT_NONE && (d.res || d.bias) && !d.accumulate && !d.c_pre)
      e = PEPI_BIAS_RES;
    else if (!odd && d.act == F2G_ACT_NONE && !d.res && !d.bias && !d.c_pre)
      e = PEPI_PLAIN;
    if (d.c_f16 && e != PEPI_BIAS_ACT) e = PEPI_GENERIC;
    if (f16 && e == PEPI_BIAS_ACT && !d.c_f16) mlp_ok = false;
    if (e != PEPI_BIAS_ACT && e != PEPI_BIAS_RES) mlp_ok = false;
    epi = (epi == -1 || epi == e) ? e : PEPI_GENERIC;
  }
  if (f16 && epi == PEPI_GENERIC && mlp_ok) epi = PEPI_MLP;   // mixed pwconv1 / pwconv2 problems
  static const int force_generic = bringup_int("F2G_PAIR_FORCE_GENERIC", 0);
  if (force_generic) epi = PEPI_GENERIC;     // timing experiments only
  if (f16) {   // K-major only; the three epilogues the inference blocks use
    if (epi == PEPI_BIAS_ACT) return pair_launch<0, 0, PEPI_BIAS_ACT, 1>(g, stream);
    if (epi == PEPI_BIAS_RES) return pair_launch<0, 0, PEPI_BIAS_RES, 1>(g, stream);
    if (epi == PEPI_MLP) return pair_launch<0, 0, PEPI_MLP, 1>(g, stream);
    return pair_launch<0, 0, PEPI_GENERIC, 1>(g, stream);
  }

#define F2G_PDISPATCH_E(AM_, BM_)                                                        \
  {                                                                                      \
    if (epi == PEPI_BIAS_ACT) return pair_launch<AM_, BM_, PEPI_BIAS_ACT>(g, stream);    \
    if (epi == PEPI_BIAS_RES) return pair_launch<AM_, BM_, PEPI_BIAS_RES>(g, stream);    \
    if (epi == PEPI_PLAIN) return pair_launch<AM_, BM_, PEPI_PLAIN>(g, stream);          \
    return pair_launch<AM_, BM_, PEPI_GENERIC>(g, stream);                               \
  }
  if (!a_mn && !b_mn) F2G_PDISPATCH_E(0, 0)
  if (!a_mn && b_mn) F2G_PDISPATCH_E(0, 1)
  if (a_mn && b_mn) F2G_PDISPATCH_E(1, 1)
  F2G_PDISPATCH_E(1, 0)
#undef F2G_PDISPATCH_E
}

}  // namespace f2g

// out: {use_sched, n_problems, total_tiles, pairs used} ; per problem (launch order) 12 ints {M, N, K, bn,
// m_tiles, n_tiles, tile_begin, split_k, waits (0/1), wait_count, publishes (0/1), tma_c} ; pairs + 1 offsets ;
// then total_tiles entries -- packed (pdecode_packed's format) when use_sched, else absent.
extern "C" int f2g_gemm_plan(const F2GGemm* problems, int n_problems, int pairs, int* out, int out_ints) {
  using namespace f2g;
  if (pairs < 1 || pairs > P_MAX_PAIRS) {
    set_error("f2g_gemm_plan: pairs=%d out of range (1..%d)", pairs, P_MAX_PAIRS);
    return F2G_EINVAL;
  }
  static thread_local PGroup g;
  PlanOut po = {pairs, &g, 0};
  g_plan_out = &po;
  const int rc = gemm_pair_group(problems, n_problems, nullptr);
  g_plan_out = nullptr;
  if (rc) return rc;
  const int need = 4 + 12 * g.n_problems + po.used_pairs + 1 + (g.use_sched ? g.total_tiles : 0);
  if (out_ints < need) {
    set_error("f2g_gemm_plan: out needs %d ints", need);
    return F2G_EINVAL;
  }
  int* o = out;
  *o++ = g.use_sched; *o++ = g.n_problems; *o++ = g.total_tiles; *o++ = po.used_pairs;
  for (int i = 0; i < g.n_problems; ++i) {
    const PProblem& p = g.p[i];
    *o++ = p.M; *o++ = p.N; *o++ = p.K; *o++ = p.bn; *o++ = p.m_tiles; *o++ = p.n_tiles; *o++ = p.tile_begin;
    *o++ = p.split_k; *o++ = p.wait ? 1 : 0; *o++ = p.wait_count; *o++ = p.done ? 1 : 0; *o++ = p.tma_c;
  }
  for (int i = 0; i <= po.used_pairs; ++i) *o++ = g.use_sched ? (int)g.pair_off[i] : 0;
  if (g.use_sched)
    for (int i = 0; i < g.total_tiles; ++i) *o++ = (int)g.sched[i];
  return need;
}
