// Grouped, persistent TF32 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )          fp32 in HBM, TF32 operands, fp32 accumulate
//
// This is the contraction behind every 1x1 conv / Linear on the Flow2GAN hot path
// (reference: nn.Conv1d(k=1) / nn.Linear call sites flow2gan/models/modules.py:443-451,
// 563-593) and, through strided "overlapping-row" tensor maps, the (5,1)/(3,9) Conv2d stacks
// of the discriminators (flow2gan/models/discriminators.py:65-76,171-184).
//
// Structure (one CTA per SM, persistent over a tile list that can span several problems):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : MMA issuer     (one elected thread, tcgen05.mma kind::tf32, D in TMEM)
//   warps 2-5: epilogue       (tcgen05.ld -> regs -> smem transpose -> fused epilogue -> HBM)
// TMEM holds two accumulator buffers so the epilogue of tile i overlaps the main loop of
// tile i+1.  Either operand may be K-major or MN-major (needed by dgrad / wgrad).
#include "common.cuh"
#include "gemm_tf32.h"

#include <stdio.h>
#include <stdlib.h>

namespace f2g {

constexpr int BM = 128;
constexpr int BK = 32;  // 32 fp32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;
constexpr int GEMM_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int EPI_SCRATCH_BYTES = 4 * 32 * 36 * 4;
constexpr int EPI_PARAM_BYTES = 3 * 256 * 4;   // bias / slope / residual scale of one tile
constexpr int SMEM_LIMIT = 227 * 1024;

// A pipeline stage holds KATOMS 32-wide K slices (one 128B-swizzle atom each).  With BN <= 128 the
// four MMAs of one slice (256 tensor cycles) are shorter than the issue + barrier round trip
// (~375 cycles measured), so two slices share a stage / barrier; BN = 256 keeps one (smem).
template <int BN>
struct TileCfg {
  static constexpr int KATOMS = BN <= 128 ? 2 : 1;
  static constexpr int BK_STAGE = BK * KATOMS;
  static constexpr int B_TILE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = KATOMS * (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int STAGES = (SMEM_LIMIT - EPI_SCRATCH_BYTES - EPI_PARAM_BYTES - 2048) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_SCRATCH_BYTES + EPI_PARAM_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
};

struct alignas(64) DevProblem {
  CUtensorMap map_a;
  CUtensorMap map_b;
  float* c;
  float* c_pre;
  const float* bias;
  const float* slope;
  const float* res;
  const float* res_scale;
  const float* row_scale;
  const float* gate;
  int ldc, ld_res, ld_gate, ld_pre;
  int M, N, K;
  int m_tiles, n_tiles, tile_begin;
  int split_k, kb_total, kb_per;
  int act, round_tf32, accumulate;
  float leaky, alpha;
};

struct alignas(64) DevGroup {
  DevProblem p[F2G_GEMM_MAX_PROBLEMS];
  int n_problems;
  int total_tiles;
  // MN-major operand descriptor geometry (bytes); defaults in gemm_tf32_group(), overridable
  // through F2G_MN_* environment variables for bring-up experiments.
  int mn_lbo, mn_sbo, mn_layout, mn_kstep;
  int dbg;  // bring-up only: bit0 = skip MMA issue, bit1 = skip TMA loads, bit2 = skip epilogue stores
};

struct TileCoord {
  int prob, m0, n0, kb0, kb1;
};

F2G_DEVINL TileCoord decode_tile(const DevGroup& g, int tile) {
  int pi = 0;
#pragma unroll 1
  for (int i = 1; i < g.n_problems; ++i)
    if (tile >= g.p[i].tile_begin) pi = i;
  int local = tile - g.p[pi].tile_begin;
  const int nt = g.p[pi].n_tiles;
  const int mn = nt * g.p[pi].m_tiles;
  const int ks = local / mn;          // K-split index (0 when split_k == 1)
  local -= ks * mn;
  TileCoord t;
  t.prob = pi;
  t.m0 = (local / nt) * BM;
  t.n0 = (local % nt);
  t.kb0 = ks * g.p[pi].kb_per;
  t.kb1 = min(t.kb0 + g.p[pi].kb_per, g.p[pi].kb_total);
  return t;
}

enum { EPI_GENERIC = 0, EPI_BIAS_ACT = 1, EPI_BIAS_RES = 2, EPI_PLAIN = 3 };

template <int BN, int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ DevGroup g, const DevGroup* __restrict__ gmaps) {
  using Cfg = TileCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;

  // 128B-swizzled operand tiles need 1024 B alignment; keep every access in the shared state
  // space (a uintptr_t round-up would turn them into generic LD/ST -- measured 10x slower).
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("f2g gemm: dynamic smem base not 1024B aligned\n");
    __trap();
  }

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < g.n_problems; ++i) {
      tma_prefetch_desc(&g.p[i].map_a);
      tma_prefetch_desc(&g.p[i].map_b);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(g, tile);
        const DevProblem& pr = g.p[tc.prob];
        const int n0 = tc.n0 * BN;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa0 = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb0 = sa0 + Cfg::KATOMS * A_TILE_BYTES;
          if (g.dbg & 2) {
            mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const CUtensorMap* ma = gmaps ? &gmaps->p[tc.prob].map_a : &pr.map_a;
          const CUtensorMap* mb = gmaps ? &gmaps->p[tc.prob].map_b : &pr.map_b;
#pragma unroll
          for (int at = 0; at < Cfg::KATOMS; ++at) {
            uint8_t* sa = sa0 + at * A_TILE_BYTES;
            uint8_t* sb = sb0 + at * Cfg::B_TILE_BYTES;
            const int kc = kb * Cfg::BK_STAGE + at * BK;      // beyond K: TMA zero-fills the slice
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 32; ++j)
                tma_load_2d(sa + j * 4096, ma, &full_bar[stage], tc.m0 + 32 * j, kc);
            } else {
              tma_load_2d(sa, ma, &full_bar[stage], kc, tc.m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 32; ++j)
                tma_load_2d(sb + j * 4096, mb, &full_bar[stage], n0 + 32 * j, kc);
            } else {
              tma_load_2d(sb, mb, &full_bar[stage], kc, n0);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(BM, BN, A_MN, B_MN);
      const uint32_t a_addr0 = smem_u32(smem), b_addr0 = a_addr0 + Cfg::KATOMS * A_TILE_BYTES;
      const uint64_t adesc0 = A_MN ? make_smem_desc(a_addr0, g.mn_lbo, g.mn_sbo, g.mn_layout)
                                   : make_smem_desc_sw128(a_addr0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(b_addr0, g.mn_lbo, g.mn_sbo, g.mn_layout)
                                   : make_smem_desc_sw128(b_addr0, 16, 1024);
      const int a_kstep = A_MN ? g.mn_kstep : 32, b_kstep = B_MN ? g.mn_kstep : 32;
      const bool dbg_no_mma = (g.dbg & 1) != 0;
      int stage = 0;
      uint32_t phase = 0;
      int ab = 0;
      uint32_t ab_phase = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(g, tile);
        mbar_wait(&tmem_empty_bar[ab], ab_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + ab * BN;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // Descriptors differ only in the 14-bit (address >> 4) field: add the stage / k-step
          // byte offset (>> 4) to the stage-0 descriptor (no carry: smem addresses < 2^18).
          // K-major (SWIZZLE_128B): the 4 k-steps live inside one 128B swizzle row -> +32 B.
          // MN-major (SWIZZLE_128B_BASE32B): a k-step is 8 rows = two 4-row (512 B) swizzle
          // atoms -> +1024 B; 32-wide MN blocks (one TMA box each) are 4096 B apart.
          const uint32_t soff = (uint32_t)(stage * Cfg::STAGE_BYTES) >> 4;
#pragma unroll
          for (int at = 0; at < Cfg::KATOMS; ++at) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = adesc0 + soff + (uint32_t)((at * A_TILE_BYTES + k * a_kstep) >> 4);
              const uint64_t bdesc = bdesc0 + soff + (uint32_t)((at * Cfg::B_TILE_BYTES + k * b_kstep) >> 4);
              if (!dbg_no_mma) umma_tf32(tmem_d, adesc, bdesc, idesc, (kb > tc.kb0 || at || k) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[ab]);  // accumulator complete -> epilogue
        ab ^= 1;
        if (ab == 0) ab_phase ^= 1;
      }
    }
  } else {
    // ------------------------------- epilogue warps -----------------------------------
    // TMEM lane = output row.  Each warp drains its 32 rows in 32-column chunks: tcgen05.ld ->
    // 8 x STS.128 into a padded (stride 36) smem scratch -> re-read as 4 rows x 8 column-quads
    // per pass so that every LDG/STG is a 128-bit access and a warp instruction covers four full
    // 128 B row segments.  Per-column parameters of the tile are staged in smem while the main
    // loop still runs.  The common epilogues are compile-time specialisations (EPI): with ONE
    // epilogue warp per SM sub-partition there is no thread-level parallelism to hide a long
    // branchy dependent chain, so the fast paths are branch-free and issue their 8 LDS.128 /
    // LDG.128 up front (ncu showed the generic path latency-bound at ~500 cycles per row group).
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    float* const scratch = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES) + (warp - 2) * (32 * 36);
    float* const sparam = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + EPI_SCRATCH_BYTES);
    const int cg = lane & 7, rsub = lane >> 3;
    const int et = (warp - 2) * 32 + lane;
    int ab = 0;
    uint32_t ab_phase = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(g, tile);
      const DevProblem& pr = g.p[tc.prob];
      // hoist every epilogue parameter out of the constant bank once per tile
      const int n0 = tc.n0 * BN;
      const int row_base = tc.m0 + q * 32;
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
      const bool skip = (g.dbg & 4) != 0;
      // 128-bit global accesses need 16 B aligned rows
      const bool vec_c = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(cbase) & 15) == 0);
      const bool vec_res = !res_p || (((ld_res & 3) == 0) && ((reinterpret_cast<uintptr_t>(res_p) & 15) == 0));
      const bool vec_gate = !gate_p || (((ld_gate & 3) == 0) && ((reinterpret_cast<uintptr_t>(gate_p) & 15) == 0));
      const bool vec_pre = !pre_p || (((ld_pre & 3) == 0) && ((reinterpret_cast<uintptr_t>(pre_p) & 15) == 0));
      const bool vec_all = vec_c && vec_res && vec_gate && vec_pre;

      // stage bias / slope / residual-scale of this tile's BN columns (overlaps the main loop)
      {
        const float* const bias_p = pr.bias;
        const float* const slope_p = pr.slope;
        const float* const rsc_p = pr.res_scale;
        const float leaky = pr.leaky;
        for (int c = et; c < BN; c += 128) {
          const int colc = n0 + c;
          const bool okc = colc < N;
          sparam[c] = (bias_p && okc) ? __ldg(bias_p + colc) : 0.f;
          sparam[BN + c] = (slope_p && okc) ? __ldg(slope_p + colc) : leaky;
          sparam[2 * BN + c] = (rsc_p && okc) ? __ldg(rsc_p + colc) : 1.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }

      mbar_wait(&tmem_full_bar[ab], ab_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= N || skip) break;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ab * BN + c0, v);
        tmem_ld_wait();
        if (rows <= 0) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(scratch + lane * 36 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
        __syncwarp();
        const int col = n0 + c0 + 4 * cg;
        const int ncol = min(4, N - col);          // valid columns of this lane's quad (<= 0: none)
        const float4 bias4 = *reinterpret_cast<const float4*>(sparam + c0 + 4 * cg);
        const float4 slope4 = *reinterpret_cast<const float4*>(sparam + BN + c0 + 4 * cg);
        const float4 rsc4 = *reinterpret_cast<const float4*>(sparam + 2 * BN + c0 + 4 * cg);
        const bool chunk_full = vec_all && (n0 + c0 + 32 <= N) && pr.split_k == 1;   // warp-uniform

        if (EPI != EPI_GENERIC && chunk_full) {
          // ---------------- specialised, branch-free fast paths -------------------------
          float4 xv[8], rv[8];
          bool okr[8];
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int i = rr * 4 + rsub;
            okr[rr] = i < rows;
            xv[rr] = *reinterpret_cast<const float4*>(scratch + i * 36 + 4 * cg);
          }
          if (EPI == EPI_BIAS_RES) {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr)
              rv[rr] = (okr[rr] && res_p) ? __ldg(reinterpret_cast<const float4*>(
                                     res_p + (size_t)(row_base + rr * 4 + rsub) * ld_res + col))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (EPI == EPI_PLAIN && do_acc) {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr)
              rv[rr] = okr[rr] ? *reinterpret_cast<const float4*>(
                                     cbase + (size_t)(row_base + rr * 4 + rsub) * ldc + col)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            float4 x = xv[rr];
            x.x = fmaf(x.x, alpha, bias4.x); x.y = fmaf(x.y, alpha, bias4.y);
            x.z = fmaf(x.z, alpha, bias4.z); x.w = fmaf(x.w, alpha, bias4.w);
            const size_t row = (size_t)(row_base + rr * 4 + rsub);
            if (EPI == EPI_BIAS_ACT) {
              if (pre_p && okr[rr]) *reinterpret_cast<float4*>(pre_p + row * ld_pre + col) = x;
              x.x = x.x > 0.f ? x.x : x.x * slope4.x; x.y = x.y > 0.f ? x.y : x.y * slope4.y;
              x.z = x.z > 0.f ? x.z : x.z * slope4.z; x.w = x.w > 0.f ? x.w : x.w * slope4.w;
            } else if (EPI == EPI_BIAS_RES) {
              x.x = fmaf(rsc4.x, rv[rr].x, x.x); x.y = fmaf(rsc4.y, rv[rr].y, x.y);
              x.z = fmaf(rsc4.z, rv[rr].z, x.z); x.w = fmaf(rsc4.w, rv[rr].w, x.w);
            } else if (do_acc) {
              x.x += rv[rr].x; x.y += rv[rr].y; x.z += rv[rr].z; x.w += rv[rr].w;
            }
            if (do_round) {
              x.x = tf32_rna(x.x); x.y = tf32_rna(x.y); x.z = tf32_rna(x.z); x.w = tf32_rna(x.w);
            }
            if (okr[rr]) *reinterpret_cast<float4*>(cbase + row * ldc + col) = x;
          }
          __syncwarp();
          continue;
        }

        const float bias[4] = {bias4.x, bias4.y, bias4.z, bias4.w};
        const float slope[4] = {slope4.x, slope4.y, slope4.z, slope4.w};
        const float rsc[4] = {rsc4.x, rsc4.y, rsc4.z, rsc4.w};
        if (pr.split_k > 1) {   // partial-K tile: atomically accumulate into the pre-zeroed C
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
        // generic path (any epilogue combination, N tails, unaligned leading dimensions); kept
        // compact on purpose: fully unrolled it was instruction-fetch bound (ncu: stall_no_inst)
        if (quad) {
#pragma unroll 2
          for (int rr = 0; rr < 8; ++rr) {
            const int i = rr * 4 + rsub;
            if (i >= rows) continue;
            const int row = row_base + i;
            const float4 xv = *reinterpret_cast<const float4*>(scratch + i * 36 + 4 * cg);
            float x[4] = {fmaf(xv.x, alpha, bias[0]), fmaf(xv.y, alpha, bias[1]), fmaf(xv.z, alpha, bias[2]),
                          fmaf(xv.w, alpha, bias[3])};
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
            float* dst = cbase + (size_t)row * ldc + col;
            if (do_acc) {
              const float4 t = *reinterpret_cast<const float4*>(dst);
              x[0] += t.x; x[1] += t.y; x[2] += t.z; x[3] += t.w;
            }
            if (do_round) {
              x[0] = tf32_rna(x[0]); x[1] = tf32_rna(x[1]); x[2] = tf32_rna(x[2]); x[3] = tf32_rna(x[3]);
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
              float* dst = cbase + (size_t)row * ldc + col + e;
              if (do_acc) x += *dst;
              *dst = do_round ? tf32_rna(x) : x;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[ab]);
      asm volatile("bar.sync 1, 128;" ::: "memory");   // all warps done with this tile's staged parameters
      ab ^= 1;
      if (ab == 0) ab_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%d)", (int)e);
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D fp32 tensor map, 128B swizzle.  dims/strides innermost first; box = {32, box_rows}.
static int env_int(const char* name, int dflt) { return bringup_int(name, dflt); }   // -DF2G_BRINGUP builds only

static int encode_2d(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer,
                     uint64_t outer_stride_elems, uint32_t box_rows, bool mn_major) {
  EncodeTiledFn fn = get_encode_fn();
  static const int l2promo = env_int("F2G_TMA_L2PROMO", (int)CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  static const int mn_swz = env_int("F2G_MN_TMA_SWIZZLE", (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (!fn) return F2G_EDRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (outer_stride_elems % 4) != 0) {
    set_error("gemm operand must be 16B aligned with a leading dimension multiple of 4 floats "
              "(ptr=%p ld=%llu)", (const void*)base, (unsigned long long)outer_stride_elems);
    return F2G_EINVAL;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {outer_stride_elems * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? (CUtensorMapSwizzle)mn_swz : CU_TENSOR_MAP_SWIZZLE_128B,
                  (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", (int)r,
              (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)outer_stride_elems);
    return F2G_EDRIVER;
  }
  return 0;
}

template <int BN, int A_MN, int B_MN, int EPI>
static int launch(const DevGroup& g, int num_sms, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  static bool configured = false;
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN, EPI>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int grid = g.total_tiles < num_sms ? g.total_tiles : num_sms;
  const DevGroup* gmaps = nullptr;
  static const int desc_global = env_int("F2G_DESC_GLOBAL", 0);
  if (desc_global) {   // bring-up experiment: tensor maps fetched from global instead of param space
    static DevGroup* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, sizeof(DevGroup));
    cudaMemcpyAsync(dbuf, &g, sizeof(DevGroup), cudaMemcpyHostToDevice, stream);
    gmaps = dbuf;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(g, gmaps);
  return check_launch("gemm_tf32");
}

static int g_num_sms = 0;

int gemm_tf32_group(const F2GGemm* descs, int n, cudaStream_t stream) {
  // default: the CTA-pair kernel; F2G_GEMM_V1=1 keeps this one-CTA kernel for A/B comparisons
  static const int use_v1 = env_int("F2G_GEMM_V1", 0);
  // (problems with at most 128 rows -- e.g. weight gradients of the 32-channel MRD convs --
  // would leave half of every 256-row pair tile empty)
  int max_m = 0;
  for (int i = 0; i < n && i < F2G_GEMM_MAX_PROBLEMS; ++i) max_m = descs[i].M > max_m ? descs[i].M : max_m;
  bool windowed = false;
  for (int i = 0; i < n && i < F2G_GEMM_MAX_PROBLEMS; ++i) windowed |= descs[i].a_seg_len != 0 || descs[i].ab_f16 != 0;
  if (windowed || (!use_v1 && max_m > 128)) return gemm_pair_group(descs, n, stream);
  if (n < 1 || n > F2G_GEMM_MAX_PROBLEMS) {
    set_error("gemm group size %d out of range", n);
    return F2G_EINVAL;
  }
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  DevGroup g;
  memset(&g, 0, sizeof(g));
  const int bn = descs[0].bn, a_mn = descs[0].a_mn, b_mn = descs[0].b_mn;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    const F2GGemm& d = descs[i];
    if (d.bn != bn || d.a_mn != a_mn || d.b_mn != b_mn) {
      set_error("all problems of a gemm group must share bn / operand majors");
      return F2G_EINVAL;
    }
    if (d.M <= 0 || d.N <= 0 || d.K <= 0) {
      set_error("gemm problem %d has empty shape %dx%dx%d", i, d.M, d.N, d.K);
      return F2G_EINVAL;
    }
    DevProblem& p = g.p[i];
    int rc;
    rc = a_mn ? encode_2d(&p.map_a, d.a, d.M, d.K, d.lda, 32, true)
              : encode_2d(&p.map_a, d.a, d.K, d.M, d.lda, BM, false);
    if (rc) return rc;
    rc = b_mn ? encode_2d(&p.map_b, d.b, d.N, d.K, d.ldb, 32, true)
              : encode_2d(&p.map_b, d.b, d.K, d.N, d.ldb, bn, false);
    if (rc) return rc;
    p.c = d.c; p.ldc = d.ldc; p.c_pre = d.c_pre; p.ld_pre = d.ld_pre;
    p.bias = d.bias; p.slope = d.slope; p.res = d.res; p.res_scale = d.res_scale;
    p.row_scale = d.row_scale; p.gate = d.gate;
    p.ld_res = d.ld_res; p.ld_gate = d.ld_gate;
    p.M = d.M; p.N = d.N; p.K = d.K;
    p.m_tiles = (d.M + BM - 1) / BM;
    p.n_tiles = (d.N + bn - 1) / bn;
    p.tile_begin = tiles;
    p.act = d.act; p.round_tf32 = d.round_tf32; p.accumulate = d.accumulate;
    p.leaky = d.leaky; p.alpha = d.alpha == 0.f ? 1.f : d.alpha;
    const int bk_stage = bn <= 128 ? 2 * BK : BK;
    p.kb_total = (d.K + bk_stage - 1) / bk_stage;
    int sk = d.split_k < 1 ? 1 : d.split_k;
    if (sk > p.kb_total) sk = p.kb_total;
    p.kb_per = (p.kb_total + sk - 1) / sk;
    p.split_k = (p.kb_total + p.kb_per - 1) / p.kb_per;      // every split owns >= 1 k-block
    if (p.split_k > 1 && (d.bias || d.act || d.res || d.gate || d.row_scale || d.c_pre || d.round_tf32)) {
      set_error("split-K gemm supports the plain (alpha) epilogue only");
      return F2G_EINVAL;
    }
    tiles += p.m_tiles * p.n_tiles * p.split_k;
  }
  g.n_problems = n;
  g.total_tiles = tiles;
  static const int mn_lbo = env_int("F2G_MN_LBO", 4096), mn_sbo = env_int("F2G_MN_SBO", 512),
                   mn_layout = env_int("F2G_MN_LAYOUT", 1), mn_kstep = env_int("F2G_MN_KSTEP", 1024);
  g.mn_lbo = mn_lbo; g.mn_sbo = mn_sbo; g.mn_layout = mn_layout; g.mn_kstep = mn_kstep;
  g.dbg = env_int("F2G_GEMM_DBG", 0);

  // epilogue specialisation shared by the whole group (else the generic epilogue)
  int epi = -1;
  for (int i = 0; i < n; ++i) {
    const F2GGemm& d = descs[i];
    int e = EPI_GENERIC;
    const bool odd = d.gate || d.row_scale || d.act == F2G_ACT_SILU;
    if (!odd && d.bias && (d.act == F2G_ACT_PRELU || d.act == F2G_ACT_LEAKY) && !d.res && !d.accumulate)
      e = EPI_BIAS_ACT;
    else if (!odd && d.act == F2G_ACT_NONE && (d.res || d.bias) && !d.accumulate && !d.c_pre)
      e = EPI_BIAS_RES;            // bias [+ scaled residual]
    else if (!odd && d.act == F2G_ACT_NONE && !d.res && !d.bias && !d.c_pre)
      e = EPI_PLAIN;
    epi = (epi == -1 || epi == e) ? e : EPI_GENERIC;
  }
  static const int force_generic = env_int("F2G_GEMM_GENERIC_EPI", 0);
  if (force_generic) epi = EPI_GENERIC;

#define F2G_DISPATCH_E(BN_, AM_, BM_)                                                       \
  {                                                                                         \
    if (epi == EPI_BIAS_ACT) return launch<BN_, AM_, BM_, EPI_BIAS_ACT>(g, g_num_sms, stream); \
    if (epi == EPI_BIAS_RES) return launch<BN_, AM_, BM_, EPI_BIAS_RES>(g, g_num_sms, stream); \
    if (epi == EPI_PLAIN) return launch<BN_, AM_, BM_, EPI_PLAIN>(g, g_num_sms, stream);       \
    return launch<BN_, AM_, BM_, EPI_GENERIC>(g, g_num_sms, stream);                         \
  }
#define F2G_DISPATCH(BN_)                                  \
  if (bn == BN_) {                                         \
    if (!a_mn && !b_mn) F2G_DISPATCH_E(BN_, 0, 0)          \
    if (!a_mn && b_mn) F2G_DISPATCH_E(BN_, 0, 1)           \
    if (a_mn && b_mn) F2G_DISPATCH_E(BN_, 1, 1)            \
    if (a_mn && !b_mn) F2G_DISPATCH_E(BN_, 1, 0)           \
  }
  F2G_DISPATCH(64)
  F2G_DISPATCH(128)
  F2G_DISPATCH(256)
#undef F2G_DISPATCH
#undef F2G_DISPATCH_E
  set_error("unsupported gemm tile width bn=%d (64/128/256)", bn);
  return F2G_EINVAL;
}

}  // namespace f2g
