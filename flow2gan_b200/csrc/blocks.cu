// ConvNeXt-block SIMT pieces on channel-last rows: BiasNorm, fused depthwise-conv prologue,
// small dense layers, sinusoidal time embedding, packing helpers.  All HBM/L2-bound.
// Reference: flow2gan/models/modules.py:217-232 (SinusoidalPosEmb), :286-416 (BiasNorm),
// :456-495 (ConvNeXtBlock.forward), :523-542 (CondEncoder), :595-627 (ConvNeXtDecoder).
#include "simt.cuh"
#include "../../include/flow2gan_b200.h"

#include <string.h>
#include <stdlib.h>

namespace f2g {

constexpr int MAX_CHUNKS = 8;  // channels <= 8 * 128 = 1024 per row

F2G_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
F2G_DEVINL void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// One warp per row.  y = x * (mean_c (x - bias_c)^2)^-1/2 * exp(log_scale).  No epsilon
// (modules.py:309-312).
__global__ void biasnorm_kernel(const float* __restrict__ x, int rows, int C, int ld,
                                const float* __restrict__ bias,
                                const float* __restrict__ log_scale, float* __restrict__ y,
                                int ld_y, float* __restrict__ inv_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int chunks = C >> 7;
  float4 v[MAX_CHUNKS];
  float ssq = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_CHUNKS; ++j) {
    if (j < chunks) {
      const int c = j * 128 + lane * 4;
      v[j] = ld4(x + (size_t)row * ld + c);
      const float4 b = ld4(bias + c);
      const float d0 = v[j].x - b.x, d1 = v[j].y - b.y, d2 = v[j].z - b.z, d3 = v[j].w - b.w;
      ssq += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
  ssq = warp_sum(ssq);
  const float scale = (1.0f / sqrtf(ssq / (float)C)) * expf(*log_scale);
  if (inv_out && lane == 0) inv_out[row] = scale;
#pragma unroll
  for (int j = 0; j < MAX_CHUNKS; ++j) {
    if (j < chunks) {
      const int c = j * 128 + lane * 4;
      st4(y + (size_t)row * ld_y + c,
          make_float4(v[j].x * scale, v[j].y * scale, v[j].z * scale, v[j].w * scale));
    }
  }
}

// Fused ConvNeXt-block prologue, up to 4 independent problems (the three branches' same-depth
// blocks) per launch.
//   y = dwconv7(x * mask) + b ; z = BiasNorm(y) + cond_row ; z *= 1 + ts[b] ; out = tf32(z)
// One CTA = 384 threads = TL token lanes x C/4 channel quads; every lane slides over PRE_S
// consecutive tokens: it loads PRE_S+6 input rows ONCE (all loads in flight together) and the 7
// depthwise taps once, instead of 7 rows + 7 taps per token -- the per-token version was bound by
// L1 wavefronts (18 LDG.128 per token quad; 33 us for the three branches of one block).
constexpr int PRE_THREADS = 384;
constexpr int PRE_MAX_TL = 4;       // C >= 384

struct BlockPreArgs {
  F2GBlockPre p[4];
  int cta_begin[5];     // prefix sums of CTAs per problem
  int tl[4];            // token lanes per CTA
  int ctas_t[4];        // CTAs along T per batch element
  int fshift[4];        // log2(factor) if factor is a power of two, else -1
  float inv_ctas_t[4], inv_c4[4];
  int n;
};

template <int PRE_S, int MINB>
__global__ void __launch_bounds__(PRE_THREADS, MINB) block_pre_kernel(const __grid_constant__ BlockPreArgs a) {
  __shared__ float red[PRE_MAX_TL * PRE_S][8];
  __shared__ float4 stage[PRE_S + 2][PRE_THREADS];   // per-thread slots: cond rows, time scale, norm bias
  int pi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < a.n && (int)blockIdx.x >= a.cta_begin[i]) pi = i;
  const F2GBlockPre& P = a.p[pi];
  const int C = P.C, T = P.T;
  const int c4 = C >> 2;
  const int TL = a.tl[pi];
  const int local = blockIdx.x - a.cta_begin[pi];
  // two small exact integer divisions done in fp32 (operands < 2^20)
  const int bi = __float2int_rd(((float)local + 0.5f) * a.inv_ctas_t[pi]);
  const int ty = __float2int_rd(((float)threadIdx.x + 0.5f) * a.inv_c4[pi]);
  const int tx = threadIdx.x - ty * c4;
  const int t0 = ((local - bi * a.ctas_t[pi]) * TL + ty) * PRE_S;
  const bool live = ty < TL && t0 < T;
  const int c = tx * 4;
  const int wid = tx >> 5, nw = c4 >> 5;
  const size_t rb = (size_t)bi * T;
  const float* __restrict__ x = P.x;
  const float* __restrict__ row_mask = P.row_mask;
  const int ld_x = P.ld_x;
  pdl_wait();
  pdl_launch();
  if (blockIdx.x == 0 && a.p[0].zero_ptr)      // chaining counters of the GEMM group that follows
    for (int i = threadIdx.x; i < a.p[0].zero_n; i += PRE_THREADS) a.p[0].zero_ptr[i] = 0;

  float4 acc[PRE_S];
  float4 cv[PRE_S];
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
  float ssq[PRE_S];
  float log_scale = 0.f;
#pragma unroll
  for (int s = 0; s < PRE_S; ++s) {
    acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    cv[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    ssq[s] = 0.f;
  }
  if (live) {
    float4 xv[PRE_S + 6];
    if (!row_mask && t0 >= 3 && t0 + PRE_S + 3 <= T) {
      // interior window without a length mask (every CTA but the two at the sequence ends of an
      // unmasked batch): one base pointer, no per-row bounds or mask tests
      const float* xp = x + (rb + (size_t)(t0 - 3)) * ld_x + c;
#pragma unroll
      for (int j = 0; j < PRE_S + 6; ++j) xv[j] = ld4(xp + (size_t)j * ld_x);
    } else {
#pragma unroll
      for (int j = 0; j < PRE_S + 6; ++j) {
        const int tt = t0 - 3 + j;
        xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tt >= 0 && tt < T) {
          const float mk = row_mask ? row_mask[rb + tt] : 1.f;
          if (mk != 0.f) {
            const float4 v = ld4(x + (rb + tt) * ld_x + c);
            xv[j] = make_float4(v.x * mk, v.y * mk, v.z * mk, v.w * mk);
          }
        }
      }
    }
    // everything the epilogue of this thread needs, as LDGSTS into its own shared-memory slots: in
    // flight together with the window, no registers held, no second L2 round trip after the taps
    if (P.cond) {
      // conditioning row of token t: frame t / factor (factor is 1, 2 or 4 in every released
      // config: shift), or the all-zero row past the last mel frame
      const int fsh = a.fshift[pi];
      const int t_cond = P.cond_T * P.factor;
      const float* cbase = P.cond + c;
#pragma unroll
      for (int s = 0; s < PRE_S; ++s) {
        const int t = t0 + s;
        if (t < T) {
          const int fr = fsh >= 0 ? (t >> fsh) : t / P.factor;
          const int crow = t < t_cond ? bi * P.cond_T + fr : P.zero_row;
          cp_async16(&stage[s][threadIdx.x], cbase + (size_t)crow * P.ld_cond);
        }
      }
    }
    if (P.tscale) cp_async16(&stage[PRE_S][threadIdx.x], P.tscale + (size_t)bi * P.ld_ts + c);
    cp_async16(&stage[PRE_S + 1][threadIdx.x], P.bn_bias + c);
    log_scale = *P.bn_log_scale;
    const float4 b0 = ld4(P.dw_b + c);
#pragma unroll
    for (int s = 0; s < PRE_S; ++s) acc[s] = b0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float4 w = ld4(P.dw_wT + k * C + c);
#pragma unroll
      for (int s = 0; s < PRE_S; ++s) {
        acc[s].x = fmaf(xv[s + k].x, w.x, acc[s].x);
        acc[s].y = fmaf(xv[s + k].y, w.y, acc[s].y);
        acc[s].z = fmaf(xv[s + k].z, w.z, acc[s].z);
        acc[s].w = fmaf(xv[s + k].w, w.w, acc[s].w);
      }
    }
    // conditioning rows, time scale and norm bias were requested before the window (below): they
    // arrived under the window's round trip
    cp_async_wait_all();
    if (P.cond) {
#pragma unroll
      for (int s = 0; s < PRE_S; ++s)
        if (t0 + s < T) cv[s] = stage[s][threadIdx.x];
    }
    if (P.tscale) sc = stage[PRE_S][threadIdx.x];
    const float4 bb = stage[PRE_S + 1][threadIdx.x];
#pragma unroll
    for (int s = 0; s < PRE_S; ++s) {
      const float d0 = acc[s].x - bb.x, d1 = acc[s].y - bb.y, d2 = acc[s].z - bb.z, d3 = acc[s].w - bb.w;
      ssq[s] = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      if (P.conv_out && t0 + s < T) st4(P.conv_out + (rb + t0 + s) * C + c, acc[s]);
    }
  }
  // c4 is a multiple of 32, so each warp belongs to exactly one token lane
  if (PRE_S == 4) {
    // the four tokens' sums in ONE butterfly (6 shuffles instead of 20): fold tokens {2,3} onto the
    // upper half-warp, then {1} / {3} onto the odd quarter-warps, then reduce within 8 lanes;
    // lane 8k ends up with the warp total of token k
    const int lane = threadIdx.x & 31;
    const bool hi = (lane & 16) != 0;
    const float k0 = hi ? ssq[2] : ssq[0], k1 = hi ? ssq[3] : ssq[1];
    const float s0 = hi ? ssq[0] : ssq[2], s1 = hi ? ssq[1] : ssq[3];
    const float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
    const float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool b8 = (lane & 8) != 0;
    float q = (b8 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, b8 ? r0 : r1, 8);
    q += __shfl_xor_sync(0xffffffffu, q, 4);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    if ((lane & 7) == 0 && ty < PRE_MAX_TL) red[ty * PRE_S + (lane >> 3)][wid] = q;
  } else {
#pragma unroll
    for (int s = 0; s < PRE_S; ++s) {
      const float r = warp_sum(ssq[s]);
      if ((threadIdx.x & 31) == 0 && ty < PRE_MAX_TL) red[ty * PRE_S + s][wid] = r;
    }
  }
  __syncthreads();
  if (!live) return;
  const float gain = expf(log_scale);
  const float inv_c = 1.0f / (float)C;
  uint32_t sat_acc = 0;  // fp16 range guard: running per-half maximum of the converted |bits| (common.cuh)
#pragma unroll
  for (int s = 0; s < PRE_S; ++s) {
    const int t = t0 + s;
    if (t >= T) break;
    float tot = 0.f;
    for (int j = 0; j < nw; ++j) tot += red[ty * PRE_S + s][j];
    const float inv = rsqrtf(tot * inv_c) * gain;      // 2 ulp; the result is rounded to 11 bits below
    if (P.inv_rms_out && tx == 0) P.inv_rms_out[rb + t] = inv;
    float4 z = make_float4(acc[s].x * inv, acc[s].y * inv, acc[s].z * inv, acc[s].w * inv);
    z.x += cv[s].x; z.y += cv[s].y; z.z += cv[s].z; z.w += cv[s].w;
    z.x *= 1.f + sc.x; z.y *= 1.f + sc.y; z.z *= 1.f + sc.z; z.w *= 1.f + sc.w;
    if (P.out_f16) {   // operand of a kind::f16 GEMM: same 11-bit significand as the TF32 rounding
      __half* o = reinterpret_cast<__half*>(P.out) + (rb + t) * (size_t)P.ld_out + c;
      const uint2 hw = pack_half4(z);
      *reinterpret_cast<uint2*>(o) = hw;
      sat_acc = half2_track(half2_track(sat_acc, hw.x), hw.y);
    } else {
      st4(P.out + (rb + t) * P.ld_out + c, make_float4(tf32_rna(z.x), tf32_rna(z.y), tf32_rna(z.z), tf32_rna(z.w)));
    }
  }
  if (P.sat_flag && half2_out_of_range(sat_acc)) atomicOr(P.sat_flag, 2);
}

// Batched small dense layers (up to 4 independent problems per launch, blockIdx.y = problem):
// out[b,o] = act(bias[o] + sum_k in[b,k] W[o,k]).  One warp per output column o: the W row is
// held in registers (float4 per lane per 128 columns), every batch row is then one float4 dot +
// warp reduction, so the weight matrix is streamed exactly once.
struct LinSmallArgs {
  const float* in[4];
  const float* W[4];
  const float* bias[4];
  float* out[4];
  int K[4], O[4], ld_in[4], ldw[4], ld_out[4];
  int n, B, act;
};
constexpr int LIN_MAX_K4 = 12;  // K <= 12 * 128 = 1536

__global__ void linear_small_kernel(const LinSmallArgs a) {
  const int p = blockIdx.y;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= a.O[p]) return;
  const int K = a.K[p];
  const int chunks = K >> 7;
  const float* __restrict__ w = a.W[p] + (size_t)o * a.ldw[p];
  const float* __restrict__ in = a.in[p];
  const int ld_in = a.ld_in[p];
  float4 wr[LIN_MAX_K4];
#pragma unroll
  for (int j = 0; j < LIN_MAX_K4; ++j)
    if (j < chunks) wr[j] = __ldg(reinterpret_cast<const float4*>(w + j * 128 + lane * 4));
  const float bias = a.bias[p] ? a.bias[p][o] : 0.f;
  for (int b = 0; b < a.B; ++b) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < LIN_MAX_K4; ++j) {
      if (j < chunks) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(in + (size_t)b * ld_in + j * 128 + lane * 4));
        acc = fmaf(x.x, wr[j].x, acc);
        acc = fmaf(x.y, wr[j].y, acc);
        acc = fmaf(x.z, wr[j].z, acc);
        acc = fmaf(x.w, wr[j].w, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + bias;
      if (a.act == F2G_ACT_SILU) v = v / (1.f + expf(-v));
      a.out[p][(size_t)b * a.ld_out[p] + o] = v;
    }
  }
}

__global__ void time_sinusoid_kernel(const float* __restrict__ t, int B, int half,
                                     const float* __restrict__ freqs, float scale,
                                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float a = scale * t[b] * freqs[k];
  out[(size_t)b * 2 * half + k] = sinf(a);
  out[(size_t)b * 2 * half + half + k] = cosf(a);
}

__global__ void pack2d_kernel(const float* __restrict__ src, long long src_rs, long long src_cs,
                              int rows, int cols, float* __restrict__ dst, int ld, int ld_fill,
                              int round_tf32) {
  const long long total = (long long)rows * ld_fill;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_fill), c = (int)(i - (long long)r * ld_fill);
    float v = 0.f;
    if (c < cols) v = src[r * src_rs + c * src_cs];
    dst[(size_t)r * ld + c] = round_tf32 ? tf32_rna(v) : v;
  }
}

__global__ void im2col_cf_kernel(const float* __restrict__ x, int B, int C, int T, int ktaps,
                                 float* __restrict__ out, int ld, int round_tf32) {
  const long long total = (long long)B * T * ld;
  const int pad = ktaps / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % ld);
    const long long row = i / ld;
    const int t = (int)(row % T), b = (int)(row / T);
    float v = 0.f;
    if (col < ktaps * C) {
      const int k = col / C, c = col - k * C;
      const int tt = t + k - pad;
      if (tt >= 0 && tt < T) v = x[((size_t)b * C + c) * T + tt];
    }
    out[i] = round_tf32 ? tf32_rna(v) : v;
  }
}

__global__ void frame_mask_kernel(const int* __restrict__ lens, int B, int frames, int hop,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * frames) return;
  const int b = i / frames, f = i - b * frames;
  out[i] = f < 1 + lens[b] / hop ? 1.f : 0.f;
}

}  // namespace f2g

using namespace f2g;

static int check_channels(const char* who, int C, int ld) {
  if (C % 128 != 0 || C > MAX_CHUNKS * 128 || ld % 4 != 0) {
    set_error("%s: channels=%d must be a multiple of 128 (<= %d) and ld=%d a multiple of 4", who, C,
              MAX_CHUNKS * 128, ld);
    return F2G_EINVAL;
  }
  return 0;
}

extern "C" int f2g_biasnorm(const float* x, int rows, int C, int ld, const float* bias,
                            const float* log_scale, float* y, int ld_y, float* inv_out, void* stream) {
  if (int rc = check_channels("f2g_biasnorm", C, ld | ld_y)) return rc;
  const int wpb = 4;
  F2G_LAUNCH_COOP(biasnorm_kernel, (rows + wpb - 1) / wpb, wpb * 32, static_cast<cudaStream_t>(stream), x, rows, C,
                  ld, bias, log_scale, y, ld_y, inv_out);
  return check_launch("f2g_biasnorm");
}

extern "C" int f2g_block_pre_group(const F2GBlockPre* probs, int n, void* stream) {
  if (n < 1 || n > 4) {
    set_error("f2g_block_pre_group: n=%d out of range (1..4)", n);
    return F2G_EINVAL;
  }
  BlockPreArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n;
  // sliding-window length per lane 4, two resident CTAs per SM (round-1 sweep: <2,3>, <2,2>, <8,1> lost)
  const int S = 4;
  int ctas = 0;
  for (int i = 0; i < n; ++i) {
    F2GBlockPre p = probs[i];
    if (int rc = check_channels("f2g_block_pre", p.C, p.ld_x | p.ld_out | (p.cond ? p.ld_cond : 0) | (p.tscale ? p.ld_ts : 0)))
      return rc;
    if (p.factor < 1) p.factor = 1;
    a.p[i] = p;
    if (p.C < PRE_THREADS) {
      set_error("f2g_block_pre: channels=%d must be >= %d", p.C, PRE_THREADS);
      return F2G_EINVAL;
    }
    a.fshift[i] = -1;
    for (int sft = 0; sft < 16; ++sft)
      if (p.factor == (1 << sft)) a.fshift[i] = sft;
    int tl = PRE_THREADS / (p.C / 4);
    if (tl > PRE_MAX_TL) tl = PRE_MAX_TL;
    a.tl[i] = tl;
    a.ctas_t[i] = (p.T + tl * S - 1) / (tl * S);
    a.inv_ctas_t[i] = 1.0f / (float)a.ctas_t[i];
    a.inv_c4[i] = 1.0f / (float)(p.C / 4);
    a.cta_begin[i] = ctas;
    ctas += a.ctas_t[i] * p.B;
  }
  for (int i = n; i <= 4; ++i) a.cta_begin[i] = ctas;
  void (*kern)(BlockPreArgs) = block_pre_kernel<4, 2>;
  cudaError_t le = launch_pdl(kern, dim3(ctas), dim3(PRE_THREADS), 0, static_cast<cudaStream_t>(stream), a);
  if (le != cudaSuccess) {
    set_error("f2g_block_pre launch: %s", cudaGetErrorString(le));
    return (int)le;
  }
  return check_launch("f2g_block_pre");
}

extern "C" int f2g_block_pre(const float* x, int B, int T, int C, int ld_x, const float* dw_wT,
                             const float* dw_b, const float* bn_bias, const float* bn_log_scale,
                             const float* row_mask, const float* cond, int ld_cond, int cond_T,
                             int factor, int zero_row, const float* tscale, int ld_ts, float* out,
                             int ld_out, float* conv_out, float* inv_rms_out, void* stream) {
  F2GBlockPre p;
  p.x = x; p.B = B; p.T = T; p.C = C; p.ld_x = ld_x; p.dw_wT = dw_wT; p.dw_b = dw_b;
  p.bn_bias = bn_bias; p.bn_log_scale = bn_log_scale; p.row_mask = row_mask; p.cond = cond;
  p.ld_cond = ld_cond; p.cond_T = cond_T; p.factor = factor; p.zero_row = zero_row;
  p.tscale = tscale; p.ld_ts = ld_ts; p.out = out; p.ld_out = ld_out; p.conv_out = conv_out;
  p.inv_rms_out = inv_rms_out;
  p.out_f16 = 0; p.zero_ptr = nullptr; p.zero_n = 0; p.sat_flag = nullptr;
  return f2g_block_pre_group(&p, 1, stream);
}

extern "C" int f2g_linear_small(const F2GLinear* probs, int n, int B, int act, void* stream) {
  if (n < 1 || n > 4) {
    set_error("f2g_linear_small: n=%d out of range (1..4)", n);
    return F2G_EINVAL;
  }
  LinSmallArgs a;
  a.n = n; a.B = B; a.act = act;
  int max_o = 0;
  for (int i = 0; i < n; ++i) {
    const F2GLinear& p = probs[i];
    if (p.K % 128 != 0 || p.K > LIN_MAX_K4 * 128 || p.ld_in % 4 != 0 || p.ldw % 4 != 0) {
      set_error("f2g_linear_small: K=%d must be a multiple of 128 (<= %d), ld_in/ldw multiples of 4",
                p.K, LIN_MAX_K4 * 128);
      return F2G_EINVAL;
    }
    a.in[i] = p.in; a.W[i] = p.W; a.bias[i] = p.bias; a.out[i] = p.out;
    a.K[i] = p.K; a.O[i] = p.O; a.ld_in[i] = p.ld_in; a.ldw[i] = p.ldw; a.ld_out[i] = p.ld_out;
    if (p.O > max_o) max_o = p.O;
  }
  const int wpb = 4;
  dim3 grid((max_o + wpb - 1) / wpb, n);
  F2G_LAUNCH_COOP(linear_small_kernel, grid, wpb * 32, static_cast<cudaStream_t>(stream), a);
  return check_launch("f2g_linear_small");
}

extern "C" int f2g_time_sinusoid(const float* t, int B, int dim, const float* freqs, float scale,
                                 float* out, void* stream) {
  const int half = dim / 2;
  const int total = B * half;
  F2G_LAUNCH(time_sinusoid_kernel, (total + 255) / 256, 256, static_cast<cudaStream_t>(stream), t, B, half, freqs,
                  scale, out);
  return check_launch("f2g_time_sinusoid");
}

extern "C" int f2g_pack2d(const float* src, long long src_rs, long long src_cs, int rows, int cols,
                          float* dst, int ld, int ld_fill, int round_tf32, void* stream) {
  if (ld_fill < cols || ld_fill > ld) {
    set_error("f2g_pack2d: need cols <= ld_fill <= ld (%d, %d, %d)", cols, ld_fill, ld);
    return F2G_EINVAL;
  }
  const long long total = (long long)rows * ld_fill;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  F2G_LAUNCH(pack2d_kernel, blocks, 256, static_cast<cudaStream_t>(stream), src, src_rs, src_cs, rows, cols, dst,
                  ld, ld_fill, round_tf32);
  return check_launch("f2g_pack2d");
}

extern "C" int f2g_im2col_cf(const float* x, int B, int C, int T, int ktaps, float* out, int ld,
                             int round_tf32, void* stream) {
  if (ld < ktaps * C) {
    set_error("f2g_im2col_cf: ld=%d < ktaps*C=%d", ld, ktaps * C);
    return F2G_EINVAL;
  }
  const long long total = (long long)B * T * ld;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  F2G_LAUNCH(im2col_cf_kernel, blocks, 256, static_cast<cudaStream_t>(stream), x, B, C, T, ktaps, out, ld,
                  round_tf32);
  return check_launch("f2g_im2col_cf");
}

extern "C" int f2g_frame_mask(const int* lens, int B, int frames, int hop, float* out, void* stream) {
  const int total = B * frames;
  F2G_LAUNCH(frame_mask_kernel, (total + 255) / 256, 256, static_cast<cudaStream_t>(stream), lens, B, frames, hop,
                  out);
  return check_launch("f2g_frame_mask");
}
