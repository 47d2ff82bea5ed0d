// ConvNeXt-block SIMT pieces on channel-last rows: BiasNorm, fused depthwise-conv prologue,
// small dense layers, sinusoidal time embedding, packing helpers.  All HBM/L2-bound.
// Reference: flow2gan/models/modules.py:217-232 (SinusoidalPosEmb), :286-416 (BiasNorm),
// :456-495 (ConvNeXtBlock.forward), :523-542 (CondEncoder), :595-627 (ConvNeXtDecoder).
#include "common.cuh"
#include "../../include/flow2gan_b200.h"

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

// Fused ConvNeXt-block prologue.  One CTA = PRE_TOK consecutive tokens of one batch element
// processed CONCURRENTLY (threadIdx.y = token, threadIdx.x = 4-channel group): the 7-row windows
// of neighbouring tokens overlap and are served by that SM's L1, so L2 sees (PRE_TOK+6)/PRE_TOK
// reads per row instead of 7 (the one-token-per-CTA version ran at L2 bandwidth, 10 us/launch),
// while the thread count stays that of the per-token mapping (a serial strip was slower: 14 us).
//   y = dwconv7(x * mask) + b ; z = BiasNorm(y) + cond_row ; z *= 1 + ts[b] ; out = tf32(z)
constexpr int PRE_TOK = 4;

__global__ void block_pre_kernel(const float* __restrict__ x, int B, int T, int C, int ld_x,
                                 const float* __restrict__ dw_wT, const float* __restrict__ dw_b,
                                 const float* __restrict__ bn_bias,
                                 const float* __restrict__ bn_log_scale,
                                 const float* __restrict__ row_mask,
                                 const float* __restrict__ cond, int ld_cond, int cond_T,
                                 int factor, int zero_row, const float* __restrict__ tscale,
                                 int ld_ts, float* __restrict__ out, int ld_out,
                                 float* __restrict__ conv_out, float* __restrict__ inv_out) {
  __shared__ float red[PRE_TOK][8];
  const int bi = blockIdx.y;
  const int ty = threadIdx.y;
  const int t = blockIdx.x * PRE_TOK + ty;
  const bool live = t < T;
  const int c = threadIdx.x * 4;
  const int wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const size_t rb = (size_t)bi * T;
  const size_t row = rb + t;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float ssq = 0.f;
  if (live) {
    acc = ld4(dw_b + c);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int tt = t + k - 3;
      if (tt >= 0 && tt < T) {
        const float mk = row_mask ? row_mask[rb + tt] : 1.f;
        if (mk != 0.f) {
          const float4 xv = ld4(x + (rb + tt) * ld_x + c);
          const float4 w = ld4(dw_wT + k * C + c);
          acc.x = fmaf(xv.x * mk, w.x, acc.x);
          acc.y = fmaf(xv.y * mk, w.y, acc.y);
          acc.z = fmaf(xv.z * mk, w.z, acc.z);
          acc.w = fmaf(xv.w * mk, w.w, acc.w);
        }
      }
    }
    const float4 bb = ld4(bn_bias + c);
    const float d0 = acc.x - bb.x, d1 = acc.y - bb.y, d2 = acc.z - bb.z, d3 = acc.w - bb.w;
    ssq = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    if (conv_out) st4(conv_out + row * C + c, acc);
  }
  // blockDim.x (= C/4) is a multiple of 32, so each warp belongs to exactly one token
  ssq = warp_sum(ssq);
  if ((threadIdx.x & 31) == 0) red[ty][wid] = ssq;
  __syncthreads();
  if (!live) return;
  ssq = 0.f;
  for (int j = 0; j < nw; ++j) ssq += red[ty][j];
  const float inv = (1.0f / sqrtf(ssq / (float)C)) * expf(*bn_log_scale);
  if (inv_out && threadIdx.x == 0) inv_out[row] = inv;
  float4 z = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  if (cond) {
    const int crow = t < cond_T * factor ? bi * cond_T + t / factor : zero_row;
    const float4 cv = ld4(cond + (size_t)crow * ld_cond + c);
    z.x += cv.x; z.y += cv.y; z.z += cv.z; z.w += cv.w;
  }
  if (tscale) {
    const float4 sc = ld4(tscale + (size_t)bi * ld_ts + c);
    z.x *= 1.f + sc.x; z.y *= 1.f + sc.y; z.z *= 1.f + sc.z; z.w *= 1.f + sc.w;
  }
  st4(out + row * ld_out + c, make_float4(tf32_rna(z.x), tf32_rna(z.y), tf32_rna(z.z), tf32_rna(z.w)));
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
  biasnorm_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, C, ld, bias, log_scale, y, ld_y, inv_out);
  return check_launch("f2g_biasnorm");
}

extern "C" int f2g_block_pre(const float* x, int B, int T, int C, int ld_x, const float* dw_wT,
                             const float* dw_b, const float* bn_bias, const float* bn_log_scale,
                             const float* row_mask, const float* cond, int ld_cond, int cond_T,
                             int factor, int zero_row, const float* tscale, int ld_ts, float* out,
                             int ld_out, float* conv_out, float* inv_rms_out, void* stream) {
  if (int rc = check_channels("f2g_block_pre", C, ld_x | ld_out | (cond ? ld_cond : 0) | (tscale ? ld_ts : 0)))
    return rc;
  dim3 grid((T + PRE_TOK - 1) / PRE_TOK, B);
  dim3 block(C / 4, PRE_TOK);
  block_pre_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      x, B, T, C, ld_x, dw_wT, dw_b, bn_bias, bn_log_scale, row_mask, cond, ld_cond, cond_T,
      factor < 1 ? 1 : factor, zero_row, tscale, ld_ts, out, ld_out, conv_out, inv_rms_out);
  return check_launch("f2g_block_pre");
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
  linear_small_kernel<<<grid, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("f2g_linear_small");
}

extern "C" int f2g_time_sinusoid(const float* t, int B, int dim, const float* freqs, float scale,
                                 float* out, void* stream) {
  const int half = dim / 2;
  const int total = B * half;
  time_sinusoid_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      t, B, half, freqs, scale, out);
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
  pack2d_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, src_rs, src_cs, rows,
                                                                      cols, dst, ld, ld_fill,
                                                                      round_tf32);
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
  im2col_cf_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, B, C, T, ktaps, out,
                                                                         ld, round_tf32);
  return check_launch("f2g_im2col_cf");
}

extern "C" int f2g_frame_mask(const int* lens, int B, int frames, int hop, float* out, void* stream) {
  const int total = B * frames;
  frame_mask_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      lens, B, frames, hop, out);
  return check_launch("f2g_frame_mask");
}
