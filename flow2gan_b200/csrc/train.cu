// Backward-pass SIMT kernels of the generator path (channel-last rows): ConvNeXt block prologue
// adjoints, PReLU/SiLU adjoints with fused parameter-gradient column reductions, column sums
// (bias grads), conditioning scatter, iSTFT / STFT adjoints.  The contractions (dgrad / wgrad)
// run on the tcgen05 GEMM (gemm_tf32.cu) with MN-major operands.
// Reference forward definitions: flow2gan/models/modules.py:286-339 (BiasNormFunction),
// :456-495 (ConvNeXtBlock.forward), :668-680 (upsample_cond), :69-116 (STFT/ISTFT).
#include "simt.cuh"
#include "../../include/flow2gan_b200.h"

namespace f2g {

constexpr int MAXC4 = 8;
F2G_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
F2G_DEVINL void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------------------
// Block prologue backward, stage A (one warp per token):
//   forward: u = y*inv + cond ; a1 = u * (1 + ts)      (y = dwconv output, inv = rsqrt(mean((y-beta)^2)) e^ls)
//   in : da1 ; out: du = da1*(1+ts) (optional), dy = inv*du - coef*(y-beta), coef = G*inv/(C*m),
//        G = sum_c du*y, m = e^(2 ls)/inv^2 ; per row: coef[r], gs[r] = G*inv (d log_scale terms)
// ---------------------------------------------------------------------------------------
__global__ void block_bwd_a_kernel(const float* __restrict__ da1, int ld_da, const float* __restrict__ y,
                                   const float* __restrict__ inv, const float* __restrict__ bn_bias,
                                   const float* __restrict__ log_scale, const float* __restrict__ tscale,
                                   int ld_ts, int B, int T, int C, float* __restrict__ dy,
                                   float* __restrict__ du, float* __restrict__ coef_out,
                                   float* __restrict__ gs_out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T) return;
  const int bi = row / T;
  const int chunks = C >> 7;
  const float s = inv[row];
  float4 g[MAXC4], yv[MAXC4];
  float G = 0.f;
#pragma unroll
  for (int j = 0; j < MAXC4; ++j) {
    if (j < chunks) {
      const int c = j * 128 + lane * 4;
      float4 d = ld4(da1 + (size_t)row * ld_da + c);
      if (tscale) {
        const float4 t = ld4(tscale + (size_t)bi * ld_ts + c);
        d.x *= 1.f + t.x; d.y *= 1.f + t.y; d.z *= 1.f + t.z; d.w *= 1.f + t.w;
      }
      g[j] = d;
      yv[j] = ld4(y + (size_t)row * C + c);
      G += d.x * yv[j].x + d.y * yv[j].y + d.z * yv[j].z + d.w * yv[j].w;
      if (du) st4(du + (size_t)row * C + c, d);
    }
  }
  G = warp_sum(G);
  const float e2 = expf(2.f * (*log_scale));
  const float m = e2 / (s * s);
  const float coef = G * s / ((float)C * m);
  if (lane == 0) {
    coef_out[row] = coef;
    gs_out[row] = G * s;
  }
#pragma unroll
  for (int j = 0; j < MAXC4; ++j) {
    if (j < chunks) {
      const int c = j * 128 + lane * 4;
      const float4 b = ld4(bn_bias + c);
      float4 o;
      o.x = s * g[j].x - coef * (yv[j].x - b.x);
      o.y = s * g[j].y - coef * (yv[j].y - b.y);
      o.z = s * g[j].z - coef * (yv[j].z - b.z);
      o.w = s * g[j].w - coef * (yv[j].w - b.w);
      st4(dy + (size_t)row * C + c, o);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Stage C: all per-channel parameter-gradient reductions of one block, thread = channel,
// CTA = (128-channel slab) x (chunk of rows inside ONE batch element):
//   d dw_w[k][c] += sum_t dy[t,c] * xm[t+k-3,c] ; d dw_b[c] += sum dy ; d beta[c] += sum coef*(y-beta)
//   d ts[b,c] += sum_t da1 * (y*inv + cond) ; d rs[c] += sum dxo * x ; d b2[c] += sum dxo
//   d log_scale += sum gs            (dxo = gradient w.r.t. the block output)
// Any pointer may be NULL to skip its term.
// ---------------------------------------------------------------------------------------
struct BlockBwdC {
  const float* dy; const float* x; int ld_x; const float* row_mask; const float* y;
  const float* coef; const float* gs; const float* bn_bias; const float* da1; int ld_da;
  const float* inv; const float* cond; int ld_cond; int cond_T; int factor; int zero_row;
  const float* dxo; int ld_dxo;
  float* g_dww; float* g_dwb; float* g_beta; float* g_ls; float* g_ts; int ld_gts; float* g_rs; float* g_b2;
  int B, T, C, rows_per_cta;
};

__global__ void block_bwd_c_kernel(const BlockBwdC a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;     // channel
  const int bi = blockIdx.z;
  const int t0 = blockIdx.y * a.rows_per_cta;
  const int t1 = min(t0 + a.rows_per_cta, a.T);
  if (c >= a.C) return;
  const size_t rb = (size_t)bi * a.T;
  float w7[7] = {0, 0, 0, 0, 0, 0, 0};
  float s_dy = 0.f, s_beta = 0.f, s_ts = 0.f, s_rs = 0.f, s_b2 = 0.f, s_ls = 0.f;
  const float beta = a.bn_bias ? a.bn_bias[c] : 0.f;
  // sliding window of masked inputs x[t-3 .. t+3]
  float win[7];
  if (a.g_dww) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int tt = t0 + k - 3;
      float v = 0.f;
      if (tt >= 0 && tt < a.T) {
        v = a.x[(rb + tt) * a.ld_x + c];
        if (a.row_mask) v *= a.row_mask[rb + tt];
      }
      win[k] = v;
    }
  }
  for (int t = t0; t < t1; ++t) {
    const size_t r = rb + t;
    const float dyv = a.dy ? a.dy[r * a.C + c] : 0.f;
    if (a.g_dww) {
#pragma unroll
      for (int k = 0; k < 7; ++k) w7[k] = fmaf(dyv, win[k], w7[k]);
#pragma unroll
      for (int k = 0; k < 6; ++k) win[k] = win[k + 1];
      const int tn = t + 4;
      float v = 0.f;
      if (tn < a.T) {
        v = a.x[(rb + tn) * a.ld_x + c];
        if (a.row_mask) v *= a.row_mask[rb + tn];
      }
      win[6] = v;
    }
    s_dy += dyv;
    float yv = 0.f;
    if (a.y) yv = a.y[r * a.C + c];
    if (a.g_beta) s_beta = fmaf(a.coef[r], yv - beta, s_beta);
    if (a.g_ts) {
      float u = yv * a.inv[r];
      if (a.cond) {
        const int crow = t < a.cond_T * a.factor ? bi * a.cond_T + t / a.factor : a.zero_row;
        u += a.cond[(size_t)crow * a.ld_cond + c];
      }
      s_ts = fmaf(a.da1[r * a.ld_da + c], u, s_ts);
    }
    if (a.g_rs || a.g_b2) {
      const float d = a.dxo[r * a.ld_dxo + c];
      s_b2 += d;
      if (a.g_rs) s_rs = fmaf(d, a.x[r * a.ld_x + c], s_rs);
    }
    if (a.g_ls && c == 0) s_ls += a.gs[r];
  }
  if (a.g_dww) {
#pragma unroll
    for (int k = 0; k < 7; ++k) atomicAdd(a.g_dww + (size_t)k * a.C + c, w7[k]);
  }
  if (a.g_dwb) atomicAdd(a.g_dwb + c, s_dy);
  if (a.g_beta) atomicAdd(a.g_beta + c, s_beta);
  if (a.g_ts) atomicAdd(a.g_ts + (size_t)bi * a.ld_gts + c, s_ts);
  if (a.g_rs) atomicAdd(a.g_rs + c, s_rs);
  if (a.g_b2) atomicAdd(a.g_b2 + c, s_b2);
  if (a.g_ls && c == 0) atomicAdd(a.g_ls, s_ls);
}

// Stage B (one warp per token): dx[t,c] = mask[t] * sum_k w[c,k] dy[t-k+3,c] + rs[c] * dxo[t,c]
__global__ void block_bwd_b_kernel(const float* __restrict__ dy, const float* __restrict__ dw_wT,
                                   const float* __restrict__ row_mask, const float* __restrict__ dxo,
                                   int ld_dxo, const float* __restrict__ rs, int B, int T, int C,
                                   float* __restrict__ dx, int ld_dx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T) return;
  const int bi = row / T;
  const int t = row - bi * T;
  const int chunks = C >> 7;
  const float mk = row_mask ? row_mask[row] : 1.f;
#pragma unroll
  for (int j = 0; j < MAXC4; ++j) {
    if (j < chunks) {
      const int c = j * 128 + lane * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dy) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const int tt = t - k + 3;
          if (tt >= 0 && tt < T) {
            const float4 d = ld4(dy + (size_t)(bi * T + tt) * C + c);
            const float4 w = ld4(dw_wT + k * C + c);
            acc.x = fmaf(d.x, w.x, acc.x); acc.y = fmaf(d.y, w.y, acc.y);
            acc.z = fmaf(d.z, w.z, acc.z); acc.w = fmaf(d.w, w.w, acc.w);
          }
        }
        acc.x *= mk; acc.y *= mk; acc.z *= mk; acc.w *= mk;
      }
      if (dxo) {
        const float4 d = ld4(dxo + (size_t)row * ld_dxo + c);
        float4 r = make_float4(1.f, 1.f, 1.f, 1.f);
        if (rs) r = ld4(rs + c);
        acc.x = fmaf(d.x, r.x, acc.x); acc.y = fmaf(d.y, r.y, acc.y);
        acc.z = fmaf(d.z, r.z, acc.z); acc.w = fmaf(d.w, r.w, acc.w);
      }
      st4(dx + (size_t)row * ld_dx + c, acc);
    }
  }
}

// PReLU / LeakyReLU / SiLU backward over rows with fused column reductions (thread = column):
//   dz = dh * act'(z) (written, optionally TF32-rounded); g_bias[c] += sum dz ; g_slope[c] += sum dh*min(z,0)
__global__ void act_bwd_kernel(const float* __restrict__ dh, int ld_dh, const float* __restrict__ z,
                               int ld_z, const float* __restrict__ slope, float leaky, int act, int rows,
                               int cols, int rows_per_cta, float* __restrict__ dz, int ld_dz,
                               float* __restrict__ g_bias, float* __restrict__ g_slope, int round_tf32) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  const float sl = slope ? slope[c] : leaky;
  float sb = 0.f, ss = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float d = dh[(size_t)r * ld_dh + c];
    const float zv = z ? z[(size_t)r * ld_z + c] : 1.f;
    float o;
    if (act == F2G_ACT_SILU) {
      const float sg = 1.f / (1.f + expf(-zv));
      o = d * sg * (1.f + zv * (1.f - sg));
    } else if (act == F2G_ACT_NONE) {
      o = d;
    } else {
      o = zv > 0.f ? d : d * sl;
      ss = fmaf(d, fminf(zv, 0.f), ss);
    }
    sb += o;
    if (dz) dz[(size_t)r * ld_dz + c] = round_tf32 ? tf32_rna(o) : o;
  }
  if (g_bias) atomicAdd(g_bias + c, sb);
  if (g_slope) atomicAdd(g_slope + c, ss);
}

// The same with one column QUAD per thread and several rows per CTA pass (TPR threads per row,
// 128 / TPR rows per pass), four passes in flight: the column-per-thread version keeps two 4-byte
// loads per thread outstanding (10 % of HBM bandwidth on the discriminator tensors, where 32
// columns also left 3 of 4 warps idle).  Column sums: registers -> shared -> one atomic per column.
template <int TPR>
__global__ void __launch_bounds__(128) act_bwd_vec_kernel(
    const float* __restrict__ dh, int ld_dh, const float* __restrict__ z, int ld_z,
    const float* __restrict__ slope, float leaky, int act, int rows, int cols, int rows_per_cta,
    float* __restrict__ dz, int ld_dz, float* __restrict__ g_bias, float* __restrict__ g_slope,
    int round_tf32) {
  constexpr int RPP = 128 / TPR;                  // rows per pass
  __shared__ float4 red[2][128];
  const int cq = threadIdx.x % TPR, rs = threadIdx.x / TPR;
  const int c = (blockIdx.x * TPR + cq) * 4;
  const bool active = c < cols;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  float4 sl = make_float4(leaky, leaky, leaky, leaky);
  if (slope && active) sl = ld4(slope + c);
  float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), ss = sb;
  if (active) {
#pragma unroll 4
    for (int r = r0 + rs; r < r1; r += RPP) {
      const float4 d = ld4(dh + (size_t)r * ld_dh + c);
      const float4 zv = z ? ld4(z + (size_t)r * ld_z + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      float4 o;
      if (act == F2G_ACT_SILU) {
        const float g0 = 1.f / (1.f + expf(-zv.x)), g1 = 1.f / (1.f + expf(-zv.y));
        const float g2 = 1.f / (1.f + expf(-zv.z)), g3 = 1.f / (1.f + expf(-zv.w));
        o = make_float4(d.x * g0 * (1.f + zv.x * (1.f - g0)), d.y * g1 * (1.f + zv.y * (1.f - g1)),
                        d.z * g2 * (1.f + zv.z * (1.f - g2)), d.w * g3 * (1.f + zv.w * (1.f - g3)));
      } else if (act == F2G_ACT_NONE) {
        o = d;
      } else {
        o = make_float4(zv.x > 0.f ? d.x : d.x * sl.x, zv.y > 0.f ? d.y : d.y * sl.y,
                        zv.z > 0.f ? d.z : d.z * sl.z, zv.w > 0.f ? d.w : d.w * sl.w);
        ss.x = fmaf(d.x, fminf(zv.x, 0.f), ss.x); ss.y = fmaf(d.y, fminf(zv.y, 0.f), ss.y);
        ss.z = fmaf(d.z, fminf(zv.z, 0.f), ss.z); ss.w = fmaf(d.w, fminf(zv.w, 0.f), ss.w);
      }
      sb.x += o.x; sb.y += o.y; sb.z += o.z; sb.w += o.w;
      if (dz) {
        if (round_tf32) o = make_float4(tf32_rna(o.x), tf32_rna(o.y), tf32_rna(o.z), tf32_rna(o.w));
        st4(dz + (size_t)r * ld_dz + c, o);
      }
    }
  }
  red[0][threadIdx.x] = sb;
  red[1][threadIdx.x] = ss;
  __syncthreads();
  if (rs == 0 && active) {
    float4 tb = make_float4(0.f, 0.f, 0.f, 0.f), ts = tb;
#pragma unroll
    for (int j = 0; j < RPP; ++j) {
      const float4 a = red[0][j * TPR + cq], b = red[1][j * TPR + cq];
      tb.x += a.x; tb.y += a.y; tb.z += a.z; tb.w += a.w;
      ts.x += b.x; ts.y += b.y; ts.z += b.z; ts.w += b.w;
    }
    if (g_bias) {
      atomicAdd(g_bias + c, tb.x); atomicAdd(g_bias + c + 1, tb.y);
      atomicAdd(g_bias + c + 2, tb.z); atomicAdd(g_bias + c + 3, tb.w);
    }
    if (g_slope) {
      atomicAdd(g_slope + c, ts.x); atomicAdd(g_slope + c + 1, ts.y);
      atomicAdd(g_slope + c + 2, ts.z); atomicAdd(g_slope + c + 3, ts.w);
    }
  }
}

// Activation backward for the windowed convs (convwin.py): the GEMM-row space is (n, line, r) with
// Hl lines of R columns per image, of which only line < Ho, r < Wo are real outputs.  Gathers the
// incoming gradient from its (N, Ho, Wo, C) view, applies act'(z), and writes EVERY row of dz
// (zeros on the padding rows / columns and on `guard` rows in front) plus the bias-gradient column
// sums -- one pass instead of zero-fill + strided copy + in-place act_bwd.
template <int TPR>
__global__ void __launch_bounds__(128) act_bwd_win_kernel(
    const float* __restrict__ dy, long long s_n, long long s_h, long long s_w, int Hl, int R, int Ho, int Wo,
    const float* __restrict__ z, int ld_z, float leaky, int act, int rows, int cols, int cols_total,
    int rows_per_cta, float* __restrict__ dz, int ld_dz, int guard, float* __restrict__ g_bias, int round_tf32) {
  constexpr int RPP = 128 / TPR;
  __shared__ float4 red[128];
  const int cq = threadIdx.x % TPR, rs = threadIdx.x / TPR;
  const int c = (blockIdx.x * TPR + cq) * 4;
  const bool active = c < cols_total, real_col = c < cols;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sb = zero4;
  if (active) {
    if (blockIdx.y == 0)
      for (int r = rs; r < guard; r += RPP) st4(dz - (size_t)(r + 1) * ld_dz + c, zero4);
    const int per_img = Hl * R;
#pragma unroll 4
    for (int r = r0 + rs; r < r1; r += RPP) {
      const int n = r / per_img;
      const int rem = r - n * per_img;
      const int line = rem / R;
      const int rr = rem - line * R;
      float4 o = zero4;
      if (real_col && line < Ho && rr < Wo) {
        const float4 d = ld4(dy + n * s_n + line * s_h + rr * s_w + c);
        if (act == F2G_ACT_NONE) {
          o = d;
        } else {
          const float4 zv = ld4(z + (size_t)r * ld_z + c);
          o = make_float4(zv.x > 0.f ? d.x : d.x * leaky, zv.y > 0.f ? d.y : d.y * leaky,
                          zv.z > 0.f ? d.z : d.z * leaky, zv.w > 0.f ? d.w : d.w * leaky);
        }
        sb.x += o.x; sb.y += o.y; sb.z += o.z; sb.w += o.w;
        if (round_tf32) o = make_float4(tf32_rna(o.x), tf32_rna(o.y), tf32_rna(o.z), tf32_rna(o.w));
      }
      st4(dz + (size_t)r * ld_dz + c, o);
    }
  }
  red[threadIdx.x] = sb;
  __syncthreads();
  if (rs == 0 && real_col && g_bias) {
    float4 tb = zero4;
#pragma unroll
    for (int j = 0; j < RPP; ++j) {
      const float4 a = red[j * TPR + cq];
      tb.x += a.x; tb.y += a.y; tb.z += a.z; tb.w += a.w;
    }
    atomicAdd(g_bias + c, tb.x); atomicAdd(g_bias + c + 1, tb.y);
    atomicAdd(g_bias + c + 2, tb.z); atomicAdd(g_bias + c + 3, tb.w);
  }
}

// dcp[b*cond_T + tm, c] = sum_{f<factor} du[b, tm*factor+f, c]; zero row = all remaining frames
__global__ void cond_reduce_kernel(const float* __restrict__ du, int B, int T, int C, int cond_T, int factor,
                                   int zero_row, float* __restrict__ out, int ld_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;     // 0 .. B*cond_T (last = zero row)
  if (c >= C) return;
  float s = 0.f;
  if (row < B * cond_T) {
    const int bi = row / cond_T, tm = row - bi * cond_T;
    for (int f = 0; f < factor; ++f) {
      const int t = tm * factor + f;
      if (t < T) s += du[((size_t)bi * T + t) * C + c];
    }
    out[(size_t)row * ld_out + c] = s;
  } else {
    for (int bi = 0; bi < B; ++bi)
      for (int t = cond_T * factor; t < T; ++t) s += du[((size_t)bi * T + t) * C + c];
    out[(size_t)zero_row * ld_out + c] = s;
  }
}

// iSTFT adjoint, first half: gs[b, p] = g[b, s] * scale / env(p) at p = s + n/2 for s < hop*(F-1), else 0
__global__ void istft_bwd_prep_kernel(const float* __restrict__ g, int T, int n, int hop, int frames,
                                      float scale, float* __restrict__ gs, int Lp) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int bi = blockIdx.y;
  if (p >= Lp) return;
  const int s = p - (n >> 1);
  float v = 0.f;
  if (s >= 0 && s < hop * (frames - 1) && s < T) {
    const int f_hi = min(p / hop, frames - 1);
    const int f_lo = p >= n ? (p - n) / hop + 1 : 0;
    float env = 0.f;
    for (int f = f_lo; f <= f_hi; ++f) {
      const float w = 0.5f - 0.5f * cospif(2.0f * (float)(p - f * hop) / (float)n);
      env += w * w;
    }
    v = g[(size_t)bi * T + s] * scale / env;
  }
  gs[(size_t)bi * Lp + p] = v;
}

// STFT adjoint, second half: fold windowed frame gradients back onto the (reflect padded) signal:
//   dx[s] = OLA(s + n/2) + [1 <= s <= n/2] OLA(n/2 - s) + [T-1-n/2 <= s <= T-2] OLA(n/2 + 2(T-1) - s)
__global__ void stft_bwd_fold_kernel(const float* __restrict__ fr, int n, int hop, int frames, int T,
                                     float* __restrict__ dx, int accumulate) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int bi = blockIdx.y;
  if (s >= T) return;
  const int h = n >> 1;
  int pos[3];
  int np = 0;
  pos[np++] = s + h;
  if (s >= 1 && s <= h) pos[np++] = h - s;
  if (s >= T - 1 - h && s <= T - 2) pos[np++] = h + 2 * (T - 1) - s;
  float acc = 0.f;
  for (int q = 0; q < np; ++q) {
    const int p = pos[q];
    const int f_hi = min(p / hop, frames - 1);
    const int f_lo = p >= n ? (p - n) / hop + 1 : 0;
    for (int f = f_lo; f <= f_hi; ++f) acc += fr[((size_t)bi * frames + f) * n + (p - f * hop)];
  }
  float* o = dx + (size_t)bi * T + s;
  *o = accumulate ? *o + acc : acc;
}

__global__ void colsum_kernel(const float* __restrict__ x, int ld, int rows, int cols, int rows_per_cta,
                              float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += x[(size_t)r * ld + c];
  atomicAdd(out + c, s);
}

}  // namespace f2g

using namespace f2g;

static int pick_rows_per_cta(int rows, int col_blocks) {
  // aim for ~4 CTAs per SM overall
  int want = (148 * 4 + col_blocks - 1) / col_blocks;
  if (want < 1) want = 1;
  int rpc = (rows + want - 1) / want;
  return rpc < 8 ? 8 : rpc;
}

extern "C" int f2g_block_bwd_a(const float* da1, int ld_da, const float* y, const float* inv,
                               const float* bn_bias, const float* log_scale, const float* tscale,
                               int ld_ts, int B, int T, int C, float* dy, float* du, float* coef,
                               float* gs, void* stream) {
  if (C % 128 != 0 || C > MAXC4 * 128) {
    set_error("f2g_block_bwd_a: channels=%d must be a multiple of 128 (<= %d)", C, MAXC4 * 128);
    return F2G_EINVAL;
  }
  const int rows = B * T;
  F2G_LAUNCH_COOP(block_bwd_a_kernel, (rows + 3) / 4, 128, static_cast<cudaStream_t>(stream),  da1, ld_da, y, inv, bn_bias, log_scale, tscale, ld_ts, B, T, C, dy, du, coef, gs);
  return check_launch("f2g_block_bwd_a");
}

extern "C" int f2g_block_bwd_c(const F2GBlockBwdC* p, void* stream) {
  BlockBwdC a;
  a.dy = p->dy; a.x = p->x; a.ld_x = p->ld_x; a.row_mask = p->row_mask; a.y = p->y;
  a.coef = p->coef; a.gs = p->gs; a.bn_bias = p->bn_bias; a.da1 = p->da1; a.ld_da = p->ld_da;
  a.inv = p->inv; a.cond = p->cond; a.ld_cond = p->ld_cond; a.cond_T = p->cond_T;
  a.factor = p->factor < 1 ? 1 : p->factor; a.zero_row = p->zero_row;
  a.dxo = p->dxo; a.ld_dxo = p->ld_dxo;
  a.g_dww = p->g_dww; a.g_dwb = p->g_dwb; a.g_beta = p->g_beta; a.g_ls = p->g_ls; a.g_ts = p->g_ts;
  a.ld_gts = p->ld_gts; a.g_rs = p->g_rs; a.g_b2 = p->g_b2;
  a.B = p->B; a.T = p->T; a.C = p->C;
  const int cb = (a.C + 127) / 128;
  int want = (148 * 4) / (cb * a.B);
  if (want < 1) want = 1;
  a.rows_per_cta = (a.T + want - 1) / want;
  if (a.rows_per_cta < 8) a.rows_per_cta = 8;
  dim3 grid(cb, (a.T + a.rows_per_cta - 1) / a.rows_per_cta, a.B);
  F2G_LAUNCH_COOP(block_bwd_c_kernel, grid, 128, static_cast<cudaStream_t>(stream), a);
  return check_launch("f2g_block_bwd_c");
}

extern "C" int f2g_block_bwd_b(const float* dy, const float* dw_wT, const float* row_mask,
                               const float* dxo, int ld_dxo, const float* rs, int B, int T, int C,
                               float* dx, int ld_dx, void* stream) {
  if (C % 128 != 0 || C > MAXC4 * 128) {
    set_error("f2g_block_bwd_b: channels=%d must be a multiple of 128 (<= %d)", C, MAXC4 * 128);
    return F2G_EINVAL;
  }
  const int rows = B * T;
  F2G_LAUNCH_COOP(block_bwd_b_kernel, (rows + 3) / 4, 128, static_cast<cudaStream_t>(stream),  dy, dw_wT, row_mask, dxo, ld_dxo, rs, B, T, C, dx, ld_dx);
  return check_launch("f2g_block_bwd_b");
}

extern "C" int f2g_act_bwd(const float* dh, int ld_dh, const float* z, int ld_z, const float* slope,
                           float leaky, int act, int rows, int cols, float* dz, int ld_dz,
                           float* g_bias, float* g_slope, int round_tf32, void* stream) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (cols & 3) == 0 && (ld_dh & 3) == 0 && (!z || (ld_z & 3) == 0) && (!dz || (ld_dz & 3) == 0) &&
                   al16(dh) && al16(z) && al16(dz) && al16(slope) && al16(g_bias) && al16(g_slope);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) {
    const int quads = cols >> 2;
    const int tpr = quads >= 32 ? 32 : (quads > 8 ? 16 : 8);
    const int cb = (quads + tpr - 1) / tpr;
    int rpc = pick_rows_per_cta(rows, cb);
    const int rpp = 128 / tpr;                 // whole passes only, >= 4 passes in flight per CTA
    rpc = ((rpc + 4 * rpp - 1) / (4 * rpp)) * (4 * rpp);
    dim3 grid(cb, (rows + rpc - 1) / rpc);
    if (tpr == 32)
      F2G_LAUNCH_COOP(act_bwd_vec_kernel<32>, grid, 128, st, dh, ld_dh, z, ld_z, slope, leaky, act, rows, cols, rpc, dz, ld_dz, g_bias, g_slope, round_tf32);
    else if (tpr == 16)
      F2G_LAUNCH_COOP(act_bwd_vec_kernel<16>, grid, 128, st, dh, ld_dh, z, ld_z, slope, leaky, act, rows, cols, rpc, dz, ld_dz, g_bias, g_slope, round_tf32);
    else
      F2G_LAUNCH_COOP(act_bwd_vec_kernel<8>, grid, 128, st, dh, ld_dh, z, ld_z, slope, leaky, act, rows, cols, rpc, dz, ld_dz, g_bias, g_slope, round_tf32);
    return check_launch("f2g_act_bwd");
  }
  const int cb = (cols + 127) / 128;
  const int rpc = pick_rows_per_cta(rows, cb);
  dim3 grid(cb, (rows + rpc - 1) / rpc);
  F2G_LAUNCH_COOP(act_bwd_kernel, grid, 128, st,  dh, ld_dh, z, ld_z, slope, leaky, act, rows, cols, rpc, dz, ld_dz, g_bias, g_slope, round_tf32);
  return check_launch("f2g_act_bwd");
}

extern "C" int f2g_act_bwd_win(const float* dy, long long s_n, long long s_h, long long s_w, int Nb, int Hl,
                               int R, int Ho, int Wo, const float* z, int ld_z, float leaky, int act, int cols,
                               int cols_total, float* dz, int ld_dz, int guard_rows, float* g_bias,
                               int round_tf32, void* stream) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((cols & 3) || (cols_total & 3) || cols > cols_total || cols_total > ld_dz || (ld_dz & 3) || (z && (ld_z & 3)) ||
      (s_n & 3) || (s_h & 3) || (s_w & 3) || !al16(dy) || !al16(z) || !al16(dz) || !al16(g_bias) ||
      (act != F2G_ACT_NONE && act != F2G_ACT_LEAKY) || (act == F2G_ACT_LEAKY && !z) || Ho > Hl || Wo > R) {
    set_error("f2g_act_bwd_win: needs 4-float aligned columns / strides / pointers and act NONE or LEAKY "
              "(cols=%d/%d ld_dz=%d)", cols, cols_total, ld_dz);
    return F2G_EINVAL;
  }
  const long long rows_ll = (long long)Nb * Hl * R;
  if (rows_ll <= 0 || rows_ll > 0x7fffffffLL) {
    set_error("f2g_act_bwd_win: %lld rows out of range", rows_ll);
    return F2G_EINVAL;
  }
  const int rows = (int)rows_ll;
  const int quads = cols_total >> 2;
  const int tpr = quads >= 32 ? 32 : (quads > 8 ? 16 : 8);
  const int cb = (quads + tpr - 1) / tpr;
  int rpc = pick_rows_per_cta(rows, cb);
  const int rpp = 128 / tpr;
  rpc = ((rpc + 4 * rpp - 1) / (4 * rpp)) * (4 * rpp);
  dim3 grid(cb, (rows + rpc - 1) / rpc);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define F2G_ABW(T_)                                                                                            \
  F2G_LAUNCH_COOP(act_bwd_win_kernel<T_>, grid, 128, st, dy, s_n, s_h, s_w, Hl, R, Ho, Wo, z, ld_z, leaky, act, rows, \
                  cols, cols_total, rpc, dz, ld_dz, guard_rows, g_bias, round_tf32)
  if (tpr == 32) F2G_ABW(32);
  else if (tpr == 16) F2G_ABW(16);
  else F2G_ABW(8);
#undef F2G_ABW
  return check_launch("f2g_act_bwd_win");
}

extern "C" int f2g_cond_reduce(const float* du, int B, int T, int C, int cond_T, int factor, int zero_row,
                               float* out, int ld_out, void* stream) {
  dim3 grid((C + 127) / 128, B * cond_T + 1);
  F2G_LAUNCH_COOP(cond_reduce_kernel, grid, 128, static_cast<cudaStream_t>(stream), du, B, T, C, cond_T, factor < 1 ? 1 : factor, zero_row, out, ld_out);
  return check_launch("f2g_cond_reduce");
}

extern "C" int f2g_istft_bwd_prep(const float* g, int B, int T, int n_fft, int hop, int frames,
                                  float scale, float* gs, void* stream) {
  const int Lp = n_fft + hop * (frames - 1);
  dim3 grid((Lp + 255) / 256, B);
  F2G_LAUNCH(istft_bwd_prep_kernel, grid, 256, static_cast<cudaStream_t>(stream), g, T, n_fft, hop, frames, scale, gs, Lp);
  return check_launch("f2g_istft_bwd_prep");
}

extern "C" int f2g_stft_bwd_fold(const float* frames_grad, int B, int T, int n_fft, int hop, int frames,
                                 float* dx, int accumulate, void* stream) {
  dim3 grid((T + 255) / 256, B);
  F2G_LAUNCH(stft_bwd_fold_kernel, grid, 256, static_cast<cudaStream_t>(stream), frames_grad, n_fft, hop, frames, T, dx, accumulate);
  return check_launch("f2g_stft_bwd_fold");
}

extern "C" int f2g_colsum(const float* x, int ld, int rows, int cols, float* out, void* stream) {
  const int cb = (cols + 127) / 128;
  const int rpc = pick_rows_per_cta(rows, cb);
  dim3 grid(cb, (rows + rpc - 1) / rpc);
  F2G_LAUNCH_COOP(colsum_kernel, grid, 128, static_cast<cudaStream_t>(stream), x, ld, rows, cols, rpc, out);
  return check_launch("f2g_colsum");
}
