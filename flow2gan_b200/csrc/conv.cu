// Conv2d of the discriminators as gather + tensor-core GEMM on channel-last tensors:
//   im2col2d : (Nb, H, W, C) [a W-band of a wider tensor allowed] -> (Nb*Ho*Wo, kh*kw*C [+pad])
//   col2im2d : adjoint (gather form), accumulates every (ih, iw) tap that touches an input pixel
//   conv weight pack / unpack between the parameter layout (Co, Ci, kh, kw) and the GEMM layout
//   (Co_pad, (ih*kw + iw)*Ci + ci)
// Reference: torch.nn.Conv2d call sites flow2gan/models/discriminators.py:65-76,95-104,171-184,
// 203-217 ((5,1)/(3,1) strided convs of DiscriminatorP, (3,9)/(3,3) convs of DiscriminatorR).
#include "simt.cuh"
#include "../../include/flow2gan_b200.h"

namespace f2g {

struct ConvGeom {
  int Nb, H, W, C;        // logical input (band) dims
  long long pitch_h;      // elements between consecutive h rows of the underlying buffer
  long long pitch_n;      // elements between consecutive n images
  int kh, kw, sh, sw, ph, pw, Ho, Wo;
  int ldk;                // leading dim of the col matrix (>= kh*kw*C, multiple of 4)
};

// One warp per output row m = (n, ho, wo): the (n, ho, wo) decomposition is done once per row,
// lanes stride over K.  VEC: C % 4 == 0 -> 128-bit loads/stores of 4 channels.
template <bool VEC>
__global__ void im2col2d_kernel(const float* __restrict__ x, ConvGeom g, float* __restrict__ col,
                                int round_tf32) {
  const long long M = (long long)g.Nb * g.Ho * g.Wo;
  const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int wo = (int)(m % g.Wo);
  const long long t = m / g.Wo;
  const int ho = (int)(t % g.Ho);
  const int n = (int)(t / g.Ho);
  const int h0 = ho * g.sh - g.ph, w0 = wo * g.sw - g.pw;
  const float* xb = x + (long long)n * g.pitch_n;
  float* crow = col + m * g.ldk;
  const int K = g.kh * g.kw * g.C;
  if (VEC) {
    const int K4 = K >> 2, C4 = g.C >> 2;
    for (int kq = lane; kq < (g.ldk >> 2); kq += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kq < K4) {
        const int tap = kq / C4, c = (kq - tap * C4) << 2;
        const int ih = tap / g.kw, iw = tap - ih * g.kw;
        const int h = h0 + ih, w = w0 + iw;
        if (h >= 0 && h < g.H && w >= 0 && w < g.W)
          v = *reinterpret_cast<const float4*>(xb + (long long)h * g.pitch_h + (long long)w * g.C + c);
        if (round_tf32) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
      }
      *reinterpret_cast<float4*>(crow + (kq << 2)) = v;
    }
  } else {
    for (int k = lane; k < g.ldk; k += 32) {
      float v = 0.f;
      if (k < K) {
        const int tap = k / g.C, c = k - tap * g.C;
        const int ih = tap / g.kw, iw = tap - ih * g.kw;
        const int h = h0 + ih, w = w0 + iw;
        if (h >= 0 && h < g.H && w >= 0 && w < g.W)
          v = xb[(long long)h * g.pitch_h + (long long)w * g.C + c];
        if (round_tf32) v = tf32_rna(v);
      }
      crow[k] = v;
    }
  }
}

// adjoint: one thread per input element (VEC: per 4 channels) gathers every tap that read it
template <bool VEC>
__global__ void col2im2d_kernel(const float* __restrict__ dcol, ConvGeom g, float* __restrict__ dx,
                                int accumulate) {
  const int CV = VEC ? (g.C >> 2) : g.C;
  const long long total = (long long)g.Nb * g.H * g.W * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long t = i / CV;
    const int w = (int)(t % g.W);
    t /= g.W;
    const int h = (int)(t % g.H);
    const int n = (int)(t / g.H);
    const int c = VEC ? (cv << 2) : cv;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ih = 0; ih < g.kh; ++ih) {
      const int hn = h + g.ph - ih;
      if (hn < 0 || hn % g.sh != 0) continue;
      const int ho = hn / g.sh;
      if (ho >= g.Ho) continue;
      for (int iw = 0; iw < g.kw; ++iw) {
        const int wn = w + g.pw - iw;
        if (wn < 0 || wn % g.sw != 0) continue;
        const int wo = wn / g.sw;
        if (wo >= g.Wo) continue;
        const long long m = ((long long)n * g.Ho + ho) * g.Wo + wo;
        const float* src = dcol + m * g.ldk + (ih * g.kw + iw) * g.C + c;
        if (VEC) {
          const float4 v = *reinterpret_cast<const float4*>(src);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        } else {
          acc.x += *src;
        }
      }
    }
    float* o = dx + (long long)n * g.pitch_n + (long long)h * g.pitch_h + (long long)w * g.C + c;
    if (VEC) {
      float4 r = acc;
      if (accumulate) { const float4 p = *reinterpret_cast<const float4*>(o); r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w; }
      *reinterpret_cast<float4*>(o) = r;
    } else {
      *o = accumulate ? *o + acc.x : acc.x;
    }
  }
}


// Zero-padded (and TF32-rounded) copy feeding the windowed convs: the GEMM then reads its A
// operand straight from this buffer through an overlapping-row tensor map, so the im2col
// matrix (kh*kw / (sh*sw) times larger) is never written.  One thread per 4 channels.
__global__ void pad2d_kernel(const float* __restrict__ x, int Nb, int H, int W, int C4, long long pitch_n,
                             long long pitch_h, long long pitch_w, int Hl, int Wp, int ph, int pw,
                             long long total4, long long body4, int round_tf32, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < body4) {
      const int cq = (int)(i % C4);
      long long t = i / C4;
      const int w = (int)(t % Wp) - pw;
      t /= Wp;
      const int h = (int)(t % Hl) - ph;
      const int n = (int)(t / Hl);
      if (h >= 0 && h < H && w >= 0 && w < W) {
        v = *reinterpret_cast<const float4*>(x + n * pitch_n + h * pitch_h + w * pitch_w + (cq << 2));
        if (round_tf32) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
      }
    }
    *reinterpret_cast<float4*>(out + (i << 2)) = v;
  }
}

// dir 0: param (Co, Ci, kh*kw) -> packed (Co_pad rows, ld), TF32 rounded, padding zeroed
// dir 1: packed gradient -> param layout (no rounding)
__global__ void conv_w_pack_kernel(const float* __restrict__ src, int Co, int Ci, int taps, int Co_pad,
                                   int ld, float* __restrict__ dst, int dir) {
  if (dir == 0) {
    const long long total = (long long)Co_pad * ld;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int k = (int)(i % ld), co = (int)(i / ld);
      float v = 0.f;
      if (co < Co && k < taps * Ci) {
        const int ci = k % Ci, tap = k / Ci;
        v = src[((long long)co * Ci + ci) * taps + tap];
      }
      dst[i] = tf32_rna(v);
    }
  } else {
    const long long total = (long long)Co * Ci * taps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int tap = (int)(i % taps);
      const long long t = i / taps;
      const int ci = (int)(t % Ci), co = (int)(t / Ci);
      dst[i] = src[(long long)co * ld + tap * Ci + ci];
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Direct convolution for the FIRST layer of every discriminator stack (Cin = 1: DiscriminatorP, the
// (5,1) stride-3 conv run along the contiguous axis; Cin = 2: DiscriminatorR, (3,9) on [re, im]),
// Cout = 32, LeakyReLU fused (flow2gan/models/discriminators.py:65,171).  K = kh*kw*Cin <= 64, so a
// GEMM formulation has to MATERIALISE an im2col matrix 27-54x the size of the input (1 GB per D+G pair,
// 20 % of the pair's GPU time for 2 % of its FLOPs, tools/train_glue_census.py) -- here the taps are
// read from the input itself (L1-resident), in fp32 (no TF32 rounding on this layer at all).
// Thread = (output pixel, 4 output channels): 8 lanes per pixel, 128-byte coalesced pixel rows.
// ---------------------------------------------------------------------------------------------
struct SmallConv {
  int Nb, H, W, Ho, Wo;
  long long pitch_n, pitch_h, pitch_w;      // input strides in elements (channel stride 1)
  int kh, kw, sh, sw, ph, pw;
  float leaky;                              // < 0: no activation
};
constexpr int SC_CO = 32;
constexpr int SC_MAX_K = 64;
constexpr int SC_THREADS = 256;
constexpr int SCW_THREADS = 64;
constexpr int SCW_TILE = 32;
constexpr int SCW_TAPS = 7;
constexpr int SCW_MAX_PATCH = 3 * ((SCW_TILE - 1) * 3 + 9) * 2 + 64;     // kh <= 3 rows, sw <= 3, kw <= 9, Cin <= 2

// weights (Co, Cin, kh, kw) -> shared [k = (s*kw + t)*CIN + ci][co]
template <int CIN>
F2G_SIMT_DEV void sc_load_weights(const float* __restrict__ w, int kh, int kw, float* sw_) {
  const int taps = kh * kw;
  for (int i = threadIdx.x; i < taps * CIN * SC_CO; i += blockDim.x) {
    const int co = i % SC_CO, k = i / SC_CO;
    const int ci = k % CIN, tap = k / CIN;
    sw_[i] = w[((long long)co * CIN + ci) * taps + tap];
  }
}

// Forward.  Block = 64 threads = 8 channel quads x 8 pixel groups and owns a tile of 32 consecutive output
// pixels of one output row: the zero-padded input patch is staged in shared memory once, a thread then
// produces 4 pixels x 4 channels (one LDS.128 of weights + 4 LDS.32 of taps per 16 FMA, no bounds tests).
constexpr int SCF_THREADS = 64;
template <int CIN>
F2G_KERNEL void conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bias, SmallConv g, float* __restrict__ y,
                                      int tiles_per_row, long long n_tiles) {
  __shared__ float sw_[SC_MAX_K * SC_CO];
  __shared__ float s_x[SCW_MAX_PATCH];
  sc_load_weights<CIN>(w, g.kh, g.kw, sw_);
  const int q = threadIdx.x & 7, pg = threadIdx.x >> 3;           // pixels pg, pg + 8, pg + 16, pg + 24 of the tile
  const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * q);
  const int patch_w = (SCW_TILE - 1) * g.sw + g.kw;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_per_row);
    const long long row = tile / tiles_per_row;
    const int ho = (int)(row % g.Ho), n = (int)(row / g.Ho);
    const int wo0 = tx * SCW_TILE;
    __syncthreads();
    {
      const float* xb = x + n * g.pitch_n;
      const int h0 = ho * g.sh - g.ph, w0 = wo0 * g.sw - g.pw;
      for (int i = threadIdx.x; i < g.kh * patch_w * CIN; i += SCF_THREADS) {
        const int ci = i % CIN, col = (i / CIN) % patch_w, rr = (i / CIN) / patch_w;
        const int h = h0 + rr, wi = w0 + col;
        s_x[i] = (h >= 0 && h < g.H && wi >= 0 && wi < g.W) ? xb[h * g.pitch_h + wi * g.pitch_w + ci] : 0.f;
      }
    }
    __syncthreads();
    float4 acc[4] = {b4, b4, b4, b4};
    const int jstep = 8 * g.sw * CIN;
    for (int ss = 0; ss < g.kh; ++ss) {
      const float* xrow = s_x + (ss * patch_w + pg * g.sw) * CIN;
      const float* wrow = sw_ + ss * g.kw * CIN * SC_CO + 4 * q;
#pragma unroll 3
      for (int tt = 0; tt < g.kw; ++tt) {
        const float* xp = xrow + tt * CIN;
        const float* wp = wrow + tt * CIN * SC_CO;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + ci * SC_CO);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xv = xp[j * jstep + ci];
            acc[j].x = fmaf(xv, w4.x, acc[j].x); acc[j].y = fmaf(xv, w4.y, acc[j].y);
            acc[j].z = fmaf(xv, w4.z, acc[j].z); acc[j].w = fmaf(xv, w4.w, acc[j].w);
          }
        }
      }
    }
    float* yrow = y + (row * g.Wo + wo0) * SC_CO + 4 * q;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int px = pg + 8 * j;
      if (wo0 + px >= g.Wo) continue;
      float4 a = acc[j];
      if (g.leaky >= 0.f) {
        a.x = a.x > 0.f ? a.x : a.x * g.leaky; a.y = a.y > 0.f ? a.y : a.y * g.leaky;
        a.z = a.z > 0.f ? a.z : a.z * g.leaky; a.w = a.w > 0.f ? a.w : a.w * g.leaky;
      }
      *reinterpret_cast<float4*>(yrow + px * SC_CO) = a;
    }
  }
}

// dz = dy * act'(y): LeakyReLU passes slope where the OUTPUT is <= 0 (same sign as the pre-activation)
F2G_SIMT_DEV float4 sc_dz(const float* __restrict__ dy, const float* __restrict__ y, long long off, float leaky) {
  float4 d = *reinterpret_cast<const float4*>(dy + off);
  if (leaky >= 0.f) {
    const float4 o = *reinterpret_cast<const float4*>(y + off);
    d.x *= o.x > 0.f ? 1.f : leaky; d.y *= o.y > 0.f ? 1.f : leaky;
    d.z *= o.z > 0.f ? 1.f : leaky; d.w *= o.w > 0.f ? 1.f : leaky;
  }
  return d;
}

// Weight + bias gradient.  Block = 64 threads = 8 channel quads x 8 tap groups; a thread owns 7 taps
// (7 * 8 = 56 >= K) of its 4 channels = 28 accumulators.  The block walks over tiles of 32 consecutive
// output pixels of one output row: dz = dy * act'(y) (32 x 32) and the zero-padded input patch the tile's
// taps touch are staged in shared memory once, then every pixel costs one LDS.128 + 7 LDS.32 + 28 FMA per
// thread with no bounds tests.  Partial sums go to part[block][K*32 + 32] (no atomics: the reduction
// kernel below adds the blocks in a fixed order, so the gradient is bit-reproducible).

template <int CIN>
F2G_KERNEL void conv_small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                        const float* __restrict__ y, SmallConv g, float* __restrict__ part,
                                        int tiles_per_row, long long n_tiles) {
  __shared__ float s_dz[SCW_TILE * SC_CO];
  __shared__ float s_x[SCW_MAX_PATCH];
  const int K = g.kh * g.kw * CIN;
  const int q = threadIdx.x & 7, kg = threadIdx.x >> 3;
  const int patch_w = (SCW_TILE - 1) * g.sw + g.kw;              // input columns a tile's taps touch
  int off[SCW_TAPS];                                              // patch offset of each owned tap (pixel 0)
  bool on[SCW_TAPS];
#pragma unroll
  for (int j = 0; j < SCW_TAPS; ++j) {
    const int k = kg * SCW_TAPS + j;
    on[j] = k < K;
    const int ci = k % CIN, tap = k / CIN;
    const int tt = tap % g.kw, ss = tap / g.kw;
    off[j] = on[j] ? (ss * patch_w + tt) * CIN + ci : 0;
  }
  float acc[SCW_TAPS][4];
#pragma unroll
  for (int j = 0; j < SCW_TAPS; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  float4 ab = make_float4(0.f, 0.f, 0.f, 0.f);
  const int step = SCW_TILE > 0 ? g.sw * CIN : 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_per_row);
    const long long row = tile / tiles_per_row;                   // (n, ho)
    const int ho = (int)(row % g.Ho), n = (int)(row / g.Ho);
    const int wo0 = tx * SCW_TILE;
    const int npx = g.Wo - wo0 < SCW_TILE ? g.Wo - wo0 : SCW_TILE;
    __syncthreads();                                              // previous tile fully consumed
    {   // dz tile: pixels beyond the row end contribute zeros
      const long long m0 = (row * g.Wo + wo0) * SC_CO;
      for (int i = threadIdx.x; i < SCW_TILE * SC_CO / 4; i += SCW_THREADS) {
        const int px = i >> 3;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (px < npx) d = sc_dz(dy, y, m0 + (long long)i * 4, g.leaky);
        *reinterpret_cast<float4*>(s_dz + i * 4) = d;
      }
      // input patch with the zero padding materialised
      const float* xb = x + n * g.pitch_n;
      const int h0 = ho * g.sh - g.ph, w0 = wo0 * g.sw - g.pw;
      for (int i = threadIdx.x; i < g.kh * patch_w * CIN; i += SCW_THREADS) {
        const int ci = i % CIN, col = (i / CIN) % patch_w, rr = (i / CIN) / patch_w;
        const int h = h0 + rr, wi = w0 + col;
        s_x[i] = (h >= 0 && h < g.H && wi >= 0 && wi < g.W) ? xb[h * g.pitch_h + wi * g.pitch_w + ci] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < SCW_TILE; ++px) {
      const float4 d = *reinterpret_cast<const float4*>(s_dz + px * SC_CO + 4 * q);
      if (kg == 0) { ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w; }
      const float* xp = s_x + px * step;
#pragma unroll
      for (int j = 0; j < SCW_TAPS; ++j) {
        const float xv = xp[off[j]];
        acc[j][0] = fmaf(xv, d.x, acc[j][0]); acc[j][1] = fmaf(xv, d.y, acc[j][1]);
        acc[j][2] = fmaf(xv, d.z, acc[j][2]); acc[j][3] = fmaf(xv, d.w, acc[j][3]);
      }
    }
  }
  float* o = part + (long long)blockIdx.x * (K * SC_CO + SC_CO);
#pragma unroll
  for (int j = 0; j < SCW_TAPS; ++j)
    if (on[j])
      *reinterpret_cast<float4*>(o + (kg * SCW_TAPS + j) * SC_CO + 4 * q) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
  if (kg == 0) *reinterpret_cast<float4*>(o + K * SC_CO + 4 * q) = ab;
}

// gw[i] += sum_b part[b][i]  (i < K*32: packed weight gradient, then 32 bias sums).  Block = 32 values x 8
// slices of the partial records; the slices are added in a fixed order (bit-reproducible).
F2G_KERNEL void conv_small_wgrad_reduce_kernel(const float* __restrict__ part, int n_blocks, int n_vals,
                                               float* __restrict__ gw, float* __restrict__ gb, int n_w) {
  __shared__ float red[8][32];
  const int v = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + v;
  float s = 0.f;
  if (i < n_vals)
    for (int b = sl; b < n_blocks; b += 8) s += part[(long long)b * n_vals + i];
  red[sl][v] = s;
  __syncthreads();
  if (sl == 0 && i < n_vals) {
    float t = red[0][v];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][v];
    if (i < n_w) gw[i] += t;
    else if (gb) gb[i - n_w] += t;
  }
}

// Input gradient (needed only when the waveform itself carries a gradient: the fake half of the G phase).
// Block = 256 threads = 32 consecutive input pixels of one input row x 8 channel quads.  dz = dy * act'(y) of
// the (<= 3) output rows x (<= 41) output columns whose taps read the tile is staged in shared memory once;
// every pixel then gathers its taps from there (LDS only), 8-lane butterfly at the end.
constexpr int SCD_COLS = 41;
template <int CIN>
F2G_KERNEL void conv_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                        const float* __restrict__ w, SmallConv g, float* __restrict__ dx,
                                        int tiles_per_row, long long n_tiles) {
  __shared__ float sw_[SC_MAX_K * SC_CO];
  __shared__ float s_dz[3 * SCD_COLS * SC_CO];
  __shared__ int s_ho[3];
  sc_load_weights<CIN>(w, g.kh, g.kw, sw_);
  const int q = threadIdx.x & 7, px = threadIdx.x >> 3;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_per_row);
    const long long row = tile / tiles_per_row;                  // (n, h)
    const int h = (int)(row % g.H), n = (int)(row / g.H);
    const int wi0 = tx * SCW_TILE;
    // output columns whose window can touch input columns [wi0, wi0 + 32)
    int lo_num = wi0 + g.pw - (g.kw - 1);
    int wo_lo = lo_num <= 0 ? 0 : (lo_num + g.sw - 1) / g.sw;
    int wo_hi = (wi0 + SCW_TILE - 1 + g.pw) / g.sw;
    if (wo_hi > g.Wo - 1) wo_hi = g.Wo - 1;
    const int ncols = wo_hi - wo_lo + 1;                          // <= SCD_COLS (kw <= 9, sw >= 1)
    __syncthreads();
    if (threadIdx.x < 3) {
      const int ss = threadIdx.x;
      const int hn = h + g.ph - ss;
      int ho = -1;
      if (ss < g.kh && hn >= 0 && hn % g.sh == 0 && hn / g.sh < g.Ho) ho = hn / g.sh;
      s_ho[ss] = ho;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.kh * ncols * 8; i += SC_THREADS) {
      const int qq = i & 7, c = (i >> 3) % ncols, ss = (i >> 3) / ncols;
      const int ho = s_ho[ss];
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ho >= 0) d = sc_dz(dy, y, (((long long)n * g.Ho + ho) * g.Wo + wo_lo + c) * SC_CO + 4 * qq, g.leaky);
      *reinterpret_cast<float4*>(s_dz + (ss * SCD_COLS + c) * SC_CO + 4 * qq) = d;
    }
    __syncthreads();
    const int wi = wi0 + px;
    float acc[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc[ci] = 0.f;
    if (wi < g.W) {
      for (int ss = 0; ss < g.kh; ++ss) {
        if (s_ho[ss] < 0) continue;
        for (int tt = 0; tt < g.kw; ++tt) {
          const int wn = wi + g.pw - tt;
          if (wn < 0) continue;
          int wo = wn;
          if (g.sw != 1) {                       // strided conv (DiscriminatorP): only every sw-th tap lands on an output
            if (wn % g.sw != 0) continue;
            wo = wn / g.sw;
          }
          if (wo > wo_hi) continue;
          const float4 d = *reinterpret_cast<const float4*>(s_dz + (ss * SCD_COLS + wo - wo_lo) * SC_CO + 4 * q);
          const float* wp = sw_ + ((ss * g.kw + tt) * CIN) * SC_CO + 4 * q;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + ci * SC_CO);
            acc[ci] = fmaf(d.x, w4.x, fmaf(d.y, w4.y, fmaf(d.z, w4.z, fmaf(d.w, w4.w, acc[ci]))));
          }
        }
      }
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float v = acc[ci];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (q == 0 && wi < g.W) dx[(row * g.W + wi) * CIN + ci] = v;
    }
  }
}

}  // namespace f2g

using namespace f2g;

static int fill_geom(ConvGeom& g, const F2GConv2d* p) {
  g.Nb = p->Nb; g.H = p->H; g.W = p->W; g.C = p->C;
  g.pitch_h = p->pitch_h; g.pitch_n = p->pitch_n;
  g.kh = p->kh; g.kw = p->kw; g.sh = p->sh; g.sw = p->sw; g.ph = p->ph; g.pw = p->pw;
  g.Ho = (p->H + 2 * p->ph - p->kh) / p->sh + 1;
  g.Wo = (p->W + 2 * p->pw - p->kw) / p->sw + 1;
  g.ldk = p->ldk;
  if (g.Ho < 1 || g.Wo < 1 || g.ldk < g.kh * g.kw * g.C || (g.ldk & 3)) {
    set_error("conv2d geometry invalid (Ho=%d Wo=%d ldk=%d K=%d)", g.Ho, g.Wo, g.ldk, g.kh * g.kw * g.C);
    return F2G_EINVAL;
  }
  return 0;
}

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 148LL * 32;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

// Weights of the phase-decomposed transposed convolution (input gradient of a stride-(1, sw)
// windowed conv, convwin.py): for phase p in [0, sw) with ntp = ceil((kw - p) / sw) taps,
//   out_p[ci][s][jj][co] = W[co][ci][s][p + sw * (ntp - 1 - jj)]   (co < Co, else 0), TF32-rounded,
// all phases in one launch, phase blocks back to back (block p starts at Ci*kh*Cop*sum_{q<p} ntq).
__global__ void conv_w_pack_dgrad_kernel(const float* __restrict__ w, int Co, int Ci, int kh, int kw, int sw,
                                         int Cop, float* __restrict__ out) {
  const long long total = (long long)Ci * kh * kw * Cop;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long rem = i;
    int p = 0, ntp = 0;
    for (; p < sw; ++p) {
      ntp = (kw - p + sw - 1) / sw;
      const long long blk = (long long)Ci * kh * ntp * Cop;
      if (rem < blk) break;
      rem -= blk;
    }
    const int co = (int)(rem % Cop);
    rem /= Cop;
    const int jj = (int)(rem % ntp);
    rem /= ntp;
    const int s = (int)(rem % kh);
    const int ci = (int)(rem / kh);
    float v = 0.f;
    if (co < Co) v = w[(((long long)co * Ci + ci) * kh + s) * kw + p + sw * (ntp - 1 - jj)];
    out[i] = tf32_rna(v);
  }
}

static int sc_fill(SmallConv& g, int Nb, int H, int W, long long pitch_n, long long pitch_h, long long pitch_w,
                   int kh, int kw, int sh, int sw, int ph, int pw, int Cin, int Co, float leaky, const char* who) {
  g.Nb = Nb; g.H = H; g.W = W; g.pitch_n = pitch_n; g.pitch_h = pitch_h; g.pitch_w = pitch_w;
  g.kh = kh; g.kw = kw; g.sh = sh; g.sw = sw; g.ph = ph; g.pw = pw; g.leaky = leaky;
  g.Ho = (H + 2 * ph - kh) / sh + 1;
  g.Wo = (W + 2 * pw - kw) / sw + 1;
  if ((Cin != 1 && Cin != 2) || Co != SC_CO || kh * kw * Cin > SC_MAX_K || g.Ho < 1 || g.Wo < 1 || sh < 1 || sw < 1) {
    set_error("%s: needs Cin in {1, 2}, Cout = 32, kh*kw*Cin <= 64 (Cin=%d Cout=%d kh=%d kw=%d)", who, Cin, Co, kh, kw);
    return F2G_EINVAL;
  }
  return 0;
}

extern "C" int f2g_conv_small_fwd(const float* x, int Nb, int H, int W, int Cin, long long pitch_n, long long pitch_h,
                                  long long pitch_w, const float* w, const float* bias, int Co, int kh, int kw,
                                  int sh, int sw, int ph, int pw, float leaky, float* y, void* stream) {
  SmallConv g;
  if (int rc = sc_fill(g, Nb, H, W, pitch_n, pitch_h, pitch_w, kh, kw, sh, sw, ph, pw, Cin, Co, leaky, "f2g_conv_small_fwd"))
    return rc;
  if (kh > 3 || sw > 3 || kw > 9) {
    set_error("f2g_conv_small_fwd: needs kh <= 3, kw <= 9, sw <= 3 (kh=%d kw=%d sw=%d)", kh, kw, sw);
    return F2G_EINVAL;
  }
  const int tiles_per_row = (g.Wo + SCW_TILE - 1) / SCW_TILE;
  const long long n_tiles = (long long)Nb * g.Ho * tiles_per_row;
  const int grid = (int)(n_tiles < 148 * 32 ? n_tiles : 148 * 32);
  if (Cin == 1)
    F2G_LAUNCH_COOP(conv_small_fwd_kernel<1>, grid, SCF_THREADS, static_cast<cudaStream_t>(stream), x, w, bias, g, y, tiles_per_row, n_tiles);
  else
    F2G_LAUNCH_COOP(conv_small_fwd_kernel<2>, grid, SCF_THREADS, static_cast<cudaStream_t>(stream), x, w, bias, g, y, tiles_per_row, n_tiles);
  return check_launch("f2g_conv_small_fwd");
}

extern "C" int f2g_conv_small_bwd(const float* x, int Nb, int H, int W, int Cin, long long pitch_n, long long pitch_h,
                                  long long pitch_w, const float* w, int Co, int kh, int kw, int sh, int sw, int ph,
                                  int pw, float leaky, const float* dy, const float* y, float* gw_packed, float* gb,
                                  float* dx, float* scratch, long long scratch_floats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SmallConv g;
  if (int rc = sc_fill(g, Nb, H, W, pitch_n, pitch_h, pitch_w, kh, kw, sh, sw, ph, pw, Cin, Co, leaky, "f2g_conv_small_bwd"))
    return rc;
  if (gw_packed) {        // caller zeroed gw_packed (kh*kw*Cin x 32) and gb (32); scratch: see the header
    const int K = kh * kw * Cin, n_vals = K * SC_CO + SC_CO;
    if (kh > 3 || sw > 3 || kw > 9 || !scratch || scratch_floats < (long long)n_vals) {
      set_error("f2g_conv_small_bwd: weight gradient needs kh <= 3, sw <= 3, kw <= 9 and a scratch buffer of >= %d floats",
                n_vals);
      return F2G_EINVAL;
    }
    const int tiles_per_row = (g.Wo + SCW_TILE - 1) / SCW_TILE;
    const long long n_tiles = (long long)Nb * g.Ho * tiles_per_row;
    long long blocks = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
    if (blocks > scratch_floats / n_vals) blocks = scratch_floats / n_vals;
    if (Cin == 1)
      F2G_LAUNCH_COOP(conv_small_wgrad_kernel<1>, (int)blocks, SCW_THREADS, stream, x, dy, y, g, scratch, tiles_per_row, n_tiles);
    else
      F2G_LAUNCH_COOP(conv_small_wgrad_kernel<2>, (int)blocks, SCW_THREADS, stream, x, dy, y, g, scratch, tiles_per_row, n_tiles);
    if (int rc = check_launch("f2g_conv_small_bwd(wgrad)")) return rc;
    F2G_LAUNCH_COOP(conv_small_wgrad_reduce_kernel, (n_vals + 31) / 32, 256, stream, scratch, (int)blocks, n_vals, gw_packed,
                    gb, K * SC_CO);
    if (int rc = check_launch("f2g_conv_small_bwd(wgrad reduce)")) return rc;
  }
  if (dx) {
    if (kh > 3 || kw > 9) {
      set_error("f2g_conv_small_bwd: input gradient needs kh <= 3, kw <= 9 (kh=%d kw=%d)", kh, kw);
      return F2G_EINVAL;
    }
    const int tiles_per_row = (W + SCW_TILE - 1) / SCW_TILE;
    const long long n_tiles = (long long)Nb * H * tiles_per_row;
    const int grid = (int)(n_tiles < 148 * 16 ? n_tiles : 148 * 16);
    if (Cin == 1) F2G_LAUNCH_COOP(conv_small_dgrad_kernel<1>, grid, SC_THREADS, stream, dy, y, w, g, dx, tiles_per_row, n_tiles);
    else F2G_LAUNCH_COOP(conv_small_dgrad_kernel<2>, grid, SC_THREADS, stream, dy, y, w, g, dx, tiles_per_row, n_tiles);
    if (int rc = check_launch("f2g_conv_small_bwd(dgrad)")) return rc;
  }
  return F2G_OK;
}

extern "C" int f2g_im2col2d(const float* x, const F2GConv2d* p, float* col, int round_tf32, void* stream) {
  ConvGeom g;
  if (int rc = fill_geom(g, p)) return rc;
  const long long M = (long long)g.Nb * g.Ho * g.Wo;
  const bool vec = (g.C % 4 == 0) && (g.pitch_h % 4 == 0) && (g.pitch_n % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(col) & 15) == 0);
  const unsigned blocks = (unsigned)((M + 7) / 8);
  if (vec) F2G_LAUNCH(im2col2d_kernel<true>, blocks, 256, static_cast<cudaStream_t>(stream), x, g, col, round_tf32);
  else F2G_LAUNCH(im2col2d_kernel<false>, blocks, 256, static_cast<cudaStream_t>(stream), x, g, col, round_tf32);
  return check_launch("f2g_im2col2d");
}

extern "C" int f2g_col2im2d(const float* dcol, const F2GConv2d* p, float* dx, int accumulate, void* stream) {
  ConvGeom g;
  if (int rc = fill_geom(g, p)) return rc;
  const bool vec = (g.C % 4 == 0) && (g.pitch_h % 4 == 0) && (g.pitch_n % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(dx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dcol) & 15) == 0);
  const long long total = (long long)g.Nb * g.H * g.W * (vec ? g.C / 4 : g.C);
  if (vec) F2G_LAUNCH(col2im2d_kernel<true>, grid_for(total), 256, static_cast<cudaStream_t>(stream), dcol, g, dx, accumulate);
  else F2G_LAUNCH(col2im2d_kernel<false>, grid_for(total), 256, static_cast<cudaStream_t>(stream), dcol, g, dx, accumulate);
  return check_launch("f2g_col2im2d");
}

extern "C" int f2g_pad2d(const float* x, int Nb, int H, int W, int C, long long pitch_n, long long pitch_h,
                         long long pitch_w, int Hl, int Wp, int ph, int pw, long long slack, int round_tf32,
                         float* out, void* stream) {
  if ((C & 3) || (pitch_n & 3) || (pitch_h & 3) || (pitch_w & 3) || (slack & 3) || slack < 0 ||
      (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || Hl < H + ph ||
      Wp < W + pw) {
    set_error("f2g_pad2d: need C, pitches, slack multiples of 4, 16B-aligned pointers, Hl >= H+ph, Wp >= W+pw "
              "(C=%d Hl=%d H=%d ph=%d Wp=%d W=%d pw=%d)", C, Hl, H, ph, Wp, W, pw);
    return F2G_EINVAL;
  }
  const long long body4 = (long long)Nb * Hl * Wp * (C >> 2);
  const long long total4 = body4 + (slack >> 2);
  F2G_LAUNCH(pad2d_kernel, grid_for(total4), 256, static_cast<cudaStream_t>(stream),  x, Nb, H, W, C >> 2, pitch_n, pitch_h, pitch_w, Hl, Wp, ph, pw, total4, body4, round_tf32, out);
  return check_launch("f2g_pad2d");
}

extern "C" int f2g_conv_w_pack(const float* src, int Co, int Ci, int taps, int Co_pad, int ld, float* dst,
                               int dir, void* stream) {
  if (ld < taps * Ci || (ld & 3) || Co_pad < Co) {
    set_error("f2g_conv_w_pack: bad ld=%d (need >= %d, multiple of 4)", ld, taps * Ci);
    return F2G_EINVAL;
  }
  const long long total = dir == 0 ? (long long)Co_pad * ld : (long long)Co * Ci * taps;
  F2G_LAUNCH(conv_w_pack_kernel, grid_for(total), 256, static_cast<cudaStream_t>(stream), src, Co, Ci, taps, Co_pad, ld, dst, dir);
  return check_launch("f2g_conv_w_pack");
}

extern "C" int f2g_conv_w_pack_dgrad(const float* w, int Co, int Ci, int kh, int kw, int sw, int Cop, float* out,
                                     void* stream) {
  if (sw < 1 || kw < sw || Cop < Co || kh < 1) {
    set_error("f2g_conv_w_pack_dgrad: bad geometry (kw=%d sw=%d Co=%d Cop=%d)", kw, sw, Co, Cop);
    return F2G_EINVAL;
  }
  const long long total = (long long)Ci * kh * kw * Cop;
  F2G_LAUNCH(conv_w_pack_dgrad_kernel, grid_for(total), 256, static_cast<cudaStream_t>(stream), w, Co, Ci, kh, kw, sw, Cop, out);
  return check_launch("f2g_conv_w_pack_dgrad");
}
