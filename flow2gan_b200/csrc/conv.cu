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

template <int CIN>
F2G_KERNEL void conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bias, SmallConv g, float* __restrict__ y) {
  __shared__ float sw_[SC_MAX_K * SC_CO];
  sc_load_weights<CIN>(w, g.kh, g.kw, sw_);
  __syncthreads();
  const long long M = (long long)g.Nb * g.Ho * g.Wo;
  const int q = threadIdx.x & 7;
  const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * q);
  for (long long p = (long long)blockIdx.x * (SC_THREADS / 8) + (threadIdx.x >> 3); p < M;
       p += (long long)gridDim.x * (SC_THREADS / 8)) {
    const int wo = (int)(p % g.Wo);
    const long long t = p / g.Wo;
    const int ho = (int)(t % g.Ho), n = (int)(t / g.Ho);
    const float* xb = x + n * g.pitch_n;
    float4 acc = b4;
    for (int s = 0; s < g.kh; ++s) {
      const int h = ho * g.sh - g.ph + s;
      if (h < 0 || h >= g.H) continue;
      for (int tt = 0; tt < g.kw; ++tt) {
        const int wi = wo * g.sw - g.pw + tt;
        if (wi < 0 || wi >= g.W) continue;
        const float* xp = xb + h * g.pitch_h + wi * g.pitch_w;
        const float* wp = sw_ + ((s * g.kw + tt) * CIN) * SC_CO + 4 * q;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float xv = xp[ci];
          const float4 w4 = *reinterpret_cast<const float4*>(wp + ci * SC_CO);
          acc.x = fmaf(xv, w4.x, acc.x); acc.y = fmaf(xv, w4.y, acc.y);
          acc.z = fmaf(xv, w4.z, acc.z); acc.w = fmaf(xv, w4.w, acc.w);
        }
      }
    }
    if (g.leaky >= 0.f) {
      acc.x = acc.x > 0.f ? acc.x : acc.x * g.leaky; acc.y = acc.y > 0.f ? acc.y : acc.y * g.leaky;
      acc.z = acc.z > 0.f ? acc.z : acc.z * g.leaky; acc.w = acc.w > 0.f ? acc.w : acc.w * g.leaky;
    }
    *reinterpret_cast<float4*>(y + p * SC_CO + 4 * q) = acc;
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

// Weight + bias gradient.  Thread = (tap group kq in [0, 32), channel quad q): it owns taps kq and kq+32 for
// its 4 channels and walks over the block's share of output pixels; one atomicAdd per owned value at the
// end (grid = a few hundred blocks).  gw is the PACKED layout [k][co]; gb[co].
template <int CIN>
F2G_KERNEL void conv_small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                        const float* __restrict__ y, SmallConv g, float* __restrict__ gw,
                                        float* __restrict__ gb) {
  const int K = g.kh * g.kw * CIN;
  const int q = threadIdx.x & 7, kq = threadIdx.x >> 3;
  const long long M = (long long)g.Nb * g.Ho * g.Wo;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, ab = a0;
  int s0 = 0, t0 = 0, c0 = 0, s1 = 0, t1 = 0, c1 = 0;
  const bool on0 = kq < K, on1 = kq + 32 < K;
  if (on0) { c0 = kq % CIN; t0 = (kq / CIN) % g.kw; s0 = (kq / CIN) / g.kw; }
  if (on1) { c1 = (kq + 32) % CIN; t1 = ((kq + 32) / CIN) % g.kw; s1 = ((kq + 32) / CIN) / g.kw; }
  const long long per = (M + gridDim.x - 1) / gridDim.x;
  const long long p_begin = (long long)blockIdx.x * per, p_end = p_begin + per < M ? p_begin + per : M;
  int wo = 0, ho = 0, n = 0;
  if (p_begin < M) {
    wo = (int)(p_begin % g.Wo);
    const long long t = p_begin / g.Wo;
    ho = (int)(t % g.Ho); n = (int)(t / g.Ho);
  }
  for (long long p = p_begin; p < p_end; ++p) {
    const float4 d = sc_dz(dy, y, p * SC_CO + 4 * q, g.leaky);
    if (kq == 0) { ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w; }
    const float* xb = x + n * g.pitch_n;
    if (on0) {
      const int h = ho * g.sh - g.ph + s0, wi = wo * g.sw - g.pw + t0;
      if (h >= 0 && h < g.H && wi >= 0 && wi < g.W) {
        const float xv = xb[h * g.pitch_h + wi * g.pitch_w + c0];
        a0.x = fmaf(xv, d.x, a0.x); a0.y = fmaf(xv, d.y, a0.y); a0.z = fmaf(xv, d.z, a0.z); a0.w = fmaf(xv, d.w, a0.w);
      }
    }
    if (on1) {
      const int h = ho * g.sh - g.ph + s1, wi = wo * g.sw - g.pw + t1;
      if (h >= 0 && h < g.H && wi >= 0 && wi < g.W) {
        const float xv = xb[h * g.pitch_h + wi * g.pitch_w + c1];
        a1.x = fmaf(xv, d.x, a1.x); a1.y = fmaf(xv, d.y, a1.y); a1.z = fmaf(xv, d.z, a1.z); a1.w = fmaf(xv, d.w, a1.w);
      }
    }
    if (++wo == g.Wo) { wo = 0; if (++ho == g.Ho) { ho = 0; ++n; } }
  }
  if (on0) {
    float* o = gw + kq * SC_CO + 4 * q;
    atomicAdd(o, a0.x); atomicAdd(o + 1, a0.y); atomicAdd(o + 2, a0.z); atomicAdd(o + 3, a0.w);
  }
  if (on1) {
    float* o = gw + (kq + 32) * SC_CO + 4 * q;
    atomicAdd(o, a1.x); atomicAdd(o + 1, a1.y); atomicAdd(o + 2, a1.z); atomicAdd(o + 3, a1.w);
  }
  if (kq == 0 && gb) {
    float* o = gb + 4 * q;
    atomicAdd(o, ab.x); atomicAdd(o + 1, ab.y); atomicAdd(o + 2, ab.z); atomicAdd(o + 3, ab.w);
  }
}

// Input gradient (needed only when the waveform itself carries a gradient: the fake half of the G phase).
// Thread = (input pixel, channel quad): gathers every tap that read the pixel, 8-lane butterfly at the end.
template <int CIN>
F2G_KERNEL void conv_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                        const float* __restrict__ w, SmallConv g, float* __restrict__ dx) {
  __shared__ float sw_[SC_MAX_K * SC_CO];
  sc_load_weights<CIN>(w, g.kh, g.kw, sw_);
  __syncthreads();
  const long long P = (long long)g.Nb * g.H * g.W;
  const int q = threadIdx.x & 7;
  const long long span = (long long)gridDim.x * (SC_THREADS / 8);
  const long long rounds = (P + span - 1) / span;          // every lane runs the same number of rounds (shuffles)
  for (long long r = 0; r < rounds; ++r) {
    const long long p = r * span + (long long)blockIdx.x * (SC_THREADS / 8) + (threadIdx.x >> 3);
    float acc[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc[ci] = 0.f;
    int wi = 0, h = 0, n = 0;
    if (p < P) {
      wi = (int)(p % g.W);
      const long long t = p / g.W;
      h = (int)(t % g.H); n = (int)(t / g.H);
      for (int s = 0; s < g.kh; ++s) {
        const int hn = h + g.ph - s;
        if (hn < 0 || hn % g.sh != 0) continue;
        const int ho = hn / g.sh;
        if (ho >= g.Ho) continue;
        for (int tt = 0; tt < g.kw; ++tt) {
          const int wn = wi + g.pw - tt;
          if (wn < 0 || wn % g.sw != 0) continue;
          const int wo = wn / g.sw;
          if (wo >= g.Wo) continue;
          const long long m = ((long long)n * g.Ho + ho) * g.Wo + wo;
          const float4 d = sc_dz(dy, y, m * SC_CO + 4 * q, g.leaky);
          const float* wp = sw_ + ((s * g.kw + tt) * CIN) * SC_CO + 4 * q;
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
      if (q == 0 && p < P) dx[p * CIN + ci] = v;
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

static int sc_grid(long long pixels) {
  long long b = (pixels + SC_THREADS / 8 - 1) / (SC_THREADS / 8);
  const long long cap = 148LL * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

extern "C" int f2g_conv_small_fwd(const float* x, int Nb, int H, int W, int Cin, long long pitch_n, long long pitch_h,
                                  long long pitch_w, const float* w, const float* bias, int Co, int kh, int kw,
                                  int sh, int sw, int ph, int pw, float leaky, float* y, void* stream) {
  SmallConv g;
  if (int rc = sc_fill(g, Nb, H, W, pitch_n, pitch_h, pitch_w, kh, kw, sh, sw, ph, pw, Cin, Co, leaky, "f2g_conv_small_fwd"))
    return rc;
  const int grid = sc_grid((long long)Nb * g.Ho * g.Wo);
  if (Cin == 1) F2G_LAUNCH_COOP(conv_small_fwd_kernel<1>, grid, SC_THREADS, static_cast<cudaStream_t>(stream), x, w, bias, g, y);
  else F2G_LAUNCH_COOP(conv_small_fwd_kernel<2>, grid, SC_THREADS, static_cast<cudaStream_t>(stream), x, w, bias, g, y);
  return check_launch("f2g_conv_small_fwd");
}

extern "C" int f2g_conv_small_bwd(const float* x, int Nb, int H, int W, int Cin, long long pitch_n, long long pitch_h,
                                  long long pitch_w, const float* w, int Co, int kh, int kw, int sh, int sw, int ph,
                                  int pw, float leaky, const float* dy, const float* y, float* gw_packed, float* gb,
                                  float* dx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SmallConv g;
  if (int rc = sc_fill(g, Nb, H, W, pitch_n, pitch_h, pitch_w, kh, kw, sh, sw, ph, pw, Cin, Co, leaky, "f2g_conv_small_bwd"))
    return rc;
  if (gw_packed) {        // caller zeroed gw_packed (kh*kw*Cin x 32) and gb (32)
    const long long M = (long long)Nb * g.Ho * g.Wo;
    long long blocks = M / 512;                        // >= 512 pixels per block: few atomics, long FMA runs
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    if (Cin == 1) F2G_LAUNCH_COOP(conv_small_wgrad_kernel<1>, (int)blocks, SC_THREADS, stream, x, dy, y, g, gw_packed, gb);
    else F2G_LAUNCH_COOP(conv_small_wgrad_kernel<2>, (int)blocks, SC_THREADS, stream, x, dy, y, g, gw_packed, gb);
    if (int rc = check_launch("f2g_conv_small_bwd(wgrad)")) return rc;
  }
  if (dx) {
    const int grid = sc_grid((long long)Nb * H * W);
    if (Cin == 1) F2G_LAUNCH_COOP(conv_small_dgrad_kernel<1>, grid, SC_THREADS, stream, dy, y, w, g, dx);
    else F2G_LAUNCH_COOP(conv_small_dgrad_kernel<2>, grid, SC_THREADS, stream, dy, y, w, g, dx);
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
