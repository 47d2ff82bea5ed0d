// STFT / iSTFT family: framing + periodic-hann window + shared-memory Stockham FFT, fused
// packing / magnitude / filterbank / log epilogues; inverse path = irfft frames + overlap-add
// + envelope divide + branch fusion + Euler update.  All HBM-bound SIMT kernels.
// Reference semantics: torch.stft / torch.istft (center=True, reflect pad, onesided) as called
// from flow2gan/models/modules.py:68-84,105-116 and the torchaudio (Mel)Spectrogram wrappers.
#include "simt.cuh"
#include "../../include/flow2gan_b200.h"
#include "fft_warp.cuh"

#include <string.h>
#include <stdlib.h>

namespace f2g {

F2G_DEVINL float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Radix-2 Stockham autosort FFT over n = 1 << logn complex points held in shared memory.
// tw[k] = exp(+2*pi*i*k/n), k < n/2.  Natural order in, natural order out.  Returns the buffer
// that holds the result.  All threads of the block must call it.
template <bool INVERSE>
F2G_DEVINL float2* block_fft(float2* a, float2* b, const float2* tw, int n, int logn) {
  const int half = n >> 1;
  for (int s = 0; s < logn; ++s) {
    const int ns = 1 << s;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      const int k = j & (ns - 1);
      float2 w = tw[k * (half >> s)];
      if (!INVERSE) w.y = -w.y;
      const float2 v0 = a[j];
      const float2 v1 = cmul(a[j + half], w);
      const int j0 = ((j - k) << 1) + k;
      b[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
      b[j0 + ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
    }
    __syncthreads();
    float2* t = a;
    a = b;
    b = t;
  }
  return a;
}

// (a precomputed global twiddle table was measured: no change -- these kernels are bound by the
// per-frame CTA latency chain, not by the sincospif calls)
F2G_DEVINL void fill_twiddles(float2* tw, int n) {
  for (int k = threadIdx.x; k < (n >> 1); k += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)k / (float)n, &s, &c);
    tw[k] = make_float2(c, s);
  }
}

// periodic hann from the twiddle table: w[i] = 0.5 - 0.5 cos(2 pi i / n)
F2G_DEVINL float hann_from_tw(const float2* tw, int i, int n) {
  const int h = n >> 1;
  return i < h ? 0.5f - 0.5f * tw[i].x : 0.5f + 0.5f * tw[i - h].x;
}

F2G_DEVINL void stft_frame(const float* __restrict__ audio, int T, int ld_audio, int n, int logn,
                           int hop, int frames, int mode, const float* __restrict__ pre,
                           const float* __restrict__ fb, int n_filt, float log_clip,
                           float* __restrict__ out, int ld_out, int round_tf32, int center,
                           int adjoint_scale, const float* __restrict__ row_mask, int row,
                           const int* __restrict__ fb_rng = nullptr) {
  F2G_DYN_SMEM(float2, sm);
  float2* a = sm;
  float2* b = sm + n;
  float2* tw = sm + 2 * n;
  float* spec = reinterpret_cast<float*>(sm + 2 * n + (n >> 1));

  const int bi = row / frames;
  const int f = row - bi * frames;
  const int nb = (n >> 1) + 1;

  fill_twiddles(tw, n);
  __syncthreads();

  float sub = 0.f, mul = 1.f;
  if (pre) {
    sub = pre[2 * bi];
    mul = pre[2 * bi + 1];
  }
  const float* x = audio + (size_t)bi * ld_audio;
  const int start = center ? f * hop - (n >> 1) : f * hop;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int p = start + i;
    if (p < 0) p = -p;
    if (p >= T) p = 2 * (T - 1) - p;
    const float v = (x[p] - sub) * mul;
    a[i] = make_float2(v * hann_from_tw(tw, i, n), 0.f);
  }
  __syncthreads();
  const float2* X = block_fft<false>(a, b, tw, n, logn);

  float* o = out + (size_t)row * ld_out;
  if (mode == F2G_SPEC_PACKED) {
    const float rm = row_mask ? row_mask[row] : 1.f;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
      float2 v = X[k];
      if (adjoint_scale) {   // adjoint of the C2R transform: c_k / n, Im(DC) = Im(Nyquist) = 0
        const bool edge = (k == 0) || (k == (n >> 1));
        const float c = (edge ? 1.f : 2.f) / (float)n;
        v.x *= c;
        v.y = edge ? 0.f : v.y * c;
      }
      v.x *= rm; v.y *= rm;
      o[k] = round_tf32 ? tf32_rna(v.x) : v.x;
      o[nb + k] = round_tf32 ? tf32_rna(v.y) : v.y;
    }
    for (int k = 2 * nb + threadIdx.x; k < ld_out; k += blockDim.x) o[k] = 0.f;
    return;
  }
  if (mode == F2G_SPEC_COMPLEX_BANDS) {   // interleaved (re, im) per bin: (rows, freq, 2) channel-last
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
      const float2 v = X[k];
      o[2 * k] = v.x;
      o[2 * k + 1] = v.y;
    }
    return;
  }
  // magnitude / power, optionally contracted with a filterbank
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    const float2 v = X[k];
    const float p2 = v.x * v.x + v.y * v.y;
    const float m = mode == F2G_SPEC_POWER ? p2 : sqrtf(p2);
    if (fb) spec[k] = m;
    else o[k] = round_tf32 ? tf32_rna(m) : m;
  }
  if (!fb) return;
  __syncthreads();
  // Triangular filterbanks are banded: fb_rng (optional) = per filter m the bin range [lo, hi) outside
  // which fb[:, m] is exactly zero -- same ascending summation, the skipped terms add 0.0
  for (int m = threadIdx.x; m < n_filt; m += blockDim.x) {
    float acc = 0.f;
    const int k0 = fb_rng ? fb_rng[2 * m] : 0, k1 = fb_rng ? fb_rng[2 * m + 1] : nb;
    for (int k = k0; k < k1; ++k) acc = fmaf(spec[k], __ldg(fb + (size_t)k * n_filt + m), acc);
    if (log_clip > 0.f) acc = logf(fmaxf(acc, log_clip));
    o[m] = round_tf32 ? tf32_rna(acc) : acc;
  }
}

__global__ void stft_kernel(const float* __restrict__ audio, int T, int ld_audio, int n, int logn,
                            int hop, int frames, int mode, const float* __restrict__ pre,
                            const float* __restrict__ fb, int n_filt, float log_clip,
                            float* __restrict__ out, int ld_out, int round_tf32, int center,
                            int adjoint_scale, const float* __restrict__ row_mask,
                            const int* __restrict__ fb_rng) {
  stft_frame(audio, T, ld_audio, n, logn, hop, frames, mode, pre, fb, n_filt, log_clip, out, ld_out,
             round_tf32, center, adjoint_scale, row_mask, blockIdx.x, fb_rng);
}

// Up to 4 packed-mode STFTs / inverse transforms of the SAME signal batch in one launch (the three
// branch resolutions of AudioConvNeXt, modules.py:699-719): blockIdx.x ranges select the problem;
// the block size is that of the largest transform (smaller ones leave threads idle).
struct SpecGroupArgs {
  const float* in[4];     // stft: audio (B, ld_in) ; irfft: packed rows
  float* out[4];
  int n[4], logn[4], hop[4], frames[4], ld_in[4], ld_out[4];
  int row_begin[5];
  int np, T, round_tf32;
};

__global__ void stft_group_kernel(const __grid_constant__ SpecGroupArgs g) {
  int pi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < g.np && (int)blockIdx.x >= g.row_begin[i]) pi = i;
  pdl_wait();
  pdl_launch();
  stft_frame(g.in[pi], g.T, g.ld_in[pi], g.n[pi], g.logn[pi], g.hop[pi], g.frames[pi], F2G_SPEC_PACKED,
             nullptr, nullptr, 0, 0.f, g.out[pi], g.ld_out[pi], g.round_tf32, 1, 0, nullptr,
             blockIdx.x - g.row_begin[pi]);
}

F2G_DEVINL void irfft_frame(const float* __restrict__ packed, int ld, int n, int logn,
                            float* __restrict__ frames_out, int row) {
  F2G_DYN_SMEM(float2, sm);
  float2* a = sm;
  float2* b = sm + n;
  float2* tw = sm + 2 * n;
  const int nb = (n >> 1) + 1;
  const float* p = packed + (size_t)row * ld;

  fill_twiddles(tw, n);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    float re, im;
    if (k < nb) {
      re = p[k];
      im = (k == 0 || k == (n >> 1)) ? 0.f : p[nb + k];   // C2R ignores Im(DC), Im(Nyquist)
    } else {
      const int kk = n - k;
      re = p[kk];
      im = -p[nb + kk];
    }
    a[k] = make_float2(re, im);
  }
  __syncthreads();
  const float2* y = block_fft<true>(a, b, tw, n, logn);
  const float inv_n = 1.0f / (float)n;
  float* o = frames_out + (size_t)row * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    o[i] = y[i].x * inv_n * hann_from_tw(tw, i, n);
}

__global__ void irfft_frames_kernel(const float* __restrict__ packed, int ld, int n, int logn,
                                    float* __restrict__ frames_out) {
  irfft_frame(packed, ld, n, logn, frames_out, blockIdx.x);
}

__global__ void irfft_group_kernel(const __grid_constant__ SpecGroupArgs g) {
  int pi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < g.np && (int)blockIdx.x >= g.row_begin[i]) pi = i;
  pdl_wait();
  pdl_launch();
  irfft_frame(g.in[pi], g.ld_in[pi], g.n[pi], g.logn[pi], g.out[pi], blockIdx.x - g.row_begin[pi]);
}

// ---- warp-per-frame versions (n_fft <= 1024): 8 frames per CTA, no block barriers ---------------
constexpr int FFT_WARPS = 8;

__global__ void __launch_bounds__(32 * FFT_WARPS) stft_group_warp_kernel(const __grid_constant__ SpecGroupArgs g) {
  __shared__ float srow[FFT_WARPS][2][FFT_WARP_ROW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_all = blockIdx.x * FFT_WARPS + warp;
  pdl_wait();
  pdl_launch();
  if (row_all >= g.row_begin[g.np]) return;
  int pi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < g.np && row_all >= g.row_begin[i]) pi = i;
  const int row = row_all - g.row_begin[pi];
  const int n = g.n[pi], frames = g.frames[pi];
  const int bi = row / frames, f = row - bi * frames;
  const float* x = g.in[pi] + (size_t)bi * g.ld_in[pi];
  float* o = g.out[pi] + (size_t)row * g.ld_out[pi];
  const int start = f * g.hop[pi] - (n >> 1);
  float* sre = srow[warp][0];
  float* sim = srow[warp][1];
  const bool rnd = g.round_tf32 != 0;
  switch (n) {
    case 128: warp_rfft_packed<2>(x, g.T, start, o, g.ld_out[pi], rnd, sre, sim, lane); break;
    case 256: warp_rfft_packed<4>(x, g.T, start, o, g.ld_out[pi], rnd, sre, sim, lane); break;
    case 512: warp_rfft_packed<8>(x, g.T, start, o, g.ld_out[pi], rnd, sre, sim, lane); break;
    default: warp_rfft_packed<16>(x, g.T, start, o, g.ld_out[pi], rnd, sre, sim, lane); break;
  }
}

__global__ void __launch_bounds__(32 * FFT_WARPS) irfft_group_warp_kernel(const __grid_constant__ SpecGroupArgs g) {
  __shared__ float srow[FFT_WARPS][2][FFT_WARP_ROW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_all = blockIdx.x * FFT_WARPS + warp;
  pdl_wait();
  pdl_launch();
  if (row_all >= g.row_begin[g.np]) return;
  int pi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < g.np && row_all >= g.row_begin[i]) pi = i;
  const int row = row_all - g.row_begin[pi];
  const int n = g.n[pi];
  const float* pk = g.in[pi] + (size_t)row * g.ld_in[pi];
  float* fr = g.out[pi] + (size_t)row * n;
  float* sre = srow[warp][0];
  float* sim = srow[warp][1];
  switch (n) {
    case 128: warp_irfft_frame<2>(pk, fr, sre, sim, lane); break;
    case 256: warp_irfft_frame<4>(pk, fr, sre, sim, lane); break;
    case 512: warp_irfft_frame<8>(pk, fr, sre, sim, lane); break;
    default: warp_irfft_frame<16>(pk, fr, sre, sim, lane); break;
  }
}

struct OlaArgs {
  const float* fr[4];
  int n[4], hop[4], frames[4];
  int nb;
};

__global__ void ola_combine_kernel(OlaArgs a, const float* __restrict__ weight,
                                   const float* __restrict__ x, float* __restrict__ out, int T,
                                   int euler, float t, float dt, int clamp) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int bi = blockIdx.y;
  if (s >= T) return;
  float pred = 0.f;
  for (int j = 0; j < a.nb; ++j) {
    const int n = a.n[j], hop = a.hop[j], F = a.frames[j];
    float v = 0.f;
    if (s < hop * (F - 1)) {
      const int p = s + (n >> 1);
      const int f_hi = min(p / hop, F - 1);
      const int f_lo = p >= n ? (p - n) / hop + 1 : 0;
      float acc = 0.f, env = 0.f;
      for (int f = f_lo; f <= f_hi; ++f) {
        const int i = p - f * hop;
        acc += a.fr[j][((size_t)bi * F + f) * n + i];
        const float w = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)n);
        env += w * w;
      }
      v = acc / env;
    }
    const float wgt = weight ? weight[bi * a.nb + j] : 1.0f / (float)a.nb;
    pred += v * wgt;
  }
  float r = pred;
  if (euler) {
    const float xv = x[(size_t)bi * T + s];
    r = xv + ((pred - xv) / (1.0f - t)) * dt;
  }
  if (clamp) r = fminf(fmaxf(r, -1.0f), 1.0f);
  out[(size_t)bi * T + s] = r;
}

// STFT adjoint, per frame: fr[i] = w[i] * Re( sum_{k<=n/2} (dRe_k + i dIm_k) e^{+2 pi i k i / n} )
__global__ void stft_bwd_frames_kernel(const float* __restrict__ dpacked, int ld, int n, int logn,
                                       float* __restrict__ frames_out, int interleaved) {
  F2G_DYN_SMEM(float2, sm);
  float2* a = sm;
  float2* b = sm + n;
  float2* tw = sm + 2 * n;
  const int row = blockIdx.x;
  const int nb = (n >> 1) + 1;
  const float* p = dpacked + (size_t)row * ld;
  fill_twiddles(tw, n);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (k < nb) v = interleaved ? make_float2(p[2 * k], p[2 * k + 1]) : make_float2(p[k], p[nb + k]);
    a[k] = v;
  }
  __syncthreads();
  const float2* y = block_fft<true>(a, b, tw, n, logn);
  float* o = frames_out + (size_t)row * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = y[i].x * hann_from_tw(tw, i, n);
}

// Fused backward of a filterbank spectrogram loss term (MelSpectrogram / LinearFilterSpectrogram
// + optional safe_log): recompute the frame spectrum, push dF (grad w.r.t. the [log-]filterbank
// output) through log -> fb -> |.| or |.|^2 -> rFFT -> window, emit windowed frame gradients.
__global__ void spec_loss_bwd_kernel(const float* __restrict__ audio, int T, int ld_audio, int n, int logn,
                                     int hop, int frames, int mode, const float* __restrict__ fb,
                                     int n_filt, float log_clip, const float* __restrict__ dF, int ld_dF,
                                     float* __restrict__ frames_out, const int* __restrict__ fb_rng) {
  F2G_DYN_SMEM(float2, sm);
  float2* a = sm;
  float2* b = sm + n;
  float2* tw = sm + 2 * n;
  float* spec = reinterpret_cast<float*>(sm + 2 * n + (n >> 1));   // nb
  float* dfilt = spec + (n >> 1) + 1;                              // n_filt
  const int row = blockIdx.x;
  const int bi = row / frames, f = row - bi * frames;
  const int nb = (n >> 1) + 1;
  fill_twiddles(tw, n);
  __syncthreads();
  const float* x = audio + (size_t)bi * ld_audio;
  const int start = f * hop - (n >> 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int p = start + i;
    if (p < 0) p = -p;
    if (p >= T) p = 2 * (T - 1) - p;
    a[i] = make_float2(x[p] * hann_from_tw(tw, i, n), 0.f);
  }
  __syncthreads();
  float2* X = block_fft<false>(a, b, tw, n, logn);
  float2* other = (X == a) ? b : a;
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    const float2 v = X[k];
    const float p2 = v.x * v.x + v.y * v.y;
    spec[k] = mode == F2G_SPEC_POWER ? p2 : sqrtf(p2);
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_filt; m += blockDim.x) {
    float g = dF[(size_t)row * ld_dF + m];
    if (log_clip > 0.f) {
      float acc = 0.f;
      const int k0 = fb_rng ? fb_rng[2 * m] : 0, k1 = fb_rng ? fb_rng[2 * m + 1] : nb;
      for (int k = k0; k < k1; ++k) acc = fmaf(spec[k], __ldg(fb + (size_t)k * n_filt + m), acc);
      g = acc > log_clip ? g / acc : 0.f;
    }
    dfilt[m] = g;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    float2 o = make_float2(0.f, 0.f);
    if (k < nb) {
      float ds = 0.f;      // per bin k the filters [m0, m1) that touch it (second half of fb_rng)
      const int m0 = fb_rng ? fb_rng[2 * (n_filt + k)] : 0, m1 = fb_rng ? fb_rng[2 * (n_filt + k) + 1] : n_filt;
      for (int m = m0; m < m1; ++m) ds = fmaf(__ldg(fb + (size_t)k * n_filt + m), dfilt[m], ds);
      const float2 v = X[k];
      if (mode == F2G_SPEC_POWER) {
        o = make_float2(2.f * v.x * ds, 2.f * v.y * ds);
      } else {
        const float mag = spec[k];
        if (mag > 0.f) o = make_float2(v.x / mag * ds, v.y / mag * ds);
      }
    }
    other[k] = o;
  }
  __syncthreads();
  const float2* y = block_fft<true>(other, X, tw, n, logn);
  float* o = frames_out + (size_t)row * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = y[i].x * hann_from_tw(tw, i, n);
}

__global__ void dc_peak_kernel(const float* __restrict__ audio, int T, int ld, float* __restrict__ pre) {
  __shared__ float red[32];
  __shared__ float bc;
  const float* x = audio + (size_t)blockIdx.x * ld;
  float s = 0.f;
  for (int i = threadIdx.x; i < T; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) bc = v / (float)T;
  }
  __syncthreads();
  const float mean = bc;
  float m = 0.f;
  for (int i = threadIdx.x; i < T; i += blockDim.x) m = fmaxf(m, fabsf(x[i] - mean));
  m = warp_max(m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_max(v);
    if (threadIdx.x == 0) {
      pre[2 * blockIdx.x] = mean;
      pre[2 * blockIdx.x + 1] = 0.8f / (v + 1e-9f);
    }
  }
}

static int ilog2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return (1 << l) == n ? l : -1;
}

}  // namespace f2g

using namespace f2g;

static int stft_launch(const float* audio, int B, int T, int ld_audio, int n_fft, int hop, int mode,
                       const float* pre, const float* fb, int n_filt, float log_clip, float* out,
                       int ld_out, int round_tf32, int center, int adjoint_scale, const float* row_mask,
                       void* stream, const int* fb_rng = nullptr) {
  const int logn = ilog2_exact(n_fft);
  if (logn < 5 || n_fft > 2048) {
    set_error("f2g_stft: n_fft=%d must be a power of two in [32, 2048]", n_fft);
    return F2G_EINVAL;
  }
  if (center ? (n_fft / 2 >= T) : (n_fft > T)) {
    set_error("f2g_stft: signal too short (n_fft=%d, T=%d, center=%d)", n_fft, T, center);
    return F2G_EINVAL;
  }
  if (mode != F2G_SPEC_PACKED && mode != F2G_SPEC_MAG && mode != F2G_SPEC_POWER &&
      mode != F2G_SPEC_COMPLEX_BANDS) {
    set_error("f2g_stft: bad mode %d", mode);
    return F2G_EINVAL;
  }
  const int frames = center ? 1 + T / hop : 1 + (T - n_fft) / hop;
  const int threads = n_fft / 2 < 32 ? 32 : n_fft / 2;
  const size_t smem = (size_t)(2 * n_fft + n_fft / 2) * sizeof(float2) + (n_fft / 2 + 1) * sizeof(float);
  F2G_LAUNCH_COOP_SMEM(stft_kernel, B * frames, threads, smem, static_cast<cudaStream_t>(stream), audio, T, ld_audio,
                       n_fft, logn, hop, frames, mode, pre, fb, n_filt, log_clip, out, ld_out, round_tf32, center,
                       adjoint_scale, row_mask, fb ? fb_rng : nullptr);
  return check_launch("f2g_stft");
}

extern "C" int f2g_stft(const float* audio, int B, int T, int ld_audio, int n_fft, int hop, int mode,
                        const float* pre, const float* fb, int n_filt, float log_clip, float* out,
                        int ld_out, int round_tf32, const int* fb_ranges, void* stream) {
  return stft_launch(audio, B, T, ld_audio, n_fft, hop, mode, pre, fb, n_filt, log_clip, out, ld_out,
                     round_tf32, 1, 0, nullptr, stream, fb_ranges);
}

static int spec_group_fill(SpecGroupArgs& g, const F2GSpecProblem* probs, int np, const char* who,
                           int* max_n) {
  if (np < 1 || np > 4) {
    set_error("%s: %d problems (1..4)", who, np);
    return F2G_EINVAL;
  }
  memset(&g, 0, sizeof(g));
  g.np = np;
  int rows = 0;
  *max_n = 0;
  for (int i = 0; i < np; ++i) {
    const F2GSpecProblem& p = probs[i];
    const int logn = ilog2_exact(p.n_fft);
    if (logn < 5 || p.n_fft > 2048) {
      set_error("%s: n_fft=%d must be a power of two in [32, 2048]", who, p.n_fft);
      return F2G_EINVAL;
    }
    g.in[i] = p.in; g.out[i] = p.out; g.n[i] = p.n_fft; g.logn[i] = logn; g.hop[i] = p.hop;
    g.frames[i] = p.frames; g.ld_in[i] = p.ld_in; g.ld_out[i] = p.ld_out;
    g.row_begin[i] = rows;
    rows += p.rows;
    if (p.n_fft > *max_n) *max_n = p.n_fft;
  }
  for (int i = np; i <= 4; ++i) g.row_begin[i] = rows;
  return 0;
}

// The warp-per-frame kernels cover n_fft 128..1024 (F2G_FFT_SMEM=1 forces the shared-memory
// Stockham kernels -- A/B testing).
static bool spec_group_warp_ok(const F2GSpecProblem* probs, int np) {
  static const int force_smem = bringup_int("F2G_FFT_SMEM", 0);
  if (force_smem) return false;
  for (int i = 0; i < np; ++i) {
    const int n = probs[i].n_fft;
    if (n != 128 && n != 256 && n != 512 && n != 1024) return false;
    if ((probs[i].ld_out & 1) || (reinterpret_cast<uintptr_t>(probs[i].out) & 7)) return false;
  }
  return true;
}

extern "C" int f2g_stft_group(const F2GSpecProblem* probs, int np, int B, int T, int round_tf32,
                              void* stream) {
  SpecGroupArgs g;
  int max_n = 0;
  if (int rc = spec_group_fill(g, probs, np, "f2g_stft_group", &max_n)) return rc;
  for (int i = 0; i < np; ++i) {
    if (probs[i].n_fft / 2 >= T) {
      set_error("f2g_stft_group: signal too short (n_fft=%d, T=%d)", probs[i].n_fft, T);
      return F2G_EINVAL;
    }
    g.frames[i] = 1 + T / probs[i].hop;
    if (probs[i].rows != B * g.frames[i]) {
      set_error("f2g_stft_group: rows=%d != B*frames=%d", probs[i].rows, B * g.frames[i]);
      return F2G_EINVAL;
    }
  }
  g.T = T;
  g.round_tf32 = round_tf32;
  if (spec_group_warp_ok(probs, np)) {      // warp-per-frame shuffle FFT
    if (int rc = fft_roots_init()) return rc;
    const int rows = g.row_begin[np];
    cudaError_t le = launch_pdl(stft_group_warp_kernel, dim3((rows + FFT_WARPS - 1) / FFT_WARPS),
                                dim3(32 * FFT_WARPS), 0, static_cast<cudaStream_t>(stream), g);
    if (le != cudaSuccess) {
      set_error("f2g_stft_group launch: %s", cudaGetErrorString(le));
      return (int)le;
    }
    return check_launch("f2g_stft_group");
  }
  const int threads = max_n / 2 < 32 ? 32 : max_n / 2;
  const size_t smem = (size_t)(2 * max_n + max_n / 2) * sizeof(float2) + (max_n / 2 + 1) * sizeof(float);
  cudaError_t le = launch_pdl(stft_group_kernel, dim3(g.row_begin[np]), dim3(threads), smem,
                              static_cast<cudaStream_t>(stream), g);
  if (le != cudaSuccess) {
    set_error("f2g_stft_group launch: %s", cudaGetErrorString(le));
    return (int)le;
  }
  return check_launch("f2g_stft_group");
}

extern "C" int f2g_irfft_group(const F2GSpecProblem* probs, int np, void* stream) {
  SpecGroupArgs g;
  int max_n = 0;
  if (int rc = spec_group_fill(g, probs, np, "f2g_irfft_group", &max_n)) return rc;
  if (spec_group_warp_ok(probs, np)) {
    if (int rc = fft_roots_init()) return rc;
    const int rows = g.row_begin[np];
    cudaError_t le = launch_pdl(irfft_group_warp_kernel, dim3((rows + FFT_WARPS - 1) / FFT_WARPS),
                                dim3(32 * FFT_WARPS), 0, static_cast<cudaStream_t>(stream), g);
    if (le != cudaSuccess) {
      set_error("f2g_irfft_group launch: %s", cudaGetErrorString(le));
      return (int)le;
    }
    return check_launch("f2g_irfft_group");
  }
  const int threads = max_n / 2 < 32 ? 32 : max_n / 2;
  const size_t smem = (size_t)(2 * max_n + max_n / 2) * sizeof(float2);
  cudaError_t le = launch_pdl(irfft_group_kernel, dim3(g.row_begin[np]), dim3(threads), smem,
                              static_cast<cudaStream_t>(stream), g);
  if (le != cudaSuccess) {
    set_error("f2g_irfft_group launch: %s", cudaGetErrorString(le));
    return (int)le;
  }
  return check_launch("f2g_irfft_group");
}

extern "C" int f2g_istft_bwd_spec(const float* gs, int B, int Lp, int n_fft, int hop,
                                  const float* row_mask, float* dpacked, int ld, int round_tf32,
                                  void* stream) {
  return stft_launch(gs, B, Lp, Lp, n_fft, hop, F2G_SPEC_PACKED, nullptr, nullptr, 0, 0.f, dpacked, ld,
                     round_tf32, 0, 1, row_mask, stream);
}

extern "C" int f2g_stft_bwd_frames(const float* dpacked, int rows, int ld, int n_fft, float* frames_out,
                                   int interleaved, void* stream) {
  const int logn = ilog2_exact(n_fft);
  if (logn < 5 || n_fft > 2048) {
    set_error("f2g_stft_bwd_frames: n_fft=%d must be a power of two in [32, 2048]", n_fft);
    return F2G_EINVAL;
  }
  const int threads = n_fft / 2 < 32 ? 32 : n_fft / 2;
  const size_t smem = (size_t)(2 * n_fft + n_fft / 2) * sizeof(float2);
  F2G_LAUNCH_COOP_SMEM(stft_bwd_frames_kernel, rows, threads, smem, static_cast<cudaStream_t>(stream), dpacked, ld,
                       n_fft, logn, frames_out, interleaved);
  return check_launch("f2g_stft_bwd_frames");
}

extern "C" int f2g_spec_loss_bwd(const float* audio, int B, int T, int ld_audio, int n_fft, int hop,
                                 int mode, const float* fb, int n_filt, float log_clip, const float* dF,
                                 int ld_dF, float* frames_out, const int* fb_ranges, void* stream) {
  const int logn = ilog2_exact(n_fft);
  if (logn < 5 || n_fft > 2048 || n_fft / 2 >= T || !fb) {
    set_error("f2g_spec_loss_bwd: bad arguments (n_fft=%d, T=%d)", n_fft, T);
    return F2G_EINVAL;
  }
  const int frames = 1 + T / hop;
  const int threads = n_fft / 2 < 32 ? 32 : n_fft / 2;
  const size_t smem = (size_t)(2 * n_fft + n_fft / 2) * sizeof(float2) +
                      (size_t)(n_fft / 2 + 1 + n_filt) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(spec_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr = true;
  }
  F2G_LAUNCH_COOP_SMEM(spec_loss_bwd_kernel, B * frames, threads, smem, static_cast<cudaStream_t>(stream), audio, T,
                       ld_audio, n_fft, logn, hop, frames, mode, fb, n_filt, log_clip, dF, ld_dF, frames_out, fb_ranges);
  return check_launch("f2g_spec_loss_bwd");
}

extern "C" int f2g_dc_peak(const float* audio, int B, int T, int ld_audio, float* pre, void* stream) {
  F2G_LAUNCH_COOP(dc_peak_kernel, B, 512, static_cast<cudaStream_t>(stream), audio, T, ld_audio, pre);
  return check_launch("f2g_dc_peak");
}

extern "C" int f2g_irfft_frames(const float* packed, int rows, int ld, int n_fft, float* frames_out,
                                void* stream) {
  const int logn = ilog2_exact(n_fft);
  if (logn < 5 || n_fft > 2048) {
    set_error("f2g_irfft_frames: n_fft=%d must be a power of two in [32, 2048]", n_fft);
    return F2G_EINVAL;
  }
  const int threads = n_fft / 2 < 32 ? 32 : n_fft / 2;
  const size_t smem = (size_t)(2 * n_fft + n_fft / 2) * sizeof(float2);
  F2G_LAUNCH_COOP_SMEM(irfft_frames_kernel, rows, threads, smem, static_cast<cudaStream_t>(stream), packed, ld, n_fft,
                       logn, frames_out);
  return check_launch("f2g_irfft_frames");
}

extern "C" int f2g_ola_combine(const float* const* frames, const int* n_ffts, const int* hops,
                               const int* n_frames, int nb, const float* weight, const float* x,
                               float* out, int B, int T, int euler, float t, float dt, int clamp,
                               void* stream) {
  if (nb < 1 || nb > 4) {
    set_error("f2g_ola_combine: nb=%d out of range", nb);
    return F2G_EINVAL;
  }
  OlaArgs a;
  a.nb = nb;
  for (int j = 0; j < nb; ++j) {
    a.fr[j] = frames[j];
    a.n[j] = n_ffts[j];
    a.hop[j] = hops[j];
    a.frames[j] = n_frames[j];
  }
  dim3 grid((T + 255) / 256, B);
  F2G_LAUNCH(ola_combine_kernel, grid, 256, static_cast<cudaStream_t>(stream), a, weight, x, out, T, euler, t, dt,
                  clamp);
  return check_launch("f2g_ola_combine");
}
