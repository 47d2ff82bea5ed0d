// Warp-per-frame real FFT / inverse real FFT for the generator's STFT <-> packed-spectrum hops
// (torch.stft / torch.istft as called from flow2gan/models/modules.py:68-84,105-116 with
// n_fft in {128, ..., 1024}).  One warp transforms one frame of n = 64*R samples as a complex FFT
// of N = n/2 = 32*R points on z[j] = x[2j] + i x[2j+1]:
//     lane l, register r hold z[32 r + l]
//  -> R-point DFT over r in registers (radix-2 DIF, compile-time twiddles)
//  -> twiddle w_N^(l q)
//  -> 32-point DFT across lanes with __shfl_xor butterflies (5 DIF stages; lane l ends up with
//     component bitrev5(l))
//  -> Z[q + R p] ; the real-input split X[k] = (Z[k] + Z*[N-k])/2 - i/2 w_n^k (Z[k] - Z*[N-k])
//     runs out of a per-warp shared-memory row, which also turns the stores into full lines.
// The inverse runs the same network backwards (conjugate twiddles).  No block-wide barriers, no
// per-frame twiddle generation (one 2048-entry root-of-unity table in global memory): the
// shared-memory Stockham kernel this replaces was bound by its 10-stage __syncthreads chain.
#pragma once
#include "simt.cuh"

namespace f2g {

__device__ float2 g_fft_roots[2048];     // exp(-2 pi i k / 2048), filled by fft_roots_init()

static int fft_roots_init() {
  static bool done = false;
  if (done) return 0;
  static float2 host[2048];
  for (int k = 0; k < 2048; ++k) {
    const double a = -2.0 * 3.14159265358979323846 * (double)k / 2048.0;
    host[k] = make_float2((float)cos(a), (float)sin(a));
  }
  cudaError_t e = cudaMemcpyToSymbol(g_fft_roots, host, sizeof(host));
  if (e != cudaSuccess) {
    set_error("fft root table upload: %s", cudaGetErrorString(e));
    return (int)e;
  }
  done = true;
  return 0;
}

F2G_DEVINL float2 fft_root(int idx) { return __ldg(&g_fft_roots[idx & 2047]); }
F2G_DEVINL float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// exp(-2 pi i j / 32), j < 16: compile-time constants for the in-register butterflies
__device__ constexpr float kCos32[16] = {1.0f,          0.98078528f,  0.92387953f,  0.83146961f,
                                         0.70710678f,   0.55557023f,  0.38268343f,  0.19509032f,
                                         0.0f,          -0.19509032f, -0.38268343f, -0.55557023f,
                                         -0.70710678f,  -0.83146961f, -0.92387953f, -0.98078528f};
__device__ constexpr float kSin32[16] = {0.0f,          0.19509032f,  0.38268343f,  0.55557023f,
                                         0.70710678f,   0.83146961f,  0.92387953f,  0.98078528f,
                                         1.0f,          0.98078528f,  0.92387953f,  0.83146961f,
                                         0.70710678f,   0.55557023f,  0.38268343f,  0.19509032f};

constexpr int fft_bitrev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
constexpr int fft_log2(int x) { return x <= 1 ? 0 : 1 + fft_log2(x >> 1); }

// R-point DFT over the register index, natural order in and out.  INV: conjugate twiddles.
template <int R, bool INV>
F2G_DEVINL void fft_regs(float2 (&v)[R]) {
#pragma unroll
  for (int d = R / 2; d >= 1; d >>= 1) {
#pragma unroll
    for (int base = 0; base < R; base += 2 * d) {
#pragma unroll
      for (int m = 0; m < d; ++m) {
        const float2 a = v[base + m], b = v[base + m + d];
        const int j = m * (16 / d);                       // angle 2 pi m / (2d) in units of 2 pi / 32
        const float wc = kCos32[j], ws = INV ? kSin32[j] : -kSin32[j];
        const float2 t = make_float2(a.x - b.x, a.y - b.y);
        v[base + m] = make_float2(a.x + b.x, a.y + b.y);
        v[base + m + d] = make_float2(t.x * wc - t.y * ws, t.x * ws + t.y * wc);
      }
    }
  }
  float2 o[R];
#pragma unroll
  for (int i = 0; i < R; ++i) o[fft_bitrev(i, fft_log2(R))] = v[i];
#pragma unroll
  for (int i = 0; i < R; ++i) v[i] = o[i];
}

// 32-point DFT across the lanes of every register slot: natural lane order in, lane l holds
// component bitrev5(l) out.  tw[s] = this lane's twiddle of stage s (distance 16 >> s).
template <int R, bool INV>
F2G_DEVINL void fft_lanes(float2 (&v)[R], int lane) {
  float2 tw[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int d = 16 >> s;
    float2 w = fft_root((lane & (d - 1)) * (1024 / d));     // exp(-2 pi i (lane mod d) / (2d))
    if (INV) w.y = -w.y;
    tw[s] = w;
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int d = 16 >> s;
    const bool upper = (lane & d) != 0;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const float ox = __shfl_xor_sync(0xffffffffu, v[q].x, d);
      const float oy = __shfl_xor_sync(0xffffffffu, v[q].y, d);
      if (!upper) {
        v[q] = make_float2(v[q].x + ox, v[q].y + oy);
      } else {
        const float2 t = make_float2(ox - v[q].x, oy - v[q].y);
        v[q] = cmulf(t, tw[s]);
      }
    }
  }
}

// padded index into the per-warp shared row: conflict-free for the stride-R accesses
F2G_DEVINL int fft_pad(int k) { return k + (k >> 5); }
constexpr int FFT_WARP_MAX_N = 1024;                                   // largest n_fft of this path
constexpr int FFT_WARP_ROW = FFT_WARP_MAX_N / 2 + 1 + FFT_WARP_MAX_N / 64 + 1;   // floats per re / im row

// Forward: frame `f` of the reflect-padded, hann-windowed signal -> packed row [Re(0..N) | Im(0..N)].
template <int R>
F2G_DEVINL void warp_rfft_packed(const float* __restrict__ x, int T, int start, float* __restrict__ o,
                                 int ld_out, bool round_tf32, float* __restrict__ sre,
                                 float* __restrict__ sim, int lane) {
  constexpr int N = 32 * R, n = 64 * R, ST = 2048 / n;
  float2 v[R];
  const bool interior = start >= 0 && start + n <= T && ((reinterpret_cast<uintptr_t>(x + start) & 7) == 0);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = 64 * r + 2 * lane;
    float2 s;
    if (interior) {
      s = *reinterpret_cast<const float2*>(x + start + i0);
    } else {
      int p0 = start + i0, p1 = p0 + 1;
      if (p0 < 0) p0 = -p0;
      if (p0 >= T) p0 = 2 * (T - 1) - p0;
      if (p1 < 0) p1 = -p1;
      if (p1 >= T) p1 = 2 * (T - 1) - p1;
      s = make_float2(x[p0], x[p1]);
    }
    const float w0 = 0.5f - 0.5f * fft_root(i0 * ST).x;            // periodic hann
    const float w1 = 0.5f - 0.5f * fft_root((i0 + 1) * ST).x;
    v[r] = make_float2(s.x * w0, s.y * w1);
  }
  fft_regs<R, false>(v);
#pragma unroll
  for (int q = 1; q < R; ++q) v[q] = cmulf(v[q], fft_root(lane * q * (2048 / N)));
  fft_lanes<R, false>(v, lane);
  const int p = (int)(__brev((unsigned)lane) >> 27);                 // component held by this lane
#pragma unroll
  for (int q = 0; q < R; ++q) {
    const int k = q + R * p;
    sre[fft_pad(k)] = v[q].x;
    sim[fft_pad(k)] = v[q].y;
  }
  __syncwarp();
  constexpr int nb = N + 1;
#pragma unroll
  for (int t = 0; t < R; ++t) {
    const int k = 32 * t + lane;
    const int kn = (N - k) & (N - 1);
    const float2 zk = make_float2(sre[fft_pad(k)], sim[fft_pad(k)]);
    const float2 zn = make_float2(sre[fft_pad(kn)], sim[fft_pad(kn)]);
    const float2 e = fft_root(k * ST);                                // exp(-2 pi i k / n)
    const float ax = zk.x + zn.x, ay = zk.y - zn.y;                   // Z[k] + conj Z[N-k]
    const float bx = zk.x - zn.x, by = zk.y + zn.y;                   // Z[k] - conj Z[N-k]
    const float ebx = e.x * bx - e.y * by, eby = e.x * by + e.y * bx;
    float re = 0.5f * (ax + eby), im = 0.5f * (ay - ebx);
    if (k == 0) im = 0.f;
    o[k] = round_tf32 ? tf32_rna(re) : re;
    o[nb + k] = round_tf32 ? tf32_rna(im) : im;
  }
  if (lane == 0) {                                                     // Nyquist bin
    const float ny = sre[0] - sim[0];
    o[N] = round_tf32 ? tf32_rna(ny) : ny;
    o[nb + N] = 0.f;
  }
  for (int k = 2 * nb + lane; k < ld_out; k += 32) o[k] = 0.f;
  __syncwarp();
}

// Inverse: packed row -> hann-windowed time frame  fr[i] = w[i] * irfft(X)[i]  (norm 1/n).
template <int R>
F2G_DEVINL void warp_irfft_frame(const float* __restrict__ pk, float* __restrict__ fr,
                                 float* __restrict__ sre, float* __restrict__ sim, int lane) {
  constexpr int N = 32 * R, n = 64 * R, ST = 2048 / n, nb = N + 1;
#pragma unroll
  for (int t = 0; t < R; ++t) {
    const int k = 32 * t + lane;
    sre[fft_pad(k)] = pk[k];
    sim[fft_pad(k)] = (k == 0) ? 0.f : pk[nb + k];                   // C2R ignores Im(DC), Im(Nyquist)
  }
  if (lane == 0) {
    sre[fft_pad(N)] = pk[N];
    sim[fft_pad(N)] = 0.f;
  }
  __syncwarp();
  float2 v[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    const int k = R * lane + q;
    const float2 xk = make_float2(sre[fft_pad(k)], sim[fft_pad(k)]);
    const float2 xn = make_float2(sre[fft_pad(N - k)], sim[fft_pad(N - k)]);
    float2 e = fft_root(k * ST);
    e.y = -e.y;                                                        // exp(+2 pi i k / n)
    const float ax = xk.x + xn.x, ay = xk.y - xn.y;                   // X[k] + conj X[N-k]
    const float bx = xk.x - xn.x, by = xk.y + xn.y;                   // X[k] - conj X[N-k]
    const float ebx = e.x * bx - e.y * by, eby = e.x * by + e.y * bx;
    v[q] = make_float2(0.5f * (ax - eby), 0.5f * (ay + ebx));          // + i e B
  }
  __syncwarp();
  fft_lanes<R, true>(v, lane);
  const int l = (int)(__brev((unsigned)lane) >> 27);                 // time-index residue held by this lane
#pragma unroll
  for (int q = 1; q < R; ++q) {
    float2 w = fft_root(l * q * (2048 / N));
    w.y = -w.y;
    v[q] = cmulf(v[q], w);
  }
  fft_regs<R, true>(v);
  constexpr float inv = 1.0f / (float)N;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = 64 * r + 2 * l;
    const float w0 = 0.5f - 0.5f * fft_root(i0 * ST).x;
    const float w1 = 0.5f - 0.5f * fft_root((i0 + 1) * ST).x;
    *reinterpret_cast<float2*>(fr + i0) = make_float2(v[r].x * inv * w0, v[r].y * inv * w1);
  }
}

}  // namespace f2g
