// Data path either side of the generator:
//   * f2g_pcm_decode     wav payload -> mono float32 segment + RMS / peak statistics
//                        (flow2gan/dataset.py:122-160 load_audio/offset/duration/is_silence/mean over
//                         channels; bin/infer_dir.py:217-220, test_from_wav.py:62-66)
//   * f2g_gain_resample  sox `norm <dB>` peak normalisation + torchaudio sinc-Hann polyphase
//                        resampling in one pass (dataset.py:164-173)
//   * f2g_pcm16_encode   float -> PCM16 as soundfile/libsndfile writes it (bin/infer.py:208-212,
//                        bin/infer_dir.py:237)
//   * f2g_average_update fp64 running model average / checkpoint interval average, multi-tensor
//                        (flow2gan/checkpoint.py:378-409,452-531 average_state_dict)
// All four are thread-independent HBM streams written against simt.cuh, so the same source also
// compiles for the host (tests only) where its arithmetic is checked without a GPU.
#include "simt.cuh"
#include "../../include/flow2gan_b200.h"

namespace f2g {

constexpr int DP_THREADS = 256;
constexpr int AVG_CHUNK = 4096;

F2G_SIMT_DEV float pcm_sample(const unsigned char* __restrict__ p, int fmt, long long idx) {
  switch (fmt) {
    case F2G_PCM_S16: {
      const short v = reinterpret_cast<const short*>(p)[idx];
      return (float)v * (1.0f / 32768.0f);
    }
    case F2G_PCM_S24: {
      const unsigned char* q = p + idx * 3;
      int v = (int)q[0] | ((int)q[1] << 8) | ((int)q[2] << 16);
      v = (int)((unsigned)v << 8) >> 8;                   // sign-extend 24 -> 32 bits
      return (float)v * (1.0f / 8388608.0f);
    }
    case F2G_PCM_S32: {
      const int v = reinterpret_cast<const int*>(p)[idx];
      return (float)v * (1.0f / 2147483648.0f);
    }
    default:
      return reinterpret_cast<const float*>(p)[idx];
  }
}

// mono[i] = mean_c pcm[(first + i) * channels + c];  stats[0] += sum mono^2, stats[1] = max |mono|
F2G_KERNEL void pcm_decode_kernel(const unsigned char* __restrict__ pcm, int fmt, int channels,
                                  long long first, long long n, float* __restrict__ mono,
                                  float* __restrict__ stats) {
  float ss = 0.f, pk = 0.f;
  const float fc = (float)channels;
  for (long long i = F2G_GTID; i < n; i += F2G_GSTRIDE) {
    const long long base = (first + i) * channels;
    float s = pcm_sample(pcm, fmt, base);
    for (int c = 1; c < channels; ++c) s += pcm_sample(pcm, fmt, base + c);
    if (channels > 1) s = s / fc;
    mono[i] = s;
    ss += simt_fmul(s, s);
    pk = fmaxf(pk, fabsf(s));
  }
  if (stats) {
    simt_block_sum(ss, stats);
    simt_block_max_nonneg(pk, stats + 1);
  }
}

// out[j*new_r + p] = sum_k taps[p][k] * g * x[j*orig_r + k - width]   (zero outside [0, n_in))
// = F.conv1d(pad(g*x, (width, width + orig_r)), taps, stride=orig_r) flattened phase-minor,
// truncated to n_out (torchaudio.functional._apply_sinc_resample_kernel).  orig_r == new_r == 1 with
// width == 0 and taps == {1} is the pure gain pass.
F2G_KERNEL void gain_resample_kernel(const float* __restrict__ x, long long n_in,
                                     const float* __restrict__ stats, float target_lin, int orig_r,
                                     int new_r, int width, const float* __restrict__ taps,
                                     float* __restrict__ out, long long n_out) {
  float g = 1.f;
  if (stats) g = target_lin / fmaxf(stats[1], 1.0e-30f);      // sox `norm`: peak -> target level
  const int kw = 2 * width + orig_r;
  for (long long o = F2G_GTID; o < n_out; o += F2G_GSTRIDE) {
    const long long j = o / new_r;
    const int p = (int)(o - j * new_r);
    const float* __restrict__ tp = taps + (long long)p * kw;
    const long long x0 = j * orig_r - width;
    int k0 = 0, k1 = kw;
    if (x0 < 0) k0 = (int)(-x0);
    if (x0 + kw > n_in) k1 = (int)(n_in - x0);
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc += simt_fmul(tp[k], simt_fmul(x[x0 + k], g));
    out[o] = acc;
  }
}

// libsndfile f2s_array with norm_float: lrintf(x * 32767); `clamp` bounds x to [-1, 1] first
// (the reference writes infer(clamp_pred=True) output, which is already bounded)
F2G_KERNEL void pcm16_encode_kernel(const float* __restrict__ x, long long n, int clamp,
                                    short* __restrict__ out) {
  for (long long i = F2G_GTID; i < n; i += F2G_GSTRIDE) {
    float v = x[i];
    if (clamp) v = fminf(fmaxf(v, -1.f), 1.f);
    out[i] = (short)simt_rint(simt_fmul(v, 32767.0f));
  }
}

// avg = (avg * w_avg + cur * w_cur) * scale with every rounding step of the reference's in-place
// sequence (v *= w1; v += cur * w2; v *= scale) as torch evaluates it:
//   fp64 accumulator (model_avg as created, finetune.py:902): products / sum in fp64; an fp32 `cur`
//     is first multiplied in fp32 (tensor * python float keeps the tensor dtype);
//   fp32 accumulator (model_avg after the first save_checkpoint, which calls
//     model_avg.to(torch.float32) in place, checkpoint.py:94-95): scalars are cast to fp32; adding an
//     fp64 `cur` term promotes the sum to fp64 and rounds the result back to fp32.
F2G_KERNEL void average_update_kernel(const F2GAvgTensor* __restrict__ tab, const int* __restrict__ chunks,
                                      double w_avg, double w_cur, double scale) {
  const int ti = chunks[2 * blockIdx.x], ci = chunks[2 * blockIdx.x + 1];
  const F2GAvgTensor t = tab[ti];
  const long long base = (long long)ci * AVG_CHUNK;
  const long long end = min(base + (long long)AVG_CHUNK, t.numel);
  const float w_avg_f = (float)w_avg, w_cur_f = (float)w_cur, scale_f = (float)scale;
  if (!t.avg_is_f32) {
    double* __restrict__ avg = reinterpret_cast<double*>(t.avg);
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
      double term;
      if (t.cur_is_f64)
        term = simt_dmul(reinterpret_cast<const double*>(t.cur)[i], w_cur);
      else
        term = (double)simt_fmul(reinterpret_cast<const float*>(t.cur)[i], w_cur_f);
      double v = simt_dmul(avg[i], w_avg);
      v = simt_dadd(v, term);
      avg[i] = simt_dmul(v, scale);
    }
  } else {
    float* __restrict__ avg = reinterpret_cast<float*>(t.avg);
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
      float v = simt_fmul(avg[i], w_avg_f);
      if (t.cur_is_f64)
        v = (float)simt_dadd((double)v, simt_dmul(reinterpret_cast<const double*>(t.cur)[i], w_cur));
      else
        v = simt_fadd(v, simt_fmul(reinterpret_cast<const float*>(t.cur)[i], w_cur_f));
      avg[i] = simt_fmul(v, scale_f);
    }
  }
}

static int stream_grid(long long n) {
  long long g = (n + DP_THREADS - 1) / DP_THREADS;
  const long long cap = 148LL * 16;                 // 16 resident 256-thread CTAs per SM
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace f2g

using namespace f2g;

extern "C" {

int f2g_pcm_decode(const void* pcm, int sample_format, int channels, long long first_frame,
                   long long n_frames, float* mono, float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (sample_format != F2G_PCM_S16 && sample_format != F2G_PCM_S24 && sample_format != F2G_PCM_S32 &&
      sample_format != F2G_PCM_F32) {
    set_error("f2g_pcm_decode: unsupported sample format %d", sample_format);
    return F2G_EINVAL;
  }
  if (channels < 1 || first_frame < 0 || n_frames < 0 || !pcm || !mono) {
    set_error("f2g_pcm_decode: bad arguments (channels %d, first %lld, n %lld)", channels, first_frame,
              n_frames);
    return F2G_EINVAL;
  }
  if (n_frames == 0) return F2G_OK;
  F2G_LAUNCH(pcm_decode_kernel, stream_grid(n_frames), DP_THREADS, stream,
             static_cast<const unsigned char*>(pcm), sample_format, channels, first_frame, n_frames, mono,
             stats);
  return check_launch("f2g_pcm_decode");
}

int f2g_gain_resample(const float* x, long long n_in, const float* stats, float norm_db, int orig_r,
                      int new_r, int width, const float* taps, float* out, long long n_out,
                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (orig_r < 1 || new_r < 1 || width < 0 || !x || !taps || !out || n_in < 0 || n_out < 0) {
    set_error("f2g_gain_resample: bad arguments (orig %d, new %d, width %d)", orig_r, new_r, width);
    return F2G_EINVAL;
  }
  // every output must come from the padded signal of the reference: (n_in + 2*width + orig_r - kw)
  // / orig_r + 1 conv positions, new_r outputs each
  const long long max_out = (n_in / orig_r + 1) * (long long)new_r;
  if (n_out > max_out) {
    set_error("f2g_gain_resample: n_out %lld exceeds the %lld samples the padded input yields", n_out,
              max_out);
    return F2G_EINVAL;
  }
  if (n_out == 0) return F2G_OK;
  const float target_lin = powf(10.0f, norm_db / 20.0f);
  F2G_LAUNCH(gain_resample_kernel, stream_grid(n_out), DP_THREADS, stream, x, n_in, stats, target_lin,
             orig_r, new_r, width, taps, out, n_out);
  return check_launch("f2g_gain_resample");
}

int f2g_pcm16_encode(const float* x, long long n, int clamp, short* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || !x || !out) {
    set_error("f2g_pcm16_encode: bad arguments");
    return F2G_EINVAL;
  }
  if (n == 0) return F2G_OK;
  F2G_LAUNCH(pcm16_encode_kernel, stream_grid(n), DP_THREADS, stream, x, n, clamp, out);
  return check_launch("f2g_pcm16_encode");
}

int f2g_average_update(const F2GAvgTensor* tab_dev, const int* chunks_dev, int n_chunks, double w_avg,
                       double w_cur, double scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!tab_dev || !chunks_dev || n_chunks <= 0) {
    set_error("f2g_average_update: empty tensor table");
    return F2G_EINVAL;
  }
  F2G_LAUNCH(average_update_kernel, n_chunks, DP_THREADS, stream, tab_dev, chunks_dev, w_avg, w_cur,
             scale);
  return check_launch("f2g_average_update");
}

}  // extern "C"
