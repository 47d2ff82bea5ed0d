// Fused multi-tensor ScaledAdam step (reference: flow2gan/optim.py:125-255 basic/scaling/
// momentum_step, :451-507 ScaledAdam.step, :509-619 _get_clipping_scale).
//
// The reference stacks same-shape parameters and grads every step (torch.stack copies) and
// calls .item() for the clipping scale.  Here every parameter tensor is visited in place by
// three launches over a static chunk table; all per-tensor scalars (param_rms, scale_grads,
// scale_exp_avg_sq, scale_step) and the clip factor live on the device:
//   1. reduce : per tensor  sum g^2, sum p*g, sum p^2                (HBM: read p, g)
//   2. scalars: one CTA -- grad-norm, clip factor, scale_grads / param_rms / size-step logic
//   3. update : exp_avg_sq EMA, -lr*g/(sqrt(v)+eps) * rms, + p*scale_step, momentum, p += delta
//               (HBM: read p, g, v, d; write p, v, d -> 28 B/param, pure streaming)
#include "simt.cuh"
#include "../../include/flow2gan_b200.h"

namespace f2g {

constexpr int CHUNK = 4096;        // elements per CTA pass
constexpr int OPT_THREADS = 256;

F2G_KERNEL void adam_reduce_kernel(const F2GAdamTensor* __restrict__ tab, const int2* __restrict__ chunks,
                                   float* __restrict__ acc /* [n_tensors][3] */) {
  const int2 ck = chunks[blockIdx.x];
  const F2GAdamTensor t = tab[ck.x];
  const long long base = (long long)ck.y * CHUNK;
  const long long end = min(base + CHUNK, t.numel);
  float sg = 0.f, spg = 0.f, sp = 0.f;
  for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
    const float p = t.p[i];
    const float g = t.g ? t.g[i] : 0.f;
    sg = fmaf(g, g, sg);
    spg = fmaf(p, g, spg);
    sp = fmaf(p, p, sp);
  }
  __shared__ float red[3][OPT_THREADS / 32];
  sg = warp_sum(sg); spg = warp_sum(spg); sp = warp_sum(sp);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = sg; red[1][w] = spg; red[2][w] = sp; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < OPT_THREADS / 32; ++i) s += red[threadIdx.x][i];
    atomicAdd(acc + (size_t)ck.x * 3 + threadIdx.x, s);
  }
}

// state scalars per tensor: [0]=param_rms [1]=scale_exp_avg_sq [2..5]=scale_grads[0..3] [6]=scale_step
// group scalars: [0]=tot_norm (out) [1]=clip (out) [2]=threshold (<0: unset) [3]=num_clipped (counter,
// optim.py:607-608; the host reads and clears it when it refreshes the threshold)
F2G_KERNEL void adam_norm_kernel(const F2GAdamTensor* __restrict__ tab, int n, const float* __restrict__ acc,
                                 float* __restrict__ ts, float* __restrict__ gs, int step,
                                 float scalar_lr_scale, float* __restrict__ model_norms, int period) {
  __shared__ float red[OPT_THREADS / 32];
  float tot = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float sg = acc[(size_t)i * 3];
    if (tab[i].is_scalar) {
      tot += sg * scalar_lr_scale * scalar_lr_scale;
    } else {
      float rms = ts[(size_t)i * 8];
      if (step == 0) {   // first step: param_rms is initialised from p (optim.py:171-173)
        rms = sqrtf(acc[(size_t)i * 3 + 2] / (float)tab[i].numel);
        ts[(size_t)i * 8] = rms;
      }
      tot += sg * rms * rms;
    }
  }
  tot = warp_sum(tot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < OPT_THREADS / 32; ++i) s += red[i];
    const float norm = sqrtf(s);
    gs[0] = norm;
    if (step > 0 && model_norms) model_norms[step % period] = norm;
  }
}

F2G_KERNEL void adam_scalars_kernel(const F2GAdamTensor* __restrict__ tab, int n, const float* __restrict__ acc,
                                    float* __restrict__ ts, float* __restrict__ gs, int step,
                                    int use_clip, float lr, float scalar_lr_scale, float beta2,
                                    float eps, float min_rms, float max_rms, int size_period) {
  // clip factor (optim.py:605-617); gs[2] < 0 means "threshold not yet set"
  float clip = 1.f;
  if (use_clip && step > 0 && gs[2] >= 0.f) {
    clip = fminf(1.f, gs[2] / (gs[0] + 1.0e-20f));
    if (clip != clip) clip = 0.f;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    gs[1] = clip;
    if (clip < 1.f) gs[3] += 1.f;
  }
  const int slot = step % size_period;
  const bool refresh = slot == size_period - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float* s = ts + (size_t)i * 8;
    s[6] = 0.f;
    if (tab[i].is_scalar) continue;
    const float dot = (clip == 0.f) ? 0.f : acc[(size_t)i * 3 + 1] * clip;   // (p * grad_clipped).sum()
    s[2 + slot] = dot;
    if (refresh) s[0] = sqrtf(acc[(size_t)i * 3 + 2] / (float)tab[i].numel);
    if (refresh && step > 0) {
      const float b2c = powf(beta2, (float)size_period);
      float msq = 0.f, sum = 0.f;
      for (int k = 0; k < size_period; ++k) { msq += s[2 + k] * s[2 + k]; sum += s[2 + k]; }
      msq /= (float)size_period;
      s[1] = s[1] * b2c + (1.f - b2c) * msq;
      const int size_step = (step + 1) / size_period;
      const float bc2 = 1.f - powf(b2c, (float)size_step);
      const float denom = sqrtf(s[1]) + eps;
      float ss = -(lr * scalar_lr_scale) * sqrtf(bc2) * sum / denom;
      const float rms = s[0];
      if (rms < min_rms) ss = 0.f;
      ss = fminf(fmaxf(ss, -0.1f), 0.1f);
      ss = fminf(ss, (max_rms - rms) / rms);
      s[6] = ss;
    }
  }
}

F2G_KERNEL void adam_update_kernel(const F2GAdamTensor* __restrict__ tab, const int2* __restrict__ chunks,
                                   const float* __restrict__ ts, const float* __restrict__ gs, int step,
                                   float lr, float scalar_lr_scale, float beta1, float beta2, float eps,
                                   float min_rms, float scalar_max) {
  const int2 ck = chunks[blockIdx.x];
  const F2GAdamTensor t = tab[ck.x];
  const long long base = (long long)ck.y * CHUNK;
  const long long end = min(base + CHUNK, t.numel);
  const float clip = gs[1];
  const float* s = ts + (size_t)ck.x * 8;
  const float rms_mult = t.is_scalar ? 1.f : fmaxf(s[0], min_rms);
  const float ss = t.is_scalar ? 0.f : s[6];
  const float lr_eff = t.is_scalar ? lr * scalar_lr_scale : lr;
  const float bc2 = 1.f - powf(beta2, (float)(step + 1));
  const float vscale = bc2 < 0.99f ? 1.f / bc2 : 1.f;
  for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
    float g = t.g ? t.g[i] : 0.f;
    g = (clip == 0.f) ? 0.f : g * clip;
    const float p = t.p[i];
    const float v = t.v[i] * beta2 + (1.f - beta2) * g * g;
    t.v[i] = v;
    const float denom = sqrtf(v * vscale) + eps;
    float delta = -lr_eff * g / denom;
    delta *= rms_mult;
    delta = fmaf(p, ss, delta);
    const float d = t.d[i] * beta1 + (1.f - beta1) * delta;
    t.d[i] = d;
    float pn = p + d;
    if (t.is_scalar) pn = fminf(fmaxf(pn, -scalar_max), scalar_max);
    t.p[i] = pn;
  }
}

}  // namespace f2g

using namespace f2g;

extern "C" int f2g_scaled_adam_step(const F2GAdamTensor* tab_dev, int n_tensors, const int* chunks_dev,
                                    int n_chunks, float* acc_dev, float* tensor_state_dev,
                                    float* group_state_dev, float* model_norms_dev, int step,
                                    int phase, const F2GAdamHyper* h, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_tensors <= 0 || n_chunks <= 0) {
    set_error("f2g_scaled_adam_step: empty parameter table");
    return F2G_EINVAL;
  }
  if (h->size_update_period < 1 || h->size_update_period > 4) {   // scale_grads is a fixed 4-slot row
    set_error("f2g_scaled_adam_step: size_update_period %d outside 1..4", h->size_update_period);
    return F2G_EINVAL;
  }
  const int2* chunks = reinterpret_cast<const int2*>(chunks_dev);
  if (phase == 0) {         // reductions + gradient norm (host may then refresh the threshold)
    if (int rc = simt_memset_async(acc_dev, 0, sizeof(float) * 3 * (size_t)n_tensors, stream)) return rc;
    F2G_LAUNCH_COOP(adam_reduce_kernel, n_chunks, OPT_THREADS, stream, tab_dev, chunks, acc_dev);
    F2G_LAUNCH_COOP(adam_norm_kernel, 1, OPT_THREADS, stream, tab_dev, n_tensors, acc_dev, tensor_state_dev,
                    group_state_dev, step, h->scalar_lr_scale, model_norms_dev, h->clipping_update_period);
    return check_launch("f2g_scaled_adam_step(reduce)");
  }
  const int sb = (n_tensors + OPT_THREADS - 1) / OPT_THREADS;
  F2G_LAUNCH_COOP(adam_scalars_kernel, sb, OPT_THREADS, stream, tab_dev, n_tensors, acc_dev, tensor_state_dev,
                  group_state_dev, step, h->use_clipping, h->lr, h->scalar_lr_scale, h->beta2, h->eps,
                  h->param_min_rms, h->param_max_rms, h->size_update_period);
  F2G_LAUNCH_COOP(adam_update_kernel, n_chunks, OPT_THREADS, stream, tab_dev, chunks, tensor_state_dev,
                  group_state_dev, step, h->lr, h->scalar_lr_scale, h->beta1, h->beta2, h->eps, h->param_min_rms,
                  h->scalar_max);
  return check_launch("f2g_scaled_adam_step(update)");
}
