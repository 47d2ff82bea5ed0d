// Fused multi-tensor GAN loss reductions (reference:
// flow2gan/models/gan.py:57-99):
//   L1 terms    sum_i mean|a_i - b_i|            feature-matching loss (:72-87, a = real.detach()),
//                                                 multi-scale log-mel reconstruction loss (:89-99)
//   hinge terms sum_i mean(clamp(1 + s_i*x_i, 0)) discriminator_loss (:57-63) / generator_loss (:65-70)
// The reference evaluates every term with 3-4 element-wise torch launches forward and as many
// backward (~90 feature maps + 16 scores + 7 mels per iteration); here up to 24 terms share one
// launch each way.  Terms are addressed through 4-D strides because the feature maps are permuted /
// channel-sliced views of the channel-last conv outputs.  Thread-independent kernels written against
// simt.cuh (host-emulated in tests).
#include "simt.cuh"
#include "../../include/flow2gan_b200.h"

namespace f2g {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_PER_THREAD = 8;       // units per thread; a unit = one element (scalar terms) or one float4 (vector terms)
constexpr int LOSS_CHUNK = LOSS_THREADS * LOSS_PER_THREAD;
constexpr int LOSS_MAX_TERMS = 24;     // 24 x 128 B descriptors + tables < 4 KB of kernel params

struct LossArgs {
  F2GLossTerm t[LOSS_MAX_TERMS];
  int block_begin[LOSS_MAX_TERMS + 1];
  // vector terms: innermost dimension contiguous in both operands, a multiple of 4 long, every other
  // stride / the base addresses 16-byte aligned -> one index decomposition and one 128-bit access per
  // four elements (the feature maps are channel-last views: C = 32..1024 innermost).  The index math is
  // unsigned 32-bit (numel < 2^31, checked on the host): the 64-bit divisions of the first version cost
  // ~600 instructions per element and left the kernels ALU-bound at a tenth of the HBM rate.
  unsigned char vec[LOSS_MAX_TERMS];
  int n;
};

// logical index (in units) -> element offset; d3 = innermost extent in units
F2G_SIMT_DEV long long loss_offset(const F2GLossTerm& t, const long long* __restrict__ stride, unsigned i, unsigned d3,
                                   unsigned unit) {
  const unsigned i3 = i % d3;
  i /= d3;
  const unsigned i2 = i % (unsigned)t.dims[2];
  i /= (unsigned)t.dims[2];
  const unsigned i1 = i % (unsigned)t.dims[1];
  const unsigned i0 = i / (unsigned)t.dims[1];
  return (long long)i0 * stride[0] + (long long)i1 * stride[1] + (long long)i2 * stride[2] + (long long)(i3 * unit) * stride[3];
}

F2G_SIMT_DEV int loss_find_term(const LossArgs& a, int block) {
  int ti = 0;
  while (ti + 1 < a.n && block >= a.block_begin[ti + 1]) ++ti;
  return ti;
}

F2G_SIMT_DEV float loss_elem_fwd(const F2GLossTerm& t, float av, float bv) {
  return t.mode == F2G_LOSS_L1 ? fabsf(av - bv) : fmaxf(1.0f + simt_fmul(t.sign, av), 0.0f);
}
F2G_SIMT_DEV float loss_elem_bwd(const F2GLossTerm& t, float av, float bv, float g) {
  if (t.mode == F2G_LOSS_L1) {
    const float diff = bv - av;                       // d|b - a| / db = sign(b - a), sign(0) = 0
    return diff > 0.f ? g : (diff < 0.f ? -g : 0.f);
  }
  // torch.clamp(x, min=0) passes the gradient where x >= 0
  return (1.0f + simt_fmul(t.sign, av) >= 0.0f) ? simt_fmul(t.sign, g) : 0.f;
}

// out[0] += sum over the launch's terms (out is zeroed by the entry point before the first group)
F2G_KERNEL void loss_terms_fwd_kernel(const F2G_GRID_CONSTANT LossArgs a, float* __restrict__ out) {
  const int ti = loss_find_term(a, (int)blockIdx.x);
  const F2GLossTerm& t = a.t[ti];
  const bool vec = a.vec[ti] != 0;
  const unsigned unit = vec ? 4u : 1u;
  const unsigned units = (unsigned)(t.numel / unit), d3 = (unsigned)t.dims[3] / unit;
  const unsigned base = (unsigned)((int)blockIdx.x - a.block_begin[ti]) * LOSS_CHUNK;
  const bool l1 = t.mode == F2G_LOSS_L1;
  float acc = 0.f;
  for (int k = 0; k < LOSS_PER_THREAD; ++k) {
    const unsigned i = base + (unsigned)k * LOSS_THREADS + threadIdx.x;
    if (i >= units) break;
    const float* ap = t.a + loss_offset(t, t.stride_a, i, d3, unit);
    const float* bp = l1 ? t.b + loss_offset(t, t.stride_b, i, d3, unit) : ap;
    if (vec) {
      const float4 av = *reinterpret_cast<const float4*>(ap);
      const float4 bv = *reinterpret_cast<const float4*>(bp);
      acc += (loss_elem_fwd(t, av.x, bv.x) + loss_elem_fwd(t, av.y, bv.y)) +
             (loss_elem_fwd(t, av.z, bv.z) + loss_elem_fwd(t, av.w, bv.w));
    } else {
      acc += loss_elem_fwd(t, *ap, *bp);
    }
  }
  simt_block_sum(simt_fmul(acc, t.scale), out);
}

// grad (contiguous, logical index order) of the summed loss w.r.t. b (L1) / a (hinge), times gout[0]
F2G_KERNEL void loss_terms_bwd_kernel(const F2G_GRID_CONSTANT LossArgs a, const float* __restrict__ gout) {
  const int ti = loss_find_term(a, (int)blockIdx.x);
  const F2GLossTerm& t = a.t[ti];
  const bool vec = a.vec[ti] != 0;
  const unsigned unit = vec ? 4u : 1u;
  const unsigned units = (unsigned)(t.numel / unit), d3 = (unsigned)t.dims[3] / unit;
  const unsigned base = (unsigned)((int)blockIdx.x - a.block_begin[ti]) * LOSS_CHUNK;
  const bool l1 = t.mode == F2G_LOSS_L1;
  const float g = simt_fmul(gout[0], t.scale);
  for (int k = 0; k < LOSS_PER_THREAD; ++k) {
    const unsigned i = base + (unsigned)k * LOSS_THREADS + threadIdx.x;
    if (i >= units) break;
    const float* ap = t.a + loss_offset(t, t.stride_a, i, d3, unit);
    const float* bp = l1 ? t.b + loss_offset(t, t.stride_b, i, d3, unit) : ap;
    if (vec) {
      const float4 av = *reinterpret_cast<const float4*>(ap);
      const float4 bv = *reinterpret_cast<const float4*>(bp);
      *reinterpret_cast<float4*>(t.grad + (size_t)i * 4) =
          make_float4(loss_elem_bwd(t, av.x, bv.x, g), loss_elem_bwd(t, av.y, bv.y, g),
                      loss_elem_bwd(t, av.z, bv.z, g), loss_elem_bwd(t, av.w, bv.w, g));
    } else {
      t.grad[i] = loss_elem_bwd(t, *ap, *bp, g);
    }
  }
}

static int loss_fill_args(const F2GLossTerm* terms, int n, int backward, LossArgs* a) {
  memset(a, 0, sizeof(*a));
  a->n = n;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    const F2GLossTerm& t = terms[i];
    if (t.numel <= 0 || !t.a || (t.mode == F2G_LOSS_L1 && !t.b) || (backward && !t.grad) ||
        (t.mode != F2G_LOSS_L1 && t.mode != F2G_LOSS_HINGE)) {
      set_error("f2g_loss_terms: term %d is malformed (numel %lld, mode %d)", i, t.numel, t.mode);
      return -1;
    }
    if ((long long)t.dims[0] * t.dims[1] * t.dims[2] * t.dims[3] != t.numel) {
      set_error("f2g_loss_terms: term %d dims do not multiply to numel", i);
      return -1;
    }
    if (t.numel >= (1ll << 31)) {
      set_error("f2g_loss_terms: term %d has %lld elements (limit 2^31 - 1)", i, t.numel);
      return -1;
    }
    bool vec = t.dims[3] % 4 == 0 && t.stride_a[3] == 1 && (reinterpret_cast<uintptr_t>(t.a) & 15) == 0 &&
               (!backward || (reinterpret_cast<uintptr_t>(t.grad) & 15) == 0);
    for (int d = 0; d < 3; ++d) vec = vec && t.stride_a[d] % 4 == 0;
    if (t.mode == F2G_LOSS_L1) {
      vec = vec && t.stride_b[3] == 1 && (reinterpret_cast<uintptr_t>(t.b) & 15) == 0;
      for (int d = 0; d < 3; ++d) vec = vec && t.stride_b[d] % 4 == 0;
    }
    a->t[i] = t;
    a->vec[i] = vec ? 1 : 0;
    a->block_begin[i] = blocks;
    const long long units = t.numel / (vec ? 4 : 1);
    blocks += (int)((units + LOSS_CHUNK - 1) / LOSS_CHUNK);
  }
  for (int i = n; i <= LOSS_MAX_TERMS; ++i) a->block_begin[i] = blocks;
  return blocks;
}

}  // namespace f2g

using namespace f2g;

extern "C" int f2g_loss_terms(const F2GLossTerm* terms, int n_terms, int backward, float* out,
                              const float* gout, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_terms < 1 || !terms || (!backward && !out) || (backward && !gout)) {
    set_error("f2g_loss_terms: bad arguments (n_terms %d)", n_terms);
    return F2G_EINVAL;
  }
  if (!backward) {
    if (int rc = simt_memset_async(out, 0, sizeof(float), stream)) return rc;
  }
  for (int first = 0; first < n_terms; first += LOSS_MAX_TERMS) {
    const int n = n_terms - first < LOSS_MAX_TERMS ? n_terms - first : LOSS_MAX_TERMS;
    LossArgs a;
    const int blocks = loss_fill_args(terms + first, n, backward, &a);
    if (blocks < 0) return F2G_EINVAL;
    if (backward)
      F2G_LAUNCH(loss_terms_bwd_kernel, blocks, LOSS_THREADS, stream, a, gout);
    else
      F2G_LAUNCH(loss_terms_fwd_kernel, blocks, LOSS_THREADS, stream, a, out);
    if (int rc = check_launch("f2g_loss_terms")) return rc;
  }
  return F2G_OK;
}
