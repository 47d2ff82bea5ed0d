// TEST INFRASTRUCTURE ONLY -- a plain host restatement of the f2g_gemm_tf32 contract
// (include/flow2gan_b200.h, F2GGemm) so that the Python host layer can be dry-run end to end on a
// GPU-less box next to the host-emulated SIMT kernels (tests/_emul.py).  The product's contraction is
// the tcgen05 kernel of csrc/gemm_pair.cu / gemm_tf32.cu; nothing under flow2gan_b200/ links this.
// Operands are taken as stored (the callers round them to TF32 / fp16), products accumulate in
// double, the epilogue runs in fp32 in the documented order.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../include/flow2gan_b200.h"

static float load_half(const void* p, long long i) {
  _Float16 h;
  memcpy(&h, static_cast<const unsigned char*>(p) + 2 * i, 2);
  return (float)h;
}
static void store_half_sat(void* p, long long i, float v) {
  v = fminf(fmaxf(v, -65504.0f), 65504.0f);
  const _Float16 h = (_Float16)v;
  memcpy(static_cast<unsigned char*>(p) + 2 * i, &h, 2);
}
static float tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

static float a_elem(const F2GGemm& d, int m, int k) {
  if (d.ab_f16) return load_half(d.a, (long long)m * d.lda + k);
  if (!d.a_seg_len) return d.a_mn ? d.a[(long long)k * d.lda + m] : d.a[(long long)m * d.lda + k];
  long long row, col;
  if (!d.a_mn) {
    row = m + (long long)(k / d.a_seg_len) * d.a_seg_shift;
    col = k % d.a_seg_len;
  } else {
    row = k + (long long)(m / d.a_seg_len) * d.a_seg_shift;
    col = m % d.a_seg_len;
  }
  if (row < 0 || row >= d.a_rows) return 0.f;           // TMA out-of-bounds fill
  return d.a[row * d.lda + col];
}
static float b_elem(const F2GGemm& d, int n, int k) {
  if (d.ab_f16) return load_half(d.b, (long long)n * d.ldb + k);
  return d.b_mn ? d.b[(long long)k * d.ldb + n] : d.b[(long long)n * d.ldb + k];
}

extern "C" int f2g_gemm_tf32(const F2GGemm* problems, int n_problems, void*) {
  if (n_problems < 1 || n_problems > F2G_GEMM_MAX_PROBLEMS) return -1;
  for (int pi = 0; pi < n_problems; ++pi) {        // producers are listed before their consumers
    const F2GGemm& d = problems[pi];
    const float alpha = d.alpha == 0.f ? 1.f : d.alpha;
    std::vector<float> arow((size_t)d.K), bmat((size_t)d.N * d.K);
    for (int n = 0; n < d.N; ++n)
      for (int k = 0; k < d.K; ++k) bmat[(size_t)n * d.K + k] = b_elem(d, n, k);
#pragma omp parallel for schedule(static) firstprivate(arow)
    for (int m = 0; m < d.M; ++m) {
      for (int k = 0; k < d.K; ++k) arow[k] = a_elem(d, m, k);
      for (int n = 0; n < d.N; ++n) {
        const float* bp = &bmat[(size_t)n * d.K];
        double acc = 0.0;
        for (int k = 0; k < d.K; ++k) acc += (double)arow[k] * (double)bp[k];
        float x = alpha * (float)acc + (d.bias ? d.bias[n] : 0.f);
        if (d.c_pre) d.c_pre[(long long)m * d.ld_pre + n] = x;
        if (d.act == F2G_ACT_PRELU) x = x > 0.f ? x : x * d.slope[n];
        else if (d.act == F2G_ACT_LEAKY) x = x > 0.f ? x : x * d.leaky;
        else if (d.act == F2G_ACT_SILU) x = x / (1.f + expf(-x));
        if (d.gate) x *= d.gate[(long long)m * d.ld_gate + n] > 0.f ? 1.f : d.slope[n];
        if (d.res) x += (d.res_scale ? d.res_scale[n] : 1.f) * d.res[(long long)m * d.ld_res + n];
        if (d.row_scale) x *= d.row_scale[m];
        if (d.c_f16) {
          store_half_sat(d.c, (long long)m * d.ldc + n, x);
          if (d.sat_flag && !(fabsf(x) <= 65504.f)) {      // fp16 range guard (F2GGemm::sat_flag, bit 0)
#pragma omp atomic
            *d.sat_flag |= 1;
          }
          continue;
        }
        float* cp = d.c + (long long)m * d.ldc + n;
        if (d.accumulate || d.split_k > 1) x += *cp;
        *cp = d.round_tf32 ? tf32_rna(x) : x;
      }
    }
  }
  return 0;
}
