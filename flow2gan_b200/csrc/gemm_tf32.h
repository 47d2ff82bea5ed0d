// internal: host entry of the grouped tcgen05 TF32 GEMM
#pragma once
#include <cuda_runtime.h>
#include <string.h>
#include "../../include/flow2gan_b200.h"

namespace f2g {
int gemm_tf32_group(const F2GGemm* descs, int n, cudaStream_t stream);
// CTA-pair (cta_group::2, 256-row tiles) variant, gemm_pair.cu; the default path
int gemm_pair_group(const F2GGemm* descs, int n, cudaStream_t stream);
}
