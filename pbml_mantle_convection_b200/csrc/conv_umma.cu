// tcgen05 / TMEM implicit-GEMM convolution (placeholder until the tensor-core path lands).
#include "common.cuh"

namespace pbmc {
bool conv_umma_supported(const pbmc_conv_desc&) { return false; }
int conv_umma_dispatch(const pbmc_conv_desc&, cudaStream_t) { return PBMC_ERR_UNSUPPORTED; }
}  // namespace pbmc
