// K3 kernel instantiations, GEN = 0 (cosine / dot_product keys, no row mask).
#define VS_GEMM_INSTANTIATE
#include "gemm_kernel.cuh"

namespace vs {
int launch_gemm_plain(const GemmPlan& plan, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int grid,
                      cudaStream_t stream) {
  return launch_gemm_variant<0>(plan, mq, mx, p, grid, stream);
}
}  // namespace vs
