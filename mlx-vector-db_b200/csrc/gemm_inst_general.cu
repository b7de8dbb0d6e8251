// K3 kernel instantiations, GEN = 1 (euclidean keys 2 s - ||x||^2 and / or a row mask).
#define VS_GEMM_INSTANTIATE
#include "gemm_kernel.cuh"

namespace vs {
int launch_gemm_general(const GemmPlan& plan, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int grid,
                        cudaStream_t stream) {
  return launch_gemm_variant<1>(plan, mq, mx, p, grid, stream);
}
}  // namespace vs
