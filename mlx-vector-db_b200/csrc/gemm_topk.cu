// K3 placeholder: the tcgen05 path is not wired in yet; AUTO never selects it.
#include "gemm_topk.cuh"
#include "store.cuh"

namespace vs {

bool gemm_supported(const vs_store*, int64_t, int, int) { return false; }

int gemm_path(vs_store*, int64_t, const float*, int, int, bool, bool, float*, int32_t*, int64_t,
              cudaStream_t) {
  set_error("the GEMM search path is not available in this build");
  return VS_ERR_STATE;
}

}  // namespace vs
