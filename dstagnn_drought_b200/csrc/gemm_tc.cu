// tcgen05 tensor-core contraction engine (placeholder until the UMMA pipeline lands).
#include "common.cuh"

namespace dstag {

bool gemm_tc_supported(const dstag_gemm_t*) { return false; }

int gemm_tc(const dstag_gemm_t*, cudaStream_t) {
  set_error("dstag_gemm: tensor-core engine not available for this shape");
  return DSTAG_ERR_UNSUPPORTED;
}

}  // namespace dstag
