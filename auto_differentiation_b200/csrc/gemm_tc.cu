// tcgen05 / TMA GEMM — placeholder until the tensor-core kernel lands (reports "unsupported" so that
// B200AD_GEMM_AUTO resolves to the fp32 SIMT kernel).
#include "common.cuh"
#include "gemm_internal.cuh"

namespace b200ad {
bool gemm_tc_supported(const GemmArgs&, int, const char** why) {
  if (why) *why = "tcgen05 path not built";
  return false;
}
int64_t gemm_tc_workspace_bytes(const GemmArgs&, int) { return 0; }
int gemm_tc_launch(const GemmArgs&, int, void*, int64_t, cudaStream_t) {
  return set_error(B200AD_ERR_UNSUPPORTED, "gemm(tcgen05): not built");
}
}  // namespace b200ad
