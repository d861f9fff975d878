// Fused tcgen05 news-encoder forward (inference).  Placeholder until the fused kernel lands:
// reports "not applicable" so encoder.cu takes the decomposed path with tc_gemm_nt.
#include "tc_common.cuh"
#include "tc_api.cuh"

namespace nrms {

size_t tc_fused_workspace_bytes(int64_t, int) { return (size_t)-1; }

int tc_news_encoder_fused(const int64_t*, int64_t, const float*, int64_t, const float*, const float*, const float*,
                          const float*, const float*, float*, void*, size_t, cudaStream_t) {
  set_error("fused tcgen05 news encoder not built");
  return NRMS_E_UNSUPPORTED;
}

}  // namespace nrms
