// placeholder until the tcgen05 tier lands (next commit): loud failure, never a fallback
#include "common.cuh"
namespace htcn {
int32_t score_bf16(const ScoreArgs&, cudaStream_t) {
  set_error("score: bf16 tier not built");
  return HTCN_ERR_UNSUPPORTED;
}
int32_t target_logit_bf16(const void*, const void*, const float*, const int*, int, int, int, float*, cudaStream_t) {
  set_error("target_logit: bf16 tier not built");
  return HTCN_ERR_UNSUPPORTED;
}
int32_t tcn_forward_bf16(const void*, int, const float*, const float*, const float* const*, const float* const*, int,
                         int, const SlotTable&, int, int, const int*, void*, int, float*, cudaStream_t) {
  set_error("tcn_forward: bf16 tier not built");
  return HTCN_ERR_UNSUPPORTED;
}
}  // namespace htcn
