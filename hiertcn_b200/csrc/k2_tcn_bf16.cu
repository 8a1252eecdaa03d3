// placeholder until the tcgen05 conv stack lands: loud failure, never a fallback
#include "common.cuh"
namespace htcn {
int32_t tcn_forward_bf16(const void*, int, const float*, const float*, const float* const*, const float* const*, int,
                         int, const SlotTable&, int, int, const int*, void*, int, float*, cudaStream_t) {
  set_error("tcn_forward: the bf16 (tcgen05) conv stack is not built; use precision HTCN_F32");
  return HTCN_ERR_UNSUPPORTED;
}
}  // namespace htcn
