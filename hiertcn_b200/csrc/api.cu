// ABI bookkeeping: version, thread-local error string, device probe.
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace htcn {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return HTCN_ERR_CUDA;
}
}  // namespace htcn

extern "C" {
int32_t htcn_abi_version(void) { return HTCN_ABI_VERSION; }
const char* htcn_last_error(void) { return htcn::g_err; }
int32_t htcn_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}
}
