// lib.cu — library-level entry points of libvgslm: version, error string, device probe.
#include <stdarg.h>
#include "common.cuh"

namespace vg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int g_pdl_mode = 0;
}  // namespace vg

extern "C" int vg_version(void) { return VG_VERSION; }

extern "C" const char* vg_last_error_string(void) { return vg::g_err; }

extern "C" int vg_device_is_sm100(void) {
  int dev = 0;
  VG_CUDA(cudaGetDevice(&dev));
  int major = 0;
  VG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}

extern "C" int vg_set_pdl_mode(int mode) {
  VG_REQUIRE(mode >= 0 && mode <= 3, -3, "vg_set_pdl_mode: mode %d not in [0,3]", mode);
  vg::g_pdl_mode = mode;
  return 0;
}
