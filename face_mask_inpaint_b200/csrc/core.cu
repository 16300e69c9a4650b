// Error plumbing, version and device check of the fmi_b200 C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void fmi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fmi_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return FMI_OK;
  fmi_set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return FMI_ECUDA;
}

extern "C" int fmi_version(void) { return 100; }

extern "C" const char* fmi_last_error(void) { return g_err; }

extern "C" int fmi_device_check(void) {
  int dev = -1;
  FMI_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  FMI_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FMI_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    fmi_set_error("device %d is sm_%d%d; fmi_b200 kernels are built for sm_100a only", dev, major, minor);
    return FMI_EARCH;
  }
  return FMI_OK;
}
