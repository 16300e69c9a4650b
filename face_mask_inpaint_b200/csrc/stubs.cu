// Entry points declared in include/fmi_b200.h whose kernels are not written yet: fail loudly.
#include "common.cuh"

#define FMI_NOT_YET(name)                                   \
  do {                                                      \
    fmi_set_error(name ": kernel not implemented yet");     \
    return FMI_EINVAL;                                      \
  } while (0)

extern "C" int64_t fmi_modconv_workspace_bytes(int, int, int, int, int, int, int) { return -1; }
extern "C" int fmi_modconv_fwd(const void*, const float*, const float*, const float*, int, const float*, const float*,
                               const void*, const float*, void*, int, int, int, int, int, int, int, int, int, int, void*,
                               int64_t, void*) {
  FMI_NOT_YET("fmi_modconv_fwd");
}
