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

// ---- launch accounting and optional CUDA-event profiling of the dominant kernels ------------------------------
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfRec {
  cudaEvent_t e0, e1;
  double flops, bytes;
};
static std::vector<ProfRec> g_prof_events[FMI_PROF_KINDS];
static const char* const g_prof_names[FMI_PROF_KINDS] = {
    "attn_fwd", "conv_gemm", "attn_fwd_fallback", "out_conv_tanh", "instnorm_stats", "norm_act", "wgrad_gemm", "attn_bwd",
    "upfirdn2d", "bias_act", "blur_act_nhwc", "conv_gemm_ir", "se_pool_scale", "streaming_other", "attn_prologue", "torgb"};

int fmi_launched(const char* kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return FMI_OK;
  fmi_set_error("launch of %s failed: CUDA error %d (%s)", kernel_name, (int)e, cudaGetErrorString(e));
  return FMI_ECUDA;
}

FmiProfScope::FmiProfScope(int kind, cudaStream_t st, double flops, double bytes)
    : kind_(kind), st_(st), e0_(nullptr), e1_(nullptr), on_(false), flops_(flops), bytes_(bytes) {
  if (!g_prof_on.load(std::memory_order_relaxed) || kind < 0 || kind >= FMI_PROF_KINDS) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st_, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;  // no timing inside a graph
  if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) return;
  on_ = true;
  cudaEventRecord(e0_, st_);
}
FmiProfScope::~FmiProfScope() {
  if (!on_) return;
  cudaEventRecord(e1_, st_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_events[kind_].push_back(ProfRec{e0_, e1_, flops_, bytes_});
}

extern "C" long long fmi_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int fmi_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return FMI_OK;
}

extern "C" int fmi_profile_kinds(void) { return FMI_PROF_KINDS; }

extern "C" const char* fmi_profile_kind_name(int kind) {
  return (kind >= 0 && kind < FMI_PROF_KINDS) ? g_prof_names[kind] : "";
}

// Synchronises on the recorded events; returns the summed duration (ms) and number of launches of `kind`.
extern "C" int fmi_profile_collect(int kind, double* total_ms, int* launches) {
  FMI_REQUIRE(kind >= 0 && kind < FMI_PROF_KINDS && total_ms && launches, "profile_collect: bad arguments");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double sum = 0;
  int n = 0;
  for (auto& pr : g_prof_events[kind]) {
    float ms = 0.f;
    cudaEventSynchronize(pr.e1);
    if (cudaEventElapsedTime(&ms, pr.e0, pr.e1) == cudaSuccess) {
      sum += ms;
      ++n;
    }
    cudaEventDestroy(pr.e0);
    cudaEventDestroy(pr.e1);
  }
  g_prof_events[kind].clear();
  *total_ms = sum;
  *launches = n;
  return FMI_OK;
}

// Per-launch records of `kind` in launch order: duration (ms), algorithmic FLOPs and algorithmic bytes the launcher stated.
// Writes at most `cap` records, returns the number written in *n, and clears the record of that kind.
extern "C" int fmi_profile_dump(int kind, double* ms_out, double* flops_out, double* bytes_out, int cap, int* n) {
  FMI_REQUIRE(kind >= 0 && kind < FMI_PROF_KINDS && n && (cap == 0 || (ms_out && flops_out && bytes_out)),
              "profile_dump: bad arguments");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int k = 0;
  for (auto& pr : g_prof_events[kind]) {
    float ms = 0.f;
    cudaEventSynchronize(pr.e1);
    const bool ok = cudaEventElapsedTime(&ms, pr.e0, pr.e1) == cudaSuccess;
    if (ok && k < cap) {
      ms_out[k] = ms;
      flops_out[k] = pr.flops;
      bytes_out[k] = pr.bytes;
      ++k;
    }
    cudaEventDestroy(pr.e0);
    cudaEventDestroy(pr.e1);
  }
  g_prof_events[kind].clear();
  *n = k;
  return FMI_OK;
}

extern "C" int fmi_version(void) { return 100; }

extern "C" const char* fmi_last_error(void) { return g_err; }

// ---- error-compensated TF32 ("3xTF32") switch ---------------------------------------------------------------------------------
// Off (default): in FMI_MMA_TF32 mode the producers of GEMM operands (weight re-layouts, normalise + activate passes) round to
// tf32. On: they keep the exact fp32 value, and the caller feeds each GEMM the split operands of fmi_tf32_split3 instead.
static std::atomic<int> g_tf32_exact{0};
bool fmi_tf32_exact_on() { return g_tf32_exact.load(std::memory_order_relaxed) != 0; }
extern "C" int fmi_set_tf32_exact(int on) { return g_tf32_exact.exchange(on ? 1 : 0); }
extern "C" int fmi_get_tf32_exact(void) { return g_tf32_exact.load(); }

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
