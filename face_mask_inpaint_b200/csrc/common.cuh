// Shared host/device helpers for the fmi_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fmi_b200.h"

#define FMI_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- error plumbing -------------------------------------------------------------------------
void fmi_set_error(const char* fmt, ...);
int fmi_check_cuda(cudaError_t e, const char* what);

// Called right after every kernel launch: counts it (fmi_kernel_launch_count) and surfaces launch errors.
int fmi_launched(const char* kernel_name);
// fmi_set_tf32_exact: TF32-mode producers keep exact fp32 values (error-compensated 3xTF32 operands, fmi_tf32_split3)
bool fmi_tf32_exact_on();

// Optional CUDA-event timing of kernels on their launch stream (fmi_profile_enable / _collect / _dump). A scope brackets one
// launch (or one entry point's launches) with two events and carries the ALGORITHMIC work of that launch — the FLOPs and the
// minimum HBM bytes (inputs read once + outputs written once) — so bench.py can put every launch on its roofline.
// kinds: see g_prof_names in core.cu (0 attention main kernel, 1 implicit-GEMM conv, 2 attention fallback kernel, ...).
#define FMI_PROF_KINDS 16
enum {
  FMI_PROF_ATTN = 0, FMI_PROF_GEMM = 1, FMI_PROF_ATTN_FALLBACK = 2, FMI_PROF_OUTCONV = 3, FMI_PROF_INSTATS = 4,
  FMI_PROF_NORMACT = 5, FMI_PROF_WGRAD = 6, FMI_PROF_ATTN_BWD = 7, FMI_PROF_UPFIRDN = 8, FMI_PROF_BIASACT = 9,
  FMI_PROF_BLURACT = 10, FMI_PROF_GEMM_IR = 11, FMI_PROF_SE = 12, FMI_PROF_STREAM = 13, FMI_PROF_ATTN_PRO = 14,
  FMI_PROF_TORGB = 15
};
struct FmiProfScope {
  FmiProfScope(int kind, cudaStream_t st, double flops = 0.0, double bytes = 0.0);
  ~FmiProfScope();
  int kind_;
  cudaStream_t st_;
  cudaEvent_t e0_, e1_;
  bool on_;
  double flops_, bytes_;
};

// One-time per-DEVICE setup of a kernel (cudaFuncSetAttribute is per device): a process that drives several GPUs must opt every
// one of them into the larger dynamic shared memory (a per-process flag left the second device at the default limit).
#include <atomic>
struct FmiPerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  static unsigned long long bit() {
    int d = 0;
    cudaGetDevice(&d);
    return 1ull << (d & 63);
  }
  bool need() const { return !(mask.load(std::memory_order_acquire) & bit()); }
  void done() { mask.fetch_or(bit(), std::memory_order_release); }
};

#define FMI_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      fmi_set_error(__VA_ARGS__);   \
      return FMI_EINVAL;            \
    }                               \
  } while (0)

#define FMI_CUDA(expr)                              \
  do {                                              \
    int _rc = fmi_check_cuda((expr), #expr);        \
    if (_rc) return _rc;                            \
  } while (0)

#define FMI_LAUNCH_CHECK(name) FMI_CUDA((cudaPeekAtLastError(), cudaGetLastError()))

static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int fmi_dtype_size(int dtype) { return dtype == FMI_F32 ? 4 : 2; }
static inline bool fmi_dtype_ok(int dtype) { return dtype == FMI_F32 || dtype == FMI_BF16 || dtype == FMI_F16; }
static inline bool fmi_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

__device__ __forceinline__ bool fmi_aligned_dev(const void* p, size_t a) {
  return (reinterpret_cast<uintptr_t>(p) % a) == 0;
}

// ---- element conversion ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// 16-byte vector of T
template <typename T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 u;
    T e[16 / sizeof(T)];
  };
};

template <typename T> __device__ __forceinline__ Vec16<T> ld_vec16(const T* p) {
  Vec16<T> v;
  v.u = *reinterpret_cast<const uint4*>(p);
  return v;
}
// streaming (read-once) 16-byte load that does not allocate in L1
template <typename T> __device__ __forceinline__ Vec16<T> ld_vec16_stream(const T* p) {
  Vec16<T> v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.u.x), "=r"(v.u.y), "=r"(v.u.z), "=r"(v.u.w)
               : "l"(p));
  return v;
}
template <typename T> __device__ __forceinline__ void st_vec16(T* p, const Vec16<T>& v) {
  *reinterpret_cast<uint4*>(p) = v.u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// dispatch on the runtime dtype code
#define FMI_DISPATCH_DTYPE(dtype, T, ...)              \
  switch (dtype) {                                     \
    case FMI_F32: { using T = float; __VA_ARGS__; } break;          \
    case FMI_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
    case FMI_F16: { using T = __half; __VA_ARGS__; } break;         \
    default: fmi_set_error("unsupported dtype %d", dtype); return FMI_EINVAL; \
  }
