// a5: fused bias + leaky-ReLU (and its backward with the bias-gradient reduction fused in).
// HBM-bound streaming kernels: 16-byte vector loads, 4 vectors in flight per thread,
// persistent grid sized in multiples of the SM count.
//
// Semantics follow modules/psp/stylegan2/op/fused_bias_act_kernel.cu:18-49 (act*10+grad switch,
// bias index (i/step_b)%size_b, scale applied last) and op/fused_act.py:18-38 (backward).
#include <type_traits>

#include "common.cuh"

namespace {

// The reference adds the bias in scalar_t (fused_bias_act_kernel.cu:28-30): when bias and
// activation share a 16-bit type the sum is rounded to it; fp32 biases keep fp32.
template <typename T, typename BT> __device__ __forceinline__ float add_bias(float x, float b) {
  if constexpr (std::is_same<T, BT>::value) return to_f32<T>(from_f32<T>(x + b));
  else return x + b;
}

__device__ __forceinline__ float act_apply(float x, float ref, int code, float alpha) {
  // code = act*10 + grad (fused_bias_act_kernel.cu:36-45)
  switch (code) {
    case 30: return x > 0.f ? x : x * alpha;
    case 31: return ref > 0.f ? x : x * alpha;
    case 12:
    case 32: return 0.f;
    default: return x;  // 10, 11 and the reference's `default`
  }
}

template <typename T, typename BT, bool kVecBias>
__global__ void __launch_bounds__(256) bias_act_vec_kernel(const T* __restrict__ x, const BT* __restrict__ b,
                                                           const T* __restrict__ ref, T* __restrict__ y,
                                                           int code, float alpha, float scale, int64_t n_vec,
                                                           int64_t step_b, int64_t size_b) {
  constexpr int VN = Vec16<T>::N;
  constexpr int UNROLL = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; v < n_vec; v += stride * UNROLL) {
    Vec16<T> xv[UNROLL], rv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int64_t vi = v + u * stride;
      if (vi < n_vec) {
        xv[u] = ld_vec16_stream(x + vi * VN);
        if (ref) rv[u] = ld_vec16_stream(ref + vi * VN);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int64_t vi = v + u * stride;
      if (vi >= n_vec) continue;
      int64_t base = vi * VN;
      float bias0 = 0.f;
      if (b && kVecBias) bias0 = to_f32<BT>(b[(base / step_b) % size_b]);
      Vec16<T> out;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        float xf = to_f32<T>(xv[u].e[e]);
        if (b) {
          float bb = kVecBias ? bias0 : to_f32<BT>(b[((base + e) / step_b) % size_b]);
          xf = add_bias<T, BT>(xf, bb);
        }
        float rf = ref ? to_f32<T>(rv[u].e[e]) : 0.f;
        out.e[e] = from_f32<T>(act_apply(xf, rf, code, alpha) * scale);
      }
      st_vec16(y + base, out);
    }
  }
}

template <typename T, typename BT>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(const T* __restrict__ x, const BT* __restrict__ b,
                                                              const T* __restrict__ ref, T* __restrict__ y,
                                                              int code, float alpha, float scale, int64_t begin,
                                                              int64_t n, int64_t step_b, int64_t size_b) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float xf = to_f32<T>(x[i]);
    if (b) xf = add_bias<T, BT>(xf, to_f32<BT>(b[(i / step_b) % size_b]));
    float rf = ref ? to_f32<T>(ref[i]) : 0.f;
    y[i] = from_f32<T>(act_apply(xf, rf, code, alpha) * scale);
  }
}

// ---- backward ---------------------------------------------------------------------------------
// One block per (row, chunk): row = (outer, channel) plane of `inner` contiguous elements.
template <typename T>
__global__ void __launch_bounds__(256) bias_act_bwd_block_kernel(const T* __restrict__ g, const T* __restrict__ out,
                                                                 T* __restrict__ gx, float* __restrict__ gb,
                                                                 float alpha, float scale, int64_t inner,
                                                                 int64_t size_b, int chunks_per_row, int64_t chunk) {
  constexpr int VN = Vec16<T>::N;
  const int64_t row = blockIdx.x / chunks_per_row;
  const int ck = blockIdx.x % chunks_per_row;
  const int64_t lo = ck * chunk;
  const int64_t hi = min(inner, lo + chunk);
  const T* gr = g + row * inner;
  const T* outr = out + row * inner;
  T* gxr = gx + row * inner;
  float acc = 0.f;
  // rows are 16-byte aligned when inner % VN == 0 (checked on the host)
  for (int64_t i = lo + (int64_t)threadIdx.x * VN; i < hi; i += 256 * VN) {
    if (i + VN <= hi) {
      Vec16<T> gv = ld_vec16_stream(gr + i), ov = ld_vec16_stream(outr + i), r;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        float v = to_f32<T>(gv.e[e]);
        v = (to_f32<T>(ov.e[e]) > 0.f ? v : v * alpha) * scale;
        r.e[e] = from_f32<T>(v);
        acc += to_f32<T>(r.e[e]);
      }
      st_vec16(gxr + i, r);
    } else {
      for (int64_t j = i; j < hi; ++j) {
        float v = to_f32<T>(gr[j]);
        v = (to_f32<T>(outr[j]) > 0.f ? v : v * alpha) * scale;
        T r = from_f32<T>(v);
        gxr[j] = r;
        acc += to_f32<T>(r);
      }
    }
  }
  if (gb) {
    __shared__ float part[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float s = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
      s = warp_sum(s);
      if (threadIdx.x == 0) atomicAdd(gb + (row % size_b), s);
    }
  }
}

// One warp per row (small planes and the 2-D [N,C] case where inner == 1).
template <typename T>
__global__ void __launch_bounds__(256) bias_act_bwd_warp_kernel(const T* __restrict__ g, const T* __restrict__ out,
                                                                T* __restrict__ gx, float* __restrict__ gb,
                                                                float alpha, float scale, int64_t rows,
                                                                int64_t inner, int64_t size_b) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    float acc = 0.f;
    for (int64_t i = lane; i < inner; i += 32) {
      int64_t idx = row * inner + i;
      float v = to_f32<T>(g[idx]);
      v = (to_f32<T>(out[idx]) > 0.f ? v : v * alpha) * scale;
      T r = from_f32<T>(v);
      gx[idx] = r;
      acc += to_f32<T>(r);
    }
    if (gb) {
      acc = warp_sum(acc);
      if (lane == 0) atomicAdd(gb + (row % size_b), acc);
    }
  }
}

template <typename T, typename BT>
int launch_fwd(const void* x, const void* b, const void* ref, void* y, int code, float alpha, float scale,
               int64_t size_x, int64_t step_b, int64_t size_b, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  const T* xp = (const T*)x;
  const T* rp = (const T*)ref;
  const BT* bp = (const BT*)b;
  T* yp = (T*)y;
  bool vec_ok = fmi_aligned(x, 16) && fmi_aligned(y, 16) && (!ref || fmi_aligned(ref, 16));
  int64_t n_vec = vec_ok ? size_x / VN : 0;
  if (n_vec > 0) {
    int64_t want = (n_vec + 256 * 4 - 1) / (256 * 4);
    int grid = (int)imin64(want, (int64_t)FMI_NUM_SMS * 16);
    if (!b || step_b % VN == 0)
      bias_act_vec_kernel<T, BT, true><<<grid, 256, 0, st>>>(xp, bp, rp, yp, code, alpha, scale, n_vec, step_b, size_b);
    else
      bias_act_vec_kernel<T, BT, false><<<grid, 256, 0, st>>>(xp, bp, rp, yp, code, alpha, scale, n_vec, step_b, size_b);
  }
  int64_t tail_begin = n_vec * VN;
  if (tail_begin < size_x) {
    int64_t n_tail = size_x - tail_begin;
    int grid = (int)imin64((n_tail + 255) / 256, (int64_t)FMI_NUM_SMS * 16);
    bias_act_scalar_kernel<T, BT><<<grid, 256, 0, st>>>(xp, bp, rp, yp, code, alpha, scale, tail_begin, size_x, step_b, size_b);
  }
  return fmi_launched("fused_bias_act");
}

template <typename T>
int launch_bwd(const void* grad_out, const void* out, void* grad_in, float* grad_bias, float alpha, float scale,
               int64_t rows, int64_t inner, int64_t size_b, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  bool vec_ok = inner % VN == 0 && fmi_aligned(grad_out, 16) && fmi_aligned(out, 16) && fmi_aligned(grad_in, 16);
  if (inner >= 2048 && vec_ok) {
    const int64_t chunk = 256 * VN * 8;
    int chunks_per_row = (int)((inner + chunk - 1) / chunk);
    int64_t blocks = rows * chunks_per_row;
    FMI_REQUIRE(blocks < (1ll << 31), "bias_act_bwd: tensor too large");
    bias_act_bwd_block_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)grad_out, (const T*)out, (T*)grad_in,
                                                                     grad_bias, alpha, scale, inner, size_b,
                                                                     chunks_per_row, chunk);
  } else {
    int grid = (int)imin64((rows + 7) / 8, (int64_t)FMI_NUM_SMS * 16);
    bias_act_bwd_warp_kernel<T><<<grid, 256, 0, st>>>((const T*)grad_out, (const T*)out, (T*)grad_in, grad_bias, alpha,
                                                        scale, rows, inner, size_b);
  }
  return fmi_launched("bias_act_bwd");
}

}  // namespace

extern "C" int fmi_fused_bias_act(const void* x, const void* b, const void* ref, void* y, int act, int grad,
                                  float alpha, float scale, int64_t size_x, int64_t step_b, int64_t size_b,
                                  int dtype, int bias_dtype, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype), "fused_bias_act: unsupported dtype %d", dtype);
  FMI_REQUIRE(size_x >= 0, "fused_bias_act: negative size");
  if (size_x == 0) return FMI_OK;
  FMI_REQUIRE(x && y, "fused_bias_act: null input/output");
  FMI_REQUIRE(act == 1 || act == 3, "fused_bias_act: act must be 1 (linear) or 3 (lrelu), got %d", act);
  FMI_REQUIRE(grad >= 0 && grad <= 2, "fused_bias_act: grad must be 0..2, got %d", grad);
  if (b) {
    FMI_REQUIRE(step_b >= 1 && size_b >= 1, "fused_bias_act: bad bias geometry step_b=%lld size_b=%lld",
                (long long)step_b, (long long)size_b);
    FMI_REQUIRE(bias_dtype == dtype || bias_dtype == FMI_F32, "fused_bias_act: bias dtype must equal dtype or be fp32");
  } else {
    step_b = 1;
    size_b = 1;
  }
  const int code = act * 10 + grad;
  cudaStream_t st = (cudaStream_t)stream;
  if (!b || bias_dtype == FMI_F32) {
    FMI_DISPATCH_DTYPE(dtype, T, return (launch_fwd<T, float>(x, b, ref, y, code, alpha, scale, size_x, step_b, size_b, st)));
  } else {
    FMI_DISPATCH_DTYPE(dtype, T, return (launch_fwd<T, T>(x, b, ref, y, code, alpha, scale, size_x, step_b, size_b, st)));
  }
  return FMI_OK;
}

extern "C" int fmi_bias_act_bwd(const void* grad_out, const void* out, void* grad_in, float* grad_bias, float alpha,
                                float scale, int64_t size_x, int64_t step_b, int64_t size_b, int dtype,
                                void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype), "bias_act_bwd: unsupported dtype %d", dtype);
  if (size_x == 0) return FMI_OK;
  FMI_REQUIRE(grad_out && out && grad_in, "bias_act_bwd: null pointer");
  FMI_REQUIRE(step_b >= 1 && size_b >= 1 && size_x % step_b == 0, "bias_act_bwd: bad geometry");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t inner = step_b;
  const int64_t rows = size_x / inner;
  FMI_DISPATCH_DTYPE(dtype, T,
                     return (launch_bwd<T>(grad_out, out, grad_in, grad_bias, alpha, scale, rows, inner, size_b, st)));
  return FMI_OK;
}
