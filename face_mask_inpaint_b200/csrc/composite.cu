// a7: masked source/reference compositing with the bilinear (align_corners=True) mask resize fused in.
//   scale_img  : modules/model.py:10-12  (F.interpolate bilinear, align_corners=True)
//   blend      : modules/model.py:99     enc = (1 - m) * src + m * ref
//                modules/psp/encoders/psp_encoders.py:135-138   c = m * r + (1 - m) * c
// HBM-bound: per element 2 reads + 1 write; the mask is sampled once per pixel vector and reused
// across a group of channels held in flight (coalesced 16-byte accesses along W).
#include "common.cuh"

namespace {

struct Bilin {
  int i0, i1;
  float l0, l1;
};

// ATen's align_corners=True source index: scale = (in-1)/(out-1) (0 when out == 1); src = scale * dst.
__device__ __forceinline__ Bilin bilin_coord(int dst, int in_size, float scale) {
  Bilin b;
  float r = scale * (float)dst;
  b.i0 = (int)r;
  if (b.i0 > in_size - 1) b.i0 = in_size - 1;
  b.i1 = b.i0 + ((b.i0 < in_size - 1) ? 1 : 0);
  b.l1 = r - (float)b.i0;
  b.l0 = 1.f - b.l1;
  return b;
}

__device__ __forceinline__ float sample_mask(const float* __restrict__ m, int Wm, const Bilin& by, const Bilin& bx) {
  const float* r0 = m + (int64_t)by.i0 * Wm;
  const float* r1 = m + (int64_t)by.i1 * Wm;
  return by.l0 * (bx.l0 * __ldg(r0 + bx.i0) + bx.l1 * __ldg(r0 + bx.i1)) +
         by.l1 * (bx.l0 * __ldg(r1 + bx.i0) + bx.l1 * __ldg(r1 + bx.i1));
}

__global__ void __launch_bounds__(256) scale_mask_kernel(const float* __restrict__ mask, float* __restrict__ out, int N,
                                                         int Hm, int Wm, int H, int W, float sh, float sw) {
  int64_t total = (int64_t)N * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    int h = (int)((i / W) % H);
    int n = (int)(i / ((int64_t)W * H));
    Bilin by = bilin_coord(h, Hm, sh), bx = bilin_coord(w, Wm, sw);
    out[i] = sample_mask(mask + (int64_t)n * Hm * Wm, Wm, by, bx);
  }
}

constexpr int CG = 8;  // channels per thread (loads kept in flight)

// MODE 0: out = (1-m)*a + m*b        (forward; a = src, b = ref)
// MODE 1: ga = (1-m)*g, gb = m*g     (backward; a = grad_out, outputs o0/o1 nullable)
template <typename T, int VN, int MODE>
__global__ void __launch_bounds__(256) composite_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                        const float* __restrict__ mask, T* __restrict__ o0,
                                                        T* __restrict__ o1, int N, int C, int H, int W, int Hm, int Wm,
                                                        float sh, float sw) {
  const int Wv = W / VN;
  const int64_t pix_vecs = (int64_t)N * H * Wv;
  const int c_groups = (C + CG - 1) / CG;
  const int64_t total = pix_vecs * c_groups;
  const int64_t HW = (int64_t)H * W;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    // pixel vectors vary fastest so a warp touches contiguous addresses for a fixed channel
    const int64_t pv = t % pix_vecs;
    const int cg = (int)(t / pix_vecs);
    const int wv = (int)(pv % Wv);
    const int h = (int)((pv / Wv) % H);
    const int n = (int)(pv / ((int64_t)Wv * H));
    const float* mp = mask + (int64_t)n * Hm * Wm;
    Bilin by = bilin_coord(h, Hm, sh);
    float m[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) m[e] = sample_mask(mp, Wm, by, bilin_coord(wv * VN + e, Wm, sw));
    const int c0 = cg * CG;
    const int64_t base = ((int64_t)n * C + c0) * HW + (int64_t)h * W + wv * VN;
    if (MODE == 0) {
      __align__(16) T av[CG][VN];
      __align__(16) T bv[CG][VN];
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        if (c0 + j >= C) continue;
        if (VN * sizeof(T) == 16) {
          *reinterpret_cast<uint4*>(av[j]) = ld_vec16_stream(a + base + j * HW).u;
          *reinterpret_cast<uint4*>(bv[j]) = ld_vec16_stream(b + base + j * HW).u;
        } else {
          av[j][0] = a[base + j * HW];
          bv[j][0] = b[base + j * HW];
        }
      }
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        if (c0 + j >= C) continue;
        __align__(16) T ov[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          // separate roundings, as the reference's mul/mul/add kernels do
          float t1 = __fmul_rn(__fsub_rn(1.f, m[e]), to_f32<T>(av[j][e]));
          float t2 = __fmul_rn(m[e], to_f32<T>(bv[j][e]));
          ov[e] = from_f32<T>(__fadd_rn(t1, t2));
        }
        if (VN * sizeof(T) == 16) *reinterpret_cast<uint4*>(o0 + base + j * HW) = *reinterpret_cast<uint4*>(ov);
        else o0[base + j * HW] = ov[0];
      }
    } else {
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        if (c0 + j >= C) continue;
        __align__(16) T gv[VN];
        __align__(16) T r0[VN];
        __align__(16) T r1[VN];
        if (VN * sizeof(T) == 16) *reinterpret_cast<uint4*>(gv) = ld_vec16_stream(a + base + j * HW).u;
        else gv[0] = a[base + j * HW];
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          float g = to_f32<T>(gv[e]);
          r0[e] = from_f32<T>(__fmul_rn(__fsub_rn(1.f, m[e]), g));
          r1[e] = from_f32<T>(__fmul_rn(m[e], g));
        }
        if (VN * sizeof(T) == 16) {
          if (o0) *reinterpret_cast<uint4*>(o0 + base + j * HW) = *reinterpret_cast<uint4*>(r0);
          if (o1) *reinterpret_cast<uint4*>(o1 + base + j * HW) = *reinterpret_cast<uint4*>(r1);
        } else {
          if (o0) o0[base + j * HW] = r0[0];
          if (o1) o1[base + j * HW] = r1[0];
        }
      }
    }
  }
}

inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

template <typename T, int MODE>
int launch_composite(const void* a, const void* b, const float* mask, void* o0, void* o1, int N, int C, int H, int W,
                     int Hm, int Wm, cudaStream_t st) {
  constexpr int VN = 16 / sizeof(T);
  const float sh = ac_scale(Hm, H), sw = ac_scale(Wm, W);
  bool vec = (W % VN == 0) && fmi_aligned(a, 16) && (!b || fmi_aligned(b, 16)) && (!o0 || fmi_aligned(o0, 16)) &&
             (!o1 || fmi_aligned(o1, 16));
  const int c_groups = (C + CG - 1) / CG;
  if (vec) {
    int64_t total = (int64_t)N * H * (W / VN) * c_groups;
    int grid = (int)imin64((total + 255) / 256, (int64_t)FMI_NUM_SMS * 16);
    composite_kernel<T, VN, MODE><<<grid, 256, 0, st>>>((const T*)a, (const T*)b, mask, (T*)o0, (T*)o1, N, C, H, W, Hm,
                                                         Wm, sh, sw);
  } else {
    int64_t total = (int64_t)N * H * W * c_groups;
    int grid = (int)imin64((total + 255) / 256, (int64_t)FMI_NUM_SMS * 16);
    composite_kernel<T, 1, MODE><<<grid, 256, 0, st>>>((const T*)a, (const T*)b, mask, (T*)o0, (T*)o1, N, C, H, W, Hm,
                                                        Wm, sh, sw);
  }
  return fmi_launched("composite");
}

}  // namespace

extern "C" int fmi_scale_mask(const float* mask, float* out, int N, int Hm, int Wm, int H, int W, void* stream) {
  FMI_REQUIRE(N >= 0 && Hm >= 1 && Wm >= 1 && H >= 1 && W >= 1, "scale_mask: bad shape");
  if (N == 0) return FMI_OK;
  FMI_REQUIRE(mask && out, "scale_mask: null pointer");
  int64_t total = (int64_t)N * H * W;
  int grid = (int)imin64((total + 255) / 256, (int64_t)FMI_NUM_SMS * 16);
  scale_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, out, N, Hm, Wm, H, W, ac_scale(Hm, H), ac_scale(Wm, W));
  return fmi_launched("scale_mask");
}

extern "C" int fmi_composite(const void* src, const void* ref, const float* mask, void* out, int N, int C, int H,
                             int W, int Hm, int Wm, int dtype, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype), "composite: unsupported dtype %d", dtype);
  FMI_REQUIRE(N >= 0 && C >= 0 && H >= 1 && W >= 1 && Hm >= 1 && Wm >= 1, "composite: bad shape");
  if (N == 0 || C == 0) return FMI_OK;
  FMI_REQUIRE(src && ref && mask && out, "composite: null pointer");
  FMI_DISPATCH_DTYPE(dtype, T,
                     return (launch_composite<T, 0>(src, ref, mask, out, nullptr, N, C, H, W, Hm, Wm, (cudaStream_t)stream)));
  return FMI_OK;
}

extern "C" int fmi_composite_bwd(const void* grad_out, const float* mask, void* grad_src, void* grad_ref, int N, int C,
                                 int H, int W, int Hm, int Wm, int dtype, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype), "composite_bwd: unsupported dtype %d", dtype);
  FMI_REQUIRE(N >= 0 && C >= 0 && H >= 1 && W >= 1 && Hm >= 1 && Wm >= 1, "composite_bwd: bad shape");
  if (N == 0 || C == 0) return FMI_OK;
  FMI_REQUIRE(grad_out && mask, "composite_bwd: null pointer");
  FMI_DISPATCH_DTYPE(dtype, T, return (launch_composite<T, 1>(grad_out, nullptr, mask, grad_src, grad_ref, N, C, H, W,
                                                               Hm, Wm, (cudaStream_t)stream)));
  return FMI_OK;
}

// ---- k x k average pooling of fp32 NCHW planes (psp.py:33,113-114 `face_pool`: AdaptiveAvgPool2d((256, 256)) of the 1024^2
// synthesis = an exact 4 x 4 mean; model.py:79,111 likewise when the decoder's pooling is not fused into its Output kernel).
// HBM-bound: every input element read once as part of a 16-byte vector, one output per thread (k = 4: one float4 per row).
namespace {
template <int K>
__global__ void __launch_bounds__(256) avgpool_planes_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W,
                                                             int64_t total) {
  const int OW = W / K, OH = H / K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(e % OW);
    const int64_t t = e / OW;
    const int oy = (int)(t % OH);
    const int64_t plane = t / OH;
    const float* src = x + (plane * H + (int64_t)oy * K) * W + (int64_t)ox * K;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      if constexpr (K == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)r * W));
        acc += (v.x + v.y) + (v.z + v.w);
      } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src + (int64_t)r * W));
        acc += v.x + v.y;
      }
    }
    y[e] = acc * (1.f / (K * K));
  }
}
}  // namespace

extern "C" int fmi_avgpool_planes(const float* x, float* y, int64_t planes, int H, int W, int k, void* stream) {
  FMI_REQUIRE(planes >= 0 && H >= 1 && W >= 1 && (k == 2 || k == 4), "avgpool_planes: k must be 2 or 4");
  if (planes == 0) return FMI_OK;
  FMI_REQUIRE(x && y && H % k == 0 && W % k == 0 && W % 4 == 0 && fmi_aligned(x, 16),
              "avgpool_planes: H, W must be multiples of k (W of 4) and x 16-byte aligned");
  const int64_t total = planes * (H / k) * (W / k);
  const int grid = (int)imin64((total + 255) / 256, (int64_t)FMI_NUM_SMS * 32);
  cudaStream_t st = (cudaStream_t)stream;
  FmiProfScope prof(FMI_PROF_STREAM, st, (double)planes * H * W, (double)planes * H * W * 4.0 * (1.0 + 1.0 / (k * k)));
  if (k == 4) avgpool_planes_kernel<4><<<grid, 256, 0, st>>>(x, y, H, W, total);
  else avgpool_planes_kernel<2><<<grid, 256, 0, st>>>(x, y, H, W, total);
  return fmi_launched("avgpool_planes");
}
