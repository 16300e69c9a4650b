// f1: the PICNet decoder conv blocks — ResBlockDecoder / Output of
// modules/pluralistic_model/base_function.py:308-398 as built by ResGenerator (network.py:175-268) — on the tcgen05
// implicit-GEMM kernel of modconv_gemm.cuh (weights shared by the batch instead of per-sample) plus three streaming kernels.
//
// One ResBlockDecoder, input x [B,H,W,Cin] -> output [B,2H,2W,Co], NHWC in the tensor-core operand type:
//   a1 = lrelu(IN(x))                    instnorm_stats + instnorm_finalize + norm_act      (base_function.py:343-345)
//   h  = conv3x3(a1) + b1                implicit GEMM, 9 taps, TMA zero fill = padding 1    (:336,345)
//   a2 = lrelu(IN(h))                    written into channels [0,Ch) of the buffer that holds x in channels [Ch,Ch+Cin)
//   y  = convT(a2) + b2 + convT_s(x) + bs ONE implicit GEMM over the concatenated channels [a2 | x] with the concatenated
//                                        weights [W2 | Ws]: both are ConvTranspose2d(3, stride 2, padding 1, output_padding 1)
//                                        (:337-338,350), so main path and shortcut share taps and the sum is the accumulator.
//                                        Stride 2 is run as its 4 output-parity classes: out[2m+py, 2n+px] reads
//                                        ky = 1 (py = 0) or ky in {0 (row m+1), 2 (row m)} (py = 1) — 1/2/2/4 taps, 9/4 taps
//                                        per output pixel, the FLOPs of the reference; rows m+1 = H are TMA zero fill.
//   The epilogue writes y straight into the channel slice of the NEXT block's [a2 | x] buffer, or — last block — applies the
//   Output block's leaky-ReLU (:389-392) and writes the interior of a reflection-padded buffer; `reflect_border` completes
//   ReflectionPad2d(1) and the Output conv is a 9-tap valid conv with tanh, storing the fp32 NCHW image.
// InstanceNorm2d(affine=True, eps=1e-5; base_function.py:46): biased variance over H*W per (sample, channel).
#include "modconv_gemm.cuh"

using namespace sm100;
using namespace fmi_conv;

namespace {

// ---- Wp[t][o][i_off + i] = OT(w[o,i,t])  (Conv2d weight [O,I,3,3])  or  OT(w[i,o,t])  (ConvTranspose2d weight [I,O,3,3]) ----
// merged (transposed only): slab = input shift 2*dy + dx, row = (2*py + px) * O + o, where tap ky serves output parity py =
// (ky != 1) from input row m + dy, dy = (ky == 0)  (out[2m+py] = sum x[iy] w[ky] with 2*iy - 1 + ky = 2m + py); wp is
// [4][4*O][I_row], rows of classes that do not use a shift stay zero.
template <typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) conv_weight_prep_kernel(const float* __restrict__ w, OT* __restrict__ wp, int O, int I,
                                                               int transposed, int O_rows, int I_row, int i_off, int merged,
                                                               int T) {
  const int total = O * I * T;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int t = e % T, r = e / T;
    int o, i;
    if (transposed) { i = r / O; o = r - i * O; }
    else { o = r / I; i = r - o * I; }
    float v = w[e];
    if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
    int slab = t, row = o;
    if (merged) {
      const int ky = t / 3, kx = t - ky * 3;
      slab = (ky == 0 ? 2 : 0) + (kx == 0 ? 1 : 0);
      row = ((ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0)) * O + o;
    }
    wp[((int64_t)slab * O_rows + row) * I_row + i_off + i] = from_f32<OT>(v);
  }
}

// ---- per (sample, channel) sum and sum of squares over the pixels of an NHWC tensor (or channel slice) ----------------
// Thread = one 16-byte channel vector x a strided set of pixels; fp32 partials per thread (<= ~64 pixels), combined in
// double (shared memory, then one atomicAdd(double) pair per channel and block).
template <typename OT>
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const OT* __restrict__ x, int64_t pix_stride, int64_t img_stride,
                                                             int HW, int C, double* __restrict__ sums /*[B][C][2]*/) {
  constexpr int VEC = 16 / sizeof(OT);
  const int nvec = C / VEC;
  const int lanes = 256 / nvec;              // pixel lanes per block
  const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  const int b = blockIdx.y;
  float s[VEC], q[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) s[k] = q[k] = 0.f;
  if (pl < lanes) {
    const OT* xb = x + (int64_t)b * img_stride + v * VEC;
    // four independent 16-byte loads in flight per thread: one per iteration left the kernel at ~0.4 of the HBM rate (2048
    // threads x 16 bytes = 32 KB in flight per SM)
    const int step = gridDim.x * lanes;
    int pp = blockIdx.x * lanes + pl;
    for (; pp + 3 * step < HW; pp += 4 * step) {
      Vec16<OT> t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = ld_vec16_stream(xb + (int64_t)(pp + u * step) * pix_stride);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float f = to_f32<OT>(t[u].e[k]);
          s[k] += f;
          q[k] = fmaf(f, f, q[k]);
        }
    }
    for (; pp < HW; pp += step) {
      const Vec16<OT> t = ld_vec16_stream(xb + (int64_t)pp * pix_stride);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float f = to_f32<OT>(t.e[k]);
        s[k] += f;
        q[k] = fmaf(f, f, q[k]);
      }
    }
  }
  __shared__ float sh[256][2 * VEC + 1];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { sh[threadIdx.x][k] = s[k]; sh[threadIdx.x][VEC + k] = q[k]; }
  __syncthreads();
  // thread e < C*2 reduces one (channel, sum|sq) over the pixel lanes
  for (int e = threadIdx.x; e < 2 * C; e += 256) {
    const int c = e >> 1, which = e & 1;
    const int vv = c / VEC, k = c % VEC;
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += (double)sh[l * nvec + vv][which * VEC + k];
    atomicAdd(&sums[((int64_t)b * C + c) * 2 + which], acc);
  }
}

// scale[b,c] = gamma[c] * rstd, shift[b,c] = beta[c] - mean * scale   (so that IN(x) = x * scale + shift)
__global__ void __launch_bounds__(256) instnorm_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float* __restrict__ ss, int B,
                                                                int C, double inv_n, float eps) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * C) return;
  const int c = e % C;
  const double mean = sums[2 * (int64_t)e] * inv_n;
  double var = sums[2 * (int64_t)e + 1] * inv_n - mean * mean;   // biased variance (F.instance_norm)
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = (gamma ? gamma[c] : 1.f) * rstd;
  ss[2 * (int64_t)e] = sc;
  ss[2 * (int64_t)e + 1] = (beta ? beta[c] : 0.f) - (float)mean * sc;
}

// ---- y = lrelu(x * scale + shift)  (scale/shift NULL: y = lrelu(x)); x and y are NHWC tensors or channel slices ----------
template <typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) norm_act_kernel(const OT* __restrict__ x, int64_t x_pix, int64_t x_img,
                                                       OT* __restrict__ y, int64_t y_pix, int64_t y_img,
                                                       const float* __restrict__ ss /*[B][C][2]*/, int HW, int C, float slope) {
  constexpr int VEC = 16 / sizeof(OT);
  const int nvec = C / VEC;
  const int b = blockIdx.y;
  const int64_t total = (int64_t)HW * nvec;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](const Vec16<OT>& t, int64_t pp, int c) {
    Vec16<OT> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float f = to_f32<OT>(t.e[k]);
      if (ss) {
        const float2 a = *reinterpret_cast<const float2*>(ss + 2 * ((int64_t)b * C + c + k));
        f = fmaf(f, a.x, a.y);
      }
      f = f > 0.f ? f : f * slope;
      if (ROUND_TF32) f = __uint_as_float(f32_to_tf32_rna(f));
      o.e[k] = from_f32<OT>(f);
    }
    st_vec16(y + (int64_t)b * y_img + pp * y_pix + c, o);
  };
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; e + 3 * step < total; e += 4 * step) {   // four independent loads in flight per thread (see instnorm_stats_kernel)
    Vec16<OT> t[4];
    int64_t pp[4];
    int c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t eu = e + u * step;
      pp[u] = eu / nvec;
      c[u] = (int)(eu - pp[u] * nvec) * VEC;
      t[u] = ld_vec16_stream(x + (int64_t)b * x_img + pp[u] * x_pix + c[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(t[u], pp[u], c[u]);
  }
  for (; e < total; e += step) {
    const int64_t pp = e / nvec;
    const int c = (int)(e - pp * nvec) * VEC;
    one(ld_vec16_stream(x + (int64_t)b * x_img + pp * x_pix + c), pp, c);
  }
}

// ---- backward of y = lrelu(IN(x)) (training: ResBlockDecoder's norm + activation pairs, base_function.py:338-344) -------------
// z = x * scale + shift (scale = gamma * rstd, shift = beta - mean * scale), y = lrelu(z):
//   dz = dy * (z > 0 ? 1 : slope);  S1[b,c] = sum_p dz;  S2[b,c] = sum_p dz * xhat,  xhat = (x - mean) * rstd
//   dx = scale * (dz - S1 / HW - xhat * S2 / HW);  dgamma[c] = sum_b S2;  dbeta[c] = sum_b S1
// pass 1 (same thread layout as instnorm_stats_kernel): the two sums per (sample, channel) in double
__global__ void __launch_bounds__(256) instnorm_act_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                     const float* __restrict__ ss, const float* __restrict__ mr,
                                                                     int HW, int C, float slope, double* __restrict__ sums) {
  constexpr int VEC = 4;
  const int nvec = C / VEC;
  const int lanes = 256 / nvec;
  const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  const int b = blockIdx.y;
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) s1[k] = s2[k] = 0.f;
  if (pl < lanes) {
    const int64_t base = (int64_t)b * HW * C + v * VEC;
    float sc[VEC], sh[VEC], mu[VEC], rs[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int64_t e = (int64_t)b * C + v * VEC + k;
      sc[k] = ss[2 * e]; sh[k] = ss[2 * e + 1]; mu[k] = mr[2 * e]; rs[k] = mr[2 * e + 1];
    }
    const int step = gridDim.x * lanes;
    int pp = blockIdx.x * lanes + pl;
    auto acc = [&](const Vec16<float>& g, const Vec16<float>& t) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float z = fmaf(t.e[k], sc[k], sh[k]);
        const float dz = z > 0.f ? g.e[k] : g.e[k] * slope;
        s1[k] += dz;
        s2[k] = fmaf(dz, (t.e[k] - mu[k]) * rs[k], s2[k]);
      }
    };
    for (; pp + step < HW; pp += 2 * step) {   // four independent 16-byte loads in flight per thread
      const Vec16<float> g0 = ld_vec16_stream(dy + base + (int64_t)pp * C), t0 = ld_vec16_stream(x + base + (int64_t)pp * C);
      const Vec16<float> g1 = ld_vec16_stream(dy + base + (int64_t)(pp + step) * C),
                         t1 = ld_vec16_stream(x + base + (int64_t)(pp + step) * C);
      acc(g0, t0);
      acc(g1, t1);
    }
    for (; pp < HW; pp += step) acc(ld_vec16_stream(dy + base + (int64_t)pp * C), ld_vec16_stream(x + base + (int64_t)pp * C));
  }
  __shared__ float shm[256][2 * VEC + 1];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { shm[threadIdx.x][k] = s1[k]; shm[threadIdx.x][VEC + k] = s2[k]; }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * C; e += 256) {
    const int c = e >> 1, which = e & 1;
    const int vv = c / VEC, k = c % VEC;
    double a = 0.0;
    for (int l = 0; l < lanes; ++l) a += (double)shm[l * nvec + vv][which * VEC + k];
    atomicAdd(&sums[((int64_t)b * C + c) * 2 + which], a);
  }
}

// pass 2: dx = scale * (dz - S1 / HW - xhat * S2 / HW)
__global__ void __launch_bounds__(256) instnorm_act_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                     const float* __restrict__ ss, const float* __restrict__ mr,
                                                                     const double* __restrict__ sums, float* __restrict__ dx,
                                                                     int HW, int C, float slope, float inv_hw) {
  constexpr int VEC = 4;
  const int nvec = C / VEC;
  const int b = blockIdx.y;
  const int64_t total = (int64_t)HW * nvec;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t img = (int64_t)b * HW * C;
  auto one = [&](const Vec16<float>& g, const Vec16<float>& t, int64_t off, int c) {
    Vec16<float> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int64_t e = (int64_t)b * C + c + k;
      const float2 a = *reinterpret_cast<const float2*>(ss + 2 * e);
      const float2 m = *reinterpret_cast<const float2*>(mr + 2 * e);
      const float m1 = (float)sums[2 * e] * inv_hw, m2 = (float)sums[2 * e + 1] * inv_hw;
      const float z = fmaf(t.e[k], a.x, a.y);
      const float dz = z > 0.f ? g.e[k] : g.e[k] * slope;
      o.e[k] = a.x * (dz - m1 - (t.e[k] - m.x) * m.y * m2);
    }
    st_vec16(dx + off, o);
  };
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; e + step < total; e += 2 * step) {
    const int64_t p0 = e / nvec, p1 = (e + step) / nvec;
    const int c0 = (int)(e - p0 * nvec) * VEC, c1 = (int)(e + step - p1 * nvec) * VEC;
    const int64_t o0 = img + p0 * C + c0, o1 = img + p1 * C + c1;
    const Vec16<float> g0 = ld_vec16_stream(dy + o0), t0 = ld_vec16_stream(x + o0), g1 = ld_vec16_stream(dy + o1),
                       t1 = ld_vec16_stream(x + o1);
    one(g0, t0, o0, c0);
    one(g1, t1, o1, c1);
  }
  for (; e < total; e += step) {
    const int64_t pp = e / nvec;
    const int c = (int)(e - pp * nvec) * VEC;
    const int64_t o = img + pp * C + c;
    one(ld_vec16_stream(dy + o), ld_vec16_stream(x + o), o, c);
  }
}

// ---- ReflectionPad2d(1) border of a [B, H+2, W+2, C] buffer whose interior is already written ------------------------
template <typename OT>
__global__ void __launch_bounds__(256) reflect_border_kernel(OT* __restrict__ y, int H, int W, int C) {
  constexpr int VEC = 16 / sizeof(OT);
  const int nvec = C / VEC;
  const int PH = H + 2, PW = W + 2;
  const int border = 2 * PW + 2 * H;  // top row, bottom row, then left / right columns of the interior rows
  const int b = blockIdx.y;
  const int64_t total = (int64_t)border * nvec;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e / nvec), c = (int)(e - (int64_t)j * nvec) * VEC;
    int py, px;
    if (j < PW) { py = 0; px = j; }
    else if (j < 2 * PW) { py = PH - 1; px = j - PW; }
    else { const int r = j - 2 * PW; py = 1 + (r >> 1); px = (r & 1) ? PW - 1 : 0; }
    // padded (py, px) <- image (reflect(py - 1), reflect(px - 1)), image pixel (iy, ix) lives at padded (iy + 1, ix + 1)
    int iy = py - 1, ix = px - 1;
    iy = iy < 0 ? -iy : (iy >= H ? 2 * H - 2 - iy : iy);
    ix = ix < 0 ? -ix : (ix >= W ? 2 * W - 2 - ix : ix);
    OT* base = y + (int64_t)b * PH * PW * C;
    st_vec16(base + ((int64_t)py * PW + px) * C + c, ld_vec16(base + ((int64_t)(iy + 1) * PW + ix + 1) * C + c));
  }
}

template <typename TI, typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) nchw_to_nhwc_slice_kernel(const TI* __restrict__ x, OT* __restrict__ y, int C, int HW,
                                                                 int64_t y_pix, int64_t y_img) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    t[j][tx] = (c < C && pp < HW) ? to_f32<TI>(x[((int64_t)b * C + c) * HW + pp]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    if (pp < HW && c < C) {
      float v = t[tx][j];
      if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
      y[(int64_t)b * y_img + (int64_t)pp * y_pix + c] = from_f32<OT>(v);
    }
  }
}

inline int stream_grid(int64_t work_items, int per_block) {
  int64_t g = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)FMI_NUM_SMS * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" int fmi_conv_weight_prep(const float* weight, void* wp, int O, int I, int transposed, int O_rows, int I_row,
                                    int i_off, int merged, int ksize, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "conv_weight_prep: bad mma");
  FMI_REQUIRE(ksize == 3 || (ksize == 1 && !merged), "conv_weight_prep: ksize must be 3, or 1 (not merged)");
  const int T = ksize * ksize;
  FMI_REQUIRE(!merged || (transposed && O_rows == 4 * O), "conv_weight_prep: merged layout is for transposed convs, O_rows = 4*O");
  FMI_REQUIRE(weight && wp && O >= 1 && I >= 1 && O_rows >= O && i_off >= 0 && I_row >= i_off + I,
              "conv_weight_prep: bad arguments (O=%d I=%d O_rows=%d I_row=%d i_off=%d)", O, I, O_rows, I_row, i_off);
  const int grid = stream_grid((int64_t)O * I * T, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (mma == FMI_MMA_TF32 && fmi_tf32_exact_on())
    conv_weight_prep_kernel<float, false><<<grid, 256, 0, st>>>(weight, (float*)wp, O, I, transposed, O_rows, I_row, i_off,
                                                                merged, T);
  else if (mma == FMI_MMA_TF32)
    conv_weight_prep_kernel<float, true><<<grid, 256, 0, st>>>(weight, (float*)wp, O, I, transposed, O_rows, I_row, i_off, merged,
                                                               T);
  else
    conv_weight_prep_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(weight, (__nv_bfloat16*)wp, O, I, transposed, O_rows,
                                                                        I_row, i_off, merged, T);
  return fmi_launched("conv_weight_prep");
}

extern "C" int fmi_nchw_to_nhwc_slice(const void* x, void* y, int B, int C, int H, int W, int64_t y_pixel_stride, int dtype,
                                      int round_y, int mma, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype) && (mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16), "nchw_to_nhwc_slice: bad dtype/mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && y && C >= 1 && H >= 1 && W >= 1 && y_pixel_stride >= C, "nchw_to_nhwc_slice: bad arguments");
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t yi = (int64_t)HW * y_pixel_stride;
  FMI_DISPATCH_DTYPE(dtype, T, {
    if (mma == FMI_MMA_TF32 && round_y)
      nchw_to_nhwc_slice_kernel<T, float, true><<<grid, 256, 0, st>>>((const T*)x, (float*)y, C, HW, y_pixel_stride, yi);
    else if (mma == FMI_MMA_TF32)
      nchw_to_nhwc_slice_kernel<T, float, false><<<grid, 256, 0, st>>>((const T*)x, (float*)y, C, HW, y_pixel_stride, yi);
    else
      nchw_to_nhwc_slice_kernel<T, __nv_bfloat16, false><<<grid, 256, 0, st>>>((const T*)x, (__nv_bfloat16*)y, C, HW,
                                                                               y_pixel_stride, yi);
  });
  return fmi_launched("nchw_to_nhwc_slice");
}

// IN statistics of x [B, H*W, C] (pixel stride x_pixel_stride elements) folded with the affine parameters:
// scale_shift [B][C][2] fp32. `sums` is B*C*2 doubles of scratch.
extern "C" int fmi_instnorm_stats_nhwc(const void* x, int64_t x_pixel_stride, const float* gamma, const float* beta,
                                       float* scale_shift, double* sums, int B, int C, int HW, float eps, int mma,
                                       void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "instnorm_stats: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(x && scale_shift && sums && HW >= 1 && C >= vec && C % vec == 0 && C / vec <= 256 && x_pixel_stride >= C &&
                  x_pixel_stride % vec == 0 && fmi_aligned(x, 16),
              "instnorm_stats: unsupported shape C=%d stride=%lld", C, (long long)x_pixel_stride);
  cudaStream_t st = (cudaStream_t)stream;
  FMI_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * C * 2 * sizeof(double), st));
  const int lanes = 256 / (C / vec);
  int gx = (HW + lanes * 32 - 1) / (lanes * 32);   // ~32 pixels per thread
  const int cap = (FMI_NUM_SMS * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const int64_t img = (int64_t)HW * x_pixel_stride;
  FmiProfScope prof(FMI_PROF_INSTATS, st, 3.0 * B * HW * C, (double)B * HW * C * esz_of(mma));
  if (mma == FMI_MMA_TF32)
    instnorm_stats_kernel<float><<<dim3(gx, B), 256, 0, st>>>((const float*)x, x_pixel_stride, img, HW, C, sums);
  else
    instnorm_stats_kernel<__nv_bfloat16><<<dim3(gx, B), 256, 0, st>>>((const __nv_bfloat16*)x, x_pixel_stride, img, HW, C, sums);
  int rc = fmi_launched("instnorm_stats");
  if (rc) return rc;
  instnorm_finalize_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(sums, gamma, beta, scale_shift, B, C, 1.0 / (double)HW, eps);
  return fmi_launched("instnorm_finalize");
}

// y = lrelu_slope(x * scale + shift) per (sample, channel); scale_shift NULL = activation only.
extern "C" int fmi_norm_act_nhwc(const void* x, int64_t x_pixel_stride, void* y, int64_t y_pixel_stride,
                                 const float* scale_shift, int B, int C, int HW, float slope, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "norm_act: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(x && y && HW >= 1 && C >= vec && C % vec == 0 && x_pixel_stride % vec == 0 && y_pixel_stride % vec == 0 &&
                  x_pixel_stride >= C && y_pixel_stride >= C && fmi_aligned(x, 16) && fmi_aligned(y, 16),
              "norm_act: unsupported shape C=%d strides %lld / %lld", C, (long long)x_pixel_stride, (long long)y_pixel_stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t work = (int64_t)HW * (C / vec);
  int gx = stream_grid(work, 256 * 4);
  const int cap = (FMI_NUM_SMS * 16 + B - 1) / B;
  if (gx > cap) gx = cap;
  FmiProfScope prof(FMI_PROF_NORMACT, st, 3.0 * B * HW * C, 2.0 * B * HW * C * esz_of(mma));
  if (mma == FMI_MMA_TF32 && fmi_tf32_exact_on())
    norm_act_kernel<float, false><<<dim3(gx, B), 256, 0, st>>>((const float*)x, x_pixel_stride, (int64_t)HW * x_pixel_stride,
                                                               (float*)y, y_pixel_stride, (int64_t)HW * y_pixel_stride,
                                                               scale_shift, HW, C, slope);
  else if (mma == FMI_MMA_TF32)
    norm_act_kernel<float, true><<<dim3(gx, B), 256, 0, st>>>((const float*)x, x_pixel_stride, (int64_t)HW * x_pixel_stride,
                                                              (float*)y, y_pixel_stride, (int64_t)HW * y_pixel_stride,
                                                              scale_shift, HW, C, slope);
  else
    norm_act_kernel<__nv_bfloat16, false><<<dim3(gx, B), 256, 0, st>>>(
        (const __nv_bfloat16*)x, x_pixel_stride, (int64_t)HW * x_pixel_stride, (__nv_bfloat16*)y, y_pixel_stride,
        (int64_t)HW * y_pixel_stride, scale_shift, HW, C, slope);
  return fmi_launched("norm_act");
}

// Backward of fmi_instnorm_stats_nhwc + fmi_norm_act_nhwc (y = lrelu(IN(x)), fp32 dense NHWC, training): dx [B,HW,C] and
// sums [B][C][2] doubles = (sum dz, sum dz * xhat) per sample and channel, from which the caller takes dgamma / dbeta.
// scale_shift as produced by fmi_instnorm_stats_nhwc; mean_rstd [B][C][2] fp32 (mean, 1 / sqrt(var + eps)).
extern "C" int fmi_instnorm_act_bwd_nhwc(const float* dy, const float* x, const float* scale_shift, const float* mean_rstd,
                                         float* dx, double* sums, int B, int C, int HW, float slope, void* stream) {
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(dy && x && scale_shift && mean_rstd && dx && sums && HW >= 1 && C >= 4 && C % 4 == 0 && C / 4 <= 256 &&
                  fmi_aligned(dy, 16) && fmi_aligned(x, 16) && fmi_aligned(dx, 16),
              "instnorm_act_bwd: unsupported shape C=%d", C);
  cudaStream_t st = (cudaStream_t)stream;
  FMI_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * C * 2 * sizeof(double), st));
  const int lanes = 256 / (C / 4);
  int gx = (HW + lanes * 32 - 1) / (lanes * 32);
  int cap = (FMI_NUM_SMS * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  FmiProfScope prof(FMI_PROF_NORMACT, st, 12.0 * B * HW * C, 5.0 * B * HW * C * 4);
  instnorm_act_bwd_stats_kernel<<<dim3(gx, B), 256, 0, st>>>(dy, x, scale_shift, mean_rstd, HW, C, slope, sums);
  int rc = fmi_launched("instnorm_act_bwd_stats");
  if (rc) return rc;
  int ga = stream_grid((int64_t)HW * (C / 4), 256 * 4);
  cap = (FMI_NUM_SMS * 16 + B - 1) / B;
  if (ga > cap) ga = cap;
  instnorm_act_bwd_apply_kernel<<<dim3(ga, B), 256, 0, st>>>(dy, x, scale_shift, mean_rstd, sums, dx, HW, C, slope,
                                                             1.0f / (float)HW);
  return fmi_launched("instnorm_act_bwd_apply");
}

extern "C" int fmi_reflect_border_nhwc(void* y, int B, int C, int H, int W, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "reflect_border: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(y && H >= 2 && W >= 2 && C % vec == 0 && fmi_aligned(y, 16), "reflect_border: unsupported shape");
  const int64_t work = (int64_t)(2 * (W + 2) + 2 * H) * (C / vec);
  const int gx = stream_grid(work, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (mma == FMI_MMA_TF32) reflect_border_kernel<float><<<dim3(gx, B), 256, 0, st>>>((float*)y, H, W, C);
  else reflect_border_kernel<__nv_bfloat16><<<dim3(gx, B), 256, 0, st>>>((__nv_bfloat16*)y, H, W, C);
  return fmi_launched("reflect_border");
}

// 3x3 convolution / stride-2 transposed convolution with batch-shared weights on the implicit-GEMM kernel.
//   mode 0: Conv2d(3, stride 1, padding 1)                              x [B,H,W,*]      -> y [B,H,W,*]
//   mode 1: Conv2d(3, stride 1, padding 0) on a pre-padded input        x [B,H+2,W+2,*]  -> y [B,H,W,*]
//   mode 2: ConvTranspose2d(3, stride 2, padding 1, output_padding 1)   x [B,H,W,*]      -> y [B,2H,2W,*]
//   mode 3: the same transposed conv as ONE GEMM (O <= 64): N = 4*O columns = the 4 output-parity classes, 4 taps = the 4
//           input shifts, wp [4][4*O][I] from fmi_conv_weight_prep(merged = 1). 4/9 of the MMA instructions of mode 2 —
//           narrow layers sit on the per-instruction floor of the tensor pipe (~85 clk for any N <= 64), not on its FLOPs.
//   mode 4: Conv2d(1, stride 1, padding 0): wp [1][O][I] (fmi_conv_weight_prep with ksize 1)
//   act + 10 (with act 2): y += acc + bias — the residual sum of a ResBlock (conv2(a2) onto the stored shortcut).
//   x: I channels per pixel out of x_pixel_stride; wp [9][O][I] from fmi_conv_weight_prep (O a multiple of 32, rows >= the
//   real output channels zero); bias [O] fp32 or NULL; act 2: y = acc + bias, 1: lrelu_slope(acc + bias), 3: tanh(acc + bias).
//   round_y = 0 (TF32 mode): y keeps the exact fp32 accumulator instead of its tf32 rounding — for outputs that feed
//   InstanceNorm (a rounding of x is amplified by |mean| / std there); an MMA reading such a tensor truncates it instead.
//   y: NHWC in the operand type, O channels per pixel out of y_pixel_stride, written at the interior of a buffer padded by
//   y_pad pixels on each side (0 or 1); may be NULL when y_nchw is given. y_nchw: fp32 [B, nchw_C, OH, OW] or NULL.
extern "C" int fmi_conv3x3_nhwc(const void* x, int64_t x_pixel_stride, const void* wp, const float* bias, void* y,
                                int64_t y_pixel_stride, int y_pad, float* y_nchw, int nchw_C, int B, int I, int O, int H,
                                int W, int mode, int act, float slope, int round_y, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "conv3x3: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && wp && (y || y_nchw), "conv3x3: null pointer");
  const int add_y = act >= 10;          // act + 10: y += result (residual sum; y must hold the other term already)
  if (add_y) act -= 10;
  FMI_REQUIRE(mode >= 0 && mode <= 4 && act >= 1 && act <= 3 && (y_pad == 0 || y_pad == 1), "conv3x3: bad mode/act/pad");
  FMI_REQUIRE(!add_y || (y && act == 2 && mode != 3), "conv3x3: the residual sum needs an NHWC output and act 2");
  FMI_REQUIRE(mode != 3 || (O <= 64 && y && !y_nchw), "conv3x3: mode 3 (merged parity classes) needs O <= 64 and an NHWC output");
  const int esz = esz_of(mma);
  FMI_REQUIRE(I >= 16 && O >= 32 && O % 32 == 0 && (O <= 256 || O % 256 == 0) && H >= 1 && W >= 1,
              "conv3x3: unsupported shape I=%d O=%d H=%d W=%d (O must be a multiple of 32, <= 256 or a multiple of 256)", I, O, H, W);
  FMI_REQUIRE((I * esz) % 16 == 0 && (x_pixel_stride * esz) % 16 == 0 && x_pixel_stride >= I && fmi_aligned(x, 16) &&
                  fmi_aligned(wp, 16),
              "conv3x3: input rows must be 16-byte multiples");
  FMI_REQUIRE(!y || ((y_pixel_stride * esz) % 16 == 0 && y_pixel_stride >= O && fmi_aligned(y, 16)),
              "conv3x3: output rows must be 16-byte multiples");
  FMI_REQUIRE(!y_nchw || (nchw_C >= 1 && nchw_C <= O), "conv3x3: bad nchw_C");
  int rc = fmi_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tf32 = mma == FMI_MMA_TF32;
  const uint32_t epa = 128 / esz;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int IH = mode == 1 ? H + 2 : H, IW = mode == 1 ? W + 2 : W;   // input extents
  const bool up = mode == 2 || mode == 3;
  const int OH = up ? 2 * H : H, OW = up ? 2 * W : W;

  ConvGemmParams p{};
  p.B = B; p.I = I; p.O = O; p.H = IH; p.W = IW; p.T = 9;
  p.w_shared = 1;
  p.raw_out = !round_y;
  p.add_out = add_y;
  if (mode == 4) p.T = 1;               // 1x1 convolution: one tap, wp [1][O][I]
  p.n_tile = O <= 256 ? O : 256;
  if (mode != 3) {   // few pixel tiles: narrower output tiles on more SMs (pick_n_tile)
    const TilePlan t0 = pick_tile(H, W);
    p.n_tile = pick_n_tile(O, p.n_tile, (int64_t)B * t0.tiles_h * t0.tiles_w);
  }
  if (mode == 3) {   // one GEMM for the 4 parity classes: N = 4*O, weights [4 shifts][4*O][I]
    p.merge_o = O;
    p.O = 4 * O;
    p.n_tile = 4 * O;
    p.T = 4;
  }
  p.k_chunks = (I + epa - 1) / epa;
  p.bias = bias; p.act = act; p.slope = slope; p.gain = 1.f;
  p.OH = OH; p.OW = OW;
  p.nchw_out = y_nchw; p.nchw_C = nchw_C;
  if (y) {
    const int64_t PW = OW + 2 * y_pad, PH = OH + 2 * y_pad;
    p.out_pstride = (int)y_pixel_stride;
    p.out_rstride = PW * y_pixel_stride;
    p.out_bstride = PH * PW * y_pixel_stride;
    p.out = (uint8_t*)y + ((int64_t)y_pad * PW + y_pad) * y_pixel_stride * esz;
  } else {
    p.out = nullptr;
    p.out_pstride = O; p.out_rstride = (int64_t)OW * O; p.out_bstride = (int64_t)OH * OW * O;
  }

  CUtensorMap mw;
  {
    uint64_t dims[2] = {(uint64_t)I, (uint64_t)p.T * p.O};
    uint64_t str[1] = {(uint64_t)I * esz};
    uint32_t box[2] = {epa, (uint32_t)p.n_tile};
    int e = make_tensor_map(&mw, dt, 2, wp, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "conv3x3: cuTensorMapEncodeTiled(weights) failed (%d)", e);
  }
  auto make_x_map = [&](CUtensorMap* m, int th, int tw) -> int {
    uint64_t dims[4] = {(uint64_t)I, (uint64_t)IW, (uint64_t)IH, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)x_pixel_stride * esz, (uint64_t)IW * x_pixel_stride * esz,
                       (uint64_t)IH * IW * x_pixel_stride * esz};
    uint32_t box[4] = {epa, (uint32_t)tw, (uint32_t)th, 1};
    return make_tensor_map(m, dt, 4, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };

  if (mode == 3) {
    p.Mh = H; p.Mw = W; p.sy = p.sx = 2; p.py = p.px = 0;
    p.ntaps = 4;
    for (int t = 0; t < 4; ++t) { p.tap_dy[t] = t >> 1; p.tap_dx[t] = t & 1; p.tap_slab[t] = t; }
    TilePlan tp = pick_tile(p.Mh, p.Mw);
    CUtensorMap mx;
    int e = make_x_map(&mx, tp.TH, tp.TW);
    FMI_REQUIRE(e == 0, "conv3x3: cuTensorMapEncodeTiled(x) failed (%d)", e);
    return tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
  }
  if (mode == 4) {
    p.Mh = H; p.Mw = W; p.sy = p.sx = 1; p.py = p.px = 0;
    p.ntaps = 1;
    p.tap_dy[0] = p.tap_dx[0] = p.tap_slab[0] = 0;
    TilePlan tp = pick_tile(p.Mh, p.Mw);
    CUtensorMap mx;
    int e = make_x_map(&mx, tp.TH, tp.TW);
    FMI_REQUIRE(e == 0, "conv3x3: cuTensorMapEncodeTiled(x) failed (%d)", e);
    return tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
  }
  if (mode != 2) {
    p.Mh = H; p.Mw = W; p.sy = p.sx = 1; p.py = p.px = 0;
    p.ntaps = 9;
    const int off = mode == 1 ? 0 : -1;   // valid conv on the padded input: taps 0..2; padding 1: taps -1..1 (TMA zero fill)
    for (int t = 0; t < 9; ++t) { p.tap_dy[t] = t / 3 + off; p.tap_dx[t] = t % 3 + off; p.tap_slab[t] = t; }
    TilePlan tp = pick_tile(p.Mh, p.Mw);
    // halo mode (modconv_gemm.cuh): one 130-pixel TMA box per kernel row instead of three 128-pixel boxes. ncu on the
    // 64 -> 32 @512^2 layer without it: 6.0 GB through TMA for 0.54 GB of DRAM reads, L2 throughput 64 % (max 80 %) — these
    // fp32-operand layers are bound by L2 -> SM operand ingest. FMI_CONV_HALO=0 switches it off (A/B measurements).
    static const bool halo_off = [] { const char* e = getenv("FMI_CONV_HALO"); return e && e[0] == '0'; }();
    p.halo = mode == 0 && !halo_off && W >= 128 && p.n_tile <= 128;
    if (p.halo) {
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) p.halo_slab[a][c] = a * 3 + c;
      tp = TilePlan{1, 130, 0, 0};
    }
    CUtensorMap mx;
    int e = make_x_map(&mx, tp.TH, tp.TW);
    FMI_REQUIRE(e == 0, "conv3x3: cuTensorMapEncodeTiled(x) failed (%d)", e);
    return tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
  }
  p.sy = p.sx = 2; p.Mh = H; p.Mw = W;
  TilePlan tp = pick_tile(p.Mh, p.Mw);
  CUtensorMap mx;
  int e = make_x_map(&mx, tp.TH, tp.TW);
  FMI_REQUIRE(e == 0, "conv3x3: cuTensorMapEncodeTiled(x) failed (%d)", e);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      // out[2m+py, 2n+px] = sum over (ky, iy) with 2*iy - 1 + ky = 2m + py:  py = 0: (1, m);  py = 1: (0, m+1), (2, m)
      p.py = py; p.px = px;
      int nt = 0;
      for (int a = 0; a < (py ? 2 : 1); ++a)
        for (int c = 0; c < (px ? 2 : 1); ++c) {
          const int ky = py ? (a ? 2 : 0) : 1, kx = px ? (c ? 2 : 0) : 1;
          p.tap_dy[nt] = ky == 0 ? 1 : 0;
          p.tap_dx[nt] = kx == 0 ? 1 : 0;
          p.tap_slab[nt] = ky * 3 + kx;
          ++nt;
        }
      p.ntaps = nt;
      rc = tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
      if (rc) return rc;
    }
  return FMI_OK;
}

// =====================================================================================================================
// Output block convolution (base_function.py:369-398): Conv2d(C -> <=4 channels, 3x3, padding 0) on the reflection-padded
// activations + tanh. 2*9*C*3 FLOPs per pixel for 128*C bytes read: HBM-bound, and on the tensor pipe every 128-pixel
// tile costs 9*C/8 instructions of ~85 clk for N = 3 useful columns (measured 0.91 ms at 1024^2 x 8, the largest launch of
// the decoder). SIMT instead: weights live in __constant__ memory so every FFMA takes its weight as a constant-bank
// operand (no load instruction), activations are staged per 16-channel chunk in shared memory as [channel vector][row][col]
// float4 planes (a warp reads 32 consecutive float4: conflict-free), each thread owns a 4-row x 1-column strip x 3 outputs
// (12 independent accumulators; 18 LDS.128 per 432 FFMA). fp32 accumulation of operand-type inputs: more accurate than the
// TF32 GEMM it replaces. Optional fused 4x4 average pooling (AdaptiveAvgPool2d(256) of the 1024^2 image, model.py:111).
// =====================================================================================================================
namespace {
constexpr int OC_MAX_C = 64;
__constant__ float c_out_w[3 * 9 * OC_MAX_C];  // [o][tap][c]
__constant__ float c_out_b[4];

__global__ void __launch_bounds__(256) out_weight_to_const_layout_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                                                         float* __restrict__ dst, int O, int C) {
  // w [O][C][3][3] -> dst [3][9][C] (rows o >= O zero), then 4 bias floats at dst + 3*9*C
  for (int e = threadIdx.x; e < 3 * 9 * C; e += blockDim.x) {
    const int c = e % C, t = (e / C) % 9, o = e / (9 * C);
    dst[e] = o < O ? w[((int64_t)o * C + c) * 9 + t] : 0.f;
  }
  if (threadIdx.x < 4) dst[3 * 9 * C + threadIdx.x] = (b && (int)threadIdx.x < O) ? b[threadIdx.x] : 0.f;
}

constexpr int OC_TH = 16, OC_TW = 32;                       // output tile per CTA (128 threads: 4 warps x 4 rows)
constexpr int OC_PLANE = ((OC_TH + 2) * (OC_TW + 2) + 7) / 8 * 8 + 1;   // float4 per channel-vector plane, = 1 mod 8

template <typename OT, int C>
__global__ void __launch_bounds__(128) out_conv_tanh_kernel(const OT* __restrict__ xpad /*[B,H+2,W+2,C]*/,
                                                            float* __restrict__ img /*[B,O,H,W] or NULL*/,
                                                            float* __restrict__ pooled /*[B,O,H/4,W/4] or NULL*/, int O, int H,
                                                            int W) {
  constexpr int CH = 16;                 // channels per staged chunk (4 float4 planes)
  __shared__ float4 tile[4 * OC_PLANE];
  const int b = blockIdx.z, y0 = blockIdx.y * OC_TH, x0 = blockIdx.x * OC_TW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int PW = W + 2;
  const OT* xb = xpad + (int64_t)b * (H + 2) * PW * C;
  float acc[4][3];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int o = 0; o < 3; ++o) acc[r][o] = c_out_b[o];
#pragma unroll
  for (int cc = 0; cc < C; cc += CH) {
    if (cc) __syncthreads();
    // stage (TH+2) x (TW+2) pixels x 16 channels: thread = (pixel, channel vector of 4), channel vector fastest
    for (int e = threadIdx.x; e < (OC_TH + 2) * (OC_TW + 2) * 4; e += 128) {
      const int v = e & 3, pix = e >> 2;
      const int ty = pix / (OC_TW + 2), tx = pix - ty * (OC_TW + 2);
      const int gy = y0 + ty, gx = x0 + tx;   // padded coordinates
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy < H + 2 && gx < PW) {
        const OT* src = xb + ((int64_t)gy * PW + gx) * C + cc + v * 4;
        if constexpr (sizeof(OT) == 4) {
          val = *reinterpret_cast<const float4*>(src);
        } else {
          const uint2 u = *reinterpret_cast<const uint2*>(src);
          const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
          const __nv_bfloat162 c2 = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
          val = make_float4(__low2float(a), __high2float(a), __low2float(c2), __high2float(c2));
        }
      }
      tile[v * OC_PLANE + pix] = val;
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4* plane = tile + v * OC_PLANE + (warp * 4) * (OC_TW + 2) + lane;
#pragma unroll
      for (int ry = 0; ry < 6; ++ry) {        // input rows of the strip (4 outputs + 2)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 d = plane[ry * (OC_TW + 2) + kx];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int ky = ry - r;
            if (ky < 0 || ky > 2) continue;
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              const float* wv = c_out_w + (o * 9 + ky * 3 + kx) * C + cc + v * 4;
              acc[r][o] = fmaf(d.x, wv[0], acc[r][o]);
              acc[r][o] = fmaf(d.y, wv[1], acc[r][o]);
              acc[r][o] = fmaf(d.z, wv[2], acc[r][o]);
              acc[r][o] = fmaf(d.w, wv[3], acc[r][o]);
            }
          }
        }
      }
    }
  }
  const int x = x0 + lane, yb = y0 + warp * 4;
  float ps[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const float t = tanhf(acc[r][o]);
      ps[o] += t;
      if (img && o < O && x < W && yb + r < H) img[(((int64_t)b * O + o) * H + yb + r) * W + x] = t;
    }
  if (pooled) {   // 4 rows in the thread, 4 columns across lanes (H, W multiples of 4: checked by the host)
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      ps[o] += __shfl_xor_sync(0xffffffffu, ps[o], 1);
      ps[o] += __shfl_xor_sync(0xffffffffu, ps[o], 2);
    }
    if ((lane & 3) == 0 && x < W && yb < H)
      for (int o = 0; o < O; ++o)
        pooled[(((int64_t)b * O + o) * (H / 4) + yb / 4) * (W / 4) + x / 4] = ps[o] * (1.f / 16.f);
  }
}
}  // namespace

// Output block: img = tanh(conv3x3_valid(xpad) + bias), xpad [B,H+2,W+2,C] NHWC operand type (leaky-ReLU and reflection
// padding already applied), weight [O,C,3,3] fp32 (O <= 3), bias [O] or NULL. img [B,O,H,W] fp32 and / or pooled
// [B,O,H/4,W/4] fp32 (4x4 means; H, W multiples of 4) — either may be NULL. scratch: (27*C + 4) floats.
extern "C" int fmi_output_conv_tanh(const void* xpad, const float* weight, const float* bias, float* img, float* pooled,
                                    float* scratch, int B, int C, int O, int H, int W, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "output_conv_tanh: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(xpad && weight && scratch && (img || pooled), "output_conv_tanh: null pointer");
  FMI_REQUIRE((C == 32 || C == 64 || C == 16) && O >= 1 && O <= 3 && H >= 1 && W >= 1 && B <= 65535,
              "output_conv_tanh: unsupported shape C=%d O=%d (C in {16,32,64}, O <= 3)", C, O);
  FMI_REQUIRE(!pooled || (H % 4 == 0 && W % 4 == 0), "output_conv_tanh: pooling needs H, W multiples of 4");
  FMI_REQUIRE(fmi_aligned(xpad, 16), "output_conv_tanh: xpad must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  out_weight_to_const_layout_kernel<<<1, 256, 0, st>>>(weight, bias, scratch, O, C);
  int rc = fmi_launched("out_weight_layout");
  if (rc) return rc;
  FMI_CUDA(cudaMemcpyToSymbolAsync(c_out_w, scratch, (size_t)27 * C * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
  FMI_CUDA(cudaMemcpyToSymbolAsync(c_out_b, scratch + 27 * C, 4 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
  dim3 grid((W + OC_TW - 1) / OC_TW, (H + OC_TH - 1) / OC_TH, B);
  FMI_REQUIRE(grid.y <= 65535, "output_conv_tanh: image too tall");
  FmiProfScope prof(FMI_PROF_OUTCONV, st, 2.0 * 9 * C * O * (double)B * H * W,
                    (double)B * (H + 2) * (W + 2) * C * esz_of(mma) + (double)B * O * H * W * 4.0 * (img ? 1.0 : 0.0) +
                        (pooled ? (double)B * O * H * W * 4.0 / 16.0 : 0.0));
#define FMI_OC_LAUNCH(OT, CC) out_conv_tanh_kernel<OT, CC><<<grid, 128, 0, st>>>((const OT*)xpad, img, pooled, O, H, W)
  if (mma == FMI_MMA_TF32) {
    if (C == 16) FMI_OC_LAUNCH(float, 16); else if (C == 32) FMI_OC_LAUNCH(float, 32); else FMI_OC_LAUNCH(float, 64);
  } else {
    if (C == 16) FMI_OC_LAUNCH(__nv_bfloat16, 16); else if (C == 32) FMI_OC_LAUNCH(__nv_bfloat16, 32); else FMI_OC_LAUNCH(__nv_bfloat16, 64);
  }
#undef FMI_OC_LAUNCH
  return fmi_launched("out_conv_tanh");
}

// =====================================================================================================================
// SpectralNorm power iteration (external_function.py:44-57) in two kernels instead of ~13 ATen launches per convolution:
//   v_raw = W^T u                                  (sn_wt_u_kernel)
//   v = v_raw / (|v_raw| + eps);  u_raw = W v      (sn_w_v_kernel; |v_raw| recomputed per CTA, CTA 0 stores v)
//   u = u_raw / (|u_raw| + eps);  sigma = u . (W v) = u . u_raw   -> fmi_conv_weight_prep_sn divides by it and stores u.
// W = w_bar viewed as [Hh][Wd] (Hh = weight.shape[0]).
// =====================================================================================================================
namespace {
__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// v_raw = W^T u. CTA = 32 columns x 8 row slices (a warp reads 32 consecutive floats of a row); one thread per column walking all
// Hh rows was a chain of Hh / 4 dependent L2 round trips (13 us for a 256-row weight, x 130 wrapped convolutions per GAN step)
__global__ void __launch_bounds__(256) sn_wt_u_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                                      float* __restrict__ v_raw, int Hh, int Wd) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (j < Wd) {
    int i = ty;
    for (; i + 24 < Hh; i += 32) {
      a0 = fmaf(w[(int64_t)i * Wd + j], u[i], a0);
      a1 = fmaf(w[(int64_t)(i + 8) * Wd + j], u[i + 8], a1);
      a2 = fmaf(w[(int64_t)(i + 16) * Wd + j], u[i + 16], a2);
      a3 = fmaf(w[(int64_t)(i + 24) * Wd + j], u[i + 24], a3);
    }
    for (; i < Hh; i += 8) a0 = fmaf(w[(int64_t)i * Wd + j], u[i], a0);
  }
  part[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && j < Wd) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += part[r][tx];
    v_raw[j] = t;
  }
}

__global__ void __launch_bounds__(256) sn_w_v_kernel(const float* __restrict__ w, const float* __restrict__ v_raw,
                                                     float* __restrict__ v, float* __restrict__ u_raw, int Hh, int Wd) {
  __shared__ float red[8];
  float q = 0.f;
  for (int j = threadIdx.x; j < Wd; j += 256) q = fmaf(v_raw[j], v_raw[j], q);
  const float inv = 1.f / (sqrtf(block_sum_256(q, red)) + 1e-12f);
  if (blockIdx.x == 0)
    for (int j = threadIdx.x; j < Wd; j += 256) v[j] = v_raw[j] * inv;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= Hh) return;
  float a = 0.f;
  for (int j = lane; j < Wd; j += 32) a = fmaf(w[(int64_t)row * Wd + j], v_raw[j] * inv, a);
  a = warp_sum(a);
  if (lane == 0) u_raw[row] = a;
}

// conv_weight_prep with the SpectralNorm division: sigma from u_raw (Hh <= 1024 values), CTA 0 stores the new u
template <typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) conv_weight_prep_sn_kernel(const float* __restrict__ w, OT* __restrict__ wp, int O, int I,
                                                                  int transposed, int O_rows, int I_row, int i_off, int merged,
                                                                  const float* __restrict__ u_raw, float* __restrict__ u, int Hh,
                                                                  int T) {
  __shared__ float red[8];
  float q = 0.f;
  for (int i = threadIdx.x; i < Hh; i += 256) q = fmaf(u_raw[i], u_raw[i], q);
  const float n2 = block_sum_256(q, red);
  const float inv = 1.f / (sqrtf(n2) + 1e-12f);
  const float sigma = n2 * inv;                       // u . u_raw with u = u_raw * inv
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < Hh; i += 256) u[i] = u_raw[i] * inv;
  const int total = O * I * T;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int t = e % T, r = e / T;
    int o, i;
    if (transposed) { i = r / O; o = r - i * O; }
    else { o = r / I; i = r - o * I; }
    float v = w[e] / sigma;
    if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
    int slab = t, row = o;
    if (merged) {
      const int ky = t / 3, kx = t - ky * 3;
      slab = (ky == 0 ? 2 : 0) + (kx == 0 ? 1 : 0);
      row = ((ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0)) * O + o;
    }
    wp[((int64_t)slab * O_rows + row) * I_row + i_off + i] = from_f32<OT>(v);
  }
}
}  // namespace

// fmi_conv_weight_prep for a SpectralNorm-wrapped 3x3 conv: one power iteration on (w_bar, u, v) — u and v are updated in
// place exactly as SpectralNorm._update_u_v does — and wp receives w_bar / sigma. scratch: (Wd + Hh) floats, Hh =
// w_bar.shape[0], Wd = numel / Hh.
extern "C" int fmi_conv_weight_prep_sn(const float* w_bar, float* u, float* v, float* scratch, void* wp, int O, int I,
                                       int transposed, int O_rows, int I_row, int i_off, int merged, int ksize, int mma,
                                       void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "conv_weight_prep_sn: bad mma");
  FMI_REQUIRE(ksize == 3 || (ksize == 1 && !merged), "conv_weight_prep_sn: ksize must be 3, or 1 (not merged)");
  const int T = ksize * ksize;
  FMI_REQUIRE(w_bar && u && v && scratch && wp && O >= 1 && I >= 1 && O_rows >= O && i_off >= 0 && I_row >= i_off + I,
              "conv_weight_prep_sn: bad arguments");
  FMI_REQUIRE(!merged || (transposed && O_rows == 4 * O), "conv_weight_prep_sn: merged layout is for transposed convs");
  const int Hh = transposed ? I : O;       // weight.shape[0]
  const int Wd = (transposed ? O : I) * T;
  cudaStream_t st = (cudaStream_t)stream;
  float* v_raw = scratch;
  float* u_raw = scratch + Wd;
  sn_wt_u_kernel<<<(Wd + 31) / 32, 256, 0, st>>>(w_bar, u, v_raw, Hh, Wd);
  int rc = fmi_launched("sn_wt_u");
  if (rc) return rc;
  sn_w_v_kernel<<<(Hh + 7) / 8, 256, 0, st>>>(w_bar, v_raw, v, u_raw, Hh, Wd);
  rc = fmi_launched("sn_w_v");
  if (rc) return rc;
  const int grid = stream_grid((int64_t)O * I * T, 256);
  if (mma == FMI_MMA_TF32 && fmi_tf32_exact_on())
    conv_weight_prep_sn_kernel<float, false><<<grid, 256, 0, st>>>(w_bar, (float*)wp, O, I, transposed, O_rows, I_row, i_off,
                                                                   merged, u_raw, u, Hh, T);
  else if (mma == FMI_MMA_TF32)
    conv_weight_prep_sn_kernel<float, true><<<grid, 256, 0, st>>>(w_bar, (float*)wp, O, I, transposed, O_rows, I_row, i_off,
                                                                  merged, u_raw, u, Hh, T);
  else
    conv_weight_prep_sn_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(w_bar, (__nv_bfloat16*)wp, O, I, transposed, O_rows,
                                                                           I_row, i_off, merged, u_raw, u, Hh, T);
  return fmi_launched("conv_weight_prep_sn");
}

// ---- SpectralNorm in TRAINING (external_function.py:30-42): the same power iteration, then the plain weight w = w_bar / sigma in
// the parameter's own layout for the framework's convolution, and its backward. With u, v held constant (they are updated from
// .data) sigma = u^T W v, so  dL/dW_bar = (g - <g, w> u v^T) / sigma  for an upstream gradient g of w.
namespace {
__global__ void __launch_bounds__(256) sn_divide_kernel(const float* __restrict__ w, float* __restrict__ w_out,
                                                        const float* __restrict__ u_raw, float* __restrict__ u,
                                                        const float* __restrict__ v, float* __restrict__ snap, int Hh, int Wd) {
  __shared__ float red[8];
  float q = 0.f;
  for (int i = threadIdx.x; i < Hh; i += 256) q = fmaf(u_raw[i], u_raw[i], q);
  const float n2 = block_sum_256(q, red);
  const float inv = 1.f / (sqrtf(n2) + 1e-12f);
  const float sigma = n2 * inv;                       // u . u_raw with u = u_raw * inv
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < Hh; i += 256) {
      const float un = u_raw[i] * inv;
      u[i] = un;
      snap[i] = un;
    }
    for (int j = threadIdx.x; j < Wd; j += 256) snap[Hh + j] = v[j];
    if (threadIdx.x == 0) snap[Hh + Wd] = sigma;
  }
  const int64_t total = (int64_t)Hh * Wd;
  const float rs = 1.f / sigma;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    w_out[e] = w[e] * rs;
}

__global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const float* __restrict__ g, const float* __restrict__ w_out,
                                                         float* __restrict__ dot, int64_t total) {
  __shared__ float red[8];
  float a = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    a = fmaf(g[e], w_out[e], a);
  a = block_sum_256(a, red);
  if (threadIdx.x == 0) atomicAdd(dot, a);
}

__global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ snap,
                                                           const float* __restrict__ dot, float* __restrict__ grad, int Hh, int Wd) {
  const int64_t total = (int64_t)Hh * Wd;
  const float rs = 1.f / snap[Hh + Wd], d = *dot;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / Wd), c = (int)(e - (int64_t)r * Wd);
    grad[e] = (g[e] - d * snap[r] * snap[Hh + c]) * rs;
  }
}
}  // namespace

// One power iteration on (w_bar [Hh][Wd], u [Hh], v [Wd]) — u, v updated in place as SpectralNorm._update_u_v does — and
// w_out = w_bar / sigma (same layout). snap [Hh + Wd + 1] receives the new u, the new v and sigma for the backward.
// scratch: (Wd + Hh) floats.
extern "C" int fmi_spectral_norm_fwd(const float* w_bar, float* u, float* v, float* scratch, float* w_out, float* snap, int Hh,
                                     int Wd, void* stream) {
  FMI_REQUIRE(w_bar && u && v && scratch && w_out && snap && Hh >= 1 && Wd >= 1, "spectral_norm_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  float* v_raw = scratch;
  float* u_raw = scratch + Wd;
  sn_wt_u_kernel<<<(Wd + 31) / 32, 256, 0, st>>>(w_bar, u, v_raw, Hh, Wd);
  int rc = fmi_launched("sn_wt_u");
  if (rc) return rc;
  sn_w_v_kernel<<<(Hh + 7) / 8, 256, 0, st>>>(w_bar, v_raw, v, u_raw, Hh, Wd);
  rc = fmi_launched("sn_w_v");
  if (rc) return rc;
  sn_divide_kernel<<<stream_grid((int64_t)Hh * Wd, 256 * 4), 256, 0, st>>>(w_bar, w_out, u_raw, u, v, snap, Hh, Wd);
  return fmi_launched("sn_divide");
}

// grad_w_bar = (g - <g, w_out> u v^T) / sigma with (u, v, sigma) = snap of the forward; dot: one zero-initialised float.
extern "C" int fmi_spectral_norm_bwd(const float* g, const float* w_out, const float* snap, float* dot, float* grad_w_bar, int Hh,
                                     int Wd, void* stream) {
  FMI_REQUIRE(g && w_out && snap && dot && grad_w_bar && Hh >= 1 && Wd >= 1, "spectral_norm_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)Hh * Wd;
  FMI_CUDA(cudaMemsetAsync(dot, 0, sizeof(float), st));
  const int grid = stream_grid(total, 256 * 4);
  sn_bwd_dot_kernel<<<grid, 256, 0, st>>>(g, w_out, dot, total);
  int rc = fmi_launched("sn_bwd_dot");
  if (rc) return rc;
  sn_bwd_apply_kernel<<<grid, 256, 0, st>>>(g, snap, dot, grad_w_bar, Hh, Wd);
  return fmi_launched("sn_bwd_apply");
}

// ---- AvgPool2d(2, 2) on NHWC (the 'down' ResBlocks and the first encoder block, base_function.py:238-239, 290-298) ------
namespace {
template <typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) avgpool2_nhwc_kernel(const OT* __restrict__ x, int64_t x_pix, OT* __restrict__ y,
                                                            int64_t y_pix, int C, int H, int W) {
  constexpr int VEC = 16 / sizeof(OT);
  const int nvec = C / VEC, OH = H / 2, OW = W / 2;
  const int b = blockIdx.y;
  const int64_t total = (int64_t)OH * OW * nvec;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pp = e / nvec;
    const int c = (int)(e - pp * nvec) * VEC;
    const int oy = (int)(pp / OW), ox = (int)(pp - (int64_t)oy * OW);
    const OT* src = x + ((int64_t)b * H * W + (int64_t)(2 * oy) * W + 2 * ox) * x_pix + c;
    const Vec16<OT> a = ld_vec16_stream(src), bb = ld_vec16_stream(src + x_pix);
    const Vec16<OT> cc = ld_vec16_stream(src + (int64_t)W * x_pix), d = ld_vec16_stream(src + (int64_t)(W + 1) * x_pix);
    Vec16<OT> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float f = 0.25f * ((to_f32<OT>(a.e[k]) + to_f32<OT>(bb.e[k])) + (to_f32<OT>(cc.e[k]) + to_f32<OT>(d.e[k])));
      if (ROUND_TF32) f = __uint_as_float(f32_to_tf32_rna(f));
      o.e[k] = from_f32<OT>(f);
    }
    st_vec16(y + ((int64_t)b * OH * OW + pp) * y_pix + c, o);
  }
}
}  // namespace

extern "C" int fmi_avgpool2_nhwc(const void* x, int64_t x_pixel_stride, void* y, int64_t y_pixel_stride, int B, int C, int H,
                                 int W, int round_y, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "avgpool2: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(x && y && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C >= vec && C % vec == 0 &&
                  x_pixel_stride % vec == 0 && y_pixel_stride % vec == 0 && x_pixel_stride >= C && y_pixel_stride >= C &&
                  fmi_aligned(x, 16) && fmi_aligned(y, 16) && B <= 65535,
              "avgpool2: unsupported shape C=%d H=%d W=%d", C, H, W);
  const int64_t work = (int64_t)(H / 2) * (W / 2) * (C / vec);
  int gx = stream_grid(work, 256 * 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (mma == FMI_MMA_TF32 && round_y)
    avgpool2_nhwc_kernel<float, true><<<dim3(gx, B), 256, 0, st>>>((const float*)x, x_pixel_stride, (float*)y, y_pixel_stride, C, H, W);
  else if (mma == FMI_MMA_TF32)
    avgpool2_nhwc_kernel<float, false><<<dim3(gx, B), 256, 0, st>>>((const float*)x, x_pixel_stride, (float*)y, y_pixel_stride, C, H, W);
  else
    avgpool2_nhwc_kernel<__nv_bfloat16, false><<<dim3(gx, B), 256, 0, st>>>((const __nv_bfloat16*)x, x_pixel_stride,
                                                                            (__nv_bfloat16*)y, y_pixel_stride, C, H, W);
  return fmi_launched("avgpool2");
}

// =====================================================================================================================
// Batched SpectralNorm + weight layout: the power iterations and re-layouts of ALL convolutions of a network do not depend
// on activations, so they run as three launches at the start of the forward (blockIdx.y = convolution) instead of three
// small launches per convolution (69 convs in the PICNet generator: 207 launches, 1.9 ms of mostly idle GPU).
// =====================================================================================================================
struct FmiSnPrepDesc {           // mirrored by picnet_fast._WeightPlan (struct format "6Q12i", 96 bytes)
  const float* w_bar;
  float* u;
  float* v;
  float* v_part;                 // [4][Wd] partial products of W^T u (row quarters), summed in fixed order
  float* u_raw;                  // [Hh]
  void* wp;
  int O, I, transposed, O_rows, I_row, i_off, merged, T, Hh, Wd, pad0, pad1;
};

namespace {
constexpr int SN_ZSPLIT = 4;

__global__ void __launch_bounds__(128) sn_batch_wt_u_kernel(const FmiSnPrepDesc* __restrict__ descs) {
  const FmiSnPrepDesc d = descs[blockIdx.y];
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= d.Wd) return;
  const int rows = (d.Hh + SN_ZSPLIT - 1) / SN_ZSPLIT;
  const int i0 = blockIdx.z * rows, i1 = min(d.Hh, i0 + rows);
  float a[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) a[q] = 0.f;
  int i = i0;
  for (; i + 8 <= i1; i += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fmaf(d.w_bar[(int64_t)(i + q) * d.Wd + j], d.u[i + q], a[q]);
  }
  for (; i < i1; ++i) a[0] = fmaf(d.w_bar[(int64_t)i * d.Wd + j], d.u[i], a[0]);
  d.v_part[(int64_t)blockIdx.z * d.Wd + j] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

__global__ void __launch_bounds__(256) sn_batch_w_v_kernel(const FmiSnPrepDesc* __restrict__ descs) {
  const FmiSnPrepDesc d = descs[blockIdx.y];
  if ((int)blockIdx.x * 8 >= d.Hh) return;
  __shared__ float red[8];
  extern __shared__ float sv[];     // v_raw of this convolution (Wd floats)
  float q = 0.f;
  for (int j = threadIdx.x; j < d.Wd; j += 256) {
    const float t = (d.v_part[j] + d.v_part[d.Wd + j]) + (d.v_part[2 * (int64_t)d.Wd + j] + d.v_part[3 * (int64_t)d.Wd + j]);
    sv[j] = t;
    q = fmaf(t, t, q);
  }
  const float inv = 1.f / (sqrtf(block_sum_256(q, red)) + 1e-12f);   // block_sum_256 synchronises: sv is complete
  if (blockIdx.x == 0)
    for (int j = threadIdx.x; j < d.Wd; j += 256) d.v[j] = sv[j] * inv;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= d.Hh) return;
  float a0 = 0.f, a1 = 0.f;
  int j = lane;
  for (; j + 32 < d.Wd; j += 64) {
    a0 = fmaf(d.w_bar[(int64_t)row * d.Wd + j], sv[j], a0);
    a1 = fmaf(d.w_bar[(int64_t)row * d.Wd + j + 32], sv[j + 32], a1);
  }
  if (j < d.Wd) a0 = fmaf(d.w_bar[(int64_t)row * d.Wd + j], sv[j], a0);
  const float a = warp_sum(a0 + a1) * inv;
  if (lane == 0) d.u_raw[row] = a;
}

template <typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) sn_batch_prep_kernel(const FmiSnPrepDesc* __restrict__ descs) {
  const FmiSnPrepDesc d = descs[blockIdx.y];
  const int total = d.O * d.I * d.T;
  if ((int)blockIdx.x * 256 >= total) return;
  __shared__ float red[8];
  float q = 0.f;
  for (int i = threadIdx.x; i < d.Hh; i += 256) q = fmaf(d.u_raw[i], d.u_raw[i], q);
  const float n2 = block_sum_256(q, red);
  const float inv = 1.f / (sqrtf(n2) + 1e-12f);
  const float sigma = n2 * inv;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < d.Hh; i += 256) d.u[i] = d.u_raw[i] * inv;
  OT* wp = (OT*)d.wp;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
    const int t = e % d.T, r = e / d.T;
    int o, i;
    if (d.transposed) { i = r / d.O; o = r - i * d.O; }
    else { o = r / d.I; i = r - o * d.I; }
    float v = d.w_bar[e] / sigma;
    if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
    int slab = t, row = o;
    if (d.merged) {
      const int ky = t / 3, kx = t - ky * 3;
      slab = (ky == 0 ? 2 : 0) + (kx == 0 ? 1 : 0);
      row = ((ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0)) * d.O + o;
    }
    wp[((int64_t)slab * d.O_rows + row) * d.I_row + d.i_off + i] = from_f32<OT>(v);
  }
}
}  // namespace

// descs: n FmiSnPrepDesc in DEVICE memory (include/fmi_b200.h documents the layout); max_wd / max_hh / max_elems: the largest
// Wd, Hh and O*I*T among them. Same results as n calls of fmi_conv_weight_prep_sn.
extern "C" int fmi_conv_weight_prep_sn_batch(const void* descs, int n, int max_wd, int max_hh, int max_elems, int mma,
                                             void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "conv_weight_prep_sn_batch: bad mma");
  if (n == 0) return FMI_OK;
  FMI_REQUIRE(descs && n > 0 && n <= 65535 && max_wd >= 1 && max_hh >= 1 && max_elems >= 1 && max_wd <= 12288,
              "conv_weight_prep_sn_batch: bad arguments");
  static_assert(sizeof(FmiSnPrepDesc) == 96, "descriptor layout");
  cudaStream_t st = (cudaStream_t)stream;
  const FmiSnPrepDesc* d = (const FmiSnPrepDesc*)descs;
  sn_batch_wt_u_kernel<<<dim3((max_wd + 127) / 128, n, SN_ZSPLIT), 128, 0, st>>>(d);
  int rc = fmi_launched("sn_batch_wt_u");
  if (rc) return rc;
  sn_batch_w_v_kernel<<<dim3((max_hh + 7) / 8, n), 256, (size_t)max_wd * sizeof(float), st>>>(d);
  rc = fmi_launched("sn_batch_w_v");
  if (rc) return rc;
  int gx = (max_elems + 256 * 8 - 1) / (256 * 8);
  if (gx < 1) gx = 1;
  if (mma == FMI_MMA_TF32 && fmi_tf32_exact_on()) sn_batch_prep_kernel<float, false><<<dim3(gx, n), 256, 0, st>>>(d);
  else if (mma == FMI_MMA_TF32) sn_batch_prep_kernel<float, true><<<dim3(gx, n), 256, 0, st>>>(d);
  else sn_batch_prep_kernel<__nv_bfloat16, false><<<dim3(gx, n), 256, 0, st>>>(d);
  return fmi_launched("sn_batch_prep");
}

// =====================================================================================================================
// Error-compensated TF32 operands ("3xTF32"). kind::tf32 reads the upper 19 bits of an fp32 operand; with x = hi + lo,
// hi = x with the low 13 mantissa bits cleared (exact in tf32) and lo = x - hi (exact in fp32, |lo| < 2^-10 |x|),
//     x . w  =  hi_x . hi_w  +  hi_x . lo_w  +  lo_x . hi_w  +  O(2^-20 |x| |w|)
// and every product is accumulated in fp32. The three terms are ONE implicit GEMM over three times the input channels:
// activations [hi | hi | lo], weights [hi | lo | hi] along I — no change to the GEMM kernel, K is 3x. This is what the PICNet
// conv blocks run when the caller asked for strict fp32 convolutions (torch.backends.cudnn.allow_tf32 = False) or
// FMI_PRECISION=tf32x3: whole-image error vs the fp32 reference <= 1e-3 with every convolution on the tensor cores.
// =====================================================================================================================
namespace {
template <int ORDER>   // 0: [hi | hi | lo] (activations), 1: [hi | lo | hi] (weights)
__global__ void __launch_bounds__(256) tf32_split3_kernel(const float* __restrict__ x, int64_t x_stride, float* __restrict__ y,
                                                          int64_t rows, int C) {
  const int nvec = C >> 2;
  const int64_t total = rows * nvec;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / nvec;
    const int c = (int)(e - r * nvec) << 2;
    const float4 v = *reinterpret_cast<const float4*>(x + r * x_stride + c);
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
    float* o = y + r * 3 * (int64_t)C + c;
    *reinterpret_cast<float4*>(o) = hi;
    *reinterpret_cast<float4*>(o + C) = ORDER ? lo : hi;
    *reinterpret_cast<float4*>(o + 2 * (int64_t)C) = ORDER ? hi : lo;
  }
}
}  // namespace

// x: `rows` rows of C fp32 values, x_stride elements apart -> y [rows][3*C] fp32. order 0: [hi | hi | lo], 1: [hi | lo | hi].
extern "C" int fmi_tf32_split3(const float* x, int64_t x_stride, float* y, int64_t rows, int C, int order, void* stream) {
  if (rows == 0) return FMI_OK;
  FMI_REQUIRE(x && y && rows > 0 && C >= 4 && C % 4 == 0 && x_stride >= C && x_stride % 4 == 0 && (order == 0 || order == 1) &&
                  fmi_aligned(x, 16) && fmi_aligned(y, 16),
              "tf32_split3: unsupported arguments (rows=%lld C=%d stride=%lld)", (long long)rows, C, (long long)x_stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = stream_grid(rows * (C / 4), 256 * 4);
  FmiProfScope prof(FMI_PROF_STREAM, st, 4.0 * rows * C, 16.0 * rows * C);
  if (order) tf32_split3_kernel<1><<<grid, 256, 0, st>>>(x, x_stride, y, rows, C);
  else tf32_split3_kernel<0><<<grid, 256, 0, st>>>(x, x_stride, y, rows, C);
  return fmi_launched("tf32_split3");
}
