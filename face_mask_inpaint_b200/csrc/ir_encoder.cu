// f2: the pSp encoder — IR-SE50 trunk + the 18 map2style heads + FPN adds of GradualStyleEncoder
// (modules/psp/encoders/psp_encoders.py:13-37, 100-152; units: encoders/helpers.py:56-119) — in inference, on the tcgen05
// implicit-GEMM kernel of modconv_gemm.cuh plus five streaming kernels. Activations are NHWC in the tensor-core operand type
// (bf16, or tf32-rounded fp32) from the input image to the style codes.
//
// One bottleneck_IR_SE unit (helpers.py:97-119), eval-mode BatchNorm folded on the host side (modules/psp_fast.py):
//   a1 = PReLU(conv3x3(BN1(x)))        ONE GEMM: BN1's scale goes into the weights per INPUT channel; its shift cannot (the conv
//                                      zero-pads BN1's OUTPUT, so the shift only reaches a pixel through the taps that are inside
//                                      the image): it becomes a bias per output channel AND border class (9 classes: top / inside /
//                                      bottom x left / inside / right), added in the epilogue together with the per-channel PReLU
//   r  = BN2(conv3x3_stride_s(a1))     ONE GEMM, BN2 folded into weights (per output channel) + bias. Stride 2: a1 is first
//                                      re-laid as its 4 pixel-parity planes [4][B][H/2][W/2][C] (space_to_planes); input row
//                                      2m + dy then is row m + (dy < 0 ? -1 : 0) of plane (dy & 1), i.e. the 9 taps are plain
//                                      TMA boxes again (zero fill at -1 = the padding)
//   s  = sigmoid(fc2(relu(fc1(mean_hw(r)))))     channel_sum (slab partial sums, no atomics) + se_gate (one CTA per image)
//   y  = r * s + shortcut(x)           se_scale_add; shortcut = x subsampled (MaxPool2d(1, s)) or BN(conv1x1_stride_s(x)) = ONE
//                                      GEMM reading x through a strided tensor map
// map2style heads (psp_encoders.py:13-37): log2(spatial) x [conv3x3 stride 2 + bias + LeakyReLU(0.01)] down to 1x1, then an
// EqualLinear. All heads that read the same pyramid level run as ONE GEMM per depth: first level = one conv with the heads'
// weights concatenated along O; deeper levels = heads as extra batch entries with one weight set per head (w_group), several
// whole images per 128-row tile once the planes are tiny (TB).
#include "modconv_gemm.cuh"

using namespace sm100;
using namespace fmi_conv;

namespace {
inline int sgrid(int64_t work_items, int per_block) {
  int64_t g = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)FMI_NUM_SMS * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// x [B][H][W][heads*C] (pixel stride xs) -> y [4][heads*B][H/2][W/2][C]: plane p = 2*(row & 1) + (col & 1), batch entry
// head * B + b. 16-byte vectors; one thread per output vector.
template <typename T>
__global__ void __launch_bounds__(256) space_to_planes_kernel(const T* __restrict__ x, int64_t xs, T* __restrict__ y, int B, int C,
                                                              int H, int W, int heads) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V, h2 = H / 2, w2 = W / 2;
  const int64_t total = (int64_t)4 * heads * B * h2 * w2 * cv;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e;
    const int c = (int)(t % cv); t /= cv;
    const int n = (int)(t % w2); t /= w2;
    const int m = (int)(t % h2); t /= h2;
    const int hb = (int)(t % ((int64_t)heads * B)); t /= (int64_t)heads * B;
    const int pl = (int)t;
    const int head = hb / B, b = hb - head * B;
    const T* src = x + (((int64_t)b * H + 2 * m + (pl >> 1)) * W + 2 * n + (pl & 1)) * xs + (int64_t)head * C + c * V;
    *reinterpret_cast<uint4*>(y + e * V) = *reinterpret_cast<const uint4*>(src);
  }
}

// Sum over a slab of pixel rows per (image, channel): x [B][HW][C] dense -> part [B][gridDim.x][C] fp32. No atomics: thread
// partials go through shared memory and are added in a fixed order, so the result is bit-reproducible (eager run == CUDA-graph
// replay; the sums feed a sigmoid gate whose product is then rounded to the operand type — a last-bit difference of the mean
// would flip roundings downstream).
constexpr int kSeMaxSlabs = 32;
template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* __restrict__ x, float* __restrict__ part, int C, int HW,
                                                          int rows_per_block) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float sacc[256 * V];             // [lanes][C], lanes * C = 256 * V
  const int cv = C / V;                       // vectors per pixel (<= 256)
  const int b = blockIdx.y;
  const int lanes = 256 / cv;                 // pixel rows handled concurrently
  const int vc = threadIdx.x % cv, lane = threadIdx.x / cv;
  if (lane < lanes) {
    const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    const T* base = x + (int64_t)b * HW * C + vc * V;
    for (int r = r0 + lane; r < r1; r += lanes) {
      const Vec16<T> v = ld_vec16(base + (int64_t)r * C);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += to_f32<T>(v.e[k]);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) sacc[lane * C + vc * V + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += sacc[l * C + c];
    part[((int64_t)b * gridDim.x + blockIdx.x) * C + c] = t;
  }
}

// SEModule gate (helpers.py:56-74): m = mean_hw r (the slab sums added in slab order), g[b][c] = sigmoid(W2 relu(W1 m[b])),
// W1 [R][C], W2 [C][R]; one CTA per image.
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ part, int slabs, float inv_hw,
                                                      const float* __restrict__ w1, const float* __restrict__ w2,
                                                      float* __restrict__ mean, float* __restrict__ gate, int C, int R) {
  extern __shared__ float sm[];   // [C] mean, [R] hidden
  float* m = sm;
  float* hid = sm + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int s = 0; s < slabs; ++s) t += part[((int64_t)b * slabs + s) * C + c];
    m[c] = t * inv_hw;
    mean[(int64_t)b * C + c] = m[c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += 8) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(w1[(int64_t)r * C + c], m[c], a);
    a = warp_sum(a);
    if (lane == 0) hid[r] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f;
    for (int r = 0; r < R; ++r) a = fmaf(w2[(int64_t)c * R + r], hid[r], a);
    gate[(int64_t)b * C + c] = 1.f / (1.f + __expf(-a));
  }
}

// y = r * gate[b][c] + sc,  r / y dense [B][H][W][C], sc read through (pixel, row, image) strides (a subsampled x or a dense tensor)
template <typename T, bool ROUND_TF32>
__global__ void __launch_bounds__(256) se_scale_add_kernel(const T* __restrict__ r, const float* __restrict__ gate,
                                                           const T* __restrict__ sc, int64_t sc_ps, int64_t sc_rs, int64_t sc_is,
                                                           T* __restrict__ y, int C, int H, int W, int64_t total) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e;
    const int c = (int)(t % cv) * V; t /= cv;
    const int n = (int)(t % W); t /= W;
    const int m = (int)(t % H); t /= H;
    const int b = (int)t;
    const Vec16<T> rv = ld_vec16(r + e * V);
    const Vec16<T> sv = ld_vec16(sc + (int64_t)b * sc_is + (int64_t)m * sc_rs + (int64_t)n * sc_ps + c);
    const float* g = gate + (int64_t)b * C + c;
    Vec16<T> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float v = fmaf(to_f32<T>(rv.e[k]), __ldg(g + k), to_f32<T>(sv.e[k]));
      if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
      o.e[k] = from_f32<T>(v);
    }
    st_vec16(y + e * V, o);
  }
}

// y[b][oy][ox][:] = bilinear(x[b], align_corners = True)(oy, ox) + add[b][oy][ox][:]   (psp_encoders.py:83-98 _upsample_add)
template <typename T, bool ROUND_TF32>
__global__ void __launch_bounds__(256) upsample_add_kernel(const T* __restrict__ x, const T* __restrict__ add, T* __restrict__ y,
                                                           int C, int h, int w, int OH, int OW, float ry, float rx, int64_t total) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e;
    const int c = (int)(t % cv) * V; t /= cv;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH); t /= OH;
    const int b = (int)t;
    const float fy = oy * ry, fx = ox * rx;
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - y0, lx = fx - x0;
    const T* xb = x + (int64_t)b * h * w * C + c;
    const Vec16<T> v00 = ld_vec16(xb + ((int64_t)y0 * w + x0) * C), v01 = ld_vec16(xb + ((int64_t)y0 * w + x1) * C);
    const Vec16<T> v10 = ld_vec16(xb + ((int64_t)y1 * w + x0) * C), v11 = ld_vec16(xb + ((int64_t)y1 * w + x1) * C);
    const Vec16<T> av = ld_vec16(add + e * V);
    Vec16<T> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      // ATen's upsample_bilinear2d order: (1-ly) * ((1-lx) v00 + lx v01) + ly * ((1-lx) v10 + lx v11)
      const float top = (1.f - lx) * to_f32<T>(v00.e[k]) + lx * to_f32<T>(v01.e[k]);
      const float bot = (1.f - lx) * to_f32<T>(v10.e[k]) + lx * to_f32<T>(v11.e[k]);
      float v = (1.f - ly) * top + ly * bot + to_f32<T>(av.e[k]);
      if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
      o.e[k] = from_f32<T>(v);
    }
    st_vec16(y + e * V, o);
  }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// General NHWC convolution on the implicit-GEMM kernel (see include/fmi_b200.h).
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int fmi_conv_nhwc(const void* x, int64_t x_pixel_stride, int64_t x_row_stride, int64_t x_img_stride, const void* wp,
                             const float* bias, int bias_classes, const float* slope_c, float slope, void* y,
                             int64_t y_pixel_stride, int B, int I, int O, int H, int W, int ksize, int planes, int w_group,
                             int bias_per_set, int act, int add_y, int round_y, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "conv_nhwc: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && wp && y, "conv_nhwc: null pointer");
  FMI_REQUIRE((ksize == 3 || ksize == 1) && (planes == 0 || (planes == 1 && ksize == 3)), "conv_nhwc: ksize 1 or 3; planes only with 3");
  FMI_REQUIRE(act == 1 || act == 2 || act == 4, "conv_nhwc: act must be 1 (leaky relu), 2 (bias) or 4 (PReLU)");
  FMI_REQUIRE(act != 4 || slope_c, "conv_nhwc: act 4 needs slope_c");
  FMI_REQUIRE(bias_classes == 1 || (bias_classes == 9 && ksize == 3 && !planes && H >= 2 && W >= 2 && bias),
              "conv_nhwc: bias_classes is 1, or 9 for a stride-1 3x3 conv on a plane of at least 2x2");
  FMI_REQUIRE(!add_y || act == 2, "conv_nhwc: the residual sum needs act 2");
  const int esz = esz_of(mma);
  FMI_REQUIRE(I >= 16 && O >= 32 && O % 32 == 0 && (O <= 256 || O % 256 == 0) && H >= 1 && W >= 1,
              "conv_nhwc: unsupported shape I=%d O=%d H=%d W=%d (O must be a multiple of 32, <= 256 or a multiple of 256)", I, O, H, W);
  FMI_REQUIRE((I * esz) % 16 == 0 && (x_pixel_stride * esz) % 16 == 0 && (x_row_stride * esz) % 16 == 0 &&
                  (x_img_stride * esz) % 16 == 0 && x_pixel_stride >= I && fmi_aligned(x, 16) && fmi_aligned(wp, 16),
              "conv_nhwc: input rows / strides must be 16-byte multiples");
  FMI_REQUIRE((y_pixel_stride * esz) % 16 == 0 && y_pixel_stride >= O && fmi_aligned(y, 16), "conv_nhwc: output rows must be 16-byte multiples");
  FMI_REQUIRE(w_group >= 0 && (w_group <= 1 || B % w_group == 0), "conv_nhwc: B must be a multiple of w_group");
  int rc = fmi_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tf32 = mma == FMI_MMA_TF32;
  const uint32_t epa = 128 / esz;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int T = ksize * ksize;

  ConvGemmParams p{};
  p.B = B; p.I = I; p.O = O; p.H = H; p.W = W; p.T = T;
  p.OH = H; p.OW = W; p.Mh = H; p.Mw = W; p.sy = p.sx = 1;
  p.w_shared = w_group == 0;
  p.w_group = w_group;
  p.raw_out = !round_y;
  p.add_out = add_y;
  p.n_tile = O <= 256 ? O : 256;
  p.k_chunks = (I + epa - 1) / epa;
  p.bias = bias; p.bias_classes = bias_classes; p.slope_c = slope_c;
  p.bias_set_stride = (bias_per_set && w_group >= 1) ? O * bias_classes : 0;
  p.act = act; p.slope = slope; p.gain = 1.f;
  p.out = y;
  p.out_pstride = (int)y_pixel_stride;
  p.out_rstride = (int64_t)W * y_pixel_stride;
  p.out_bstride = (int64_t)H * W * y_pixel_stride;
  p.prof_kind = FMI_PROF_GEMM_IR;
  p.ntaps = T;
  const int n_sets = w_group == 0 ? 1 : B / (w_group > 1 ? w_group : 1);
  for (int t = 0; t < T; ++t) {
    const int dy = ksize == 3 ? t / 3 - 1 : 0, dx = ksize == 3 ? t % 3 - 1 : 0;
    p.tap_slab[t] = t;
    if (planes) {   // input row 2m + dy = row m + (dy < 0 ? -1 : 0) of parity plane (dy & 1); x is [4][B][H][W][*]
      p.tap_dy[t] = dy < 0 ? -1 : 0;
      p.tap_dx[t] = dx < 0 ? -1 : 0;
      p.tap_boff[t] = ((dy & 1) * 2 + (dx & 1)) * B;
    } else {
      p.tap_dy[t] = dy;
      p.tap_dx[t] = dx;
    }
  }
  // several whole images per tile once a plane is much smaller than the 128-row tile
  int tb = 1;
  if (H * W <= 64) {
    const int lim = w_group == 0 ? B : (w_group > 1 ? w_group : 1);
    while (tb * 2 * H * W <= 128 && tb * 2 <= lim && lim % (tb * 2) == 0) tb *= 2;
  }
  p.TB = tb;
  TilePlan tp = tb > 1 ? TilePlan{H, W, 1, 1} : pick_tile(H, W);
  p.n_tile = pick_n_tile(O, p.n_tile, (int64_t)((B + tb - 1) / tb) * tp.tiles_h * tp.tiles_w);   // few tiles: narrower, on more SMs
  static const bool halo_off = [] { const char* e = getenv("FMI_CONV_HALO"); return e && e[0] == '0'; }();
  p.halo = ksize == 3 && !planes && tb == 1 && !halo_off && W >= 128 && p.n_tile <= 128 && w_group == 0;
  if (p.halo) {
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) p.halo_slab[a][c] = a * 3 + c;
    tp = TilePlan{1, 130, 0, 0};
  }
  CUtensorMap mw, mx;
  {
    uint64_t dims[2] = {(uint64_t)I, (uint64_t)n_sets * T * O};
    uint64_t str[1] = {(uint64_t)I * esz};
    uint32_t box[2] = {epa, (uint32_t)p.n_tile};
    int e = make_tensor_map(&mw, dt, 2, wp, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "conv_nhwc: cuTensorMapEncodeTiled(weights) failed (%d)", e);
  }
  {
    uint64_t dims[4] = {(uint64_t)I, (uint64_t)W, (uint64_t)H, (uint64_t)(planes ? 4 * B : B)};
    uint64_t str[3] = {(uint64_t)x_pixel_stride * esz, (uint64_t)x_row_stride * esz, (uint64_t)x_img_stride * esz};
    uint32_t box[4] = {epa, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tb};
    int e = make_tensor_map(&mx, dt, 4, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "conv_nhwc: cuTensorMapEncodeTiled(x) failed (%d)", e);
  }
  return tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
}

extern "C" int fmi_space_to_planes_nhwc(const void* x, int64_t x_pixel_stride, void* y, int B, int C, int H, int W, int heads,
                                        int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "space_to_planes: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(x && y && heads >= 1 && C >= vec && C % vec == 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 &&
                  x_pixel_stride >= (int64_t)heads * C && x_pixel_stride % vec == 0 && fmi_aligned(x, 16) && fmi_aligned(y, 16),
              "space_to_planes: unsupported shape C=%d H=%d W=%d heads=%d", C, H, W, heads);
  const int64_t total = (int64_t)heads * B * H * W * (C / vec);
  cudaStream_t st = (cudaStream_t)stream;
  FmiProfScope prof(FMI_PROF_STREAM, st, 0.0, 2.0 * total * 16);
  if (mma == FMI_MMA_TF32)
    space_to_planes_kernel<float><<<sgrid(total, 256 * 2), 256, 0, st>>>((const float*)x, x_pixel_stride, (float*)y, B, C, H, W, heads);
  else
    space_to_planes_kernel<__nv_bfloat16><<<sgrid(total, 256 * 2), 256, 0, st>>>((const __nv_bfloat16*)x, x_pixel_stride,
                                                                                (__nv_bfloat16*)y, B, C, H, W, heads);
  return fmi_launched("space_to_planes");
}

extern "C" int fmi_se_gate_nhwc(const void* r, const float* w1, const float* w2, float* scratch, float* mean, float* gate, int B,
                                int C, int R, int HW, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "se_gate: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(r && w1 && w2 && scratch && mean && gate && C >= vec && C % vec == 0 && C / vec <= 256 && R >= 1 && R <= 256 &&
                  HW >= 1 && B <= 65535 && fmi_aligned(r, 16),
              "se_gate: unsupported shape C=%d R=%d", C, R);
  cudaStream_t st = (cudaStream_t)stream;
  FmiProfScope prof(FMI_PROF_SE, st, 2.0 * B * HW * C, (double)B * HW * C * esz_of(mma));
  const int lanes = 256 / (C / vec);
  int gx = (HW + lanes * 16 - 1) / (lanes * 16);     // ~16 pixel rows per thread
  int cap = (FMI_NUM_SMS * 8 + B - 1) / B;
  if (cap > kSeMaxSlabs) cap = kSeMaxSlabs;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const int rows = (HW + gx - 1) / gx;
  gx = (HW + rows - 1) / rows;
  if (mma == FMI_MMA_TF32)
    channel_sum_kernel<float><<<dim3(gx, B), 256, 0, st>>>((const float*)r, scratch, C, HW, rows);
  else
    channel_sum_kernel<__nv_bfloat16><<<dim3(gx, B), 256, 0, st>>>((const __nv_bfloat16*)r, scratch, C, HW, rows);
  int rc = fmi_launched("channel_sum");
  if (rc) return rc;
  se_gate_kernel<<<B, 256, (size_t)(C + R) * sizeof(float), st>>>(scratch, gx, 1.f / (float)HW, w1, w2, mean, gate, C, R);
  return fmi_launched("se_gate");
}

extern "C" int fmi_se_scale_add_nhwc(const void* r, const float* gate, const void* sc, int64_t sc_pixel_stride,
                                     int64_t sc_row_stride, int64_t sc_img_stride, void* y, int B, int C, int H, int W,
                                     int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "se_scale_add: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(r && gate && sc && y && C >= vec && C % vec == 0 && H >= 1 && W >= 1 && sc_pixel_stride % vec == 0 &&
                  sc_row_stride % vec == 0 && sc_img_stride % vec == 0 && fmi_aligned(r, 16) && fmi_aligned(sc, 16) &&
                  fmi_aligned(y, 16),
              "se_scale_add: unsupported shape C=%d", C);
  const int64_t total = (int64_t)B * H * W * (C / vec);
  cudaStream_t st = (cudaStream_t)stream;
  FmiProfScope prof(FMI_PROF_SE, st, 2.0 * total * vec, 3.0 * total * 16);
  if (mma == FMI_MMA_TF32)
    se_scale_add_kernel<float, true><<<sgrid(total, 256 * 2), 256, 0, st>>>((const float*)r, gate, (const float*)sc, sc_pixel_stride,
                                                                           sc_row_stride, sc_img_stride, (float*)y, C, H, W, total);
  else
    se_scale_add_kernel<__nv_bfloat16, false><<<sgrid(total, 256 * 2), 256, 0, st>>>(
        (const __nv_bfloat16*)r, gate, (const __nv_bfloat16*)sc, sc_pixel_stride, sc_row_stride, sc_img_stride, (__nv_bfloat16*)y, C,
        H, W, total);
  return fmi_launched("se_scale_add");
}

extern "C" int fmi_upsample_add_nhwc(const void* x, const void* add, void* y, int B, int C, int h, int w, int OH, int OW, int mma,
                                     void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "upsample_add: bad mma");
  if (B == 0) return FMI_OK;
  const int vec = 16 / esz_of(mma);
  FMI_REQUIRE(x && add && y && C >= vec && C % vec == 0 && h >= 1 && w >= 1 && OH >= 1 && OW >= 1 && fmi_aligned(x, 16) &&
                  fmi_aligned(add, 16) && fmi_aligned(y, 16),
              "upsample_add: unsupported shape C=%d", C);
  const int64_t total = (int64_t)B * OH * OW * (C / vec);
  const float ry = OH > 1 ? (float)(h - 1) / (float)(OH - 1) : 0.f, rx = OW > 1 ? (float)(w - 1) / (float)(OW - 1) : 0.f;
  cudaStream_t st = (cudaStream_t)stream;
  FmiProfScope prof(FMI_PROF_STREAM, st, 8.0 * total * vec, 2.0 * total * 16 + (double)B * h * w * C * esz_of(mma));
  if (mma == FMI_MMA_TF32)
    upsample_add_kernel<float, true><<<sgrid(total, 256 * 2), 256, 0, st>>>((const float*)x, (const float*)add, (float*)y, C, h, w, OH,
                                                                           OW, ry, rx, total);
  else
    upsample_add_kernel<__nv_bfloat16, false><<<sgrid(total, 256 * 2), 256, 0, st>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)add, (__nv_bfloat16*)y, C, h, w, OH, OW, ry, rx, total);
  return fmi_launched("upsample_add");
}
