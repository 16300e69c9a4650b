// a4: upfirdn2d — zero-insert upsample, pad/crop, FIR with the flipped kernel, decimate.
// Semantics: modules/psp/stylegan2/op/upfirdn2d_kernel.cu:52-137 (tap flip :77, extent :167-168) and
// the restatement op/upfirdn2d.py:150-184.
//
// Two kernels:
//  * upfirdn2d_fir_tile_kernel — the live hot path (up=1, down=1, taps <= 4x4, minor == 1: the Blur after
//    every up-sampling modulated conv and its backward). HBM-bound: one 32x128 output tile per CTA, the
//    (32+3)x(128+3) input tile staged once in shared memory with coalesced loads, 4x4 outputs per thread
//    from a 7x7 register window, 16-byte stores.
//  * upfirdn2d_generic_kernel — every other (up, down, pad, kernel <= 32x32, minor) configuration, gather
//    form, taps in shared memory. (The reference launches nothing for configurations outside its six
//    template modes; here all are computed.)
#include "common.cuh"

namespace {

constexpr int kMaxTaps = 32;

__host__ __device__ __forceinline__ int floor_div_i(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

struct UfdParams {
  int64_t major;
  int in_h, in_w, minor, kh, kw;
  int up_x, up_y, down_x, down_y;
  int pad_x0, pad_x1, pad_y0, pad_y1;
  int out_h, out_w;
};

template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(const T* __restrict__ x, const float* __restrict__ k,
                                                                T* __restrict__ y, UfdParams p) {
  __shared__ float sk[kMaxTaps * kMaxTaps];
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    int ky = i / p.kw, kx = i % p.kw;
    sk[i] = k[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];  // flipped taps
  }
  __syncthreads();
  const int64_t total = p.major * p.out_h * (int64_t)p.out_w * p.minor;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int m = (int)(idx % p.minor);
    int64_t t = idx / p.minor;
    int ox = (int)(t % p.out_w);
    t /= p.out_w;
    int oy = (int)(t % p.out_h);
    int64_t mj = t / p.out_h;
    const T* xp = x + mj * p.in_h * (int64_t)p.in_w * p.minor + m;
    float acc = 0.f;
    const int y0 = oy * p.down_y - p.pad_y0;
    const int x0 = ox * p.down_x - p.pad_x0;
    for (int ky = 0; ky < p.kh; ++ky) {
      int yy = y0 + ky;
      if (yy < 0) continue;
      int iy = yy / p.up_y;
      if (iy * p.up_y != yy || iy >= p.in_h) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        int xx = x0 + kx;
        if (xx < 0) continue;
        int ix = xx / p.up_x;
        if (ix * p.up_x != xx || ix >= p.in_w) continue;
        acc += to_f32<T>(xp[((int64_t)iy * p.in_w + ix) * p.minor]) * sk[ky * p.kw + kx];
      }
    }
    y[idx] = from_f32<T>(acc);
  }
}

constexpr int TOH = 32, TOW = 128;          // output tile
constexpr int TIH = TOH + 3, TIW = TOW + 4;  // staged input tile (row padded to a multiple of 4)

template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_fir_tile_kernel(const T* __restrict__ x, const float* __restrict__ k,
                                                                 T* __restrict__ y, UfdParams p, int tiles_x,
                                                                 int tiles_y) {
  __shared__ float sk[16];
  __shared__ __align__(16) float sx[TIH][TIW];
  int bid = blockIdx.x;
  const int tx_i = bid % tiles_x;
  bid /= tiles_x;
  const int ty_i = bid % tiles_y;
  const int64_t plane = bid / tiles_y;
  const int oy0 = ty_i * TOH, ox0 = tx_i * TOW;
  const int iy0 = oy0 - p.pad_y0, ix0 = ox0 - p.pad_x0;

  if (threadIdx.x < 16) {
    int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    float v = 0.f;
    if (ky < p.kh && kx < p.kw) v = k[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
    sk[threadIdx.x] = v;
  }
  const T* xp = x + plane * p.in_h * (int64_t)p.in_w;
  for (int i = threadIdx.x; i < TIH * TIW; i += 256) {
    int r = i / TIW, c = i % TIW;
    int iy = iy0 + r, ix = ix0 + c;
    float v = 0.f;
    if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = to_f32<T>(xp[(int64_t)iy * p.in_w + ix]);
    sx[r][c] = v;
  }
  __syncthreads();

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 column groups x 8 row groups
  const int lx = tx * 4, ly = ty * 4;
  float kk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kk[i] = sk[i];
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    float4 w0 = *reinterpret_cast<const float4*>(&sx[ly + r][lx]);
    float4 w1 = *reinterpret_cast<const float4*>(&sx[ly + r][lx + 4]);
    float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ky = r - a;  // output row a uses input row a+ky
      if (ky < 0 || ky > 3) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) acc[a][b] = fmaf(w[b + kx], kk[ky * 4 + kx], acc[a][b]);
    }
  }
  T* yp = y + plane * p.out_h * (int64_t)p.out_w;
  const int ox = ox0 + lx;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int oy = oy0 + ly + a;
    if (oy >= p.out_h || ox >= p.out_w) continue;
    T* dst = yp + (int64_t)oy * p.out_w + ox;
    if (ox + 4 <= p.out_w && fmi_aligned_dev(dst, 4 * sizeof(T))) {
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      } else {
        union { uint2 u; T e[4]; } pk;
#pragma unroll
        for (int b = 0; b < 4; ++b) pk.e[b] = from_f32<T>(acc[a][b]);
        *reinterpret_cast<uint2*>(dst) = pk.u;
      }
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (ox + b < p.out_w) dst[b] = from_f32<T>(acc[a][b]);
    }
  }
}

}  // namespace

extern "C" int fmi_upfirdn2d_out_size(int in, int up, int down, int pad0, int pad1, int k) {
  if (up < 1 || down < 1 || k < 1) return -1;
  int num = in * up + pad0 + pad1 - k + down;
  if (num <= 0) return -1;
  return num / down;
}

extern "C" int fmi_upfirdn2d(const void* x, const float* kernel, void* y, int64_t major, int in_h, int in_w, int minor,
                             int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                             int pad_y0, int pad_y1, int dtype, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype), "upfirdn2d: unsupported dtype %d", dtype);
  FMI_REQUIRE(major >= 0 && in_h >= 1 && in_w >= 1 && minor >= 1, "upfirdn2d: bad input shape");
  FMI_REQUIRE(kh >= 1 && kw >= 1 && kh <= kMaxTaps && kw <= kMaxTaps, "upfirdn2d: kernel %dx%d outside 1..%d", kh, kw,
              kMaxTaps);
  FMI_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
  UfdParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_x1 = pad_x1; p.pad_y0 = pad_y0; p.pad_y1 = pad_y1;
  p.out_h = fmi_upfirdn2d_out_size(in_h, up_y, down_y, pad_y0, pad_y1, kh);
  p.out_w = fmi_upfirdn2d_out_size(in_w, up_x, down_x, pad_x0, pad_x1, kw);
  FMI_REQUIRE(p.out_h >= 1 && p.out_w >= 1, "upfirdn2d: empty output extent (%d x %d)", p.out_h, p.out_w);
  if (major == 0) return FMI_OK;
  FMI_REQUIRE(x && kernel && y, "upfirdn2d: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool fir_tile = up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kh <= 4 && kw <= 4 && minor == 1;
  FMI_DISPATCH_DTYPE(dtype, T, {
    if (fir_tile) {
      int tiles_x = (p.out_w + TOW - 1) / TOW, tiles_y = (p.out_h + TOH - 1) / TOH;
      int64_t blocks = (int64_t)tiles_x * tiles_y * major;
      FMI_REQUIRE(blocks < (1ll << 31), "upfirdn2d: tensor too large");
      upfirdn2d_fir_tile_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, kernel, (T*)y, p, tiles_x, tiles_y);
    } else {
      int64_t total = major * p.out_h * (int64_t)p.out_w * minor;
      int grid = (int)imin64((total + 255) / 256, (int64_t)FMI_NUM_SMS * 32);
      upfirdn2d_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)x, kernel, (T*)y, p);
    }
  });
  return fmi_launched("upfirdn2d");
}
