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
#include <stdlib.h>

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

// ---- tiled polyphase FIR: up in {1,2}, down in {1,2}, taps <= 4x4, minor == 1 ------------------------------------------
// One CTA = 256 threads = 8 x 32 thread grid, OT x OT outputs per thread (output tile 8*OT rows x 32*OT columns). The
// input window of the tile is staged once in shared memory (row-wise, coalesced, zero-filled outside the image = the
// padding); every thread then reads its register window with vector loads and applies only the taps that hit a real
// sample: with zero-insertion by UP the tap (ky, kx) of output (a, b) is live iff (a*DOWN + ky + CY) % UP == 0, which is
// static because tile and thread origins are multiples of UP and CY = (-pad0) mod UP is a launch constant.
template <int UP, int DOWN, int OT>
struct FirTile {
  static constexpr int TOH = 8 * OT, TOW = 32 * OT;
  static constexpr int WR = ((OT - 1) * DOWN + 3 + (UP - 1)) / UP + 1;  // register window rows/cols per thread
  static constexpr int STEP = OT * DOWN / UP;                            // window origin step between adjacent threads
  static constexpr int WCV = (WR + 3) / 4 * 4;                           // window columns rounded up to whole vectors
  static constexpr int TIH = 7 * STEP + WR;
  static constexpr int TIW = (31 * STEP + WCV + 3) / 4 * 4;
  static_assert((OT * DOWN) % UP == 0, "thread origin must land on a real sample");
};

template <typename T, int UP, int DOWN, int OT, int CY, int CX>
__global__ void __launch_bounds__(256) upfirdn2d_tile_kernel(const T* __restrict__ x, const float* __restrict__ k,
                                                             T* __restrict__ y, UfdParams p, int tiles_x, float inv_tiles_x) {
  using F = FirTile<UP, DOWN, OT>;
  __shared__ float sk[16];
  __shared__ __align__(16) float sx[F::TIH][F::TIW];
  // grid (tiles of one plane, planes): tile rows via a float reciprocal (exact below 2^23 tiles) — the four integer
  // divisions of a flat 1-D decode were ~15 % of the kernel's instructions, and a 3-D grid measured 6 % slower
  const int64_t plane = blockIdx.y;
  const int ty_i = (int)(((float)blockIdx.x + 0.5f) * inv_tiles_x);
  const int oy0 = ty_i * F::TOH, ox0 = ((int)blockIdx.x - ty_i * tiles_x) * F::TOW;
  // first input sample the tile can touch: floor((o0*DOWN - pad0) / UP); the remainder is CY / CX by construction
  const int iy0 = (oy0 * DOWN - p.pad_y0 - CY) / UP, ix0 = (ox0 * DOWN - p.pad_x0 - CX) / UP;

  if (threadIdx.x < 16) {
    const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    float v = 0.f;
    if (ky < p.kh && kx < p.kw) v = k[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];  // flipped taps (upfirdn2d_kernel.cu:77)
    sk[threadIdx.x] = v;
  }
  const T* xp = x + plane * p.in_h * (int64_t)p.in_w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (iy0 >= 0 && iy0 + F::TIH <= p.in_h && ix0 >= 0 && ix0 + F::TIW <= p.in_w) {
    // interior tile (all but the image border): no bounds tests, one add per element
    const T* src = xp + (int64_t)iy0 * p.in_w + ix0 + lane;
    for (int r = warp; r < F::TIH; r += 8) {
      const T* row = src + r * p.in_w;
#pragma unroll
      for (int c = 0; c < F::TIW; c += 32)
        if (c + 32 <= F::TIW || lane < F::TIW - c) sx[r][c + lane] = to_f32<T>(row[c]);
    }
  } else {
    for (int r = warp; r < F::TIH; r += 8) {
      const int iy = iy0 + r;
      const bool row_ok = iy >= 0 && iy < p.in_h;
      const T* row = xp + (int64_t)iy * p.in_w;
#pragma unroll
      for (int c = lane; c < F::TIW; c += 32) {
        const int ix = ix0 + c;
        float v = 0.f;
        if (row_ok && ix >= 0 && ix < p.in_w) v = to_f32<T>(row[ix]);
        sx[r][c] = v;
      }
    }
  }
  __syncthreads();

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int wy = ty * F::STEP, wx = tx * F::STEP;  // window origin in the staged tile
  float kk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kk[i] = sk[i];
  float acc[OT][OT];
#pragma unroll
  for (int a = 0; a < OT; ++a)
#pragma unroll
    for (int b = 0; b < OT; ++b) acc[a][b] = 0.f;
#pragma unroll
  for (int r = 0; r < F::WR; ++r) {
    float w[F::WCV];
    if constexpr (F::STEP % 4 == 0) {
#pragma unroll
      for (int v = 0; v < F::WCV; v += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&sx[wy + r][wx + v]);
        w[v] = t.x; w[v + 1] = t.y; w[v + 2] = t.z; w[v + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int v = 0; v < F::WCV; v += 2) {
        const float2 t = *reinterpret_cast<const float2*>(&sx[wy + r][wx + v]);
        w[v] = t.x; w[v + 1] = t.y;
      }
    }
#pragma unroll
    for (int a = 0; a < OT; ++a)
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        if ((a * DOWN + ky + CY) % UP != 0 || (a * DOWN + ky + CY) / UP != r) continue;
#pragma unroll
        for (int b = 0; b < OT; ++b)
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            if ((b * DOWN + kx + CX) % UP != 0) continue;
            acc[a][b] = fmaf(w[(b * DOWN + kx + CX) / UP], kk[ky * 4 + kx], acc[a][b]);
          }
      }
  }
  T* yp = y + plane * p.out_h * (int64_t)p.out_w;
  const int ox = ox0 + tx * OT;
#pragma unroll
  for (int a = 0; a < OT; ++a) {
    const int oy = oy0 + ty * OT + a;
    if (oy >= p.out_h || ox >= p.out_w) continue;
    T* dst = yp + (int64_t)oy * p.out_w + ox;
    if (ox + OT <= p.out_w && fmi_aligned_dev(dst, OT * sizeof(T))) {
      union { uint4 u4; uint2 u2; uint32_t u1; T e[OT]; } pk;
#pragma unroll
      for (int b = 0; b < OT; ++b) pk.e[b] = from_f32<T>(acc[a][b]);
      if constexpr (OT * sizeof(T) == 16) *reinterpret_cast<uint4*>(dst) = pk.u4;
      else if constexpr (OT * sizeof(T) == 8) *reinterpret_cast<uint2*>(dst) = pk.u2;
      else *reinterpret_cast<uint32_t*>(dst) = pk.u1;
    } else {
#pragma unroll
      for (int b = 0; b < OT; ++b)
        if (ox + b < p.out_w) dst[b] = from_f32<T>(acc[a][b]);
    }
  }
}

// ---- separable strip kernel: up = down = 1, taps <= 4x4, minor == 1 — Blur forward (pad 1,1) and backward (pad 2,2) ----------
// The 16-FMA-per-output tile kernel above is instruction-issue bound AND keeps too few bytes in flight (its staging loads pass
// through registers a row at a time): 0.68 / 0.38 of the HBM roofline in fp32 / bf16. Here:
//  * staging is asynchronous: every thread issues ALL its global -> shared copies of the tile (cp.async, 4-byte units, zero fill
//    outside the image = the padding) before it waits, so a CTA has its whole 67 x 131 input window in flight and the other
//    CTAs of the SM compute meanwhile. Rows of the (2H+1)-wide tensors are not 16-byte aligned, 4-byte units always are for
//    fp32; a 16-bit row is copied as the aligned 32-bit words that cover it, which leaves the row shifted by one slot when its
//    first element sits at an odd index (the window reader funnel-shifts it back).
//  * every blur kernel the models build is an outer product (stylegan2/model.py:19-27 make_kernel), so the taps are factored on
//    the device (pivot row x pivot column / pivot; a CTA whose taps are not rank one takes the 16-tap sum from the same staged
//    tile): a horizontal 4-tap pass on each staged row, then a vertical 4-tap pass over a rotating register window.
//    64 x 128 output tile per CTA, warp = 8 output rows x 128 columns, lane = 4 adjacent columns.
//  * an odd-width OUTPUT row is bounced through a per-warp shared-memory row so that the scalar stores of a warp are 32
//    consecutive elements; aligned rows take one 16-byte (fp32) / 8-byte (16-bit) store per lane and row.
constexpr int kStripW = 128, kStripH = 64, kStripIH = kStripH + 3, kStripSlots = 132;

__device__ __forceinline__ void cp_async_4(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float half_lo_to_f32(uint32_t w, __nv_bfloat16) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float half_hi_to_f32(uint32_t w, __nv_bfloat16) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float half_lo_to_f32(uint32_t w, __half) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ float half_hi_to_f32(uint32_t w, __half) { return __high2float(*reinterpret_cast<const __half2*>(&w)); }

// columns 4*lane .. 4*lane + 7 of staged row `row` as fp32 (shift = 16 when the row was staged one slot to the right)
__device__ __forceinline__ void strip_window(const float* row, int lane, uint32_t, float (&w)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(row + 4 * lane), b = *reinterpret_cast<const float4*>(row + 4 * lane + 4);
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void strip_window(const T* row, int lane, uint32_t shift, float (&w)[8]) {
  const uint2 a = *reinterpret_cast<const uint2*>(row + 4 * lane), b = *reinterpret_cast<const uint2*>(row + 4 * lane + 4);
  const uint32_t w0 = __funnelshift_r(a.x, a.y, shift), w1 = __funnelshift_r(a.y, b.x, shift),
                 w2 = __funnelshift_r(b.x, b.y, shift), w3 = b.y >> shift;
  w[0] = half_lo_to_f32(w0, T()); w[1] = half_hi_to_f32(w0, T()); w[2] = half_lo_to_f32(w1, T()); w[3] = half_hi_to_f32(w1, T());
  w[4] = half_lo_to_f32(w2, T()); w[5] = half_hi_to_f32(w2, T()); w[6] = half_lo_to_f32(w3, T()); w[7] = half_hi_to_f32(w3, T());
}
__device__ __forceinline__ void store4(float* dst, const float (&o)[4]) {
  *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* dst, const float (&o)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
__device__ __forceinline__ void store4(__half* dst, const float (&o)[4]) {
  const __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

template <typename T, bool VEC_OUT>
__global__ void __launch_bounds__(256) upfirdn2d_blur_strip_kernel(const T* __restrict__ x, const float* __restrict__ k,
                                                                   T* __restrict__ y, UfdParams p, int tiles_x,
                                                                   float inv_tiles_x) {
  constexpr bool B16 = sizeof(T) == 2;
  __shared__ __align__(16) T sx[kStripIH][kStripSlots];
  __shared__ __align__(16) float sout[VEC_OUT ? 1 : 8][kStripW];
  __shared__ float sk[16], skr[4], skc[4];
  __shared__ int s_sep;
  const int64_t plane = blockIdx.y;
  const int ty_i = (int)(((float)blockIdx.x + 0.5f) * inv_tiles_x);
  const int oy0 = ty_i * kStripH, ox0 = ((int)blockIdx.x - ty_i * tiles_x) * kStripW;
  const int iy0 = oy0 - p.pad_y0, ix0 = ox0 - p.pad_x0;  // image coordinates of tile column 0 / row 0
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t plane_off = plane * p.in_h * (int64_t)p.in_w;
  const bool interior = iy0 >= 0 && iy0 + kStripIH <= p.in_h && ix0 >= 1 && ix0 + kStripW + 4 <= p.in_w;
  // 16-bit rows: slot s of staged row r holds tile column s - m_r, m_r = parity of the row's first element index
  const int m0 = B16 ? (int)((plane_off + (int64_t)iy0 * p.in_w + ix0) & 1) : 0;
  const int modd = B16 ? (p.in_w & 1) : 0;

  // ---- stage: all copies of the thread in flight, then one wait. Interior tiles (no bounds tests): one pointer per thread,
  // immediate offsets; border tiles: zero fill through the source size (the address of a zero-size copy is not read).
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(&sx[0][0]);
  if constexpr (!B16) {
    const T* src = x + plane_off + (int64_t)(iy0 + warp) * p.in_w + ix0 + lane;
    uint32_t dst = s_base + (uint32_t)((warp * kStripSlots + lane) * 4);
    if (interior) {
      for (int r = warp; r < kStripIH; r += 8, src += 8 * (int64_t)p.in_w, dst += 8 * kStripSlots * 4) {
#pragma unroll
        for (int c0 = 0; c0 < kStripW; c0 += 32) cp_async_4(dst + c0 * 4, src + c0, 4u);
        if (lane < 3) cp_async_4(dst + kStripW * 4, src + kStripW, 4u);
      }
    } else {
      const uint32_t span = (uint32_t)p.in_w;
      const int t = ix0 + lane;  // image column of tile column `lane`
      for (int r = warp; r < kStripIH; r += 8, src += 8 * (int64_t)p.in_w, dst += 8 * kStripSlots * 4) {
        const uint32_t row_bytes = (iy0 + r >= 0 && iy0 + r < p.in_h) ? 4u : 0u;
#pragma unroll
        for (int c0 = 0; c0 < kStripW; c0 += 32) cp_async_4(dst + c0 * 4, src + c0, (uint32_t)(t + c0) < span ? row_bytes : 0u);
        if (lane < 3) cp_async_4(dst + kStripW * 4, src + kStripW, (uint32_t)(t + kStripW) < span ? row_bytes : 0u);
      }
    }
  } else {
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
    int64_t e = plane_off + (int64_t)(iy0 + warp) * p.in_w + ix0;  // element index of tile column 0 of row r (< 0 possible at the border)
    uint32_t dst = s_base + (uint32_t)(warp * kStripSlots * 2 + lane * 4);
    if (interior) {
      for (int r = warp; r < kStripIH; r += 8, e += 8 * (int64_t)p.in_w, dst += 8 * kStripSlots * 2) {
        const uint32_t* src = xw + (e >> 1) + lane;   // aligned word that holds element e (floor: e may be odd)
        cp_async_4(dst, src, 4u);
        cp_async_4(dst + 128, src + 32, 4u);
        if (lane < 2) cp_async_4(dst + 256, src + 64, 4u);
      }
    } else {
      const uint32_t span = (uint32_t)p.in_w + 1u;      // a word is copied when at least one of its two columns is inside
      for (int r = warp; r < kStripIH; r += 8, e += 8 * (int64_t)p.in_w, dst += 8 * kStripSlots * 2) {
        const uint32_t* src = xw + (e >> 1) + lane;
        const uint32_t row_bytes = (iy0 + r >= 0 && iy0 + r < p.in_h) ? 4u : 0u;
        const int t = ix0 - (int)(e & 1) + 2 * lane + 1;  // image column of the word's high half
        cp_async_4(dst, src, (uint32_t)t < span ? row_bytes : 0u);
        cp_async_4(dst + 128, src + 32, (uint32_t)(t + 64) < span ? row_bytes : 0u);
        if (lane < 2) cp_async_4(dst + 256, src + 64, (uint32_t)(t + 128) < span ? row_bytes : 0u);
      }
    }
  }
  if (warp == 0) {
    // taps (flipped, upfirdn2d_kernel.cu:77) and their rank-one factors: lane i < 16 holds tap (i >> 2, i & 3)
    const int ky = lane >> 2, kx = lane & 3;
    float v = 0.f;
    if (lane < 16 && ky < p.kh && kx < p.kw) v = k[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
    float best = fabsf(v);
    int piv = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, piv, o);
      if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
    }
    piv &= 15;
    const float pv = __shfl_sync(0xffffffffu, v, piv);
    const float inv = pv != 0.f ? 1.0f / pv : 0.f;
    const float kc = __shfl_sync(0xffffffffu, v, (piv & 12) | kx);          // pivot row
    const float kr = __shfl_sync(0xffffffffu, v, (ky << 2) | (piv & 3)) * inv;  // pivot column / pivot
    const float dev = warp_max(lane < 16 ? fabsf(v - kr * kc) : 0.f);
    if (lane < 16) sk[lane] = v;
    if (lane < 4) skc[lane] = kc;           // lanes 0..3: ky = 0, kx = lane
    if (lane < 16 && kx == 0) skr[ky] = kr;
    if (lane == 0) s_sep = dev <= 1e-6f * fabsf(pv);
  }
  cp_async_wait_all();
  __syncthreads();
  if constexpr (B16) {
    if (!interior) {
      // a word that straddles the image border brought one element of the neighbouring row / plane: the padding is zero
      if (threadIdx.x < kStripIH) {
        const int r = threadIdx.x;
        const int m = m0 ^ (r & modd);
        const int s_lo = -1 - ix0 + m, s_hi = p.in_w - ix0 + m;  // slots of image columns -1 and in_w
        if (s_lo >= 0 && s_lo < kStripSlots) sx[r][s_lo] = from_f32<T>(0.f);
        if (s_hi >= 0 && s_hi < kStripSlots) sx[r][s_hi] = from_f32<T>(0.f);
      }
      __syncthreads();
    }
  }

  T* yp = y + plane * p.out_h * (int64_t)p.out_w;
  const int ox = ox0 + 4 * lane;
  auto store_row = [&](int a, const float (&o)[4]) {  // output row a (0..7) of this warp, columns ox .. ox + 3 of this lane
    const int oy = oy0 + warp * 8 + a;
    if constexpr (VEC_OUT) {
      if (oy < p.out_h && ox < p.out_w) store4(yp + (int64_t)oy * p.out_w + ox, o);  // out_w % 4 == 0
    } else {
      *reinterpret_cast<float4*>(&sout[warp][4 * lane]) = make_float4(o[0], o[1], o[2], o[3]);
      __syncwarp();
      if (oy < p.out_h) {
        T* dst = yp + (int64_t)oy * p.out_w + ox0 + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ox0 + 32 * j + lane < p.out_w) dst[32 * j] = from_f32<T>(sout[warp][32 * j + lane]);
      }
      __syncwarp();
    }
  };
  const int r0 = warp * 8;
  if (s_sep) {
    float kr[4], kc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      kr[i] = skr[i];
      kc[i] = skc[i];
    }
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll
    for (int r = 0; r < 11; ++r) {
      float w[8];
      strip_window(&sx[r0 + r][0], lane, (uint32_t)((m0 ^ ((r0 + r) & modd)) << 4), w);
      float h[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        h[b] = kc[0] * w[b];
#pragma unroll
        for (int kx = 1; kx < 4; ++kx) h[b] = fmaf(kc[kx], w[b + kx], h[b]);
      }
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int a = r - ky;  // input row r is tap ky of output row a
        if (a < 0 || a >= 8) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a & 3][b] = fmaf(kr[ky], h[b], acc[a & 3][b]);
      }
      if (r >= 3) {
        store_row(r - 3, acc[(r - 3) & 3]);
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[(r - 3) & 3][b] = 0.f;
      }
    }
  } else {
    for (int a = 0; a < 8; ++a) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      for (int ky = 0; ky < 4; ++ky) {
        float w[8];
        strip_window(&sx[r0 + a + ky][0], lane, (uint32_t)((m0 ^ ((r0 + a + ky) & modd)) << 4), w);
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
#pragma unroll
          for (int b = 0; b < 4; ++b) o[b] = fmaf(sk[ky * 4 + kx], w[b + kx], o[b]);
      }
      store_row(a, o);
    }
  }
}

template <typename T>
bool blur_strip_ok(const T* x, const UfdParams& p) {
  // 16-bit rows are staged as aligned 32-bit words: the base must be word aligned and the tensor must end on a word
  return sizeof(T) == 4 || (fmi_aligned(x, 4) && (p.major * p.in_h * (int64_t)p.in_w) % 2 == 0);
}

template <typename T>
int launch_blur_strip(const T* x, const float* k, T* y, const UfdParams& p, cudaStream_t st) {
  const int tiles_x = (p.out_w + kStripW - 1) / kStripW, tiles_y = (p.out_h + kStripH - 1) / kStripH;
  FMI_REQUIRE((int64_t)tiles_x * tiles_y < (1 << 23), "upfirdn2d: plane too large");
  const dim3 grid(tiles_x * tiles_y, (unsigned)p.major);
  const float inv_tx = 1.0f / (float)tiles_x;
  const bool vec_out = p.out_w % 4 == 0 && fmi_aligned(y, 4 * sizeof(T)) && ((int64_t)p.out_h * p.out_w) % 4 == 0;
  if (vec_out) upfirdn2d_blur_strip_kernel<T, true><<<grid, 256, 0, st>>>(x, k, y, p, tiles_x, inv_tx);
  else upfirdn2d_blur_strip_kernel<T, false><<<grid, 256, 0, st>>>(x, k, y, p, tiles_x, inv_tx);
  return FMI_OK;
}

template <typename T, int UP, int DOWN, int OT>
int launch_tile(const T* x, const float* k, T* y, const UfdParams& p, cudaStream_t st) {
  using F = FirTile<UP, DOWN, OT>;
  const int tiles_x = (p.out_w + F::TOW - 1) / F::TOW, tiles_y = (p.out_h + F::TOH - 1) / F::TOH;
  FMI_REQUIRE((int64_t)tiles_x * tiles_y < (1 << 23), "upfirdn2d: plane too large");
  const dim3 grid(tiles_x * tiles_y, (unsigned)p.major);
  const float inv_tx = 1.0f / (float)tiles_x;
  // CY = (o0*DOWN - pad0) mod UP for every tile origin o0 (a multiple of UP since the tile extent is)
  const int cy = ((-p.pad_y0) % UP + UP) % UP, cx = ((-p.pad_x0) % UP + UP) % UP;
#define FMI_UFD_CASE(CYV, CXV)                                                                                  \
  if (cy == CYV && cx == CXV)                                                                                   \
    upfirdn2d_tile_kernel<T, UP, DOWN, OT, CYV, CXV><<<grid, 256, 0, st>>>(x, k, y, p, tiles_x, inv_tx);
  FMI_UFD_CASE(0, 0)
  if constexpr (UP == 2) {
    FMI_UFD_CASE(0, 1)
    FMI_UFD_CASE(1, 0)
    FMI_UFD_CASE(1, 1)
  }
#undef FMI_UFD_CASE
  return FMI_OK;
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
  // live call sites: Blur fwd/bwd (1,1), Upsample of the RGB skip (up 2), its backward / Downsample (down 2)
  const bool tile = up_x == up_y && down_x == down_y && kh <= 4 && kw <= 4 && minor == 1 && major <= 65535 &&
                    ((up_x == 1 && down_x <= 2) || (up_x == 2 && down_x == 1));
  // FMI_UPFIRDN_STRIP=0 (read per call) restores the 16-tap tile kernel for the blur configurations
  const bool strip = tile && up_x == 1 && down_x == 1 && [] { const char* e = getenv("FMI_UPFIRDN_STRIP"); return !(e && e[0] == '0'); }();
  FMI_DISPATCH_DTYPE(dtype, T, {
    if (tile) {
      int rc;
      if (up_x == 2) rc = launch_tile<T, 2, 1, 4>((const T*)x, kernel, (T*)y, p, st);
      else if (down_x == 2) rc = launch_tile<T, 1, 2, 2>((const T*)x, kernel, (T*)y, p, st);
      else if (strip && blur_strip_ok<T>((const T*)x, p)) rc = launch_blur_strip<T>((const T*)x, kernel, (T*)y, p, st);
      else rc = launch_tile<T, 1, 1, 4>((const T*)x, kernel, (T*)y, p, st);
      if (rc) return rc;
    } else {
      int64_t total = major * p.out_h * (int64_t)p.out_w * minor;
      int grid = (int)imin64((total + 255) / 256, (int64_t)FMI_NUM_SMS * 32);
      upfirdn2d_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)x, kernel, (T*)y, p);
    }
  });
  return fmi_launched("upfirdn2d");
}
