// a3/a6: StyleGAN2 modulated convolution (ModulatedConv2d.forward, modules/psp/stylegan2/model.py:241-279) with the
// StyledConv / NoiseInjection / FusedLeakyReLU / ToRGB glue (:289-294, :340-346, :360-369) fused around it.
//
// Reformulation. The reference builds per-sample weights w[b] = scale*W*s[b] (*demod[b]) and runs a grouped conv
// (groups = batch) through cuDNN. Here the same per-sample weights are staged once per layer in the tensor-core
// operand type, tap-major and K-major ([b][tap][o][i], i contiguous), and the convolution is an implicit GEMM on
// tcgen05:  D[pixel, o] = sum_{tap, i} X[b, pixel + tap, i] * Wp[b][tap][o][i]
//   A operand: activations kept in NHWC between layers (channels contiguous = K-major); the im2col gather is a 4-D TMA
//              box {64 ch, TW, TH, 1} at the tap-shifted coordinate — out-of-bounds rows are zero-filled by TMA, which
//              is exactly the conv zero padding. One box = 128 rows x 128 bytes, SWIZZLE_128B.
//   B operand: Wp tile {64 ch, N_tile} by 2-D TMA.   Accumulator: 128 x N_tile fp32 in TMEM.
//   Epilogue (StyledConv): sqrt2 * lrelu_0.2(acc + noise_w * noise[b|0, p] + act_bias[o])  -> NHWC store.
// upsample=True (:255-263): conv_transpose2d(stride 2) is split into its 4 output-parity classes (1, 2, 2 and 4 taps:
//   9/4 taps per output pixel, the same FLOPs as the reference), each an implicit GEMM writing the (2H+1)^2 NHWC
//   intermediate; a second streaming kernel applies the 4x4 blur (pad 1,1) fused with noise + bias + leaky-ReLU.
// ToRGB (:360-369): 1x1 modulated conv to 3 channels without demodulation — HBM-bound, so a warp-shuffle SIMT kernel
//   that also fuses `+ bias` and `+ upfirdn2d(skip, up=2)` (Upsample, :30-49).
#include <stdlib.h>

#include "modconv_gemm.cuh"

using namespace sm100;
using namespace fmi_conv;

namespace {


// ---- s[b, i] = latent[b, :] . Wm[i, :] / sqrt(K) + bm[i]   (EqualLinear, model.py:159-167, lr_mul = 1) -----------
__global__ void __launch_bounds__(256) style_mod_kernel(const float* __restrict__ latent, int64_t ld,
                                                        const float* __restrict__ mw, const float* __restrict__ mb,
                                                        float* __restrict__ s, int B, int K, int I, float scale) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int idx = blockIdx.x * warps_per_block + (threadIdx.x >> 5); idx < B * I; idx += gridDim.x * warps_per_block) {
    const int b = idx / I, i = idx % I;
    float acc = 0.f;
    // the reference multiplies the weight by `scale` first (F.linear(input, weight * scale)): same rounding here
    for (int k = lane; k < K; k += 32) acc = fmaf(latent[(int64_t)b * ld + k], mw[(int64_t)i * K + k] * scale, acc);
    acc = warp_sum(acc);
    if (lane == 0) s[idx] = acc + (mb ? mb[i] : 0.f);
  }
}

// ---- per-sample modulated (and demodulated) weights in operand layout -------------------------------------------
// Wp[b][t][o][i] = OT( scale * W[o, i, t] * s[b, i] * demod[b, o] ),
// demod[b, o] = rsqrt( sum_{i,t} (scale * W[o,i,t] * s[b,i])^2 + 1e-8 )   (model.py:245-249). One block per (o, b).
template <typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ w, const float* __restrict__ s,
                                                          OT* __restrict__ wp, int B, int I, int O, int T,
                                                          float scale, int demodulate) {
  // One block per output channel o; W[o] (I*T floats, <= 18 KB) is staged ONCE in shared memory with coalesced loads and
  // pre-multiplied by `scale`, then every sample of the batch is produced from it (the first version re-read W[o] from
  // global memory with a stride-T gather once per sample: 59 us per 512x512 layer, 10 % of the decoder forward).
  extern __shared__ float sw[];  // [I*T] scale * W[o], original (i, t) order
  __shared__ float red[8];
  __shared__ float demod_s;
  const int o = blockIdx.x;
  const int n = I * T;
  const float* wo = w + (int64_t)o * n;
  for (int e = threadIdx.x; e < n; e += 256) sw[e] = scale * wo[e];
  __syncthreads();
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const float* sb = s + (int64_t)b * I;
    float d = 1.f;
    if (demodulate) {
      float acc = 0.f;
      for (int i = threadIdx.x; i < I; i += 256) {
        const float si = sb[i];
        float q = 0.f;
        for (int t = 0; t < T; ++t) {
          const float v = sw[i * T + t];
          q = fmaf(v, v, q);
        }
        acc = fmaf(q, si * si, acc);
      }
      acc = warp_sum(acc);
      __syncthreads();  // red / demod_s of the previous sample have been read
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x < 32) {
        float tt = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
        tt = warp_sum(tt);
        if (threadIdx.x == 0) demod_s = rsqrtf(tt + 1e-8f);
      }
      __syncthreads();
      d = demod_s;
    }
    if ((I & 7) == 0) {
      // 8 consecutive input channels per thread: one 16-byte (bf16) / two 16-byte (tf32) stores and one division per vector.
      // (The scalar version below wrote 2 bytes per lane: 50 us for a 512 x 512 layer whose 38 MB take 8 us at HBM speed.)
      const int vec_per_tap = I >> 3;
      for (int ev = threadIdx.x; ev < T * vec_per_tap; ev += 256) {
        const int t = ev / vec_per_tap, i0 = (ev - t * vec_per_tap) << 3;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = sw[(i0 + k) * T + t] * sb[i0 + k] * d;
          if (ROUND_TF32) v[k] = __uint_as_float(f32_to_tf32_rna(v[k]));
        }
        OT* dst = wp + (((int64_t)b * T + t) * O + o) * I + i0;
        if constexpr (sizeof(OT) == 2) {
          uint4 u;
          u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(dst) = u;
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      continue;
    }
    for (int e = threadIdx.x; e < n; e += 256) {
      const int t = e / I, i = e - t * I;  // write order: i fastest (coalesced); smem read is a stride-T gather (T = 9: no conflicts)
      float v = sw[i * T + t] * sb[i] * d;
      if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
      wp[(((int64_t)b * T + t) * O + o) * I + i] = from_f32<OT>(v);
    }
  }
}

// ---- layout changes at the module boundary -------------------------------------------------------------------------
template <typename TI, typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const TI* __restrict__ x, OT* __restrict__ y, int C, int HW,
                                                           int64_t x_bstride) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    t[j][tx] = (c < C && pp < HW) ? to_f32<TI>(x[(int64_t)b * x_bstride + (int64_t)c * HW + pp]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    if (pp < HW && c < C) {
      float v = t[tx][j];
      if (ROUND_TF32) v = __uint_as_float(f32_to_tf32_rna(v));
      y[((int64_t)b * HW + pp) * C + c] = from_f32<OT>(v);
    }
  }
}

template <typename OT, typename TO>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const OT* __restrict__ x, TO* __restrict__ y, int C, int HW) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int pp = p0 + j, c = c0 + tx;
    t[j][tx] = (c < C && pp < HW) ? to_f32<OT>(x[((int64_t)b * HW + pp) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, pp = p0 + tx;
    if (c < C && pp < HW) y[((int64_t)b * C + c) * HW + pp] = from_f32<TO>(t[tx][j]);
  }
}

// ---- 4x4 blur (pad 1,1) of the (2H+1)^2 NHWC intermediate fused with noise + bias + leaky relu --------------------
// out[b,y,x,c] = gain * lrelu( sum_{a,e} kf[a][e] * mid[b, y+a-1, x+e-1, c] + nw*noise[b|0,y,x] + bias[c] )
template <typename OT, int VEC>
__global__ void __launch_bounds__(256, 4) blur_act_nhwc_kernel(const OT* __restrict__ mid, OT* __restrict__ out,
                                                            const float* __restrict__ kf /*4x4 blur.kernel*/,
                                                            const float* __restrict__ noise, int noise_batched,
                                                            const float* __restrict__ noise_w,
                                                            const float* __restrict__ bias, int B, int C, int OH, int OW,
                                                            int act, float slope, float gain) {
  // Each thread owns a vertical strip of RY outputs of one (x, channel-vector) column and walks the RY+3 input rows once:
  // every loaded vector feeds up to 4 output rows from registers (5.5 loads per output instead of 16). Consecutive
  // threads take consecutive channel vectors, then consecutive x: every warp-level load is one contiguous 512-byte run.
  constexpr int RY = 4;
  const int MH = OH + 1, MW = OW + 1;
  const int cv = C / VEC;
  __shared__ float sk[16];
  if (threadIdx.x < 16) sk[threadIdx.x] = kf[15 - threadIdx.x];  // flipped taps (upfirdn2d_kernel.cu:77)
  __syncthreads();
  const float nw = noise_w ? *noise_w : 1.f;
  // grid (chunks of OW*cv, strips, B): one 32-bit division per thread. (A flat 64-bit index decode — three 64-bit
  // div/mod per output vector — made the ALU pipe the busiest unit of the first version: ncu 48 % ALU vs 27 % FMA.)
  const int col = blockIdx.x * blockDim.x + threadIdx.x;  // x * cv + channel vector
  if (col >= OW * cv) return;
  {
    const int x = col / cv;
    const int c = (col - x * cv) * VEC;
    const int y0 = blockIdx.y * RY;
    const int b = blockIdx.z;
    float acc[RY][VEC];
#pragma unroll
    for (int r = 0; r < RY; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
#pragma unroll
    for (int rr = 0; rr < RY + 3; ++rr) {
      const int yy = y0 + rr - 1;
      if (yy < 0 || yy >= MH) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int xx = x + e - 1;
        if (xx < 0 || xx >= MW) continue;
        const Vec16<OT> v = ld_vec16(mid + (((int64_t)b * MH + yy) * MW + xx) * C + c);
        float f[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) f[k] = to_f32<OT>(v.e[k]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int r = rr - a;  // output row y0 + r reads input row (y0 + r) + a - 1
          if (r < 0 || r >= RY) continue;
          const float kv = sk[a * 4 + e];
#pragma unroll
          for (int k = 0; k < VEC; ++k) acc[r][k] = fmaf(f[k], kv, acc[r][k]);
        }
      }
    }
    float bv[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) bv[k] = (act && bias) ? __ldg(bias + c + k) : 0.f;
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int y = y0 + r;
      if (y >= OH) break;
      Vec16<OT> o;
      if (act) {
        const float nz = noise ? nw * noise[(int64_t)(noise_batched ? b : 0) * OH * OW + (int64_t)y * OW + x] : 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float u = acc[r][k] + nz + bv[k];
          o.e[k] = from_f32<OT>((u > 0.f ? u : u * slope) * gain);
        }
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.e[k] = from_f32<OT>(acc[r][k]);
      }
      st_vec16(out + (((int64_t)b * OH + y) * OW + x) * C + c, o);
    }
  }
}

// ---- ToRGB: 1x1 modulated conv to 3 channels (no demod) + bias + upsampled skip -------------------------------------
// rgb[b,o,y,x] = sum_i (scale*W[o,i]*s[b,i]) * act[b,y,x,i] + bias[o] + upfirdn2d(skip, k, up=2, pad=(2,1))[b,o,y,x]
// LANES lanes cooperate on one pixel (16-byte channel vectors), reduced with warp shuffles.
template <typename OT, int VEC>
__global__ void __launch_bounds__(256) torgb_nhwc_kernel(const OT* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ s, const float* __restrict__ bias,
                                                         const float* __restrict__ skip, const float* __restrict__ kf,
                                                         float* __restrict__ rgb, int B, int I, int H, int W, float scale,
                                                         int lanes) {
  extern __shared__ float sw[];  // [3][I] modulated weights of this image, then 16 taps
  const int b = blockIdx.y;
  float* sk = sw + 3 * I;
  for (int e = threadIdx.x; e < 3 * I; e += blockDim.x) sw[e] = scale * w[e] * s[(int64_t)b * I + (e % I)];
  if (threadIdx.x < 16) sk[threadIdx.x] = kf ? kf[15 - threadIdx.x] : 0.f;  // flipped taps
  __syncthreads();
  const int sub = threadIdx.x % lanes;                 // lane within the pixel group
  const int groups_per_block = blockDim.x / lanes;
  const int HW = H * W;
  const int h2 = H / 2, w2 = W / 2;
  for (int pix = blockIdx.x * groups_per_block + threadIdx.x / lanes; pix < HW; pix += gridDim.x * groups_per_block) {
    const OT* xp = x + ((int64_t)b * HW + pix) * I;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int c = sub * VEC; c < I; c += lanes * VEC) {
      Vec16<OT> v = ld_vec16_stream(xp + c);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float f = to_f32<OT>(v.e[k]);
        a0 = fmaf(f, sw[c + k], a0);
        a1 = fmaf(f, sw[I + c + k], a1);
        a2 = fmaf(f, sw[2 * I + c + k], a2);
      }
    }
    for (int off = lanes >> 1; off > 0; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if (sub == 0) {
      const int y = pix / W, xx = pix % W;
      float r[3] = {a0 + bias[0], a1 + bias[1], a2 + bias[2]};
      if (skip) {
        // Upsample: zero-insert x2, pad (2,1), flipped 4x4 taps (model.py:30-49, upfirdn2d_kernel.cu:77)
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int uy = y + ky - 2;
          if (uy < 0 || (uy & 1) || (uy >> 1) >= h2) continue;
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            const int ux = xx + kx - 2;
            if (ux < 0 || (ux & 1) || (ux >> 1) >= w2) continue;
            const float kv = sk[ky * 4 + kx];
            const int64_t si = ((int64_t)b * 3 * h2 + (uy >> 1)) * w2 + (ux >> 1);
#pragma unroll
            for (int o = 0; o < 3; ++o) r[o] = fmaf(kv, skip[si + (int64_t)o * h2 * w2], r[o]);
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 3; ++o) rgb[((int64_t)b * 3 + o) * HW + pix] = r[o];
    }
  }
}


// ---- combined weights of "conv_transpose2d(stride 2, 3x3) then 4x4 blur (pad 1,1)" per output-parity class ------------
// out[2m+py, 2n+px] = sum_{dy,dx in {-1,0,1}} x[m+dy, n+dx] * Wc[py - 2dy][px - 2dx],
// Wc[u][v] = sum_{a,e} kfl[a][e] * w[u+a-1][v+e-1]  (w = the 3x3 per-sample weights, zero outside; kfl = flipped blur kernel,
// upfirdn2d_kernel.cu:77). wpc[b][shift = (dy+1)*3 + (dx+1)][cls*O + o][i], cls = 2*py + px. One block per (o, b).
template <typename OT, bool ROUND_TF32>
__global__ void __launch_bounds__(128) upblur_weight_kernel(const OT* __restrict__ wp /*[B][9][O][I]*/,
                                                            const float* __restrict__ kf /*4x4 blur.kernel*/,
                                                            OT* __restrict__ wpc /*[B][9][4*O][I]*/, int O, int I) {
  __shared__ float skf[16];
  if (threadIdx.x < 16) skf[threadIdx.x] = kf[15 - threadIdx.x];  // kfl[a][e] = kf[3-a][3-e]
  __syncthreads();
  const int o = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    float w[3][3];
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t / 3][t % 3] = to_f32<OT>(wp[(((int64_t)b * 9 + t) * O + o) * I + i]);
#pragma unroll
    for (int sh = 0; sh < 9; ++sh) {
      const int dy = sh / 3 - 1, dx = sh % 3 - 1;
#pragma unroll
      for (int cls = 0; cls < 4; ++cls) {
        const int u = (cls >> 1) - 2 * dy, v = (cls & 1) - 2 * dx;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int ky = u + a - 1;
          if (ky < 0 || ky > 2) continue;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int kx = v + e - 1;
            if (kx < 0 || kx > 2) continue;
            acc = fmaf(skf[a * 4 + e], w[ky][kx], acc);
          }
        }
        if (ROUND_TF32) acc = __uint_as_float(f32_to_tf32_rna(acc));
        wpc[(((int64_t)b * 9 + sh) * (4 * O) + cls * O + o) * I + i] = from_f32<OT>(acc);
      }
    }
  }
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" int fmi_nchw_to_nhwc(const void* x, void* y, int B, int C, int H, int W, int dtype, int mma, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype) && (mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16), "nchw_to_nhwc: bad dtype/mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && y && C >= 1 && H >= 1 && W >= 1, "nchw_to_nhwc: bad arguments");
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t bs = (int64_t)C * HW;
  FMI_DISPATCH_DTYPE(dtype, T, {
    if (mma == FMI_MMA_TF32) nchw_to_nhwc_kernel<T, float, true><<<grid, 256, 0, st>>>((const T*)x, (float*)y, C, HW, bs);
    else nchw_to_nhwc_kernel<T, __nv_bfloat16, false><<<grid, 256, 0, st>>>((const T*)x, (__nv_bfloat16*)y, C, HW, bs);
  });
  return fmi_launched("nchw_to_nhwc");
}

extern "C" int fmi_nhwc_to_nchw(const void* x, void* y, int B, int C, int H, int W, int mma, int dtype, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype) && (mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16), "nhwc_to_nchw: bad dtype/mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && y && C >= 1 && H >= 1 && W >= 1, "nhwc_to_nchw: bad arguments");
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  cudaStream_t st = (cudaStream_t)stream;
  FMI_DISPATCH_DTYPE(dtype, T, {
    if (mma == FMI_MMA_TF32) nhwc_to_nchw_kernel<float, T><<<grid, 256, 0, st>>>((const float*)x, (T*)y, C, HW);
    else nhwc_to_nchw_kernel<__nv_bfloat16, T><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (T*)y, C, HW);
  });
  return fmi_launched("nhwc_to_nchw");
}

extern "C" int fmi_style_modulation(const float* latent, int64_t latent_row_stride, const float* mod_weight,
                                    const float* mod_bias, float* s, int B, int K, int I, void* stream) {
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(latent && mod_weight && s && K >= 1 && I >= 1, "style_modulation: bad arguments");
  int grid = (B * I + 7) / 8;
  if (grid > FMI_NUM_SMS * 8) grid = FMI_NUM_SMS * 8;
  style_mod_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(latent, latent_row_stride, mod_weight, mod_bias, s, B, K, I,
                                                            1.0f / sqrtf((float)K));
  return fmi_launched("style_modulation");
}

extern "C" int64_t fmi_modconv_weight_bytes(int B, int I, int O, int ksize, int mma) {
  return (int64_t)B * ksize * ksize * O * I * esz_of(mma);
}

extern "C" int fmi_modconv_weight_prep(const float* weight, const float* s, void* wp, int B, int I, int O, int ksize,
                                       int demodulate, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "modconv_weight_prep: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(weight && s && wp && I >= 1 && O >= 1 && (ksize == 1 || ksize == 3), "modconv_weight_prep: bad arguments");
  const int T = ksize * ksize;
  const float scale = 1.0f / sqrtf((float)(I * T));  // model.py:225-226
  // blocks: O x (batch slices); each block stages W[o] once and loops over its samples
  int by = B;
  while (by > 1 && (int64_t)O * by > (int64_t)FMI_NUM_SMS * 8) by = (by + 1) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)I * T * sizeof(float);
  FMI_REQUIRE(smem <= 48 * 1024, "modconv_weight_prep: I*k*k = %d floats exceed the staging buffer", I * T);
  if (mma == FMI_MMA_TF32)
    weight_prep_kernel<float, true><<<dim3(O, by, 1), 256, smem, st>>>(weight, s, (float*)wp, B, I, O, T, scale, demodulate);
  else
    weight_prep_kernel<__nv_bfloat16, false><<<dim3(O, by, 1), 256, smem, st>>>(weight, s, (__nv_bfloat16*)wp, B, I, O, T, scale, demodulate);
  return fmi_launched("modconv_weight_prep");
}

extern "C" int64_t fmi_styled_conv_workspace_bytes(int B, int O, int H, int W, int upsample, int mma) {
  if (!upsample) return 0;
  return (int64_t)B * (2 * H + 1) * (2 * W + 1) * O * esz_of(mma);
}

namespace {
struct RgbFuse {  // fused ToRGB of a plain StyledConv (fmi_styled_conv_torgb_nhwc)
  const float* w;     // [B,3,O] modulated weights (fmi_torgb_weights)
  const float* bias;  // [3]
  const float* skip;  // [B,3,H/2,W/2] or NULL
  const float* kf;    // 4x4 upsample kernel (needed with skip)
  float* out;         // [B,3,H,W]
};
int styled_conv_impl(const void* x, const void* wp, void* y, const float* noise, int noise_batched, const float* noise_w,
                     const float* act_bias, const float* blur_k, int B, int I, int O, int H, int W, int upsample, int act,
                     int mma, void* workspace, int64_t workspace_bytes, void* stream, const RgbFuse* rgb);
}  // namespace

extern "C" int fmi_styled_conv_nhwc(const void* x, const void* wp, void* y, const float* noise, int noise_batched,
                                    const float* noise_w, const float* act_bias, const float* blur_k, int B, int I, int O,
                                    int H, int W, int upsample, int act, int mma, void* workspace, int64_t workspace_bytes,
                                    void* stream) {
  return styled_conv_impl(x, wp, y, noise, noise_batched, noise_w, act_bias, blur_k, B, I, O, H, W, upsample, act, mma,
                          workspace, workspace_bytes, stream, nullptr);
}

// rgb_w[b][o][c] = W[o,c] * s[b,c] / sqrt(C)   (ToRGB's modulated, not demodulated, 1x1 weights; model.py:357,245)
__global__ void __launch_bounds__(256) torgb_weights_kernel(const float* __restrict__ w, const float* __restrict__ s,
                                                            float* __restrict__ out, int B, int C, float scale) {
  const int total = B * 3 * C;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e % C, o = (e / C) % 3, b = e / (3 * C);
    out[e] = scale * w[o * C + c] * s[b * C + c];
  }
}

extern "C" int fmi_torgb_weights(const float* weight, const float* s, float* rgb_w, int B, int C, void* stream) {
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(weight && s && rgb_w && C >= 1, "torgb_weights: bad arguments");
  torgb_weights_kernel<<<(B * 3 * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(weight, s, rgb_w, B, C,
                                                                                   1.0f / sqrtf((float)C));
  return fmi_launched("torgb_weights");
}

// A plain (non-upsampling) StyledConv with the following ToRGB fused into its epilogue (O <= 256):
//   y as fmi_styled_conv_nhwc(act = 1); rgb[b,o,p] = sum_c rgb_w[b,o,c] * y[b,p,c] + rgb_bias[o] + upsample(skip)[b,o,p].
extern "C" int fmi_styled_conv_torgb_nhwc(const void* x, const void* wp, void* y, const float* noise, int noise_batched,
                                          const float* noise_w, const float* act_bias, const float* rgb_w,
                                          const float* rgb_bias, const float* rgb_skip, const float* rgb_kernel, float* rgb,
                                          int B, int I, int O, int H, int W, int mma, void* stream) {
  FMI_REQUIRE(rgb_w && rgb_bias && rgb, "styled_conv_torgb: null pointer");
  FMI_REQUIRE(O <= 256, "styled_conv_torgb: fused ToRGB needs O <= 256 (got %d); use fmi_torgb_nhwc", O);
  FMI_REQUIRE(!rgb_skip || (rgb_kernel && H % 2 == 0 && W % 2 == 0), "styled_conv_torgb: skip needs the kernel and even H, W");
  RgbFuse f{rgb_w, rgb_bias, rgb_skip, rgb_kernel, rgb};
  return styled_conv_impl(x, wp, y, noise, noise_batched, noise_w, act_bias, nullptr, B, I, O, H, W, 0, 1, mma, nullptr, 0,
                          stream, &f);
}

namespace {
int styled_conv_impl(const void* x, const void* wp, void* y, const float* noise, int noise_batched, const float* noise_w,
                     const float* act_bias, const float* blur_k, int B, int I, int O, int H, int W, int upsample, int act,
                     int mma, void* workspace, int64_t workspace_bytes, void* stream, const RgbFuse* rgb) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "styled_conv: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && wp && (y || (rgb && !upsample)), "styled_conv: null pointer");   // y may be NULL when only the fused ToRGB is wanted
  FMI_REQUIRE(I >= 16 && O >= 32 && O % 32 == 0 && H >= 1 && W >= 1,
              "styled_conv: unsupported shape I=%d O=%d H=%d W=%d (O must be a multiple of 32)", I, O, H, W);
  const int esz = esz_of(mma);
  FMI_REQUIRE((I * esz) % 16 == 0 && (O * esz) % 16 == 0, "styled_conv: channel counts must give 16-byte rows");
  FMI_REQUIRE(O <= 256 || O % 256 == 0, "styled_conv: O=%d must be <= 256 or a multiple of 256", O);
  FMI_REQUIRE(fmi_aligned(x, 16) && fmi_aligned(wp, 16) && (!y || fmi_aligned(y, 16)), "styled_conv: buffers must be 16-byte aligned");
  int rc = fmi_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tf32 = mma == FMI_MMA_TF32;
  const uint32_t epa = 128 / esz;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int T = 9;

  ConvGemmParams p{};
  p.B = B; p.I = I; p.O = O; p.H = H; p.W = W; p.T = T;
  p.n_tile = O <= 256 ? O : 256;
  if (!rgb) {   // (the fused ToRGB needs all O channels of a pixel in one tile) few pixel tiles: narrower tiles on more SMs
    const TilePlan t0 = pick_tile(H, W);
    p.n_tile = pick_n_tile(O, p.n_tile, (int64_t)B * t0.tiles_h * t0.tiles_w);
  }
  p.k_chunks = (I + epa - 1) / epa;
  p.noise = noise; p.noise_batched = noise_batched; p.noise_w = noise_w; p.bias = act_bias;
  p.slope = 0.2f; p.gain = 1.4142135623730951f;

  CUtensorMap mw;
  {
    uint64_t dims[2] = {(uint64_t)I, (uint64_t)B * T * O};
    uint64_t str[1] = {(uint64_t)I * esz};
    uint32_t box[2] = {epa, (uint32_t)p.n_tile};
    int e = make_tensor_map(&mw, dt, 2, wp, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "styled_conv: cuTensorMapEncodeTiled(weights) failed (%d)", e);
  }
  auto make_x_map = [&](CUtensorMap* m, int th, int tw) -> int {
    uint64_t dims[4] = {(uint64_t)I, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)I * esz, (uint64_t)W * I * esz, (uint64_t)H * W * I * esz};
    uint32_t box[4] = {epa, (uint32_t)tw, (uint32_t)th, 1};
    return make_tensor_map(m, dt, 4, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };

  if (!upsample) {
    p.OH = H; p.OW = W; p.Mh = H; p.Mw = W; p.sy = p.sx = 1; p.py = p.px = 0;
    p.ntaps = 9;
    for (int t = 0; t < 9; ++t) { p.tap_dy[t] = t / 3 - 1; p.tap_dx[t] = t % 3 - 1; p.tap_slab[t] = t; }
    p.act = act; p.out = y;
    if (rgb) {
      p.rgb_w = rgb->w; p.rgb_bias = rgb->bias; p.rgb_skip = rgb->skip; p.rgb_kf = rgb->kf; p.rgb_out = rgb->out;
    }
    TilePlan tp = pick_tile(p.Mh, p.Mw);
    // halo mode: wide, narrow-channel layers re-read every activation once per tap from L2 (ncu: 10 GB of L2 reads for 0.6 GB
    // of DRAM reads on 32 -> 32 @1024^2); one 130-pixel box per kernel row cuts that 3x. On by default for N <= 32 (round 2,
    // after the epilogue stopped bounding the kernel: 32 -> 32 @1024^2 + ToRGB 866 -> 656 us; N = 64 layers get slower with it:
    // 316 -> 417 us, one CTA per SM). FMI_MODCONV_HALO=1 forces it for every N <= 128, =0 switches it off.
    const int halo_env = [] { const char* e = getenv("FMI_MODCONV_HALO"); return e ? (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : -1)) : -1; }();
    p.halo = W >= 128 && p.n_tile <= 128 && (halo_env == 1 || (halo_env == -1 && p.n_tile <= 32));
    if (p.halo) {
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) p.halo_slab[a][c] = a * 3 + c;
      tp = TilePlan{1, 130, 0, 0};  // box: 130 pixels of one row
    }
    CUtensorMap mx;
    int e = make_x_map(&mx, tp.TH, tp.TW);
    FMI_REQUIRE(e == 0, "styled_conv: cuTensorMapEncodeTiled(x) failed (%d)", e);
    return tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
  }
  FMI_REQUIRE(!rgb, "styled_conv: ToRGB fusion is for plain (non-upsampling) layers");
  FMI_REQUIRE(blur_k, "styled_conv: upsample needs the 4x4 blur kernel");

  // ---- upsample, narrow layers (O <= 64): conv_transpose2d(stride 2, 3x3) followed by the 4x4 blur IS one transposed conv
  // with the 6x6 kernel Wc = w (*) blur, i.e. per output-parity class a 3x3 conv on the input (i = m-1..m+1). All four
  // classes read the same nine input shifts, so the layer is ONE implicit GEMM with N = 4*O columns (merged parity classes,
  // modconv_gemm.cuh) whose epilogue adds noise + bias + leaky-ReLU: the same number of MMA instructions as the four
  // per-class launches below (9 taps, now 4x wider), no (2H+1)^2 intermediate and no blur pass (0.6 ms of a batch-8
  // 1024^2 layer). The combined per-sample weights are built from wp in the workspace. FMI_UPBLUR_FUSED=0: the two-pass path.
  {
    const bool off = [] { const char* e = getenv("FMI_UPBLUR_FUSED"); return e && e[0] == '0'; }();  // read per call
    const int64_t need_w = (int64_t)B * 9 * 4 * O * I * esz;
    if (!off && O <= 64 && workspace && workspace_bytes >= need_w) {
      if (tf32)
        upblur_weight_kernel<float, true><<<dim3(O, B), 128, 0, st>>>((const float*)wp, blur_k, (float*)workspace, O, I);
      else
        upblur_weight_kernel<__nv_bfloat16, false><<<dim3(O, B), 128, 0, st>>>((const __nv_bfloat16*)wp, blur_k,
                                                                              (__nv_bfloat16*)workspace, O, I);
      rc = fmi_launched("upblur_weight");
      if (rc) return rc;
      CUtensorMap mwc;
      {
        uint64_t dims[2] = {(uint64_t)I, (uint64_t)B * 9 * 4 * O};
        uint64_t str[1] = {(uint64_t)I * esz};
        uint32_t box[2] = {epa, (uint32_t)(4 * O)};
        int e = make_tensor_map(&mwc, dt, 2, workspace, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        FMI_REQUIRE(e == 0, "styled_conv: cuTensorMapEncodeTiled(combined weights) failed (%d)", e);
      }
      p.merge_o = O; p.O = 4 * O; p.n_tile = 4 * O; p.T = 9;
      p.OH = 2 * H; p.OW = 2 * W; p.sy = p.sx = 2; p.py = p.px = 0; p.Mh = H; p.Mw = W;
      p.out_pstride = O; p.out_rstride = (int64_t)p.OW * O; p.out_bstride = (int64_t)p.OH * p.OW * O;
      p.act = act; p.out = y;
      p.ntaps = 9;
      for (int t = 0; t < 9; ++t) { p.tap_dy[t] = t / 3 - 1; p.tap_dx[t] = t % 3 - 1; p.tap_slab[t] = t; }
      TilePlan tp = pick_tile(p.Mh, p.Mw);
      CUtensorMap mx;
      int e = make_x_map(&mx, tp.TH, tp.TW);
      FMI_REQUIRE(e == 0, "styled_conv: cuTensorMapEncodeTiled(x) failed (%d)", e);
      return tf32 ? launch_gemm_class<true>(mx, mwc, p, st) : launch_gemm_class<false>(mx, mwc, p, st);
    }
  }

  // ---- upsample: conv_transpose2d(stride 2) by output parity class, then blur + epilogue
  const int MH = 2 * H + 1, MW = 2 * W + 1;
  const int64_t need = (int64_t)B * MH * MW * O * esz;
  FMI_REQUIRE(workspace && workspace_bytes >= need, "styled_conv: workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)need);
  p.OH = MH; p.OW = MW; p.sy = p.sx = 2; p.act = 0; p.out = workspace;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      // out[2m+py, 2n+px]: ky in {0,2} (py==0) or {1}; input row m - ky/2
      p.py = py; p.px = px;
      p.Mh = py == 0 ? H + 1 : H;
      p.Mw = px == 0 ? W + 1 : W;
      int nt = 0;
      for (int ky = py; ky < 3; ky += 2)
        for (int kx = px; kx < 3; kx += 2) {
          p.tap_dy[nt] = -(ky / 2);
          p.tap_dx[nt] = -(kx / 2);
          p.tap_slab[nt] = ky * 3 + kx;
          ++nt;
        }
      p.ntaps = nt;
      TilePlan tp = pick_tile(p.Mh, p.Mw);
      CUtensorMap mx;
      int e = make_x_map(&mx, tp.TH, tp.TW);
      FMI_REQUIRE(e == 0, "styled_conv: cuTensorMapEncodeTiled(x) failed (%d)", e);
      rc = tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
      if (rc) return rc;
    }
  // blur: flipped taps (upfirdn2d_kernel.cu:77)
  const int OH = 2 * H, OW = 2 * W;
  const int cvn = O * esz / 16;                                // 16-byte channel vectors per pixel
  const dim3 grid((OW * cvn + 255) / 256, (OH + 3) / 4, B);    // one thread per (x, channel vector, 4-row strip)
  FMI_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "styled_conv: blur grid too large");
  if (tf32)
    blur_act_nhwc_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)workspace, (float*)y, blur_k, noise, noise_batched,
                                                         noise_w, act_bias, B, O, OH, OW, act, p.slope, p.gain);
  else
    blur_act_nhwc_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)workspace, (__nv_bfloat16*)y, blur_k,
                                                                 noise, noise_batched, noise_w, act_bias, B, O, OH, OW,
                                                                 act, p.slope, p.gain);
  return fmi_launched("blur_act");
}
}  // namespace

extern "C" int fmi_torgb_nhwc(const void* x, const float* weight, const float* s, const float* bias, const float* skip,
                              const float* blur_k, float* rgb, int B, int I, int H, int W, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "torgb: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && weight && s && bias && rgb, "torgb: null pointer");
  FMI_REQUIRE(!skip || (blur_k && H % 2 == 0 && W % 2 == 0), "torgb: skip needs the blur kernel and even H, W");
  const int esz = esz_of(mma);
  const int vec = 16 / esz;
  FMI_REQUIRE(I % vec == 0, "torgb: I=%d must be a multiple of %d", I, vec);
  int lanes = I / vec;  // lanes per pixel, power of two <= 32
  if (lanes > 32) lanes = 32;
  int l2 = 1;
  while (l2 * 2 <= lanes) l2 *= 2;
  lanes = l2;
  const int groups = 256 / lanes;
  int gx = (H * W + groups - 1) / groups;
  if (gx > FMI_NUM_SMS * 8) gx = FMI_NUM_SMS * 8;
  dim3 grid(gx, B);
  const size_t smem = (size_t)(3 * I + 16) * sizeof(float);
  const float scale = 1.0f / sqrtf((float)I);
  cudaStream_t st = (cudaStream_t)stream;
  if (mma == FMI_MMA_TF32)
    torgb_nhwc_kernel<float, 4><<<grid, 256, smem, st>>>((const float*)x, weight, s, bias, skip, blur_k, rgb, B, I, H, W, scale, lanes);
  else
    torgb_nhwc_kernel<__nv_bfloat16, 8><<<grid, 256, smem, st>>>((const __nv_bfloat16*)x, weight, s, bias, skip, blur_k, rgb, B, I, H, W, scale, lanes);
  return fmi_launched("torgb");
}
