// Implicit-GEMM convolution on tcgen05 shared by the modulated-conv forward (modconv.cu) and its data gradient
// (modconv_bwd.cu): kernel, tile planner and launcher. See modconv.cu for the formulation.
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace fmi_conv {
using namespace sm100;


constexpr int kGemmThreads = 320;   // TMA warp, MMA warp, two epilogue warpgroups (one per TMEM accumulator)
constexpr int kGemmStaticSmem = 10240;  // static shared memory of the kernel (barriers, 2 x 4 KB fused-ToRGB weights), rounded up
__device__ __align__(16) const float kZeroBias[32] = {};   // stands in for an absent bias vector in the epilogue
constexpr int kMaxTaps = 9;
constexpr int A_STAGE_BYTES = 128 * 128;
constexpr int A_HALO_BYTES = 17 * 1024;   // 130 pixel rows of 128 bytes, padded to the 1024-byte swizzle period

struct ConvGemmParams {
  int B, I, O, H, W;   // input NHWC
  int OH, OW;          // output NHWC extents
  int Mh, Mw;          // iteration domain of this launch (per image)
  int TH, TW, tiles_w; // pixel tile (TH*TW <= 128)
  int tiles_per_img, n_otiles, total_tiles;  // persistent tile loop: t -> (sample, output-channel tile, pixel tile)
  // halo mode (plain 3x3 convs on wide images): the pixel tile is 128 consecutive pixels of ONE image row; per kernel row
  // dy a single TMA box of 130 pixels (x0-1 .. x0+128, zero filled outside) serves the three horizontal taps — the A
  // descriptor of tap dx starts (dx+1) * 128 bytes into the SWIZZLE_128B tile (tools/umma_probe.cu tests 9-12: the MMA unit
  // swizzles on address bits, so a row-shifted start address reads the right rows). 3 instead of 9 activation loads.
  int halo;
  int halo_slab[3][3];  // weight slab of tap (dy-1, dx-1)
  int ntaps, tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_slab[kMaxTaps];
  int tap_boff[kMaxTaps];  // added to the batch coordinate of the A box (parity planes of the up-conv data gradient)
  int T;               // weight slabs per sample
  int sy, sx, py, px;  // output pixel = (m*sy + py, n*sx + px)
  int n_tile, k_chunks, stages, tmem_cols;
  int act;             // 0: raw accumulator, 1: noise + bias + leaky relu, 2: + bias, 3: tanh(+ bias)
  int w_shared;        // 1: one weight set [T][O][I] for the whole batch (plain convolutions, conv_blocks.cu)
  // output addressing in elements (0 = dense NHWC [B, OH, OW, O]): a channel slice of a wider buffer and / or the interior
  // of a padded one — the base pointer `out` is pre-offset by the caller
  int64_t out_bstride, out_rstride;
  int out_pstride;
  // merged parity classes of a stride-2 transposed conv (conv_blocks.cu): the N tile is 4 x merge_o columns, column block
  // cls = 2*py + px holds the merge_o output channels of output pixel (2m + py, 2n + px); taps are the 4 input shifts
  int merge_o;
  // several whole images per pixel tile (tiny planes: the map2style heads of the pSp encoder run 3x3 convs down to 1x1): the
  // A box is {channels, Mw, Mh, TB} — TB consecutive images, TB * Mh * Mw <= 128 rows — and the tile index walks image groups
  int TB;              // images per tile (0 / 1: one image, tiled by TH x TW)
  int w_group;         // per-sample weights: images [g * w_group, (g+1) * w_group) share weight set g (0 / 1: one set per image)
  int bias_classes;    // 9: bias is [9][O], indexed by the border class 3 * vy + vx of the output pixel (vy = 0 top row, 2 bottom
                       // row, 1 inside) — an input-side BatchNorm folded into a zero-padded 3x3 conv (ir_encoder.cu); else [O]
  int bias_set_stride; // per-sample weights: the bias of weight set g starts g * bias_set_stride floats into `bias` (0: one bias)
  const float* slope_c;  // act 4: PReLU, negative slope per output channel
  int prof_kind;       // FMI_PROF_* of this launch for the optional event timing (0: FMI_PROF_GEMM)
  int st256;           // output rows are 32-byte aligned: 256-bit stores / residual loads (set by the launcher)
  int add_out;         // the epilogue adds the values already stored at the output location (residual sum: y += conv(x))
  int raw_out;         // TF32 mode: store the fp32 accumulator as is instead of rounding it to tf32 (the consumer is not an MMA,
                       // or must see the exact value: InstanceNorm statistics amplify a rounding of x by |mean| / std)
  float* nchw_out;     // optional fp32 NCHW copy of the first nchw_C output channels ([B, nchw_C, OH, OW]); `out` may be NULL
  int nchw_C;
  const float* noise;
  int noise_batched;
  const float* noise_w;
  const float* bias;
  float slope, gain;
  void* out;
  // fused ToRGB (plain layers with a single N tile): rgb_w [B,3,O] modulated weights, rgb_out [B,3,OH,OW] fp32
  const float* rgb_w;
  const float* rgb_bias;
  const float* rgb_skip;
  const float* rgb_kf;
  float* rgb_out;
};

// Persistent: each CTA walks output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (pixel tile fastest, then output-
// channel tile, then sample). The TMA ring runs ahead across tile boundaries and the accumulator is double buffered in
// TMEM, so the epilogue of tile i overlaps the MMAs of tile i+1. (One tile per CTA left the tensor pipe 5 % active on the
// 32 -> 32 @1024^2 layer — ncu: 27 % warps active, the CTA lifetime was prologue + TMA latency + epilogue.)
template <bool TF32>
__global__ void __launch_bounds__(kGemmThreads, 2)
    modconv_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                        const ConvGemmParams p) {
  constexpr int EPA = TF32 ? 32 : 64;
  using OT = typename std::conditional<TF32, float, __nv_bfloat16>::type;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = p.n_tile * 128;
  const int a_bytes = p.halo ? A_HALO_BYTES : A_STAGE_BYTES;
  const int stage_bytes = a_bytes + (p.halo ? 3 : 1) * b_stage_bytes;
  __shared__ uint64_t full[8], empty[8], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float4 s_rgbw[2][256];  // fused ToRGB weights of the current image, per epilogue group: (w_r, w_g, w_b, -) per channel

  const int tid = threadIdx.x, warp = tid >> 5;
  const int iters = (p.halo ? 3 : p.ntaps) * p.k_chunks;
  const int per_img = p.tiles_per_img * p.n_otiles;
  const int TBi = p.TB > 1 ? p.TB : 1;
  const int wgrp = p.w_group > 1 ? p.w_group : 1;
  auto decode = [&](int t, int& b, int& o0, int& m0, int& n0) {
    b = t / per_img;
    const int rem = t - b * per_img;
    b *= TBi;
    const int oi = rem / p.tiles_per_img;
    const int tile = rem - oi * p.tiles_per_img;
    o0 = oi * p.n_tile;
    const int ty = tile / p.tiles_w;
    m0 = ty * p.TH;
    n0 = (tile - ty * p.tiles_w) * p.TW;
  };

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, p.tmem_cols);  // two accumulators; narrow layers leave TMEM for co-resident CTAs
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&map_x);
      tma_prefetch_desc(&map_w);
      const uint32_t bytes = (uint32_t)(p.TH * p.TW * TBi * 128 + b_stage_bytes);
      // ring position and parity are carried as counters: `g % stages`, `g / stages`, `it / k_chunks` with run-time
      // divisors cost ~30 instructions each on the single issuing thread of this warp
      int st = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int b, o0, m0, n0;
        decode(t, b, o0, m0, n0);
        const int wb = p.w_shared ? 0 : b / wgrp;
        int tap = 0, kc = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&empty[st], ph ^ 1);
          uint8_t* sA = smem + st * stage_bytes;
          const int tap_c = tap, kc_c = kc, st_c = st;
          if (++kc == p.k_chunks) { kc = 0; ++tap; }
          if (++st == p.stages) { st = 0; ph ^= 1; }
          if (p.halo) {  // `tap` is the kernel row here
            mbar_arrive_expect_tx(&full[st_c], (uint32_t)(130 * 128 + 3 * b_stage_bytes));
            tma_load_4d(sA, &map_x, &full[st_c], kc_c * EPA, n0 - 1, m0 + tap_c - 1, b);
            for (int dxi = 0; dxi < 3; ++dxi)
              tma_load_2d(sA + A_HALO_BYTES + dxi * b_stage_bytes, &map_w, &full[st_c], kc_c * EPA,
                          (wb * p.T + p.halo_slab[tap_c][dxi]) * p.O + o0);
            continue;
          }
          mbar_arrive_expect_tx(&full[st_c], bytes);
          tma_load_4d(sA, &map_x, &full[st_c], kc_c * EPA, n0 + p.tap_dx[tap_c], m0 + p.tap_dy[tap_c], b + p.tap_boff[tap_c]);
          tma_load_2d(sA + A_STAGE_BYTES, &map_w, &full[st_c], kc_c * EPA, (wb * p.T + p.tap_slab[tap_c]) * p.O + o0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, 128, p.n_tile);
      const int last_valid = p.I - (p.k_chunks - 1) * EPA;               // channels in the last chunk
      const int last_ksteps = (last_valid + EPA / 4 - 1) / (EPA / 4);     // MMA K = EPA / 4 elements
      // shared-memory descriptors advance by a constant per ring stage (the address field holds addr >> 4): one add per
      // stage instead of rebuilding them; ring position / parity / channel-chunk index are counters (no run-time div / mod)
      const uint64_t adesc0 = make_sdesc_k_sw128(smem_u32(smem));
      const uint64_t bdesc0 = make_sdesc_k_sw128(smem_u32(smem + A_STAGE_BYTES));
      const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
      int i = 0, st = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
        const int as = i & 1;
        mbar_wait(&acc_empty[as], ((i >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem + as * p.n_tile;
        int kc = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full[st], ph);
          tc_fence_after();
          uint8_t* sA = smem + st * stage_bytes;
          const bool last_chunk = kc == p.k_chunks - 1;
          const int st_c = st;
          if (++kc == p.k_chunks) kc = 0;
          if (++st == p.stages) { st = 0; ph ^= 1; }
          if (p.halo) {
            const int ksteps_h = last_chunk ? last_ksteps : 4;
            for (int dxi = 0; dxi < 3; ++dxi) {
              const uint64_t ad = make_sdesc_k_sw128(smem_u32(sA) + dxi * 128);  // window shifted by dxi pixel rows
              const uint64_t bd = make_sdesc_k_sw128(smem_u32(sA + A_HALO_BYTES + dxi * b_stage_bytes));
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                if (s >= ksteps_h) break;
                const uint32_t acc = (it > 0 || dxi > 0 || s > 0) ? 1u : 0u;
                if (TF32) mma_ss_tf32(d_tmem, ad + 2 * s, bd + 2 * s, idesc, acc);
                else mma_ss_f16(d_tmem, ad + 2 * s, bd + 2 * s, idesc, acc);
              }
            }
            tc_commit(&empty[st_c]);
            continue;
          }
          const uint64_t adesc = adesc0 + stage_step * (uint64_t)st_c;
          const uint64_t bdesc = bdesc0 + stage_step * (uint64_t)st_c;
          // the last channel chunk of a narrow layer is partly TMA zero fill (I = 32 bf16 fills half a 128-byte row):
          // skip the K steps that would only multiply zeros (each N <= 64 MMA costs ~85 clk whatever it multiplies)
          const int ksteps = last_chunk ? last_ksteps : 4;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            if (s >= ksteps) break;
            const uint32_t acc = (it > 0 || s > 0) ? 1u : 0u;
            if (TF32) mma_ss_tf32(d_tmem, adesc + 2 * s, bdesc + 2 * s, idesc, acc);
            else mma_ss_f16(d_tmem, adesc + 2 * s, bdesc + 2 * s, idesc, acc);
          }
          tc_commit(&empty[st_c]);
        }
        tc_commit(&acc_full[as]);
      }
    }
    __syncwarp();
  } else {
    // Two epilogue warpgroups: warps 2-5 drain accumulator 0 (even tiles of this CTA), warps 6-9 accumulator 1 (odd tiles), so
    // two tiles' epilogues run concurrently. (ncu: with one group the epilogue — TMEM load, bias / noise loads, activation,
    // conversion, stores: ~800 instructions and several load latencies per tile — set the tile rate of every narrow layer.)
    const int grp = (warp - 2) >> 2;
    const int gtid = tid - 64 - grp * 128;
    const int lane_base = (warp & 3) * 32;
    const int r = lane_base + (tid & 31);  // tile row == TMEM lane
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    int cur_b = -1, cur_o0 = -1, i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const int as = i & 1;
      if (as != grp) continue;
      int b, o0, m0, n0;
      decode(t, b, o0, m0, n0);
      const int ppi = p.TH * p.TW;                 // rows of one image in the tile
      const int bl = TBi > 1 ? r / ppi : 0, rr = r - bl * ppi;
      const int m = m0 + rr / p.TW, n = n0 + rr % p.TW;
      const bool valid = r < ppi * TBi && b + bl < p.B && m < p.Mh && n < p.Mw;
      b += bl;
      const int oy = m * p.sy + p.py, ox = n * p.sx + p.px;
      const float* biasp = p.bias;
      if (biasp && p.bias_set_stride) biasp += (int64_t)(b / wgrp) * p.bias_set_stride;
      if (p.bias_classes == 9 && biasp)
        biasp += ((oy == 0 ? 0 : (oy == p.OH - 1 ? 2 : 1)) * 3 + (ox == 0 ? 0 : (ox == p.OW - 1 ? 2 : 1))) * p.O;
      float nz = 0.f;
      if (p.act == 1 && valid && p.noise) {
        const float nw = p.noise_w ? *p.noise_w : 1.f;
        nz = nw * p.noise[(int64_t)(p.noise_batched ? b : 0) * p.OH * p.OW + (int64_t)oy * p.OW + ox];
      }
      OT* out = (OT*)p.out + (int64_t)b * p.out_bstride + (int64_t)oy * p.out_rstride + (int64_t)ox * p.out_pstride + o0;
      // fused ToRGB (model.py:360-369): the 1x1 modulated conv to 3 channels reads exactly the activations this thread
      // holds, so it is 3 dot products in the epilogue instead of a kernel that re-reads the whole layer output from HBM
      if (p.rgb_out && (b != cur_b || o0 != cur_o0)) {   // (never with TB > 1: b is the tile's image here)
        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");  // readers of the previous image's weights are done
        for (int e = gtid; e < p.n_tile; e += 128) {
          const float* w = p.rgb_w + ((int64_t)b * 3) * p.O + o0 + e;
          s_rgbw[grp][e] = make_float4(w[0], w[p.O], w[2 * p.O], 0.f);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
        cur_b = b;
        cur_o0 = o0;
      }
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      mbar_wait(&acc_full[as], (i >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + as * p.n_tile + c0, v);
        tc_wait_ld();
        if (c0 + 32 >= p.n_tile) {  // last read of this accumulator: hand it back to the MMA warp before the stores
          tc_fence_before();
          mbar_arrive(&acc_empty[as]);
        }
        if (!valid) continue;
        OT* outp = out + c0;
        int bofs = o0 + c0;
        float nzc = nz;
        if (p.merge_o) {
          const int cls = c0 / p.merge_o;
          bofs = c0 - cls * p.merge_o;
          outp = out + (cls >> 1) * p.out_rstride + (cls & 1) * p.out_pstride + bofs;
          if (p.act == 1 && p.noise && cls)   // the noise plane is indexed by the output pixel of this parity class
            nzc = (p.noise_w ? *p.noise_w : 1.f) * p.noise[(int64_t)(p.noise_batched ? b : 0) * p.OH * p.OW +
                                                          (int64_t)(oy + (cls >> 1)) * p.OW + ox + (cls & 1)];
        }
        float f[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
        // Bias (and PReLU slopes) of this 32-channel chunk as 8 independent 16-byte loads issued together. One scalar
        // `ptr ? __ldg(ptr + k) : 0` per element compiles to a branch per element, which serialises 32 dependent-latency loads
        // per chunk: ncu (source page, 32 -> 32 @512^2 bf16) showed 820 instructions and ~8.6 k cycles per tile in the epilogue
        // warps for ~1.5 k cycles of MMA — the epilogue, not the mainloop, bounded every implicit GEMM.
        if (p.act != 0) {
          // no bias: the same loads from a zero vector, so that the loop body has no branch (bias vectors are 16-byte aligned:
          // cudaMalloc'd parameter tensors at offsets that are multiples of 32 floats)
          const float4* b4 = reinterpret_cast<const float4*>(biasp ? biasp + bofs : kZeroBias);
          if (p.act == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t4 = __ldg(b4 + q);
              const float bq[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float tt = f[4 * q + j] + nzc + bq[j];
                f[4 * q + j] = (tt > 0.f ? tt : tt * p.slope) * p.gain;
              }
            }
          } else if (p.act == 4) {
            const float4* s4 = reinterpret_cast<const float4*>(p.slope_c + bofs);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t4 = __ldg(b4 + q), sl = __ldg(s4 + q);
              const float bq[4] = {t4.x, t4.y, t4.z, t4.w}, sv[4] = {sl.x, sl.y, sl.z, sl.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float tt = f[4 * q + j] + bq[j];
                f[4 * q + j] = tt > 0.f ? tt : tt * sv[j];
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t4 = __ldg(b4 + q);
              const float bq[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float tt = f[4 * q + j] + bq[j];
                f[4 * q + j] = p.act == 3 ? tanhf(tt) : tt;
              }
            }
          }
        }
        if (p.nchw_out) {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (o0 + c0 + k < p.nchw_C)
              p.nchw_out[(((int64_t)b * p.nchw_C + o0 + c0 + k) * p.OH + oy) * p.OW + ox] = f[k];
          if (!p.out) continue;
        }
        if (p.rgb_out) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float4 w = s_rgbw[grp][c0 + k];
            const float a = TF32 ? __uint_as_float(f32_to_tf32_rna(f[k])) : __bfloat162float(__float2bfloat16_rn(f[k]));
            rgb0 = fmaf(a, w.x, rgb0);
            rgb1 = fmaf(a, w.y, rgb1);
            rgb2 = fmaf(a, w.z, rgb2);
          }
        }
        if (!p.out) continue;   // fused ToRGB only (the last StyledConv of an inference forward: nothing else reads y)
        const int ncols = min(32, p.n_tile - c0);
        if (p.st256) {   // whole 32-byte sectors per lane (see st_global_v8)
          constexpr int EPV = TF32 ? 8 : 16;   // elements per 32 bytes
          if (p.add_out) {
#pragma unroll
            for (int k = 0; k < 32; k += EPV)
              if (k < ncols) {
                uint32_t pv[8];
                ld_global_v8(outp + k, pv);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  if constexpr (TF32) {
                    f[k + q] += __uint_as_float(pv[q]);
                  } else {
                    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&pv[q]);
                    f[k + 2 * q] += __low2float(h2);
                    f[k + 2 * q + 1] += __high2float(h2);
                  }
                }
              }
          }
#pragma unroll
          for (int k = 0; k < 32; k += EPV)
            if (k < ncols) {
              uint32_t u[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if constexpr (TF32) u[q] = p.raw_out ? __float_as_uint(f[k + q]) : f32_to_tf32_rna(f[k + q]);
                else u[q] = pack_bf16x2(f[k + 2 * q], f[k + 2 * q + 1]);
              }
              st_global_v8(outp + k, u);
            }
          continue;
        }
        if (p.add_out) {
          if constexpr (TF32) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              if (k < ncols) {
                const float4 prev = *reinterpret_cast<const float4*>(outp + k);
                f[k] += prev.x; f[k + 1] += prev.y; f[k + 2] += prev.z; f[k + 3] += prev.w;
              }
          } else {
#pragma unroll
            for (int k = 0; k < 32; k += 8)
              if (k < ncols) {
                const uint4 prev = *reinterpret_cast<const uint4*>(outp + k);
                const uint32_t pw[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&pw[q]);
                  f[k + 2 * q] += __low2float(h2);
                  f[k + 2 * q + 1] += __high2float(h2);
                }
              }
          }
        }
        if constexpr (TF32) {
#pragma unroll
          for (int k = 0; k < 32; k += 4)
            if (k < ncols)
              *reinterpret_cast<float4*>(outp + k) =
                  p.raw_out ? make_float4(f[k], f[k + 1], f[k + 2], f[k + 3])
                            : make_float4(__uint_as_float(f32_to_tf32_rna(f[k])), __uint_as_float(f32_to_tf32_rna(f[k + 1])),
                                          __uint_as_float(f32_to_tf32_rna(f[k + 2])), __uint_as_float(f32_to_tf32_rna(f[k + 3])));
        } else {
#pragma unroll
          for (int k = 0; k < 32; k += 8)
            if (k < ncols) {
              uint4 u;
              u.x = pack_bf16x2(f[k], f[k + 1]);
              u.y = pack_bf16x2(f[k + 2], f[k + 3]);
              u.z = pack_bf16x2(f[k + 4], f[k + 5]);
              u.w = pack_bf16x2(f[k + 6], f[k + 7]);
              *reinterpret_cast<uint4*>(outp + k) = u;
            }
        }
      }
      if (p.rgb_out && valid) {
        float r3[3] = {rgb0 + p.rgb_bias[0], rgb1 + p.rgb_bias[1], rgb2 + p.rgb_bias[2]};
        const int HW = p.OH * p.OW;
        if (p.rgb_skip) {
          // Upsample of the previous RGB: zero-insert x2, pad (2,1), flipped 4x4 taps (model.py:30-49). Tap ky reaches a real
          // sample iff oy + ky is even: ky in {py, py + 2} with py = oy & 1, source rows sy0 = (oy + py) / 2 - 1 and sy0 + 1 (same
          // in x): exactly 2 x 2 source pixels. Out-of-range ones are clamped and weighted 0, so the 4 tap loads and 12 pixel
          // loads are unconditional and issue together (same summation order as the tap loops).
          const int h2 = p.OH / 2, w2 = p.OW / 2;
          const int py = oy & 1, px = ox & 1;
          const int sy0 = ((oy + py) >> 1) - 1, sx0 = ((ox + px) >> 1) - 1;
          const float* sk = p.rgb_skip + (int64_t)b * 3 * h2 * w2;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int sy = sy0 + a, ky = py + 2 * a;
            const bool yok = sy >= 0 && sy < h2;
            const int syc = min(max(sy, 0), h2 - 1);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int sx = sx0 + c, kx = px + 2 * c;
              const bool ok = yok && sx >= 0 && sx < w2;
              const int sxc = min(max(sx, 0), w2 - 1);
              const float kv = ok ? __ldg(p.rgb_kf + 15 - (ky * 4 + kx)) : 0.f;
              const int64_t si = (int64_t)syc * w2 + sxc;
#pragma unroll
              for (int o = 0; o < 3; ++o) r3[o] = fmaf(kv, __ldg(sk + si + (int64_t)o * h2 * w2), r3[o]);
            }
          }
        }
#pragma unroll
        for (int o = 0; o < 3; ++o) p.rgb_out[((int64_t)b * 3 + o) * HW + (int64_t)oy * p.OW + ox] = r3[o];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
}

inline int esz_of(int mma) { return mma == FMI_MMA_TF32 ? 4 : 2; }

struct TilePlan { int TH, TW, tiles_h, tiles_w; };
inline TilePlan pick_tile(int Mh, int Mw) {
  TilePlan best{};
  int64_t best_area = -1;
  const int cands[5][2] = {{4, 32}, {8, 16}, {16, 8}, {32, 4}, {2, 64}};
  for (auto& c : cands) {
    int th = c[0], tw = c[1];
    if (tw > 256 || th > 256) continue;
    int tiles_h = (Mh + th - 1) / th, tiles_w = (Mw + tw - 1) / tw;
    int64_t area = (int64_t)tiles_h * tiles_w;
    if (best_area < 0 || area < best_area) {
      best_area = area;
      best = {th, tw, tiles_h, tiles_w};
    }
  }
  // small images: one tile covering the whole plane when it fits
  if ((int64_t)Mh * Mw <= 128 && Mw <= 128) best = {Mh, Mw, 1, 1};
  return best;
}

// Output-channel tile of a launch. The mainloop of a CTA runs at the SM's L2 -> shared-memory ingest (~45 B/clk measured: a
// 128 x 256 x 2304 tile = 36 stages of 48 KB takes ~26 us whatever the tensor pipe could do), so a launch that leaves SMs idle is
// faster with NARROWER tiles on more SMs, although every A tile is then fetched by several CTAs: per SM the cost is
// ceil(CTAs / 148) * (A bytes + n * 128) per K step. Candidates: n0, n0/2, ... >= 32 (divisors of O that are multiples of 32).
inline int pick_n_tile(int O, int n0, int64_t pixel_tiles) {
  static const bool off = [] { const char* e = getenv("FMI_GEMM_NTILE_AUTO"); return e && e[0] == '0'; }();
  if (off) return n0;
  int best = n0;
  int64_t best_cost = -1;
  for (int n = n0; n >= 32; n >>= 1) {
    if (O % n != 0 || n % 32 != 0) break;
    const int64_t ctas = pixel_tiles * (O / n);
    const int64_t cost = ((ctas + FMI_NUM_SMS - 1) / FMI_NUM_SMS) * (A_STAGE_BYTES + (int64_t)n * 128);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = n; }
  }
  return best;
}

template <bool TF32>
int launch_gemm_class(const CUtensorMap& mx, const CUtensorMap& mw, ConvGemmParams p, cudaStream_t st) {
  auto kern = modconv_gemm_kernel<TF32>;
  static FmiPerDeviceOnce attr_once;  // per translation unit and template instance
  if (attr_once.need()) {
    FMI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kGemmStaticSmem));  // static: 5 KB
    attr_once.done();
  }
  const int stage_bytes = p.halo ? A_HALO_BYTES + 3 * p.n_tile * 128 : A_STAGE_BYTES + p.n_tile * 128;
  // Two accumulators per CTA in TMEM. Narrow tiles (N <= 64: 128 columns) leave room for 3 co-resident CTAs, N <= 128 for
  // 2; the TMA ring is sized to the CTA's share of shared memory.
  p.tmem_cols = p.n_tile <= 16 ? 32 : p.n_tile <= 32 ? 64 : p.n_tile <= 64 ? 128 : p.n_tile <= 128 ? 256 : 512;
  // (a halo stage is 17 KB + three weight tiles and carries three taps: fewer, fatter CTAs)
  int ctas_per_sm = p.halo ? (p.n_tile <= 32 ? 2 : 1) : (p.n_tile <= 128 ? 2 : 1);   // 320 threads x <= 102 registers: 2 per SM
  {
    // a grid that does not fill the co-resident slots gets fewer CTAs per SM and a deeper TMA ring instead: a launch of <= 148
    // tiles is a chain of k-iterations per CTA, and with 3 stages of 32 KB in flight it runs at TMA latency, not bandwidth
    // (ncu, PICNet encoder convs at 32^2: 32 CTAs, 25-48 us for 5 us of MMA issue)
    const int tb = p.TB > 1 ? p.TB : 1;
    const TilePlan t0 = p.halo ? TilePlan{1, 128, p.Mh, (p.Mw + 127) / 128} : (tb > 1 ? TilePlan{p.Mh, p.Mw, 1, 1} : pick_tile(p.Mh, p.Mw));
    const int64_t tiles = (int64_t)t0.tiles_h * t0.tiles_w * (p.O / p.n_tile) * ((p.B + tb - 1) / tb);
    static const bool deep_off = [] { const char* e = getenv("FMI_GEMM_DEEP_RING"); return e && e[0] == '0'; }();
    while (!deep_off && ctas_per_sm > 1 && tiles <= (int64_t)FMI_NUM_SMS * (ctas_per_sm - 1)) --ctas_per_sm;
  }
  int stages = ((232448 - kGemmStaticSmem) / ctas_per_sm - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  FMI_REQUIRE(stages >= 2, "modconv_gemm: stage of %d bytes does not fit twice", stage_bytes);
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  TilePlan tp = pick_tile(p.Mh, p.Mw);
  if (p.halo) tp = TilePlan{1, 128, p.Mh, (p.Mw + 127) / 128};
  if (p.TB > 1) {
    FMI_REQUIRE(!p.halo && !p.rgb_out && !p.merge_o && p.Mh * p.Mw * p.TB <= 128 && (p.w_shared || p.w_group % p.TB == 0),
                "modconv_gemm: bad multi-image tile (TB=%d, %dx%d, w_group=%d)", p.TB, p.Mh, p.Mw, p.w_group);
    tp = TilePlan{p.Mh, p.Mw, 1, 1};
  }
  p.TH = tp.TH; p.TW = tp.TW; p.tiles_w = tp.tiles_w;
  p.tiles_per_img = tp.tiles_h * tp.tiles_w;
  p.n_otiles = p.O / p.n_tile;
  const int64_t total = (int64_t)p.tiles_per_img * p.n_otiles * ((p.B + (p.TB > 1 ? p.TB : 1) - 1) / (p.TB > 1 ? p.TB : 1));
  FMI_REQUIRE(total < (1ll << 31), "modconv_gemm: too many tiles");
  p.total_tiles = (int)total;
  if (p.out_pstride == 0) {
    p.out_pstride = p.O;
    p.out_rstride = (int64_t)p.OW * p.O;
    p.out_bstride = (int64_t)p.OH * p.OW * p.O;
  }
  {
    static const bool st256_off = [] { const char* e = getenv("FMI_GEMM_ST256"); return e && e[0] == '0'; }();
    const int64_t esz = TF32 ? 4 : 2;
    p.st256 = !st256_off && p.out && ((uintptr_t)p.out & 31) == 0 && (p.out_pstride * esz) % 32 == 0 && (p.out_rstride * esz) % 32 == 0 &&
              (p.out_bstride * esz) % 32 == 0 && p.n_tile % 16 == 0 && (!p.merge_o || p.merge_o % 16 == 0);
  }
  int grid = (int)imin64(total, (int64_t)FMI_NUM_SMS * ctas_per_sm);
  {  // debug: FMI_MODCONV_ONE_TILE=1 launches one CTA per tile (no tile loop) — must be bit-identical to the persistent run
    static const bool one_tile = [] { const char* e = getenv("FMI_MODCONV_ONE_TILE"); return e && e[0] == '1'; }();
    if (one_tile) grid = (int)total;
  }
  // algorithmic work of this launch: every input element read once, every output element written once, the weights once;
  // FLOPs = 2 * pixels * taps * I * N (merged parity classes: the zero weight rows are not counted)
  const double esz = TF32 ? 4.0 : 2.0;
  const double pix = (double)p.B * p.Mh * p.Mw;
  const double n_real = p.merge_o ? (double)p.merge_o * 9.0 / 4.0 : (double)p.O;   // merged: 9 real taps over 4 classes
  const double taps = p.halo ? 9.0 : (double)p.ntaps;
  const double wflops = 2.0 * pix * taps * p.I * (p.merge_o ? n_real * 4.0 / taps : n_real);
  const double out_elems = pix * (p.merge_o ? 4.0 * p.merge_o : (double)p.O);
  const double wbytes = ((double)p.B * p.H * p.W * p.I + (p.out ? out_elems * (p.add_out ? 2.0 : 1.0) : 0.0) +
                         (double)p.T * p.O * p.I * (p.w_shared ? 1.0 : (double)(p.B / (p.w_group > 1 ? p.w_group : 1)))) * esz +
                        (p.nchw_out ? pix * p.nchw_C * 4.0 : 0.0) + (p.rgb_out ? pix * 3 * 4.0 * (p.rgb_skip ? 1.25 : 1.0) : 0.0);
  FmiProfScope prof(p.prof_kind ? p.prof_kind : FMI_PROF_GEMM, st, wflops, wbytes);
  kern<<<grid, kGemmThreads, smem, st>>>(mx, mw, p);
  return fmi_launched("modconv_gemm");
}

}  // namespace fmi_conv
