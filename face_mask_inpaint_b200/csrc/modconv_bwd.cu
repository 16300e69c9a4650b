// a3/a6 backward: gradients of the fused StyledConv (modulated conv + noise + bias + leaky-ReLU, modules/psp/stylegan2/
// model.py:241-279, :289-294, :340-346) and of ToRGB (:360-369), needed by train_psp.py (train_decoder).
//
// Forward (modconv.cu):  y = gain * lrelu( conv(x, Wp[b]) + nw * noise + bias ),  Wp[b][t][o][i] = scale W[o,i,t] s[b,i] d[b,o],
//                        d[b,o] = rsqrt( sum_{i,t} (scale W s)^2 + 1e-8 );  up-sampling layers: convT(stride 2) -> 4x4 blur.
// Backward, all in the NHWC operand layout of the forward:
//   1. act_bwd_nhwc_kernel       g = dy * (y > 0 ? gain : gain*slope);  dbias[o] = sum g;  dnoise_w = sum g * noise
//   2. blur_bwd_planes_kernel    (up only) gmid = blur^T g, written as the 4 output-parity planes of the (2H+1)^2 grid
//                                P[py][px][b][m][n][o] = gmid[b, 2m+py, 2n+px, o]  (zero where that falls outside)
//   3. data gradient             plain: dx[p,i] = sum_{t,o} g[p - off_t, o] WpT[t][i][o]      -> the forward's implicit-GEMM
//                                up:    dx[m,n,i] = sum_{t,o} P[par_t][m+ky/2][n+kx/2][o] WpT   kernel (modconv_gemm.cuh) with
//                                transposed weights WpT[b][t][i][o] (transpose_wp_kernel)
//   4. wgrad_gemm_kernel         G[b][t][o][i] = dL/dWp = sum_p g[p,o] x[p+off_t, i]  (up: sum_{m,n} P[..][o] x[m,n,i]):
//                                tcgen05 GEMM whose contraction index is the PIXEL, so both operands are MN-major tiles —
//                                exactly the NHWC box {64 ch, pixels} TMA already delivers. Split-K over pixel ranges,
//                                fp32 accumulators in TMEM (up to three taps per CTA share the un-shifted operand),
//                                reduced into G with 16-byte red.global.add.
//   5. weight_bwd_kernel         through the modulation/demodulation:  du = d (G - Wp c), c[b,o] = sum_{i,t} G Wp;
//                                dW[o,i,t] = scale sum_b s[b,i] du ;  ds[b,i] = scale sum_{o,t} W[o,i,t] du
//   ToRGB: torgb_bwd_nhwc_kernel  dx[p,c] = sum_o rgb_w[b,o,c] drgb[b,o,p];  d rgb_w[b,o,c] = sum_p drgb[b,o,p] x[p,c]
#include <stdlib.h>

#include "modconv_gemm.cuh"

using namespace sm100;
using namespace fmi_conv;

namespace {

// ---- 1. activation backward + bias / noise-weight gradients -----------------------------------------------------------
// Thread = one 16-byte channel vector, fixed for the thread; it walks pixels, so the per-channel partial sums stay in
// registers until one block-level reduction and C atomics per block.
template <typename OT, int VEC>
__global__ void __launch_bounds__(256) act_bwd_nhwc_kernel(const OT* __restrict__ dy, const OT* __restrict__ y,
                                                           const float* __restrict__ noise, int noise_batched,
                                                           OT* __restrict__ g, float* __restrict__ dbias,
                                                           float* __restrict__ dnw, int C, int HW, int64_t npix, float slope,
                                                           float gain) {
  const int cv = C / VEC;          // power of two <= 256 (host-checked)
  const int ppb = 256 / cv;        // pixels per block iteration
  const int c = (threadIdx.x % cv) * VEC;
  const int prow = threadIdx.x / cv;
  float accb[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) accb[k] = 0.f;
  float accn = 0.f;
  for (int64_t pix = (int64_t)blockIdx.x * ppb + prow; pix < npix; pix += (int64_t)gridDim.x * ppb) {
    const Vec16<OT> vd = ld_vec16_stream(dy + pix * C + c);
    const Vec16<OT> vy = ld_vec16_stream(y + pix * C + c);
    Vec16<OT> vg;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float gv = to_f32<OT>(vd.e[k]) * (to_f32<OT>(vy.e[k]) > 0.f ? gain : gain * slope);
      vg.e[k] = from_f32<OT>(gv);
      accb[k] += gv;
      s += gv;
    }
    st_vec16(g + pix * C + c, vg);
    if (noise) accn = fmaf(s, noise[noise_batched ? pix : pix % HW], accn);
  }
  __shared__ float sb[256 * VEC];
  __shared__ float sn[8];
#pragma unroll
  for (int k = 0; k < VEC; ++k) sb[threadIdx.x * VEC + k] = accb[k];
  accn = warp_sum(accn);
  if ((threadIdx.x & 31) == 0) sn[threadIdx.x >> 5] = accn;
  __syncthreads();
  if (threadIdx.x < cv) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float t = 0.f;
      for (int r = 0; r < ppb; ++r) t += sb[(r * cv + threadIdx.x) * VEC + k];
      atomicAdd(dbias + c + k, t);
    }
  }
  if (threadIdx.x == 0 && noise) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sn[w];
    atomicAdd(dnw, t);
  }
}

// ---- 2. blur backward into output-parity planes -------------------------------------------------------------------------
// forward (blur_act_nhwc_kernel): out[y,x] = sum_{a,e} sk[a][e] mid[y+a-1, x+e-1], sk = flipped blur.kernel
// => gmid[Y,X] = sum_{a,e} sk[a][e] g[Y-a+1, X-e+1].   planes: [4][B][H+1][W+1][C], plane = py*2+px.
template <typename OT, int VEC>
__global__ void __launch_bounds__(256) blur_bwd_planes_kernel(const OT* __restrict__ g, OT* __restrict__ planes,
                                                              const float* __restrict__ kf, int B, int C, int H, int W) {
  __shared__ float sk[16];
  if (threadIdx.x < 16) sk[threadIdx.x] = kf[15 - threadIdx.x];
  __syncthreads();
  const int cv = C / VEC, PH = H + 1, PW = W + 1, OH = 2 * H, OW = 2 * W;
  const int64_t total = (int64_t)4 * B * PH * PW * cv;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * VEC;
    int64_t t = idx / cv;
    const int n = (int)(t % PW);
    t /= PW;
    const int m = (int)(t % PH);
    t /= PH;
    const int b = (int)(t % B);
    const int par = (int)(t / B);
    const int Y = 2 * m + (par >> 1), X = 2 * n + (par & 1);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    if (Y <= OH && X <= OW) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int yy = Y - a + 1;
        if (yy < 0 || yy >= OH) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int xx = X - e + 1;
          if (xx < 0 || xx >= OW) continue;
          const Vec16<OT> v = ld_vec16(g + (((int64_t)b * OH + yy) * OW + xx) * C + c);
          const float kv = sk[a * 4 + e];
#pragma unroll
          for (int k = 0; k < VEC; ++k) acc[k] = fmaf(to_f32<OT>(v.e[k]), kv, acc[k]);
        }
      }
    }
    Vec16<OT> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.e[k] = from_f32<OT>(acc[k]);
    st_vec16(planes + idx * VEC, o);
  }
}

// ---- 3. WpT[bt][i][o] = Wp[bt][o][i]  (operand type, 32x32 shared-memory tiles) ----------------------------------------
template <typename OT>
__global__ void __launch_bounds__(256) transpose_wp_kernel(const OT* __restrict__ wp, OT* __restrict__ wpt, int O, int I) {
  __shared__ OT t[32][33];
  const int64_t base = (int64_t)blockIdx.z * O * I;
  const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int o = o0 + j, i = i0 + tx;
    if (o < O && i < I) t[j][tx] = wp[base + (int64_t)o * I + i];
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int i = i0 + j, o = o0 + tx;
    if (o < O && i < I) wpt[base + (int64_t)i * O + o] = t[tx][j];
  }
}

// ---- 4. weight gradient: pixel-contraction GEMM on tcgen05 with MN-major operands ---------------------------------------
struct WgradParams {
  int B, I, O, H, W;
  int TH, TW, kt_w, kt_total, ksplit;  // K tile = TH x TW pixels, kt_w tiles per image row of tiles
  int G, ngroups;                      // taps per CTA, tap groups (G * ngroups == 9)
  int m_tiles, n_tiles, n_tile;
  int a_atoms, b_atoms;                // 128-byte channel atoms loaded per A / B tile
  int a_shared;                        // 1: one A tile per stage, G shifted B tiles (plain); 0: G A tiles, one B tile (up)
  int stack, apt;                      // plain 3x3, O <= 64: `stack` kernel ROWS share one A tile (apt atoms each), see wgrad_stack()
  int tap_ady[9], tap_adx[9], tap_aboff[9], tap_bdy[9], tap_bdx[9];
  int stages, tmem_cols;
  float* dwp;                          // [B][9][O][I] fp32, zero-initialised (dwp_img_stride = 9*O*I), or one [taps][O][I]
  int64_t dwp_img_stride;              // shared by the batch (dwp_img_stride = 0: every image adds into the same gradient)
};

constexpr int kWgradThreads = 192;

template <bool TF32>
__global__ void __launch_bounds__(kWgradThreads, 1)
    wgrad_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const WgradParams p) {
  constexpr int EPA = TF32 ? 32 : 64;        // channels per 128-byte atom row
  constexpr int KMMA = TF32 ? 8 : 16;        // pixels per MMA
  constexpr int A_ATOMS_FULL = 128 / EPA;    // the A tile always reserves M = 128 rows (unloaded atoms feed unused D rows)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8], acc_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int KT = p.TH * p.TW;
  const int box_bytes = KT * 128;
  const bool ustack = p.stack > 1 && !p.a_shared;   // up-sampling layers: the taps of a group stacked along M (one A tile, one MMA)
  const int nA = (p.a_shared || ustack) ? 1 : p.G, nB = p.a_shared ? p.G : 1;
  const int nacc = ustack ? 1 : p.G;                // accumulators (of n_tile columns) per CTA
  const int a_tile_bytes = A_ATOMS_FULL * box_bytes, b_tile_bytes = p.b_atoms * box_bytes;
  const int stage_bytes = nA * a_tile_bytes + nB * b_tile_bytes;
  // tap group fastest: the CTAs that read the SAME pixel range of x / g (3 tap groups x output tiles, tap-shifted by one pixel) are
  // launched next to each other and share it through L2 — with the split index fastest a 268 MB activation was streamed from
  // DRAM once per tap group and tap (ncu-free evidence: 727 us for 403 MB of algorithmic bytes, 0.09 of the HBM roofline)
  const int split = blockIdx.y, b = blockIdx.z;
  const int grp = blockIdx.x % p.ngroups;
  const int mt = (blockIdx.x / p.ngroups) % p.m_tiles;
  const int nt = blockIdx.x / (p.ngroups * p.m_tiles);
  const int kt0 = (int)((int64_t)p.kt_total * split / p.ksplit);
  const int kt1 = (int)((int64_t)p.kt_total * (split + 1) / p.ksplit);

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (elect_one() && kt1 > kt0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_b);
      const uint32_t bytes = (uint32_t)((nA * p.a_atoms + nB * p.b_atoms) * box_bytes);
      for (int kt = kt0; kt < kt1; ++kt) {
        const int it = kt - kt0, st = it % p.stages;
        const int m0 = (kt / p.kt_w) * p.TH, n0 = (kt % p.kt_w) * p.TW;
        mbar_wait(&empty[st], ((it / p.stages) & 1) ^ 1);
        uint8_t* sA = smem + st * stage_bytes;
        uint8_t* sB = sA + nA * a_tile_bytes;
        if (ustack) {
          // stacked taps: atom group j of the one A tile = the parity plane / shift of tap grp * stack + j, one un-shifted B tile
          const int nvalid = (9 - grp * p.stack) < p.stack ? (9 - grp * p.stack) : p.stack;
          mbar_arrive_expect_tx(&full[st], (uint32_t)((nvalid * p.apt + p.b_atoms) * box_bytes));
          for (int j = 0; j < nvalid; ++j) {
            const int tap = grp * p.stack + j;
            for (int at = 0; at < p.apt; ++at)
              tma_load_4d(sA + (j * p.apt + at) * box_bytes, &map_a, &full[st], at * EPA, n0 + p.tap_adx[tap], m0 + p.tap_ady[tap],
                          b + p.tap_aboff[tap]);
          }
          for (int at = 0; at < p.b_atoms; ++at)
            tma_load_4d(sB + at * box_bytes, &map_b, &full[st], nt * p.n_tile + at * EPA, n0, m0, b);
          continue;
        }
        if (p.stack > 1) {
          // stacked kernel rows: atom group j of the A tile = g shifted by 1 - ky rows (ky = grp * stack + j), the three B tiles
          // = x shifted by kx - 1 columns: D rows [j * apt * EPA, ...) of accumulator kx are the gradient of tap (ky, kx)
          const int nvalid = (3 - grp * p.stack) < p.stack ? (3 - grp * p.stack) : p.stack;
          mbar_arrive_expect_tx(&full[st], (uint32_t)((nvalid * p.apt + 3 * p.b_atoms) * box_bytes));
          for (int j = 0; j < nvalid; ++j)
            for (int at = 0; at < p.apt; ++at)
              tma_load_4d(sA + (j * p.apt + at) * box_bytes, &map_a, &full[st], at * EPA, n0, m0 + 1 - (grp * p.stack + j), b);
          for (int tb = 0; tb < 3; ++tb)
            for (int at = 0; at < p.b_atoms; ++at)
              tma_load_4d(sB + tb * b_tile_bytes + at * box_bytes, &map_b, &full[st], nt * p.n_tile + at * EPA, n0 + tb - 1, m0, b);
          continue;
        }
        mbar_arrive_expect_tx(&full[st], bytes);
        for (int ta = 0; ta < nA; ++ta) {
          const int tap = grp * p.G + ta;
          for (int at = 0; at < p.a_atoms; ++at)
            tma_load_4d(sA + ta * a_tile_bytes + at * box_bytes, &map_a, &full[st], mt * 128 + at * EPA,
                        n0 + p.tap_adx[tap], m0 + p.tap_ady[tap], b + p.tap_aboff[tap]);
        }
        for (int tb = 0; tb < nB; ++tb) {
          const int tap = grp * p.G + tb;
          for (int at = 0; at < p.b_atoms; ++at)
            tma_load_4d(sB + tb * b_tile_bytes + at * box_bytes, &map_b, &full[st], nt * p.n_tile + at * EPA,
                        n0 + p.tap_bdx[tap], m0 + p.tap_bdy[tap], b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one() && kt1 > kt0) {
      const uint32_t idesc = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, 128, p.n_tile, 1, 1);  // both operands MN-major
      const int ksteps = KT / KMMA;
      for (int kt = kt0; kt < kt1; ++kt) {
        const int it = kt - kt0, st = it % p.stages;
        mbar_wait(&full[st], (it / p.stages) & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + st * stage_bytes);
        const uint32_t sB = sA + nA * a_tile_bytes;
        for (int tl = 0; tl < nacc; ++tl) {
          const uint32_t aBase = sA + (p.a_shared ? 0 : tl) * a_tile_bytes;
          const uint32_t bBase = sB + (p.a_shared ? tl : 0) * b_tile_bytes;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t adesc = make_sdesc_mn_sw128(aBase + ks * (KMMA * 128), (uint32_t)box_bytes, TF32);
            const uint64_t bdesc = make_sdesc_mn_sw128(bBase + ks * (KMMA * 128), (uint32_t)box_bytes, TF32);
            const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
            if (TF32) mma_ss_tf32(tmem + tl * p.n_tile, adesc, bdesc, idesc, acc);
            else mma_ss_f16(tmem + tl * p.n_tile, adesc, bdesc, idesc, acc);
          }
        }
        tc_commit(&empty[st]);
      }
      tc_commit(&acc_full);
    }
    __syncwarp();
  } else if (kt1 > kt0) {
    const int lane_base = (warp & 3) * 32;
    const int o = mt * 128 + lane_base + (tid & 31);
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int row = lane_base + (tid & 31);          // accumulator row of this thread
    const int sj = p.stack > 1 ? row / (p.apt * EPA) : 0, so = p.stack > 1 ? row % (p.apt * EPA) : o;
    const int sky = grp * p.stack + sj;               // plain: kernel row; up: tap
    const bool s_ok = p.stack > 1 ? (sj < p.stack && sky < (ustack ? 9 : 3) && so < p.O) : (o < p.O);
    for (int tl = 0; tl < nacc; ++tl) {
      const int tap = ustack ? sky : (p.stack > 1 ? sky * 3 + tl : grp * p.G + tl);
      float* dst = p.dwp + (int64_t)b * p.dwp_img_stride + ((int64_t)tap * p.O + so) * p.I + nt * p.n_tile;
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + tl * p.n_tile + c0, v);
        tc_wait_ld();
        if (s_ok) {
#pragma unroll
          for (int k = 0; k < 32; k += 4)
            red_add_v4(dst + c0 + k, __uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]),
                       __uint_as_float(v[k + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
}

// ---- 5. through modulation / demodulation --------------------------------------------------------------------------------
// One block per output channel o; loops over the batch so dW[o] needs no atomics. ds partials go to ds_part[o][b][i].
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(256) weight_bwd_kernel(const float* __restrict__ G, const float* __restrict__ w,
                                                         const float* __restrict__ s, float* __restrict__ dw,
                                                         float* __restrict__ ds_part, int B, int I, int O, float scale,
                                                         int demodulate) {
  constexpr int T = 9, MAXI = 2;  // I <= 512: at most 2 input channels per thread
  const int o = blockIdx.x;
  const float* wo = w + (int64_t)o * I * T;
  __shared__ float red[8];
  float dwacc[MAXI][T];
#pragma unroll
  for (int j = 0; j < MAXI; ++j)
#pragma unroll
    for (int t = 0; t < T; ++t) dwacc[j][t] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* sb = s + (int64_t)b * I;
    const float* Gb = G + ((int64_t)b * T * O + o) * I;  // + t * O * I + i
    float d = 1.f, c = 0.f;
    if (demodulate) {
      float q = 0.f, cc = 0.f;
#pragma unroll
      for (int j = 0; j < MAXI; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < I) {
          const float si = scale * sb[i];
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const float u = wo[i * T + t] * si;
            q = fmaf(u, u, q);
            cc = fmaf(Gb[(int64_t)t * O * I + i], u, cc);
          }
        }
      }
      q = block_sum_256(q, red);
      cc = block_sum_256(cc, red);
      d = rsqrtf(q + 1e-8f);
      c = cc * d;  // sum G * Wp, Wp = u d
    }
#pragma unroll
    for (int j = 0; j < MAXI; ++j) {
      const int i = threadIdx.x + j * 256;
      if (i < I) {
        const float si = scale * sb[i];
        float dsi = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float wv = wo[i * T + t];
          const float du = d * (Gb[(int64_t)t * O * I + i] - wv * si * d * c);
          dwacc[j][t] = fmaf(du, si, dwacc[j][t]);
          dsi = fmaf(du, wv, dsi);
        }
        ds_part[((int64_t)o * B + b) * I + i] = dsi * scale;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXI; ++j) {
    const int i = threadIdx.x + j * 256;
    if (i < I)
#pragma unroll
      for (int t = 0; t < T; ++t) dw[((int64_t)o * I + i) * T + t] = dwacc[j][t];
  }
}

// out[e] = sum_r part[r][e]
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ part, float* __restrict__ out, int rows,
                                                          int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float a = 0.f;
  for (int r = 0; r < rows; ++r) a += part[(int64_t)r * n + e];
  out[e] = a;
}

// ---- ToRGB backward -----------------------------------------------------------------------------------------------------
// grid (blocks, B). Thread = fixed channel vector; dx written, d rgb_w[b][o][c] and dbias[o] reduced per block.
template <typename OT, int VEC>
__global__ void __launch_bounds__(256) torgb_bwd_nhwc_kernel(const OT* __restrict__ x, const float* __restrict__ drgb,
                                                             const float* __restrict__ rgb_w, OT* __restrict__ dx,
                                                             float* __restrict__ d_rgbw, float* __restrict__ dbias, int C,
                                                             int HW) {
  const int b = blockIdx.y;
  const int cv = C / VEC, ppb = 256 / cv;
  const int c = (threadIdx.x % cv) * VEC, prow = threadIdx.x / cv;
  float w[3][VEC], acc[3][VEC], accb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      w[o][k] = rgb_w[((int64_t)b * 3 + o) * C + c + k];
      acc[o][k] = 0.f;
    }
  const float* dr = drgb + (int64_t)b * 3 * HW;
  for (int pix = blockIdx.x * ppb + prow; pix < HW; pix += gridDim.x * ppb) {
    const float g0 = dr[pix], g1 = dr[HW + pix], g2 = dr[2 * HW + pix];
    const Vec16<OT> vx = ld_vec16_stream(x + ((int64_t)b * HW + pix) * C + c);
    Vec16<OT> vo;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float xv = to_f32<OT>(vx.e[k]);
      vo.e[k] = from_f32<OT>(fmaf(g0, w[0][k], fmaf(g1, w[1][k], g2 * w[2][k])));
      acc[0][k] = fmaf(g0, xv, acc[0][k]);
      acc[1][k] = fmaf(g1, xv, acc[1][k]);
      acc[2][k] = fmaf(g2, xv, acc[2][k]);
    }
    st_vec16(dx + ((int64_t)b * HW + pix) * C + c, vo);
    if (c == 0) {
      accb[0] += g0;
      accb[1] += g1;
      accb[2] += g2;
    }
  }
  __shared__ float sb[256 * VEC];
  for (int o = 0; o < 3; ++o) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VEC; ++k) sb[threadIdx.x * VEC + k] = acc[o][k];
    __syncthreads();
    if (threadIdx.x < cv) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float t = 0.f;
        for (int r = 0; r < ppb; ++r) t += sb[(r * cv + threadIdx.x) * VEC + k];
        atomicAdd(d_rgbw + ((int64_t)b * 3 + o) * C + c + k, t);
      }
    }
  }
  if (c == 0) {
#pragma unroll
    for (int o = 0; o < 3; ++o) atomicAdd(dbias + o, accb[o]);
  }
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

struct BwdLayout {  // workspace carve-up, every block 1024-byte aligned
  int64_t g, planes, wp, wpt, dwp, ds_part, total;
};
BwdLayout bwd_layout(int B, int I, int O, int H, int W, int upsample, int act, int mma) {
  const int esz = esz_of(mma);
  const int OH = upsample ? 2 * H : H, OW = upsample ? 2 * W : W;
  BwdLayout L{};
  int64_t off = 0;
  L.g = off;
  off += act ? align_up((int64_t)B * OH * OW * O * esz, 1024) : 0;
  L.planes = off;
  off += upsample ? align_up((int64_t)4 * B * (H + 1) * (W + 1) * O * esz, 1024) : 0;
  L.wp = off;
  off += align_up((int64_t)B * 9 * O * I * esz, 1024);
  L.wpt = off;
  off += align_up((int64_t)B * 9 * O * I * esz, 1024);
  L.dwp = off;
  off += align_up((int64_t)B * 9 * O * I * 4, 1024);
  L.ds_part = off;
  off += align_up((int64_t)O * B * I * 4, 1024);
  L.total = off;
  return L;
}

// Plain 3x3 convolutions with few output channels: the A tile always spans M = 128 accumulator rows, of which O <= 64 were in use —
// and the kernel is bound by the MMA-issuing thread (ncu, 64 -> 32 @512^2: 536 us, tensor pipe 28 %, DRAM 10 %: 12 instructions of
// 128 x N x 8 per 32 pixels). Stacking kernel ROWS along M (row ky = the gradient tile shifted by 1 - ky image rows, a different
// TMA coordinate per atom group) makes one instruction produce up to three taps: 3 instead of 9 per 8 pixels for O <= 32
// (tf32), 6 for O <= 64, and the activations are read once per CTA instead of once per tap group.
inline void wgrad_stack(WgradParams& p, int O, int epa, int taps) {
  static const bool off = [] { const char* e = getenv("FMI_WGRAD_STACK"); return e && e[0] == '0'; }();
  p.stack = 0; p.apt = 0;
  if (off || taps != 9 || p.m_tiles != 1) return;
  const int apt = (O + epa - 1) / epa, fit = (128 / epa) / apt;
  if (fit < 2) return;
  if (p.a_shared) {
    if (p.G != 3) return;
    p.stack = fit > 3 ? 3 : fit;
    p.apt = apt;
    p.ngroups = (3 + p.stack - 1) / p.stack;
  } else {
    // up-sampling layers (one A tile per tap from the gradient's parity planes, one un-shifted B tile): the taps of a CTA stacked
    // along M — one MMA per 8 pixels yields `stack` taps and the CTA needs one accumulator
    static const bool uoff = [] { const char* e = getenv("FMI_WGRAD_STACK_UP"); return e && e[0] == '0'; }();
    if (uoff) return;
    p.stack = fit > 3 ? 3 : fit;
    p.apt = apt;
    p.G = p.stack;
    p.ngroups = (9 + p.stack - 1) / p.stack;
  }
}

// The pixels of a K tile per stage: 32 (tf32) / 64 (bf16) pixels make 4 KB TMA boxes, and the single producer thread issues about one
// box per 200 ns (measured: 5 boxes per stage -> 1.35 us per stage whatever the MMAs do) — twice the pixels per box halves the
// instruction count. Grown while at least 4 stages still fit in shared memory. Call after wgrad_stack().
inline void wgrad_grow_ktile(WgradParams& p, bool tf32, int H, int W) {
  static const bool off = [] { const char* e = getenv("FMI_WGRAD_KTILE"); return e && e[0] == '0'; }();
  if (off) return;
  const bool ustack = p.stack > 1 && !p.a_shared;
  const int nA = (p.a_shared || ustack) ? 1 : p.G, nB = p.a_shared ? p.G : 1;
  const int atoms = nA * (tf32 ? 4 : 2) + nB * p.b_atoms;
  int KT = p.TH * p.TW;
  while (2 * KT <= 256 && 2 * KT <= H * W && (int64_t)atoms * 2 * KT * 128 * 4 <= 232448 - 4096) {
    const int tw = W < 2 * KT ? W : 2 * KT, th = 2 * KT / tw;
    if (W % tw || H % th || (H / th) * (W / tw) < 8) break;
    KT *= 2;
    p.TW = tw; p.TH = th;
  }
  p.kt_w = W / p.TW;
  p.kt_total = (H / p.TH) * p.kt_w;
}

template <bool TF32>
int launch_wgrad(const CUtensorMap& ma, const CUtensorMap& mb, WgradParams p, cudaStream_t st) {
  auto kern = wgrad_gemm_kernel<TF32>;
  static FmiPerDeviceOnce attr_once;
  if (attr_once.need()) {
    FMI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 2048));
    attr_once.done();
  }
  constexpr int EPA = TF32 ? 32 : 64;
  const int box_bytes = p.TH * p.TW * 128;
  const bool ustack = p.stack > 1 && !p.a_shared;
  const int nA = (p.a_shared || ustack) ? 1 : p.G, nB = p.a_shared ? p.G : 1;
  const int stage_bytes = nA * (128 / EPA) * box_bytes + nB * p.b_atoms * box_bytes;
  int stages = (232448 - 4096) / stage_bytes;
  if (stages > 8) stages = 8;
  FMI_REQUIRE(stages >= 2, "modconv wgrad: stage of %d bytes does not fit twice in shared memory", stage_bytes);
  p.stages = stages;
  const int cols = (ustack ? 1 : p.G) * p.n_tile;
  p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  dim3 grid(p.ngroups * p.m_tiles * p.n_tiles, p.ksplit, p.B);
  kern<<<grid, kWgradThreads, (size_t)stages * stage_bytes + 1024, st>>>(ma, mb, p);
  return fmi_launched("modconv_wgrad");
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" int64_t fmi_styled_conv_bwd_workspace_bytes(int B, int I, int O, int H, int W, int upsample, int act, int mma) {
  return bwd_layout(B, I, O, H, W, upsample, act, mma).total;
}

extern "C" int fmi_styled_conv_bwd_nhwc(const void* x, const void* y, const void* dy, const float* weight, const float* s,
                                        const float* noise, int noise_batched, const float* blur_k, void* dx,
                                        float* dweight, float* ds, float* dnoise_w, float* dbias, int B, int I, int O, int H,
                                        int W, int upsample, int act, int demodulate, int mma, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "styled_conv_bwd: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && dy && weight && s && dweight && ds && workspace, "styled_conv_bwd: null pointer");
  FMI_REQUIRE(!act || (y && dbias && dnoise_w), "styled_conv_bwd: act needs y, dbias and dnoise_w");
  FMI_REQUIRE(I >= 32 && I % 32 == 0 && O >= 32 && O % 32 == 0 && I <= 512 && (I <= 256 || I % 256 == 0) &&
                  (O <= 128 || O % 128 == 0) && (O <= 256 || O % 256 == 0),
              "styled_conv_bwd: unsupported channel counts I=%d O=%d", I, O);
  FMI_REQUIRE(is_pow2(H) && is_pow2(W) && H * W >= 16 && W >= 4, "styled_conv_bwd: H=%d W=%d must be powers of two (>= 4)", H, W);
  FMI_REQUIRE(!upsample || blur_k, "styled_conv_bwd: upsample needs the 4x4 blur kernel");
  const BwdLayout L = bwd_layout(B, I, O, H, W, upsample, act, mma);
  FMI_REQUIRE(workspace_bytes >= L.total, "styled_conv_bwd: workspace too small (%lld < %lld)", (long long)workspace_bytes,
              (long long)L.total);
  FMI_REQUIRE(fmi_aligned(workspace, 256) && fmi_aligned(x, 16) && fmi_aligned(dy, 16) && (!dx || fmi_aligned(dx, 16)),
              "styled_conv_bwd: buffers must be 16-byte aligned (workspace 256)");
  int rc = fmi_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // whole backward of one StyledConv (activation + blur backward, data gradient GEMM, weight-gradient GEMM, chain rule)
  FmiProfScope prof(FMI_PROF_WGRAD, st, 4.0 * B * (double)H * W * (upsample ? 1.0 : 1.0) * 9.0 * I * O,
                    (double)B * H * W * (2.0 * I + (upsample ? 8.0 : 2.0) * O) * (mma == FMI_MMA_TF32 ? 4.0 : 2.0));

  const bool tf32 = mma == FMI_MMA_TF32;
  const int esz = esz_of(mma);
  const int vec = 16 / esz;
  const uint32_t epa = 128 / esz;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int OH = upsample ? 2 * H : H, OW = upsample ? 2 * W : W;
  uint8_t* ws = (uint8_t*)workspace;
  const float slope = 0.2f, gain = 1.4142135623730951f;

  // ---- 1. g = dy * act'(y), dbias, dnoise_w
  const void* g = dy;
  if (act) {
    FMI_REQUIRE(is_pow2(O / vec) && O / vec <= 256, "styled_conv_bwd: O=%d must give a power-of-two vector count", O);
    FMI_CUDA(cudaMemsetAsync(dbias, 0, (size_t)O * 4, st));
    FMI_CUDA(cudaMemsetAsync(dnoise_w, 0, 4, st));
    const int64_t npix = (int64_t)B * OH * OW;
    const int ppb = 256 / (O / vec);
    const int grid = (int)imin64((npix + ppb - 1) / ppb, (int64_t)FMI_NUM_SMS * 8);
    void* gout = ws + L.g;
    if (tf32)
      act_bwd_nhwc_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)dy, (const float*)y, noise, noise_batched,
                                                          (float*)gout, dbias, dnoise_w, O, OH * OW, npix, slope, gain);
    else
      act_bwd_nhwc_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, noise,
                                                                  noise_batched, (__nv_bfloat16*)gout, dbias, dnoise_w, O,
                                                                  OH * OW, npix, slope, gain);
    rc = fmi_launched("act_bwd");
    if (rc) return rc;
    g = gout;
  }
  // ---- 2. blur backward into parity planes
  const void* gsrc = g;  // tensor the two GEMMs read: g [B,H,W,O] (plain) or planes [4B,H+1,W+1,O] (up)
  if (upsample) {
    const int64_t total = (int64_t)4 * B * (H + 1) * (W + 1) * (O / vec);
    const int grid = (int)imin64((total + 255) / 256, (int64_t)FMI_NUM_SMS * 32);
    void* pl = ws + L.planes;
    if (tf32) blur_bwd_planes_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)g, (float*)pl, blur_k, B, O, H, W);
    else blur_bwd_planes_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (__nv_bfloat16*)pl, blur_k, B, O, H, W);
    rc = fmi_launched("blur_bwd_planes");
    if (rc) return rc;
    gsrc = pl;
  }
  const int GH = upsample ? H + 1 : H, GW = upsample ? W + 1 : W, GB = upsample ? 4 * B : B;
  int tap_dy[9], tap_dx[9], tap_boff[9];  // where tap t reads `gsrc` relative to the x pixel (wgrad) — dgrad negates (plain)
  for (int t = 0; t < 9; ++t) {
    const int ky = t / 3, kx = t % 3;
    if (upsample) { tap_dy[t] = ky / 2; tap_dx[t] = kx / 2; tap_boff[t] = ((ky & 1) * 2 + (kx & 1)) * B; }
    else { tap_dy[t] = ky - 1; tap_dx[t] = kx - 1; tap_boff[t] = 0; }
  }

  // ---- 3. data gradient (skipped when dx == NULL, e.g. the constant input needs it but a frozen trunk does not)
  if (dx) {
    rc = fmi_modconv_weight_prep(weight, s, ws + L.wp, B, I, O, 3, demodulate, mma, stream);
    if (rc) return rc;
    dim3 tg((I + 31) / 32, (O + 31) / 32, B * 9);
    if (tf32) transpose_wp_kernel<float><<<tg, 256, 0, st>>>((const float*)(ws + L.wp), (float*)(ws + L.wpt), O, I);
    else transpose_wp_kernel<__nv_bfloat16><<<tg, 256, 0, st>>>((const __nv_bfloat16*)(ws + L.wp), (__nv_bfloat16*)(ws + L.wpt), O, I);
    rc = fmi_launched("transpose_wp");
    if (rc) return rc;
    ConvGemmParams p{};
    p.B = B; p.I = O; p.O = I; p.H = GH; p.W = GW; p.T = 9;
    p.n_tile = I <= 256 ? I : 256;
    p.k_chunks = (O + epa - 1) / epa;
    p.OH = H; p.OW = W; p.Mh = H; p.Mw = W; p.sy = p.sx = 1; p.py = p.px = 0;
    p.ntaps = 9; p.act = 0; p.out = dx; p.slope = slope; p.gain = gain;
    for (int t = 0; t < 9; ++t) {
      // plain: dx[p] += g[p - off_t] WpT[t]; up: dx[m,n] += P[par_t][m + ky/2][n + kx/2] WpT[t]
      p.tap_dy[t] = upsample ? tap_dy[t] : -tap_dy[t];
      p.tap_dx[t] = upsample ? tap_dx[t] : -tap_dx[t];
      p.tap_boff[t] = tap_boff[t];
      p.tap_slab[t] = t;
    }
    CUtensorMap mw, mx;
    {
      uint64_t dims[2] = {(uint64_t)O, (uint64_t)B * 9 * I};
      uint64_t str[1] = {(uint64_t)O * esz};
      uint32_t box[2] = {epa, (uint32_t)p.n_tile};
      int e = make_tensor_map(&mw, dt, 2, ws + L.wpt, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
      FMI_REQUIRE(e == 0, "styled_conv_bwd: cuTensorMapEncodeTiled(WpT) failed (%d)", e);
    }
    {
      TilePlan tp = pick_tile(p.Mh, p.Mw);
      const bool halo_env = [] { const char* e = getenv("FMI_MODCONV_HALO"); return e && e[0] == '1'; }();  // opt-in (read per call): measured 4-7 % slower
      p.halo = halo_env && !upsample && W >= 128 && p.n_tile <= 128;
      if (p.halo) {
        // dx[p] += g[p + (oy, ox)] WpT[t] with (oy, ox) = -(tap offset of t): kernel row dyi reads offset dyi - 1
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c) p.halo_slab[a][c] = (2 - a) * 3 + (2 - c);
        tp = TilePlan{1, 130, 0, 0};
      }
      uint64_t dims[4] = {(uint64_t)O, (uint64_t)GW, (uint64_t)GH, (uint64_t)GB};
      uint64_t str[3] = {(uint64_t)O * esz, (uint64_t)GW * O * esz, (uint64_t)GH * GW * O * esz};
      uint32_t box[4] = {epa, (uint32_t)tp.TW, (uint32_t)tp.TH, 1};
      int e = make_tensor_map(&mx, dt, 4, gsrc, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
      FMI_REQUIRE(e == 0, "styled_conv_bwd: cuTensorMapEncodeTiled(g) failed (%d)", e);
    }
    rc = tf32 ? launch_gemm_class<true>(mx, mw, p, st) : launch_gemm_class<false>(mx, mw, p, st);
    if (rc) return rc;
  }

  // ---- 4. per-sample weight gradient G[b][t][o][i]
  {
    FMI_CUDA(cudaMemsetAsync(ws + L.dwp, 0, (size_t)B * 9 * O * I * 4, st));
    WgradParams p{};
    p.B = B; p.I = I; p.O = O; p.H = H; p.W = W;
    int KT = tf32 ? 32 : 64;
    if (KT > H * W) KT = H * W;
    p.TW = W < KT ? W : KT;
    p.TH = KT / p.TW;
    p.kt_w = W / p.TW;
    p.kt_total = (H / p.TH) * p.kt_w;
    p.n_tile = I <= 256 ? I : 256;
    p.n_tiles = I / p.n_tile;
    p.m_tiles = (O + 127) / 128;
    p.G = p.n_tile <= 128 ? 3 : 1;
    p.ngroups = 9 / p.G;
    p.a_atoms = ((O < 128 ? O : 128) + (int)epa - 1) / (int)epa;
    p.b_atoms = (p.n_tile + (int)epa - 1) / (int)epa;
    p.a_shared = upsample ? 0 : 1;
    wgrad_stack(p, O, (int)epa, 9);
    wgrad_grow_ktile(p, tf32, H, W);
    for (int t = 0; t < 9; ++t) {
      if (upsample) { p.tap_ady[t] = tap_dy[t]; p.tap_adx[t] = tap_dx[t]; p.tap_aboff[t] = tap_boff[t]; }
      else { p.tap_bdy[t] = tap_dy[t]; p.tap_bdx[t] = tap_dx[t]; }
    }
    const int base = B * p.ngroups * p.m_tiles * p.n_tiles;
    int ksplit = (2 * FMI_NUM_SMS) / base;                 // one CTA per SM: whole waves (300 CTAs ran as 3 waves of 148 + 148 + 4)
    if (ksplit > p.kt_total / 4) ksplit = p.kt_total / 4;  // at least 4 K tiles per CTA
    if (ksplit < 1) ksplit = 1;
    p.ksplit = ksplit;
    p.dwp = (float*)(ws + L.dwp);
    p.dwp_img_stride = (int64_t)9 * O * I;
    CUtensorMap ma, mb;
    // MN-major tf32 operands need the 32-byte-atom flavour of the 128-byte swizzle (sm100.cuh make_sdesc_mn_sw128)
    const CUtensorMapSwizzle wswz = tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    {
      uint64_t dims[4] = {(uint64_t)O, (uint64_t)GW, (uint64_t)GH, (uint64_t)GB};
      uint64_t str[3] = {(uint64_t)O * esz, (uint64_t)GW * O * esz, (uint64_t)GH * GW * O * esz};
      uint32_t box[4] = {epa, (uint32_t)p.TW, (uint32_t)p.TH, 1};
      int e = make_tensor_map(&ma, dt, 4, gsrc, dims, str, box, wswz);
      FMI_REQUIRE(e == 0, "styled_conv_bwd: cuTensorMapEncodeTiled(wgrad A) failed (%d)", e);
    }
    {
      uint64_t dims[4] = {(uint64_t)I, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      uint64_t str[3] = {(uint64_t)I * esz, (uint64_t)W * I * esz, (uint64_t)H * W * I * esz};
      uint32_t box[4] = {epa, (uint32_t)p.TW, (uint32_t)p.TH, 1};
      int e = make_tensor_map(&mb, dt, 4, x, dims, str, box, wswz);
      FMI_REQUIRE(e == 0, "styled_conv_bwd: cuTensorMapEncodeTiled(wgrad B) failed (%d)", e);
    }
    rc = tf32 ? launch_wgrad<true>(ma, mb, p, st) : launch_wgrad<false>(ma, mb, p, st);
    if (rc) return rc;
  }
  // ---- 5. dW, ds
  {
    const float scale = 1.0f / sqrtf((float)(I * 9));
    weight_bwd_kernel<<<O, 256, 0, st>>>((const float*)(ws + L.dwp), weight, s, dweight, (float*)(ws + L.ds_part), B, I, O,
                                         scale, demodulate);
    rc = fmi_launched("weight_bwd");
    if (rc) return rc;
    reduce_rows_kernel<<<(B * I + 255) / 256, 256, 0, st>>>((const float*)(ws + L.ds_part), ds, O, B * I);
    rc = fmi_launched("reduce_rows");
    if (rc) return rc;
  }
  return FMI_OK;
}

// Weight gradient of a batch-shared Conv2d(ksize 1 or 3, stride 1, padding ksize/2) — the PICNet conv blocks in training
// (SpectralNorm-wrapped nn.Conv2d of base_function.py:207-305 / network.py:73-365, differentiated by ATen / cuDNN in the reference):
//   dwp[t][o][i] += sum_{b, pixels} dy[b, p, o] * x[b, p + off_t, i]      (fp32, accumulated into the caller's zeroed buffer)
// the pixel-contraction GEMM above with every image adding into ONE gradient. x [B,H,W,I], dy [B,H,W,O] dense NHWC in the
// operand type (fp32 read as tf32 / bf16). transposed = 1: the weight gradient of ConvTranspose2d(3, stride 2, padding 1,
// output_padding 1) (ResBlockDecoder's conv2 / bypass, base_function.py:330-336): `dy` is then the 4 pixel-parity planes
// [4][B][H][W][O] of the [B,2H,2W,O] output gradient, dwp[t][o][i] = sum dy[b, 2m-1+ky, 2n-1+kx, o] * x[b,m,n,i].
extern "C" int fmi_conv_wgrad_nhwc(const void* x, const void* dy, float* dwp, int B, int I, int O, int H, int W, int ksize,
                                   int transposed, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "conv_wgrad: bad mma");
  FMI_REQUIRE(ksize == 1 || ksize == 3, "conv_wgrad: ksize must be 1 or 3");
  FMI_REQUIRE(!transposed || ksize == 3, "conv_wgrad: the transposed convolution is 3x3");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && dy && dwp, "conv_wgrad: null pointer");
  FMI_REQUIRE(I >= 32 && I % 32 == 0 && O >= 32 && O % 32 == 0 && (I <= 256 || I % 256 == 0) && (O <= 128 || O % 128 == 0),
              "conv_wgrad: unsupported channel counts I=%d O=%d", I, O);
  FMI_REQUIRE(is_pow2(H) && is_pow2(W) && H * W >= 16 && W >= 4, "conv_wgrad: H=%d W=%d must be powers of two (>= 4)", H, W);
  FMI_REQUIRE(fmi_aligned(x, 16) && fmi_aligned(dy, 16) && fmi_aligned(dwp, 16), "conv_wgrad: buffers must be 16-byte aligned");
  int rc = fmi_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tf32 = mma == FMI_MMA_TF32;
  const int esz = esz_of(mma);
  const uint32_t epa = 128 / esz;
  const int taps = ksize * ksize;
  FmiProfScope prof(FMI_PROF_WGRAD, st, 2.0 * B * (double)H * W * taps * I * O, (double)B * H * W * (I + O) * esz + 4.0 * taps * I * O);
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  WgradParams p{};
  p.B = B; p.I = I; p.O = O; p.H = H; p.W = W;
  int KT = tf32 ? 32 : 64;
  if (KT > H * W) KT = H * W;
  p.TW = W < KT ? W : KT;
  p.TH = KT / p.TW;
  p.kt_w = W / p.TW;
  p.kt_total = (H / p.TH) * p.kt_w;
  p.n_tile = I <= 256 ? I : 256;
  p.n_tiles = I / p.n_tile;
  p.m_tiles = (O + 127) / 128;
  p.G = (taps == 9 && p.n_tile <= 128) ? 3 : 1;
  p.ngroups = taps / p.G;
  p.a_atoms = ((O < 128 ? O : 128) + (int)epa - 1) / (int)epa;
  p.b_atoms = (p.n_tile + (int)epa - 1) / (int)epa;
  p.a_shared = transposed ? 0 : 1;
  wgrad_stack(p, O, (int)epa, taps);
  wgrad_grow_ktile(p, tf32, H, W);
  for (int t = 0; t < taps; ++t) {
    const int ky = t / 3, kx = t % 3;
    if (transposed) {
      // dy[2m - 1 + ky] = plane (ky + 1) & 1 at row m - (ky == 0): the parity planes of fmi_space_to_planes_nhwc
      p.tap_ady[t] = ky == 0 ? -1 : 0;
      p.tap_adx[t] = kx == 0 ? -1 : 0;
      p.tap_aboff[t] = ((((ky + 1) & 1) << 1) | ((kx + 1) & 1)) * B;
    } else {
      p.tap_bdy[t] = ksize == 3 ? ky - 1 : 0;
      p.tap_bdx[t] = ksize == 3 ? kx - 1 : 0;
    }
  }
  const int base = B * p.ngroups * p.m_tiles * p.n_tiles;
  int ksplit = (2 * FMI_NUM_SMS) / base;     // whole waves of one CTA per SM
  if (ksplit > p.kt_total / 4) ksplit = p.kt_total / 4;
  if (ksplit < 1) ksplit = 1;
  p.ksplit = ksplit;
  p.dwp = dwp;
  p.dwp_img_stride = 0;
  CUtensorMap ma, mb;
  const CUtensorMapSwizzle wswz = tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)W, (uint64_t)H, (uint64_t)(transposed ? 4 * B : B)};
    uint64_t str[3] = {(uint64_t)O * esz, (uint64_t)W * O * esz, (uint64_t)H * W * O * esz};
    uint32_t box[4] = {epa, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    int e = make_tensor_map(&ma, dt, 4, dy, dims, str, box, wswz);
    FMI_REQUIRE(e == 0, "conv_wgrad: cuTensorMapEncodeTiled(dy) failed (%d)", e);
  }
  {
    uint64_t dims[4] = {(uint64_t)I, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)I * esz, (uint64_t)W * I * esz, (uint64_t)H * W * I * esz};
    uint32_t box[4] = {epa, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    int e = make_tensor_map(&mb, dt, 4, x, dims, str, box, wswz);
    FMI_REQUIRE(e == 0, "conv_wgrad: cuTensorMapEncodeTiled(x) failed (%d)", e);
  }
  return tf32 ? launch_wgrad<true>(ma, mb, p, st) : launch_wgrad<false>(ma, mb, p, st);
}

extern "C" int fmi_torgb_bwd_nhwc(const void* x, const float* drgb, const float* rgb_w, void* dx, float* d_rgbw,
                                  float* dbias, int B, int I, int H, int W, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "torgb_bwd: bad mma");
  if (B == 0) return FMI_OK;
  FMI_REQUIRE(x && drgb && rgb_w && dx && d_rgbw && dbias, "torgb_bwd: null pointer");
  const int esz = esz_of(mma), vec = 16 / esz;
  FMI_REQUIRE(I % vec == 0 && is_pow2(I / vec) && I / vec <= 256, "torgb_bwd: I=%d must give a power-of-two vector count", I);
  cudaStream_t st = (cudaStream_t)stream;
  FMI_CUDA(cudaMemsetAsync(d_rgbw, 0, (size_t)B * 3 * I * 4, st));
  FMI_CUDA(cudaMemsetAsync(dbias, 0, 12, st));
  const int HW = H * W, ppb = 256 / (I / vec);
  int gx = (HW + ppb - 1) / ppb;
  const int cap = (FMI_NUM_SMS * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  dim3 grid(gx, B);
  if (mma == FMI_MMA_TF32)
    torgb_bwd_nhwc_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)x, drgb, rgb_w, (float*)dx, d_rgbw, dbias, I, HW);
  else
    torgb_bwd_nhwc_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, drgb, rgb_w, (__nv_bfloat16*)dx,
                                                                  d_rgbw, dbias, I, HW);
  return fmi_launched("torgb_bwd");
}
