// Backward of the fused attention (ExampleGuidedAttention / Auto_Attn), first version: the S x S maps are
// materialised ONE IMAGE AT A TIME in the tensor-core operand type and every contraction is a tcgen05 GEMM
// (gemm.cuh). Round-1 design choice: correct, native, tensor-core; a flash-style (never materialised) backward is the
// follow-up.
//
// With O[c,i] = sum_j P[i,j] V[c,j], out_g = a_g w_i O_g + r_i v_g  (include/fmi_b200.h), upstream grads dOut_g:
//   dO'[c,i]  = a_g w_i dOut_g[c,i]                                  (effective grad of O)          prep kernel
//   delta_i   = sum_c dO'[c,i] O[c,i]                                (O saved by the forward)       prep kernel
//   da_g      = sum_{c,i} w_i O_g[c,i] dOut_g[c,i]                   (grad of gamma / alpha)        prep kernel
//   E = q^T q (hi/lo split, K = 3 dpad), P = exp(E - lse_i), PT[j,i] = P[i,j] = exp(E[j,i] - lse_i)  GEMM 1 (EPI_EXP_SYM)
//   dP[i,j]   = sum_c dO'[c,i] V[c,j] ;  dE = P o (dP - delta_i)                                     GEMM 2 (EPI_DS)
//   dV[c,j]   = sum_i dO'[c,i] P[i,j]  (+ r_j dOut_g[c,j], the blend's direct path)                  GEMM 3 (EPI_ADD_COLSCALE)
//   G = dE + dE^T  (keys == queries: the same q plays the row and the column role)                  transpose-add kernel
//   dq[:,t]   = sum_j G[t,j] q[:,j]                                                                  GEMM 4
// dq goes back to the caller as [N, S, dpad] fp32; the 1x1-conv parameter/input gradients from dq are plain GEMMs that
// the Python side leaves to the library (cuBLAS).  Formulas: SURVEY.md §8.1 (verified there against autograd in fp64).
#include <stdlib.h>

#include "gemm.cuh"

using namespace fmi_gemm;

namespace {

template <typename OT, bool TF32> __device__ __forceinline__ OT to_op(float v) {
  if constexpr (TF32) return __uint_as_float(sm100::f32_to_tf32_rna(v));
  else return __float2bfloat16_rn(v);
}

// dO' (operand type, [N,Cv,S]), delta [N,S], da0/da1 (atomic). One thread per (n, i); loops over channels.
template <typename T, typename OT, bool TF32>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const T* __restrict__ dout0, int64_t dout0_bs,
                                                            const T* __restrict__ dout1, int64_t dout1_bs,
                                                            const T* __restrict__ o_saved, const float* __restrict__ mask,
                                                            const float* __restrict__ a0, const float* __restrict__ a1,
                                                            int masked0, int masked1, OT* __restrict__ dop,
                                                            float* __restrict__ delta, float* __restrict__ da0,
                                                            float* __restrict__ da1, int C0, int C1, int S) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float acc_delta = 0.f, acc_a0 = 0.f, acc_a1 = 0.f;
  if (i < S) {
    const float m = mask ? mask[(int64_t)n * S + i] : 0.f;
    const float w0 = masked0 ? 1.f - m : 1.f, w1 = masked1 ? 1.f - m : 1.f;
    const float s0 = (a0 ? *a0 : 1.f) * w0, s1 = (a1 ? *a1 : 1.f) * w1;
    const int Cv = C0 + C1;
    for (int c = 0; c < Cv; ++c) {
      const bool g1 = c >= C0;
      const float g = g1 ? to_f32<T>(dout1[(int64_t)n * dout1_bs + (int64_t)(c - C0) * S + i])
                         : to_f32<T>(dout0[(int64_t)n * dout0_bs + (int64_t)c * S + i]);
      const float o = to_f32<T>(o_saved[((int64_t)n * Cv + c) * S + i]);
      const OT dr = to_op<OT, TF32>((g1 ? s1 : s0) * g);
      dop[((int64_t)n * Cv + c) * S + i] = dr;
      // delta from the ROUNDED operand: dE = P o (dP - delta) then cancels dP's rounding exactly where P is peaked
      acc_delta = fmaf(to_f32<OT>(dr), o, acc_delta);
      if (g1) acc_a1 = fmaf(w1 * o, g, acc_a1);
      else acc_a0 = fmaf(w0 * o, g, acc_a0);
    }
    delta[(int64_t)n * S + i] = acc_delta;
  }
  __shared__ float red[2][8];
  acc_a0 = warp_sum(acc_a0);
  acc_a1 = warp_sum(acc_a1);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = acc_a0;
    red[1][threadIdx.x >> 5] = acc_a1;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float x0 = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.f, x1 = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.f;
    x0 = warp_sum(x0);
    x1 = warp_sum(x1);
    if (threadIdx.x == 0) {
      if (da0) atomicAdd(da0, x0);
      if (da1 && C1 > 0) atomicAdd(da1, x1);
    }
  }
}

// batched transpose [R, C] -> [C, R] of operand-type matrices (ldi / ldo in elements)
template <typename OT>
__global__ void __launch_bounds__(256) transpose_kernel(const OT* __restrict__ in, int64_t ldi, int64_t in_bs,
                                                        OT* __restrict__ out, int64_t ldo, int64_t out_bs, int R, int C) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    t[j][tx] = (r < R && c < C) ? to_f32<OT>(in[(int64_t)b * in_bs + (int64_t)r * ldi + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < C && r < R) out[(int64_t)b * out_bs + (int64_t)c * ldo + r] = from_f32<OT>(t[tx][j]);
  }
}

// G = X + X^T for square operand-type matrices (out-of-place, blockIdx.z = image; strides in elements)
template <typename OT, bool TF32>
__global__ void __launch_bounds__(256) sym_add_kernel(const OT* __restrict__ x, OT* __restrict__ g, int S, int64_t x_bs,
                                                      int64_t g_bs) {
  __shared__ float t[32][33];
  x += (int64_t)blockIdx.z * x_bs;
  g += (int64_t)blockIdx.z * g_bs;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) t[j][tx] = to_f32<OT>(x[(int64_t)(c0 + j) * S + r0 + tx]);  // block (c0, r0) of X
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t idx = (int64_t)(r0 + j) * S + c0 + tx;
    g[idx] = to_op<OT, TF32>(to_f32<OT>(x[idx]) + t[tx][j]);
  }
}

// G is written into the (fp32-sized) P buffer: per-image stride of that buffer in OT elements
template <typename OT> inline int64_t P_stride_elems(int64_t SS) { return SS * 4 / (int64_t)sizeof(OT); }

// From the staged Qt [nb, S, qrow] bf16 ([hi | lo]) (blockIdx.y = image of the group):
//   TF32: QA = [hi | hi | lo], QB = [hi | lo | hi]  (fp32, K = 3 dpad) so that QA.QB^T = hi.hi + hi.lo + lo.hi;
//         Qn [dpad, S] = tf32(hi + lo) transposed
//   BF16: QA = QB = hi (K = dpad), Qn [dpad, S] = hi transposed
template <typename OT, bool TF32>
__global__ void __launch_bounds__(256) q_operands_kernel(const __nv_bfloat16* __restrict__ qt, OT* __restrict__ qa,
                                                         OT* __restrict__ qb, OT* __restrict__ qn, OT* __restrict__ qr, int S,
                                                         int dpad, int split) {
  const int rowlen = dpad * (1 + split);
  const int kq = TF32 ? 3 * dpad : dpad;
  const int64_t total = (int64_t)S * dpad;
  const int b = blockIdx.y;
  qt += (int64_t)b * S * rowlen;
  qa += (int64_t)b * S * kq;
  if (TF32) qb += (int64_t)b * S * kq;
  qn += (int64_t)b * dpad * S;
  qr += (int64_t)b * S * dpad;    // Qr [S, dpad] = the same values as Qn, pixel-major (the column-role term of dq contracts over i)
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = e / dpad;
    const int k = (int)(e % dpad);
    const float hi = __bfloat162float(qt[s * rowlen + k]);
    const float lo = split ? __bfloat162float(qt[s * rowlen + dpad + k]) : 0.f;
    if constexpr (TF32) {
      qa[s * kq + k] = hi;
      qa[s * kq + dpad + k] = hi;
      qa[s * kq + 2 * dpad + k] = lo;
      qb[s * kq + k] = hi;
      qb[s * kq + dpad + k] = lo;
      qb[s * kq + 2 * dpad + k] = hi;
      qn[(int64_t)k * S + s] = to_op<OT, TF32>(hi + lo);
      qr[s * dpad + k] = to_op<OT, TF32>(hi + lo);
    } else {
      qa[s * kq + k] = __float2bfloat16_rn(hi);
      qn[(int64_t)k * S + s] = __float2bfloat16_rn(hi);
      qr[s * dpad + k] = __float2bfloat16_rn(hi);
    }
  }
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

inline int64_t al(int64_t b) { return (b + 1023) / 1024 * 1024; }

struct BwdPlan {
  int dpad, split, esz, kq, nb;  // nb = images processed per group of batched launches
  int64_t qt, vcat, qa, qb, qn, qr, dop, dot, vt, delta, rvec, p, pt, ds, total;
};

int make_bwd_plan(int N, int d, int C0, int C1, int S, int mma, BwdPlan* pl) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "attn_bwd: bad mma");
  FMI_REQUIRE(N >= 1 && d >= 1 && d <= 256 && C0 >= 32 && C1 >= 0 && C0 % 32 == 0 && C1 % 32 == 0, "attn_bwd: bad shape");
  FMI_REQUIRE(S >= 128 && S % 128 == 0, "attn_bwd: S=%d must be a positive multiple of 128", S);
  const int esz = mma == FMI_MMA_TF32 ? 4 : 2;
  const int Cv = C0 + C1;
  pl->esz = esz;
  pl->split = mma == FMI_MMA_TF32 ? 1 : 0;
  pl->dpad = (d + 63) / 64 * 64;
  pl->kq = pl->split ? 3 * pl->dpad : pl->dpad;
  // the S x S maps dominate the scratch: as many images per group as fit in ~8 GiB (at least one)
  const int64_t per_image = (int64_t)S * S * (4 + 2 * esz);
  int64_t nb = (8ll << 30) / per_image;
  if (nb < 1) nb = 1;
  if (nb > N) nb = N;
  pl->nb = (int)nb;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += al(bytes); return o; };
  pl->qt = take((int64_t)N * S * pl->dpad * (1 + pl->split) * 2);
  pl->vcat = take((int64_t)N * Cv * S * esz);
  pl->dop = take((int64_t)N * Cv * S * esz);
  pl->delta = take((int64_t)N * S * 4);
  pl->rvec = take((int64_t)S * 4);
  // per-group buffers (reused for every group of nb images)
  pl->qa = take(nb * S * pl->kq * esz);
  pl->qb = take(nb * S * pl->kq * esz);
  pl->qn = take(nb * pl->dpad * S * esz);
  pl->qr = take(nb * S * pl->dpad * esz);
  pl->dot = take(nb * S * Cv * esz);
  pl->vt = take(nb * S * Cv * esz);
  pl->p = take(nb * S * S * 4);  // P stays fp32 (unrounded); later reused for G
  pl->pt = take(nb * S * S * esz);
  pl->ds = take(nb * S * S * esz);
  pl->total = off;
  return FMI_OK;
}

// forward's staging kernels live in attention.cu
}  // namespace

int fmi_attn_stage_operands(const void* x, const float* wq, const float* bq, const void* v0, const void* v1, void* qt,
                            void* vcat, int N, int C, int d, int C0, int C1, int S, int dtype, int mma, cudaStream_t st);

namespace {

template <typename T, bool TF32>
int run_bwd(const BwdPlan& pl, uint8_t* ws, const float* mask, const float* a0, float b0, int masked0, const float* a1,
            float b1, int masked1, const void* o_saved, const float* lse, const void* dout0, int64_t dout0_bs,
            const void* dout1, int64_t dout1_bs, float* dq, float* dv0, float* dv1, float* da0, float* da1, int N, int C0,
            int C1, int S, int dtype, cudaStream_t st) {
  using OT = typename std::conditional<TF32, float, __nv_bfloat16>::type;
  const int Cv = C0 + C1;
  OT* vcat = (OT*)(ws + pl.vcat);
  OT* dop = (OT*)(ws + pl.dop);
  float* delta = (float*)(ws + pl.delta);
  float* rvec = (float*)(ws + pl.rvec);
  OT *qa = (OT*)(ws + pl.qa), *qb = TF32 ? (OT*)(ws + pl.qb) : qa, *qn = (OT*)(ws + pl.qn), *qr = (OT*)(ws + pl.qr);
  OT *dot = (OT*)(ws + pl.dot), *vt = (OT*)(ws + pl.vt), *PT = (OT*)(ws + pl.pt), *dS = (OT*)(ws + pl.ds);
  float* P = (float*)(ws + pl.p);   // fp32 P; the buffer is reused for G (operand type) once dE exists
  OT* G = (OT*)(ws + pl.p);
  const __nv_bfloat16* qt = (const __nv_bfloat16*)(ws + pl.qt);
  const int64_t SS = (int64_t)S * S;
  int rc;

  {
    dim3 grid((S + 255) / 256, N);
    attn_bwd_prep_kernel<T, OT, TF32><<<grid, 256, 0, st>>>((const T*)dout0, dout0_bs, (const T*)dout1, dout1_bs,
                                                             (const T*)o_saved, mask, a0, a1, masked0, masked1, dop, delta,
                                                             da0, da1, C0, C1, S);
    if ((rc = fmi_launched("attn_bwd_prep"))) return rc;
  }
  const int rowlen = pl.dpad * (1 + pl.split);
  for (int n0 = 0; n0 < N; n0 += pl.nb) {
    const int nb = N - n0 < pl.nb ? N - n0 : pl.nb;
    {
      dim3 qg((unsigned)imin64(((int64_t)S * pl.dpad + 255) / 256, (int64_t)FMI_NUM_SMS * 8), nb);
      q_operands_kernel<OT, TF32><<<qg, 256, 0, st>>>(qt + (int64_t)n0 * S * rowlen, qa, qb, qn, qr, S, pl.dpad, pl.split);
      if ((rc = fmi_launched("q_operands"))) return rc;
      dim3 tg((S + 31) / 32, (Cv + 31) / 32, nb);
      transpose_kernel<OT><<<tg, 256, 0, st>>>(dop + (int64_t)n0 * Cv * S, S, (int64_t)Cv * S, dot, Cv, (int64_t)S * Cv, Cv, S);
      if ((rc = fmi_launched("transpose"))) return rc;
      transpose_kernel<OT><<<tg, 256, 0, st>>>(vcat + (int64_t)n0 * Cv * S, S, (int64_t)Cv * S, vt, Cv, (int64_t)S * Cv, Cv, S);
      if ((rc = fmi_launched("transpose"))) return rc;
    }
    GemmParams g{};
    // 1: P (fp32), PT (operand type)
    g.M = S; g.N = S; g.K = pl.kq; g.epi = EPI_EXP_SYM; g.out0 = P; g.out1 = PT; g.ldo = S; g.out_bs = SS;
    g.rowvec = lse + (int64_t)n0 * S; g.vec_bs = S;
    if ((rc = launch_gemm_nt<TF32>(qa, pl.kq, (int64_t)S * pl.kq, qb, pl.kq, (int64_t)S * pl.kq, nb, g, st))) return rc;
    // 2a: delta_i = sum_j P[i,j] dP[i,j] from exactly the P and dP that 2b uses (the algebraically equal
    //     sum_c dO'[c,i] O[c,i] from the forward differs by the forward's roundings, and for peaked attention
    //     dP - delta cancels to that difference — measured 3x gradient error)
    float* delta_g = delta + (int64_t)n0 * S;
    FMI_CUDA(cudaMemsetAsync(delta_g, 0, (size_t)nb * S * sizeof(float), st));
    g = GemmParams{};
    g.M = S; g.N = S; g.K = Cv; g.epi = EPI_ROWDOT; g.out0 = delta_g; g.ldo = S; g.out_bs = S; g.aux = P; g.ld_aux = S;
    g.aux_bs = SS;
    if ((rc = launch_gemm_nt<TF32>(dot, Cv, (int64_t)S * Cv, vt, Cv, (int64_t)S * Cv, nb, g, st))) return rc;
    // 2b: dE = P o (dP - delta)
    g = GemmParams{};
    g.M = S; g.N = S; g.K = Cv; g.epi = EPI_DS; g.out0 = dS; g.ldo = S; g.out_bs = SS; g.rowvec = delta_g; g.vec_bs = S;
    g.aux = P; g.ld_aux = S; g.aux_bs = SS;
    if ((rc = launch_gemm_nt<TF32>(dot, Cv, (int64_t)S * Cv, vt, Cv, (int64_t)S * Cv, nb, g, st))) return rc;
    // 3: dV per value group (+ r_j * dOut_g[c, j])
    for (int grp = 0; grp < (C1 ? 2 : 1); ++grp) {
      const int Cg = grp ? C1 : C0, cofs = grp ? C0 : 0;
      const bool masked = grp ? masked1 : masked0;
      const float bconst = grp ? b1 : b0;
      float* dv = grp ? dv1 : dv0;
      if (!dv) continue;
      g = GemmParams{};
      g.M = Cg; g.N = S; g.K = S; g.out0 = dv + (int64_t)n0 * Cg * S; g.ldo = S; g.out_bs = (int64_t)Cg * S;
      if (masked || bconst != 0.f) {
        g.epi = EPI_ADD_COLSCALE;
        if (masked) {
          g.colvec = mask + (int64_t)n0 * S;
          g.vec_bs = S;
        } else {
          fill_kernel<<<(S + 255) / 256, 256, 0, st>>>(rvec, bconst, S);
          if ((rc = fmi_launched("fill"))) return rc;
          g.colvec = rvec;
          g.vec_bs = 0;
        }
        const int64_t dbs = grp ? dout1_bs : dout0_bs;
        g.aux = (const uint8_t*)(grp ? dout1 : dout0) + (int64_t)n0 * dbs * sizeof(T);
        g.ld_aux = S;
        g.aux_bs = dbs;
        g.aux_dtype = dtype;
      } else {
        g.epi = EPI_STORE_F32;
      }
      if ((rc = launch_gemm_nt<TF32>(dop + ((int64_t)n0 * Cv + cofs) * S, S, (int64_t)Cv * S, PT, S, SS, nb, g, st))) return rc;
    }
    // 4: dq[t,:] = sum_j (dE[t,j] + dE[j,t]) q[j,:]  (keys == queries: q plays the row and the column role).
    // The row-role term is a gemm_nt on dE as it lies; the column-role term sum_i dE[i,t] q[i,:] contracts over the ROW index of
    // dE — the pixel-contraction GEMM of the convolution weight gradient (fmi_conv_wgrad_nhwc: "pixels" = i, "output channels" =
    // t, "input channels" = the head dimension), accumulated onto dq with red.add. This replaces the explicit G = dE + dE^T
    // (sym_add_kernel: 2 GB read + 1 GB written per 128^2 image, 4.7 ms per PICNet GAN step). S must be a power of two for the
    // pixel tiling; other sizes keep the transpose-add.
    static const bool sym_env = [] { const char* e = getenv("FMI_ATTN_BWD_SYM"); return e && e[0] == '1'; }();
    const bool pix = !sym_env && (S & (S - 1)) == 0 && S >= 1024;
    if (dq && pix) {
      g = GemmParams{};
      g.M = S; g.N = pl.dpad; g.K = S; g.epi = EPI_STORE_F32; g.out0 = dq + (int64_t)n0 * S * pl.dpad; g.ldo = pl.dpad;
      g.out_bs = (int64_t)S * pl.dpad;
      if ((rc = launch_gemm_nt<TF32>(dS, S, SS, qn, S, (int64_t)pl.dpad * S, nb, g, st))) return rc;
      for (int b = 0; b < nb; ++b) {
        rc = fmi_conv_wgrad_nhwc(qr + (int64_t)b * S * pl.dpad, dS + (int64_t)b * SS, dq + (int64_t)(n0 + b) * S * pl.dpad, 1, pl.dpad, S,
                                 S / 128, 128, 1, 0, TF32 ? FMI_MMA_TF32 : FMI_MMA_BF16, (void*)st);
        if (rc) return rc;
      }
    } else if (dq) {
      dim3 sg(S / 32, S / 32, nb);
      sym_add_kernel<OT, TF32><<<sg, 256, 0, st>>>(dS, G, S, SS, P_stride_elems<OT>(SS));
      if ((rc = fmi_launched("sym_add"))) return rc;
      g = GemmParams{};
      g.M = S; g.N = pl.dpad; g.K = S; g.epi = EPI_STORE_F32; g.out0 = dq + (int64_t)n0 * S * pl.dpad; g.ldo = pl.dpad;
      g.out_bs = (int64_t)S * pl.dpad;
      if ((rc = launch_gemm_nt<TF32>(G, S, P_stride_elems<OT>(SS), qn, S, (int64_t)pl.dpad * S, nb, g, st))) return rc;
    }
  }
  return FMI_OK;
}

}  // namespace

extern "C" int64_t fmi_attn_bwd_workspace_bytes(int N, int C, int d, int C0, int C1, int S, int mma) {
  (void)C;
  BwdPlan pl;
  if (make_bwd_plan(N, d, C0, C1, S, mma, &pl)) return -1;
  return pl.total;
}

extern "C" int fmi_attn_bwd(const void* x, const float* wq, const float* bq, const void* v0, const void* v1,
                            const float* mask, const float* a0, float b0, int masked0, const float* a1, float b1,
                            int masked1, const void* o_saved, const float* lse, const void* dout0, int64_t dout0_bs,
                            const void* dout1, int64_t dout1_bs, float* dq, float* dv0, float* dv1, float* da0, float* da1,
                            int N, int C, int d, int C0, int C1, int S, int dtype, int mma, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  FMI_REQUIRE(dtype == FMI_F32 || dtype == FMI_BF16, "attn_bwd: dtype must be fp32 or bf16");
  if (N == 0) return FMI_OK;
  BwdPlan pl;
  int rc = make_bwd_plan(N, d, C0, C1, S, mma, &pl);
  if (rc) return rc;
  FMI_REQUIRE(x && wq && v0 && o_saved && lse && dout0 && workspace, "attn_bwd: null pointer");
  FMI_REQUIRE((C1 == 0) == (v1 == nullptr) && (C1 == 0 || dout1), "attn_bwd: v1/dout1 must be given exactly when C1 > 0");
  FMI_REQUIRE(!(masked0 || masked1) || mask, "attn_bwd: masked group without a mask");
  FMI_REQUIRE(workspace_bytes >= pl.total && fmi_aligned(workspace, 1024), "attn_bwd: workspace too small or misaligned");
  rc = fmi_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // whole backward of one attention call (all of its launches): 5 S x S products per value group + the row-role / column-role dQ
  FmiProfScope prof(FMI_PROF_ATTN_BWD, st, (double)N * (4.0 * S * (double)S * d + 4.0 * S * (double)S * (C0 + C1)) * 1.0,
                    (double)N * S * (d + 3.0 * (C0 + C1)) * (dtype == FMI_F32 ? 4.0 : 2.0));
  uint8_t* ws = (uint8_t*)workspace;
  rc = fmi_attn_stage_operands(x, wq, bq, v0, v1, ws + pl.qt, ws + pl.vcat, N, C, d, C0, C1, S, dtype, mma, st);
  if (rc) return rc;
  const bool tf32 = mma == FMI_MMA_TF32;
#define FMI_RUN_BWD(T, TF)                                                                                              \
  run_bwd<T, TF>(pl, ws, mask, a0, b0, masked0, a1, b1, masked1, o_saved, lse, dout0, dout0_bs, dout1, dout1_bs, dq, dv0, \
                 dv1, da0, da1, N, C0, C1, S, dtype, st)
  if (dtype == FMI_F32) return tf32 ? FMI_RUN_BWD(float, true) : FMI_RUN_BWD(float, false);
  return tf32 ? FMI_RUN_BWD(__nv_bfloat16, true) : FMI_RUN_BWD(__nv_bfloat16, false);
#undef FMI_RUN_BWD
}

// Batched C[b] = A[b] B[b]^T (fp32 out) on the tcgen05 GEMM above, as a C entry: the loss-side S x S and Gram products (SURVEY 8f
// rank 3: contextual_loss' cosine similarities, GramMatrix — torch.bmm in external_function.py:180-185,249-252, fp32 SIMT GEMMs
// in the reference since matmul TF32 is off by default). A [batch, M, K], B [batch, N, K], K contiguous, in the operand type.
extern "C" int fmi_gemm_nt(const void* A, int64_t lda, int64_t a_bs, const void* B, int64_t ldb, int64_t b_bs, float* C,
                           int64_t ldc, int64_t c_bs, int batch, int M, int N, int K, int accumulate, int mma, void* stream) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "gemm_nt: bad mma");
  if (batch == 0 || M == 0 || N == 0) return FMI_OK;
  FMI_REQUIRE(A && B && C && ldc >= N && ldc % 4 == 0 && fmi_aligned(C, 16), "gemm_nt: bad output");
  int rc = fmi_device_check();
  if (rc) return rc;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.epi = EPI_STORE_F32;
  p.accumulate = accumulate;
  p.out0 = C; p.ldo = ldc; p.out_bs = c_bs;
  cudaStream_t st = (cudaStream_t)stream;
  return mma == FMI_MMA_TF32 ? launch_gemm_nt<true>(A, lda, a_bs, B, ldb, b_bs, batch, p, st)
                             : launch_gemm_nt<false>(A, lda, a_bs, B, ldb, b_bs, batch, p, st);
}
