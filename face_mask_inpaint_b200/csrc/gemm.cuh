// gemm_nt: batched C[b] = A[b] (M x K) * B[b]^T (N x K), both operands K-major in the tensor-core operand type
// (bf16, or tf32-rounded fp32), fp32 accumulation in TMEM, on tcgen05 fed by TMA (SWIZZLE_128B) — the building block
// of the attention backward pass (and, next, the modulated-conv weight gradient). One 128 x n_tile output tile per
// CTA; 192 threads: TMA warp, MMA warp, four epilogue warps (thread = output row).
//
// Epilogues
//   EPI_STORE_F32   out0[r,c] = acc (+ out0[r,c] if accumulate)                                   fp32
//   EPI_STORE_OP    out0[r,c] = acc                                                                operand type
//   EPI_EXP_SYM     out0[r,c] = exp(acc - rowvec[r]) (fp32, unrounded) ; out1[r,c] = exp(acc - rowvec[c]) (operand type)
//                   (P and P^T of the attention from the symmetric logits E = Q^T Q and the saved row lse)
//   EPI_DS          out0[r,c] = aux[r,c] * (acc - rowvec[r])     (dE = P o (dP - delta)), aux = P in fp32
//   EPI_ADD_COLSCALE out0[r,c] = acc + colvec[c] * aux_f[r,c]    fp32 (dV = dO'.P + r_i * dOut), aux_f in `aux_dtype`
#pragma once
#include "common.cuh"
#include "sm100.cuh"

namespace fmi_gemm {

using namespace sm100;

enum { EPI_STORE_F32 = 0, EPI_STORE_OP = 1, EPI_EXP_SYM = 2, EPI_DS = 3, EPI_ADD_COLSCALE = 4, EPI_ROWDOT = 5 };

struct GemmParams {
  int M, N, K, n_tile, stages, epi, accumulate, aux_dtype;
  void* out0;
  void* out1;
  int64_t ldo, out_bs;        // leading dimension (elements) and batch stride of out0/out1
  const float* rowvec;        // [batch, M]  (EPI_EXP_SYM: also indexed by column, so M == N there)
  const float* colvec;        // [batch, N]
  int64_t vec_bs;
  const void* aux;            // [batch, M, ld_aux]
  int64_t ld_aux, aux_bs;
};

constexpr int kGemmThreads = 192;
constexpr int A_TILE_BYTES = 128 * 128;

// 32 consecutive floats as 4 independent 32-byte loads (src 32-byte aligned) or 8 16-byte loads. An epilogue thread owns one
// row, so every access of a warp touches 32 different lines: 32 bytes per lane moves whole sectors (sm100.cuh, st_global_v8).
__device__ __forceinline__ void ld32f(const float* __restrict__ src, float (&dst)[32]) {
  if ((reinterpret_cast<uintptr_t>(src) & 31) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t t[8];
      sm100::ld_global_v8(src + 8 * q, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[8 * q + e] = __uint_as_float(t[e]);
    }
    return;
  }
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = s4[q];
    dst[4 * q] = t.x; dst[4 * q + 1] = t.y; dst[4 * q + 2] = t.z; dst[4 * q + 3] = t.w;
  }
}

// ncols (a multiple of 4, <= 32) floats of x to dst (16-byte aligned): 32-byte stores when dst and ncols allow
template <bool ROUND_TF32>
__device__ __forceinline__ void st32f(float* __restrict__ dst, const float (&x)[32], int ncols) {
  if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0 && (ncols & 7) == 0) {
#pragma unroll
    for (int k = 0; k < 32; k += 8)
      if (k < ncols) {
        uint32_t u[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] = ROUND_TF32 ? sm100::f32_to_tf32_rna(x[k + e]) : __float_as_uint(x[k + e]);
        sm100::st_global_v8(dst + k, u);
      }
    return;
  }
#pragma unroll
  for (int k = 0; k < 32; k += 4)
    if (k < ncols) {
      if (ROUND_TF32)
        *reinterpret_cast<float4*>(dst + k) = make_float4(__uint_as_float(sm100::f32_to_tf32_rna(x[k])), __uint_as_float(sm100::f32_to_tf32_rna(x[k + 1])),
                                                          __uint_as_float(sm100::f32_to_tf32_rna(x[k + 2])), __uint_as_float(sm100::f32_to_tf32_rna(x[k + 3])));
      else
        *reinterpret_cast<float4*>(dst + k) = make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
    }
}

template <bool TF32>
__global__ void __launch_bounds__(kGemmThreads, 2)
    gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const GemmParams p) {
  constexpr int EPA = TF32 ? 32 : 64;
  using OT = typename std::conditional<TF32, float, __nv_bfloat16>::type;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int b_tile_bytes = p.n_tile * 128;
  const int stage_bytes = A_TILE_BYTES + b_tile_bytes;
  __shared__ uint64_t full[8], empty[8], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.x * p.n_tile, m0 = blockIdx.y * 128, bz = blockIdx.z;
  const int iters = (p.K + EPA - 1) / EPA;

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 128);  // n_tile <= 128 accumulator columns; leaves room for 4 CTAs per SM
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_b);
      for (int it = 0; it < iters; ++it) {
        const int st = it % p.stages;
        mbar_wait(&empty[st], ((it / p.stages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[st], (uint32_t)stage_bytes);
        uint8_t* sA = smem + st * stage_bytes;
        tma_load_3d(sA, &map_a, &full[st], it * EPA, m0, bz);
        tma_load_3d(sA + A_TILE_BYTES, &map_b, &full[st], it * EPA, n0, bz);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, 128, p.n_tile);
      for (int it = 0; it < iters; ++it) {
        const int st = it % p.stages;
        mbar_wait(&full[st], (it / p.stages) & 1);
        tc_fence_after();
        uint8_t* sA = smem + st * stage_bytes;
        const uint64_t adesc = make_sdesc_k_sw128(smem_u32(sA));
        const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(sA + A_TILE_BYTES));
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint32_t acc = (it > 0 || s > 0) ? 1u : 0u;
          if (TF32) mma_ss_tf32(tmem, adesc + 2 * s, bdesc + 2 * s, idesc, acc);
          else mma_ss_f16(tmem, adesc + 2 * s, bdesc + 2 * s, idesc, acc);
        }
        tc_commit(&empty[st]);
      }
      tc_commit(&acc_full);
    }
    __syncwarp();
  } else {
    const int lane_base = (warp & 3) * 32;
    const int r = m0 + lane_base + (tid & 31);
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const bool row_ok = r < p.M;
    const float rv = (p.rowvec && row_ok) ? p.rowvec[(int64_t)bz * p.vec_bs + r] : 0.f;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_addr + c0, v);
      tc_wait_ld();
      const int cbase = n0 + c0;
      if (!row_ok || cbase >= p.N) continue;
      const int ncols = min(32, p.N - cbase);  // N is a multiple of 4 (host-checked): whole 16-byte groups
      float f[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
      const int64_t off = (int64_t)bz * p.out_bs + (int64_t)r * p.ldo + cbase;
      if (p.epi == EPI_ROWDOT) {
        // out0[r] += sum_c aux[r,c] * acc[r,c]   (delta_i = sum_j P[i,j] dP[i,j] from the SAME P and dP as EPI_DS uses)
        const float* pa = (const float*)p.aux + (int64_t)bz * p.aux_bs + (int64_t)r * p.ld_aux + cbase;
        float part = 0.f;
        if (ncols == 32 && fmi_aligned_dev(pa, 16)) {   // 8 independent 16-byte loads (a guarded scalar load per element
          float a[32];                                    // compiles to a branch per element: 32 serialised load latencies)
          ld32f(pa, a);
#pragma unroll
          for (int k = 0; k < 32; ++k) part = fmaf(a[k], f[k], part);
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k < ncols) part = fmaf(pa[k], f[k], part);
        }
        atomicAdd((float*)p.out0 + (int64_t)bz * p.out_bs + r, part);
        continue;
      }
      if (p.epi == EPI_STORE_F32 || p.epi == EPI_ADD_COLSCALE) {
        float* o = (float*)p.out0 + off;
        if (p.epi == EPI_ADD_COLSCALE) {
          const float* cv = p.colvec + (int64_t)bz * p.vec_bs + cbase;
          const int64_t aoff = (int64_t)bz * p.aux_bs + (int64_t)r * p.ld_aux + cbase;
          if (ncols == 32 && p.aux_dtype == FMI_F32 && fmi_aligned_dev((const float*)p.aux + aoff, 16) && fmi_aligned_dev(cv, 16)) {
            float a[32], c[32];
            ld32f((const float*)p.aux + aoff, a);
            ld32f(cv, c);
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] = fmaf(c[k], a[k], f[k]);
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < ncols) {
                const float a = p.aux_dtype == FMI_F32 ? ((const float*)p.aux)[aoff + k]
                                                       : __bfloat162float(((const __nv_bfloat16*)p.aux)[aoff + k]);
                f[k] = fmaf(cv[k], a, f[k]);
              }
          }
        }
        if (p.accumulate) {
          if (ncols == 32) {
            float old[32];
            ld32f(o, old);
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] += old[k];
          } else {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              if (k < ncols) {
                const float4 old = *reinterpret_cast<const float4*>(o + k);
                f[k] += old.x; f[k + 1] += old.y; f[k + 2] += old.z; f[k + 3] += old.w;
              }
          }
        }
        st32f<false>(o, f, ncols);
      } else {
        OT* o0 = (OT*)p.out0 + off;
        float g[32];
        if (p.epi == EPI_EXP_SYM) {
          const float* lc = p.rowvec + (int64_t)bz * p.vec_bs + cbase;
          if (ncols == 32 && fmi_aligned_dev(lc, 16)) {
            float l[32];
            ld32f(lc, l);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              g[k] = __expf(f[k] - l[k]);
              f[k] = __expf(f[k] - rv);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              g[k] = k < ncols ? __expf(f[k] - lc[k]) : 0.f;
              f[k] = __expf(f[k] - rv);
            }
          }
          // P (out0) is only ever an elementwise factor of dE: keep it UNROUNDED fp32 — rounding it before forming
          // P o (dP - delta) is what dominated the gradient error for peaked attention; P^T (out1) is a GEMM operand.
          float* pf = (float*)p.out0 + off;
          st32f<false>(pf, f, ncols);
        } else if (p.epi == EPI_DS) {
          const float* pa = (const float*)p.aux + (int64_t)bz * p.aux_bs + (int64_t)r * p.ld_aux + cbase;
          if (ncols == 32 && fmi_aligned_dev(pa, 16)) {
            float a[32];
            ld32f(pa, a);
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] = a[k] * (f[k] - rv);
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] = k < ncols ? pa[k] * (f[k] - rv) : 0.f;
          }
        }
        auto store_row = [&](OT* dst, const float (&x)[32]) {
          if constexpr (TF32) {
            st32f<true>(reinterpret_cast<float*>(dst), x, ncols);
          } else if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0 && (ncols & 15) == 0) {
#pragma unroll
            for (int k = 0; k < 32; k += 16)
              if (k < ncols) {
                uint32_t u[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) u[e] = pack_bf16x2(x[k + 2 * e], x[k + 2 * e + 1]);
                st_global_v8(dst + k, u);
              }
          } else {
#pragma unroll
            for (int k = 0; k < 32; k += 8)
              if (k < ncols) {
                uint4 u;
                u.x = pack_bf16x2(x[k], x[k + 1]);
                u.y = pack_bf16x2(x[k + 2], x[k + 3]);
                u.z = pack_bf16x2(x[k + 4], x[k + 5]);
                u.w = pack_bf16x2(x[k + 6], x[k + 7]);
                *reinterpret_cast<uint4*>(dst + k) = u;
              }
          }
        };
        if (p.epi == EPI_EXP_SYM) store_row((OT*)p.out1 + off, g);
        else store_row(o0, f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

// Host launcher. A: [batch, M, K] with row pitch lda, B: [batch, N, K] with row pitch ldb (elements of the operand type).
template <bool TF32>
int launch_gemm_nt(const void* A, int64_t lda, int64_t a_bs, const void* B, int64_t ldb, int64_t b_bs, int batch,
                   GemmParams p, cudaStream_t st) {
  constexpr int esz = TF32 ? 4 : 2;
  constexpr uint32_t epa = 128 / esz;
  FMI_REQUIRE(p.M >= 1 && p.N >= 1 && p.K >= 1 && batch >= 1, "gemm_nt: bad shape");
  FMI_REQUIRE(p.N % 8 == 0, "gemm_nt: N=%d must be a multiple of 8", p.N);
  FMI_REQUIRE((lda * esz) % 16 == 0 && (ldb * esz) % 16 == 0 && fmi_aligned(A, 16) && fmi_aligned(B, 16),
              "gemm_nt: operands must have 16-byte aligned rows");
  int n_tile = p.N >= 128 ? 128 : (p.N + 15) / 16 * 16;
  p.n_tile = n_tile;
  const int stage_bytes = A_TILE_BYTES + n_tile * 128;
  int stages = (232448 - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  // Short-K products (the S x S maps of the attention backward, K = 64..512) are latency-bound per CTA (ncu: tensor pipe
  // 5 %, one CTA per SM): a 3-deep ring keeps the footprint under 100 KB so two CTAs share an SM and one's epilogue
  // overlaps the other's loads.
  const int iters = (p.K + (int)epa - 1) / (int)epa;
  if (iters <= 16 && stages > 3) stages = 3;
  p.stages = stages;
  const CUtensorMapDataType dt = TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ma, mb;
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.M, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)lda * esz, (uint64_t)(batch > 1 ? a_bs : (int64_t)p.M * lda) * esz};
    uint32_t box[3] = {epa, 128, 1};
    int e = make_tensor_map(&ma, dt, 3, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "gemm_nt: cuTensorMapEncodeTiled(A) failed (%d)", e);
  }
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.N, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)ldb * esz, (uint64_t)(batch > 1 ? b_bs : (int64_t)p.N * ldb) * esz};
    uint32_t box[3] = {epa, (uint32_t)n_tile, 1};
    int e = make_tensor_map(&mb, dt, 3, B, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "gemm_nt: cuTensorMapEncodeTiled(B) failed (%d)", e);
  }
  auto kern = gemm_nt_kernel<TF32>;
  static FmiPerDeviceOnce attr_once;
  if (attr_once.need()) {
    FMI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    attr_once.done();
  }
  dim3 grid((p.N + n_tile - 1) / n_tile, (p.M + 127) / 128, batch);
  kern<<<grid, kGemmThreads, (size_t)stages * stage_bytes, st>>>(ma, mb, p);
  return fmi_launched("gemm_nt");
}

}  // namespace fmi_gemm
