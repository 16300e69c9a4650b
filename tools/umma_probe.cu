// Standalone probe of the sm_100a primitives in csrc/sm100.cuh: one 128 x N x K tcgen05 GEMM per test,
// checked against a host reference. Validates (without the rest of the kernels in the way):
//   A sources : 2-D TMA tile (SWIZZLE_128B, K-major) | tensor memory (tcgen05.st, "TS" MMA) |
//               4-D TMA im2col window over an NHWC tensor with negative start coordinates (zero fill)
//   kinds     : bf16 (kind::f16), tf32 (kind::tf32)
//   K         : several 128-byte swizzle atoms (descriptor stepping inside and across atoms)
// Usage: umma_probe <test>   (each test in its own process: a trap kills the context)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include tools/umma_probe.cu -o build/umma_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cuda_bf16.h>

#include "../face_mask_inpaint_b200/csrc/sm100.cuh"

using namespace sm100;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

enum { A_TMA2D = 0, A_TMEM = 1, A_IM2COL = 2 };

// One CTA, 128 threads. TF32: element = 4 bytes (32 per atom row); BF16: 2 bytes (64 per atom row).
template <bool TF32, int AMODE>
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap mapA,
                                                    const __grid_constant__ CUtensorMap mapB, const float* __restrict__ Araw,
                                                    float* __restrict__ D, int N, int katoms) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int EPA = TF32 ? 32 : 64;  // elements per 128-byte atom row
  uint8_t* sA = smem;                  // katoms x [128 rows x 128 B]
  uint8_t* sB = smem + katoms * 128 * 128;
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_D = tmem;        // columns [0, N)
  const uint32_t tmem_A = tmem + 256;  // columns [256, ...)

  if (tid == 0) {
    uint32_t bytes = katoms * N * 128;
    if (AMODE != A_TMEM) bytes += katoms * 128 * 128;
    mbar_arrive_expect_tx(&bar_load, bytes);
    for (int a = 0; a < katoms; ++a) {
      if (AMODE == A_TMA2D) tma_load_2d(sA + a * 128 * 128, &mapA, &bar_load, a * EPA, 0);
      if (AMODE == A_IM2COL) tma_load_4d(sA + a * 128 * 128, &mapA, &bar_load, a * EPA, -1, -1, 0);
      tma_load_2d(sB + a * N * 128, &mapB, &bar_load, a * EPA, 0);
    }
  }
  if (AMODE == A_TMEM) {
    // thread t owns row t: write A[t, :] into tensor memory
    const int K = katoms * EPA;
    const float* arow = Araw + (size_t)tid * K;
    const uint32_t lane_addr = tmem_A + ((uint32_t)(warp * 32) << 16);
    if (TF32) {
      for (int c0 = 0; c0 < K; c0 += 32) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = f32_to_tf32_rna(arow[c0 + j]);
        tmem_st32(lane_addr + c0, v);
      }
    } else {
      for (int c0 = 0; c0 < K / 2; c0 += 32) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = pack_bf16x2(arow[2 * (c0 + j)], arow[2 * (c0 + j) + 1]);
        tmem_st32(lane_addr + c0, v);
      }
    }
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, 128, N);
    uint32_t acc = 0;
    for (int a = 0; a < katoms; ++a) {
      const uint64_t adesc = make_sdesc_k_sw128(smem_u32(sA + a * 128 * 128));
      const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(sB + a * N * 128));
      for (int s = 0; s < 4; ++s) {  // 4 MMAs of 32 bytes of K per 128-byte atom row
        const uint64_t koff = (uint64_t)((s * 32) >> 4);
        if (AMODE == A_TMEM) {
          const uint32_t a_t = tmem_A + (a * 4 + s) * 8;
          if (TF32) mma_ts_tf32(tmem_D, a_t, bdesc + koff, idesc, acc);
          else mma_ts_f16(tmem_D, a_t, bdesc + koff, idesc, acc);
        } else {
          if (TF32) mma_ss_tf32(tmem_D, adesc + koff, bdesc + koff, idesc, acc);
          else mma_ss_f16(tmem_D, adesc + koff, bdesc + koff, idesc, acc);
        }
        acc = 1;
      }
    }
    tc_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_D + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float tf32_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;  // round-to-nearest (ties away), matches cvt.rna
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

template <bool TF32, int AMODE>
int run(const char* name, int N, int katoms) {
  const int EPA = TF32 ? 32 : 64;
  const int K = katoms * EPA;
  const int esz = TF32 ? 4 : 2;
  // im2col source: NHWC [1,16,16,K]; A rows are the 8x16 window starting at (-1,-1)
  const int IH = 16, IW = 16;
  std::vector<float> A((size_t)128 * K), B((size_t)N * K), X((size_t)IH * IW * K);
  srand(1234);
  auto rnd = [] { return (float)(rand() % 2001 - 1000) / 1000.f; };
  for (auto& v : X) v = TF32 ? tf32_round(rnd()) : bf16_round(rnd());
  for (auto& v : B) v = TF32 ? tf32_round(rnd()) : bf16_round(rnd());
  if (AMODE == A_IM2COL) {
    for (int r = 0; r < 128; ++r) {
      int h = r / 16 - 1, w = r % 16 - 1;
      for (int k = 0; k < K; ++k) A[(size_t)r * K + k] = (h < 0 || w < 0) ? 0.f : X[((size_t)h * IW + w) * K + k];
    }
  } else {
    for (auto& v : A) v = TF32 ? tf32_round(rnd()) : bf16_round(rnd());
  }
  std::vector<float> ref((size_t)128 * N);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      ref[(size_t)m * N + n] = (float)s;
    }
  // device buffers in the operand element type
  auto upload = [&](const std::vector<float>& src) -> void* {
    void* d;
    if (TF32) {
      CK(cudaMalloc(&d, src.size() * 4));
      CK(cudaMemcpy(d, src.data(), src.size() * 4, cudaMemcpyHostToDevice));
    } else {
      std::vector<__nv_bfloat16> h(src.size());
      for (size_t i = 0; i < src.size(); ++i) h[i] = __float2bfloat16_rn(src[i]);
      CK(cudaMalloc(&d, src.size() * 2));
      CK(cudaMemcpy(d, h.data(), src.size() * 2, cudaMemcpyHostToDevice));
    }
    return d;
  };
  void* dA = upload(AMODE == A_IM2COL ? X : A);
  void* dB = upload(B);
  float* dAraw;
  CK(cudaMalloc(&dAraw, A.size() * 4));
  CK(cudaMemcpy(dAraw, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  float* dD;
  CK(cudaMalloc(&dD, ref.size() * 4));
  CK(cudaMemset(dD, 0xFF, ref.size() * 4));

  CUtensorMapDataType dt = TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap mA, mB;
  int rc;
  if (AMODE == A_IM2COL) {
    uint64_t dims[4] = {(uint64_t)K, (uint64_t)IW, (uint64_t)IH, 1};
    uint64_t str[3] = {(uint64_t)K * esz, (uint64_t)IW * K * esz, (uint64_t)IH * IW * K * esz};
    uint32_t box[4] = {(uint32_t)EPA, 16, 8, 1};
    rc = make_tensor_map(&mA, dt, 4, dA, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    uint64_t dims[2] = {(uint64_t)K, 128};
    uint64_t str[1] = {(uint64_t)K * esz};
    uint32_t box[2] = {(uint32_t)EPA, 128};
    rc = make_tensor_map(&mA, dt, 2, dA, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) { printf("%s: tensor map A failed rc=%d\n", name, rc); return 1; }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * esz};
    uint32_t box[2] = {(uint32_t)EPA, (uint32_t)N};
    rc = make_tensor_map(&mB, dt, 2, dB, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) { printf("%s: tensor map B failed rc=%d\n", name, rc); return 1; }

  size_t smem = (size_t)katoms * (128 + N) * 128 + 1024;
  auto kern = probe_kernel<TF32, AMODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<1, 128, smem>>>(mA, mB, dAraw, dD, N, katoms);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: KERNEL FAILED: %s\n", name, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> out(ref.size());
  CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  int bad_m = -1, bad_n = -1;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double d = fabs((double)out[(size_t)m * N + n] - ref[(size_t)m * N + n]);
      if (!(d <= maxerr)) { maxerr = d; bad_m = m; bad_n = n; }
      maxref = fmax(maxref, fabs((double)ref[(size_t)m * N + n]));
    }
  bool pass = maxerr <= 1e-3 * maxref;
  printf("%s: N=%d K=%d max_abs_err=%.3e max_ref=%.3e worst=(%d,%d) got=%.5f want=%.5f -> %s\n", name, N, K, maxerr, maxref,
         bad_m, bad_n, out[(size_t)bad_m * N + bad_n], ref[(size_t)bad_m * N + bad_n], pass ? "PASS" : "FAIL");
  if (!pass) {
    printf("  row0 got : ");
    for (int n = 0; n < 8; ++n) printf("%9.4f ", out[n]);
    printf("\n  row0 want: ");
    for (int n = 0; n < 8; ++n) printf("%9.4f ", ref[n]);
    printf("\n  row1 got : ");
    for (int n = 0; n < 8; ++n) printf("%9.4f ", out[N + n]);
    printf("\n  row1 want: ");
    for (int n = 0; n < 8; ++n) printf("%9.4f ", ref[N + n]);
    printf("\n");
  }
  return pass ? 0 : 1;
}

// ---- halo test: ONE TMA box of 130 consecutive pixels (NHWC row segment starting at x = -1, zero filled) serves the three
// horizontal taps of a 3x3 convolution: the A descriptor of tap dx simply starts (dx + 1) pixel rows = (dx + 1) * 128 bytes
// further into the SWIZZLE_128B tile. Valid iff the MMA unit applies the 128-byte swizzle to ADDRESS bits (as TMA does
// when it writes), not to row indices relative to the descriptor start.
template <bool TF32>
__global__ void __launch_bounds__(128) halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                   const __grid_constant__ CUtensorMap mapB, float* __restrict__ D, int N,
                                                   int base_off_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int EPA = TF32 ? 32 : 64;
  uint8_t* sA = smem;                 // 130 rows x 128 B (padded to 17 KB)
  uint8_t* sB = smem + 17 * 1024;     // N rows x 128 B
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_load, 130 * 128 + N * 128);
    tma_load_4d(sA, &mapA, &bar_load, 0, -1, 1, 0);   // pixels x = -1 .. 128 of image row 1
    tma_load_2d(sB, &mapB, &bar_load, 0, 0);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, 128, N);
    for (int j = 0; j < 3; ++j) {
      const uint32_t a_addr = smem_u32(sA) + j * 128;
      uint64_t adesc = make_sdesc_k_sw128(a_addr);
      if (base_off_mode) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;  // PTX "matrix base offset" field
      const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(sB));
      for (int s = 0; s < 4; ++s) {
        if (TF32) mma_ss_tf32(tmem + j * N, adesc + 2 * s, bdesc + 2 * s, idesc, s > 0);
        else mma_ss_f16(tmem + j * N, adesc + 2 * s, bdesc + 2 * s, idesc, s > 0);
      }
    }
    tc_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 3 * N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * 3 * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <bool TF32>
int run_halo(const char* name, int base_off_mode) {
  const int EPA = TF32 ? 32 : 64, K = EPA, N = 64, esz = TF32 ? 4 : 2;
  const int IH = 3, IW = 160;
  std::vector<float> X((size_t)IH * IW * K), B((size_t)N * K);
  srand(4321);
  auto rnd = [] { return (float)(rand() % 2001 - 1000) / 1000.f; };
  for (auto& v : X) v = TF32 ? tf32_round(rnd()) : bf16_round(rnd());
  for (auto& v : B) v = TF32 ? tf32_round(rnd()) : bf16_round(rnd());
  std::vector<float> ref((size_t)128 * 3 * N);
  for (int j = 0; j < 3; ++j)
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        const int x = m + j - 1;
        double s = 0;
        if (x >= 0 && x < IW)
          for (int k = 0; k < K; ++k) s += (double)X[((size_t)1 * IW + x) * K + k] * B[(size_t)n * K + k];
        ref[(size_t)m * 3 * N + j * N + n] = (float)s;
      }
  auto upload = [&](const std::vector<float>& src) -> void* {
    void* d;
    if (TF32) {
      CK(cudaMalloc(&d, src.size() * 4));
      CK(cudaMemcpy(d, src.data(), src.size() * 4, cudaMemcpyHostToDevice));
    } else {
      std::vector<__nv_bfloat16> h(src.size());
      for (size_t i = 0; i < src.size(); ++i) h[i] = __float2bfloat16_rn(src[i]);
      CK(cudaMalloc(&d, src.size() * 2));
      CK(cudaMemcpy(d, h.data(), src.size() * 2, cudaMemcpyHostToDevice));
    }
    return d;
  };
  void* dX = upload(X);
  void* dB = upload(B);
  float* dD;
  CK(cudaMalloc(&dD, ref.size() * 4));
  CK(cudaMemset(dD, 0xFF, ref.size() * 4));
  CUtensorMapDataType dt = TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap mA, mB;
  {
    uint64_t dims[4] = {(uint64_t)K, (uint64_t)IW, (uint64_t)IH, 1};
    uint64_t str[3] = {(uint64_t)K * esz, (uint64_t)IW * K * esz, (uint64_t)IH * IW * K * esz};
    uint32_t box[4] = {(uint32_t)EPA, 130, 1, 1};
    if (make_tensor_map(&mA, dt, 4, dX, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("%s: map A failed\n", name); return 1; }
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * esz};
    uint32_t box[2] = {(uint32_t)EPA, (uint32_t)N};
    if (make_tensor_map(&mB, dt, 2, dB, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("%s: map B failed\n", name); return 1; }
  }
  const size_t smem = 17 * 1024 + N * 128 + 1024;
  auto kern = halo_kernel<TF32>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<1, 128, smem>>>(mA, mB, dD, N, base_off_mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: KERNEL FAILED: %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> out(ref.size());
  CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
  int rc = 0;
  for (int j = 0; j < 3; ++j) {
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        const size_t i = (size_t)m * 3 * N + j * N + n;
        maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
        maxref = fmax(maxref, fabs((double)ref[i]));
      }
    const bool pass = maxerr <= 1e-3 * maxref;
    printf("%s (base_offset field %s): tap shift %d rows: max_abs_err=%.3e max_ref=%.3e -> %s\n", name,
           base_off_mode ? "set" : "0", j, maxerr, maxref, pass ? "PASS" : "FAIL");
    rc |= !pass;
  }
  return rc;
}

int main(int argc, char** argv) {
  int t = argc > 1 ? atoi(argv[1]) : 0;
  switch (t) {
    case 0: return run<false, A_TMA2D>("ss_bf16_k64", 128, 1);
    case 1: return run<false, A_TMA2D>("ss_bf16_k256_n256", 256, 4);
    case 2: return run<true, A_TMA2D>("ss_tf32_k64", 128, 2);
    case 3: return run<true, A_TMA2D>("ss_tf32_k128_n256", 256, 4);
    case 4: return run<false, A_TMEM>("ts_bf16_k128_n256", 256, 2);
    case 5: return run<true, A_TMEM>("ts_tf32_k128_n256", 256, 4);
    case 6: return run<false, A_IM2COL>("im2col_bf16_k128_n64", 64, 2);
    case 7: return run<false, A_TMA2D>("ss_bf16_k64_n16", 16, 1);
    case 8: return run<true, A_IM2COL>("im2col_tf32_k64_n64", 64, 2);
    case 9: return run_halo<false>("halo_bf16", 0);
    case 10: return run_halo<false>("halo_bf16", 1);
    case 11: return run_halo<true>("halo_tf32", 0);
    case 12: return run_halo<true>("halo_tf32", 1);
    default: printf("unknown test %d\n", t); return 3;
  }
}
