// Issue-rate probe for the tcgen05.mma shapes the attention kernels use: every SM runs one CTA whose elected thread issues
// `iters` back-to-back MMAs of one shape on fixed operands (no TMA, no softmax, no dependencies except the accumulator),
// then commits and waits. Reports MAC/clk/SM and the fraction of the nominal dense rate (bf16 4096, tf32 2048 MAC/clk/SM).
// This separates "the instruction shape cannot go faster" from "the kernel around it starves the tensor pipe".
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/umma_rate.cu -o build/umma_rate
//   run:   build/umma_rate
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../face_mask_inpaint_b200/csrc/sm100.cuh"

using namespace sm100;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

struct Shape {
  const char* name;
  int tf32;      // 0: kind::f16 (bf16), 1: kind::tf32
  int a_tmem;    // A operand from tensor memory (.ts) instead of shared memory
  int N;         // M = 128
  int nacc;      // number of distinct accumulators cycled through (1 = every MMA depends on the previous one's D)
  int bslots;    // number of distinct B tiles cycled through (smem footprint / bank behaviour)
  int traffic;   // what warps 1-3 do meanwhile: 0 idle, 1 tcgen05.ld + tcgen05.st loop (softmax-like TMEM traffic), 2 ld only, 3 st only
  int interleave;  // 1: consecutive MMAs go to DIFFERENT accumulators (acc = mma index % nacc) instead of 4 K-steps per accumulator
};

__global__ void __launch_bounds__(128, 1) rate_kernel(Shape sh, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) stop_flag = 0;
  // operands: A tile 128 rows x 128 B at smem[0], B tiles (N rows x 128 B each) after it; small finite values
  const int b_tile = sh.N * 128;
  for (int i = tid; i < (16384 + sh.bslots * b_tile) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // P-like A operand in TMEM columns [448, 512)
  {
    uint32_t v[32];
    for (int k = 0; k < 32; ++k) v[k] = 0x3c003c00u + k;
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 448, v);
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 480, v);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc(sh.tf32 ? KIND_TF32 : KIND_BF16, 128, sh.N);
    const uint64_t adesc = make_sdesc_k_sw128(smem_u32(smem));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(smem + 16384 + (it % sh.bslots) * b_tile));
      const uint32_t d0 = tmem + (it % sh.nacc) * sh.N;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint32_t d = sh.interleave ? tmem + ((it * 4 + s) % sh.nacc) * sh.N : d0;
        if (sh.a_tmem) {
          if (sh.tf32) mma_ts_tf32(d, tmem + 448 + s * 8, bdesc + 2 * s, idesc, 1);
          else mma_ts_f16(d, tmem + 448 + s * 8, bdesc + 2 * s, idesc, 1);
        } else {
          if (sh.tf32) mma_ss_tf32(d, adesc + 2 * s, bdesc + 2 * s, idesc, 1);
          else mma_ss_f16(d, adesc + 2 * s, bdesc + 2 * s, idesc, 1);
        }
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    *(volatile int*)&stop_flag = 1;
  } else if (warp > 0 && sh.traffic) {
    // softmax-like traffic on columns [256, 384) of this warp's lanes until the issuer is done
    uint32_t v[32];
    const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16) + 256;
    for (int k = 0; k < 32; ++k) v[k] = k;
    long long n_ops = 0;
    const long long c_start = clock64();
    while (!*(volatile int*)&stop_flag) {
      for (int c = 0; c < 128; c += 32) {
        if (sh.traffic != 3) { tmem_ld32(base + c, v); tc_wait_ld(); }
        if (sh.traffic != 2) { tmem_st32(base + c, v); tc_wait_st(); }
        ++n_ops;
      }
    }
    // average latency of one 32-column tcgen05.ld and/or .st (+ wait) of this warp while the MMAs were running
    if (warp == 1 && (tid & 31) == 0) cycles[gridDim.x + blockIdx.x] = (clock64() - c_start) / (n_ops > 0 ? n_ops : 1);
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Lean issue loop: descriptors precomputed, four MMAs per iteration fully unrolled, nothing but the MMAs in the loop body.
// `issuers` elected threads (one per warp) issue concurrently into their own accumulators. Separates "the tensor pipe needs
// this long per instruction" from "one thread cannot issue faster than this".
template <int N, bool TF32, bool ATMEM>
__global__ void __launch_bounds__(128, 1) lean_kernel(int iters, int issuers, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 2 * N * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  {
    uint32_t v[32];
    for (int k = 0; k < 32; ++k) v[k] = 0x3c003c00u + k;
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 448, v);
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 480, v);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < issuers && elect_one()) {
    const uint32_t idesc = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, 128, N);
    const uint64_t a0 = make_sdesc_k_sw128(smem_u32(smem));
    const uint64_t b0 = make_sdesc_k_sw128(smem_u32(smem + 16384 + (warp & 1) * N * 128));
    const uint32_t d = tmem + warp * N;   // N <= 64 with up to 4 issuers, N = 256 with 1
    const uint32_t at = tmem + 448;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      if (ATMEM) {
        if (TF32) { mma_ts_tf32(d, at, b0, idesc, 1); mma_ts_tf32(d, at + 8, b0 + 2, idesc, 1); mma_ts_tf32(d, at + 16, b0 + 4, idesc, 1); mma_ts_tf32(d, at + 24, b0 + 6, idesc, 1); }
        else { mma_ts_f16(d, at, b0, idesc, 1); mma_ts_f16(d, at + 8, b0 + 2, idesc, 1); mma_ts_f16(d, at + 16, b0 + 4, idesc, 1); mma_ts_f16(d, at + 24, b0 + 6, idesc, 1); }
      } else {
        if (TF32) { mma_ss_tf32(d, a0, b0, idesc, 1); mma_ss_tf32(d, a0 + 2, b0 + 2, idesc, 1); mma_ss_tf32(d, a0 + 4, b0 + 4, idesc, 1); mma_ss_tf32(d, a0 + 6, b0 + 6, idesc, 1); }
        else { mma_ss_f16(d, a0, b0, idesc, 1); mma_ss_f16(d, a0 + 2, b0 + 2, idesc, 1); mma_ss_f16(d, a0 + 4, b0 + 4, idesc, 1); mma_ss_f16(d, a0 + 6, b0 + 6, idesc, 1); }
      }
    }
    const long long t_issue = clock64() - t0;
    tc_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    if (warp == 0) {
      cycles[blockIdx.x] = clock64() - t0;
      cycles[gridDim.x + blockIdx.x] = t_issue;
    }
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, bool TF32, bool ATMEM>
void run_lean(const char* name, int issuers, int sms, long long* d_cycles) {
  const int iters = 20000;
  auto kern = lean_kernel<N, TF32, ATMEM>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const size_t smem = 1024 + 16384 + 2 * (size_t)N * 128;
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<sms, 128, smem>>>(iters, issuers, d_cycles);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(2 * sms);
  CK(cudaMemcpy(h.data(), d_cycles, 2 * sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double tot = 0, iss = 0;
  for (int i = 0; i < sms; ++i) { tot += (double)h[i]; iss += (double)h[sms + i]; }
  tot /= sms; iss /= sms;
  const double mmas = iters * 4.0 * issuers;
  printf("%-44s %10.1f clk per MMA on the SM (all issuers), issue loop alone %6.1f clk per own MMA, %d issuer(s)\n", name,
         tot / mmas, iss / (iters * 4.0), issuers);
}

// Same probe on a CTA pair (cta_group::2, M = 256): the leader issues, each CTA holds N/2 rows of B.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(Shape sh, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool leader = cluster_ctarank() == 0;
  const int b_tile = (sh.N / 2) * 128;
  for (int i = tid; i < (16384 + sh.bslots * b_tile) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 0xff);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(&tmem_base_s, 512);
    tmem_relinquish2();
  }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  {
    uint32_t v[32];
    for (int k = 0; k < 32; ++k) v[k] = 0x3c003c00u + k;
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 448, v);
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 480, v);
    tc_wait_st();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (leader && warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc(sh.tf32 ? KIND_TF32 : KIND_BF16, 256, sh.N);
    const uint64_t adesc = make_sdesc_k_sw128(smem_u32(smem));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(smem + 16384 + (it % sh.bslots) * b_tile));
      const uint32_t d = tmem + (it % sh.nacc) * sh.N;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (sh.a_tmem) {
          if (sh.tf32) mma2_ts_tf32(d, tmem + 448 + s * 8, bdesc + 2 * s, idesc, 1);
          else mma2_ts_f16(d, tmem + 448 + s * 8, bdesc + 2 * s, idesc, 1);
        } else {
          mma2_ss_f16(d, adesc + 2 * s, bdesc + 2 * s, idesc, 1);
        }
      }
    }
    tc_commit2_mc(&bar, 0x3);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x / 2] = clock64() - t0;
  } else if (!leader && warp == 0 && elect_one()) {
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 512);
}

int main() {
  const Shape shapes[] = {
      {"bf16 SS N=256 (GEMM-like), 1 acc", 0, 0, 256, 1, 1, 0, 0},
      {"bf16 TS N=256 (PV bf16), 1 acc", 0, 1, 256, 1, 4, 0, 0},
      {"tf32 TS N=256 (PV fp32 contract), 1 acc", 1, 1, 256, 1, 4, 0, 0},
      {"tf32 SS N=256, 1 acc", 1, 0, 256, 1, 4, 0, 0},
      {"bf16 SS N=64 (QK step), 4 acc", 0, 0, 64, 4, 2, 0, 0},
      {"bf16 SS N=64 (QK step), 1 acc", 0, 0, 64, 1, 2, 0, 0},
      {"bf16 SS N=128, 2 acc", 0, 0, 128, 2, 2, 0, 0},
      {"bf16 SS N=128, 1 acc", 0, 0, 128, 1, 1, 0, 0},
      {"tf32 TS N=256 + 3 warps tcgen05.ld/st", 1, 1, 256, 1, 4, 1, 0},
      {"tf32 TS N=256 + 3 warps tcgen05.ld", 1, 1, 256, 1, 4, 2, 0},
      {"tf32 TS N=256 + 3 warps tcgen05.st", 1, 1, 256, 1, 4, 3, 0},
      {"bf16 TS N=256 + 3 warps tcgen05.ld/st", 0, 1, 256, 1, 4, 1, 0},
      {"bf16 SS N=64 + 3 warps tcgen05.ld/st", 0, 0, 64, 4, 2, 1, 0},
      {"bf16 SS N=64, 4 acc INTERLEAVED per MMA", 0, 0, 64, 4, 2, 0, 1},
      {"bf16 SS N=32, 1 acc", 0, 0, 32, 1, 2, 0, 0},
      {"bf16 SS N=32, 4 acc INTERLEAVED per MMA", 0, 0, 32, 4, 2, 0, 1},
      {"bf16 SS N=32, 8 acc INTERLEAVED per MMA", 0, 0, 32, 8, 2, 0, 1},
      {"bf16 SS N=128, 2 acc INTERLEAVED per MMA", 0, 0, 128, 2, 2, 0, 1},
      {"tf32 TS N=256, 1 acc (ref)", 1, 1, 256, 1, 4, 0, 0},
  };
  int dev = 0, sms = 0, clk_khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, 2 * sms * sizeof(long long)));
  const int iters = 20000;
  printf("%d SMs, %d MHz nominal; %d iterations x 4 MMAs (K = 4 x %s) per SM\n", sms, clk_khz / 1000, iters, "16 bf16 | 8 tf32");
  printf("%-44s %12s %12s %10s\n", "shape (M=128)", "clk/MMA", "MAC/clk/SM", "of nominal");
  for (const Shape& sh : shapes) {
    const size_t smem = 1024 + 16384 + (size_t)sh.bslots * sh.N * 128;
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<<<sms, 128, smem>>>(sh, iters, d_cycles);
      CK(cudaDeviceSynchronize());
    }
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (long long c : h) avg += (double)c;
    avg /= sms;
    const double per_mma = avg / (iters * 4.0);
    const double k = sh.tf32 ? 8 : 16;
    const double mac = 128.0 * sh.N * k / per_mma;
    printf("%-44s %12.1f %12.1f %10.2f", sh.name, per_mma, mac, mac / (sh.tf32 ? 2048.0 : 4096.0));
    if (sh.traffic) {
      std::vector<long long> l(sms);
      CK(cudaMemcpy(l.data(), d_cycles + sms, sms * sizeof(long long), cudaMemcpyDeviceToHost));
      double a = 0;
      for (long long c : l) a += (double)c;
      printf("   | x32 %s round trip: %.0f clk", sh.traffic == 1 ? "ld+st" : sh.traffic == 2 ? "ld" : "st", a / sms);
    }
    printf("\n");
  }
  printf("lean issue loop (precomputed descriptors, unrolled x4):\n");
  run_lean<256, false, true>("bf16 TS N=256", 1, sms, d_cycles);
  run_lean<256, true, true>("tf32 TS N=256", 1, sms, d_cycles);
  run_lean<128, false, false>("bf16 SS N=128", 1, sms, d_cycles);
  run_lean<64, false, false>("bf16 SS N=64", 1, sms, d_cycles);
  run_lean<64, false, false>("bf16 SS N=64", 2, sms, d_cycles);
  run_lean<64, false, false>("bf16 SS N=64", 4, sms, d_cycles);
  run_lean<64, false, true>("bf16 TS N=64", 1, sms, d_cycles);
  run_lean<32, false, false>("bf16 SS N=32", 1, sms, d_cycles);
  run_lean<32, false, false>("bf16 SS N=32", 4, sms, d_cycles);
  const Shape pair_shapes[] = {
      {"pair bf16 SS N=64 (QK, Q in smem), 3 acc", 0, 0, 64, 3, 2, 0, 0},
      {"pair bf16 TS N=64 (QK, Q in TMEM), 3 acc", 0, 1, 64, 3, 2, 0, 0},
      {"pair bf16 TS N=256 (PV bf16)", 0, 1, 256, 1, 4, 0, 0},
      {"pair tf32 TS N=256 (PV fp32 contract)", 1, 1, 256, 1, 4, 0, 0},
      {"pair bf16 SS N=256", 0, 0, 256, 1, 4, 0, 0},
  };
  CK(cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("%-44s %12s %12s %10s\n", "CTA pair (M=256 over 2 SMs)", "clk/MMA", "MAC/clk/SM", "of nominal");
  for (const Shape& sh : pair_shapes) {
    const size_t smem = 1024 + 16384 + (size_t)sh.bslots * (sh.N / 2) * 128;
    const int ctas = sms / 2 * 2;
    for (int rep = 0; rep < 2; ++rep) {
      rate2_kernel<<<ctas, 128, smem>>>(sh, iters, d_cycles);
      CK(cudaDeviceSynchronize());
    }
    std::vector<long long> h(ctas / 2);
    CK(cudaMemcpy(h.data(), d_cycles, (ctas / 2) * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (long long c : h) avg += (double)c;
    avg /= (ctas / 2);
    const double per_mma = avg / (iters * 4.0);
    const double k = sh.tf32 ? 8 : 16;
    const double mac = 128.0 * sh.N * k / per_mma;  // per SM: each SM does 128 of the 256 rows
    printf("%-44s %12.1f %12.1f %10.2f\n", sh.name, per_mma, mac, mac / (sh.tf32 ? 2048.0 : 4096.0));
  }
  return 0;
}
