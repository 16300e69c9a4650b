// a1/a2: reference-guided attention (ExampleGuidedAttention, modules/example_guided_att.py:21-41) and PICNet
// Auto_Attn (modules/pluralistic_model/base_function.py:420-448) as one flash-style sm_100a kernel family.
//
//   q = Wq x (+ bq)            [d, S]  per image       (example_guided_att.py:27 / base_function.py:429)
//   E = q^T q                  [S, S]  NO 1/sqrt(d)    (example_guided_att.py:30 / base_function.py:432)
//   P = softmax_j E[i, j]                              (example_guided_att.py:30 / base_function.py:433)
//   O[c, i] = sum_j P[i, j] V[c, j]                    (example_guided_att.py:18 / base_function.py:436,443)
//   epilogue per value group (see include/fmi_b200.h)  (example_guided_att.py:34-36 / base_function.py:439,445)
//
// Data layout in HBM
//   inputs/outputs stay in the reference's NCHW ([N, C, S], S contiguous).
//   workspace (tensor-core operand staging, written by two tiny prologue kernels):
//     Qt   [N, S, qrow]   q transposed, bf16, d padded to a whole 128-byte row (dpad); with the fp32 contract the row
//                         is [hi | lo]: q = hi + lo to 16 mantissa bits, and E = hi.hi + hi.lo + lo.hi (3 bf16 MMAs)
//     Vcat [N, Cv, S]     value groups concatenated, bf16 or tf32-rounded fp32 (S contiguous = K-major B operand)
//
// Main kernel: one CTA per (128 query rows, image, 256-channel slice of V). 6 warps:
//   warp 0  TMA producer: Q tile once; per 128-key tile the K tile (same Qt tensor: keys == queries) and the V tile in
//           128-byte K-chunks, through mbarrier rings (SWIZZLE_128B, the layout tcgen05 reads directly).
//   warp 1  tcgen05.mma issuer (one elected lane): S = Q K^T into TMEM (double buffered), then O += P V^T with P read
//           straight from TMEM (A operand in tensor memory; P overwrites S in place) and V from shared memory.
//   warps 2-5  one thread per query row: tcgen05.ld S, online softmax in the log2 domain (FFMA + ex2.approx), lazy
//           rescale of the O accumulator (only when the running max grows by > 2^8), tcgen05.st P; finally the fused
//           epilogue (1/rowsum, gamma/alpha/mask blend with the fp32 residual, NCHW store coalesced along S).
//   TMEM: O 256 columns + 2 x 128 columns of S/P = 512 columns (whole SM).
//   The S x S map never leaves the SM. Key tiles are visited diagonal-first (tile j = i first): with keys == queries the
//   row maximum is almost always on the diagonal block, so the lazy rescale practically never fires.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

using namespace sm100;

namespace {

constexpr int BM = 128;          // query rows per CTA
constexpr int BN = 128;          // keys per tile
constexpr int CV_MAX = 256;      // value channels per CTA (TMEM columns of O)
constexpr int ATOM_BYTES = 128;  // swizzle atom row
constexpr int kAttnThreads = 320;  // TMA warp + MMA warp + two softmax warpgroups
constexpr int kStaticSmem = 3072;               // barriers + exchange buffer, padded to the 1024-byte alignment
constexpr int kSmemBudget = 232448 - kStaticSmem;  // 227 KB opt-in minus the static part (dynamic base is 1024-aligned)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct AttnParams {
  int N, S, C0, C1, cv_tile;
  int d_atoms;  // 128-byte (64 x bf16) atoms per Qt component
  int k_last;   // 16-element K slices of the last atom that hold data (1..4); the rest is zero padding
  int split;    // 1: Qt rows hold [hi | lo] bf16 components (fp32 contract), 0: hi only
  int k_stages, v_stages;
  const void* v0;
  const void* v1;
  const float* mask;
  const float* a0;
  const float* a1;
  float b0, b1;
  int masked0, masked1;
  void* out0;
  void* out1;
  int64_t out0_bs, out1_bs;
  float* lse;
  void* o_save;        // [N, C0+C1, S] (type T) normalised attention output before the blend, for backward; or NULL
  long long* trace;  // debug: per-tile clock64 stamps of CTA (0,0,0) (fmi_debug_set_attn_trace), else NULL
  int dbg;             // debug knobs (FMI_ATTN_DBG): skip parts of the fast kernel to attribute time; 0 in production
  const float* qmax2;  // [N] max_j |q_j|^2 per image (selects fixed-bound fast path vs this robust kernel), or NULL
};

static long long* g_attn_trace = nullptr;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kSafeQ2 = 256.f;  // attn_fwd2_kernel's fixed-bound softmax is used when max_j |q_j|^2 <= this

// CLUSTER: two CTAs (adjacent query tiles of one image) form a cluster and share every K and V tile: each CTA loads half
// of the tile and TMA-multicasts it into both shared memories, halving the L2 -> SMEM fill traffic that bounds the
// non-cluster kernel (profiles/README.md). Ring slots are released by both consumers (multicast tcgen05.commit).
template <bool TF32, typename T, bool CLUSTER>
__global__ void __launch_bounds__(kAttnThreads, 1)
    attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
  constexpr int EPA = TF32 ? 32 : 64;        // operand elements per 128-byte atom row
  constexpr int V_CHUNKS = BN / EPA;         // K-chunks of the PV product per key tile
  constexpr int P_COLS_PER_CHUNK = 32;       // TMEM columns of P per chunk (32 tf32 or 64 packed bf16)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // 1024-byte aligned by construction (checked below): SWIZZLE_128B needs it
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const int q_atoms = p.d_atoms * (1 + p.split);  // atoms per Qt row
  const int q_tile_bytes = q_atoms * BM * ATOM_BYTES;
  const int v_chunk_bytes = p.cv_tile * ATOM_BYTES;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_tile_bytes;                   // k_stages tiles
  uint8_t* sV = sK + p.k_stages * q_tile_bytes;      // v_stages chunks

  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[8], v_empty[8], s_full[2], p_full[2], pv_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float xch[2][2][BM];  // row max / row sum exchange between the two softmax warpgroups (double buffered)

  const int tid = threadIdx.x, warp = tid >> 5;
  const int i_tile = blockIdx.x, n = blockIdx.y, cv0 = blockIdx.z * p.cv_tile;
  // as the fallback of attn_fwd2_kernel this kernel only takes the images whose logits are too large for the fixed bound
  if (p.qmax2 && p.qmax2[n] <= kSafeQ2) return;
  const int NT = p.S / BN;  // key tiles
  const uint32_t cta_rank = CLUSTER ? cluster_ctarank() : 0;
  const int tile0 = CLUSTER ? (i_tile & ~1) : i_tile;  // first key tile (diagonal first; shared by the CTA pair)
  constexpr uint32_t kConsumers = CLUSTER ? 2 : 1;     // CTAs that must release a K/V ring slot
  constexpr uint16_t kMask = 0x3;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], kConsumers);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], kConsumers);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CLUSTER) cluster_sync_all();  // the peer's barriers must be initialised before any multicast reaches them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_O = tmem;  // [0, 256)
  auto tmem_S = [&](int b) { return tmem + 256 + b * 128; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      mbar_arrive_expect_tx(&q_full, q_tile_bytes);
      for (int a = 0; a < q_atoms; ++a)
        tma_load_2d(sQ + a * BM * ATOM_BYTES, &map_q, &q_full, a * 64, n * p.S + i_tile * BM);
      auto load_k = [&](int jj) {
        const int j = (tile0 + jj) % NT;
        const int slot = jj % p.k_stages;
        mbar_wait(&k_empty[slot], ((jj / p.k_stages) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[slot], q_tile_bytes);  // both halves land here
        for (int a = 0; a < q_atoms; ++a) {
          uint8_t* dst = sK + slot * q_tile_bytes + a * BN * ATOM_BYTES;
          if (CLUSTER)  // this CTA fetches key rows [64r, 64r+64) of the tile and multicasts them to both CTAs
            tma_load_2d_mc(dst + cta_rank * (BN / 2) * ATOM_BYTES, &map_k, &k_full[slot], a * 64,
                           n * p.S + j * BN + cta_rank * (BN / 2), kMask);
          else
            tma_load_2d(dst, &map_k, &k_full[slot], a * 64, n * p.S + j * BN);
        }
      };
      auto load_v = [&](int jj) {
        const int j = (tile0 + jj) % NT;
        for (int c = 0; c < V_CHUNKS; ++c) {
          const int use = jj * V_CHUNKS + c;
          const int slot = use % p.v_stages;
          mbar_wait(&v_empty[slot], ((use / p.v_stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[slot], v_chunk_bytes);
          if (CLUSTER) {  // this CTA fetches channels [r*cv/2, (r+1)*cv/2) of the chunk for both CTAs
            const int half = p.cv_tile / 2;
            tma_load_2d_mc(sV + slot * v_chunk_bytes + cta_rank * half * ATOM_BYTES, &map_v, &v_full[slot],
                           j * BN + c * EPA, n * (p.C0 + p.C1) + cv0 + cta_rank * half, kMask);
          } else {
            tma_load_2d(sV + slot * v_chunk_bytes, &map_v, &v_full[slot], j * BN + c * EPA, n * (p.C0 + p.C1) + cv0);
          }
        }
      };
      // same order as the MMA warp consumes: K(0), K(1), V(0), K(2), V(1), ...
      load_k(0);
      for (int jj = 0; jj < NT; ++jj) {
        if (jj + 1 < NT) load_k(jj + 1);
        load_v(jj);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_qk = make_idesc(KIND_BF16, BM, BN);  // logits always from bf16 (hi/lo split) operands
      const uint32_t idesc_pv = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, BM, p.cv_tile);
      auto issue_qk = [&](int jj) {
        const int slot = jj % p.k_stages, b = jj & 1;
        mbar_wait(&k_full[slot], (jj / p.k_stages) & 1);
        tc_fence_after();
        uint32_t acc = 0;
        // fp32 contract: q = hi + lo (two bf16 terms, 16 mantissa bits); E = hi.hi + hi.lo + lo.hi (error ~2^-16)
        const int npairs = p.split ? 3 : 1;
        for (int pr = 0; pr < npairs; ++pr) {
          const int ca = pr == 2 ? 1 : 0, cb = pr == 1 ? 1 : 0;  // component (0 = hi, 1 = lo) of A and B
          for (int a = 0; a < p.d_atoms; ++a) {
            const uint64_t adesc = make_sdesc_k_sw128(smem_u32(sQ + (ca * p.d_atoms + a) * BM * ATOM_BYTES));
            const uint64_t bdesc =
                make_sdesc_k_sw128(smem_u32(sK + slot * q_tile_bytes + (cb * p.d_atoms + a) * BN * ATOM_BYTES));
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              mma_ss_f16(tmem_S(b), adesc + 2 * s, bdesc + 2 * s, idesc_qk, acc);
              acc = 1;
            }
          }
        }
        if (CLUSTER) tc_commit_mc(&k_empty[slot], kMask);
        else tc_commit(&k_empty[slot]);
        tc_commit(&s_full[b]);
      };
      mbar_wait(&q_full, 0);
      issue_qk(0);
      for (int jj = 0; jj < NT; ++jj) {
        const int b = jj & 1;
        if (jj + 1 < NT) {
          // S[(jj+1)&1] still holds P(jj-1) until PV(jj-1) has completed
          if (jj >= 1) mbar_wait(&pv_done[(jj + 1) & 1], ((jj - 1) >> 1) & 1);
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && jj < 256) p.trace[jj * 16 + 7] = clock64();  // PV(jj-1) complete
          issue_qk(jj + 1);
        }
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && jj < 256;
        if (tr) p.trace[jj * 16 + 8] = clock64();   // QK(jj+1) issued, about to wait for P(jj)
        mbar_wait(&p_full[b], (jj >> 1) & 1);
        if (tr) p.trace[jj * 16 + 9] = clock64();   // P(jj) available
        tc_fence_after();
        for (int c = 0; c < V_CHUNKS; ++c) {
          const int use = jj * V_CHUNKS + c;
          const int slot = use % p.v_stages;
          mbar_wait(&v_full[slot], (use / p.v_stages) & 1);
          tc_fence_after();
          const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(sV + slot * v_chunk_bytes));
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint32_t a_t = tmem_S(b) + c * P_COLS_PER_CHUNK + s * 8;
            const uint32_t acc = (jj > 0 || c > 0 || s > 0) ? 1u : 0u;
            if (TF32) mma_ts_tf32(tmem_O, a_t, bdesc + 2 * s, idesc_pv, acc);
            else mma_ts_f16(tmem_O, a_t, bdesc + 2 * s, idesc_pv, acc);
          }
          if (CLUSTER) tc_commit_mc(&v_empty[slot], kMask);
          else tc_commit(&v_empty[slot]);
        }
        tc_commit(&pv_done[b]);
        if (tr) p.trace[jj * 16 + 10] = clock64();  // PV(jj) issued (V chunks were available)
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (one thread per row)
    // Two softmax warpgroups split every S tile by columns: thread (wg, row) owns 64 of the 128 logits of its row.
    const int wg = (warp - 2) >> 2;         // 0: key columns [0,64), 1: [64,128)
    const int lane_base = (warp & 3) * 32;  // TMEM lanes this warp may touch
    const int row = lane_base + (tid & 31);
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const int n_chunks = p.cv_tile / 32;    // 32-column chunks of O: chunk k belongs to warpgroup k & 1
    float m_used = 0.f, l = 0.f;
    for (int jj = 0; jj < NT; ++jj) {
      const int b = jj & 1;
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 64 && jj < 256;
      if (tr) p.trace[jj * 16 + 0] = clock64();  // softmax warps ready for tile jj
      mbar_wait(&s_full[b], (jj >> 1) & 1);
      if (tr) p.trace[jj * 16 + 1] = clock64();  // S(jj) available
      tc_fence_after();
      uint32_t s[64];
      tmem_ld32(tmem_S(b) + lane_addr + wg * 64, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_S(b) + lane_addr + wg * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tc_wait_ld();
      if (tr) p.trace[jj * 16 + 2] = clock64();  // S in registers
      float mx4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) mx4[q] = __uint_as_float(s[q]);
#pragma unroll
      for (int k = 4; k < 64; ++k) mx4[k & 3] = fmaxf(mx4[k & 3], __uint_as_float(s[k]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kLog2e;
      // row maximum over both halves: exchange through shared memory (double buffered by tile parity)
      xch[b][wg][row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, xch[b][wg ^ 1][row]);
      if (tr) p.trace[jj * 16 + 3] = clock64();  // row max exchanged
      if (jj == 0) {
        m_used = mx;
      } else {
        const bool need = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // O may only be touched once PV(jj-1) has landed; each warpgroup rescales its own column chunks
          mbar_wait(&pv_done[(jj - 1) & 1], ((jj - 1) >> 1) & 1);
          tc_fence_after();
          const float f = need ? ex2(m_used - mx) : 1.f;
          for (int ck = wg; ck < n_chunks; ck += 2) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_addr + ck * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
            tmem_st32(tmem_O + lane_addr + ck * 32, o);
          }
          l *= f;
          if (need) m_used = mx;
        }
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_used;
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const float pk = ex2(fmaf(__uint_as_float(s[k]), kLog2e, neg_m));
        sum4[k & 3] += pk;
        s[k] = __float_as_uint(pk);
      }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (tr) p.trace[jj * 16 + 4] = clock64();  // exponentials done
      if (TF32) {
        // the tensor core drops the 13 low mantissa bits of a tf32 operand: +0x1000 first = round to nearest
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] += 0x1000u;
        tmem_st32(tmem_S(b) + lane_addr + wg * 64, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_st32(tmem_S(b) + lane_addr + wg * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) s[k] = pack_bf16x2(__uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1]));
        tmem_st32(tmem_S(b) + lane_addr + wg * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[b]);
      if (tr) p.trace[jj * 16 + 5] = clock64();  // P(jj) published
    }
    // ---- epilogue: total row sum = both halves
    // (the exchange buffer of the other parity was last read two tiles ago: free)
    const int xb = (NT & 1);
    xch[xb][wg][row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xch[xb][wg ^ 1][row];
    mbar_wait(&pv_done[(NT - 1) & 1], ((NT - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const int i = i_tile * BM + row;
    const float m_i = p.mask ? p.mask[(int64_t)n * p.S + i] : 0.f;
    const float alpha0 = p.a0 ? *p.a0 : 1.f, alpha1 = p.a1 ? *p.a1 : 1.f;
    if (p.lse && blockIdx.z == 0 && wg == 0) p.lse[(int64_t)n * p.S + i] = (m_used + log2f(l)) * 0.6931471805599453f;
    for (int ck = wg; ck < n_chunks; ck += 2) {
      const int c0 = ck * 32;
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_addr + c0, o);
      tc_wait_ld();
      const int cg = cv0 + c0;  // channel in Vcat; a 32-channel chunk never straddles the two groups
      const bool g1 = cg >= p.C0;
      const int c_in_group = g1 ? cg - p.C0 : cg;
      const int Cg = g1 ? p.C1 : p.C0;
      const T* v = (const T*)(g1 ? p.v1 : p.v0) + ((int64_t)n * Cg + c_in_group) * p.S + i;
      T* out = (T*)(g1 ? p.out1 : p.out0) + (int64_t)n * (g1 ? p.out1_bs : p.out0_bs) + (int64_t)c_in_group * p.S + i;
      const bool masked = g1 ? p.masked1 : p.masked0;
      const float a = (g1 ? alpha1 : alpha0) * (masked ? (1.f - m_i) : 1.f);
      const float r = masked ? m_i : (g1 ? p.b1 : p.b0);
      if (p.o_save) {
        T* os = (T*)p.o_save + ((int64_t)n * (p.C0 + p.C1) + cg) * p.S + i;
#pragma unroll
        for (int k = 0; k < 32; ++k) os[(int64_t)k * p.S] = from_f32<T>(__uint_as_float(o[k]) * inv_l);
      }
      if (r != 0.f || masked) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float ov = __uint_as_float(o[k]) * inv_l;
          out[(int64_t)k * p.S] = from_f32<T>(fmaf(a, ov, r * to_f32<T>(v[(int64_t)k * p.S])));
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) out[(int64_t)k * p.S] = from_f32<T>(a * (__uint_as_float(o[k]) * inv_l));
      }
    }
  }
  tc_fence_before();
  if (CLUSTER) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it / arrive on its barriers
  else __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

#include "attention_v2.cuh"
#include "attention_v3.cuh"

// ---- prologue kernels ---------------------------------------------------------------------------
// 1x1 convolution as a small SIMT fp32 GEMM: y[n,o,s] = sum_c W[o,c] x[n,c,s] + b[o].
//   OUT_QT = false: y is NCHW of type TO.
//   OUT_QT = true : y is Qt [N, S, dpad] (transposed, rows padded with zeros to dpad), TO = bf16 or float(tf32-rounded).
constexpr int CT_O = 64, CT_S = 64, CT_K = 32;

template <typename TI, typename TO, bool OUT_QT, bool ROUND_TF32>
__global__ void __launch_bounds__(256) conv1x1_kernel(const TI* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, TO* __restrict__ y, int Cin,
                                                      int Cout, int S, int dpad) {
  __shared__ float sw[CT_K][CT_O + 1];
  __shared__ float sx[CT_K][CT_S];
  const int n = blockIdx.z, o0 = blockIdx.y * CT_O, s0 = blockIdx.x * CT_S;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  const TI* xn = x + (int64_t)n * Cin * S;
  for (int c0 = 0; c0 < Cin; c0 += CT_K) {
    for (int i = threadIdx.x; i < CT_K * CT_O; i += 256) {
      int o = i / CT_K, c = i % CT_K;
      sw[c][o] = (o0 + o < Cout && c0 + c < Cin) ? w[(int64_t)(o0 + o) * Cin + c0 + c] : 0.f;
    }
    for (int i = threadIdx.x; i < CT_K * CT_S; i += 256) {
      int c = i / CT_S, s = i % CT_S;
      sx[c][s] = (c0 + c < Cin && s0 + s < S) ? to_f32<TI>(xn[(int64_t)(c0 + c) * S + s0 + s]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < CT_K; ++c) {
      float wv[4], xv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) wv[a] = sw[c][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) xv[b] = sx[c][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(wv[a], xv[b], acc[a][b]);
    }
    __syncthreads();
  }
  if (OUT_QT) {
    // Qt rows are [hi(dpad)] or, with SPLIT (= ROUND_TF32 slot of the template), [hi(dpad) | lo(dpad)]: stage the 64 x 64 tile
    // through shared memory so that a warp writes whole rows (the direct store scattered 2-byte elements over 16 rows per
    // instruction: 129 us for the 128^2 x 64 -> 16 query projection of the PICNet decoder, 25x its bytes)
    __shared__ float stage[CT_S][CT_O + 1];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int o = o0 + ty * 4 + a;
      const float bo = (bias && o < Cout) ? bias[o] : 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) stage[tx + 16 * b][ty * 4 + a] = o < Cout ? acc[a][b] + bo : 0.f;
    }
    __syncthreads();
    const int ow = min(CT_O, dpad - o0);                  // columns of this tile inside the padded row
    const int64_t rowlen = ROUND_TF32 ? 2 * dpad : dpad;
    for (int i = threadIdx.x; i < CT_S * ow; i += 256) {
      const int sl = i / ow, o = i - sl * ow;
      if (s0 + sl >= S) continue;
      const float v = stage[sl][o];
      const TO hi = from_f32<TO>(v);
      TO* row = y + ((int64_t)n * S + s0 + sl) * rowlen + o0 + o;
      row[0] = hi;
      if (ROUND_TF32) row[dpad] = from_f32<TO>(v - to_f32<TO>(hi));
    }
    return;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int o = o0 + ty * 4 + a;
    const float bo = (bias && o < Cout) ? bias[o] : 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int s = s0 + tx + 16 * b;
      if (s >= S) continue;
      if (o < Cout) y[((int64_t)n * Cout + o) * S + s] = from_f32<TO>(acc[a][b] + bo);
    }
  }
}

// Vcat[n, c, s] = (c < C0 ? v0[n, c, s] : v1[n, c - C0, s]) converted to the operand type
template <typename TI, typename TO, bool ROUND_TF32>
__global__ void __launch_bounds__(256) pack_values_kernel(const TI* __restrict__ v0, const TI* __restrict__ v1,
                                                          TO* __restrict__ vcat, int C0, int C1, int64_t S, int64_t total4) {
  // 4 elements per thread; S % 4 == 0 so a group of 4 never crosses a row
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = q * 4;
    const int64_t s = e % S;
    const int64_t row = e / S;  // n * (C0 + C1) + c
    const int c = (int)(row % (C0 + C1));
    const int64_t n = row / (C0 + C1);
    const TI* src = c < C0 ? v0 + (n * C0 + c) * S + s : v1 + (n * C1 + (c - C0)) * S + s;
    float f[4];
    if constexpr (sizeof(TI) == 4) {
      float4 t = *reinterpret_cast<const float4*>(src);
      f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
    } else {
      union { uint2 u; TI e[4]; } t;
      t.u = *reinterpret_cast<const uint2*>(src);
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = to_f32<TI>(t.e[k]);
    }
    if constexpr (sizeof(TO) == 4) {
      float4 o;
      if (ROUND_TF32) {
        o.x = __uint_as_float(f32_to_tf32_rna(f[0])); o.y = __uint_as_float(f32_to_tf32_rna(f[1]));
        o.z = __uint_as_float(f32_to_tf32_rna(f[2])); o.w = __uint_as_float(f32_to_tf32_rna(f[3]));
      } else {
        o = make_float4(f[0], f[1], f[2], f[3]);
      }
      *reinterpret_cast<float4*>(vcat + e) = o;
    } else {
      uint2 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      *reinterpret_cast<uint2*>(vcat + e) = o;
    }
  }
}

// attn[n, i, j] = exp(q_i . q_j - lse_i) from the staged Qt (opt-in materialisation, base_function.py:448)
__global__ void __launch_bounds__(256) attn_materialize_kernel(const __nv_bfloat16* __restrict__ qt,
                                                               const float* __restrict__ lse, float* __restrict__ attn,
                                                               int S, int dpad, int split) {
  extern __shared__ float sq[];  // 16 query rows x dpad
  const int n = blockIdx.y, i0 = blockIdx.x * 16;
  const int rowlen = dpad * (1 + split);
  const __nv_bfloat16* qn = qt + (int64_t)n * S * rowlen;
  auto qval = [&](int64_t row, int k) {
    float v = __bfloat162float(qn[row * rowlen + k]);
    if (split) v += __bfloat162float(qn[row * rowlen + dpad + k]);
    return v;
  };
  for (int t = threadIdx.x; t < 16 * dpad; t += 256) sq[t] = qval(i0 + t / dpad, t % dpad);
  __syncthreads();
  for (int j = threadIdx.x; j < S; j += 256) {
    float acc[16] = {};
    for (int k = 0; k < dpad; ++k) {
      const float kv = qval(j, k);
#pragma unroll
      for (int r = 0; r < 16; ++r) acc[r] = fmaf(sq[r * dpad + k], kv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r)
      attn[((int64_t)n * S + i0 + r) * S + j] = __expf(acc[r] - lse[(int64_t)n * S + i0 + r]);
  }
}

struct AttnPlan {
  int dpad, d_atoms, split, cv_tile, k_stages, v_stages, esz;  // esz: element size of the V / P operands
  int k_stages2, v_stages2;                                    // ring depths of attn_fwd2_kernel
  int v_stages3;                                               // V ring depth of attn_fwd3_kernel (CTA pairs)
  int d;                                                       // head dimension as given (dpad is its padded size)
  int64_t qt_bytes, vcat_bytes, qmax_bytes;
  size_t smem, smem2, smem3;
};

int make_plan(int N, int d, int C0, int C1, int S, int mma, AttnPlan* pl) {
  FMI_REQUIRE(mma == FMI_MMA_TF32 || mma == FMI_MMA_BF16, "attn: mma must be FMI_MMA_TF32 or FMI_MMA_BF16");
  FMI_REQUIRE(N >= 1 && d >= 1 && d <= 256 && C0 >= 32 && C1 >= 0, "attn: bad shape N=%d d=%d C0=%d C1=%d", N, d, C0, C1);
  FMI_REQUIRE(S >= BN && S % BN == 0, "attn: S=%d must be a positive multiple of %d", S, BN);
  FMI_REQUIRE(C0 % 32 == 0 && C1 % 32 == 0, "attn: value channel counts (%d, %d) must be multiples of 32", C0, C1);
  const int esz = mma == FMI_MMA_TF32 ? 4 : 2;
  pl->esz = esz;
  pl->d = d;
  pl->split = mma == FMI_MMA_TF32 ? 1 : 0;  // fp32 contract: logits from [hi | lo] bf16 pairs
  pl->dpad = (d + 63) / 64 * 64;
  pl->d_atoms = pl->dpad / 64;
  const int Cv = C0 + C1;
  pl->cv_tile = Cv <= CV_MAX ? Cv : CV_MAX;
  FMI_REQUIRE(Cv % pl->cv_tile == 0 && pl->cv_tile % 16 == 0, "attn: C0+C1=%d must be <= 256 or a multiple of 256", Cv);
  const int q_tile = pl->d_atoms * (1 + pl->split) * BM * ATOM_BYTES;
  const int v_chunk = pl->cv_tile * ATOM_BYTES;
  pl->k_stages = (3 * q_tile + 4 * v_chunk <= kSmemBudget) ? 2 : 1;
  int vs = (kSmemBudget - (1 + pl->k_stages) * q_tile) / v_chunk;
  if (vs > 8) vs = 8;
  FMI_REQUIRE(vs >= 2, "attn: d=%d too large for shared memory", d);
  pl->v_stages = vs;
  pl->smem = (size_t)(1 + pl->k_stages) * q_tile + (size_t)vs * v_chunk;
  pl->qt_bytes = ((int64_t)N * S * pl->dpad * (1 + pl->split) * 2 + 1023) / 1024 * 1024;
  pl->vcat_bytes = ((int64_t)N * Cv * S * esz + 1023) / 1024 * 1024;
  pl->qmax_bytes = ((int64_t)N * 4 + 1023) / 1024 * 1024;
  // fast kernel: same tiles, its own (smaller) static footprint
  pl->k_stages2 = (3 * q_tile + 4 * v_chunk <= kAttn2SmemBudget) ? 2 : 1;
  int vs2 = (kAttn2SmemBudget - (1 + pl->k_stages2) * q_tile) / v_chunk;
  if (vs2 > 8) vs2 = 8;
  pl->v_stages2 = vs2;
  pl->smem2 = (size_t)(1 + pl->k_stages2) * q_tile + (size_t)vs2 * v_chunk;
  // pair kernel: each CTA holds half of every K tile (2 stages) and half of every V chunk
  int vs3 = (kAttn3SmemBudget - 2 * q_tile) / (v_chunk / 2);
  if (vs3 > 8) vs3 = 8;
  pl->v_stages3 = vs3;
  pl->smem3 = (size_t)2 * q_tile + (size_t)(vs3 > 0 ? vs3 : 0) * (v_chunk / 2);
  return FMI_OK;
}

template <typename TI>
int launch_conv1x1_any(const void* x, const float* w, const float* b, void* y, int N, int Cin, int Cout, int S, int dpad,
                       int out_mode /*0 NCHW same type, 1 Qt bf16 [hi], 2 Qt bf16 [hi | lo]*/, cudaStream_t st) {
  const int o_extent = out_mode == 0 ? Cout : dpad;
  dim3 grid((S + CT_S - 1) / CT_S, (o_extent + CT_O - 1) / CT_O, N);
  if (out_mode == 0)
    conv1x1_kernel<TI, TI, false, false><<<grid, 256, 0, st>>>((const TI*)x, w, b, (TI*)y, Cin, Cout, S, 0);
  else if (out_mode == 1)
    conv1x1_kernel<TI, __nv_bfloat16, true, false><<<grid, 256, 0, st>>>((const TI*)x, w, b, (__nv_bfloat16*)y, Cin, Cout, S, dpad);
  else
    conv1x1_kernel<TI, __nv_bfloat16, true, true><<<grid, 256, 0, st>>>((const TI*)x, w, b, (__nv_bfloat16*)y, Cin, Cout, S, dpad);
  return fmi_launched("conv1x1");
}

template <typename TI>
int launch_pack_values(const void* v0, const void* v1, void* vcat, int N, int C0, int C1, int S, int mma, cudaStream_t st) {
  const int64_t total4 = (int64_t)N * (C0 + C1) * S / 4;
  int grid = (int)imin64((total4 + 255) / 256, (int64_t)FMI_NUM_SMS * 16);
  if (mma == FMI_MMA_TF32)
    pack_values_kernel<TI, float, true><<<grid, 256, 0, st>>>((const TI*)v0, (const TI*)v1, (float*)vcat, C0, C1, S, total4);
  else
    pack_values_kernel<TI, __nv_bfloat16, false><<<grid, 256, 0, st>>>((const TI*)v0, (const TI*)v1, (__nv_bfloat16*)vcat, C0, C1, S, total4);
  return fmi_launched("pack_values");
}

// Algorithmic work of one attention launch: FLOPs = 2*S^2*d (QK^T) + 2*S^2*Cv (P.V); bytes = q + values read once +
// outputs written once + mask, in the I/O element type.
template <typename T>
static inline void attn_work(const AttnParams& prm, const AttnPlan& pl, double* flops, double* bytes) {
  const double S = prm.S, N = prm.N, Cv = prm.C0 + prm.C1;
  *flops = N * (2.0 * S * S * pl.d + 2.0 * S * S * Cv);
  *bytes = N * S * ((pl.d + 2.0 * Cv) * sizeof(T) + (prm.mask ? 4.0 : 0.0));
}

template <bool TF32, typename T, bool CLUSTER>
int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& prm,
                const AttnPlan& pl, cudaStream_t st) {
  auto kern = attn_fwd_kernel<TF32, T, CLUSTER>;
  static FmiPerDeviceOnce attr_once;  // per template instantiation
  if (attr_once.need()) {
    cudaFuncAttributes fa;
    FMI_CUDA(cudaFuncGetAttributes(&fa, kern));
    FMI_REQUIRE((int)fa.sharedSizeBytes <= kStaticSmem, "attn_fwd: static shared memory grew to %d bytes", (int)fa.sharedSizeBytes);
    FMI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attr_once.done();
  }
  dim3 grid(prm.S / BM, prm.N, (prm.C0 + prm.C1) / prm.cv_tile);
  // kind 0 = the kernel that does the work; as the fallback behind the fast kernel (qmax2 given: it exits at once for
  // every image the fast kernel took) it is timed separately so it cannot dilute the roofline average
  double wf, wb;
  attn_work<T>(prm, pl, &wf, &wb);
  // (as the fallback it exits at once for every image the fast kernel took: no algorithmic work is attributed to it)
  FmiProfScope prof(prm.qmax2 ? FMI_PROF_ATTN_FALLBACK : FMI_PROF_ATTN, st, prm.qmax2 ? 0.0 : wf, prm.qmax2 ? 0.0 : wb);
  if (CLUSTER) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kAttnThreads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    FMI_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, prm));
  } else {
    kern<<<grid, kAttnThreads, pl.smem, st>>>(mq, mk, mv, prm);
  }
  return fmi_launched("attn_fwd");
}

template <bool TF32, typename T, bool CLUSTER>
int launch_attn2(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, AttnParams prm, const AttnPlan& pl,
                 cudaStream_t st) {
  auto kern = attn_fwd2_kernel<TF32, T, CLUSTER>;
  static FmiPerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaFuncAttributes fa;
    FMI_CUDA(cudaFuncGetAttributes(&fa, kern));
    FMI_REQUIRE((int)fa.sharedSizeBytes <= kAttn2StaticSmem, "attn_fwd2: static shared memory grew to %d bytes",
                (int)fa.sharedSizeBytes);
    FMI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2SmemBudget));
    attr_once.done();
  }
  prm.k_stages = pl.k_stages2;
  prm.v_stages = pl.v_stages2;
  dim3 grid(prm.S / BM, prm.N, (prm.C0 + prm.C1) / prm.cv_tile);
  double wf, wb;
  attn_work<T>(prm, pl, &wf, &wb);
  FmiProfScope prof(FMI_PROF_ATTN, st, wf, wb);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kAttn2Threads);
  cfg.dynamicSmemBytes = pl.smem2;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLUSTER ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  FMI_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, prm));
  return fmi_launched("attn_fwd2");
}

template <bool TF32, typename T>
int launch_attn3(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, AttnParams prm, const AttnPlan& pl,
                 cudaStream_t st) {
  auto kern = attn_fwd3_kernel<TF32, T>;
  static FmiPerDeviceOnce attr_once;
  if (attr_once.need()) {
    cudaFuncAttributes fa;
    FMI_CUDA(cudaFuncGetAttributes(&fa, kern));
    FMI_REQUIRE((int)fa.sharedSizeBytes <= kAttn3StaticSmem, "attn_fwd3: static shared memory grew to %d bytes",
                (int)fa.sharedSizeBytes);
    FMI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn3SmemBudget));
    attr_once.done();
  }
  prm.k_stages = 2;
  prm.v_stages = pl.v_stages3;
  dim3 grid(prm.S / BM, prm.N, (prm.C0 + prm.C1) / prm.cv_tile);
  double wf, wb;
  attn_work<T>(prm, pl, &wf, &wb);
  FmiProfScope prof(FMI_PROF_ATTN, st, wf, wb);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kAttn3Threads);
  cfg.dynamicSmemBytes = pl.smem3;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  FMI_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, prm));
  return fmi_launched("attn_fwd3");
}

template <bool TF32, typename T>
int launch_attn2_any(bool cluster, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                     const AttnParams& prm, const AttnPlan& pl, cudaStream_t st) {
  return cluster ? launch_attn2<TF32, T, true>(mq, mk, mv, prm, pl, st) : launch_attn2<TF32, T, false>(mq, mk, mv, prm, pl, st);
}

template <bool TF32, typename T>
int launch_attn_any(bool cluster, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                    const AttnParams& prm, const AttnPlan& pl, cudaStream_t st) {
  return cluster ? launch_attn<TF32, T, true>(mq, mk, mv, prm, pl, st) : launch_attn<TF32, T, false>(mq, mk, mv, prm, pl, st);
}

}  // namespace

// Operand staging shared with the backward pass (attn_bwd.cu): Qt (bf16 [hi | lo] rows) and Vcat.
int fmi_attn_stage_operands(const void* x, const float* wq, const float* bq, const void* v0, const void* v1, void* qt,
                            void* vcat, int N, int C, int d, int C0, int C1, int S, int dtype, int mma, cudaStream_t st) {
  AttnPlan pl;
  int rc = make_plan(N, d, C0, C1, S, mma, &pl);
  if (rc) return rc;
  const bool tf32 = mma == FMI_MMA_TF32;
  if (dtype == FMI_F32) {
    rc = launch_conv1x1_any<float>(x, wq, bq, qt, N, C, d, S, pl.dpad, tf32 ? 2 : 1, st);
    if (!rc) rc = launch_pack_values<float>(v0, v1, vcat, N, C0, C1, S, mma, st);
  } else {
    rc = launch_conv1x1_any<__nv_bfloat16>(x, wq, bq, qt, N, C, d, S, pl.dpad, tf32 ? 2 : 1, st);
    if (!rc) rc = launch_pack_values<__nv_bfloat16>(v0, v1, vcat, N, C0, C1, S, mma, st);
  }
  return rc;
}

extern "C" int64_t fmi_attn_workspace_bytes(int N, int C, int d, int C0, int C1, int S, int mma) {
  (void)C;
  AttnPlan pl;
  if (make_plan(N, d, C0, C1, S, mma, &pl)) return -1;
  return pl.qt_bytes + pl.vcat_bytes + pl.qmax_bytes;
}

extern "C" int fmi_conv1x1(const void* x, const float* w, const float* b, void* y, int N, int Cin, int Cout, int S,
                           int dtype, void* stream) {
  FMI_REQUIRE(fmi_dtype_ok(dtype), "conv1x1: unsupported dtype %d", dtype);
  FMI_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1 && S >= 1, "conv1x1: bad shape");
  if (N == 0) return FMI_OK;
  FMI_REQUIRE(x && w && y, "conv1x1: null pointer");
  FMI_DISPATCH_DTYPE(dtype, T, return (launch_conv1x1_any<T>(x, w, b, y, N, Cin, Cout, S, 0, 0, (cudaStream_t)stream)));
  return FMI_OK;
}

extern "C" int fmi_attn_fwd(const void* x, const float* wq, const float* bq, const void* v0, const void* v1,
                            const float* mask, const float* a0, float b0, int masked0, const float* a1, float b1,
                            int masked1, void* out0, int64_t out0_bs, void* out1, int64_t out1_bs, float* lse, void* o_save, int N, int C,
                            int d, int C0, int C1, int S, int dtype, int mma, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  FMI_REQUIRE(dtype == FMI_F32 || dtype == FMI_BF16, "attn_fwd: dtype must be fp32 or bf16");
  if (N == 0) return FMI_OK;
  AttnPlan pl;
  int rc = make_plan(N, d, C0, C1, S, mma, &pl);
  if (rc) return rc;
  FMI_REQUIRE(x && wq && v0 && out0 && workspace, "attn_fwd: null pointer");
  FMI_REQUIRE(C >= 1, "attn_fwd: bad C");
  FMI_REQUIRE((C1 == 0) == (v1 == nullptr) && (C1 == 0 || out1), "attn_fwd: v1/out1 must be given exactly when C1 > 0");
  FMI_REQUIRE(!(masked0 || masked1) || mask, "attn_fwd: masked group without a mask");
  FMI_REQUIRE(workspace_bytes >= pl.qt_bytes + pl.vcat_bytes + pl.qmax_bytes, "attn_fwd: workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)(pl.qt_bytes + pl.vcat_bytes + pl.qmax_bytes));
  FMI_REQUIRE(fmi_aligned(workspace, 1024), "attn_fwd: workspace must be 1024-byte aligned");
  FMI_REQUIRE(fmi_aligned(v0, 16) && (!v1 || fmi_aligned(v1, 16)), "attn_fwd: value tensors must be 16-byte aligned");
  rc = fmi_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* qt = (uint8_t*)workspace;
  uint8_t* vcat = qt + pl.qt_bytes;
  float* qmax2 = (float*)(vcat + pl.vcat_bytes);
  const bool tf32 = mma == FMI_MMA_TF32;
  // FMI_ATTN_KERNEL=robust: only the online-max kernel; default: fixed-bound fast kernel + robust fallback per image
  static const bool fast_env0 = [] { const char* e = getenv("FMI_ATTN_KERNEL"); return !(e && e[0] == 'r'); }();
  const bool fast_env = fast_env0 && pl.d_atoms <= 2;  // the fast kernels keep the Q/K descriptors of <= 4 atoms in registers

  if (dtype == FMI_F32) {
    rc = launch_conv1x1_any<float>(x, wq, bq, qt, N, C, d, S, pl.dpad, tf32 ? 2 : 1, st);
    if (!rc) rc = launch_pack_values<float>(v0, v1, vcat, N, C0, C1, S, mma, st);
  } else {
    rc = launch_conv1x1_any<__nv_bfloat16>(x, wq, bq, qt, N, C, d, S, pl.dpad, tf32 ? 2 : 1, st);
    if (!rc) rc = launch_pack_values<__nv_bfloat16>(v0, v1, vcat, N, C0, C1, S, mma, st);
  }
  if (rc) return rc;
  if (fast_env) {
    FMI_CUDA(cudaMemsetAsync(qmax2, 0, (size_t)N * sizeof(float), st));
    dim3 g((S + 63) / 64 < 592 ? (S + 63) / 64 : 592, N);
    qnorm_max_kernel<<<g, 256, 0, st>>>((const __nv_bfloat16*)qt, qmax2, S, pl.dpad, pl.split);
    rc = fmi_launched("qnorm_max");
    if (rc) return rc;
  }

  // 2-CTA clusters (K/V multicast) need an even number of query tiles; FMI_ATTN_CLUSTER=0 disables them
  static const bool cluster_env = [] { const char* e = getenv("FMI_ATTN_CLUSTER"); return !(e && e[0] == '0'); }();
  const bool cluster = cluster_env && ((S / BM) % 2 == 0) && (pl.cv_tile % 16 == 0);
  CUtensorMap mq, mk, mv;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const uint32_t epa = ATOM_BYTES / pl.esz;
  {
    const uint64_t qrow = (uint64_t)pl.dpad * (1 + pl.split);  // bf16 elements per Qt row
    uint64_t dims[2] = {qrow, (uint64_t)N * S};
    uint64_t str[1] = {qrow * 2};
    uint32_t box[2] = {64, (uint32_t)BM};
    int e = make_tensor_map(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qt, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "attn_fwd: cuTensorMapEncodeTiled(Qt) failed (%d)", e);
    uint32_t kbox[2] = {64, (uint32_t)(cluster ? BN / 2 : BN)};  // each CTA of a pair fetches half of the key rows
    e = make_tensor_map(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qt, dims, str, kbox, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "attn_fwd: cuTensorMapEncodeTiled(K) failed (%d)", e);
  }
  {
    uint64_t dims[2] = {(uint64_t)S, (uint64_t)N * (C0 + C1)};
    uint64_t str[1] = {(uint64_t)S * pl.esz};
    uint32_t box[2] = {epa, (uint32_t)(cluster ? pl.cv_tile / 2 : pl.cv_tile)};
    int e = make_tensor_map(&mv, dt, 2, vcat, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "attn_fwd: cuTensorMapEncodeTiled(Vcat) failed (%d)", e);
  }
  AttnParams prm;
  prm.N = N; prm.S = S; prm.C0 = C0; prm.C1 = C1; prm.cv_tile = pl.cv_tile; prm.d_atoms = pl.d_atoms; prm.split = pl.split;
  prm.k_last = (pl.d - 64 * (pl.d_atoms - 1) + 15) / 16;
  { static const bool full = [] { const char* e = getenv("FMI_ATTN_KTRIM"); return e && e[0] == '0'; }(); if (full) prm.k_last = 4; }
  prm.k_stages = pl.k_stages; prm.v_stages = pl.v_stages;
  prm.v0 = v0; prm.v1 = v1; prm.mask = mask; prm.a0 = a0; prm.a1 = a1; prm.b0 = b0; prm.b1 = b1;
  prm.masked0 = masked0; prm.masked1 = masked1;
  prm.out0 = out0; prm.out1 = out1; prm.out0_bs = out0_bs; prm.out1_bs = out1_bs; prm.lse = lse; prm.o_save = o_save;
  prm.trace = g_attn_trace;
  prm.qmax2 = fast_env ? qmax2 : nullptr;
  { const char* e = getenv("FMI_ATTN_DBG"); prm.dbg = e ? atoi(e) : 0; }
  // CTA-pair kernel (cta_group::2, Q in tensor memory): opt-in with FMI_ATTN_PAIR=1. It is parity-green but measured
  // 5-15 % SLOWER than attn_fwd2_kernel (profiles/README.md), so it is kept as the vehicle for the round-2 work on the
  // pair path, not as the default. Needs an even number of query tiles and a V tile that splits into two MMA-N halves.
  const bool pair_env = [] { const char* e = getenv("FMI_ATTN_PAIR"); return e && e[0] == '1'; }();  // read per call
  const bool pair = fast_env && pair_env && cluster && pl.cv_tile % 32 == 0 && pl.v_stages3 >= 2 &&
                    256 - pl.d_atoms * (1 + pl.split) * 32 >= 2 * BS;  // Q in TMEM must leave two S/P buffers
  if (pair) {
    CUtensorMap mk3;
    const uint64_t qrow = (uint64_t)pl.dpad * (1 + pl.split);
    uint64_t dims[2] = {qrow, (uint64_t)N * S};
    uint64_t str[1] = {qrow * 2};
    uint32_t kbox[2] = {64, (uint32_t)(BS / 2)};  // 32 key rows: this CTA's half of one 64-key step
    int e = make_tensor_map(&mk3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qt, dims, str, kbox, CU_TENSOR_MAP_SWIZZLE_128B);
    FMI_REQUIRE(e == 0, "attn_fwd: cuTensorMapEncodeTiled(K, pair) failed (%d)", e);
    // mv already has the half-height box {epa, cv_tile / 2} of the cluster variants
    if (tf32) rc = dtype == FMI_F32 ? launch_attn3<true, float>(mq, mk3, mv, prm, pl, st)
                                    : launch_attn3<true, __nv_bfloat16>(mq, mk3, mv, prm, pl, st);
    else rc = dtype == FMI_F32 ? launch_attn3<false, float>(mq, mk3, mv, prm, pl, st)
                               : launch_attn3<false, __nv_bfloat16>(mq, mk3, mv, prm, pl, st);
    if (rc) return rc;
  } else if (fast_env) {  // fast kernel first; images with max|q|^2 > kSafeQ2 fall through to the robust kernel below
    if (tf32) rc = dtype == FMI_F32 ? launch_attn2_any<true, float>(cluster, mq, mk, mv, prm, pl, st)
                                    : launch_attn2_any<true, __nv_bfloat16>(cluster, mq, mk, mv, prm, pl, st);
    else rc = dtype == FMI_F32 ? launch_attn2_any<false, float>(cluster, mq, mk, mv, prm, pl, st)
                               : launch_attn2_any<false, __nv_bfloat16>(cluster, mq, mk, mv, prm, pl, st);
    if (rc) return rc;
  }
  if (tf32) {
    if (dtype == FMI_F32) return launch_attn_any<true, float>(cluster, mq, mk, mv, prm, pl, st);
    return launch_attn_any<true, __nv_bfloat16>(cluster, mq, mk, mv, prm, pl, st);
  }
  if (dtype == FMI_F32) return launch_attn_any<false, float>(cluster, mq, mk, mv, prm, pl, st);
  return launch_attn_any<false, __nv_bfloat16>(cluster, mq, mk, mv, prm, pl, st);
}

// Debug aid: clock64 stamps of CTA (0,0,0) of subsequent fmi_attn_fwd calls are written to dev_buffer
// (>= 256*16 long long); NULL switches tracing off. Not part of the product interface.
extern "C" int fmi_debug_set_attn_trace(void* dev_buffer) {
  g_attn_trace = (long long*)dev_buffer;
  return FMI_OK;
}

extern "C" int fmi_attn_materialize(const void* workspace, const float* lse, float* attn, int N, int d, int S, int mma,
                                    void* stream) {
  AttnPlan pl;
  int rc = make_plan(N, d, 32, 0, S, mma, &pl);
  if (rc) return rc;
  FMI_REQUIRE(workspace && lse && attn, "attn_materialize: null pointer");
  dim3 grid(S / 16, N);
  size_t smem = (size_t)16 * pl.dpad * sizeof(float);
  attn_materialize_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)workspace, lse, attn, S, pl.dpad,
                                                                      pl.split);
  return fmi_launched("attn_materialize");
}
