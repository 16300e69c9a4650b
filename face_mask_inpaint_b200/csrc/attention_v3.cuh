// attn_fwd3_kernel — attn_fwd2_kernel on CTA PAIRS (tcgen05 cta_group::2). Included by attention.cu.
//
// Why: the round-1 attribution runs (profiles/README.md) showed attn_fwd2_kernel bound by the L2 -> shared-memory operand
// stream: every 128-row CTA pulls the whole K and V of its image through its SM (with all MMAs skipped the kernel still
// took 3.5 of 4.4 ms), and TMA multicast does not help because each SM still ingests every byte. With cta_group::2 one
// MMA covers 256 query rows on two SMs and the B operand is SPLIT between them: per key step each CTA loads only 32 of
// the 64 K rows and cv_tile/2 of the V rows, so the bytes each SM ingests per FLOP halve.
//
// Structure per CTA (rank r of the pair; i_tile = its own 128 query rows):
//   warp 0   TMA producer: own Q tile; per K tile its 2 x 32 key rows; per step its cv_tile/2 value rows. K/V bytes of
//            BOTH CTAs are counted on the LEADER's full barriers (cp.async.bulk.tensor ... cta_group::2).
//   warp 1   leader: QK issuer, S[h&3] (256 x 64) = Q K_h^T, M = 256; peer: tells the leader when its Q tile landed.
//   warp 2   leader: PV issuer, O (256 x cv_tile) += P_h V_h^T, P read from each CTA's own tensor memory.
//   warps 4-11  two softmax warpgroups alternating steps on the CTA's own 128 rows (fixed-bound softmax as in v2), then
//            the fused epilogue. P-ready arrivals of both CTAs go to the leader's barrier (remote mbarrier.arrive).
//   Completion (slot free, S ready, PV done) is multicast to both CTAs by tcgen05.commit.cta_group::2.
//
// Q lives in TENSOR MEMORY. tools/umma_rate.cu (profiles/r01_umma_rate.txt) measured that an MMA with both operands in
// shared memory is bound by ~70 B/clk/SM of operand reads: the 128 x 64 x 16 QK step MMA takes 87 clk instead of 32,
// because it re-reads the 4 KB Q tile every time — with the hi/lo split that was 1049 of the ~2150 clk a key step costs.
// The softmax threads therefore copy their Q row (hi and lo) into TMEM columns [448, 512) once, and S = Q K^T runs in the
// A-from-TMEM form, reading only this CTA's 32 key rows (1 KB per MMA) from shared memory. TMEM: O 256 columns,
// (256 - q columns) / 64 = 3 S/P buffers (2 for d = 128), Q 32..128 columns.
#pragma once

constexpr int kAttn3Threads = 384;

// One 64-key step of S = Q K^T on a CTA pair with Q in tensor memory: descriptors / TMEM addresses come precomputed, the
// body is only MMAs (see qk_step_mmas in attention_v2.cuh for why that matters).
template <int D_ATOMS, int NPAIRS>
__device__ __forceinline__ void qk_step_mmas_ts2(uint32_t d_tmem, uint32_t tmem_q, const uint64_t (&kb)[4], uint64_t hoff,
                                                 uint32_t idesc) {
#pragma unroll
  for (int pr = 0; pr < NPAIRS; ++pr) {
    const int ca = pr == 2 ? 1 : 0, cb = pr == 1 ? 1 : 0;  // hi.hi, hi.lo, lo.hi
#pragma unroll
    for (int a = 0; a < D_ATOMS; ++a) {
      const uint32_t at = tmem_q + (ca * D_ATOMS + a) * 32;
      const uint64_t bd = kb[cb * D_ATOMS + a] + hoff;
      mma2_ts_f16(d_tmem, at, bd, idesc, (pr | a) ? 1u : 0u);
      mma2_ts_f16(d_tmem, at + 8, bd + 2, idesc, 1u);
      mma2_ts_f16(d_tmem, at + 16, bd + 4, idesc, 1u);
      mma2_ts_f16(d_tmem, at + 24, bd + 6, idesc, 1u);
    }
  }
}
constexpr int kAttn3StaticSmem = 2048;
constexpr int kAttn3SmemBudget = 232448 - kAttn3StaticSmem;

template <bool TF32, typename T>
__global__ void __launch_bounds__(kAttn3Threads, 1)
    attn_fwd3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
  constexpr int EPA = TF32 ? 32 : 64;   // V / P operand elements per 128-byte row
  constexpr int VC = BS / EPA;          // V chunks per step (2 for tf32, 1 for bf16)
  constexpr int KH = BS / 2;            // key rows of a step held by each CTA
  const int n = blockIdx.y;
  if (p.qmax2[n] > kSafeQ2) return;     // (uniform per image, hence per pair) the robust kernel handles it

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int q_atoms = p.d_atoms * (1 + p.split);
  const int q_tile_bytes = q_atoms * BM * ATOM_BYTES;
  const int k_half_bytes = q_atoms * (BN / 2) * ATOM_BYTES;   // this CTA's share of a 128-key tile: [atom][step][32 rows]
  const int v_half = p.cv_tile / 2;
  const int v_half_bytes = v_half * ATOM_BYTES;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_tile_bytes;
  uint8_t* sV = sK + p.k_stages * k_half_bytes;

  __shared__ uint64_t q_full, q_pair, k_full[2], k_empty[2], v_full[8], v_empty[8], s_full[4], p_full[4], pv_done[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ float xsum[BM];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int i_tile = blockIdx.x, cv0 = blockIdx.z * p.cv_tile;
  const int NS = p.S / BS;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int tile0 = i_tile & ~1;  // first 128-key tile: the pair's own (diagonal) block
  const int NT = p.S / BN;
  constexpr uint16_t kBoth = 0x3;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&q_pair, 2 * 8);    // every softmax warp of both CTAs, once its Q rows sit in tensor memory
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);    // one arrival per softmax warp of the step's warpgroup, both CTAs (leader's copy is used)
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(&tmem_base_s, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_O = tmem;
  const int q_cols = q_atoms * 32;            // packed bf16 pairs: 32 columns per 64-wide atom
  const int NB = (256 - q_cols) / BS;         // S/P buffers (host guarantees >= 2)
  const uint32_t tmem_Q = tmem + 512 - q_cols;
  auto tmem_S = [&](int b) { return tmem + 256 + b * BS; };
  const bool lane0 = elect_one();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs)
    if (lane0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      mbar_arrive_expect_tx(&q_full, q_tile_bytes);
      for (int a = 0; a < q_atoms; ++a)
        tma_load_2d(sQ + a * BM * ATOM_BYTES, &map_q, &q_full, a * 64, n * p.S + i_tile * BM);
      uint32_t kf[2], vf[8];  // the LEADER's full barriers (cluster addresses)
      for (int i = 0; i < 2; ++i) kf[i] = mapa_u32(smem_u32(&k_full[i]), 0);
      for (int i = 0; i < 8; ++i) vf[i] = mapa_u32(smem_u32(&v_full[i]), 0);
      auto load_k = [&](int t) {
        const int j = (tile0 + t) % NT;
        const int slot = t % p.k_stages;
        mbar_wait(&k_empty[slot], ((t / p.k_stages) & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(&k_full[slot], 2 * k_half_bytes);
        for (int a = 0; a < q_atoms; ++a)
          for (int hh = 0; hh < 2; ++hh)  // step hh of the tile uses keys [hh*64, hh*64+64): this CTA holds 32 of them
            tma_load_2d_pair(sK + slot * k_half_bytes + (a * 2 + hh) * KH * ATOM_BYTES, &map_k, kf[slot], a * 64,
                             n * p.S + j * BN + hh * BS + (int)rank * KH);
      };
      load_k(0);
      for (int h = 0; h < NS; ++h) {
        if ((h & 1) == 0 && h / 2 + 1 < NT) load_k(h / 2 + 1);
        const int key0 = ((tile0 + h / 2) % NT) * BN + (h & 1) * BS;
        for (int c = 0; c < VC; ++c) {
          const int use = h * VC + c;
          const int slot = use % p.v_stages;
          mbar_wait(&v_empty[slot], ((use / p.v_stages) & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(&v_full[slot], 2 * v_half_bytes);
          tma_load_2d_pair(sV + slot * v_half_bytes, &map_v, vf[slot], key0 + c * EPA,
                           n * (p.C0 + p.C1) + cv0 + (int)rank * v_half);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- QK issuer (leader): S[h % NB] = Q K_h^T  (256 x 64)
    if (lane0 && leader) {
      const uint32_t idesc_qk = make_idesc(KIND_BF16, 2 * BM, BS);
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long w_q = 0, w_pv = 0, w_k = 0, c0 = clock64(), c1;
      mbar_wait_cluster(&q_pair, 0);  // both Q tiles are in tensor memory
      tc_fence_after();
      c1 = clock64(); w_q = c1 - c0;
      // K atoms of both ring slots: [atom][step][32 rows]; a step adds the 32-row offset
      uint64_t kb[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ii = i < q_atoms ? i : 0;
        kb[0][i] = make_sdesc_k_sw128(smem_u32(sK + ii * 2 * KH * ATOM_BYTES));
        kb[1][i] = make_sdesc_k_sw128(smem_u32(sK + k_half_bytes + ii * 2 * KH * ATOM_BYTES));
      }
      const uint64_t step_off = (uint64_t)((KH * ATOM_BYTES) >> 4);
      int b = 0, use = 0;
      for (int h = 0; h < NS; ++h) {
        const int t = h >> 1, slot = t & 1;
        c0 = clock64();
        if (use > 0) mbar_wait(&pv_done[b], (use - 1) & 1);  // P(h - NB) lived in this buffer
        c1 = clock64(); w_pv += c1 - c0;
        mbar_wait(&k_full[slot], (t >> 1) & 1);
        tc_fence_after();
        c0 = clock64(); w_k += c0 - c1;
        const uint64_t hoff = (h & 1) ? step_off : 0;
        const uint32_t d_s = tmem_S(b);
        if (slot == 0) {
          if (p.split) { if (p.d_atoms == 1) qk_step_mmas_ts2<1, 3>(d_s, tmem_Q, kb[0], hoff, idesc_qk); else qk_step_mmas_ts2<2, 3>(d_s, tmem_Q, kb[0], hoff, idesc_qk); }
          else { if (p.d_atoms == 1) qk_step_mmas_ts2<1, 1>(d_s, tmem_Q, kb[0], hoff, idesc_qk); else qk_step_mmas_ts2<2, 1>(d_s, tmem_Q, kb[0], hoff, idesc_qk); }
        } else {
          if (p.split) { if (p.d_atoms == 1) qk_step_mmas_ts2<1, 3>(d_s, tmem_Q, kb[1], hoff, idesc_qk); else qk_step_mmas_ts2<2, 3>(d_s, tmem_Q, kb[1], hoff, idesc_qk); }
          else { if (p.d_atoms == 1) qk_step_mmas_ts2<1, 1>(d_s, tmem_Q, kb[1], hoff, idesc_qk); else qk_step_mmas_ts2<2, 1>(d_s, tmem_Q, kb[1], hoff, idesc_qk); }
        }
        tc_commit2_mc(&s_full[b], kBoth);
        if (h & 1) tc_commit2_mc(&k_empty[slot], kBoth);  // both steps of the K tile consumed
        if (++b == NB) { b = 0; ++use; }
      }
      if (tr) { p.trace[0] = w_q; p.trace[1] = w_pv; p.trace[2] = w_k; p.trace[3] = clock64(); }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ---------------------------------------------------------------- PV issuer (leader): O += P_h V_h^T  (256 x cv_tile)
    if (lane0 && leader) {
      const uint32_t idesc_pv = make_idesc(TF32 ? KIND_TF32 : KIND_BF16, 2 * BM, p.cv_tile);
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long w_p = 0, w_v = 0, c0, c1;
      const long long cstart = clock64();
      for (int h = 0; h < NS; ++h) {
        const int b = h % NB;
        c0 = clock64();
        mbar_wait_cluster(&p_full[b], (h / NB) & 1);
        tc_fence_after();
        c1 = clock64(); w_p += c1 - c0;
        for (int c = 0; c < VC; ++c) {
          const int use = h * VC + c;
          const int slot = use % p.v_stages;
          c0 = clock64();
          mbar_wait(&v_full[slot], (use / p.v_stages) & 1);
          tc_fence_after();
          c1 = clock64(); w_v += c1 - c0;
          const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(sV + slot * v_half_bytes));
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint32_t a_t = tmem_S(b) + c * 32 + s * 8;
            const uint32_t acc = (h > 0 || c > 0 || s > 0) ? 1u : 0u;
            if (TF32) mma2_ts_tf32(tmem_O, a_t, bdesc + 2 * s, idesc_pv, acc);
            else mma2_ts_f16(tmem_O, a_t, bdesc + 2 * s, idesc_pv, acc);
          }
          tc_commit2_mc(&v_empty[slot], kBoth);
        }
        tc_commit2_mc(&pv_done[b], kBoth);
      }
      if (tr) { p.trace[4] = w_p; p.trace[5] = w_v; p.trace[6] = clock64() - cstart; }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax warpgroups (alternate steps) + epilogue
    const int wg = (warp - 4) >> 2;
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + (tid & 31);
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    uint32_t pf[4];  // the leader's P-ready barriers
#pragma unroll
    for (int i = 0; i < 4; ++i) pf[i] = mapa_u32(smem_u32(&p_full[i]), 0);
    mbar_wait(&q_full, 0);
    // this row of the Q tile: |q|^2 for the fixed softmax shift, and a copy into tensor memory (A operand of S = Q K^T).
    // TMA wrote the row with SWIZZLE_128B: logical 16-byte chunk c sits at chunk c ^ (row & 7).
    float q2 = 0.f;
    const uint32_t q_st = tmem_Q + lane_addr;
    for (int a = 0; a < p.d_atoms; ++a) {
      const uint4* hi = reinterpret_cast<const uint4*>(sQ + a * BM * ATOM_BYTES + row * ATOM_BYTES);
      const uint4* lo = reinterpret_cast<const uint4*>(sQ + (p.d_atoms + a) * BM * ATOM_BYTES + row * ATOM_BYTES);
      uint32_t wh[32], wl[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 h4 = hi[c ^ (row & 7)];
        const uint4 l4v = p.split ? lo[c ^ (row & 7)] : make_uint4(0, 0, 0, 0);
        wh[4 * c] = h4.x; wh[4 * c + 1] = h4.y; wh[4 * c + 2] = h4.z; wh[4 * c + 3] = h4.w;
        wl[4 * c] = l4v.x; wl[4 * c + 1] = l4v.y; wl[4 * c + 2] = l4v.z; wl[4 * c + 3] = l4v.w;
      }
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float x0 = __uint_as_float(wh[e] << 16) + __uint_as_float(wl[e] << 16);
        const float x1 = __uint_as_float(wh[e] & 0xffff0000u) + __uint_as_float(wl[e] & 0xffff0000u);
        q2 = fmaf(x0, x0, q2);
        q2 = fmaf(x1, x1, q2);
      }
      if (wg == 0) tmem_st32(q_st + a * 32, wh);                               // warpgroup 0 stages hi,
      else if (p.split) tmem_st32(q_st + (p.d_atoms + a) * 32, wl);            // warpgroup 1 stages lo
    }
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&q_pair), 0));
    const float m_i = sqrtf(q2 * p.qmax2[n]) * kLog2e * 1.00001f + 1e-6f;
    const float neg_m = -m_i;
    const int n_chunks = p.cv_tile / 32;
    float l4[4] = {0.f, 0.f, 0.f, 0.f};
    const bool trs = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid & 127) == 0;
    long long w_s = 0, w_c = 0, w_ld = 0, w_ex = 0, cs0, cs1;
    for (int h = wg; h < NS; h += 2) {
      const int b = h % NB;
      cs0 = clock64();
      mbar_wait(&s_full[b], (h / NB) & 1);
      tc_fence_after();
      cs1 = clock64(); w_s += cs1 - cs0;
      uint32_t s[64];
      tmem_ld32(tmem_S(b) + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_S(b) + lane_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tc_wait_ld();
      const long long cs2 = clock64();
      w_ld += cs2 - cs1;
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const float pk = ex2(fmaf(__uint_as_float(s[k]), kLog2e, neg_m));
        l4[k & 3] += pk;
        s[k] = __float_as_uint(pk);
      }
      const long long cs3 = clock64();
      w_ex += cs3 - cs2;
      if (TF32) {
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] += 0x1000u;  // round-to-nearest for the truncating tf32 operand read
        tmem_st32(tmem_S(b) + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_st32(tmem_S(b) + lane_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) s[k] = pack_bf16x2(__uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1]));
        tmem_st32(tmem_S(b) + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      }
      tc_wait_st();
      tc_fence_before();
      // ONE remote arrival per warp: a release.cluster arrive per thread cost ~1000 clk per step (measured with the
      // trace counters below) and made the softmax, not the tensor pipe, the critical path
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive_cluster(pf[b]);
      w_c += clock64() - cs1;
    }
    if (trs) { p.trace[8 + 2 * wg] = w_s; p.trace[9 + 2 * wg] = w_c; p.trace[12 + 2 * wg] = w_ld; p.trace[13 + 2 * wg] = w_ex; }
    // ---- epilogue (identical to attn_fwd2_kernel: each CTA owns its 128 rows of O)
    float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
    if (wg == 1) xsum[row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (wg == 0) {
      l += xsum[row];
      xsum[row] = l;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l = xsum[row];
    mbar_wait(&pv_done[(NS - 1) % NB], ((NS - 1) / NB) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const int i = i_tile * BM + row;
    const float mk = p.mask ? p.mask[(int64_t)n * p.S + i] : 0.f;
    const float alpha0 = p.a0 ? *p.a0 : 1.f, alpha1 = p.a1 ? *p.a1 : 1.f;
    if (p.lse && blockIdx.z == 0 && wg == 0) p.lse[(int64_t)n * p.S + i] = (m_i + log2f(l)) * 0.6931471805599453f;
    for (int ck = wg; ck < n_chunks; ck += 2) {
      const int c0 = ck * 32;
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_addr + c0, o);
      tc_wait_ld();
      const int cg = cv0 + c0;
      const bool g1 = cg >= p.C0;
      const int c_in_group = g1 ? cg - p.C0 : cg;
      const int Cg = g1 ? p.C1 : p.C0;
      const T* v = (const T*)(g1 ? p.v1 : p.v0) + ((int64_t)n * Cg + c_in_group) * p.S + i;
      T* out = (T*)(g1 ? p.out1 : p.out0) + (int64_t)n * (g1 ? p.out1_bs : p.out0_bs) + (int64_t)c_in_group * p.S + i;
      const bool masked = g1 ? p.masked1 : p.masked0;
      const float a = (g1 ? alpha1 : alpha0) * (masked ? (1.f - mk) : 1.f);
      const float r = masked ? mk : (g1 ? p.b1 : p.b0);
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
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, 512);
}
