// attn_tc_fwd_split.cuh — EXPERIMENT (compiled only with -DVG_ATTN_FWD_SPLIT=2; included by attn_tc.cu inside namespace vg
// after its helpers): the forward attention kernel with two softmax warps per TMEM lane quadrant.  It is a copy of
// attn_tc_fwd_kernel with the column ranges, the accumulator width and the partner exchange parametrised; the product
// kernel is left textually untouched so that the default build stays byte-identical to the one validated on the GPU.
#pragma once

// -DVG_ATTN_FWD_SPLIT=2 (experiment, default 1): two softmax warps per TMEM lane quadrant, each taking half of the key
// columns of a tile and half of the output columns — four warps per scheduler instead of two hide the TMEM-load / MUFU
// latency of the two softmax passes (the same step took the backward's arithmetic phase from 1.6 k to 0.7 k cycles).
// The partner warps exchange their row maxima (and, once, their row sums) through 2 KB of shared memory and a 64-thread
// named barrier per quadrant.  Written at the end of round 1, not yet run on the GPU; the default build is unchanged.
constexpr int FWDS_SPLIT = 2;
constexpr int FWDS_THREADS = 64 + 128 * FWDS_SPLIT;
constexpr int FWDS_CHUNKS = (TK / 32) / FWDS_SPLIT;        // 32-column chunks of a key tile per softmax warp
constexpr int FWDS_OC = HD / FWDS_SPLIT;                   // output columns per softmax warp
constexpr int FWDS_XCH = FWDS_SPLIT == 2 ? 2048 : 0;       // partner exchange: 2 x 128 maxima + 2 x 128 sums
constexpr int FWDS_SMEM = 5 * TILE_BYTES /*Q, K0, K1, V*/ + TILE_BYTES /*P second half*/ + 1024 + 256 + FWDS_XCH;
// layout: Q | K0 | K1 | V | P(2 halves)

__global__ void __launch_bounds__(FWDS_THREADS, 2)
attn_tc_fwd_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   float* __restrict__ lse, const int32_t* __restrict__ kv_len, const float* __restrict__ slopes,
                   AttnTcShape sh) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;            // 2 stages
  uint8_t* sV = smem + 3 * TILE_BYTES;
  uint8_t* sP = smem + 4 * TILE_BYTES;        // 2 halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 6,
           *s_full = bars + 7, *p_full = bars + 8, *o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* xch = reinterpret_cast<float*>(smem + 6 * TILE_BYTES + 256);       // (FWDS_SPLIT == 2 only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 64) AT_MARK(0);
  // heavy (late) query tiles first: causal work grows with the tile index
  const int n_qt = (sh.Tq + TQ - 1) / TQ;
  // 1-D grid, query tile slowest: every (head, batch) CTA of the heaviest tile is dispatched before any lighter one
  const int qt = n_qt - 1 - (int)(blockIdx.x / (unsigned)(sh.H * sh.B));
  const int hb = (int)(blockIdx.x % (unsigned)(sh.H * sh.B));
  const int h = hb % sh.H, b = hb / sh.H;
  const int q0 = qt * TQ;
  const int klen = kv_len ? min(kv_len[b], sh.Tk) : sh.Tk;
  const int q_abs_last = sh.q_offset + min(q0 + TQ, sh.Tq) - 1;
  const int k_end = min(klen, q_abs_last + 1);
  const int n_kt = (k_end + TK - 1) / TK;       // may be 0 (empty sequence)

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmO);
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
    mbar_init(&k_empty[0], 1); mbar_init(&k_empty[1], 1);
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128 * FWDS_SPLIT); mbar_init(o_full, 1);
    fence_barrier_init();
    // this thread is the TMA producer: the first tiles are requested before the CTA-wide set-up (TMEM allocation,
    // barrier) so that their ~2 k cycles of latency overlap it
    if (n_kt > 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * HD, q0, b);
      mbar_arrive_expect_tx(&k_full[0], TILE_BYTES);
      tma_load_3d(sK, &tmK, &k_full[0], h * HD, 0, b);
      mbar_arrive_expect_tx(v_full, TILE_BYTES);
      tma_load_3d(sV, &tmV, v_full, h * HD, 0, b);
    }
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;            // columns [0,128): S
  const uint32_t tO = tmem_base + 128;      // columns [128,192): per-tile O_j
  if (threadIdx.x == 64) AT_MARK(1);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    for (int j = 1; j < n_kt; ++j) {
      const int s = j & 1;
      mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
      mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
      tma_load_3d(sK + s * TILE_BYTES, &tmK, &k_full[s], h * HD, j * TK, b);
      mbar_wait(v_empty, (j & 1) ^ 1);
      mbar_arrive_expect_tx(v_full, TILE_BYTES);
      tma_load_3d(sV, &tmV, v_full, h * HD, j * TK, b);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop and one ELECTED lane issues: inside a `lane == 0` branch the compiler wraps every
    // tcgen05.mma in a per-thread ELECT / R2UR loop (~80 cycles per instruction against 32 cycles of tensor time for a
    // 128x64x16 MMA); in warp-uniform code the descriptors stay in uniform registers and the MMAs issue back to back.
    // S_{j+1} is issued BEFORE P_j·V: the softmax threads start on the next key tile while the tensor core finishes
    // this one (they fold O_j one tile late).
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);     // S = Q·Kᵀ : both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);      // O = P·V  : A K-major, B (V) MN-major
    const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP), aV = smem_u32(sV), aK0 = smem_u32(sK);
    if (n_kt > 0) {
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      AT_STAMP(0, 5);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tS, desc_kmajor(aQ, k), desc_kmajor(aK0, k), idesc_s, k != 0);
        umma_commit(&k_empty[0]);
        umma_commit(s_full);
      }
      __syncwarp();
    }
    for (int j = 0; j < n_kt; ++j) {
      AT_STAMP(j, 6);
      mbar_wait(p_full, j & 1);                 // P_j written, S_j fully consumed, O_{j-1} folded
      tc_fence_after();
      AT_STAMP(j, 7);
      if (j + 1 < n_kt) {
        const int s = (j + 1) & 1;
        mbar_wait(&k_full[s], ((j + 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = aK0 + s * TILE_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tS, desc_kmajor(aQ, k), desc_kmajor(aK, k), idesc_s, k != 0);
          umma_commit(&k_empty[s]);
          umma_commit(s_full);
        }
        __syncwarp();
        AT_STAMP(j + 1, 5);
      }
      mbar_wait(v_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < TK / 16; ++k) umma_f16_ss(tO, desc_kmajor(aP, k), desc_mnmajor(aV, k), idesc_o, k != 0);
        umma_commit(v_empty);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else if (warp >= 2) {
    // ===================== softmax (thread = query row) =====================
    const int rb = (warp & 3) * 32;             // TMEM lane quadrant accessible to this warp
    const int r = rb + lane;                    // row inside the tile
    const int part = FWDS_SPLIT == 2 ? (warp - 2) >> 2 : 0;       // which half of the key / output columns
    auto pair_sync = [&]() {                    // the two warps that share a lane quadrant (FWDS_SPLIT == 2)
      asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
    };
    const int iq = q0 + r;
    const int ia = sh.q_offset + iq;
    const float slope2 = (slopes ? slopes[h] : 0.f) * kLog2e;
    const float scale2 = sh.scale * kLog2e;
    const uint32_t lane_addr = (uint32_t)rb << 16;
    const int lim = min(ia + 1, klen);          // keys [0, lim) are visible to this row
    const float rowc = -slope2 * (float)ia;
    const f32x2_t sc2 = splat2(scale2), sl2 = splat2(slope2);
    float m = -CUDART_INF_F, l = 0.f, corr_prev = 1.f;
    float o[FWDS_OC];
#pragma unroll
    for (int d = 0; d < FWDS_OC; ++d) o[d] = 0.f;
    // o = o·corr + O_j  (O_j = P_j·V_j from TMEM; this warp's FWDS_OC output columns)
    auto fold = [&](float corr) {
#pragma unroll
      for (int c = 0; c < FWDS_OC / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tO + lane_addr + part * FWDS_OC + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c * 32 + e] = o[c * 32 + e] * corr + __uint_as_float(v[e]);
      }
    };
    // One key tile.  The scores are formed two at a time with packed f32x2 FMAs: the ALiBi bias is an affine function of
    // the column, s2 = S·scale2 + slope2·(j − i) (log2 domain).  MASKED tiles (the diagonal tile, or one that crosses
    // kv_len) additionally replace columns ≥ lim by −inf; converting column indices to float per element, as the first
    // version did, ran on the same XU pipe as the exponentials and tripled the cost of those tiles.
    auto tile = [&](auto masked_tag, const int j) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int j0 = j * TK;
      float mx = -CUDART_INF_F;
#pragma unroll 1
      for (int c = part * FWDS_CHUNKS; c < (part + 1) * FWDS_CHUNKS; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + lane_addr + c * 32, v);
        tmem_ld_wait();
        const f32x2_t cb0 = splat2(fmaf(slope2, (float)(j0 + c * 32), rowc));
        const int nvalid = lim - (j0 + c * 32);
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const f32x2_t cb = fma2(sl2, pack2((float)e, (float)(e + 1)), cb0);
          float a, b;
          unpack2(fma2(pack2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sc2, cb), a, b);
          if (MASKED) {
            a = (e < nvalid) ? a : -CUDART_INF_F;
            b = (e + 1 < nvalid) ? b : -CUDART_INF_F;
          }
          mx = fmaxf(mx, fmaxf(a, b));
        }
      }
      if (FWDS_SPLIT == 2) {                     // row maximum over both halves of the tile
        xch[part * 128 + r] = mx;
        pair_sync();
        mx = fmaxf(mx, xch[(part ^ 1) * 128 + r]);
      }
      if (threadIdx.x == 64) AT_STAMP(j, 1);
      const float m_new = fmaxf(m, mx);
      const float m_use = (MASKED && m_new == -CUDART_INF_F) ? 0.f : m_new;      // unmasked: every key is visible
      const float corr = ex2_approx(m - m_use);                                   // m = -inf → 0
      if (j > 0) {                              // O_{j-1} has long landed: fold it before P_j may be overwritten
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (threadIdx.x == 64) AT_STAMP(j, 3);
        fold(corr_prev);
        tc_fence_before();
      }
      if (threadIdx.x == 64) AT_STAMP(j, 4);
      f32x2_t rs2 = splat2(0.f);
#pragma unroll 1
      for (int c = part * FWDS_CHUNKS; c < (part + 1) * FWDS_CHUNKS; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + lane_addr + c * 32, v);
        tmem_ld_wait();
        const f32x2_t cb0 = splat2(fmaf(slope2, (float)(j0 + c * 32), rowc) - m_use);
        const int nvalid = lim - (j0 + c * 32);
        float p[32];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const f32x2_t cb = fma2(sl2, pack2((float)e, (float)(e + 1)), cb0);
          float a, b;
          unpack2(fma2(pack2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), sc2, cb), a, b);
          if (MASKED) {
            a = (e < nvalid) ? a : -CUDART_INF_F;
            b = (e + 1 < nvalid) ? b : -CUDART_INF_F;
          }
          p[e] = ex2_approx(a);
          p[e + 1] = ex2_approx(b);
          rs2 = add2(rs2, pack2(p[e], p[e + 1]));
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(p[g * 8 + 0], p[g * 8 + 1]);
          pk.y = pack_bf16x2(p[g * 8 + 2], p[g * 8 + 3]);
          pk.z = pack_bf16x2(p[g * 8 + 4], p[g * 8 + 5]);
          pk.w = pack_bf16x2(p[g * 8 + 6], p[g * 8 + 7]);
          *reinterpret_cast<uint4*>(sP + sw128_piece(r, c * 4 + g)) = pk;
        }
      }
      float ra, rb2;
      unpack2(rs2, ra, rb2);
      l = l * corr + (ra + rb2);
      m = m_new;
      corr_prev = corr;
    };

    for (int j = 0; j < n_kt; ++j) {
      const int j0 = j * TK;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (threadIdx.x == 64) AT_STAMP(j, 0);
      const bool full_tile = (j0 + TK - 1 <= sh.q_offset + q0) && (j0 + TK <= klen);      // CTA-uniform
      if (full_tile) tile(std::false_type{}, j); else tile(std::true_type{}, j);
      tc_fence_before();            // our tcgen05.ld of S_j / O_{j-1} are ordered before the MMAs that overwrite them
      fence_proxy_async();          // generic-proxy writes of P visible to the tensor core (async proxy)
      mbar_arrive(p_full);
      if (threadIdx.x == 64) AT_STAMP(j, 2);
    }
    if (n_kt > 0) {
      mbar_wait(o_full, (n_kt - 1) & 1);
      tc_fence_after();
      fold(corr_prev);
      tc_fence_before();
    }
    if (threadIdx.x == 64) AT_MARK(2);
    // Output rows go through shared memory (the Q tile is dead after the last S MMA; each warp owns its 32 rows of it)
    // in the SWIZZLE_128B box layout and leave as one TMA store per warp: thread-per-row 16-byte global stores touch
    // 32 half-used sectors per request.  Rows beyond Tq are clipped by the tensor map.
    {
      if (FWDS_SPLIT == 2) {                     // l = Σ over both halves (each warp summed its own key columns)
        xch[256 + part * 128 + r] = l;
        pair_sync();
        l += xch[256 + (part ^ 1) * 128 + r];
      }
      const bool valid = ia < klen && l > 0.f;
      const float inv = valid ? 1.f / l : 0.f;
#pragma unroll
      for (int g = 0; g < FWDS_OC / 8; ++g) {
        uint4 pk;
        pk.x = pack_bf16x2(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv);
        pk.y = pack_bf16x2(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
        pk.z = pack_bf16x2(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv);
        pk.w = pack_bf16x2(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(sQ + r * 128 + (((part * (FWDS_OC / 8) + g) ^ (r & 7)) << 4)) = pk;
      }
      fence_proxy_async();
      if (FWDS_SPLIT == 2) pair_sync(); else __syncwarp();
      if (part == 0 && lane == 0 && q0 + rb < sh.Tq) {
        tma_store_3d(&tmO, sQ + rb * 128, h * HD, q0 + rb, b);
        tma_store_commit();
      }
      if (part == 0 && iq < sh.Tq) lse[((int64_t)b * sh.H + h) * sh.Tq + iq] = valid ? (m + log2f(l)) * kLn2 : 0.f;
      if (part == 0 && lane == 0) tma_store_wait_read<0>();
    }
  }
  if (threadIdx.x == 64) AT_MARK(3);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
  if (threadIdx.x == 64) AT_MARK(4);
}

