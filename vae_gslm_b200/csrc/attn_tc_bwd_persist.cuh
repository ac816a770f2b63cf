// attn_tc_bwd_persist.cuh — the default attention backward (-DVG_ATTN_BWD_PERSIST=0 disables; included by attn_tc.cu inside
// namespace vg after attn_tc_bwd_kernel): the attention backward with PERSISTENT CTAs.  Per-CTA fixed costs of the product kernel
// (first S / dP ready ~4 k cycles after CTA entry, dK / dV epilogue + exit ~2.5 k: ~28 % of an average CTA,
// profiles/r01_attention_v2.md) are hidden by letting the TMA producer and the MMA issuer run ahead into the next work
// item while the elementwise warps finish the current one.  Same data flow, shared-memory / TMEM layout and inner loops as
// attn_tc_bwd_kernel; what changes is that every mbarrier parity and ring stage derives from running counters and that
// two barriers (kv_empty, acc_empty) hand K / V and the dK / dV accumulators back across items.  Parity: tests/test_kernels_gpu.py (attention
// forward/backward vs fp32 torch, ragged lengths); 112 vs 121 us per layer at B=8, T=1000 (profiles/r02_variants.md).
#pragma once

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_bwd_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                   const __grid_constant__ CUtensorMap tmDV, const float* __restrict__ lse,
                   const float* __restrict__ delta,
                   const int32_t* __restrict__ kv_len, const float* __restrict__ slopes, AttnTcShape sh) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched kernel behind may start its prologue (vg_set_pdl_mode)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;                       // BWD_QS stages
  uint8_t* sdO = smem + (2 + BWD_QS) * TILE_BYTES;           // BWD_QS stages
  uint8_t* sP = smem + (2 + 2 * BWD_QS) * TILE_BYTES;        // 2 halves
  uint8_t* sdS = smem + (4 + 2 * BWD_QS) * TILE_BYTES;       // 2 halves
  uint8_t* sDQ = smem + (6 + 2 * BWD_QS) * TILE_BYTES;       // 8 warps x 4 KB dQ transpose scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDQ + 8 * 4096);
  uint64_t *kv_full = bars, *qdo_full = bars + 1 /* 3 */, *qdo_empty = bars + 4 /* 3 */, *sdp_full = bars + 7,
           *pds_full = bars + 8, *dq_full = bars + 9 /* 2: one per dQ buffer */, *acc_full = bars + 11,
           *mma_done = bars + 12, *sdp_free = bars + 13, *kv_empty = bars + 14, *acc_empty = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work items = (key tile, head, batch), key tile slowest (heavy first); CTA c takes items c, c + gridDim.x, …  Every role
  // walks the same list and keeps RUNNING counters (g: (key tile, query tile) iterations, kc: non-empty items), from which
  // all mbarrier parities and ring stages derive, so the pipelines run on across item boundaries: the producer fetches
  // the next item's K / V / Q / dO and the tensor core forms its first S / dP while the elementwise warps still store the
  // previous item's dK / dV.
  struct Item { int kt, h, b, j0, klen, i_first, n_it, q_valid_end; };
  const int n_items = ((sh.Tk + TK - 1) / TK) * sh.H * sh.B;
  auto item_of = [&](int w) {
    Item it;
    it.kt = w / (sh.H * sh.B);
    const int hb = w % (sh.H * sh.B);
    it.h = hb % sh.H;
    it.b = hb / sh.H;
    it.j0 = it.kt * TK;
    it.klen = kv_len ? min(kv_len[it.b], sh.Tk) : sh.Tk;
    it.i_first = (it.j0 - sh.q_offset) / TQ;
    if (it.j0 - sh.q_offset < 0) it.i_first = 0;
    it.q_valid_end = min(sh.Tq, it.klen - sh.q_offset);
    const int n_qt_total = (max(it.q_valid_end, 0) + TQ - 1) / TQ;
    it.n_it = (it.j0 < it.klen) ? max(n_qt_total - it.i_first, 0) : 0;
    return it;
  };

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
    prefetch_tensormap(&tmDQ); prefetch_tensormap(&tmDK); prefetch_tensormap(&tmDV);
    mbar_init(kv_full, 1); mbar_init(mma_done, 1);
    for (int i = 0; i < BWD_QS; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_full, 32 * BWD_EW); mbar_init(sdp_free, 32 * BWD_EW); mbar_init(&dq_full[0], 1); mbar_init(&dq_full[1], 1);
    mbar_init(acc_full, 1); mbar_init(kv_empty, 1); mbar_init(acc_empty, 32 * 8);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384;      // two buffers of 64 columns: dQ of tile it lives in buffer it & 1

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int g = 0, kc = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item im = item_of(w);
      if (im.n_it == 0) continue;
      if (kc > 0) mbar_wait(kv_empty, (kc - 1) & 1);          // the previous item's MMAs no longer read K / V
      mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_3d(sK, &tmK, kv_full, im.h * HD, im.j0, im.b);
      tma_load_3d(sV, &tmV, kv_full, im.h * HD, im.j0, im.b);
      for (int it = 0; it < im.n_it; ++it, ++g) {
        const int s = g % BWD_QS;
        const int q0 = (im.i_first + it) * TQ;
        mbar_wait(&qdo_empty[s], ((g / BWD_QS) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[s], 2 * TILE_BYTES);
        tma_load_3d(sQ + s * TILE_BYTES, &tmQ, &qdo_full[s], im.h * HD, q0, im.b);
        tma_load_3d(sdO + s * TILE_BYTES, &tmdO, &qdo_full[s], im.h * HD, q0, im.b);
      }
      ++kc;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loops, one elected lane issues) =====================
    constexpr uint32_t idesc_sp = make_idesc_bf16(128, 128, 0, 0);    // S = Q·Kᵀ, dP = dO·Vᵀ
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);      // dV = Pᵀ·dO, dK = dSᵀ·Q (A, B MN-major)
    constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);      // dQ = dS·K       (A K-major, B MN-major)
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), adS = smem_u32(sdS);
    auto issue_sdp = [&](int gi) {                 // S, dP of global iteration gi (its Q / dO stage = gi % BWD_QS)
      const int s = gi % BWD_QS;
      const uint32_t aQ = smem_u32(sQ + s * TILE_BYTES), adO = smem_u32(sdO + s * TILE_BYTES);
      mbar_wait(&qdo_full[s], (gi / BWD_QS) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tS, desc_kmajor(aQ, k), desc_kmajor(aK, k), idesc_sp, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tdP, desc_kmajor(adO, k), desc_kmajor(aV, k), idesc_sp, k != 0);
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    int g = 0, kc = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item im = item_of(w);
      if (im.n_it == 0) continue;
      mbar_wait(kv_full, kc & 1);
      if (g > 0) mbar_wait(sdp_free, (g - 1) & 1);      // S, dP of the previous item's last tile are in registers
      tc_fence_after();
      issue_sdp(g);
      for (int it = 0; it < im.n_it; ++it, ++g) {
        const int s = g & 1;                             // dQ buffer
        const int sq = g % BWD_QS;                       // Q / dO stage
        const uint32_t aQ = smem_u32(sQ + sq * TILE_BYTES), adO = smem_u32(sdO + sq * TILE_BYTES);
        const bool last = it + 1 == im.n_it;
        if (!last) {
          mbar_wait(sdp_free, g & 1);
          tc_fence_after();
          issue_sdp(g + 1);
        }
        mbar_wait(pds_full, g & 1);                      // P and dS in shared memory; dQ_{g-2} drained
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TK / 16; ++k)
            umma_f16_ss(tdQ + s * 64, desc_kmajor(adS, k), desc_mnmajor(aK, k), idesc_q, k != 0);
          umma_commit(&dq_full[s]);
        }
        __syncwarp();
        if (it == 0 && kc > 0) {                         // dK / dV of the previous item have left TMEM
          mbar_wait(acc_empty, (kc - 1) & 1);
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TQ / 16; ++k)
            umma_f16_ss(tdV, desc_mnmajor(aP, k), desc_mnmajor(adO, k), idesc_t, (it | k) != 0);
#pragma unroll
          for (int k = 0; k < TQ / 16; ++k)
            umma_f16_ss(tdK, desc_mnmajor(adS, k), desc_mnmajor(aQ, k), idesc_t, (it | k) != 0);
          umma_commit(&qdo_empty[sq]);
          umma_commit(mma_done);                         // P / dS of this tile may be overwritten
          if (last) {
            umma_commit(kv_empty);                       // K / V of this item may be overwritten
            umma_commit(acc_full);                       // dK / dV of this item are complete
          }
        }
        __syncwarp();
      }
      ++kc;
    }
  } else if (warp >= 2) {
    // ===================== elementwise warps (thread = query row of the current tile) =====================
    const int rb = (warp & 3) * 32;
    const int part = (warp - 2) >> 2;           // which 32 of the 128 key columns
    const int half = part & 1;                  // dQ drain / dK, dV epilogue: parts 0 and 1 only, 32 columns each
    const bool drains = part < 2;
    const int r = rb + lane;
    const uint32_t lane_addr = (uint32_t)rb << 16;
    const float scale2 = sh.scale * kLog2e;
    uint8_t* scr = sDQ + ((warp - 2) & 7) * 4096;       // (used by the draining warps only)
    int g = 0, kc = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const Item im = item_of(w);
      const int h = im.h, b = im.b, j0 = im.j0, klen = im.klen, i_first = im.i_first, n_it = im.n_it;
      const float slope2 = (slopes ? slopes[h] : 0.f) * kLog2e;
      const float* lse_row = lse + ((int64_t)b * sh.H + h) * sh.Tq;
      const float* delta_row = delta + ((int64_t)b * sh.H + h) * sh.Tq;
      float lse_next = 0.f, delta_next = 0.f;
      if (n_it > 0) {
        const int iq0 = i_first * TQ + r;
        if (iq0 < sh.Tq && sh.q_offset + iq0 < klen) { lse_next = lse_row[iq0]; delta_next = delta_row[iq0]; }
      }
      // dQ partial of this item's query tile `t` (global iteration gi) → TMA bulk reduce-add, see attn_tc_bwd_kernel
      auto drain_dq = [&](int t, int gi) {
        if (!drains) return;                      // warp-uniform
        mbar_wait(&dq_full[gi & 1], (gi >> 1) & 1);
        tc_fence_after();
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        uint32_t v[32];
        tmem_ld_32x32b_x32(tdQ + (gi & 1) * 64 + lane_addr + half * 32, v);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int gq = 0; gq < 8; ++gq)
          *reinterpret_cast<uint4*>(scr + lane * 128 + ((gq ^ (lane & 7)) << 4)) =
              make_uint4(v[gq * 4 + 0], v[gq * 4 + 1], v[gq * 4 + 2], v[gq * 4 + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_3d(&tmDQ, scr, h * HD + half * 32, (i_first + t) * TQ + rb, b);
          tma_store_commit();
        }
      };
      for (int it = 0; it < n_it; ++it, ++g) {
        const int q0 = (i_first + it) * TQ;
        const int iq = q0 + r;
        const int ia = sh.q_offset + iq;
        const bool row_ok = iq < sh.Tq && ia < klen;
        const float L2 = row_ok ? lse_next * kLog2e : 0.f;
        const float dl = row_ok ? delta_next : 0.f;
        {
          const int iqn = iq + TQ;
          lse_next = 0.f; delta_next = 0.f;
          if (it + 1 < n_it && iqn < sh.Tq && sh.q_offset + iqn < klen) { lse_next = lse_row[iqn]; delta_next = delta_row[iqn]; }
        }
        mbar_wait(sdp_full, g & 1);
        tc_fence_after();
        const bool full_pair = (j0 + TK - 1 <= sh.q_offset + q0) && (j0 + TK <= klen) && (q0 + TQ <= sh.Tq) &&
                               (sh.q_offset + q0 + TQ <= klen);                         // CTA-uniform
        const f32x2_t sc2 = splat2(scale2), sl2 = splat2(slope2), scl = splat2(sh.scale), ndl = splat2(-dl * sh.scale);
        const float rowc = -slope2 * (float)ia - L2;
        const int lim = row_ok ? min(ia + 1, klen) : 0;             // keys [0, lim) are visible to this row
        uint32_t pkp[16], pkd[16];
        auto pair_tile = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = part * 2 + cc;
            uint32_t vs[16], vp[16];
            tmem_ld_32x32b_x16(tS + lane_addr + c * 16, vs);
            tmem_ld_32x32b_x16(tdP + lane_addr + c * 16, vp);
            tmem_ld_wait();
            const f32x2_t cb0 = splat2(fmaf(slope2, (float)(j0 + c * 16), rowc));
            const int nvalid = lim - (j0 + c * 16);
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const f32x2_t cb = fma2(sl2, pack2((float)e, (float)(e + 1)), cb0);
              float a, bq, p0, p1, d0, d1;
              unpack2(fma2(pack2(__uint_as_float(vs[e]), __uint_as_float(vs[e + 1])), sc2, cb), a, bq);
              p0 = ex2_approx(a);
              p1 = ex2_approx(bq);
              if (MASKED) {
                p0 = (e < nvalid) ? p0 : 0.f;
                p1 = (e + 1 < nvalid) ? p1 : 0.f;
              }
              const f32x2_t g2 = fma2(pack2(__uint_as_float(vp[e]), __uint_as_float(vp[e + 1])), scl, ndl);
              unpack2(mul2(pack2(p0, p1), g2), d0, d1);
              pkp[cc * 8 + e / 2] = pack_bf16x2(p0, p1);
              pkd[cc * 8 + e / 2] = pack_bf16x2(d0, d1);
            }
          }
        };
        if (full_pair) pair_tile(std::false_type{}); else pair_tile(std::true_type{});
        tc_fence_before();                              // S / dP of this tile are consumed
        mbar_arrive(sdp_free);
        if (g > 0) mbar_wait(mma_done, (g - 1) & 1);    // the previous tile's MMAs (of this or the previous item) are done
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {
            const uint32_t off = sw128_piece(r, (part * 2 + cc) * 2 + gq);
            *reinterpret_cast<uint4*>(sP + off) =
                make_uint4(pkp[cc * 8 + gq * 4 + 0], pkp[cc * 8 + gq * 4 + 1], pkp[cc * 8 + gq * 4 + 2], pkp[cc * 8 + gq * 4 + 3]);
            *reinterpret_cast<uint4*>(sdS + off) =
                make_uint4(pkd[cc * 8 + gq * 4 + 0], pkd[cc * 8 + gq * 4 + 1], pkd[cc * 8 + gq * 4 + 2], pkd[cc * 8 + gq * 4 + 3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(pds_full);
        if (it > 0) drain_dq(it - 1, g - 1);
      }
      if (n_it > 0) drain_dq(n_it - 1, g - 1);
      // ---- dK / dV of this item (thread = key row), see attn_tc_bwd_kernel
      if (n_it > 0) {
        mbar_wait(acc_full, kc & 1);
        tc_fence_after();
      }
      if (drains) {
        if (lane == 0) tma_store_wait_read<0>();        // the last dQ reduce has finished reading the scratch
        __syncwarp();
        const uint32_t tsrc = (half == 0 ? tdV : tdK) + lane_addr;
        uint32_t v0[32], v1[32];
        if (n_it > 0) {
          tmem_ld_32x32b_x32(tsrc, v0);
          tmem_ld_32x32b_x32(tsrc + 32, v1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(acc_empty);                        // the next item's dV / dK MMAs may overwrite the accumulators
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) { v0[e] = 0u; v1[e] = 0u; }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const uint32_t* v = c == 0 ? v0 : v1;
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(v[gq * 8 + 0]), __uint_as_float(v[gq * 8 + 1]));
            pk.y = pack_bf16x2(__uint_as_float(v[gq * 8 + 2]), __uint_as_float(v[gq * 8 + 3]));
            pk.z = pack_bf16x2(__uint_as_float(v[gq * 8 + 4]), __uint_as_float(v[gq * 8 + 5]));
            pk.w = pack_bf16x2(__uint_as_float(v[gq * 8 + 6]), __uint_as_float(v[gq * 8 + 7]));
            *reinterpret_cast<uint4*>(scr + lane * 128 + (((c * 4 + gq) ^ (lane & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && j0 + rb < sh.Tk) {
          tma_store_3d(half == 0 ? &tmDV : &tmDK, scr, h * HD, j0 + rb, b);
          tma_store_commit();
        }
      }
      if (n_it > 0) ++kc;
    }
    if (drains && lane == 0) tma_store_wait_all();      // this lane's reductions and stores are complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}
