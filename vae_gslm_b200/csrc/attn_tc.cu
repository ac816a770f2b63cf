// attn_tc.cu — causal ALiBi self-attention forward/backward on the 5th-generation tensor cores (bf16).
// Reference: attention.py:52-78 (dense [B,H,T,T] additive mask + F.scaled_dot_product_attention) and
// position/alibi.py.  Same contract as attn_simt.cu (which stays as the fp32 parity backend).
//
// Forward — one CTA per (batch, head, 128-query tile), two CTAs resident per SM:
//   warp 0 lane 0 : TMA producer: Q once, then K (2-stage ring) and V tiles of 128 keys up to the causal limit
//   warp 1        : TMEM allocator; lane 0 issues tcgen05.mma:  S = Q·Kᵀ (128x128x64)  and  O_j = P·V (128x64x128)
//   warps 2..5    : softmax: thread = query row.  tcgen05.ld brings the row of S into registers, the bias
//                   −slope_h·(i−j) and the (j ≤ i, j < kv_len) mask are applied in registers, online softmax
//                   needs no shuffles, P goes to shared memory (bf16, 128B-swizzled, as the K-major A operand
//                   of the second MMA) and the per-tile O_j is folded into a register accumulator.
// Backward — one CTA per (batch, head, 128-key tile), looping over the query tiles that can see it:
//   S = Q·Kᵀ and dP = dO·Vᵀ land in TMEM; threads form P = exp2(S' − LSE) and dS = P∘(dP − δ)·scale in
//   registers and store both (bf16, swizzled) once; the SAME shared-memory bytes then serve as MN-major A
//   operand for dV += Pᵀ·dO and dK += dSᵀ·Q and as K-major A operand for dQ_i = dS·K (only the descriptor's
//   major bit differs).  dK/dV accumulate in TMEM across the loop; dQ_i is reduced into an fp32 buffer with
//   vector red.global.add (as FlashAttention-2 does) and converted to bf16 by a tail kernel.
#include <math_constants.h>
#include <type_traits>
#include "common.cuh"
#include "sm100.cuh"

namespace vg {
using namespace sm100;

int make_tmap_bf16_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                      int64_t stride2, int box0, int box1);
int make_tmap_f32_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                     int64_t stride2, int box0, int box1);

constexpr int TQ = 128;            // query rows per tile (UMMA M)
constexpr int TK = 128;            // keys per tile
constexpr int HD = 64;             // head dim
constexpr int TILE_BYTES = 128 * HD * 2;      // 16 KB: a [128 x 64] bf16 tile, rows of 128 B, SWIZZLE_128B
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct AttnTcShape {
  int B, H, Tq, Tk, q_offset;
  float scale;
  unsigned long long* trace;      // debug: clock64 phase stamps of CTA (0,0,0), 8 per loop iteration (vg_debug_attn_trace)
};
// slot k of iteration `it`; one elected thread per role stamps.  The traced CTA is the one whose linear block index is
// stored in trace[255] by the host (tools/attn_trace.py); slots 240.. are whole-CTA marks.
#define AT_ON() (sh.trace && (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) == (int)sh.trace[255])
#define AT_STAMP(it, k)                                                    \
  do {                                                                     \
    if (AT_ON() && (it) < 30) sh.trace[(it) * 8 + (k)] = clock64();        \
  } while (0)
#define AT_MARK(k)                                       \
  do {                                                   \
    if (AT_ON()) sh.trace[240 + (k)] = clock64();        \
  } while (0)

// byte offset of the 16-byte piece `piece` (0..15: 8 bf16 each) of row `row` inside a [128 x 128] bf16 tile stored
// as two [128 x 64] SWIZZLE_128B halves (what TMA would produce and what the UMMA descriptors expect)
__device__ __forceinline__ uint32_t sw128_piece(int row, int piece) {
  return (uint32_t)((piece >> 3) * TILE_BYTES + row * 128 + (((piece & 7) ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// K-major A/B operand: tile rows at 128 B pitch; k-step of 16 elements = +32 B inside the 64-wide half
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int kstep) {
  return make_smem_desc_sw128(tile_addr + (uint32_t)((kstep >> 2) * TILE_BYTES + (kstep & 3) * 32), 16, 1024);
}
// MN-major operand: tile stored [k rows x 64 mn] (+ further 64-wide mn halves every TILE_BYTES); k-step = 16 rows
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int kstep) {
  return make_smem_desc_sw128(tile_addr + (uint32_t)(kstep * 2048), TILE_BYTES, 1024);
}

// =========================================================================================== forward
constexpr int FWD_THREADS = 192;
constexpr int FWD_SMEM = 5 * TILE_BYTES /*Q, K0, K1, V*/ + TILE_BYTES /*P second half*/ + 1024 + 256;
// layout: Q | K0 | K1 | V | P(2 halves)

__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   float* __restrict__ lse, const int32_t* __restrict__ kv_len, const float* __restrict__ slopes,
                   AttnTcShape sh) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched kernel behind may start its prologue (vg_set_pdl_mode)
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
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
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
    const int iq = q0 + r;
    const int ia = sh.q_offset + iq;
    const float slope2 = (slopes ? slopes[h] : 0.f) * kLog2e;
    const float scale2 = sh.scale * kLog2e;
    const uint32_t lane_addr = (uint32_t)rb << 16;
    const int lim = min(ia + 1, klen);          // keys [0, lim) are visible to this row
    const float rowc = -slope2 * (float)ia;
    const f32x2_t sc2 = splat2(scale2), sl2 = splat2(slope2);
    float m = -CUDART_INF_F, l = 0.f, corr_prev = 1.f;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    // o = o·corr + O_j  (O_j = P_j·V_j from TMEM)
    auto fold = [&](float corr) {
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tO + lane_addr + c * 32, v);
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
      for (int c = 0; c < TK / 32; ++c) {
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
      for (int c = 0; c < TK / 32; ++c) {
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
      const bool valid = ia < klen && l > 0.f;
      const float inv = valid ? 1.f / l : 0.f;
#pragma unroll
      for (int g = 0; g < HD / 8; ++g) {
        uint4 pk;
        pk.x = pack_bf16x2(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv);
        pk.y = pack_bf16x2(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
        pk.z = pack_bf16x2(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv);
        pk.w = pack_bf16x2(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(sQ + r * 128 + ((g ^ (r & 7)) << 4)) = pk;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && q0 + rb < sh.Tq) {
        tma_store_3d(&tmO, sQ + rb * 128, h * HD, q0 + rb, b);
        tma_store_commit();
      }
      if (iq < sh.Tq) lse[((int64_t)b * sh.H + h) * sh.Tq + iq] = valid ? (m + log2f(l)) * kLn2 : 0.f;
      if (lane == 0) tma_store_wait_read<0>();
    }
  }
  if (threadIdx.x == 64) AT_MARK(3);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
  if (threadIdx.x == 64) AT_MARK(4);
}

// =========================================================================================== backward
constexpr int BWD_EW = 16;                  // elementwise warps: four per TMEM lane quadrant, 32 of the 128 key columns each
constexpr int BWD_THREADS = 64 + 32 * BWD_EW;      // + TMA producer warp + MMA issuer warp
// layout: K | V | Q x3 | dO x3 | P (2 halves) | dS (2 halves) = 12 tiles
constexpr int BWD_QS = 3;                   // Q / dO ring stages: a stage is free only after dV, dK of its tile (the LAST MMAs
                                            // of an iteration); with 2 stages the ~1.3 k cycle reload was exposed every tile
constexpr int BWD_SMEM = (6 + 2 * BWD_QS) * TILE_BYTES + 8 * 4096 /* dQ / dK / dV scratch */ + 1024 + 256;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                   const __grid_constant__ CUtensorMap tmDV, const float* __restrict__ lse,
                   const float* __restrict__ delta,
                   const int32_t* __restrict__ kv_len, const float* __restrict__ slopes, AttnTcShape sh) {
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
           *mma_done = bars + 12, *sdp_free = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 64) AT_MARK(0);
  // 1-D grid, key tile slowest: key tile 0 loops over every query tile, so all its (head, batch) CTAs go first
  const int kt = (int)(blockIdx.x / (unsigned)(sh.H * sh.B));
  const int hb = (int)(blockIdx.x % (unsigned)(sh.H * sh.B));
  const int h = hb % sh.H, b = hb / sh.H;
  const int j0 = kt * TK;
  const int klen = kv_len ? min(kv_len[b], sh.Tk) : sh.Tk;
  // query tiles that can see key j0: q_offset + iq >= j0; rows at/after klen are padded queries (P = 0)
  int i_first = (j0 - sh.q_offset) / TQ;
  if (j0 - sh.q_offset < 0) i_first = 0;
  const int q_valid_end = min(sh.Tq, klen - sh.q_offset);            // queries with ia < klen
  const int n_qt_total = (max(q_valid_end, 0) + TQ - 1) / TQ;
  const int n_it = (j0 < klen) ? max(n_qt_total - i_first, 0) : 0;   // iterations of the query-tile loop

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmdO);
    prefetch_tensormap(&tmDQ); prefetch_tensormap(&tmDK); prefetch_tensormap(&tmDV);
    mbar_init(kv_full, 1); mbar_init(mma_done, 1);
    for (int i = 0; i < BWD_QS; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_full, 32 * BWD_EW); mbar_init(sdp_free, 32 * BWD_EW); mbar_init(&dq_full[0], 1); mbar_init(&dq_full[1], 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
    // this thread is the TMA producer: first tiles requested before the CTA-wide set-up (see the forward)
    if (n_it > 0) {
      mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_3d(sK, &tmK, kv_full, h * HD, j0, b);
      tma_load_3d(sV, &tmV, kv_full, h * HD, j0, b);
      mbar_arrive_expect_tx(&qdo_full[0], 2 * TILE_BYTES);
      tma_load_3d(sQ, &tmQ, &qdo_full[0], h * HD, i_first * TQ, b);
      tma_load_3d(sdO, &tmdO, &qdo_full[0], h * HD, i_first * TQ, b);
    }
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384;      // two buffers of 64 columns: dQ of tile it lives in buffer it & 1
  if (threadIdx.x == 64) AT_MARK(1);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    for (int it = 1; it < n_it; ++it) {
      const int s = it % BWD_QS;
      const int q0 = (i_first + it) * TQ;
      mbar_wait(&qdo_empty[s], ((it / BWD_QS) & 1) ^ 1);
      mbar_arrive_expect_tx(&qdo_full[s], 2 * TILE_BYTES);
      tma_load_3d(sQ + s * TILE_BYTES, &tmQ, &qdo_full[s], h * HD, q0, b);
      tma_load_3d(sdO + s * TILE_BYTES, &tmdO, &qdo_full[s], h * HD, q0, b);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues: see the forward) ==========
    constexpr uint32_t idesc_sp = make_idesc_bf16(128, 128, 0, 0);    // S = Q·Kᵀ, dP = dO·Vᵀ
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);      // dV = Pᵀ·dO, dK = dSᵀ·Q (A, B MN-major)
    constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);      // dQ = dS·K       (A K-major, B MN-major)
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), adS = smem_u32(sdS);
    // tensor-core order:  S, dP of tile it+1 as soon as the elementwise warps hold S, dP of tile `it` in registers
    // (sdp_free; their wait for the previous MMAs and the 64 KB burst of P / dS stores overlaps these MMAs)  →  once
    // P, dS are published: dQ_it  →  dV, dK of tile it.  The N = 64 MMAs read 6 KB of shared memory per 32 tensor cycles and are
    // shared-memory-bandwidth bound (~48 cycles each).
    auto issue_sdp = [&](int it) {
      const int s = it % BWD_QS;
      const uint32_t aQ = smem_u32(sQ + s * TILE_BYTES), adO = smem_u32(sdO + s * TILE_BYTES);
      mbar_wait(&qdo_full[s], (it / BWD_QS) & 1);
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
    if (n_it > 0) {
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      AT_STAMP(0, 5);
    }
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;                        // dQ buffer
      const int sq = it % BWD_QS;                  // Q / dO stage
      const uint32_t aQ = smem_u32(sQ + sq * TILE_BYTES), adO = smem_u32(sdO + sq * TILE_BYTES);
      if (it + 1 < n_it) {
        mbar_wait(sdp_free, it & 1);               // S, dP of tile `it` are in the elementwise warps' registers
        tc_fence_after();
        issue_sdp(it + 1);
        AT_STAMP(it + 1, 5);
      }
      mbar_wait(pds_full, it & 1);                 // P and dS in shared memory; dQ_{it-2} drained
      tc_fence_after();
      AT_STAMP(it, 6);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < TK / 16; ++k)
          umma_f16_ss(tdQ + s * 64, desc_kmajor(adS, k), desc_mnmajor(aK, k), idesc_q, k != 0);
        umma_commit(&dq_full[s]);
      }
      __syncwarp();
      AT_STAMP(it, 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < TQ / 16; ++k)
          umma_f16_ss(tdV, desc_mnmajor(aP, k), desc_mnmajor(adO, k), idesc_t, (it | k) != 0);
#pragma unroll
        for (int k = 0; k < TQ / 16; ++k)
          umma_f16_ss(tdK, desc_mnmajor(adS, k), desc_mnmajor(aQ, k), idesc_t, (it | k) != 0);
        umma_commit(&qdo_empty[sq]);
        umma_commit(mma_done);                     // P / dS of this tile may be overwritten
      }
      __syncwarp();
      AT_STAMP(it, 7);
    }
    if (n_it > 0) {
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else if (warp >= 2) {
    // ===================== softmax-backward threads (thread = query row of the current tile) =====================
    // warps 2..17: TMEM lane quadrant = warp % 4 (hardware rule), column part = (warp - 2) / 4.  With one warp per
    // scheduler the elementwise phase was latency-bound (~10 k cycles per tile pair against 1.3 k cycles of MMA).
    const int rb = (warp & 3) * 32;
    const int part = (warp - 2) >> 2;           // which 32 of the 128 key columns
    const int half = part & 1;                  // dQ drain / dK, dV epilogue: parts 0 and 1 only, 32 columns each
    const bool drains = part < 2;
    const int r = rb + lane;
    const uint32_t lane_addr = (uint32_t)rb << 16;
    const float slope2 = (slopes ? slopes[h] : 0.f) * kLog2e;
    const float scale2 = sh.scale * kLog2e;
    const float* lse_row = lse + ((int64_t)b * sh.H + h) * sh.Tq;
    const float* delta_row = delta + ((int64_t)b * sh.H + h) * sh.Tq;
    // lse / delta of the NEXT query tile are fetched one iteration ahead: loaded at the top of the iteration they were
    // needed in, these two global loads were 30 % of the kernel's stall samples (one exposed L2 round trip per tile pair)
    float lse_next = 0.f, delta_next = 0.f;
    if (n_it > 0) {
      const int iq0 = i_first * TQ + r;
      if (iq0 < sh.Tq && sh.q_offset + iq0 < klen) { lse_next = lse_row[iq0]; delta_next = delta_row[iq0]; }
    }
    // dQ partial of query tile `t` (this key tile's contribution) → fp32 accumulator in global memory.  Each warp writes
    // its 32 rows x 32 columns (thread = row, 128-byte rows, 16-byte pieces XOR-swizzled: exactly the SWIZZLE_128B box
    // layout) into a private 4 KB scratch and ONE lane hands the box to the TMA unit as a bulk reduce-add
    // (cp.reduce.async.bulk.tensor); rows beyond Tq are clipped by the tensor map.  The first version read the scratch
    // back and issued eight 16-byte vector atomics per thread — 1.3-1.7 k cycles per tile on the loop's critical path.
    // It runs one tile late (dQ is double-buffered in TMEM), in the bubble after P / dS of the next tile are published.
    auto drain_dq = [&](int t) {
      if (!drains) return;                      // warp-uniform
      uint8_t* scr = sDQ + (warp - 2) * 4096;
      mbar_wait(&dq_full[t & 1], (t >> 1) & 1);
      tc_fence_after();
      if (lane == 0) tma_store_wait_read<0>();        // the previous reduce has finished reading the scratch
      __syncwarp();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tdQ + (t & 1) * 64 + lane_addr + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(scr + lane * 128 + ((g ^ (lane & 7)) << 4)) =
            make_uint4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmDQ, scr, h * HD + half * 32, (i_first + t) * TQ + rb, b);
        tma_store_commit();
      }
    };
    for (int it = 0; it < n_it; ++it) {
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
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      if (threadIdx.x == 64) AT_STAMP(it, 0);
      // (key tile, query tile) pairs entirely below the diagonal with every row / key valid need no mask: packed
      // f32x2 arithmetic, two elements per FFMA2 / FMUL2 (the elementwise threads are issue-bound, see the forward)
      const bool full_pair = (j0 + TK - 1 <= sh.q_offset + q0) && (j0 + TK <= klen) && (q0 + TQ <= sh.Tq) &&
                             (sh.q_offset + q0 + TQ <= klen);                         // CTA-uniform
      const f32x2_t sc2 = splat2(scale2), sl2 = splat2(slope2), scl = splat2(sh.scale), ndl = splat2(-dl * sh.scale);
      const float rowc = -slope2 * (float)ia - L2;
      const int lim = row_ok ? min(ia + 1, klen) : 0;             // keys [0, lim) are visible to this row
      // P = exp2(S·scale2 + slope2·(j − i) − LSE2), dS = P·(dP − δ)·scale, two elements per packed FFMA2 / FMUL2 (the
      // elementwise threads are issue- and XU-bound).  MASKED pairs zero P for columns ≥ lim with a select; the first
      // version converted column indices to float per element on the XU pipe.
      // The results stay in registers (64 columns x {P, dS} as packed bf16 pairs) until the previous tile's dQ / dV /
      // dK MMAs have finished reading P and dS from shared memory: only the short burst of stores waits for them, the
      // exponentials of this tile overlap them.
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
      if (threadIdx.x == 64) AT_STAMP(it, 1);
      if (it > 0) mbar_wait(mma_done, (it - 1) & 1);  // the previous tile's MMAs no longer read P / dS
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t off = sw128_piece(r, (part * 2 + cc) * 2 + g);
          *reinterpret_cast<uint4*>(sP + off) =
              make_uint4(pkp[cc * 8 + g * 4 + 0], pkp[cc * 8 + g * 4 + 1], pkp[cc * 8 + g * 4 + 2], pkp[cc * 8 + g * 4 + 3]);
          *reinterpret_cast<uint4*>(sdS + off) =
              make_uint4(pkd[cc * 8 + g * 4 + 0], pkd[cc * 8 + g * 4 + 1], pkd[cc * 8 + g * 4 + 2], pkd[cc * 8 + g * 4 + 3]);
        }
      }
      fence_proxy_async();
      mbar_arrive(pds_full);
      if (threadIdx.x == 64) AT_STAMP(it, 2);
      // while the tensor core forms S, dP of the next tile: dQ of the PREVIOUS tile (the other TMEM buffer; dQ of this
      // tile is only being issued now) goes to the global accumulator
      if (it > 0) drain_dq(it - 1);
      if (threadIdx.x == 64) AT_STAMP(it, 3);
    }
    if (n_it > 0) drain_dq(n_it - 1);
    if (threadIdx.x == 64) AT_MARK(2);
    // ---- dK / dV of this key tile (thread = key row): the `half` 0 warps store dV, the `half` 1 warps dK — each its
    // TMEM quadrant's 32 rows x 64 columns, bf16, through its 4 KB scratch (SWIZZLE_128B box) and one TMA store; rows
    // beyond Tk are clipped by the tensor map
    if (n_it > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    if (drains) {
      uint8_t* scr = sDQ + (warp - 2) * 4096;
      if (lane == 0) tma_store_wait_read<0>();        // the last dQ reduce has finished reading the scratch
      __syncwarp();
      const uint32_t tsrc = (half == 0 ? tdV : tdK) + lane_addr;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        if (n_it > 0) {
          tmem_ld_32x32b_x32(tsrc + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
          *reinterpret_cast<uint4*>(scr + lane * 128 + (((c * 4 + g) ^ (lane & 7)) << 4)) = pk;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && j0 + rb < sh.Tk) {
        tma_store_3d(half == 0 ? &tmDV : &tmDK, scr, h * HD, j0 + rb, b);
        tma_store_commit();
      }
      if (lane == 0) tma_store_wait_all();             // this lane's dQ reductions and its dK / dV store are complete
    }
    if (threadIdx.x == 64) AT_MARK(5);
  }
  if (threadIdx.x == 64) AT_MARK(3);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (threadIdx.x == 64) AT_MARK(4);
}

// Persistent backward (one CTA per SM walking a heavy-first item list, next item's K/V prefetched under the dK/dV
// epilogue): measured on B200 112 vs 121 us at B=8,T=1000 and 65 vs 73 us at T=640 (profiles/r02_variants.md); default.
// -DVG_ATTN_BWD_PERSIST=0 selects the one-CTA-per-key-tile kernel above.
#ifndef VG_ATTN_BWD_PERSIST
#define VG_ATTN_BWD_PERSIST 1
#endif
#if VG_ATTN_BWD_PERSIST
#include "attn_tc_bwd_persist.cuh"
#endif

// delta[b,h,i] = Σ_d dO[i,d]·O[i,d]   (bf16 inputs)
__global__ void attn_tc_delta_kernel(const __nv_bfloat16* __restrict__ dout, int64_t ld_dout,
                                     const __nv_bfloat16* __restrict__ out, int64_t ld_out, float* __restrict__ delta,
                                     int B, int H, int Tq) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched kernel behind may start its prologue (vg_set_pdl_mode)
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;      // 8 lanes per (b,h,i) row
  const int sub = threadIdx.x & 7;
  if (w >= (int64_t)B * H * Tq) return;     // whole 8-lane groups exit together; shuffles below stay in-group
  const int i = (int)(w % Tq);
  const int h = (int)((w / Tq) % H);
  const int b = (int)(w / ((int64_t)Tq * H));
  Vec8<__nv_bfloat16> a, o;
  a.load(dout + ((int64_t)b * Tq + i) * ld_dout + h * HD + sub * 8);
  o.load(out + ((int64_t)b * Tq + i) * ld_out + h * HD + sub * 8);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s = fmaf(a.v[e], o.v[e], s);
  const unsigned grp_mask = 0xffu << ((threadIdx.x & 31) & ~7);
  s += __shfl_xor_sync(grp_mask, s, 4);
  s += __shfl_xor_sync(grp_mask, s, 2);
  s += __shfl_xor_sync(grp_mask, s, 1);
  if (sub == 0) delta[w] = s;
}

__global__ void attn_tc_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int64_t ld_dq,
                                          int64_t rows, int C) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched kernel behind may start its prologue (vg_set_pdl_mode)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // over rows * C/8
  const int c8 = C / 8;
  if (i >= rows * c8) return;
  const int64_t row = i / c8;
  const int c = (int)(i % c8) * 8;
  Vec8<float> a;
  a.load(acc + row * C + c);
  Vec8<__nv_bfloat16> o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = a.v[e];
  o.store(dq + row * ld_dq + c);
}

// ---- host side ------------------------------------------------------------------------------------
static unsigned long long* g_attn_trace = nullptr;
void attn_tc_set_trace(void* buf) { g_attn_trace = (unsigned long long*)buf; }
bool attn_tc_supported(int dtype, int64_t D, int64_t ld_q, int64_t ld_kv, const void* q, const void* k, const void* v,
                       int64_t kv_batch_stride, int64_t kv_head_stride) {
  return dtype == VG_BF16 && D == HD && ld_q % 8 == 0 && ld_kv % 8 == 0 && aligned(q, 16) && aligned(k, 16) &&
         aligned(v, 16) && kv_batch_stride == 0 && kv_head_stride == 0;
}

int attn_tc_fwd_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out,
                       int64_t ld_out, float* lse, const int32_t* kv_len, const float* slopes, int B, int H, int Tq,
                       int Tk, int q_offset, float scale, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmQ, q, (int64_t)H * HD, Tq, B, ld_q, (int64_t)Tq * ld_q, HD, TQ))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmK, k, (int64_t)H * HD, Tk, B, ld_kv, (int64_t)Tk * ld_kv, HD, TK))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmV, v, (int64_t)H * HD, Tk, B, ld_kv, (int64_t)Tk * ld_kv, HD, TK))) return rc;
  CUtensorMap tmO;           // output rows leave as one [32 x 64] box per warp
  if ((rc = make_tmap_bf16_3d(&tmO, out, (int64_t)H * HD, Tq, B, ld_out, (int64_t)Tq * ld_out, HD, 32))) return rc;
  AttnTcShape sh{B, H, Tq, Tk, q_offset, scale, g_attn_trace};
  dim3 grid((unsigned)((Tq + TQ - 1) / TQ) * (unsigned)H * (unsigned)B);
  static bool set = false;
  if (!set) {
    VG_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    set = true;
  }
  attn_tc_fwd_kernel<<<grid, FWD_THREADS, FWD_SMEM, st>>>(tmQ, tmK, tmV, tmO, lse, kv_len, slopes, sh);
  VG_LAUNCH_CHECK("vg_attn_fwd(tcgen05)");
  return 0;
}

size_t attn_tc_bwd_workspace(int64_t B, int64_t H, int64_t Tq) {
  return align_up((size_t)(B * H * Tq) * sizeof(float), 256) + (size_t)(B * Tq * H * HD) * sizeof(float);
}

int attn_tc_bwd_launch(const void* dout, int64_t ld_dout, const void* q, const void* k, const void* v, int64_t ld_q,
                       int64_t ld_kv, const void* out, int64_t ld_out, const float* lse, void* dq, void* dk, void* dv,
                       int64_t ld_dq, int64_t ld_dkv, const int32_t* kv_len, const float* slopes, int B, int H, int Tq,
                       int Tk, int q_offset, float scale, void* workspace, cudaStream_t st) {
  float* delta = (float*)workspace;
  float* dq_acc = (float*)((uint8_t*)workspace + align_up((size_t)B * H * Tq * sizeof(float), 256));
  const int64_t nrows = (int64_t)B * H * Tq;
  attn_tc_delta_kernel<<<(unsigned)ceil_div(nrows * 8, 256), 256, 0, st>>>((const __nv_bfloat16*)dout, ld_dout,
                                                                          (const __nv_bfloat16*)out, ld_out, delta, B,
                                                                          H, Tq);
  VG_LAUNCH_CHECK("vg_attn_bwd(tcgen05 delta)");
  VG_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * Tq * H * HD * sizeof(float), st));
  CUtensorMap tmQ, tmK, tmV, tmdO;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmQ, q, (int64_t)H * HD, Tq, B, ld_q, (int64_t)Tq * ld_q, HD, TQ))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmK, k, (int64_t)H * HD, Tk, B, ld_kv, (int64_t)Tk * ld_kv, HD, TK))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmV, v, (int64_t)H * HD, Tk, B, ld_kv, (int64_t)Tk * ld_kv, HD, TK))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmdO, dout, (int64_t)H * HD, Tq, B, ld_dout, (int64_t)Tq * ld_dout, HD, TQ))) return rc;
  CUtensorMap tmDQ;          // fp32 dQ accumulator [B, Tq, H*64]: reduce-add boxes of 32 rows x 32 columns
  if ((rc = make_tmap_f32_3d(&tmDQ, dq_acc, (int64_t)H * HD, Tq, B, (int64_t)H * HD, (int64_t)Tq * H * HD, 32, 32)))
    return rc;
  CUtensorMap tmDK, tmDV;    // dK / dV rows leave as [32 x 64] boxes
  if ((rc = make_tmap_bf16_3d(&tmDK, dk, (int64_t)H * HD, Tk, B, ld_dkv, (int64_t)Tk * ld_dkv, HD, 32))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmDV, dv, (int64_t)H * HD, Tk, B, ld_dkv, (int64_t)Tk * ld_dkv, HD, 32))) return rc;
  AttnTcShape sh{B, H, Tq, Tk, q_offset, scale, g_attn_trace};
  static bool set = false;
#if VG_ATTN_BWD_PERSIST
  static int sms = 0;
  if (!set) {
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    VG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VG_CUDA(cudaFuncSetAttribute(attn_tc_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    set = true;
  }
  const int n_items = ((Tk + TK - 1) / TK) * H * B;
  attn_tc_bwd_persist_kernel<<<(unsigned)(n_items < sms ? n_items : sms), BWD_THREADS, BWD_SMEM, st>>>(
      tmQ, tmK, tmV, tmdO, tmDQ, tmDK, tmDV, lse, delta, kv_len, slopes, sh);
#else
  if (!set) {
    VG_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    set = true;
  }
  dim3 grid((unsigned)((Tk + TK - 1) / TK) * (unsigned)H * (unsigned)B);
  attn_tc_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM, st>>>(tmQ, tmK, tmV, tmdO, tmDQ, tmDK, tmDV, lse, delta, kv_len,
                                                          slopes, sh);
#endif
  VG_LAUNCH_CHECK("vg_attn_bwd(tcgen05)");
  const int C = H * HD;
  const int64_t n = (int64_t)B * Tq * (C / 8);
  attn_tc_dq_convert_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dq_acc, (__nv_bfloat16*)dq, ld_dq,
                                                                       (int64_t)B * Tq, C);
  VG_LAUNCH_CHECK("vg_attn_bwd(tcgen05 dq convert)");
  return 0;
}

}  // namespace vg
