// attn_decode.cu — KV-cache attention for the cached LVTR.step loop (lvtr.py:253-257 →
// attention.py:56-85).  The reference concatenates the whole K/V per layer per step and rebuilds
// an O(Tk^2) mask that it then slices to its last row; here the cache is a pre-allocated head-major
// [B,H,Tmax,D] buffer, the new token's k/v are appended in the same kernel that attends, and the
// ALiBi bias -slope_h*(pos-j) is computed in registers.
//
// HBM-bound (reads Tk*D*2 elements per (b,h)): 8 lanes share one key row with 16-byte loads, so a
// warp instruction covers 4 full 128-byte (bf16) rows; split-KV fills the machine at small batch.
#include <math_constants.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace vg {

constexpr int DD = 64;            // head dim
constexpr int DEC_WARPS = 4;
constexpr int DEC_GROUPS = DEC_WARPS * 4;   // key groups per CTA iteration (8 lanes each)

// 8 elements of a row as loaded (16 bytes of bf16 / 32 bytes of f32): kept raw while several keys are in flight
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = r; }
  __device__ __forceinline__ void zero() { r = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

template <typename T>
__global__ void __launch_bounds__(DEC_WARPS * 32, 6)
attn_decode_kernel(const T* __restrict__ qkv, T* __restrict__ k_cache, T* __restrict__ v_cache,
                   T* __restrict__ out, float* __restrict__ partial, const float* __restrict__ slopes,
                   int H, int Tmax, int pos_arg, const int32_t* __restrict__ pos_dev, int splits, float scale,
                   int* __restrict__ tickets) {
  __shared__ float sm_m[DEC_GROUPS], sm_l[DEC_GROUPS];
  __shared__ int s_last;
  // programmatic dependent launch: the weight-streaming GEMM that follows (decode_linear.cu) may start prefetching its
  // weight slab now; it still waits (griddepcontrol.wait) for this grid to finish before it reads our output
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int pos = pos_dev ? min(*pos_dev, Tmax - 1) : pos_arg;   // device-resident position for graph replay
  __shared__ float sm_o[DEC_GROUPS][DD];
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp * 4 + (lane >> 3);     // 0..15
  const int sub = lane & 7;                   // dims sub*8 .. sub*8+7
  const int HD = H * DD;
  const float slope = slopes ? slopes[h] : 0.f;

  const T* qrow = qkv + (int64_t)b * 3 * HD + h * DD;
  Vec8<T> qv;
  qv.load(qrow + sub * 8);
  const T* knew = qrow + HD;
  const T* vnew = qrow + 2 * HD;
  T* kc = k_cache + ((int64_t)b * H + h) * (int64_t)Tmax * DD;
  T* vc = v_cache + ((int64_t)b * H + h) * (int64_t)Tmax * DD;

  const int nkeys = pos + 1;
  const int chunk = (nkeys + splits - 1) / splits;
  const int j_begin = split * chunk;
  const int j_end = min(nkeys, j_begin + chunk);

  float m = -CUDART_INF_F, l = 0.f, acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;

  // the trip count is CTA-uniform and out-of-range key groups are predicated: the 8-lane shuffles below use the
  // full-warp mask, so every lane of a warp must reach them together.  DEC_UNROLL keys per group are loaded before any
  // is consumed: with one key per iteration the kernel ran at 4.3 TB/s at batch 256 (profiles/r02_decode.md) — the
  // scores of key j gate the loads of key j + 16.
  constexpr int DEC_UNROLL = 4;
  for (int base = j_begin; base < j_end; base += DEC_GROUPS * DEC_UNROLL) {
    Raw8<T> kraw[DEC_UNROLL], vraw[DEC_UNROLL];
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; ++u) {
      const int j = base + u * DEC_GROUPS + grp;
      if (j < j_end) {
        if (j == pos) {
          kraw[u].load(knew + sub * 8);
          vraw[u].load(vnew + sub * 8);
          kraw[u].store(kc + (int64_t)pos * DD + sub * 8);      // append (fused kv-cache write)
          vraw[u].store(vc + (int64_t)pos * DD + sub * 8);
        } else {
          kraw[u].load(kc + (int64_t)j * DD + sub * 8);
          vraw[u].load(vc + (int64_t)j * DD + sub * 8);
        }
      } else {
        kraw[u].zero();
        vraw[u].zero();
      }
    }
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; ++u) {
      const int j = base + u * DEC_GROUPS + grp;
      float kf[8], vf[8];
      kraw[u].unpack(kf);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) s = fmaf(qv.v[d], kf[d], s);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (j < j_end) {
        s = s * scale - slope * (float)(pos - j);
        const float mnew = fmaxf(m, s);
        const float corr = expf(m - mnew);      // m = -inf on first key → 0
        const float p = expf(s - mnew);
        l = l * corr + p;
        vraw[u].unpack(vf);
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = acc[d] * corr + p * vf[d];
        m = mnew;
      }
    }
  }
  if (sub == 0) { sm_m[grp] = m; sm_l[grp] = l; }
#pragma unroll
  for (int d = 0; d < 8; ++d) sm_o[grp][sub * 8 + d] = acc[d];
  __syncthreads();
  if (threadIdx.x < DD) {
    const int d = threadIdx.x;
    float M = -CUDART_INF_F;
#pragma unroll
    for (int g = 0; g < DEC_GROUPS; ++g) M = fmaxf(M, sm_m[g]);
    float L = 0.f, O = 0.f;
    if (M != -CUDART_INF_F) {
#pragma unroll
      for (int g = 0; g < DEC_GROUPS; ++g) {
        const float w = (sm_m[g] == -CUDART_INF_F) ? 0.f : expf(sm_m[g] - M);
        L += sm_l[g] * w;
        O += sm_o[g][d] * w;
      }
    }
    if (splits == 1) {
      out[(int64_t)b * HD + h * DD + d] = from_f32<T>(L > 0.f ? O / L : 0.f);
    } else {
      float* pp = partial + (((int64_t)b * H + h) * splits + split) * (DD + 2);
      pp[2 + d] = O;
      if (d == 0) { pp[0] = M; pp[1] = L; }
    }
  }
  if (splits > 1 && tickets) {
    // merge in the same launch: the last CTA of this (b, h) to finish combines the split partials (a separate merge
    // kernel costs a kernel boundary per layer of a ~0.5 ms step).  tickets[b·H+h] is zero on entry and left zero.
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(tickets + (int64_t)b * H + h, 1) == splits - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (threadIdx.x < DD) {
        const int d = threadIdx.x;
        const float* pp = partial + ((int64_t)b * H + h) * splits * (DD + 2);
        float M = -CUDART_INF_F;
        for (int s2 = 0; s2 < splits; ++s2) M = fmaxf(M, __ldcg(pp + s2 * (DD + 2)));
        float L = 0.f, O = 0.f;
        for (int s2 = 0; s2 < splits; ++s2) {
          const float ms = __ldcg(pp + s2 * (DD + 2));
          const float w = (ms == -CUDART_INF_F) ? 0.f : expf(ms - M);
          L += __ldcg(pp + s2 * (DD + 2) + 1) * w;
          O += __ldcg(pp + s2 * (DD + 2) + 2 + d) * w;
        }
        out[((int64_t)b * H + h) * DD + d] = from_f32<T>(L > 0.f ? O / L : 0.f);
      }
      if (threadIdx.x == 0) tickets[(int64_t)b * H + h] = 0;
    }
  }
}

// ---- bf16 product kernel: persistent CTAs, K/V rows streamed HBM → shared memory by 1-D bulk TMA ---------------------
// The register-load kernel above stalls every warp on its own loads (load 4 keys, consume 4 keys, repeat): 4.3 TB/s at
// batch 256 (profiles/r02_decode.md).  Here a producer warp keeps a ring of STAGES x (KEYS rows of K + KEYS rows of V)
// in flight per CTA with cp.async.bulk (UBLKCP) — a (b, h) slab of the head-major cache is contiguous, so a stage is two
// plain byte ranges — and runs ahead across the work items of its CTA, so neither memory latency nor item boundaries
// stop the HBM stream.  CW consumer warps read the stage 8 lanes per key row (conflict-free 512-byte warp reads), with
// one online-softmax update (exp2 domain) per U = KEYS / (4 CW) keys per group.  Work items are (b, h, kv-split); a CTA
// takes a contiguous run of them (consecutive (b, h) slabs are adjacent in the cache).  The consumers are latency-bound
// (LDS → dot → 3 shuffles → max → ex2 → FMA per stage), so what matters is the number of consumer warps per SM and that
// nothing else is exposed per item: the next item's q is prefetched, the groups of a warp merge with shuffles, and the
// cross-warp merge buffer is double-buffered (one named barrier per item).
// Launched with the programmatic-dependent-launch attribute: the producer starts streaming cache rows (written by
// EARLIER steps) before griddepcontrol.wait; only the consumers, which read this step's q / k / v, wait for the
// projection kernel in front.
constexpr float AS_LOG2E = 1.4426950408889634f;
constexpr float AS_LN2 = 0.6931471805599453f;
constexpr int AS_MAX_STAGES = 8;
constexpr int AS_MAX_CW = 8;

struct AsShared {
  uint64_t full[AS_MAX_STAGES], empty[AS_MAX_STAGES];
  float m[2][AS_MAX_CW], l[2][AS_MAX_CW];
  __align__(16) float o[2][AS_MAX_CW][DD];
  int last;
};

__device__ __forceinline__ void as_unpack(const uint4& r, float (&v)[8]) {
  v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
  v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
  v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
  v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ float as_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int KEYS, int STAGES, int CW>
__global__ void __launch_bounds__(32 * (1 + CW), 2)
attn_decode_stream_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ k_cache,
                          __nv_bfloat16* __restrict__ v_cache, __nv_bfloat16* __restrict__ out,
                          float* __restrict__ partial, const float* __restrict__ slopes, int B, int H, int Tmax,
                          int pos_arg, const int32_t* __restrict__ pos_dev, int splits, float scale,
                          int* __restrict__ tickets, int contiguous) {
  using namespace sm100;
  constexpr int STAGE_BYTES = 2 * KEYS * DD * 2;          // K rows then V rows
  constexpr int GROUPS = CW * 4;                          // 8-lane key groups per CTA
  constexpr int U = KEYS / GROUPS;                        // keys per group per stage
  static_assert(U >= 1 && U * GROUPS == KEYS && STAGES <= AS_MAX_STAGES && CW <= AS_MAX_CW, "bad ring shape");
  extern __shared__ __align__(128) uint8_t as_raw[];
  uint8_t* ring = as_raw;
  AsShared& sh = *reinterpret_cast<AsShared*>(as_raw + STAGES * STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], CW); }
    fence_barrier_init();
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // the position was advanced by a kernel many launches back (end of the previous step): complete and visible even
  // when this grid starts early
  const int pos = pos_dev ? min(*pos_dev, Tmax - 1) : pos_arg;
  const int nkeys = pos + 1;
  const int chunk = (nkeys + splits - 1) / splits;
  const int n_items = B * H * splits;
  const int HD = H * DD;
  // items of this CTA: a CONTIGUOUS run (each CTA then reads one sequential DRAM stream per tensor) or round-robin
  const int item0 = contiguous ? (int)(((int64_t)blockIdx.x * n_items) / gridDim.x) : (int)blockIdx.x;
  const int item1 = contiguous ? (int)(((int64_t)(blockIdx.x + 1) * n_items) / gridDim.x) : n_items;
  const int istep = contiguous ? 1 : (int)gridDim.x;

  if (warp == 0) {
    // ===================== producer: cached rows [j_begin, min(j_end, pos)) of every item of this CTA → ring =============
    uint32_t cnt = 0;
    for (int item = item0; item < item1; item += istep) {
      const int split = item % splits, bh = item / splits;
      const int j_begin = split * chunk;
      const int j_stop = min(min(nkeys, j_begin + chunk), pos);
      const __nv_bfloat16* kc = k_cache + (int64_t)bh * Tmax * DD;
      const __nv_bfloat16* vc = v_cache + (int64_t)bh * Tmax * DD;
      for (int j = j_begin; j < j_stop; j += KEYS) {
        const int n = min(KEYS, j_stop - j);
        const int s = cnt % STAGES;
        const uint32_t use = cnt / STAGES;
        if (use > 0) mbar_wait(&sh.empty[s], (use - 1) & 1);
        if (elect_one()) {
          const uint32_t bytes = (uint32_t)n * DD * 2;
          uint8_t* dst = ring + s * STAGE_BYTES;
          mbar_arrive_expect_tx(&sh.full[s], 2 * bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(dst)), "l"(kc + (int64_t)j * DD), "r"(bytes), "r"(smem_u32(&sh.full[s])) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(dst + KEYS * DD * 2)), "l"(vc + (int64_t)j * DD), "r"(bytes),
                       "r"(smem_u32(&sh.full[s])) : "memory");
        }
        __syncwarp();
        ++cnt;
      }
    }
    return;
  }

  // ===================== consumers ======================================================================================
  const int ct = tid - 32;                    // 0 .. 32 CW - 1
  const int cw = ct >> 5;                     // consumer warp
  const int grp = ct >> 3;                    // key group
  const int sub = lane & 7;                   // dims sub*8 .. sub*8+7
  asm volatile("griddepcontrol.wait;" ::: "memory");      // q / k / v of this step come from the kernel in front
  uint32_t cnt = 0, nitem = 0;
  uint4 q_next = make_uint4(0, 0, 0, 0);
  float slope_next = 0.f;
  if (item0 < item1) {
    const int bh = item0 / splits;
    q_next = *reinterpret_cast<const uint4*>(qkv + (int64_t)(bh / H) * 3 * HD + (bh % H) * DD + sub * 8);
    slope_next = slopes ? __ldg(slopes + bh % H) : 0.f;
  }
  for (int item = item0; item < item1; item += istep, ++nitem) {
    const int split = item % splits, bh = item / splits;
    const int h = bh % H, b = bh / H;
    const int j_begin = split * chunk;
    const int j_end = min(nkeys, j_begin + chunk);
    const int j_stop = min(j_end, pos);
    const float slope2 = slope_next * AS_LOG2E;
    const __nv_bfloat16* qrow = qkv + (int64_t)b * 3 * HD + h * DD + sub * 8;
    float q[8];
    as_unpack(q_next, q);
    {
      const float f = scale * AS_LOG2E;
#pragma unroll
      for (int d = 0; d < 8; ++d) q[d] *= f;
    }
    if (item + istep < item1) {                            // the next item's query row: its latency hides behind this item
      const int bh2 = (item + istep) / splits;
      q_next = *reinterpret_cast<const uint4*>(qkv + (int64_t)(bh2 / H) * 3 * HD + (bh2 % H) * DD + sub * 8);
      slope_next = slopes ? __ldg(slopes + bh2 % H) : 0.f;
    }
    float m = -CUDART_INF_F, l = 0.f, acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
    if (j_end == nkeys && j_begin < nkeys && grp == GROUPS - 1) {
      // the new token (key index pos, distance 0): append to the cache and start the running softmax from it.
      // One 8-lane group (the last quarter of the last consumer warp — the group with the fewest keys in a ragged
      // stage); the shuffles stay inside that quarter warp.
      const uint4 kr = *reinterpret_cast<const uint4*>(qrow + HD);
      const uint4 vr = *reinterpret_cast<const uint4*>(qrow + 2 * HD);
      __nv_bfloat16* kc = k_cache + ((int64_t)bh * Tmax + pos) * DD + sub * 8;
      __nv_bfloat16* vc = v_cache + ((int64_t)bh * Tmax + pos) * DD + sub * 8;
      *reinterpret_cast<uint4*>(kc) = kr;
      *reinterpret_cast<uint4*>(vc) = vr;
      float kf[8];
      as_unpack(kr, kf);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) s = fmaf(q[d], kf[d], s);
      s += __shfl_xor_sync(0xff000000u, s, 4);
      s += __shfl_xor_sync(0xff000000u, s, 2);
      s += __shfl_xor_sync(0xff000000u, s, 1);
      m = s; l = 1.f;
      as_unpack(vr, acc);
    }
    for (int j0 = j_begin; j0 < j_stop; j0 += KEYS) {
      const int n = min(KEYS, j_stop - j0);
      const int s_ = cnt % STAGES;
      mbar_wait(&sh.full[s_], (cnt / STAGES) & 1);
      const uint8_t* ks = ring + s_ * STAGE_BYTES;
      const uint8_t* vs = ks + KEYS * DD * 2;
      uint4 kr[U], vr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = u * GROUPS + grp;
        if (jj < n) {
          kr[u] = *reinterpret_cast<const uint4*>(ks + jj * (DD * 2) + sub * 16);
          vr[u] = *reinterpret_cast<const uint4*>(vs + jj * (DD * 2) + sub * 16);
        } else {
          kr[u] = make_uint4(0, 0, 0, 0);
          vr[u] = make_uint4(0, 0, 0, 0);
        }
      }
      float sc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float kf[8];
        as_unpack(kr[u], kf);
        float s0 = q[0] * kf[0], s1 = q[1] * kf[1];
#pragma unroll
        for (int d = 2; d < 8; d += 2) { s0 = fmaf(q[d], kf[d], s0); s1 = fmaf(q[d + 1], kf[d + 1], s1); }
        sc[u] = s0 + s1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 4);
#pragma unroll
      for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 2);
#pragma unroll
      for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 1);
      // the stage is in registers: hand it back before the softmax arithmetic
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty[s_]);
      ++cnt;
      float mnew = m;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = u * GROUPS + grp;
        sc[u] = jj < n ? sc[u] - slope2 * (float)(pos - (j0 + jj)) : -CUDART_INF_F;
        mnew = fmaxf(mnew, sc[u]);
      }
      if (mnew > -CUDART_INF_F) {
        const float corr = as_ex2(m - mnew);        // m = -inf → 0
        float p[U], psum = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) { p[u] = as_ex2(sc[u] - mnew); psum += p[u]; }
        l = fmaf(l, corr, psum);
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] *= corr;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float vf[8];
          as_unpack(vr[u], vf);
#pragma unroll
          for (int d = 0; d < 8; ++d) acc[d] = fmaf(p[u], vf[d], acc[d]);
        }
        m = mnew;
      }
    }
    // ---- merge: the four groups of a warp with shuffles, the warps through (double-buffered) shared memory
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, off);
      const float lo = __shfl_xor_sync(0xffffffffu, l, off);
      const float mn = fmaxf(m, mo);
      const float wa = (m == -CUDART_INF_F) ? 0.f : as_ex2(m - mn);
      const float wb = (mo == -CUDART_INF_F) ? 0.f : as_ex2(mo - mn);
      l = l * wa + lo * wb;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] = acc[d] * wa + __shfl_xor_sync(0xffffffffu, acc[d], off) * wb;
      m = mn;
    }
    const int buf = nitem & 1;
    if (lane < 8) {
      if (lane == 0) { sh.m[buf][cw] = m; sh.l[buf][cw] = l; }
      *reinterpret_cast<float4*>(&sh.o[buf][cw][sub * 8]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(&sh.o[buf][cw][sub * 8 + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
    // (the buffer written two items ago is free again: every thread passed the barrier of the item in between after
    //  it had finished reading it)
    if (ct < DD) {
      const int d = ct;
      float M = -CUDART_INF_F;
#pragma unroll
      for (int g = 0; g < CW; ++g) M = fmaxf(M, sh.m[buf][g]);
      float L = 0.f, O = 0.f;
      if (M != -CUDART_INF_F) {
#pragma unroll
        for (int g = 0; g < CW; ++g) {
          const float w = as_ex2(sh.m[buf][g] - M);      // -inf → 0
          L = fmaf(sh.l[buf][g], w, L);
          O = fmaf(sh.o[buf][g][d], w, O);
        }
      }
      if (splits == 1) {
        out[(int64_t)b * HD + h * DD + d] = __float2bfloat16_rn(L > 0.f ? O / L : 0.f);
      } else {
        float* pp = partial + ((int64_t)bh * splits + split) * (DD + 2);     // (M in the natural-log domain, L, O[64])
        pp[2 + d] = O;
        if (d == 0) { pp[0] = M * AS_LN2; pp[1] = L; }
      }
    }
    if (splits > 1 && tickets) {
      // the last CTA of this (b, h) to finish combines the split partials; tickets[bh] is zero on entry and left zero.
      // bar.sync orders the 64 partial stores before thread 0's acq_rel ticket (release is cumulative); the last arriver's
      // acquire + the second bar.sync order its readers after every other split's stores (read with ld.cg)
      asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
      if (ct == 0) {
        int old;
        asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(tickets + bh) : "memory");
        sh.last = (old == splits - 1);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
      const int last = sh.last;
      if (last) {
        if (ct < DD) {
          const int d = ct;
          const float* pp = partial + (int64_t)bh * splits * (DD + 2);
          float M = -CUDART_INF_F;
          for (int s2 = 0; s2 < splits; ++s2) M = fmaxf(M, __ldcg(pp + s2 * (DD + 2)));
          float L = 0.f, O = 0.f;
          for (int s2 = 0; s2 < splits; ++s2) {
            const float ms = __ldcg(pp + s2 * (DD + 2));
            const float w = (ms == -CUDART_INF_F) ? 0.f : as_ex2((ms - M) * AS_LOG2E);
            L = fmaf(__ldcg(pp + s2 * (DD + 2) + 1), w, L);
            O = fmaf(__ldcg(pp + s2 * (DD + 2) + 2 + d), w, O);
          }
          out[(int64_t)bh * DD + d] = __float2bfloat16_rn(L > 0.f ? O / L : 0.f);
        }
        if (ct == 0) tickets[bh] = 0;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");        // sh.last is rewritten by the next item
    }
  }
}

template <typename T>
__global__ void attn_decode_merge_kernel(const float* __restrict__ partial, T* __restrict__ out, int H, int splits) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // see attn_decode_kernel
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const float* pp = partial + ((int64_t)b * H + h) * splits * (DD + 2);
  float M = -CUDART_INF_F;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, pp[s * (DD + 2)]);
  float L = 0.f, O = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float ms = pp[s * (DD + 2)];
    const float w = (ms == -CUDART_INF_F) ? 0.f : expf(ms - M);
    L += pp[s * (DD + 2) + 1] * w;
    O += pp[s * (DD + 2) + 2 + d] * w;
  }
  out[((int64_t)b * H + h) * DD + d] = from_f32<T>(L > 0.f ? O / L : 0.f);
}

template <typename T>
__global__ void kv_append_kernel(const T* __restrict__ k, const T* __restrict__ v, int64_t ld, T* __restrict__ kc,
                                 T* __restrict__ vc, int B, int H, int T_, int Tmax, int pos) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*T*H*(D/8)
  const int64_t total = (int64_t)B * T_ * H * (DD / 8);
  if (idx >= total) return;
  const int c = (int)(idx % (DD / 8));
  const int h = (int)((idx / (DD / 8)) % H);
  const int t = (int)((idx / ((int64_t)(DD / 8) * H)) % T_);
  const int b = (int)(idx / ((int64_t)(DD / 8) * H * T_));
  const int64_t src = ((int64_t)b * T_ + t) * ld + h * DD + c * 8;
  const int64_t dst = (((int64_t)b * H + h) * Tmax + pos + t) * DD + c * 8;
  Vec8<T> a;
  a.load(k + src); a.store(kc + dst);
  a.load(v + src); a.store(vc + dst);
}

}  // namespace vg

using namespace vg;

extern "C" size_t vg_attn_decode_workspace(int64_t B, int64_t H, int64_t D, int64_t splits) {
  return splits <= 1 ? 0 : (size_t)(B * H * splits * (D + 2)) * sizeof(float);
}

extern "C" int vg_attn_decode(const void* qkv, void* k_cache, void* v_cache, void* out, const float* slopes,
                              int64_t B, int64_t H, int64_t D, int64_t Tmax, int64_t pos, const int32_t* pos_dev,
                              int64_t splits, float scale, int dtype, void* workspace, size_t workspace_bytes,
                              int32_t* tickets, vg_stream_t stream) {
  VG_REQUIRE(qkv && k_cache && v_cache && out, -1, "vg_attn_decode: null pointer");
  VG_REQUIRE(valid_dtype(dtype), -2, "vg_attn_decode: bad dtype");
  VG_REQUIRE(D == DD, -3, "vg_attn_decode: head dim %lld unsupported (only 64)", (long long)D);
  VG_REQUIRE(B > 0 && H > 0 && B < 65536 && H < 65536 && pos >= 0 && pos < Tmax, -3,
             "vg_attn_decode: bad shape (pos=%lld Tmax=%lld)", (long long)pos, (long long)Tmax);
  VG_REQUIRE(splits >= 1 && splits <= 64, -3, "vg_attn_decode: splits must be in [1,64]");
  VG_REQUIRE(aligned(qkv, 16) && aligned(k_cache, 16) && aligned(v_cache, 16), -4, "vg_attn_decode: unaligned");
  if (splits > 1)
    VG_REQUIRE(workspace && workspace_bytes >= vg_attn_decode_workspace(B, H, D, splits), -5,
               "vg_attn_decode: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  static const int use_stream = getenv("VG_ATTN_DECODE_STREAM") ? atoi(getenv("VG_ATTN_DECODE_STREAM")) : 1;
  if (dtype == VG_BF16 && use_stream && (splits == 1 || tickets)) {
    // VG_AD_CFG = keys per stage * 1000 + stages * 100 + consumer warps * 10 + CTAs per SM: the three ring shapes kept from
    // the sweep of profiles/r02_decode.md (64343 = product; 64542 and 128382 are the runners-up, exercised by the tests)
    static const int cfgv = getenv("VG_AD_CFG") ? atoi(getenv("VG_AD_CFG")) : 64343;   // measured: profiles/r02_decode.md
    static const int contiguous = getenv("VG_AD_CONTIG") ? atoi(getenv("VG_AD_CONTIG")) : 1;
    typedef void (*Kern)(const __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, float*, const float*, int, int,
                         int, int, const int32_t*, int, float, int*, int);
    Kern kern = nullptr;
    const int keys = cfgv / 1000, stages = (cfgv / 100) % 10, cws = (cfgv / 10) % 10, per_sm = cfgv % 10;
    if (keys == 64 && stages == 3 && cws == 4) kern = attn_decode_stream_kernel<64, 3, 4>;          // product shape
    else if (keys == 64 && stages == 5 && cws == 4) kern = attn_decode_stream_kernel<64, 5, 4>;     // deeper ring, 2 CTAs / SM
    else if (keys == 128 && stages == 3 && cws == 8) kern = attn_decode_stream_kernel<128, 3, 8>;   // 8 consumer warps, 2 CTAs / SM
    VG_REQUIRE(kern != nullptr && per_sm >= 1 && per_sm <= 4, -3, "vg_attn_decode: unknown VG_AD_CFG %d", cfgv);
    static int sms = 0;
    const size_t smem = (size_t)stages * keys * DD * 4 + sizeof(AsShared);
    if (!sms) {
      int dev = 0;
      VG_CUDA(cudaGetDevice(&dev));
      VG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int64_t n_items = B * H * splits;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_items < per_sm * sms ? n_items : per_sm * sms));
    cfg.blockDim = dim3(32 * (1 + cws));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VG_CUDA(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache,
                               (__nv_bfloat16*)out, (float*)workspace, slopes, (int)B, (int)H, (int)Tmax, (int)pos, pos_dev,
                               (int)splits, scale, (int*)tickets, contiguous));
    return 0;
  }
  dim3 grid((unsigned)splits, (unsigned)H, (unsigned)B);
  if (dtype == VG_F32) {
    attn_decode_kernel<float><<<grid, DEC_WARPS * 32, 0, st>>>((const float*)qkv, (float*)k_cache, (float*)v_cache,
                                                              (float*)out, (float*)workspace, slopes, (int)H,
                                                              (int)Tmax, (int)pos, pos_dev, (int)splits, scale, tickets);
  } else {
    attn_decode_kernel<__nv_bfloat16><<<grid, DEC_WARPS * 32, 0, st>>>(
        (const __nv_bfloat16*)qkv, (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache, (__nv_bfloat16*)out,
        (float*)workspace, slopes, (int)H, (int)Tmax, (int)pos, pos_dev, (int)splits, scale, tickets);
  }
  VG_LAUNCH_CHECK("vg_attn_decode");
  if (splits > 1 && !tickets) {
    dim3 g2((unsigned)H, (unsigned)B);
    if (dtype == VG_F32)
      attn_decode_merge_kernel<float><<<g2, DD, 0, st>>>((const float*)workspace, (float*)out, (int)H, (int)splits);
    else
      attn_decode_merge_kernel<__nv_bfloat16><<<g2, DD, 0, st>>>((const float*)workspace, (__nv_bfloat16*)out, (int)H,
                                                                 (int)splits);
    VG_LAUNCH_CHECK("vg_attn_decode(merge)");
  }
  return 0;
}

extern "C" int vg_kv_append(const void* k, const void* v, int64_t ld, void* k_cache, void* v_cache, int64_t B,
                            int64_t H, int64_t D, int64_t T, int64_t Tmax, int64_t pos, int dtype,
                            vg_stream_t stream) {
  VG_REQUIRE(k && v && k_cache && v_cache, -1, "vg_kv_append: null pointer");
  VG_REQUIRE(valid_dtype(dtype), -2, "vg_kv_append: bad dtype");
  VG_REQUIRE(D == DD, -3, "vg_kv_append: head dim %lld unsupported (only 64)", (long long)D);
  VG_REQUIRE(B > 0 && H > 0 && T > 0 && pos >= 0 && pos + T <= Tmax, -3, "vg_kv_append: range exceeds cache");
  VG_REQUIRE(ld % 8 == 0 && aligned(k, 16) && aligned(v, 16) && aligned(k_cache, 16) && aligned(v_cache, 16), -4,
             "vg_kv_append: unaligned");
  const int64_t total = B * T * H * (DD / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VG_F32)
    kv_append_kernel<float><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>((const float*)k, (const float*)v, ld,
                                                                           (float*)k_cache, (float*)v_cache, (int)B,
                                                                           (int)H, (int)T, (int)Tmax, (int)pos);
  else
    kv_append_kernel<__nv_bfloat16><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
        (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, ld, (__nv_bfloat16*)k_cache, (__nv_bfloat16*)v_cache,
        (int)B, (int)H, (int)T, (int)Tmax, (int)pos);
  VG_LAUNCH_CHECK("vg_kv_append");
  return 0;
}

namespace vg {
__global__ void add_i32_kernel(int32_t* c, int32_t d) { *c += d; }
}  // namespace vg

extern "C" int vg_add_i32(int32_t* counter, int32_t delta, vg_stream_t stream) {
  VG_REQUIRE(counter, -1, "vg_add_i32: null pointer");
  vg::add_i32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  VG_LAUNCH_CHECK("vg_add_i32");
  return 0;
}
