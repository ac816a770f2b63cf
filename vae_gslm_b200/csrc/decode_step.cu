// decode_step.cu — ONE persistent kernel for the transformer part of a cached generation step (LVTR.step,
// models/speech/lvtr.py:253-279 → modules/transformer/layers.py:134-195 with past_kv, modules/attention/attention.py:52-85,
// modules/norm.py:28-32): stack-input linear, 16 x [RMSNorm1 + QKV, cached attention (+ KV append), out-proj + residual,
// RMSNorm3 + FFN1 + GELU, FFN2 + residual], final RMSNorm, q_spliter | token_spliter (+ReLU), prior/FiLM head and logit head.
//
// Why one kernel: the step is HBM-bound (408.7 MB of bf16 weights + the KV cache) but the 92-kernel engine of
// round 1 (decode.py) was bound by kernel boundaries (~5 us each, 0.47 ms at batch 1 against 63 us of weight streaming).
// Here one CTA per SM stays resident for the whole step:
//   * WEIGHTS never wait for activations: the host packs, per CTA, the exact byte stream of weight slabs that CTA will
//     consume (in order, already in the tcgen05 SWIZZLE_128B K-major shared-memory layout), and a producer warp copies it
//     with 1-D bulk TMA (cp.async.bulk → UBLKCP) through an 8 x 16 KB ring, running ahead across phase boundaries — the HBM
//     stream does not stop at a barrier;
//   * the step is a list of PHASES separated by device-wide barriers; in a GEMM phase every CTA executes at most one UNIT
//     (R <= 128 output features x a k-range of one linear layer) described by a host-built task table — the decomposition
//     per batch size (small batch: few features x full K per CTA; large batch: 128 features x a k-slice) is host policy,
//     the kernel is an interpreter;
//   * a unit is a swap-AB tcgen05 GEMM: D[128 features x Bp batch columns] (TMEM, fp32) += W_slab[128 x 64] · Xᵀ[64 x Bp]
//     per 64-wide k-block, weights as the M operand so that batch 1..256 is one code path (N = 16..256);
//   * the X operand is formed by the CONSUMER: worker warps read the producer phase's fp32 accumulators (L2), apply what the
//     reference applies between the two linears (RMSNorm scale / 1/rms, bias + GELU / ReLU), round to bf16 and store the
//     swizzled K-major tile — so split-K partial sums (red.global.add.f32 from the epilogue) need no finishing pass;
//   * the residual stream is an fp32 [B, d] buffer that out-proj and FFN2 units reduce into directly;
//   * attention phases walk (sequence, head, kv-split) items over the head-major cache, 8 lanes per key row, online softmax
//     per 8-lane group, in-kernel merge of split partials by the last arriver.
// Every spin (mbarrier or global) is bounded: on time-out the kernel records where it was and traps instead of hanging.
#include <math_constants.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace vg {
using namespace sm100;

constexpr int DS_THREADS = 256;            // warp 0: weight producer, warp 1: MMA issuer, warps 2..7: workers (4..7: epilogue)
constexpr int DS_WORKERS = 192;
constexpr int DS_STAGE = 32768;            // weight ring stage: one k-block slab of up to 256 features (or several smaller ones)
constexpr int DS_NSTAGES = 4;
constexpr int DS_XSLOT = 32768;            // X operand slot (two of them)
constexpr int DS_HD = 64;                  // head dim
constexpr int DS_AW = 7;                   // warps 1..7 take attention items (the MMA warp has nothing else to do there)
constexpr long long DS_TIMEOUT = 4000000000LL;      // ~2 s of SM clocks

typedef vg_decode_step_task DsTask;
typedef vg_decode_step_args DsArgs;

constexpr int DS_VEC_MAX = 1024;           // floats of a unit's per-k vector (RMSNorm scale / bias) prefetched into shared memory
constexpr int DS_BIAS_MAX = 256;           // floats of a unit's output bias (R <= 256)

struct DsShared {
  uint64_t full[DS_NSTAGES], empty[DS_NSTAGES], xfull[2], xempty[2], accfull;
  uint32_t tmem_base;
  float am[DS_AW], al[DS_AW];                // cooperative attention: per-warp partial softmax state of one item
  __align__(16) float ao[DS_AW][DS_HD];
  float ss[256];
  int phase_kind[128];
  __align__(16) float vec[2][DS_VEC_MAX];
  __align__(16) float bias[2][DS_BIAS_MAX];
};

__device__ __forceinline__ void ds_die(const DsArgs& a, int code, int phase) {
  if (a.debug) {
    a.debug[0] = 0xDEAD0000ull | (unsigned)code;
    a.debug[1] = ((unsigned long long)blockIdx.x << 32) | (unsigned)phase;
    a.debug[2] = threadIdx.x;
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void ds_mbar_wait(uint64_t* bar, uint32_t parity, const DsArgs& a, int code, int phase) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > DS_TIMEOUT) ds_die(a, code, phase);
}
__device__ __forceinline__ void ds_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned ds_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ds_ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ds_bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ds_red_add(float* p, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void ds_red_add4(float* p, float x, float y, float z, float w) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void ds_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ds_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ds_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t ds_pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

#define DS_TRACE(slot, thread) \
  do { if (a.trace && threadIdx.x == (thread)) a.trace[((size_t)blockIdx.x * a.NP + p) * 8 + (slot)] = clock64(); } while (0)

// ---- device-wide barrier (warps 1..7; the producer warp never stops).
//   mode 1: one monotonic counter — red.release.gpu by one thread per CTA, the same thread polls with ld.acquire.gpu
//   mode 2: per-CTA flag words gathered by CTA 0, which publishes one "go" word that the other CTAs poll
// (every CTA polling every flag, 148 x 148 loads per round on one L2 slice, was 2x slower than either: profiles/r02_decode.md)
__device__ __forceinline__ void ds_grid_barrier(const DsArgs& a, unsigned epoch, int phase) {
  const int p = phase;
  ds_named_bar(1, DS_THREADS - 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DS_TRACE(0, 32);
  if (warp == 1) {
    const unsigned G = gridDim.x;
    if (a.barrier_mode == 1) {
      if (lane == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(a.bar_flags), "r"(1u) : "memory");
        const unsigned target = epoch * G;
        const long long t0 = clock64();
        while ((int)(ds_ld_acquire(a.bar_flags) - target) < 0)
          if (clock64() - t0 > DS_TIMEOUT) ds_die(a, 101, phase);
      }
    } else {
      unsigned* go = a.bar_flags + ((G + 63u) & ~31u);
      if (lane == 0)
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.bar_flags + blockIdx.x), "r"(epoch) : "memory");
      const long long t0 = clock64();
      if (blockIdx.x == 0) {
        for (;;) {
          bool ok = true;
          for (unsigned i = lane; i < G; i += 32) ok = ok & ((int)(ds_ld_acquire(a.bar_flags + i) - epoch) >= 0);
          if (__all_sync(0xffffffffu, ok)) break;
          if (clock64() - t0 > DS_TIMEOUT) ds_die(a, 102, phase);
        }
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(epoch) : "memory");
      } else if (lane == 0) {
        while ((int)(ds_ld_acquire(go) - epoch) < 0)
          if (clock64() - t0 > DS_TIMEOUT) ds_die(a, 103, phase);
      }
    }
  }
  DS_TRACE(1, 32);
  ds_named_bar(1, DS_THREADS - 32);
}

// ---- per-k vector (RMSNorm scale / bias of the X transform) and output bias of a unit → shared memory, ahead of time:
// these 33 + 35 vectors are touched once per step, and the 408 MB weight stream evicts them from L2 in between, so a plain
// load in the transform is an HBM miss on the critical path of every phase.  Issued by warps 2..3 with cp.async right after
// the previous unit's transform; waited for at the start of the unit's own transform.
__device__ __forceinline__ bool ds_vec_prefetched(const DsTask& t) { return t.vec != nullptr && t.nkb * 64 <= DS_VEC_MAX; }
__device__ __forceinline__ void ds_prefetch_vec(const DsTask& t, DsShared& sh, int par) {
  const int at = threadIdx.x - 64;              // 0..63
  if (ds_vec_prefetched(t))
    for (int i = at; i < t.nkb * 16; i += 64) ds_cp_async16(&sh.vec[par][i * 4], t.vec + t.k0 + i * 4);
  if (t.bias_out)
    for (int i = at; i < t.R / 4; i += 64) ds_cp_async16(&sh.bias[par][i * 4], t.bias_out + t.n0 + i * 4);
  ds_cp_async_commit();
}

// ---- X operand: worker warps form the bf16, SWIZZLE_128B K-major [Bp x 64] tiles of this unit's k-range.
//   kind 0: x bf16 [B, ldx] as is          kind 1: x f32 * vec[k]   (RMSNorm scale; 1/rms is applied by the consumer of the sums)
//   kind 2: act(x f32 * rstd[b] + vec[k])  (rstd from the row sum-of-squares the kind-1 phase accumulated)
// Items of 8 consecutive k of one batch row; a thread loads DS_XU items before it touches any of them (the sources are
// L2-resident accumulators: one dependent L2 round trip per item would be the whole cost of the transform).
constexpr int DS_XU = 4;
__device__ __forceinline__ float ds_act(float y, int act) {
  return act == VG_ACT_GELU ? gelu_fast(y) : (act == VG_ACT_RELU ? fmaxf(y, 0.f) : y);
}
__device__ __forceinline__ void ds_transform(const DsArgs& a, const DsTask& t, DsShared& sh, uint8_t* xs, uint32_t& xcount,
                                             int par, int phase) {
  const int p = phase;
  const int wt = threadIdx.x - 64;
  const int lane = threadIdx.x & 31;
  const int B = a.B, Bp = a.Bp;
  const int cap = DS_XSLOT / (Bp * 128);
  const int xkb = t.nkb < cap ? t.nkb : cap;                   // a power of two (host contract)
  const int nchunks = t.nkb / xkb;
  const int lg = 31 - __clz(xkb * 8);                          // items (8 consecutive k of one row) per batch row = 1 << lg
  const int per_b = 1 << lg;
  const int items = B << lg;
  const int items_pad = (items + 31) & ~31;
  const bool want_ss = t.ss_out != nullptr;
  const bool vec_smem = ds_vec_prefetched(t);
  const int kind = t.x_kind;
  if (wt < 64) ds_cp_async_wait_all();          // the prefetch of this unit's vectors (issued by these two warps)
  ds_named_bar(2, DS_WORKERS);
  if (kind == 3) {
    // x = the attention output, formed from the kv-split partials of the attention phase (late merge): for 8 consecutive
    // dims of one (sequence, head): o = sum_s 2^(m_s - M) o_s / sum_s 2^(m_s - M) l_s, rounded to bf16
    const int nsplit = a.nsplit, Hh = a.H;
    for (int c = 0; c < nchunks; ++c) {
      const int slot = xcount & 1;
      const uint32_t use = xcount >> 1;
      if (use > 0) ds_mbar_wait(&sh.xempty[slot], (use - 1) & 1, a, 10, phase);
      uint8_t* dst = xs + slot * DS_XSLOT;
      const int kc0 = c * xkb * 64;
      for (int item = wt; item < items; item += DS_WORKERS) {
        const int b = item >> lg;
        const int rem = item & (per_b - 1);
        const int j = rem >> 3, g = rem & 7;
        const int k = t.k0 + kc0 + (rem << 3);               // first of the 8 model dims of this item
        const int h = k >> 6, d0 = k & 63;
        const float* pp = reinterpret_cast<const float*>(t.x) + ((size_t)(b * Hh + h) * nsplit) * (DS_HD + 8);
        float M = -CUDART_INF_F;
        for (int s2 = 0; s2 < nsplit; ++s2) M = fmaxf(M, __ldcg(pp + s2 * (DS_HD + 8)));
        float L = 0.f, o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        for (int s2 = 0; s2 < nsplit; ++s2) {
          const float2 ml = __ldcg(reinterpret_cast<const float2*>(pp + s2 * (DS_HD + 8)));
          const float4 v0 = __ldcg(reinterpret_cast<const float4*>(pp + s2 * (DS_HD + 8) + 8 + d0));
          const float4 v1 = __ldcg(reinterpret_cast<const float4*>(pp + s2 * (DS_HD + 8) + 8 + d0) + 1);
          const float w = (ml.x == -CUDART_INF_F) ? 0.f : ex2_approx(ml.x - M);
          L = fmaf(ml.y, w, L);
          o[0] = fmaf(v0.x, w, o[0]); o[1] = fmaf(v0.y, w, o[1]); o[2] = fmaf(v0.z, w, o[2]); o[3] = fmaf(v0.w, w, o[3]);
          o[4] = fmaf(v1.x, w, o[4]); o[5] = fmaf(v1.y, w, o[5]); o[6] = fmaf(v1.z, w, o[6]); o[7] = fmaf(v1.w, w, o[7]);
        }
        const float inv = L > 0.f ? 1.0f / L : 0.f;
        uint4 packed;
        packed.x = ds_pack_bf16(o[0] * inv, o[1] * inv); packed.y = ds_pack_bf16(o[2] * inv, o[3] * inv);
        packed.z = ds_pack_bf16(o[4] * inv, o[5] * inv); packed.w = ds_pack_bf16(o[6] * inv, o[7] * inv);
        for (int r = 0; r < a.rep; ++r) {
          const int br = b + r * (128 / a.rep);
          *reinterpret_cast<uint4*>(dst + j * (Bp * 128) + (br >> 3) * 1024 + (br & 7) * 128 + ((g ^ (br & 7)) << 4)) = packed;
        }
      }
      fence_proxy_async();
      mbar_arrive(&sh.xfull[slot]);
      if (c == 0) DS_TRACE(7, 64);
      ++xcount;
    }
    return;
  }
  for (int c = 0; c < nchunks; ++c) {
    const int slot = xcount & 1;
    const uint32_t use = xcount >> 1;
    if (use > 0) ds_mbar_wait(&sh.xempty[slot], (use - 1) & 1, a, 10, phase);
    uint8_t* dst = xs + slot * DS_XSLOT;
    const int kc0 = c * xkb * 64;                               // first k of this chunk relative to the unit's k0
    for (int base = wt; base < items_pad; base += DS_WORKERS * DS_XU) {
      uint4 raw0[DS_XU], raw1[DS_XU];
      float ssv[DS_XU];
#pragma unroll
      for (int u = 0; u < DS_XU; ++u) {
        const int item = base + u * DS_WORKERS;
        ssv[u] = 0.f;
        if (item < items) {
          const int b = item >> lg;
          const int kl = kc0 + ((item & (per_b - 1)) << 3);
          if (kind == 0) {
            raw0[u] = __ldcg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(t.x) + (size_t)b * t.ldx +
                                                            t.k0 + kl));
          } else {
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(t.x) + (size_t)b * t.ldx + t.k0 + kl);
            raw0[u] = __ldcg(src);
            raw1[u] = __ldcg(src + 1);
            if (kind == 2 && t.ss_in) ssv[u] = __ldcg(t.ss_in + b);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < DS_XU; ++u) {
        const int item = base + u * DS_WORKERS;
        float ssq = 0.f;
        if (item < items) {
          const int b = item >> lg;
          const int rem = item & (per_b - 1);
          const int j = rem >> 3, g = rem & 7;
          uint4 packed = raw0[u];
          if (kind != 0) {
            const int kl = kc0 + (rem << 3);
            const float4* vp = vec_smem ? reinterpret_cast<const float4*>(&sh.vec[par][kl])
                                        : reinterpret_cast<const float4*>(t.vec + t.k0 + kl);
            const float4 s0 = vp[0], s1 = vp[1];
            float v[8] = {__uint_as_float(raw0[u].x), __uint_as_float(raw0[u].y), __uint_as_float(raw0[u].z),
                          __uint_as_float(raw0[u].w), __uint_as_float(raw1[u].x), __uint_as_float(raw1[u].y),
                          __uint_as_float(raw1[u].z), __uint_as_float(raw1[u].w)};
            const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            if (kind == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { ssq = fmaf(v[i], v[i], ssq); v[i] *= s[i]; }
            } else {
              const float rstd = t.ss_in ? rsqrtf(ssv[u] * t.inv_k + t.eps) : 1.0f;
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = ds_act(fmaf(v[i], rstd, s[i]), t.act);
            }
            packed.x = ds_pack_bf16(v[0], v[1]); packed.y = ds_pack_bf16(v[2], v[3]);
            packed.z = ds_pack_bf16(v[4], v[5]); packed.w = ds_pack_bf16(v[6], v[7]);
          }
          // a.rep copies of the row, 128 / a.rep rows apart: every TMEM lane quadrant then holds the batch (see the epilogue)
          for (int r = 0; r < a.rep; ++r) {
            const int br = b + r * (128 / a.rep);
            *reinterpret_cast<uint4*>(dst + j * (Bp * 128) + (br >> 3) * 1024 + (br & 7) * 128 + ((g ^ (br & 7)) << 4)) = packed;
          }
        }
        ssv[u] = ssq;
      }
      if (want_ss) {
        // row sum-of-squares: lanes of one batch row are contiguous (per_b, a power of two >= 8, lanes per row when
        // per_b < 32, else whole warps).  items_pad is a multiple of 32, so a warp is either wholly inside a round or out.
        const int seg = per_b < 32 ? per_b : 32;
#pragma unroll
        for (int u = 0; u < DS_XU; ++u) {
          const int item = base + u * DS_WORKERS;
          if (item >= items_pad) break;
          float ssq = ssv[u];
          for (int o = 1; o < seg; o <<= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
          if ((lane & (seg - 1)) == 0 && item < items) atomicAdd(&sh.ss[item >> lg], ssq);
        }
      }
    }
    fence_proxy_async();
    mbar_arrive(&sh.xfull[slot]);
    if (c == 0) DS_TRACE(7, 64);
    ++xcount;
  }
  if (want_ss) {
    ds_named_bar(2, DS_WORKERS);
    for (int b = wt; b < B; b += DS_WORKERS) {
      ds_red_add(t.ss_out + b, sh.ss[b]);
      sh.ss[b] = 0.f;
    }
  }
}

// ---- auxiliary jobs of a phase (warps 2..3, off the critical path): clear accumulators for a later phase, write H
__device__ __forceinline__ void ds_aux(const DsTask& t) {
  const int at = threadIdx.x - 64;              // 0..63
  if (t.aux_kind == 1) {                        // zero aux_i1 float4s starting at float4 index aux_i0 of aux_ptr0
    float4* p = reinterpret_cast<float4*>(t.aux_ptr0) + t.aux_i0;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = at; i < t.aux_i1; i += 64) p[i] = z;
  } else if (t.aux_kind == 2) {
    // transformer_latent rows: H[b,k] = bf16(h[b,k] * scale[k] * rstd[b]); groups of 8 elements aux_i0 .. aux_i0+aux_i1
    const float* h = reinterpret_cast<const float*>(t.aux_ptr0);
    const float* scale = reinterpret_cast<const float*>(t.aux_ptr1);
    const float* ss = reinterpret_cast<const float*>(t.aux_ptr2);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(t.aux_ptr3);
    const int d = t.aux_i2;                     // model dim
    for (int i = at; i < t.aux_i1; i += 64) {
      const int e = (t.aux_i0 + i) * 8;
      const int b = e / d, k = e - b * d;
      const float rstd = rsqrtf(__ldcg(ss + b) * t.inv_k + t.eps);
      const float4 v0 = __ldcg(reinterpret_cast<const float4*>(h + e));
      const float4 v1 = __ldcg(reinterpret_cast<const float4*>(h + e) + 1);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + k));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + k) + 1);
      uint4 q;
      q.x = ds_pack_bf16(v0.x * s0.x * rstd, v0.y * s0.y * rstd);
      q.y = ds_pack_bf16(v0.z * s0.z * rstd, v0.w * s0.w * rstd);
      q.z = ds_pack_bf16(v1.x * s1.x * rstd, v1.y * s1.y * rstd);
      q.w = ds_pack_bf16(v1.z * s1.z * rstd, v1.w * s1.w * rstd);
      *reinterpret_cast<uint4*>(out + e) = q;
    }
  }
}

// ---- attention phase.  Items are (sequence, head, kv-split).  8 lanes share a key row (16-byte loads: a warp instruction
// covers 4 full 128-byte rows); each warp keeps two sets of DS_KU x 4 keys of K and V in flight (the loads of the next
// set are issued before the current one is consumed), online softmax per 8-lane group in the exp2 domain, the four groups
// of a warp merged with shuffles.  Two ways to deal the items (host policy, a.attn_coop):
//   0: one WARP per item — many (b, h): no CTA-level synchronisation inside the phase at all;
//   1: one CTA per item, its seven attention warps take contiguous parts of the key range and merge through shared memory
//      — few (b, h): short per-warp key ranges without a global ticket round trip.
// With nsplit > 1 the last arriver of a (b, h) (ticket) merges the partials.  The new token's k / v come from the QKV sums
// of the phase before (· 1/rms, rounded to bf16 — the value that goes into the cache) and are appended here.
constexpr int DS_KU = 4;
struct DsAttnState { float m, l, acc[8]; };

__device__ __forceinline__ void ds_attn_load(uint4 (&kraw)[DS_KU], uint4 (&vraw)[DS_KU], int base, int j_end, int pos, int grp,
                                             int sub, __nv_bfloat16* kc, __nv_bfloat16* vc, const float* row, int HD, float rstd) {
#pragma unroll
  for (int u = 0; u < DS_KU; ++u) {
    const int j = base + u * 4 + grp;
    if (j < j_end && j != pos) {
      kraw[u] = __ldcs(reinterpret_cast<const uint4*>(kc + (size_t)j * DS_HD + sub * 8));
      vraw[u] = __ldcs(reinterpret_cast<const uint4*>(vc + (size_t)j * DS_HD + sub * 8));
    } else if (j == pos && j < j_end) {
      const float4 k0 = __ldcg(reinterpret_cast<const float4*>(row + HD));
      const float4 k1 = __ldcg(reinterpret_cast<const float4*>(row + HD) + 1);
      const float4 v0 = __ldcg(reinterpret_cast<const float4*>(row + 2 * HD));
      const float4 v1 = __ldcg(reinterpret_cast<const float4*>(row + 2 * HD) + 1);
      kraw[u].x = ds_pack_bf16(k0.x * rstd, k0.y * rstd); kraw[u].y = ds_pack_bf16(k0.z * rstd, k0.w * rstd);
      kraw[u].z = ds_pack_bf16(k1.x * rstd, k1.y * rstd); kraw[u].w = ds_pack_bf16(k1.z * rstd, k1.w * rstd);
      vraw[u].x = ds_pack_bf16(v0.x * rstd, v0.y * rstd); vraw[u].y = ds_pack_bf16(v0.z * rstd, v0.w * rstd);
      vraw[u].z = ds_pack_bf16(v1.x * rstd, v1.y * rstd); vraw[u].w = ds_pack_bf16(v1.z * rstd, v1.w * rstd);
      *reinterpret_cast<uint4*>(kc + (size_t)pos * DS_HD + sub * 8) = kraw[u];
      *reinterpret_cast<uint4*>(vc + (size_t)pos * DS_HD + sub * 8) = vraw[u];
    } else {
      kraw[u] = make_uint4(0, 0, 0, 0);
      vraw[u] = make_uint4(0, 0, 0, 0);
    }
  }
}
__device__ __forceinline__ void ds_attn_consume(DsAttnState& st, const uint4 (&kraw)[DS_KU], const uint4 (&vraw)[DS_KU], int base,
                                                int j_end, int pos, int grp, const float (&q)[8], float slope) {
#pragma unroll
  for (int u = 0; u < DS_KU; ++u) {
    const int j = base + u * 4 + grp;
    const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kraw[u]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(k2[i]);
      s = fmaf(q[2 * i], f.x, s);
      s = fmaf(q[2 * i + 1], f.y, s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (j < j_end) {
      s -= slope * (float)(pos - j);
      const float mnew = fmaxf(st.m, s);
      const float corr = ex2_approx(st.m - mnew);           // m = -inf on the first key → 0
      const float pr = ex2_approx(s - mnew);
      st.l = st.l * corr + pr;
      const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vraw[u]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(v2[i]);
        st.acc[2 * i] = fmaf(pr, f.x, st.acc[2 * i] * corr);
        st.acc[2 * i + 1] = fmaf(pr, f.y, st.acc[2 * i + 1] * corr);
      }
      st.m = mnew;
    }
  }
}
// keys [j0, j1) of one (b, h) by one warp; on return lanes 0..7 (grp 0) hold the warp's (m, l, acc) for dims sub*8..sub*8+7
__device__ __forceinline__ void ds_attn_range(DsAttnState& st, int j0, int j1, int pos, int grp, int sub, __nv_bfloat16* kc,
                                              __nv_bfloat16* vc, const float* row, int HD, float rstd, const float (&q)[8],
                                              float slope) {
  st.m = -CUDART_INF_F; st.l = 0.f;
#pragma unroll
  for (int d = 0; d < 8; ++d) st.acc[d] = 0.f;
  uint4 ka[DS_KU], va[DS_KU], kb[DS_KU], vb[DS_KU];
  if (j0 < j1) ds_attn_load(ka, va, j0, j1, pos, grp, sub, kc, vc, row, HD, rstd);
  for (int base = j0; base < j1; base += 8 * DS_KU) {
    const int b1 = base + 4 * DS_KU, b2 = base + 8 * DS_KU;
    if (b1 < j1) ds_attn_load(kb, vb, b1, j1, pos, grp, sub, kc, vc, row, HD, rstd);
    ds_attn_consume(st, ka, va, base, j1, pos, grp, q, slope);
    if (b2 < j1) ds_attn_load(ka, va, b2, j1, pos, grp, sub, kc, vc, row, HD, rstd);
    if (b1 < j1) ds_attn_consume(st, kb, vb, b1, j1, pos, grp, q, slope);
  }
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, st.m, off);
    const float lo = __shfl_xor_sync(0xffffffffu, st.l, off);
    const float mn = fmaxf(st.m, mo);
    const float wa = (st.m == -CUDART_INF_F) ? 0.f : ex2_approx(st.m - mn);
    const float wb = (mo == -CUDART_INF_F) ? 0.f : ex2_approx(mo - mn);
    st.l = st.l * wa + lo * wb;
#pragma unroll
    for (int d = 0; d < 8; ++d) st.acc[d] = st.acc[d] * wa + __shfl_xor_sync(0xffffffffu, st.acc[d], off) * wb;
    st.m = mn;
  }
}
// one warp publishes (m, l, o[64]) of (b, h, split) — lanes 0..7 hold it — and, if it is the last split to arrive, merges
__device__ __forceinline__ void ds_attn_finish(const DsArgs& a, const DsAttnState& st, int bh, int split, int b, int h, int lane) {
  const int HD = a.H * DS_HD, nsplit = a.nsplit;
  const int grp = lane >> 3, sub = lane & 7;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.attn_out);
  if (a.late_merge) {
    // the split partials (m, l, o[64]) are all this phase publishes: the out-projection's X transform merges and
    // normalises them when it forms its operand (x_kind 3) — no ticket, no last-arriver round trips on the critical path
    float* pp = a.attn_partial + ((size_t)bh * nsplit + split) * (DS_HD + 8);
    if (grp == 0) {
      __stcg(reinterpret_cast<float4*>(pp + 8 + sub * 8), make_float4(st.acc[0], st.acc[1], st.acc[2], st.acc[3]));
      __stcg(reinterpret_cast<float4*>(pp + 8 + sub * 8) + 1, make_float4(st.acc[4], st.acc[5], st.acc[6], st.acc[7]));
      if (sub == 0) __stcg(reinterpret_cast<float2*>(pp), make_float2(st.m, st.l));
    }
    return;
  }
  if (nsplit == 1) {
    if (grp == 0) {
      const float inv = st.l > 0.f ? 1.0f / st.l : 0.f;
      uint4 o;
      o.x = ds_pack_bf16(st.acc[0] * inv, st.acc[1] * inv); o.y = ds_pack_bf16(st.acc[2] * inv, st.acc[3] * inv);
      o.z = ds_pack_bf16(st.acc[4] * inv, st.acc[5] * inv); o.w = ds_pack_bf16(st.acc[6] * inv, st.acc[7] * inv);
      *reinterpret_cast<uint4*>(out + (size_t)b * HD + h * DS_HD + sub * 8) = o;
    }
    return;
  }
  float* pp = a.attn_partial + ((size_t)bh * nsplit + split) * (DS_HD + 8);      // [m, l, pad.., o[64]], 16-byte aligned rows
  if (grp == 0) {
    __stcg(reinterpret_cast<float4*>(pp + 8 + sub * 8), make_float4(st.acc[0], st.acc[1], st.acc[2], st.acc[3]));
    __stcg(reinterpret_cast<float4*>(pp + 8 + sub * 8) + 1, make_float4(st.acc[4], st.acc[5], st.acc[6], st.acc[7]));
    if (sub == 0) __stcg(reinterpret_cast<float2*>(pp), make_float2(st.m, st.l));
  }
  // ticket: zero on entry and left zero.  __syncwarp orders the eight writers before lane 0's release.
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(a.tickets + bh) : "memory");
    last = (old == nsplit - 1);
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  const float* p0 = a.attn_partial + (size_t)bh * nsplit * (DS_HD + 8);
  float M2 = -CUDART_INF_F;
  for (int s2 = 0; s2 < nsplit; ++s2) M2 = fmaxf(M2, __ldcg(p0 + s2 * (DS_HD + 8)));
  float L2 = 0.f, o0 = 0.f, o1 = 0.f;
  for (int s2 = 0; s2 < nsplit; ++s2) {
    const float2 ml = __ldcg(reinterpret_cast<const float2*>(p0 + s2 * (DS_HD + 8)));
    const float w = (ml.x == -CUDART_INF_F) ? 0.f : ex2_approx(ml.x - M2);
    const float2 ov = __ldcg(reinterpret_cast<const float2*>(p0 + s2 * (DS_HD + 8) + 8 + lane * 2));
    L2 = fmaf(ml.y, w, L2);
    o0 = fmaf(ov.x, w, o0);
    o1 = fmaf(ov.y, w, o1);
  }
  const float inv = L2 > 0.f ? 1.0f / L2 : 0.f;
  *reinterpret_cast<uint32_t*>(out + (size_t)b * HD + h * DS_HD + lane * 2) = ds_pack_bf16(o0 * inv, o1 * inv);
  if (lane == 0) a.tickets[bh] = 0;
}

__device__ __forceinline__ void ds_attention(const DsArgs& a, DsShared& sh, int layer, int pos, int phase) {
  const int aw = (threadIdx.x >> 5) - 1;        // 0..6 (warps 1..7)
  const int lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int H = a.H, HD = a.H * DS_HD;
  const int nsplit = a.nsplit;
  const int n_items = a.B * H * nsplit;
  const int nkeys = pos + 1;
  const int chunk = (nkeys + nsplit - 1) / nsplit;
  const float* ss = a.ss_base + (size_t)(2 * layer) * a.B;
  constexpr float LOG2E = 1.4426950408889634f;
  const int first = a.attn_coop ? (int)blockIdx.x : (int)blockIdx.x * DS_AW + aw;
  const int stride = a.attn_coop ? (int)gridDim.x : (int)gridDim.x * DS_AW;
  for (int item = first; item < n_items; item += stride) {
    const int split = item % nsplit;
    const int bh = item / nsplit;
    const int h = bh % H, b = bh / H;
    const float rstd = rsqrtf(__ldcg(ss + b) * a.inv_d + a.eps);
    const float* row = a.qkv_acc + (size_t)b * 3 * HD + h * DS_HD + sub * 8;
    float q[8];
    {
      const float4 q0 = __ldcg(reinterpret_cast<const float4*>(row));
      const float4 q1 = __ldcg(reinterpret_cast<const float4*>(row) + 1);
      const float f = rstd * a.scale * LOG2E;
      q[0] = q0.x * f; q[1] = q0.y * f; q[2] = q0.z * f; q[3] = q0.w * f;
      q[4] = q1.x * f; q[5] = q1.y * f; q[6] = q1.z * f; q[7] = q1.w * f;
    }
    const float slope = (a.slopes ? __ldg(a.slopes + h) : 0.f) * LOG2E;
    __nv_bfloat16* kc = reinterpret_cast<__nv_bfloat16*>(a.cache) + (size_t)layer * a.cache_layer_stride +
                        ((size_t)b * H + h) * (size_t)a.Tmax * DS_HD;
    __nv_bfloat16* vc = kc + a.cache_kv_stride;
    const int j_begin = split * chunk;
    const int j_end = min(nkeys, j_begin + chunk);
    DsAttnState st;
    if (!a.attn_coop) {
      ds_attn_range(st, j_begin, j_end, pos, grp, sub, kc, vc, row, HD, rstd, q, slope);
      ds_attn_finish(a, st, bh, split, b, h, lane);
    } else {
      // the seven warps of this CTA split [j_begin, j_end) into contiguous parts (multiples of 4 keys)
      const int len = max(0, j_end - j_begin);
      const int part = ((len + DS_AW - 1) / DS_AW + 3) & ~3;
      const int j0 = min(j_end, j_begin + aw * part), j1 = min(j_end, j0 + part);
      ds_attn_range(st, j0, j1, pos, grp, sub, kc, vc, row, HD, rstd, q, slope);
      if (grp == 0) {
        if (sub == 0) { sh.am[aw] = st.m; sh.al[aw] = st.l; }
        *reinterpret_cast<float4*>(&sh.ao[aw][sub * 8]) = make_float4(st.acc[0], st.acc[1], st.acc[2], st.acc[3]);
        *reinterpret_cast<float4*>(&sh.ao[aw][sub * 8 + 4]) = make_float4(st.acc[4], st.acc[5], st.acc[6], st.acc[7]);
      }
      ds_named_bar(3, DS_AW * 32);
      if (aw == 0) {
        float M = -CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < DS_AW; ++w) M = fmaxf(M, sh.am[w]);
        st.m = M; st.l = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) st.acc[d] = 0.f;
#pragma unroll
        for (int w = 0; w < DS_AW; ++w) {
          const float wgt = (sh.am[w] == -CUDART_INF_F) ? 0.f : ex2_approx(sh.am[w] - M);
          st.l = fmaf(sh.al[w], wgt, st.l);
#pragma unroll
          for (int d = 0; d < 8; ++d) st.acc[d] = fmaf(sh.ao[w][sub * 8 + d], wgt, st.acc[d]);
        }
        ds_attn_finish(a, st, bh, split, b, h, lane);
      }
      ds_named_bar(3, DS_AW * 32);             // sh.am / al / ao are reused by the next item
    }
  }
}

__global__ void __launch_bounds__(DS_THREADS, 1) decode_step_kernel(const __grid_constant__ DsArgs a) {
  extern __shared__ uint8_t ds_smem_raw[];
  // SWIZZLE_128B atoms are 1024-byte aligned
  uint8_t* smem = ds_smem_raw + ((1024u - (smem_u32(ds_smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;                                        // DS_NSTAGES x 16 KB
  uint8_t* xs = ring + DS_NSTAGES * DS_STAGE;                  // 2 x 32 KB
  DsTask* tasks = reinterpret_cast<DsTask*>(xs + 2 * DS_XSLOT);
  DsShared& sh = *reinterpret_cast<DsShared*>(reinterpret_cast<uint8_t*>(tasks) + (size_t)a.NP * sizeof(DsTask));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NP = a.NP;
  // this CTA's row of the task table → shared memory (every role reads it, phase after phase)
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.tasks + (size_t)blockIdx.x * NP);
    uint4* dst = reinterpret_cast<uint4*>(tasks);
    const int n16 = NP * (int)(sizeof(DsTask) / 16);
    for (int i = tid; i < n16; i += DS_THREADS) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    for (int s = 0; s < DS_NSTAGES; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
    mbar_init(&sh.xfull[0], DS_WORKERS); mbar_init(&sh.xfull[1], DS_WORKERS);
    mbar_init(&sh.xempty[0], 1); mbar_init(&sh.xempty[1], 1);
    mbar_init(&sh.accfull, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 256; i += DS_THREADS) sh.ss[i] = 0.f;
  for (int i = tid; i < NP; i += DS_THREADS) sh.phase_kind[i] = __ldg(a.phase_kind + i);
  if (warp == 1) {
    tmem_alloc(&sh.tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;
  const unsigned epoch0 = *a.epoch;
  const int pos = min(*a.pos_dev, a.Tmax - 1);

  if (warp == 0) {
    // ===================== weight producer: this CTA's packed stream → ring, ahead of everything else =====================
    // (the whole warp walks the loop and one elected lane issues: inside a `lane == 0` branch every uniform-datapath
    //  instruction — UBLKCP, UTCHMMA, UTCBAR — is wrapped in a per-thread ELECT / R2UR loop of ~80 cycles)
    const uint8_t* src = a.wstream + a.wstream_off[blockIdx.x];
    uint32_t cnt = 0;
    for (int p = 0; p < NP; ++p) {
      const DsTask& t = tasks[p];
      if (t.R == 0) continue;
      const int slab = t.R * 128;
      const int wkb = max(1, DS_STAGE / slab);
      for (int j = 0; j < t.nkb; j += wkb) {
        const int n = min(wkb, t.nkb - j);
        const uint32_t bytes = (uint32_t)(n * slab);
        const int s = cnt % DS_NSTAGES;
        const uint32_t use = cnt / DS_NSTAGES;
        if (use > 0) ds_mbar_wait(&sh.empty[s], (use - 1) & 1, a, 1, p);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sh.full[s], bytes);
          ds_bulk_g2s(smem_u32(ring + s * DS_STAGE), src, bytes, &sh.full[s]);
        }
        __syncwarp();
        src += bytes;
        ++cnt;
      }
    }
  } else {
    uint32_t wcount = 0, xcount = 0, ntask = 0;       // ring stages / X slots / GEMM units consumed so far (role-local copies)
    const int cap = DS_XSLOT / (a.Bp * 128);
    const int mtiles = a.Bp > 128 ? 2 : 1;            // 128 batch rows per tcgen05 M tile
    int pnext = 0;                                    // next phase whose unit's vectors have not been prefetched yet
    if (warp == 2 || warp == 3) {
      while (pnext < NP && tasks[pnext].R == 0) ++pnext;
      if (pnext < NP) ds_prefetch_vec(tasks[pnext], sh, 0);
      ++pnext;
    }
    for (int p = 0; p < NP; ++p) {
      const DsTask& t = tasks[p];
      const int kind = sh.phase_kind[p];
      if (kind >= 0) {
        ds_attention(a, sh, kind, pos, p);            // warps 1..7
      } else if (t.R > 0) {
        const int xkb = t.nkb < cap ? t.nkb : cap;
        const int par = ntask & 1;
        if (warp == 1) {
          // ===================== MMA issuer: D[batch rows x R features] += X[128 x 64] · W_slab[R x 64]ᵀ per k-block ==========
          // (the batch is the M operand: the accumulator row of a thread is a batch row, its columns are consecutive
          //  features — contiguous in the row-major accumulators, so the epilogue reduces with 16-byte vector atomics)
          const uint32_t idesc = make_idesc_bf16(128, t.R, 0, 0);
          const int slab = t.R * 128;
          const int wkb = max(1, DS_STAGE / slab);
          for (int j = 0; j < t.nkb; ++j) {
            const int jw = j % wkb, jx = j % xkb;
            const int s = wcount % DS_NSTAGES;
            const int slot = xcount & 1;
            if (jw == 0) ds_mbar_wait(&sh.full[s], (wcount / DS_NSTAGES) & 1, a, 2, p);
            if (jx == 0) ds_mbar_wait(&sh.xfull[slot], (xcount >> 1) & 1, a, 3, p);
            tc_fence_after();
            const uint32_t w_addr = smem_u32(ring + s * DS_STAGE + jw * slab);
            const uint32_t x_addr = smem_u32(xs + slot * DS_XSLOT + jx * (a.Bp * 128));
            const bool w_done = (jw == wkb - 1 || j == t.nkb - 1), x_done = (jx == xkb - 1);
            if (elect_one()) {
              for (int mt = 0; mt < mtiles; ++mt) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_ss(tmem_base + (uint32_t)(mt * t.R), make_smem_desc_sw128(x_addr + mt * 16384 + k * 32, 16, 1024),
                              make_smem_desc_sw128(w_addr + k * 32, 16, 1024), idesc, (j | k) != 0);
              }
              if (w_done) umma_commit(&sh.empty[s]);
              if (x_done) umma_commit(&sh.xempty[slot]);
              if (j == t.nkb - 1) umma_commit(&sh.accfull);
            }
            __syncwarp();
            if (w_done) ++wcount;
            if (x_done) ++xcount;
          }
          DS_TRACE(3, 32);
        } else {
          // ===================== workers: X operand, then (warps 4..7) the accumulator, (warps 2..3) prefetch + aux ==========
          ds_transform(a, t, sh, xs, xcount, par, p);
          if (warp >= 4) {
            const int q = warp - 4;
            // With a.rep copies of the batch in the accumulator (rows b + i * 128 / rep), quadrant q reads copy
            // q / (4 / rep) and takes every rep-th 32-column chunk: at batch <= 32 all four epilogue warps work on
            // different columns of the same rows instead of one warp doing everything.
            const int qper = 4 / a.rep;                         // quadrants per copy
            const int copy = q / qper, qrow = (q % qper) * 32;  // first batch row of this warp within its copy
            if (qrow < a.B || mtiles == 2) {
              ds_mbar_wait(&sh.accfull, ntask & 1, a, 4, p);
              tc_fence_after();
              DS_TRACE(4, 128);
              for (int mt = 0; mt < mtiles; ++mt) {
                const int row0 = mt * 128 + qrow;
                if (row0 >= a.B) break;
                const bool vec4 = ((t.ldacc | t.rows) & 3) == 0;  // 16-byte vector atomics: 16-byte aligned rows, whole groups
                // 32 columns at a time: thread = batch row (TMEM lane) holds 32 consecutive features.  One lane per row
                // issuing the atomics is slow when few rows are valid, so the chunk goes through a warp-private,
                // XOR-swizzled 4 KB staging tile and comes back as (row, 4 features) per lane: 8 consecutive lanes cover
                // 128 contiguous bytes of one accumulator row.
                float4* stg = reinterpret_cast<float4*>(xs) + q * 256;      // X slots are idle once the accumulator is complete
                const int rows_here = min(32, a.B - row0);
                for (int c0 = copy * 32; c0 < t.rows; c0 += 32 * a.rep) {
                  uint32_t v[32];
                  tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * t.R + c0), v);
                  tmem_ld_wait();
                  __syncwarp();
                  if (lane < rows_here) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                      stg[lane * 8 + (g ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                                                    __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
                  }
                  __syncwarp();
                  const int g = lane & 7;
                  if (c0 + 4 * g < t.rows) {
                    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t.bias_out) bb = *reinterpret_cast<const float4*>(&sh.bias[par][c0 + 4 * g]);
                    for (int r = lane >> 3; r < rows_here; r += 4) {
                      float4 val = stg[r * 8 + (g ^ (r & 7))];
                      val.x += bb.x; val.y += bb.y; val.z += bb.z; val.w += bb.w;
                      float* dst = t.acc + (size_t)(row0 + r) * t.ldacc + t.n0 + c0 + 4 * g;
                      if (!vec4) {
                        const float e[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                          if (c0 + 4 * g + i >= t.rows) break;       // ragged tail of the last feature slab
                          if (t.store) __stcg(dst + i, e[i]); else ds_red_add(dst + i, e[i]);
                        }
                      } else if (t.store) {
                        __stcg(reinterpret_cast<float4*>(dst), val);
                      } else {
                        ds_red_add4(dst, val.x, val.y, val.z, val.w);
                      }
                    }
                  }
                }
              }
              tc_fence_before();
              DS_TRACE(5, 128);
            }
          } else {
            // the NEXT unit's vectors, into the other buffer (this unit's bias is still being read by the epilogue warps)
            while (pnext < NP && tasks[pnext].R == 0) ++pnext;
            if (pnext < NP) ds_prefetch_vec(tasks[pnext], sh, par ^ 1);
            ++pnext;
          }
        }
        ++ntask;
      }
      if (t.aux_kind && (warp == 2 || warp == 3)) ds_aux(t);      // these two warps have no accumulator to drain
      if (p + 1 < NP) ds_grid_barrier(a, epoch0 + (unsigned)p + 1u, p);
    }
  }
  // ---- teardown: advance the launch epoch and the cache position for the next launch; free TMEM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
  if (blockIdx.x == 0 && tid == 0) {
    // every CTA has read *epoch / *pos_dev before it could pass its first barrier, and this CTA is past the last one
    *a.epoch = epoch0 + (unsigned)(NP - 1);                      // NP - 1 barriers per launch
    if (a.advance_pos) *a.pos_dev = pos + 1;
  }
}

}  // namespace vg

using namespace vg;

extern "C" size_t vg_decode_step_task_bytes(void) { return sizeof(vg_decode_step_task); }

extern "C" size_t vg_decode_step_smem_bytes(int32_t n_phases) {
  return 1024 + (size_t)DS_NSTAGES * DS_STAGE + 2 * DS_XSLOT + (size_t)n_phases * sizeof(vg_decode_step_task) +
         sizeof(DsShared) + 64;
}

extern "C" int vg_decode_step(const vg_decode_step_args* a, vg_stream_t stream) {
  VG_REQUIRE(a && a->tasks && a->phase_kind && a->wstream && a->wstream_off && a->epoch && a->pos_dev && a->bar_flags, -1,
             "vg_decode_step: null pointer");
  VG_REQUIRE(a->B >= 1 && a->B <= 256 && a->Bp >= 16 && a->Bp <= 256 && (a->Bp & (a->Bp - 1)) == 0 && a->Bp >= a->B, -3,
             "vg_decode_step: batch %d / padded %d (power of two in [16,256])", a->B, a->Bp);
  VG_REQUIRE(a->NP >= 1 && a->NP <= 128 && a->grid >= 1, -3, "vg_decode_step: bad phase count / grid");
  VG_REQUIRE(a->nsplit >= 1 && a->nsplit <= 64, -3, "vg_decode_step: nsplit must be in [1,64]");
  VG_REQUIRE((a->rep == 1 || a->rep == 2 || a->rep == 4) && (a->rep == 1 || (a->B * a->rep <= 128 && a->Bp == 128)), -3,
             "vg_decode_step: rep %d copies of %d rows do not fit a 128-row tile", a->rep, a->B);
  const size_t smem = vg_decode_step_smem_bytes(a->NP);
  VG_REQUIRE(smem <= 227 * 1024, -6, "vg_decode_step: %d phases need %zu bytes of shared memory", a->NP, smem);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    VG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VG_CUDA(cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  VG_REQUIRE(a->grid <= sms, -3, "vg_decode_step: grid %d exceeds the %d SMs (all CTAs must be co-resident)", a->grid, sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)a->grid);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;       // co-residency of the whole grid is guaranteed or the launch fails
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VG_CUDA(cudaLaunchKernelEx(&cfg, decode_step_kernel, *a));
  return 0;
}
