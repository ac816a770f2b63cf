// decode_linear.cu — weight-streaming linear layer for the cached generation step (LVTR.step, lvtr.py:227-286):
// every nn.Linear of the single-token step — in_proj / out_proj (attention.py:52,79), linear1 / linear2
// (transformer/layers.py:82), the stack input linear (:152), q_spliter / token_spliter / prior head / token_predictor
// (lvtr.py:171,172,194,195) — when the "M" dimension is the decode batch (1..256 rows).
//
// Such a GEMM is HBM-bound: 2·N·K bytes of bf16 weights against a few hundred KB of activations.  The training GEMM
// (gemm_tc.cu) would put an M=64 problem on N/128 SMs; here the weight matrix is cut into (Nc features x KL inputs)
// slabs, one CTA per slab, so that every SM streams a disjoint part of W exactly once:
//   1. the CTA's W slab goes global → shared memory with cp.async BEFORE the programmatic-dependent-launch wait
//      (griddepcontrol.wait): weights do not depend on the previous kernel, so the HBM stream of layer i+1 overlaps
//      the dependent tail (activation load → MMA → epilogue) of layer i;
//   2. after the wait, the [B x KL] activation slice is staged, optionally scaled by the RMSNorm weight (the
//      per-row 1/rms factor is linear and is applied in the epilogue from the row sum-of-squares the PRODUCER of x
//      accumulated — norm.py:28-32 fused away);
//   3. warps run mma.sync m16n8k16 (bf16 → f32) straight from shared memory, reduce their k-slices with shared-memory
//      atomics, and — when K is split across CTAs — with f32 atomics into a global accumulator whose last-arriving
//      CTA (ticket counter per feature range) runs the epilogue: ·rstd, +bias, ReLU/GELU, +residual, bf16/f32 store,
//      and the row sum-of-squares for the next RMSNorm.  Accumulator and counters are left zeroed for the next call.
// The tensor-core path here is mma.sync on purpose: at M ≤ 256 the kernel is bound by HBM (and by the dependent-launch
// latency chain), not by MMA issue; tcgen05 would add TMEM round trips to a 2 µs kernel.
#include <stdlib.h>
#include "common.cuh"

namespace vg {

constexpr int DL_THREADS = 256;
constexpr int DL_WARPS = DL_THREADS / 32;
constexpr int DL_PAD = 32;            // bf16 elements (64 B) of row padding: rows r, r+1 land in opposite bank halves

struct DlParams {
  const __nv_bfloat16* x; int64_t ldx;
  const __nv_bfloat16* W; int64_t ldw;
  const float* norm_scale; const float* x_ss; float norm_eps;
  const float* bias; int act;
  const __nv_bfloat16* residual; int64_t ld_res;
  __nv_bfloat16* y; int64_t ldy;
  float* y_f32; int64_t ldy32;
  float* y_ss; float* zero_ss;
  float* acc; int* counters;
  int B, N, K, Nc, KL, ksplit, Bp, ksub;
  unsigned long long* trace;          // debug: per-phase clock64 stamps of CTA 0 (vg_debug_decode_linear_trace)
};
#ifndef VG_DL_UNIFORM_ISSUE
#define VG_DL_UNIFORM_ISSUE 0
#endif
#if VG_DL_UNIFORM_ISSUE
__device__ __forceinline__ bool dl_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
#define DL_ISSUER() (dl_elect_one())
#else
#define DL_ISSUER() (lane == 0)
#endif
#define DL_STAMP(i) do { if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[i] = clock64(); } while (0)

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t dl_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// 1-D bulk copy global → shared (TMA engine, no LSU instructions per 16 bytes), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dl_smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(dl_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dl_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dl_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dl_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dl_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(dl_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t dl_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void dl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 dl_ld_cluster_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  float4 v;      // not volatile: the partial sums are final after the cluster barrier, so the loads may be batched
  asm("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra));
  return v;
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// bias + activation + residual of one output element.  Deliberately NOT inlined: the epilogue calls it eight times and
// the body then runs from the instruction cache (this kernel executes every code path exactly once per CTA, so
// straight-line code is a chain of cold instruction-cache misses).
__device__ __noinline__ float dl_finish_elem(float t, const float* bias_ptr, int act, const __nv_bfloat16* res_ptr) {
  if (bias_ptr) t += *bias_ptr;
  if (act != VG_ACT_NONE) t = apply_act(t, act);
  if (res_ptr) t += __bfloat162float(*res_ptr);
  return t;
}

// NTC = n8-tiles per CTA (Nc = 8·NTC features).  Grid = feature ranges x ksplit; the ksplit CTAs of one feature
// range form a thread-block cluster and reduce their k-slices through distributed shared memory.
template <int NTC>
__global__ void __launch_bounds__(DL_THREADS)
decode_linear_kernel(const __grid_constant__ DlParams p) {
  extern __shared__ __align__(16) uint8_t dl_smem[];
  constexpr int Nc = NTC * 8;
  const int KLP = p.KL + DL_PAD;
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(dl_smem);                       // [Nc][KLP]
  __nv_bfloat16* Xsm = Wsm + (size_t)Nc * KLP;                                          // [Bp][KLP]
  __nv_bfloat16* Ssm = Xsm + (size_t)p.Bp * KLP;                                        // [KL] norm scale (bf16)
  float* Asm = reinterpret_cast<float*>(Ssm + p.KL);                                    // [Bp][Nc] partial sums
  __shared__ __align__(8) uint64_t bars[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_range = blockIdx.x / p.ksplit, ks_cta = blockIdx.x % p.ksplit;   // ks_cta == rank in the cluster
  const int n0 = n_range * Nc;
  const int k0 = ks_cta * p.KL;
  const uint32_t row_bytes = (uint32_t)p.KL * 2;

  DL_STAMP(0);
  // Kernel parameters live in the constant bank and every launch starts with a cold constant cache: the epilogue's
  // first touch of bias / act / residual / ldy ... was a chain of dependent ~500-cycle misses (2.6 k cycles measured).
  // Touch every 64-byte line of the parameter block now, while nothing depends on it.
  {
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(&p);
    uint32_t sink = 0;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(DlParams) / 4); i += 16) sink |= pw[i];
    sink |= pw[sizeof(DlParams) / 4 - 1];
    if (sink == 0x7fc0dead && p.trace) p.trace[39] = sink;      // never true in practice; keeps the loads alive
  }
  pdl_launch_dependents();                         // the next kernel may begin ITS weight prefetch
  if (tid == 0) {
    dl_mbar_init(&bars[0], 1);
    dl_mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // ---- 1. weight slab: independent of the previous kernel.  Rows beyond N / B are never loaded: an MMA row or
  //         column only feeds its own output element, and those outputs are discarded.
  // (bulk copies are uniform-datapath instructions: a warp issues them one lane at a time, so one lane per warp
  //  issues and the rows are spread over all eight warps)
  const int w_rows = min(Nc, p.N - n0);
  if (tid == 0) dl_mbar_expect_tx(&bars[0], (uint32_t)w_rows * row_bytes);
  __syncthreads();
  // -DVG_DL_UNIFORM_ISSUE=1 (experiment, default off): elect.sync in warp-uniform code instead of a `lane == 0` branch, in
  // which the compiler wraps every UBLKCP in a per-thread ELECT / R2UR loop — the ~55 cycles per copy measured above
  // (same effect and fix as for tcgen05.mma in attn_tc.cu; not yet measured here)
  if (DL_ISSUER())
    for (int row = warp; row < w_rows; row += DL_WARPS)
      bulk_g2s(Wsm + (size_t)row * KLP, p.W + (int64_t)(n0 + row) * p.ldw + k0, row_bytes, &bars[0]);
  DL_STAMP(1);
  // ---- 2. everything below reads what the previous kernel wrote
  pdl_wait();
  DL_STAMP(2);
  if (tid == 0) dl_mbar_expect_tx(&bars[1], (uint32_t)p.B * row_bytes);     // (bars[1] is only waited on below)
  if (DL_ISSUER()) {
    // the expect_tx above must precede the copies' complete_tx only in the sense that the phase cannot complete
    // before both happened: the pending arrival count (1) is consumed by that expect_tx arrive
    for (int row = warp; row < p.B; row += DL_WARPS)
      bulk_g2s(Xsm + (size_t)row * KLP, p.x + (int64_t)row * p.ldx + k0, row_bytes, &bars[1]);
  }
  if (p.norm_scale)
    for (int i = tid; i < p.KL; i += DL_THREADS) Ssm[i] = __float2bfloat16_rn(p.norm_scale[k0 + i]);
  if (p.zero_ss && blockIdx.x == 0)
    for (int i = tid; i < p.B; i += DL_THREADS) p.zero_ss[i] = 0.f;
  DL_STAMP(3);
  __syncthreads();                                 // Ssm visible
  dl_mbar_wait(&bars[0], 0);
  if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[16 + warp] = clock64();     // per-warp: W landed
  dl_mbar_wait(&bars[1], 0);
  if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[24 + warp] = clock64();     // per-warp: x landed
  DL_STAMP(4);

  // ---- 3. MMA: item = (16-row batch tile, k sub-slice); a warp sweeps all NTC feature tiles of the CTA per item
  {
    const int r = lane >> 2, c = lane & 3;
    const int MT = p.Bp / 16;
    const int kb_per = (p.KL / 32) / p.ksub;
    const bool has_norm = p.norm_scale != nullptr;
    for (int item = warp; item < MT * p.ksub; item += DL_WARPS) {
      const int mt = item / p.ksub, ks = item - mt * p.ksub;
      float d[NTC][4];
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) { d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f; }
      const __nv_bfloat16* xa = Xsm + (size_t)(mt * 16 + r) * KLP + c * 8;
      const __nv_bfloat16* wa = Wsm + (size_t)r * KLP + c * 8;
      const __nv_bfloat16* sa = Ssm + c * 8;
      // NOT unrolled: every CTA runs this kernel's code exactly once, so each distinct instruction line is a cold
      // instruction-cache miss (measured: ~25 cycles per executed instruction in straight-line code); a compact loop
      // body pays that once and then runs from L0
#pragma unroll 1
      for (int kb = ks * kb_per; kb < (ks + 1) * kb_per; ++kb) {
        // one 16-byte load = 8 consecutive k of one row; the k → fragment-slot assignment is a permutation applied
        // identically to both operands (dot products do not care), so no ldmatrix / transposes are needed
        const int ko = kb * 32;
        uint4 alo = *reinterpret_cast<const uint4*>(xa + ko);
        uint4 ahi = *reinterpret_cast<const uint4*>(xa + (size_t)8 * KLP + ko);
        if (has_norm) {
          const uint4 s = *reinterpret_cast<const uint4*>(sa + ko);
          alo.x = hmul2_bf16(alo.x, s.x); alo.y = hmul2_bf16(alo.y, s.y);
          alo.z = hmul2_bf16(alo.z, s.z); alo.w = hmul2_bf16(alo.w, s.w);
          ahi.x = hmul2_bf16(ahi.x, s.x); ahi.y = hmul2_bf16(ahi.y, s.y);
          ahi.z = hmul2_bf16(ahi.z, s.z); ahi.w = hmul2_bf16(ahi.w, s.w);
        }
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) {
          const uint4 w = *reinterpret_cast<const uint4*>(wa + (size_t)(nt * 8) * KLP + ko);
          mma_bf16_16816(d[nt], alo.x, ahi.x, alo.y, ahi.y, w.x, w.y);
          mma_bf16_16816(d[nt], alo.z, ahi.z, alo.w, ahi.w, w.z, w.w);
        }
      }
      // every (batch tile, k sub-slice) item owns its own [16 x Nc] slab of Asm[ks][Bp][Nc]: plain stores (float
      // atomics on shared memory compile to CAS loops and were 27 % of this kernel's stall samples)
      const int row0 = mt * 16 + r;
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        float* a0 = Asm + ((size_t)ks * p.Bp + row0) * Nc + nt * 8 + 2 * c;
        float* a1 = a0 + (size_t)8 * Nc;
        *reinterpret_cast<float2*>(a0) = make_float2(d[nt][0], d[nt][1]);
        *reinterpret_cast<float2*>(a1) = make_float2(d[nt][2], d[nt][3]);
      }
    }
  }
  if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[8 + warp] = clock64();      // per-warp end of MMA
  // ---- 4. k reduction + epilogue.  First every CTA folds its own k sub-slices into slice 0 (all threads), then the
  //         cluster barrier, then every CTA finishes a 1/ksplit share of the [B x Nc] outputs, reading the other
  //         CTAs' slice 0 over DSMEM (loads batched four ranks at a time: a DSMEM load is ~250 cycles).
  __syncthreads();
  if (p.ksub > 1) {
    const int n4 = p.B * (Nc / 4);
#pragma unroll 1
    for (int i = tid; i < n4; i += DL_THREADS) {
      float4 a = *reinterpret_cast<const float4*>(Asm + (size_t)i * 4);
#pragma unroll 1
      for (int ks = 1; ks < p.ksub; ++ks) {
        const float4 q4 = *reinterpret_cast<const float4*>(Asm + ((size_t)ks * p.Bp * Nc) + (size_t)i * 4);
        a.x += q4.x; a.y += q4.y; a.z += q4.z; a.w += q4.w;
      }
      *reinterpret_cast<float4*>(Asm + (size_t)i * 4) = a;
    }
  }
  if (p.ksplit > 1) dl_cluster_sync(); else __syncthreads();
  DL_STAMP(5);
  {
    // ONE output element per thread and a rolled loop: this kernel is instruction-fetch bound (cold code runs at
    // ~10-25 cycles per instruction; an epilogue unrolled over 8 elements per thread cost 2.6 k cycles), so the work is
    // spread over as many threads as possible and every thread executes as few distinct instructions as possible.
    // Element e = b * Nc + n; the CTAs of a cluster interleave rows of 32 consecutive elements.
    if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[34] = clock64();        // epilogue loop entered
    // Four consecutive outputs per thread (one 16-byte DSMEM / shared load per partial sum: scalar remote loads were
    // transaction-bound at batch 64), warps of a cluster's CTAs interleaved; group gi covers elements 4·gi .. 4·gi+3
    // of the row-major [B x Nc] tile.
    const int total4 = p.B * (Nc / 4);
    const float inv_k = 1.0f / (float)p.K;
    const uint32_t asm_addr = dl_smem_u32(Asm);
    const int lane_e = tid & 31;
    constexpr int G4R = Nc / 4;                    // groups per batch row
#pragma unroll 1
    for (int g0 = ((tid >> 5) * p.ksplit + ks_cta) * 32; g0 < total4; g0 += (DL_THREADS / 32) * p.ksplit * 32) {
      const int gi = g0 + lane_e;                  // (no early exits: the row-sum below shuffles across the warp)
      const int b = gi / G4R, n = (gi - b * G4R) * 4;
      const int col = n0 + n;
      const bool ok = gi < total4 && col < p.N;
      const uint32_t off = (uint32_t)(ok ? gi : 0) * 16u;
      float4 v4 = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(Asm) + off);
#pragma unroll 1
      for (int q = 1; q < p.ksplit; ++q) {         // the other CTAs' partial sums (slice 0 after the local fold)
        const float4 r = dl_ld_cluster_f4(asm_addr + off, (uint32_t)((ks_cta + q) % p.ksplit));
        v4.x += r.x; v4.y += r.y; v4.z += r.z; v4.w += r.w;
      }
      if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[32] = clock64();      // after the partial-sum reads
      float ss = 0.f;
      if (ok) {
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
        const float rs = p.norm_scale ? rsqrtf(p.x_ss[b] * inv_k + p.norm_eps) : 1.0f;
        const int nj = min(4, p.N - col);
        __nv_bfloat16 qv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = v[j] * rs;
          if (j < nj) {
            if (p.bias) t += p.bias[col + j];
            if (p.act != VG_ACT_NONE) t = apply_act(t, p.act);
            if (p.residual) t += __bfloat162float(p.residual[(int64_t)b * p.ld_res + col + j]);
          }
          v[j] = t;
          qv[j] = __float2bfloat16_rn(t);
          const float f = __bfloat162float(qv[j]);
          if (j < nj) ss = fmaf(f, f, ss);
        }
        if (p.y) {
          __nv_bfloat16* yp = p.y + (int64_t)b * p.ldy + col;
          if (nj == 4 && ((reinterpret_cast<uintptr_t>(yp) & 7) == 0)) {
            *reinterpret_cast<uint2*>(yp) = *reinterpret_cast<const uint2*>(qv);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < nj) yp[j] = qv[j];
          }
        } else {
          ss = 0.f;
        }
        if (p.y_f32) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nj) p.y_f32[(int64_t)b * p.ldy32 + col + j] = v[j];
        }
      }
      if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[33] = clock64();      // after bias / activation / residual
      if (p.y_ss && p.y) {
        // lanes of a warp cover 128 consecutive elements: whole rows when Nc <= 128 divides 128 (always: Nc is a
        // power of two <= 64), i.e. 128 / Nc rows per warp → segmented sum over G4R consecutive lanes
        for (int o = 1; o < G4R && o < 32; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((lane_e & (G4R - 1)) == 0 && ok) atomicAdd(p.y_ss + b, ss);
      }
    }
  }
  DL_STAMP(6);
  if (p.ksplit > 1) dl_cluster_sync();             // peers may still be reading this CTA's partial sums
}

struct DlPlan { int Nc, KL, ksub; size_t smem; };

static int dl_ksub(int Bp, int KL) {
  const int MT = Bp / 16;
  int ksub = 1;
  while (MT * ksub < DL_WARPS && (KL / 32) % (ksub * 2) == 0) ksub *= 2;
  return ksub;
}
static size_t dl_smem_bytes(int Bp, int Nc, int KL) {
  return (size_t)(Nc + Bp) * (KL + DL_PAD) * 2 + (size_t)KL * 2 + (size_t)dl_ksub(Bp, KL) * Bp * Nc * 4 + 16;
}

static DlPlan dl_plan(int64_t B, int64_t N, int64_t K) {
  const int Bp = (int)((B + 15) / 16 * 16);
  DlPlan best{0, 0, 1, 0};
  double best_cost = 1e30;
  // CTAs of one feature range = one cluster: portable sizes up to 8; 16 (opt-in on B200) only when the activation
  // slice of a smaller split cannot fit shared memory (batch 256 with K = 4096)
  for (int max_split = 8; max_split <= 16 && !best.Nc; max_split *= 2) {
    for (int Nc = 8; Nc <= 64; Nc *= 2) {
      for (int ksplit = 1; ksplit <= max_split; ksplit *= 2) {
        if (K % (64 * ksplit) != 0) continue;
        const int64_t KL = K / ksplit;
        const size_t smem = dl_smem_bytes(Bp, Nc, (int)KL);
        if (smem > 200 * 1024) continue;
        const int64_t units = ceil_div(N, Nc) * ksplit;
        const int64_t waves = ceil_div(units, kNumSMs);
        // per-wave time ~ bytes that must enter this SM's shared memory (weight slab and activation slice take the
        // same ~40-60 B/clk path; measured: a 128 KB activation slice costs ~3.7 k cycles) + DSMEM reduction + fixed
        const double bytes = (double)Nc * KL * 2 + (double)B * KL * 2 + (ksplit > 1 ? 3.0 * Bp * Nc * 4 : 0.0);
        const double cost = (double)waves * (bytes + 32000.0);
        if (cost < best_cost) { best_cost = cost; best = DlPlan{Nc, (int)KL, 1, smem}; }
      }
    }
  }
  if (best.Nc) best.ksub = dl_ksub(Bp, best.KL);
  return best;
}

}  // namespace vg

using namespace vg;

static unsigned long long* g_dl_trace = nullptr;
extern "C" int vg_debug_decode_linear_trace(void* buf) {
  g_dl_trace = (unsigned long long*)buf;
  return 0;
}

extern "C" size_t vg_decode_linear_workspace(int64_t max_batch, int64_t max_n) {
  // the k-slices of one feature range reduce through distributed shared memory: no global scratch is needed any more;
  // the parameter stays in the ABI for layouts that will (a small non-zero size keeps allocation code uniform)
  (void)max_batch; (void)max_n;
  return 64;
}

extern "C" int vg_decode_linear(const vg_decode_linear_args* a, void* workspace, size_t workspace_bytes,
                                vg_stream_t stream) {
  VG_REQUIRE(a && a->x && a->W && (a->y || a->y_f32), -1, "vg_decode_linear: null pointer");
  VG_REQUIRE(a->B >= 1 && a->B <= 256 && a->N >= 1 && a->K >= 64, -3, "vg_decode_linear: bad shape B=%lld N=%lld K=%lld",
             (long long)a->B, (long long)a->N, (long long)a->K);
  VG_REQUIRE(a->K % 64 == 0 && a->ldx % 8 == 0 && a->ldw % 8 == 0 && aligned(a->x, 16) && aligned(a->W, 16), -4,
             "vg_decode_linear: K must be a multiple of 64 and x / W rows 16-byte aligned");
  VG_REQUIRE(!a->norm_scale || a->x_ss, -1, "vg_decode_linear: norm_scale needs the row sum-of-squares x_ss");
  VG_REQUIRE(a->act >= VG_ACT_NONE && a->act <= VG_ACT_SILU, -3, "vg_decode_linear: bad activation");
  VG_REQUIRE(workspace && workspace_bytes >= vg_decode_linear_workspace(a->B, a->N), -5,
             "vg_decode_linear: workspace too small");
  const DlPlan plan = dl_plan(a->B, a->N, a->K);
  static const bool debug = getenv("VG_DL_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "vg_decode_linear B=%lld N=%lld K=%lld -> Nc=%d KL=%d ksplit=%lld ksub=%d smem=%zu grid=%lld\n",
            (long long)a->B, (long long)a->N, (long long)a->K, plan.Nc, plan.KL, (long long)(a->K / (plan.KL ? plan.KL : 1)),
            plan.ksub, plan.smem, (long long)(ceil_div(a->N, plan.Nc ? plan.Nc : 1) * (a->K / (plan.KL ? plan.KL : 1))));
  VG_REQUIRE(plan.Nc > 0, -6, "vg_decode_linear: no slab shape fits shared memory for B=%lld K=%lld", (long long)a->B,
             (long long)a->K);
  DlParams p;
  p.x = (const __nv_bfloat16*)a->x; p.ldx = a->ldx;
  p.W = (const __nv_bfloat16*)a->W; p.ldw = a->ldw;
  p.norm_scale = a->norm_scale; p.x_ss = a->x_ss; p.norm_eps = a->norm_eps;
  p.bias = a->bias; p.act = a->act;
  p.residual = (const __nv_bfloat16*)a->residual; p.ld_res = a->ld_res;
  p.y = (__nv_bfloat16*)a->y; p.ldy = a->ldy;
  p.y_f32 = a->y_f32; p.ldy32 = a->ldy_f32;
  p.y_ss = a->y_ss; p.zero_ss = a->zero_ss;
  p.acc = nullptr;
  p.counters = nullptr;
  p.B = (int)a->B; p.N = (int)a->N; p.K = (int)a->K;
  p.Nc = plan.Nc; p.KL = plan.KL; p.ksplit = (int)(a->K / plan.KL); p.Bp = (int)((a->B + 15) / 16 * 16);
  p.ksub = plan.ksub;
  p.trace = g_dl_trace;

  void (*kern)(DlParams) = plan.Nc == 8 ? decode_linear_kernel<1> : plan.Nc == 16 ? decode_linear_kernel<2>
                           : plan.Nc == 32 ? decode_linear_kernel<4> : decode_linear_kernel<8>;
  static bool attr_set[4] = {false, false, false, false};
  const int ki = plan.Nc == 8 ? 0 : plan.Nc == 16 ? 1 : plan.Nc == 32 ? 2 : 3;
  if (!attr_set[ki]) {
    VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set[ki] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ceil_div(a->N, plan.Nc) * p.ksplit));
  cfg.blockDim = dim3(DL_THREADS);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.ksplit > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)p.ksplit;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (a->allow_overlap) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  VG_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return 0;
}
