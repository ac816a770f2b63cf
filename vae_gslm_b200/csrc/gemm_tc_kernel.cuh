// gemm_tc_kernel.cuh — the tcgen05 bf16 GEMM kernel template (see gemm_tc.cu for the overview).  Included by the
// per-operand-layout translation units gemm_tc_{kk,kmn,mnk,mnmn}.cu so that they compile in parallel.
#pragma once
#ifndef VG_GEMM_UNIFORM_ISSUE
#define VG_GEMM_UNIFORM_ISSUE 1
#endif
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "sm100.cuh"

namespace vg {
using namespace sm100;

constexpr int TBM = 128;            // tile rows  (UMMA M)
constexpr int TBK = 64;             // k-block: 64 bf16 = 128 B = one swizzle atom
constexpr int TC_THREADS = 640;
constexpr int TC_EPI_THREADS = 512;         // warps 4..19: four per TMEM lane quadrant, each a quarter of the columns
constexpr int A_TILE_BYTES = TBM * TBK * 2;   // 16 KB

// CTAS = 2: a CTA pair (cta_group::2) computes one 256 x BN tile; each CTA holds 128 rows of A and BN/2 columns of B,
// so per MMA cycle only half of B crosses each SM's shared memory (the 1-CTA kernel is shared-memory-bandwidth bound).
template <int BN, int CTAS = 1> struct TcCfg {
  static constexpr int kBBytes = (BN / CTAS) * TBK * 2;
  static constexpr int kStageBytes = A_TILE_BYTES + kBBytes;
  static constexpr int kStages = (kStageBytes >= 49152) ? 4 : (kStageBytes >= 32768 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kScratchOff = kStages * kStageBytes + 256 /* barriers */;
  static constexpr int kSmemBytes = kScratchOff + (TC_EPI_THREADS / 32) * 2048 /* epilogue scratch */ + 1024 /* align slack */;
};

// Epilogue variants with the activation resolved at compile time.  The generic variant reads act/dact at run time
// (one uniform branch per element: correct for every combination the ABI allows, but it serialises the MUFU/FMA
// chains of neighbouring elements); the specialised ones are straight-line code over 8 independent elements.
enum TcVariant : int {
  TCV_GENERIC = 0,
  TCV_PLAIN,       // act = none, dact_src = null
  TCV_RELU, TCV_GELU, TCV_SILU,   // act = X, dact_src = null
  TCV_MULT,        // act = none, dact = VG_ACT_MULT (stored derivative)
  TCV_RED,         // split-K partial: C += acc via red.global.add (f32 C, no other epilogue term)
};

struct TcEpilogue {
  EpilogueParams p;
  int vec_ok;    // every pointer/ld allows 8-wide vector access
  int variant;   // TcVariant
  int splits;    // split-K factor (1 = off); >1 requires TCV_RED
  int pdl;       // vg_set_pdl_mode at launch time (0 = plain stream order)
};

// ---- warp-private transpose scratch ---------------------------------------------------------------------------
// The accumulator arrives thread-per-row (tcgen05.ld 32x32b: lane i holds row i), but a warp-wide global access in
// that layout touches 32 different rows, 16 bytes each: 32 half-used sectors per request, which throttled every
// epilogue that moves more than the plain C tile (GELU + saved derivative ran at 0.4x the plain kernel).  Each
// epilogue warp therefore owns a 32-row x 64-byte shared-memory scratch: rows are written thread-per-row, read back
// with S lanes per row, and go to / come from global memory as full 32- or 64-byte row segments (8 or 16 rows per
// request, every sector fully used).  XOR-swizzled 16-byte slots keep both access patterns bank-conflict free.
constexpr int TC_SCR_BYTES = 2048;

template <int S> __device__ __forceinline__ uint32_t scr_off(int row, int slot) {
  return (uint32_t)(row * (S * 16) + ((slot ^ ((row >> (S == 4 ? 1 : 2)) & (S - 1))) << 4));
}
template <int S> __device__ __forceinline__ void scr_put_row(uint8_t* scr, int lane, const uint4* q) {
#pragma unroll
  for (int g = 0; g < S; ++g) *reinterpret_cast<uint4*>(scr + scr_off<S>(lane, g)) = q[g];
}
template <int S> __device__ __forceinline__ void scr_get_row(const uint8_t* scr, int lane, uint4* q) {
#pragma unroll
  for (int g = 0; g < S; ++g) q[g] = *reinterpret_cast<const uint4*>(scr + scr_off<S>(lane, g));
}
// rows [0, rows_valid) of the scratch <-> global rows of `pitch` bytes starting at gbase; S lanes per row
template <int S>
__device__ __forceinline__ void staged_store(uint8_t* scr, int lane, const uint4* q, uint8_t* gbase, int64_t pitch,
                                             int rows_valid) {
  scr_put_row<S>(scr, lane, q);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < S; ++it) {
    const int row = it * (32 / S) + lane / S, slot = lane % S;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(gbase + row * pitch + slot * 16) =
          *reinterpret_cast<const uint4*>(scr + scr_off<S>(row, slot));
  }
  __syncwarp();
}
template <int S>
__device__ __forceinline__ void staged_load(uint8_t* scr, int lane, uint4* q, const uint8_t* gbase, int64_t pitch,
                                            int rows_valid) {
#pragma unroll
  for (int it = 0; it < S; ++it) {
    const int row = it * (32 / S) + lane / S, slot = lane % S;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(scr + scr_off<S>(row, slot)) =
          *reinterpret_cast<const uint4*>(gbase + row * pitch + slot * 16);
  }
  __syncwarp();
  scr_get_row<S>(scr, lane, q);
  __syncwarp();
}
template <int S>
__device__ __forceinline__ void staged_red_f32(uint8_t* scr, int lane, const uint4* q, uint8_t* gbase, int64_t pitch,
                                               int rows_valid) {
  scr_put_row<S>(scr, lane, q);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < S; ++it) {
    const int row = it * (32 / S) + lane / S, slot = lane % S;
    if (row < rows_valid) {
      const float4 f = *reinterpret_cast<const float4*>(scr + scr_off<S>(row, slot));
      atomicAdd(reinterpret_cast<float4*>(gbase + row * pitch + slot * 16), f);
    }
  }
  __syncwarp();
}
// one 16-byte slot <-> floats (8 bf16 or 4 f32)
template <typename TC> struct Slot;
template <> struct Slot<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ uint4 pack(const float* v) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return raw;
  }
  static __device__ __forceinline__ void unpack(const uint4& raw, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
};
template <> struct Slot<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
  static __device__ __forceinline__ void unpack(const uint4& raw, float* v) {
    v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
  }
};

// One segment (S 16-byte slots = S*Slot::N columns) of one 32-row slab.  `v` holds the accumulators of this lane's
// row; every lane of the warp must call this (the staging uses __syncwarp), rows beyond M are masked by rows_valid.
template <typename TC, int ACT, int DACT, int S>   // ACT/DACT -1 = read from EpilogueParams at run time
__device__ __forceinline__ void tc_epilogue_seg(const TcEpilogue& e, uint8_t* scr, int lane, int m0w, int rows_valid,
                                                int n, float* v) {
  using SL = Slot<TC>;
  constexpr int SEGC = S * SL::N;
  const EpilogueParams& ep = e.p;
  const int act = ACT >= 0 ? ACT : ep.act;
  const int dact = DACT >= 0 ? DACT : ep.dact;
  const bool has_dact = DACT >= 0 ? (DACT != VG_ACT_NONE) : (ep.dact_src != nullptr);
  const bool row_ok = lane < rows_valid;
  const bool keep = !(ep.row_mask && row_ok && !ep.row_mask[m0w + lane]);
  uint4 q[S];
  if (ep.bias) {
#pragma unroll
    for (int j = 0; j < SEGC; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(ep.bias + n + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (act != VG_ACT_NONE || ep.preact) {
#pragma unroll
    for (int g = 0; g < S; ++g) {
      float p[SL::N];
      if constexpr (ACT == VG_ACT_GELU) {       // packed f32x2 math: two elements per FFMA2
#pragma unroll
        for (int j = 0; j < SL::N; j += 2) {
          float* x = v + g * SL::N + j;
          float y0, y1, d0, d1;
          gelu_and_grad_pair(x[0], x[1], y0, y1, d0, d1);
          p[j] = ep.preact_is_grad ? d0 : x[0];
          p[j + 1] = ep.preact_is_grad ? d1 : x[1];
          x[0] = y0; x[1] = y1;
        }
      } else {
#pragma unroll
        for (int j = 0; j < SL::N; ++j) {
          float* x = v + g * SL::N + j;
          float y, dy;
          act_and_grad_fast(*x, act, y, dy);
          p[j] = ep.preact_is_grad ? dy : *x;
          *x = y;
        }
      }
      q[g] = SL::pack(p);
    }
    if (ep.preact)
      staged_store<S>(scr, lane, q, reinterpret_cast<uint8_t*>(reinterpret_cast<TC*>(ep.preact) + (int64_t)m0w * ep.ld_preact + n),
                      ep.ld_preact * (int64_t)sizeof(TC), rows_valid);
  }
  if (has_dact) {
    staged_load<S>(scr, lane, q, reinterpret_cast<const uint8_t*>(reinterpret_cast<const TC*>(ep.dact_src) + (int64_t)m0w * ep.ld_dact + n),
                   ep.ld_dact * (int64_t)sizeof(TC), rows_valid);
#pragma unroll
    for (int g = 0; g < S; ++g) {
      float d[SL::N];
      SL::unpack(q[g], d);
#pragma unroll
      for (int j = 0; j < SL::N; ++j) v[g * SL::N + j] *= act_grad_fast(d[j], dact);
    }
  }
  if (!keep && ep.mask_first) {
#pragma unroll
    for (int j = 0; j < SEGC; ++j) v[j] = 0.f;
  }
  if (ep.residual) {
    staged_load<S>(scr, lane, q, reinterpret_cast<const uint8_t*>(reinterpret_cast<const TC*>(ep.residual) + (int64_t)m0w * ep.ld_res + n),
                   ep.ld_res * (int64_t)sizeof(TC), rows_valid);
#pragma unroll
    for (int g = 0; g < S; ++g) {
      float d[SL::N];
      SL::unpack(q[g], d);
#pragma unroll
      for (int j = 0; j < SL::N; ++j) v[g * SL::N + j] += d[j];
    }
  }
  if (!keep && !ep.mask_first) {
#pragma unroll
    for (int j = 0; j < SEGC; ++j) v[j] = 0.f;
  }
  uint8_t* cbase = reinterpret_cast<uint8_t*>(reinterpret_cast<TC*>(ep.C) + (int64_t)m0w * ep.ldc + n);
  const int64_t cpitch = ep.ldc * (int64_t)sizeof(TC);
  if (ep.beta != 0.f) {
    staged_load<S>(scr, lane, q, cbase, cpitch, rows_valid);
#pragma unroll
    for (int g = 0; g < S; ++g) {
      float d[SL::N];
      SL::unpack(q[g], d);
#pragma unroll
      for (int j = 0; j < SL::N; ++j) v[g * SL::N + j] += ep.beta * d[j];
    }
  }
#pragma unroll
  for (int g = 0; g < S; ++g) q[g] = SL::pack(v + g * SL::N);
  staged_store<S>(scr, lane, q, cbase, cpitch, rows_valid);
}

// One tcgen05.ld chunk of CW columns → segments of at most 64 bytes per row.
template <typename TC, int VARIANT, int CW>
__device__ __forceinline__ void tc_epilogue_chunk(const TcEpilogue& e, uint8_t* scr, int lane, int m0w, int M, int n0,
                                                  int N, const uint32_t* r) {
  using SL = Slot<TC>;
  constexpr int SEGC = (CW * (int)sizeof(TC) > 64) ? 64 / (int)sizeof(TC) : CW;   // columns per segment
  constexpr int S = SEGC / SL::N;                                                  // 16-byte slots per row segment
  const int rows_valid = M - m0w < 32 ? M - m0w : 32;
#pragma unroll
  for (int sg = 0; sg < CW / SEGC; ++sg) {
    const int n = n0 + sg * SEGC;
    if (n >= N) break;
    float v[SEGC];
#pragma unroll
    for (int j = 0; j < SEGC; ++j) v[j] = __uint_as_float(r[sg * SEGC + j]);
    if (e.vec_ok && n + SEGC <= N) {        // warp-uniform
      if constexpr (VARIANT == TCV_RED) {
        if constexpr (sizeof(TC) == 4) {
          uint4 q[S];
#pragma unroll
          for (int g = 0; g < S; ++g) q[g] = SL::pack(v + g * SL::N);
          staged_red_f32<S>(scr, lane, q, reinterpret_cast<uint8_t*>(reinterpret_cast<float*>(e.p.C) + (int64_t)m0w * e.p.ldc + n),
                            e.p.ldc * 4, rows_valid);
        }
      } else if constexpr (VARIANT == TCV_PLAIN) tc_epilogue_seg<TC, VG_ACT_NONE, VG_ACT_NONE, S>(e, scr, lane, m0w, rows_valid, n, v);
      else if constexpr (VARIANT == TCV_RELU) tc_epilogue_seg<TC, VG_ACT_RELU, VG_ACT_NONE, S>(e, scr, lane, m0w, rows_valid, n, v);
      else if constexpr (VARIANT == TCV_GELU) tc_epilogue_seg<TC, VG_ACT_GELU, VG_ACT_NONE, S>(e, scr, lane, m0w, rows_valid, n, v);
      else if constexpr (VARIANT == TCV_SILU) tc_epilogue_seg<TC, VG_ACT_SILU, VG_ACT_NONE, S>(e, scr, lane, m0w, rows_valid, n, v);
      else if constexpr (VARIANT == TCV_MULT) tc_epilogue_seg<TC, VG_ACT_NONE, VG_ACT_MULT, S>(e, scr, lane, m0w, rows_valid, n, v);
      else tc_epilogue_seg<TC, -1, -1, S>(e, scr, lane, m0w, rows_valid, n, v);
    } else if (lane < rows_valid) {         // ragged edge / unaligned operands: element-wise, thread-per-row
#pragma unroll
      for (int j = 0; j < SEGC; ++j) {
        if (n + j < N) {
          if constexpr (VARIANT == TCV_RED) atomicAdd(reinterpret_cast<float*>(e.p.C) + (int64_t)(m0w + lane) * e.p.ldc + n + j, v[j]);
          else epilogue_store<TC>(e.p, m0w + lane, n + j, v[j]);
        }
      }
    }
  }
}

// SMs the persistent GEMM grids may occupy (default: all 148).  Data-parallel runs leave a few SMs to the NCCL kernels
// (VG_GEMM_SMS / vg_set_gemm_sm_budget): a persistent CTA whose SM is held by a long-running all-reduce kernel cannot
// start, and with the static tile schedule its tiles wait with it.
int gemm_sm_budget();

// One work unit = (output tile, k-range).  With splits == 1 a unit is a whole tile.
struct TcUnit {
  int m0, n0, kb0, kb1;
};
__device__ __forceinline__ TcUnit tc_unit(int u, int num_m, int num_kb, int splits, int BN, int tile_rows = TBM) {
  const int tile = u / splits;
  const int sp = u - tile * splits;
  TcUnit w;
  w.m0 = (tile % num_m) * tile_rows;
  w.n0 = (tile / num_m) * BN;
  w.kb0 = (int)(((int64_t)sp * num_kb) / splits);
  w.kb1 = (int)(((int64_t)(sp + 1) * num_kb) / splits);
  return w;
}

template <int BN, typename TC, int VARIANT, int CTAS>
__device__ __forceinline__ void tc_epilogue_loop(const TcEpilogue& epi, uint32_t tmem_base, uint64_t* tmem_full,
                                                 uint64_t* tmem_empty, uint8_t* scratch, int M, int N, int num_m,
                                                 int num_kb, int num_units, int warp, int lane) {
  const int cta_rank = CTAS == 2 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / CTAS, unit_step = gridDim.x / CTAS;
  // the MMA issuer (leader CTA) waits for the accumulator to be drained by the epilogue warps of BOTH CTAs
  const uint32_t empty_addr0 = CTAS == 2 ? mapa_shared(smem_u32(&tmem_empty[0]), 0) : 0;
  constexpr int COLS = BN / 4;                    // columns per epilogue warp
  constexpr int CW = COLS < 32 ? COLS : 32;       // columns per tcgen05.ld
  const int e = warp - 4;
  const int wq = e & 3;                           // TMEM lane quadrant this warp may access (== warp % 4)
  const int cg = e >> 2;                          // column group
  uint8_t* scr = scratch + e * TC_SCR_BYTES;
  int acc = 0;
  uint32_t acc_ph = 0;
  for (int u = unit0; u < num_units; u += unit_step) {
    const TcUnit w = tc_unit(u, num_m, num_kb, epi.splits, BN, TBM * CTAS);
    mbar_wait(&tmem_full[acc], acc_ph);
    tc_fence_after();
    const int m0w = w.m0 + cta_rank * TBM + wq * 32;   // first row of this warp's 32-row slab
    const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN + cg * COLS);
#pragma unroll
    for (int c = 0; c < COLS / CW; ++c) {
      uint32_t r[CW];
      if constexpr (CW == 32) tmem_ld_32x32b_x32(t_row + (uint32_t)(c * CW), r);
      else tmem_ld_32x32b_x16(t_row + (uint32_t)(c * CW), r);
      tmem_ld_wait();
      const int nc = w.n0 + cg * COLS + c * CW;
      if (m0w < M && nc < N) tc_epilogue_chunk<TC, VARIANT, CW>(epi, scr, lane, m0w, M, nc, N, r);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (CTAS == 2) mbar_arrive_cluster(empty_addr0 + (uint32_t)(acc * 8));
      else mbar_arrive(&tmem_empty[acc]);
    }
    acc ^= 1;
    if (acc == 0) acc_ph ^= 1u;
  }
}

// VMASK: bit i set → TcVariant i has a specialised epilogue in this instantiation (others use the generic one).
template <int BN, bool A_MN, bool B_MN, typename TC, unsigned VMASK, int CTAS>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, TcEpilogue epi) {
  using Cfg = TcCfg<BN, CTAS>;
  constexpr int kStages = Cfg::kStages;
  constexpr int BNL = BN / CTAS;                 // B columns held (and loaded) by this CTA
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);   // 1024 B alignment for SW128

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = CTAS == 2 ? (int)cluster_ctarank() : 0;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], CTAS * (TC_EPI_THREADS / 32));     // one arrival per epilogue warp (of each CTA)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTAS == 2) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + TBM * CTAS - 1) / (TBM * CTAS);
  const int num_n = (N + BN - 1) / BN;
  const int num_kb = (K + TBK - 1) / TBK;
  const int num_units = num_m * num_n * epi.splits;
  const int unit0 = blockIdx.x / CTAS, unit_step = gridDim.x / CTAS;

  // Programmatic dependent launch (vg_set_pdl_mode): everything above overlapped the tail of the kernel in front; the
  // kernel behind may start its own prologue now.  Every role waits for the predecessor's memory before it touches
  // global memory — except the producer's prefetch of a STATIC B operand (weights) into the first ring stages.
  // Both instructions are no-ops in a launch without the attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int pdl_prefetched = 0;
  if (!(warp == 0 && (epi.pdl & 2) && CTAS == 1 && VG_GEMM_UNIFORM_ISSUE != 0)) asm volatile("griddepcontrol.wait;" ::: "memory");

  // VG_GEMM_UNIFORM_ISSUE (compile-time, default 1): the producer / MMA-issuer warps walk their loops as whole warps and
  // one elect.sync-elected lane issues.  Inside a `lane == 0` branch the compiler wraps every TMA / tcgen05.mma in a
  // per-thread ELECT / R2UR / BRA.U.ANY loop (~80 cycles per instruction: measured in the attention kernels, where the
  // same change made the MMAs issue back to back — profiles/r01_attention_v2.md); with 128-wide tiles (64 tensor cycles
  // per MMA) that is issue-bound.  Measured on B200 (profiles/r02_variants.md): fwd qkv 0.88 -> 0.99x cuBLAS, fwd ffn1
  // 0.99 -> 1.07x, dgrad ffn2 1.04 -> 1.13x, nothing slower; -DVG_GEMM_UNIFORM_ISSUE=0 restores the lane-0 branches.
  constexpr bool kUniformIssue = VG_GEMM_UNIFORM_ISSUE != 0;
  if (warp == 0 && (kUniformIssue || lane == 0)) {
    // ===================== TMA producer (every CTA loads its own A rows and its own slice of B) =====================
    int s = 0;
    uint32_t ph = 0;
    if constexpr (CTAS == 1) {
      if ((epi.pdl & 2) && kUniformIssue) {
        // static-weight prefetch: B tiles of the first unit's first k-blocks, before the predecessor is done
        if (unit0 < num_units) {
          const TcUnit w = tc_unit(unit0, num_m, num_kb, epi.splits, BN, TBM);
          const int npf = (w.kb1 - w.kb0) < kStages ? (w.kb1 - w.kb0) : kStages;
          if (elect_one()) {
            for (int i = 0; i < npf; ++i) {
              uint8_t* b_dst = smem + i * Cfg::kStageBytes + A_TILE_BYTES;
              const int k0 = (w.kb0 + i) * TBK;
              mbar_arrive_expect_tx(&full_bar[i], Cfg::kStageBytes);
              if constexpr (!B_MN) {
                tma_load_2d(b_dst, &tmB, &full_bar[i], k0, w.n0);
              } else {
#pragma unroll
                for (int j = 0; j < BNL / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[i], w.n0 + 64 * j, k0);
              }
            }
          }
          __syncwarp();
          pdl_prefetched = npf;
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
      }
    }
    for (int u = unit0; u < num_units; u += unit_step) {
      const TcUnit w = tc_unit(u, num_m, num_kb, epi.splits, BN, TBM * CTAS);
      const int m0 = w.m0 + cta_rank * TBM;
      const int n0 = w.n0 + cta_rank * BNL;
      for (int kb = w.kb0; kb < w.kb1; ++kb) {
        const bool b_done = pdl_prefetched > 0;      // this stage's B tile and expect_tx were issued before the wait
        if (b_done) --pdl_prefetched;
        else mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* a_dst = smem + s * Cfg::kStageBytes;
        uint8_t* b_dst = a_dst + A_TILE_BYTES;
        const int k0 = kb * TBK;
        if (b_done) {
          if (elect_one()) {
            if constexpr (!A_MN) {
              tma_load_2d(a_dst, &tmA, &full_bar[s], k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < TBM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
            }
          }
        } else
        if (!kUniformIssue || elect_one()) {
        if constexpr (CTAS == 2) {
          // both CTAs' bytes are counted on the LEADER's barrier (the MMA issuer lives there)
          const uint32_t bar = mapa_shared(smem_u32(&full_bar[s]), 0);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
          if constexpr (!A_MN) {
            tma_load_2d_pair(a_dst, &tmA, bar, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < TBM / 64; ++j) tma_load_2d_pair(a_dst + j * 8192, &tmA, bar, m0 + 64 * j, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_pair(b_dst, &tmB, bar, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BNL / 64; ++j) tma_load_2d_pair(b_dst + j * 8192, &tmB, bar, n0 + 64 * j, k0);
          }
        } else {
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, &tmA, &full_bar[s], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < TBM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d(b_dst, &tmB, &full_bar[s], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BNL / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
          }
        }
        }
        if (kUniformIssue) __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && (kUniformIssue || lane == 0) && cta_rank == 0) {
    // ===================== MMA issuer (the leader CTA issues for the pair) =====================
    constexpr uint32_t idesc = make_idesc_bf16(TBM * CTAS, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int u = unit0; u < num_units; u += unit_step) {
      const TcUnit w = tc_unit(u, num_m, num_kb, epi.splits, BN, TBM * CTAS);
      if constexpr (CTAS == 2) mbar_wait_cluster(&tmem_empty[acc], acc_ph ^ 1u);
      else mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = w.kb0; kb < w.kb1; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + A_TILE_BYTES;
        if (!kUniformIssue || elect_one()) {
#pragma unroll
        for (int k = 0; k < TBK / 16; ++k) {
          // K-major  : rows at 128 B pitch, 8-row groups every 1024 B (SBO); +32 B per 16-wide k step.
          // MN-major : 64-element MN chunks every 8192 B (LBO), 8 k-rows per 1024 B (SBO); +2048 B per k step.
          const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          if constexpr (CTAS == 2) umma_f16_ss_pair(d_tmem, da, db, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
          else umma_f16_ss(d_tmem, da, db, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
        }
        // frees the smem slot (in both CTAs) once these MMAs retire
        if constexpr (CTAS == 2) umma_commit_pair(&empty_bar[s]); else umma_commit(&empty_bar[s]);
        if (kUniformIssue && kb + 1 == w.kb1) {      // same lane as the MMAs it tracks
          if constexpr (CTAS == 2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
        }
        }
        if (kUniformIssue) __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      // accumulator complete → epilogue warps (of both CTAs)
      if (!kUniformIssue || (w.kb0 >= w.kb1 && elect_one())) {
        if constexpr (CTAS == 2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
      }
      if (kUniformIssue) __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
#define VG_TC_EPI(V)                                                                                               \
  tc_epilogue_loop<BN, TC, V, CTAS>(epi, tmem_base, tmem_full, tmem_empty, smem + Cfg::kScratchOff, M, N, num_m,    \
                                    num_kb, num_units, warp, lane)
    const int var = epi.variant;
    bool done = false;
    if constexpr ((VMASK >> TCV_PLAIN) & 1u) { if (!done && var == TCV_PLAIN) { VG_TC_EPI(TCV_PLAIN); done = true; } }
    if constexpr ((VMASK >> TCV_RELU) & 1u) { if (!done && var == TCV_RELU) { VG_TC_EPI(TCV_RELU); done = true; } }
    if constexpr ((VMASK >> TCV_GELU) & 1u) { if (!done && var == TCV_GELU) { VG_TC_EPI(TCV_GELU); done = true; } }
    if constexpr ((VMASK >> TCV_SILU) & 1u) { if (!done && var == TCV_SILU) { VG_TC_EPI(TCV_SILU); done = true; } }
    if constexpr ((VMASK >> TCV_MULT) & 1u) { if (!done && var == TCV_MULT) { VG_TC_EPI(TCV_MULT); done = true; } }
    if constexpr (((VMASK >> TCV_RED) & 1u) && sizeof(TC) == 4) {
      if (!done && var == TCV_RED) { VG_TC_EPI(TCV_RED); done = true; }
    }
    if (!done) VG_TC_EPI(TCV_GENERIC);
#undef VG_TC_EPI
  }

  tc_fence_before();
  if constexpr (CTAS == 2) {
    cluster_sync_all();          // the peer may still multicast commits into / read from this CTA's shared memory
    if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  } else {
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, typename TC, unsigned VMASK, int CTAS>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a, TcEpilogue epi,
                     cudaStream_t st) {
  using Cfg = TcCfg<BN, CTAS>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, TC, VMASK, CTAS>;
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  // a variant this instantiation does not specialise falls back to the generic epilogue (TCV_RED has no generic
  // equivalent: the host only selects it where VMASK carries it)
  const int64_t units = ceil_div(a->M, TBM * CTAS) * ceil_div(a->N, BN) * epi.splits;
  const int max_groups = gemm_sm_budget() / CTAS;
  const int grid = (int)(units < max_groups ? units : max_groups) * CTAS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CTAS == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CTAS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  epi.pdl = g_pdl_mode;
  if (epi.pdl & 1) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  VG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, (int)a->M, (int)a->N, (int)a->K, epi));
  return 0;
}

// bn: 64 / 128 / 256 (one CTA per tile) or 512 = a CTA pair on 256 x 256 tiles
template <bool A_MN, bool B_MN, unsigned VMASK, typename TC>
static int launch_tc_layout(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                            const TcEpilogue& epi, cudaStream_t st) {
  if (bn == 512) return launch_tc<256, A_MN, B_MN, TC, VMASK, 2>(tmA, tmB, a, epi, st);
  if (bn == 256) return launch_tc<256, A_MN, B_MN, TC, VMASK, 1>(tmA, tmB, a, epi, st);
  if (bn == 128) return launch_tc<128, A_MN, B_MN, TC, VMASK, 1>(tmA, tmB, a, epi, st);
  return launch_tc<64, A_MN, B_MN, TC, VMASK, 1>(tmA, tmB, a, epi, st);
}

// specialised-variant masks per operand layout (what the model launches: forward = K/K with activations,
// dgrad = K/MN with the stored derivative, wgrad = MN/MN with split-K)
constexpr unsigned TCM_KK = (1u << TCV_PLAIN) | (1u << TCV_RELU) | (1u << TCV_GELU) | (1u << TCV_SILU) | (1u << TCV_RED);
constexpr unsigned TCM_KMN = (1u << TCV_PLAIN) | (1u << TCV_MULT) | (1u << TCV_RED);
constexpr unsigned TCM_MNK = (1u << TCV_RED);
constexpr unsigned TCM_MNMN = (1u << TCV_PLAIN) | (1u << TCV_RED);

int gemm_tc_launch_kk_bf16(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);
int gemm_tc_launch_kk_f32(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);
int gemm_tc_launch_kmn_bf16(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);
int gemm_tc_launch_kmn_f32(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);
int gemm_tc_launch_mnk_bf16(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);
int gemm_tc_launch_mnk_f32(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);
int gemm_tc_launch_mnmn_bf16(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);
int gemm_tc_launch_mnmn_f32(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st);

}  // namespace vg
