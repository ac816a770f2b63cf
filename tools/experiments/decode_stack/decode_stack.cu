// decode_stack.cu — EXPERIMENT: the 16 transformer layers of one cached LVTR.step as ONE cooperative kernel (batch ≤ 8).
// Design, shared-memory budget and the race analysis of the residual stream: ../decode_stack_design.md.  Built by the
// Makefile next to this file into libds.so and checked / timed by run.py against a plain torch fp32 reference.
// Written at the end of round 1 without access to a GPU: it compiles for sm_100a, it has NOT run yet.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ds {

constexpr int DM = 1024, NH = 16, HD = 64, FF = 4096, MAXL = 16, MAXB = 8;
constexpr int THREADS = 256, WARPS = 8;
constexpr int OWNERS = 128;                 // CTAs that own GEMM slabs
constexpr int NA = 3 * DM / OWNERS;         // 24 qkv features per owner
constexpr int NC = DM / OWNERS;             // 8 out-projection features per owner
constexpr int NF = FF / OWNERS;             // 32 hidden features per owner
constexpr int XS = DM + 8;                  // padded row stride (elements) of K = 1024 tiles
constexpr int GS = NF + 8;                  // row stride of the hidden tile g [8][32]
constexpr int D2S = NF + 4;                 // row stride of the FFN2 slab [1024 n][32 k]

constexpr int SZ_A = NA * XS * 2, SZ_C = NC * XS * 2, SZ_D1 = NF * XS * 2, SZ_D2 = DM * D2S * 2;
constexpr int OFF_A = 0, OFF_C = OFF_A + SZ_A, OFF_D1 = OFF_C + SZ_C, OFF_D2 = OFF_D1 + SZ_D1;
constexpr int OFF_XA = OFF_D2 + SZ_D2, SZ_XA = MAXB * XS * 2;
constexpr int OFF_RED = OFF_XA + SZ_XA, SZ_RED = WARPS * MAXB * 32 * 4;
constexpr int OFF_G = OFF_RED + SZ_RED, SZ_G = MAXB * GS * 2;
constexpr int OFF_BAR = OFF_G + SZ_G;
constexpr int SMEM_BYTES = OFF_BAR + 64;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(SZ_A % 16 == 0 && SZ_C % 16 == 0 && SZ_D1 % 16 == 0 && SZ_D2 % 16 == 0, "bulk copies are 16-byte granular");
constexpr int MAX_KEYS = (SZ_XA + SZ_RED - 4096) / 4;       // score buffer of the attention phase (aliases xa + red)

struct Layer {
  const __nv_bfloat16 *slabA, *slabC, *slabD1, *slabD2;     // host-packed per-owner images: [OWNERS][SZ_x bytes]
  const float *n1, *n3, *b1, *b2;
  __nv_bfloat16 *kc, *vc;                                   // KV cache of this layer [B][NH][Tmax][HD]
};
struct Params {
  Layer layer[MAXL];
  int L, B, pos, Tmax;
  float eps, scale;
  const float* slopes;
  float* hres[2];          // fp32 residual stream [B][DM], double-buffered by layer parity; hres[0] = input
  float* facc[2];          // FFN2 partial-sum accumulators [B][DM] (zeroed by the host)
  float* qbuf;             // [B][DM] fp32, pre-scaled queries
  __nv_bfloat16* abuf;     // [B][DM] attention output
  unsigned* barrier;       // one word, zeroed by the host per launch
  float* out;              // [B][DM] stack output (before the final norm)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(b)), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  const uint32_t z = 0u;       // rows 8..15 of the A tile are zero (batch ≤ 8)
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(z), "r"(a2), "r"(z), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// device-wide barrier #idx (monotonic counter, zeroed by the host before the launch)
__device__ __forceinline__ void grid_barrier(unsigned* counter, int& idx) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const unsigned target = (unsigned)(idx + 1) * gridDim.x;
    while (ld_acquire(counter) < target) {
    }
  }
  ++idx;
  __syncthreads();
}
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// out[r][n] (r < 8, n < 8·NT) = Σ_k xa[r][k] · W[n][k] over K = 1024: the 8 warps split K, partials meet in `red`
template <int NT>
__device__ __forceinline__ void gemm_k1024(const __nv_bfloat16* xa, const __nv_bfloat16* W, float* red, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float c[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f;
#pragma unroll 2
  for (int ks = 0; ks < DM / 16 / WARPS; ++ks) {
    const int k0 = (warp * (DM / 16 / WARPS) + ks) * 16;
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(xa + g * XS + k0 + 2 * t);
    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(xa + g * XS + k0 + 8 + 2 * t);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const __nv_bfloat16* wr = W + (n * 8 + g) * XS + k0 + 2 * t;
      mma16816(c[n], a0, a2, *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
    }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) {            // red[warp][row g][col n·8 + 2t, +1]
    red[(warp * MAXB + g) * 32 + n * 8 + 2 * t] = c[n][0];
    red[(warp * MAXB + g) * 32 + n * 8 + 2 * t + 1] = c[n][1];
  }
}
__device__ __forceinline__ float red_sum(const float* red, int row, int col) {
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) s += red[(w * MAXB + row) * 32 + col];
  return s;
}

// x rows → RMSNorm → bf16 tile xa.  x = src0 (+ src1 + bias) read around L1 (other CTAs wrote them); thread = 4 columns.
// If dst (owner copy) is given, the columns [own0, own0 + 8) of x are also written there and zeroed in `zero`.
__device__ __forceinline__ void load_norm(const Params& p, const float* src0, const float* src1, const float* bias,
                                          const float* nscale, __nv_bfloat16* xa, float* sred, float* dst, float* zero,
                                          int own0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float4 x[MAXB];
  float ss[MAXB];
  const float4 bz = bias ? __ldg(reinterpret_cast<const float4*>(bias) + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < MAXB; ++r) {
    x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < p.B) {
      x[r] = __ldcg(reinterpret_cast<const float4*>(src0 + (size_t)r * DM) + tid);
      if (src1) {
        const float4 f = __ldcg(reinterpret_cast<const float4*>(src1 + (size_t)r * DM) + tid);
        x[r].x += f.x + bz.x; x[r].y += f.y + bz.y; x[r].z += f.z + bz.z; x[r].w += f.w + bz.w;
      }
    }
    float s = x[r].x * x[r].x + x[r].y * x[r].y + x[r].z * x[r].z + x[r].w * x[r].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    ss[r] = s;
  }
  if (lane == 0)
#pragma unroll
    for (int r = 0; r < MAXB; ++r) sred[warp * MAXB + r] = ss[r];
  __syncthreads();
  const float4 sc = __ldg(reinterpret_cast<const float4*>(nscale) + tid);
#pragma unroll
  for (int r = 0; r < MAXB; ++r) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) tot += sred[w * MAXB + r];
    const float rstd = rsqrtf(tot * (1.f / DM) + p.eps);
    __nv_bfloat162 lo = __floats2bfloat162_rn(x[r].x * rstd * sc.x, x[r].y * rstd * sc.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(x[r].z * rstd * sc.z, x[r].w * rstd * sc.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(xa + r * XS + tid * 4) = pk;
    if (dst && r < p.B && tid * 4 >= own0 && tid * 4 < own0 + NC) {
      *(reinterpret_cast<float4*>(dst + (size_t)r * DM) + tid) = x[r];
      *(reinterpret_cast<float4*>(zero + (size_t)r * DM) + tid) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 1) decode_stack_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem + OFF_A);
  __nv_bfloat16* sC = reinterpret_cast<__nv_bfloat16*>(smem + OFF_C);
  __nv_bfloat16* sD1 = reinterpret_cast<__nv_bfloat16*>(smem + OFF_D1);
  __nv_bfloat16* sD2 = reinterpret_cast<__nv_bfloat16*>(smem + OFF_D2);
  __nv_bfloat16* xa = reinterpret_cast<__nv_bfloat16*>(smem + OFF_XA);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  __nv_bfloat16* sg = reinterpret_cast<__nv_bfloat16*>(smem + OFF_G);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);      // [0] A, [1] C, [2] D1, [3] D2
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const bool owner = c < OWNERS;
  int bar_idx = 0;

  auto prefetch = [&](int which, int l) {      // thread 0 only; the region must have been fully consumed
    const Layer& ly = p.layer[l];
    const __nv_bfloat16* src = which == 0 ? ly.slabA : which == 1 ? ly.slabC : which == 2 ? ly.slabD1 : ly.slabD2;
    const int sz = which == 0 ? SZ_A : which == 1 ? SZ_C : which == 2 ? SZ_D1 : SZ_D2;
    void* dst = which == 0 ? (void*)sA : which == 1 ? (void*)sC : which == 2 ? (void*)sD1 : (void*)sD2;
    mbar_expect(&bars[which], (uint32_t)sz);
    bulk_g2s(dst, reinterpret_cast<const uint8_t*>(src) + (size_t)c * sz, (uint32_t)sz, &bars[which]);
  };

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (owner)
      for (int i = 0; i < 4; ++i) prefetch(i, 0);
  }
  for (int i = tid; i < SZ_XA / 4; i += THREADS) reinterpret_cast<uint32_t*>(xa)[i] = 0u;
  for (int i = tid; i < SZ_G / 4; i += THREADS) reinterpret_cast<uint32_t*>(sg)[i] = 0u;
  __syncthreads();

  const float scale2 = p.scale * 1.4426950408889634f;
  for (int l = 0; l < p.L; ++l) {
    const Layer& ly = p.layer[l];
    const int par = l & 1;
    // ------------------------------------------------------------------ A: RMSNorm1 + QKV slab
    if (owner) {
      load_norm(p, p.hres[par], l > 0 ? p.facc[par] : nullptr, l > 0 ? p.layer[l - 1].b2 : nullptr, ly.n1, xa, red,
                p.hres[par ^ 1], p.facc[par ^ 1], c * NC);
      mbar_wait(&bars[0], par);
      gemm_k1024<NA / 8>(xa, sA, red, warp, lane);
      __syncthreads();                                        // slab A consumed, partials in `red`
      if (tid == 0 && l + 1 < p.L) prefetch(0, l + 1);
      if (tid < MAXB * NA) {
        const int r = tid / NA, n = tid % NA;
        if (r < p.B) {
          const float v = red_sum(red, r, n);
          const int f = c * NA + n, which = f / DM, col = f % DM, h = col / HD, dd = col % HD;
          if (which == 0) {
            p.qbuf[(size_t)r * DM + col] = v * scale2;
          } else {
            __nv_bfloat16* cache = which == 1 ? ly.kc : ly.vc;
            cache[(((size_t)r * NH + h) * p.Tmax + p.pos) * HD + dd] = __float2bfloat16_rn(v);
          }
        }
      }
    }
    grid_barrier(p.barrier, bar_idx);
    // ------------------------------------------------------------------ B: attention, one (b, head) item at a time
    {
      float* sq = reinterpret_cast<float*>(smem + OFF_XA);          // [64] pre-scaled query
      float* so = sq + 64;                                          // [WARPS][64] partial outputs
      float* sst = so + WARPS * 64;                                 // [2·WARPS] block max / sum
      float* sc = sst + 64;                                         // [Tk] scores → probabilities
      const int Tk = p.pos + 1;
      for (int item = c; item < p.B * NH; item += gridDim.x) {
        const int b = item / NH, h = item % NH;
        const float slope2 = (p.slopes ? p.slopes[h] : 0.f) * 1.4426950408889634f;
        const __nv_bfloat16* K = ly.kc + ((size_t)b * NH + h) * p.Tmax * HD;
        const __nv_bfloat16* V = ly.vc + ((size_t)b * NH + h) * p.Tmax * HD;
        __syncthreads();                                            // previous item's scratch is free
        if (tid < HD) sq[tid] = __ldcg(p.qbuf + (size_t)b * DM + h * HD + tid);
        __syncthreads();
        float mx = -INFINITY;
        for (int j = tid; j < Tk; j += THREADS) {
          const uint4* kr = reinterpret_cast<const uint4*>(K + (size_t)j * HD);
          float acc = 0.f;
#pragma unroll
          for (int q8 = 0; q8 < HD / 8; ++q8) {
            const uint4 u = __ldcg(kr + q8);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h2[e]);
              acc = fmaf(f.x, sq[q8 * 8 + 2 * e], acc);
              acc = fmaf(f.y, sq[q8 * 8 + 2 * e + 1], acc);
            }
          }
          const float s = acc - slope2 * (float)(p.pos - j);
          sc[j] = s;
          mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) sst[warp] = mx;
        __syncthreads();
        mx = sst[0];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) mx = fmaxf(mx, sst[w]);
        float sum = 0.f;
        for (int j = tid; j < Tk; j += THREADS) {
          const float e = exp2f(sc[j] - mx);
          sc[j] = e;
          sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) sst[WARPS + warp] = sum;
        __syncthreads();                                            // probabilities and partial sums visible
        sum = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) sum += sst[WARPS + w];
        float o0 = 0.f, o1 = 0.f;
        for (int j = warp; j < Tk; j += WARPS) {
          const uint32_t u = __ldcg(reinterpret_cast<const uint32_t*>(V + (size_t)j * HD) + lane);
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
          const float pj = sc[j];
          o0 = fmaf(pj, f.x, o0);
          o1 = fmaf(pj, f.y, o1);
        }
        so[warp * 64 + 2 * lane] = o0;
        so[warp * 64 + 2 * lane + 1] = o1;
        __syncthreads();
        if (tid < HD) {
          float o = 0.f;
#pragma unroll
          for (int w = 0; w < WARPS; ++w) o += so[w * 64 + tid];
          p.abuf[(size_t)b * DM + h * HD + tid] = __float2bfloat16_rn(o / sum);
        }
      }
      __syncthreads();
    }
    grid_barrier(p.barrier, bar_idx);
    // ------------------------------------------------------------------ C: out-projection slab, owner adds into the residual
    if (owner) {
      for (int i = tid; i < MAXB * (DM / 8); i += THREADS) {       // abuf [B][1024] bf16 → xa (16-byte pieces)
        const int r = i / (DM / 8), q8 = i % (DM / 8);
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (r < p.B) u = __ldcg(reinterpret_cast<const uint4*>(p.abuf + (size_t)r * DM) + q8);
        *reinterpret_cast<uint4*>(xa + r * XS + q8 * 8) = u;
      }
      __syncthreads();
      mbar_wait(&bars[1], par);
      gemm_k1024<NC / 8>(xa, sC, red, warp, lane);
      __syncthreads();
      if (tid == 0 && l + 1 < p.L) prefetch(1, l + 1);
      if (tid < MAXB * NC) {
        const int r = tid / NC, n = tid % NC;
        if (r < p.B) {
          float* dst = p.hres[par ^ 1] + (size_t)r * DM + c * NC + n;
          *dst = __ldcg(dst) + red_sum(red, r, n);
        }
      }
    }
    grid_barrier(p.barrier, bar_idx);
    // ------------------------------------------------------------------ D: RMSNorm3 + FFN1 slab + GELU + FFN2 partial sums
    if (owner) {
      load_norm(p, p.hres[par ^ 1], nullptr, nullptr, ly.n3, xa, red, nullptr, nullptr, 0);
      mbar_wait(&bars[2], par);
      gemm_k1024<NF / 8>(xa, sD1, red, warp, lane);
      __syncthreads();
      if (tid == 0 && l + 1 < p.L) prefetch(2, l + 1);
      {
        const int r = tid / NF, n = tid % NF;                       // 256 threads = 8 rows x 32 hidden features
        const float v = red_sum(red, r, n) + __ldg(ly.b1 + c * NF + n);
        sg[r * GS + n] = __float2bfloat16_rn(r < p.B ? gelu_exact(v) : 0.f);
      }
      __syncthreads();
      mbar_wait(&bars[3], par);
      {
        const int g = lane >> 2, t = lane & 3;
        const uint32_t a0k0 = *reinterpret_cast<const uint32_t*>(sg + g * GS + 2 * t);
        const uint32_t a2k0 = *reinterpret_cast<const uint32_t*>(sg + g * GS + 8 + 2 * t);
        const uint32_t a0k1 = *reinterpret_cast<const uint32_t*>(sg + g * GS + 16 + 2 * t);
        const uint32_t a2k1 = *reinterpret_cast<const uint32_t*>(sg + g * GS + 24 + 2 * t);
        float* acc = p.facc[par ^ 1];
#pragma unroll 4
        for (int nt = 0; nt < DM / 8 / WARPS; ++nt) {
          const int n0 = (warp * (DM / 8 / WARPS) + nt) * 8;
          const __nv_bfloat16* wr = sD2 + (n0 + g) * D2S + 2 * t;
          float cc[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(cc, a0k0, a2k0, *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
          mma16816(cc, a0k1, a2k1, *reinterpret_cast<const uint32_t*>(wr + 16), *reinterpret_cast<const uint32_t*>(wr + 24));
          if (g < p.B) {
            atomicAdd(acc + (size_t)g * DM + n0 + 2 * t, cc[0]);
            atomicAdd(acc + (size_t)g * DM + n0 + 2 * t + 1, cc[1]);
          }
        }
      }
      __syncthreads();
      if (tid == 0 && l + 1 < p.L) prefetch(3, l + 1);
    }
    grid_barrier(p.barrier, bar_idx);
  }
  // ---- stack output: residual + last FFN accumulator + its bias (owner columns)
  if (owner && tid < MAXB * NC) {
    const int r = tid / NC, n = tid % NC, col = c * NC + n, par = p.L & 1;
    if (r < p.B)
      p.out[(size_t)r * DM + col] = __ldcg(p.hres[par] + (size_t)r * DM + col) + __ldcg(p.facc[par] + (size_t)r * DM + col) +
                                    __ldg(p.layer[p.L - 1].b2 + col);
  }
}

}  // namespace ds

extern "C" int ds_smem_bytes() { return ds::SMEM_BYTES; }
extern "C" int ds_max_keys() { return ds::MAX_KEYS; }
extern "C" int ds_slab_bytes(int which) {
  return which == 0 ? ds::SZ_A : which == 1 ? ds::SZ_C : which == 2 ? ds::SZ_D1 : ds::SZ_D2;
}
// params: host copy of ds::Params (built by run.py through ctypes).  Returns the grid size, or < 0 on error.
extern "C" int ds_launch(const void* params, void* stream) {
  ds::Params p = *reinterpret_cast<const ds::Params*>(params);
  if (p.B < 1 || p.B > ds::MAXB || p.L < 1 || p.L > ds::MAXL || p.pos + 1 > ds::MAX_KEYS || p.pos >= p.Tmax) return -3;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms < ds::OWNERS) return -4;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(ds::decode_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ds::SMEM_BYTES) !=
        cudaSuccess)
      return -5;
    set = true;
  }
  if (cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), (cudaStream_t)stream) != cudaSuccess) return -6;
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)ds::decode_stack_kernel, dim3(sms), dim3(ds::THREADS), args,
                                              ds::SMEM_BYTES, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    fprintf(stderr, "ds_launch: %s\n", cudaGetErrorString(e));
    return -7;
  }
  return sms;
}
