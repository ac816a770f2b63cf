// common.cuh — shared helpers for libvgslm (sm_100a).  Internal; the public ABI is include/vgslm.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/vgslm.h"

namespace vg {

void set_error(const char* fmt, ...);
// vg_set_pdl_mode (include/vgslm.h): bit 0 = launch vg_gemm / vg_rmsnorm_fwd with the programmatic-dependent-launch
// attribute, bit 1 = the B operand of vg_gemm is a static weight (prefetched before griddepcontrol.wait)
extern int g_pdl_mode;

#define VG_REQUIRE(cond, code, ...)                     \
  do {                                                  \
    if (!(cond)) {                                      \
      ::vg::set_error(__VA_ARGS__);                     \
      return (code);                                    \
    }                                                   \
  } while (0)

#define VG_LAUNCH_CHECK(name)                                                  \
  do {                                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess) {                                                   \
      ::vg::set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));  \
      return (int)_e;                                                          \
    }                                                                          \
  } while (0)

#define VG_CUDA(call)                                                             \
  do {                                                                            \
    cudaError_t _e = (call);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::vg::set_error("%s failed: %s", #call, cudaGetErrorString(_e));            \
      return (int)_e;                                                             \
    }                                                                             \
  } while (0)

static inline bool valid_dtype(int d) { return d == VG_F32 || d == VG_BF16; }
static inline size_t dtype_size(int d) { return d == VG_F32 ? 4 : 2; }
static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// ---- device helpers ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact (erf) GELU, torch.nn.GELU() default — modules/activations.py:11
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// Fast GELU for the bf16 tensor-core epilogues: erf by Abramowitz–Stegun 7.1.26 (|error| <= 1.5e-7, far below
// bf16 resolution) — one MUFU.RCP, one MUFU.EX2 and a 5-term Horner instead of erff's ~40 instructions.  The
// exponential e^{-x^2/2} it needs is also the Gaussian density of GELU', so the derivative costs nothing extra.
__device__ __forceinline__ void gelu_fast_parts(float x, float& cdf, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-z * z * 1.4426950408889634f));
  e = ex;                                           // e^{-x^2/2}
  const float erf_abs = 1.0f - poly * ex;           // erf(|x|/sqrt(2))
  cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, e;
  gelu_fast_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float cdf, e;
  gelu_fast_parts(x, cdf, e);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}
// ---- packed f32x2 arithmetic (Blackwell FFMA2/FMUL2/FADD2: two fp32 lanes per instruction) -------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2_t splat2(float a) { return pack2(a, a); }
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU and its derivative for two elements at once (same Abramowitz–Stegun erf as gelu_fast_parts): ~12 issue slots
// per element instead of ~21 — the bf16 GEMM epilogue is issue-bound once the tensor pipe is fed.
__device__ __forceinline__ void gelu_and_grad_pair(float x0, float x1, float& y0, float& y1, float& d0, float& d1) {
  const f32x2_t x = pack2(x0, x1);
  const f32x2_t ax = pack2(fabsf(x0), fabsf(x1));
  const f32x2_t tin = fma2(ax, splat2(0.3275911f * 0.70710678118654752f), splat2(1.0f));
  float t0, t1;
  unpack2(tin, t0, t1);
  const f32x2_t t = pack2(rcp_approx(t0), rcp_approx(t1));
  f32x2_t poly = fma2(splat2(-1.061405429f), t, splat2(1.453152027f));       // negated polynomial: -(a5 t + a4 ...)
  poly = fma2(poly, t, splat2(-1.421413741f));
  poly = fma2(poly, t, splat2(0.284496736f));
  poly = fma2(poly, t, splat2(-0.254829592f));
  poly = mul2(poly, t);
  const f32x2_t arg = mul2(mul2(x, x), splat2(-0.5f * 1.4426950408889634f));
  float a0, a1;
  unpack2(arg, a0, a1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const f32x2_t e = pack2(e0, e1);                                            // e^{-x^2/2}
  float r0, r1;
  unpack2(fma2(poly, e, splat2(1.0f)), r0, r1);                               // erf(|x|/sqrt 2)
  const f32x2_t cdf = fma2(pack2(copysignf(r0, x0), copysignf(r1, x1)), splat2(0.5f), splat2(0.5f));
  unpack2(mul2(x, cdf), y0, y1);
  unpack2(fma2(mul2(x, splat2(0.39894228040143268f)), e, cdf), d0, d1);
}
// SiLU (conv stack): x·σ(x);  d/dx = σ(x)·(1 + x·(1 − σ(x)))
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float silu_grad_f(float x) {
  const float s = sigmoid_f(x);
  return s * fmaf(x, 1.0f - s, 1.0f);
}
// activation and derivative from one evaluation (the forward epilogue can store either)
__device__ __forceinline__ void act_and_grad_fast(float v, int act, float& y, float& dy) {
  if (act == VG_ACT_RELU) { y = v > 0.f ? v : 0.f; dy = v > 0.f ? 1.f : 0.f; return; }
  if (act == VG_ACT_GELU) {
    float cdf, e;
    gelu_fast_parts(v, cdf, e);
    y = v * cdf;
    dy = fmaf(v * 0.39894228040143268f, e, cdf);
    return;
  }
  if (act == VG_ACT_SILU) {
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-v * 1.4426950408889634f));
    const float s = __frcp_rn(1.0f + ex);
    y = v * s;
    dy = s * fmaf(v, 1.0f - s, 1.0f);
    return;
  }
  y = v; dy = 1.f;
}
__device__ __forceinline__ float apply_act_fast(float v, int act) {
  float y, dy;
  act_and_grad_fast(v, act, y, dy);
  return y;
}
__device__ __forceinline__ float act_grad_fast(float pre, int act) {
  if (act == VG_ACT_MULT) return pre;       // `pre` already holds the derivative
  float y, dy;
  act_and_grad_fast(pre, act, y, dy);
  return dy;
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == VG_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == VG_ACT_GELU) return gelu_f(v);
  if (act == VG_ACT_SILU) return silu_f(v);
  return v;
}
__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == VG_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  if (act == VG_ACT_GELU) return gelu_grad_f(pre);
  if (act == VG_ACT_SILU) return silu_grad_f(pre);
  if (act == VG_ACT_MULT) return pre;
  return 1.f;
}

// 8-wide vector of activation values, loaded/stored as one 16 B (bf16) or two 16 B (f32) accesses.
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
};

}  // namespace vg
