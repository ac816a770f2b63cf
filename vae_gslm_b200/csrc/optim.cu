// optim.cu — AdamW over a flat parameter arena, fused with the bf16 shadow-weight refresh.
// Reference semantics: training_lib/optimizer.py:18-25,110-130 → torch.optim.AdamW (decoupled decay,
// bias-corrected moments).  HBM-bound: reads p,g,m,v (16 B) and writes p,m,v (+2 B bf16 shadow) per
// element in ONE pass, instead of torch's multi-tensor AdamW followed by a separate cast kernel.
#include "common.cuh"

namespace vg {

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ shadow, int64_t n, float lr, float beta1, float beta2, float eps,
             float decay_mul, float inv_bc1, float inv_sqrt_bc2, float grad_scale,
             const float* __restrict__ hyper) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (hyper) {   // device-resident {lr, 1/bias_corr1, 1/sqrt(bias_corr2), 1 - lr*wd}: graph replays see new values
    lr = hyper[0]; inv_bc1 = hyper[1]; inv_sqrt_bc2 = hyper[2]; decay_mul = hyper[3];
  }
  if (i4 + 4 <= n) {
    float4 pv = *reinterpret_cast<float4*>(p + i4);
    const float4 gv = *reinterpret_cast<const float4*>(g + i4);
    float4 mv = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * grad_scale;
      mp[j] = beta1 * mp[j] + (1.f - beta1) * gj;
      vp[j] = beta2 * vp[j] + (1.f - beta2) * gj * gj;
      const float denom = sqrtf(vp[j]) * inv_sqrt_bc2 + eps;
      pp[j] = pp[j] * decay_mul - lr * inv_bc1 * (mp[j] / denom);
    }
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(m + i4) = mv;
    *reinterpret_cast<float4*>(v + i4) = vv;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(shadow + i4) = pk;
    }
  } else {
    for (int64_t i = i4; i < n; ++i) {
      const float gj = g[i] * grad_scale;
      const float mj = beta1 * m[i] + (1.f - beta1) * gj;
      const float vj = beta2 * v[i] + (1.f - beta2) * gj * gj;
      m[i] = mj; v[i] = vj;
      const float denom = sqrtf(vj) * inv_sqrt_bc2 + eps;
      const float pj = p[i] * decay_mul - lr * inv_bc1 * (mj / denom);
      p[i] = pj;
      if (shadow) shadow[i] = __float2bfloat16_rn(pj);
    }
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n && ((reinterpret_cast<uintptr_t>(src + i4) & 15) == 0) &&
      ((reinterpret_cast<uintptr_t>(dst + i4) & 7) == 0)) {
    const float4 s = *reinterpret_cast<const float4*>(src + i4);
    __nv_bfloat162 lo = __floats2bfloat162_rn(s.x, s.y), hi = __floats2bfloat162_rn(s.z, s.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + i4) = pk;
  } else {
    for (int64_t i = i4; i < n && i < i4 + 4; ++i) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// zero a list of [offset, offset + length) element ranges of one fp32 buffer (offsets / lengths multiples of 4):
// blockIdx.y = range, blockIdx.x strides over it
__global__ void zero_segments_kernel(float* __restrict__ base, const int64_t* __restrict__ seg_off,
                                     const int64_t* __restrict__ seg_len) {
  float4* p = reinterpret_cast<float4*>(base + seg_off[blockIdx.y]);
  const int64_t n4 = seg_len[blockIdx.y] / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace vg

using namespace vg;

extern "C" int vg_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                             int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                             float bias_corr1, float bias_corr2, float grad_scale, const float* hyper_dev,
                             vg_stream_t stream) {
  VG_REQUIRE(param && grad && exp_avg && exp_avg_sq, -1, "vg_adamw_step: null pointer");
  VG_REQUIRE(n > 0, -3, "vg_adamw_step: n must be positive");
  VG_REQUIRE(aligned(param, 16) && aligned(grad, 16) && aligned(exp_avg, 16) && aligned(exp_avg_sq, 16) &&
                 (!shadow_bf16 || aligned(shadow_bf16, 8)), -4, "vg_adamw_step: arenas must be 16-byte aligned");
  VG_REQUIRE(bias_corr1 > 0.f && bias_corr2 > 0.f, -3, "vg_adamw_step: bias corrections must be > 0");
  const int64_t nthreads = ceil_div(n, 4);
  adamw_kernel<<<(unsigned)ceil_div(nthreads, 256), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, (__nv_bfloat16*)shadow_bf16, n, lr, beta1, beta2, eps,
      1.f - lr * weight_decay, 1.f / bias_corr1, 1.f / sqrtf(bias_corr2), grad_scale, hyper_dev);
  VG_LAUNCH_CHECK("vg_adamw_step");
  return 0;
}

extern "C" int vg_zero_segments(float* base, const int64_t* seg_off, const int64_t* seg_len, int n_seg,
                                vg_stream_t stream) {
  VG_REQUIRE(base && seg_off && seg_len, -1, "vg_zero_segments: null pointer");
  VG_REQUIRE(n_seg > 0 && n_seg <= 65535, -3, "vg_zero_segments: 1..65535 segments");
  VG_REQUIRE(aligned(base, 16), -4, "vg_zero_segments: base must be 16-byte aligned");
  zero_segments_kernel<<<dim3(64, (unsigned)n_seg), 256, 0, (cudaStream_t)stream>>>(base, seg_off, seg_len);
  VG_LAUNCH_CHECK("vg_zero_segments");
  return 0;
}

extern "C" int vg_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vg_stream_t stream) {
  VG_REQUIRE(src && dst, -1, "vg_cast_f32_to_bf16: null pointer");
  VG_REQUIRE(n > 0, -3, "vg_cast_f32_to_bf16: n must be positive");
  const int64_t nthreads = ceil_div(n, 4);
  cast_bf16_kernel<<<(unsigned)ceil_div(nthreads, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  VG_LAUNCH_CHECK("vg_cast_f32_to_bf16");
  return 0;
}
