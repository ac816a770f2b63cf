// attn_decode.cu — KV-cache attention for the cached LVTR.step loop (lvtr.py:253-257 →
// attention.py:56-85).  The reference concatenates the whole K/V per layer per step and rebuilds
// an O(Tk^2) mask that it then slices to its last row; here the cache is a pre-allocated head-major
// [B,H,Tmax,D] buffer, the new token's k/v are appended in the same kernel that attends, and the
// ALiBi bias -slope_h*(pos-j) is computed in registers.
//
// HBM-bound (reads Tk*D*2 elements per (b,h)): 8 lanes share one key row with 16-byte loads, so a
// warp instruction covers 4 full 128-byte (bf16) rows; split-KV fills the machine at small batch.
#include <math_constants.h>
#include "common.cuh"

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
