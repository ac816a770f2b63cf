// conv_im2col.cu — strided Conv1d of the utterance encoder as a GEMM on [B,T,C] rows: the window gather in front of the
// GEMM and its adjoint (reference: modules/conv/layers.py:549-593 ConvNormAct — Conv1d(k = 4, stride 2, padding 1) →
// channel norm → ReLU, three layers over <= 200 frames; models/speech/lvtr.py:127-136,203-207).
//
//   A[b, t, c*K + j] = f(x[b, S*t - pad + j, c])   (0 outside [0, T)),   f = identity or ReLU
//
// The column order (channel-major, tap-minor) is the memory order of the PyTorch Conv1d weight [Cout, Cin, K], so the
// convolution is ONE vg_gemm with the weight used in place — forward, dgrad and wgrad (the weight gradient lands in
// the parameter's own layout, straight in the gradient arena) — and nothing is transposed to B,C,T.  With f = ReLU the
// activation of the PREVIOUS ConvNormAct is applied on read (and its mask in the adjoint), so norm → ReLU → next
// convolution needs no elementwise kernel in between.  Sizes are tiny (<= 8 x 75 x 256 elements per layer): one
// thread per (b, t, c), coalesced over c.
#include "common.cuh"

namespace vg {

template <typename T>
__global__ void im2col_fwd_kernel(const T* __restrict__ x, T* __restrict__ a, int B, int Tn, int C, int K, int S, int pad,
                                  int Tout, int relu) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Tout * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int t = (int)((idx / C) % Tout);
  const int b = (int)(idx / ((int64_t)C * Tout));
  const T* xb = x + (int64_t)b * Tn * C + c;
  T* ar = a + ((int64_t)b * Tout + t) * ((int64_t)C * K) + (int64_t)c * K;
  for (int j = 0; j < K; ++j) {
    const int ti = S * t - pad + j;
    float v = (ti >= 0 && ti < Tn) ? to_f32(xb[(int64_t)ti * C]) : 0.f;
    if (relu) v = fmaxf(v, 0.f);
    ar[j] = from_f32<T>(v);
  }
}

// dx[b, ti, c] = [x > 0] * sum over (t, j) with S*t - pad + j = ti of dA[b, t, c*K + j]
template <typename T>
__global__ void im2col_bwd_kernel(const T* __restrict__ da, const T* __restrict__ x, T* __restrict__ dx, int B, int Tn,
                                  int C, int K, int S, int pad, int Tout) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Tn * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int ti = (int)((idx / C) % Tn);
  const int b = (int)(idx / ((int64_t)C * Tn));
  float g = 0.f;
  if (!x || to_f32(x[idx]) > 0.f) {
    for (int j = 0; j < K; ++j) {
      const int num = ti + pad - j;
      if (num < 0 || num % S) continue;
      const int t = num / S;
      if (t < Tout) g += to_f32(da[((int64_t)b * Tout + t) * ((int64_t)C * K) + (int64_t)c * K + j]);
    }
  }
  dx[idx] = from_f32<T>(g);
}

}  // namespace vg

using namespace vg;

static int im2col_check(const char* who, int64_t B, int64_t T, int64_t C, int64_t K, int64_t S, int64_t pad, int64_t Tout,
                        int dtype) {
  VG_REQUIRE(valid_dtype(dtype), -2, "%s: bad dtype", who);
  VG_REQUIRE(B > 0 && T > 0 && C > 0 && K >= 1 && K <= 16 && S >= 1 && pad >= 0 && pad < K, -3, "%s: bad shape", who);
  VG_REQUIRE(Tout == (T + 2 * pad - K) / S + 1 && Tout > 0, -3, "%s: T_out %lld does not match (T + 2 pad - K) / S + 1", who,
             (long long)Tout);
  return 0;
}

extern "C" int vg_im2col_fwd(const void* x, void* a, int64_t B, int64_t T, int64_t C, int64_t K, int64_t S, int64_t pad,
                             int64_t Tout, int relu, int dtype, vg_stream_t stream) {
  VG_REQUIRE(x && a, -1, "vg_im2col_fwd: null pointer");
  if (int rc = im2col_check("vg_im2col_fwd", B, T, C, K, S, pad, Tout, dtype)) return rc;
  const int64_t total = B * Tout * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VG_F32)
    im2col_fwd_kernel<float><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>((const float*)x, (float*)a, (int)B, (int)T, (int)C,
                                                                            (int)K, (int)S, (int)pad, (int)Tout, relu);
  else
    im2col_fwd_kernel<__nv_bfloat16><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)a, (int)B, (int)T, (int)C, (int)K, (int)S, (int)pad, (int)Tout, relu);
  VG_LAUNCH_CHECK("vg_im2col_fwd");
  return 0;
}

extern "C" int vg_im2col_bwd(const void* da, const void* x_relu, void* dx, int64_t B, int64_t T, int64_t C, int64_t K,
                             int64_t S, int64_t pad, int64_t Tout, int dtype, vg_stream_t stream) {
  VG_REQUIRE(da && dx, -1, "vg_im2col_bwd: null pointer");
  if (int rc = im2col_check("vg_im2col_bwd", B, T, C, K, S, pad, Tout, dtype)) return rc;
  const int64_t total = B * T * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VG_F32)
    im2col_bwd_kernel<float><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>((const float*)da, (const float*)x_relu, (float*)dx,
                                                                            (int)B, (int)T, (int)C, (int)K, (int)S, (int)pad,
                                                                            (int)Tout);
  else
    im2col_bwd_kernel<__nv_bfloat16><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
        (const __nv_bfloat16*)da, (const __nv_bfloat16*)x_relu, (__nv_bfloat16*)dx, (int)B, (int)T, (int)C, (int)K, (int)S,
        (int)pad, (int)Tout);
  VG_LAUNCH_CHECK("vg_im2col_bwd");
  return 0;
}
