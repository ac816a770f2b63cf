// loss.cu — token cross-entropy, diffusion q_sample / masked L1, and token sampling.
//   softmax-CE : training_lib/losses.py:30-41 (F.cross_entropy, ignore_index=-100, reduction="sum")
//   q_sample   : modules/diffusion/ddpm.py:328-334 ;  masked L1: ddpm.py:345-366 + losses.py:9-27,44-57
//   sampling   : models/speech/lvtr.py:277-285 (softmax(logits/τ) → multinomial) + greedy argmax
// All HBM-bound: one warp per row, one pass over the data, deterministic two-stage sums.
#include <math_constants.h>
#include "common.cuh"

namespace vg {

constexpr int kRowsPerBlock = 4;   // warps per CTA

// ------------------------------------------------------------------------------------ softmax-CE
template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
softmax_ce_fwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ targets,
                      const uint8_t* __restrict__ mask, float* __restrict__ lse, float* __restrict__ loss_rows,
                      float* __restrict__ partial, int64_t rows, int vocab) {
  __shared__ float red[kRowsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kRowsPerBlock + warp;
  float loss = 0.f;
  if (r < rows) {
    const T* row = logits + r * ld;
    float mx = -CUDART_INF_F;
    for (int c = lane; c < vocab; c += 32) mx = fmaxf(mx, to_f32<T>(row[c]));
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < vocab; c += 32) se += expf(to_f32<T>(row[c]) - mx);
    se = warp_sum(se);
    const float L = mx + logf(se);
    const int64_t tgt = targets[r];
    const bool use = (!mask || mask[r]) && tgt >= 0 && tgt < vocab;     // ignore_index / padded rows
    if (use) loss = L - to_f32<T>(row[tgt]);
    if (lane == 0) { lse[r] = L; loss_rows[r] = loss; }
  }
  if (lane == 0) red[warp] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRowsPerBlock; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

// fixed-order tree: one CTA, thread-strided then warp/block reduce → deterministic
__global__ void __launch_bounds__(256) sum_f32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
softmax_ce_bwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ targets,
                      const uint8_t* __restrict__ mask, const float* __restrict__ lse,
                      const float* __restrict__ d_loss, T* __restrict__ d_logits, int64_t ld_d, int64_t rows,
                      int vocab) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kRowsPerBlock + warp;
  if (r >= rows) return;
  const int64_t tgt = targets[r];
  const bool use = (!mask || mask[r]) && tgt >= 0 && tgt < vocab;
  const float g = use ? d_loss[0] : 0.f;
  const float L = lse[r];
  const T* row = logits + r * ld;
  T* drow = d_logits + r * ld_d;
  for (int c = lane; c < vocab; c += 32) {
    float v = 0.f;
    if (use) v = g * (expf(to_f32<T>(row[c]) - L) - (c == tgt ? 1.f : 0.f));
    drow[c] = from_f32<T>(v);
  }
}

// ------------------------------------------------------------------------------------ diffusion glue
__global__ void qsample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                               const int64_t* __restrict__ t, const float* __restrict__ sqrt_ac,
                               const float* __restrict__ sqrt_1mac, const uint8_t* __restrict__ mask, float x0_scale,
                               float* __restrict__ x_t, float* __restrict__ target, int64_t T, int64_t C,
                               int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t frame = i / C;
  const int64_t b = frame / T;
  const bool valid = mask[frame] != 0;
  const int64_t tb = t[b];
  const float n = noise[i];
  x_t[i] = valid ? sqrt_ac[tb] * (x0[i] * x0_scale) + sqrt_1mac[tb] * n : 0.f;
  target[i] = valid ? n : 0.f;
}

// one warp per frame: sum_c |pred - target| / C  (masked); per-CTA partials
__global__ void __launch_bounds__(kRowsPerBlock * 32)
masked_l1_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                     const uint8_t* __restrict__ mask, float* __restrict__ partial, int64_t frames, int C) {
  __shared__ float red[kRowsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t f = (int64_t)blockIdx.x * kRowsPerBlock + warp;
  float s = 0.f;
  if (f < frames && mask[f]) {
    for (int c = lane; c < C; c += 32) s += fabsf(pred[f * C + c] - target[f * C + c]);
    s = warp_sum(s) / (float)C;
  }
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kRowsPerBlock; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void masked_l1_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                     const uint8_t* __restrict__ mask, const float* __restrict__ d_loss,
                                     float* __restrict__ d_pred, int64_t C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const bool valid = mask[i / C] != 0;
  const float d = pred[i] - target[i];
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);      // torch.abs backward: sign(0) = 0
  d_pred[i] = valid ? d_loss[0] * sgn / (float)C : 0.f;
}

// ------------------------------------------------------------------------------------ token sampling
template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
sample_token_kernel(const T* __restrict__ logits, int64_t ld, const float* __restrict__ u, float temperature,
                    int64_t* __restrict__ out_ids, int64_t rows, int vocab) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kRowsPerBlock + warp;
  if (r >= rows) return;
  const T* row = logits + r * ld;
  // argmax (lowest index wins ties) — also the softmax max
  float mx = -CUDART_INF_F;
  int arg = 0x7fffffff;
  for (int c = lane; c < vocab; c += 32) {
    const float v = to_f32<T>(row[c]);
    if (v > mx) { mx = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  if (!u) {
    if (lane == 0) out_ids[r] = arg;
    return;
  }
  // inverse-CDF multinomial over softmax(logits / temperature), vocabulary order
  const float inv_t = 1.f / temperature;
  float se = 0.f;
  for (int c = lane; c < vocab; c += 32) se += expf((to_f32<T>(row[c]) - mx) * inv_t);
  se = warp_sum(se);
  const float thresh = u[r] * se;
  float run = 0.f;
  int pick = -1;
  for (int c0 = 0; c0 < vocab && pick < 0; c0 += 32) {
    const int c = c0 + lane;
    const float p = c < vocab ? expf((to_f32<T>(row[c]) - mx) * inv_t) : 0.f;
    float incl = p;                                   // inclusive scan over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float nb = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += nb;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, c < vocab && run + incl > thresh);
    if (hit) pick = c0 + __ffs(hit) - 1;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (pick < 0) pick = vocab - 1;
  if (lane == 0) out_ids[r] = pick;
}

}  // namespace vg

using namespace vg;

extern "C" size_t vg_softmax_ce_workspace(int64_t rows) {
  return (size_t)ceil_div(rows > 0 ? rows : 1, kRowsPerBlock) * sizeof(float);
}

extern "C" int vg_softmax_ce_fwd(const void* logits, int64_t ld, const int64_t* targets, const uint8_t* row_mask,
                                 float* lse, float* loss_rows, float* loss_sum, int64_t rows, int64_t vocab,
                                 int dtype, void* workspace, size_t workspace_bytes, vg_stream_t stream) {
  VG_REQUIRE(logits && targets && lse && loss_rows && loss_sum, -1, "vg_softmax_ce_fwd: null pointer");
  VG_REQUIRE(valid_dtype(dtype), -2, "vg_softmax_ce_fwd: bad dtype");
  VG_REQUIRE(rows > 0 && vocab > 0 && ld >= vocab, -3, "vg_softmax_ce_fwd: bad shape");
  VG_REQUIRE(workspace && workspace_bytes >= vg_softmax_ce_workspace(rows), -5, "vg_softmax_ce_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = (unsigned)ceil_div(rows, kRowsPerBlock);
  if (dtype == VG_F32)
    softmax_ce_fwd_kernel<float><<<nb, kRowsPerBlock * 32, 0, st>>>((const float*)logits, ld, targets, row_mask, lse,
                                                                   loss_rows, (float*)workspace, rows, (int)vocab);
  else
    softmax_ce_fwd_kernel<__nv_bfloat16><<<nb, kRowsPerBlock * 32, 0, st>>>(
        (const __nv_bfloat16*)logits, ld, targets, row_mask, lse, loss_rows, (float*)workspace, rows, (int)vocab);
  VG_LAUNCH_CHECK("vg_softmax_ce_fwd");
  sum_f32_kernel<<<1, 256, 0, st>>>((const float*)workspace, nb, loss_sum);
  VG_LAUNCH_CHECK("vg_softmax_ce_fwd(sum)");
  return 0;
}

extern "C" int vg_softmax_ce_bwd(const void* logits, int64_t ld, const int64_t* targets, const uint8_t* row_mask,
                                 const float* lse, const float* d_loss, void* d_logits, int64_t ld_d, int64_t rows,
                                 int64_t vocab, int dtype, vg_stream_t stream) {
  VG_REQUIRE(logits && targets && lse && d_loss && d_logits, -1, "vg_softmax_ce_bwd: null pointer");
  VG_REQUIRE(valid_dtype(dtype), -2, "vg_softmax_ce_bwd: bad dtype");
  VG_REQUIRE(rows > 0 && vocab > 0 && ld >= vocab && ld_d >= vocab, -3, "vg_softmax_ce_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = (unsigned)ceil_div(rows, kRowsPerBlock);
  if (dtype == VG_F32)
    softmax_ce_bwd_kernel<float><<<nb, kRowsPerBlock * 32, 0, st>>>((const float*)logits, ld, targets, row_mask, lse,
                                                                   d_loss, (float*)d_logits, ld_d, rows, (int)vocab);
  else
    softmax_ce_bwd_kernel<__nv_bfloat16><<<nb, kRowsPerBlock * 32, 0, st>>>(
        (const __nv_bfloat16*)logits, ld, targets, row_mask, lse, d_loss, (__nv_bfloat16*)d_logits, ld_d, rows,
        (int)vocab);
  VG_LAUNCH_CHECK("vg_softmax_ce_bwd");
  return 0;
}

extern "C" int vg_qsample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                          const float* sqrt_1mac, const uint8_t* mask, float x0_scale, float* x_t, float* target,
                          int64_t B, int64_t T, int64_t C, vg_stream_t stream) {
  VG_REQUIRE(x0 && noise && t && sqrt_ac && sqrt_1mac && mask && x_t && target, -1, "vg_qsample: null pointer");
  VG_REQUIRE(B > 0 && T > 0 && C > 0, -3, "vg_qsample: bad shape");
  const int64_t total = B * T * C;
  qsample_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x0, noise, t, sqrt_ac, sqrt_1mac, mask,
                                                                                 x0_scale, x_t, target, T, C, total);
  VG_LAUNCH_CHECK("vg_qsample");
  return 0;
}

extern "C" size_t vg_masked_l1_workspace(int64_t B, int64_t T, int64_t) {
  return (size_t)ceil_div(B * T > 0 ? B * T : 1, kRowsPerBlock) * sizeof(float);
}

extern "C" int vg_masked_l1_fwd(const float* pred, const float* target, const uint8_t* mask, float* loss_sum,
                                int64_t B, int64_t T, int64_t C, void* workspace, size_t workspace_bytes,
                                vg_stream_t stream) {
  VG_REQUIRE(pred && target && mask && loss_sum, -1, "vg_masked_l1_fwd: null pointer");
  VG_REQUIRE(B > 0 && T > 0 && C > 0, -3, "vg_masked_l1_fwd: bad shape");
  VG_REQUIRE(workspace && workspace_bytes >= vg_masked_l1_workspace(B, T, C), -5, "vg_masked_l1_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = (unsigned)ceil_div(B * T, kRowsPerBlock);
  masked_l1_fwd_kernel<<<nb, kRowsPerBlock * 32, 0, st>>>(pred, target, mask, (float*)workspace, B * T, (int)C);
  VG_LAUNCH_CHECK("vg_masked_l1_fwd");
  sum_f32_kernel<<<1, 256, 0, st>>>((const float*)workspace, nb, loss_sum);
  VG_LAUNCH_CHECK("vg_masked_l1_fwd(sum)");
  return 0;
}

extern "C" int vg_masked_l1_bwd(const float* pred, const float* target, const uint8_t* mask, const float* d_loss,
                                float* d_pred, int64_t B, int64_t T, int64_t C, vg_stream_t stream) {
  VG_REQUIRE(pred && target && mask && d_loss && d_pred, -1, "vg_masked_l1_bwd: null pointer");
  VG_REQUIRE(B > 0 && T > 0 && C > 0, -3, "vg_masked_l1_bwd: bad shape");
  const int64_t total = B * T * C;
  masked_l1_bwd_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(pred, target, mask, d_loss,
                                                                                       d_pred, C, total);
  VG_LAUNCH_CHECK("vg_masked_l1_bwd");
  return 0;
}

extern "C" int vg_sample_token(const void* logits, int64_t ld, const float* u, float temperature, int64_t* out_ids,
                               int64_t rows, int64_t vocab, int dtype, vg_stream_t stream) {
  VG_REQUIRE(logits && out_ids, -1, "vg_sample_token: null pointer");
  VG_REQUIRE(valid_dtype(dtype), -2, "vg_sample_token: bad dtype");
  VG_REQUIRE(rows > 0 && vocab > 0 && ld >= vocab, -3, "vg_sample_token: bad shape");
  VG_REQUIRE(!u || temperature > 0.f, -3, "vg_sample_token: temperature must be > 0 when sampling");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = (unsigned)ceil_div(rows, kRowsPerBlock);
  if (dtype == VG_F32)
    sample_token_kernel<float><<<nb, kRowsPerBlock * 32, 0, st>>>((const float*)logits, ld, u, temperature, out_ids,
                                                                 rows, (int)vocab);
  else
    sample_token_kernel<__nv_bfloat16><<<nb, kRowsPerBlock * 32, 0, st>>>((const __nv_bfloat16*)logits, ld, u,
                                                                         temperature, out_ids, rows, (int)vocab);
  VG_LAUNCH_CHECK("vg_sample_token");
  return 0;
}
