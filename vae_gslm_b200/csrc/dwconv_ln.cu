// dwconv_ln.cu — depthwise k-tap convolution along time + per-(batch,channel) add + channel LayerNorm, fused,
// on [B,T,C] rows.  Reference: the front half of every ResidualBlock of the conv encoder / diffusion UNet,
//   norm(conv1(x) [+ time_emb]) , modules/conv/layers.py:117-135,238-253,275-295 with Conv1d :13-31 (asymmetric zero
//   padding at the ARRAY ends) and the channel "InstanceNorm" of modules/norm.py:43-47 (per-frame statistics over
//   channels, UNBIASED variance).
// The reference runs this in B,C,T layout through cuDNN + ~8 elementwise/reduce kernels and two transposes; here the
// activations stay B,T,C (so the 1x1 convolutions around it are plain GEMMs on the same buffers) and one warp owns
// one frame: 16-byte vector loads of the k neighbouring frames, statistics by warp shuffles, one store.
// HBM-bound: forward reads x once (neighbour rows hit L1/L2) and writes y once.
// Backward: kernel A recomputes the conv output, produces dh = dL/d(conv output) and accumulates the parameter
// gradients (conv taps, conv bias, LN weight/bias) in per-warp shared-memory slots (each lane owns its channels →
// no atomics), reduced deterministically in two stages; kernel B is the transposed depthwise conv dx = corr(dh, w).
#include "common.cuh"

namespace vg {

constexpr int DW_WARPS = 4;
constexpr int DW_MAX_TAPS = 8;

struct DwShape {
  int B, T, C, taps, pad_left;
  float eps;
};

// conv output (+bias, +t_add) of one frame for this lane's channels: ITERS x 8 channels at (it*32+lane)*8
template <typename T, int ITERS>
__device__ __forceinline__ void dw_conv_row(const T* __restrict__ x, const float* __restrict__ w_t,
                                            const float* __restrict__ bias, const float* __restrict__ t_add,
                                            const DwShape& s, int b, int t, int lane, float (&h)[ITERS][8]) {
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (it * 32 + lane) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) h[it][j] = 0.f;
    if (c >= s.C) continue;
    if (bias) {
      Vec8<float> bv;
      bv.load(bias + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[it][j] = bv.v[j];
    }
    if (t_add) {
      Vec8<float> tv;
      tv.load(t_add + (int64_t)b * s.C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[it][j] += tv.v[j];
    }
    for (int k = 0; k < s.taps; ++k) {
      const int tt = t - s.pad_left + k;
      if (tt < 0 || tt >= s.T) continue;
      Vec8<T> xv;
      xv.load(x + ((int64_t)b * s.T + tt) * s.C + c);
      if (w_t) {
        Vec8<float> wv;
        wv.load(w_t + (int64_t)k * s.C + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) h[it][j] = fmaf(wv.v[j], xv.v[j], h[it][j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) h[it][j] += xv.v[j];     // identity "conv" (plain channel LayerNorm)
      }
    }
  }
}

template <typename T, int ITERS>
__global__ void __launch_bounds__(DW_WARPS * 32)
dwconv_ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w_t, const float* __restrict__ bias,
                     const float* __restrict__ t_add, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                     T* __restrict__ y, int64_t ld_y, float* __restrict__ mean, float* __restrict__ rstd, DwShape s) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * DW_WARPS + (threadIdx.x >> 5);
  if (row >= (int64_t)s.B * s.T) return;
  const int b = (int)(row / s.T), t = (int)(row % s.T);
  float h[ITERS][8];
  dw_conv_row<T, ITERS>(x, w_t, bias, t_add, s, b, t, lane, h);
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += h[it][j];       // lanes beyond C hold zeros
  const float mu = warp_sum(sum) / (float)s.C;
  float sq = 0.f;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (it * 32 + lane) * 8;
    if (c < s.C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = h[it][j] - mu; sq += d * d; }
    }
  }
  const float var = warp_sum(sq) / (float)(s.C - 1);     // torch.var_mean default: unbiased
  const float r = rsqrtf(var + s.eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = r; }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (it * 32 + lane) * 8;
    if (c < s.C) {
      Vec8<float> wv, bv;
      wv.load(ln_w + c);
      bv.load(ln_b + c);
      Vec8<T> o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = fmaf((h[it][j] - mu) * r, wv.v[j], bv.v[j]);
      o.store(y + row * ld_y + c);
    }
  }
}

// Backward kernel A: dh = dL/d(conv output) per frame, plus the three per-channel sums that only need the frame itself
// (d_ln_w, d_ln_b, d_bias) accumulated in REGISTERS (each lane owns fixed channels) — the tap gradients, which need the
// neighbouring frames, are a separate strip kernel (dwconv_wgrad_kernel): accumulating all (taps+3)·C sums in shared
// memory cost (taps+3)·16 read-modify-writes per lane per frame and made this the slowest kernel of the conv stack.
// partial layout per CTA (floats): d_ln_w [C] | d_ln_b [C] | d_bias [C]
template <typename T, int ITERS>
__global__ void __launch_bounds__(DW_WARPS * 32)
dwconv_ln_bwd_a_kernel(const T* __restrict__ dy, int64_t ld_dy, const T* __restrict__ x,
                       const float* __restrict__ w_t, const float* __restrict__ bias,
                       const float* __restrict__ t_add, const float* __restrict__ ln_w,
                       const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dh,
                       float* __restrict__ partial, DwShape s) {
  extern __shared__ float acc_smem[];            // [DW_WARPS][3*C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int np = 3 * s.C;
  float a_w[ITERS][8], a_b[ITERS][8], a_c[ITERS][8];
#pragma unroll
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) { a_w[it][j] = 0.f; a_b[it][j] = 0.f; a_c[it][j] = 0.f; }
  const int64_t rows = (int64_t)s.B * s.T;
  for (int64_t row = (int64_t)blockIdx.x * DW_WARPS + warp; row < rows; row += (int64_t)gridDim.x * DW_WARPS) {
    const int b = (int)(row / s.T), t = (int)(row % s.T);
    float h[ITERS][8];
    dw_conv_row<T, ITERS>(x, w_t, bias, t_add, s, b, t, lane, h);
    const float mu = mean[row], r = rstd[row];
    float g[ITERS][8];                            // dL/d(hhat)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < s.C) {
        Vec8<T> dv;
        dv.load(dy + row * ld_dy + c);
        Vec8<float> wv;
        wv.load(ln_w + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float hh = (h[it][j] - mu) * r;
          a_w[it][j] = fmaf(dv.v[j], hh, a_w[it][j]);          // d_ln_w
          a_b[it][j] += dv.v[j];                               // d_ln_b
          g[it][j] = dv.v[j] * wv.v[j];
          h[it][j] = hh;
          s1 += g[it][j];
          s2 += g[it][j] * hh;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[it][j] = 0.f;
      }
    }
    s1 = warp_sum(s1) / (float)s.C;
    s2 = warp_sum(s2) / (float)(s.C - 1);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < s.C) {
        Vec8<T> o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = r * (g[it][j] - s1 - h[it][j] * s2);      // dL/d(conv output)
          o.v[j] = d;
          a_c[it][j] += d;                                     // d_bias
        }
        o.store(dh + row * s.C + c);
      }
    }
  }
  float* acc = acc_smem + warp * np;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (it * 32 + lane) * 8;
    if (c < s.C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[c + j] = a_w[it][j];
        acc[s.C + c + j] = a_b[it][j];
        acc[2 * s.C + c + j] = a_c[it][j];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < DW_WARPS; ++w) v += acc_smem[w * np + i];
    partial[(int64_t)blockIdx.x * np + i] = v;
  }
}

// Tap gradients and the per-sequence sum of dh (gradient of the time embedding added before the norm):
//   d_w[k][c] = sum_{b,t} dh[b,t,c] * x[b, t - pad_left + k, c]        d_t[b][c] = sum_t dh[b,t,c]
// One thread owns 8 channels and walks a strip of DW_STRIP frames of one sequence; the k neighbouring x rows of
// consecutive frames overlap and come from L1.  partial layout per strip (floats): d_w [taps][C] | sum_dh [C]
constexpr int DW_STRIP = 16;
template <typename T>
__global__ void __launch_bounds__(64)
dwconv_wgrad_kernel(const T* __restrict__ dh, const T* __restrict__ x, float* __restrict__ partial, DwShape s,
                    int strips_per_seq) {
  const int c = (blockIdx.x * 64 + threadIdx.x) * 8;
  if (c >= s.C) return;
  const int b = blockIdx.y / strips_per_seq, strip = blockIdx.y - b * strips_per_seq;
  const int t0 = strip * DW_STRIP, t1 = min(s.T, t0 + DW_STRIP);
  float aw[DW_MAX_TAPS][8], asum[8];
#pragma unroll
  for (int k = 0; k < DW_MAX_TAPS; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) aw[k][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) asum[j] = 0.f;
  for (int t = t0; t < t1; ++t) {
    Vec8<T> g;
    g.load(dh + ((int64_t)b * s.T + t) * s.C + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) asum[j] += g.v[j];
#pragma unroll
    for (int k = 0; k < DW_MAX_TAPS; ++k) {
      if (k >= s.taps) break;
      const int tt = t - s.pad_left + k;
      if (tt < 0 || tt >= s.T) continue;
      Vec8<T> xv;
      xv.load(x + ((int64_t)b * s.T + tt) * s.C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) aw[k][j] = fmaf(g.v[j], xv.v[j], aw[k][j]);
    }
  }
  float* pp = partial + (int64_t)blockIdx.y * (s.taps + 1) * s.C;
#pragma unroll
  for (int k = 0; k < DW_MAX_TAPS; ++k) {
    if (k >= s.taps) break;
#pragma unroll
    for (int j = 0; j < 8; ++j) pp[(int64_t)k * s.C + c + j] = aw[k][j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) pp[(int64_t)s.taps * s.C + c + j] = asum[j];
}

// d_w[k][c] = sum over all strips; d_t[b][c] = sum over the strips of sequence b.  Fixed order → deterministic.
// One CTA = 32 outputs x 8 groups of partials (shared-memory tree), grid over (taps + B)·C / 32.
__global__ void __launch_bounds__(256)
dw_reduce_wgrad_kernel(const float* __restrict__ partial, int B, int strips_per_seq, int taps, int C,
                       float* __restrict__ dw_t, float* __restrict__ d_t) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;                      // over (taps + B) * C
  const int np = (taps + 1) * C;
  float v = 0.f;
  bool live = false;
  float* dst = nullptr;
  if (i < taps * C) {
    live = dw_t != nullptr;
    dst = dw_t + i;
    if (live)
      for (int p = ty; p < B * strips_per_seq; p += 8) v += partial[(int64_t)p * np + i];
  } else if (i < (taps + B) * C) {
    const int b = (i - taps * C) / C, c = (i - taps * C) - b * C;
    live = d_t != nullptr;
    dst = d_t + (int64_t)b * C + c;
    if (live)
      for (int p = ty; p < strips_per_seq; p += 8) v += partial[(int64_t)(b * strips_per_seq + p) * np + taps * C + c];
  }
  red[ty][tx] = v;
  __syncthreads();
  if (ty == 0 && live) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][tx];
    *dst = t;
  }
}

// dx[b,t,c] = sum_k w_t[k][c] * dh[b, t + pad_left - k, c]
template <typename T>
__global__ void __launch_bounds__(256)
dwconv_bwd_dx_kernel(const T* __restrict__ dh, const float* __restrict__ w_t, T* __restrict__ dx, DwShape s) {
  const int c8 = s.C / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)s.B * s.T * c8) return;
  const int c = (int)(i % c8) * 8;
  const int64_t row = i / c8;
  const int b = (int)(row / s.T), t = (int)(row % s.T);
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  for (int k = 0; k < s.taps; ++k) {
    const int tt = t + s.pad_left - k;
    if (tt < 0 || tt >= s.T) continue;
    Vec8<T> g;
    g.load(dh + ((int64_t)b * s.T + tt) * s.C + c);
    if (w_t) {
      Vec8<float> wv;
      wv.load(w_t + (int64_t)k * s.C + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(wv.v[j], g.v[j], a[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += g.v[j];
    }
  }
  Vec8<T> o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = a[j];
  o.store(dx + row * s.C + c);
}

__global__ void __launch_bounds__(256)
dw_reduce_partials_kernel(const float* __restrict__ partial, int nparts, int np, int taps, int C,
                          float* __restrict__ dw_t, float* __restrict__ d_ln_w, float* __restrict__ d_ln_b,
                          float* __restrict__ d_bias) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float sacc = 0.f;
  if (i < np)
    for (int p = ty; p < nparts; p += 8) sacc += partial[(int64_t)p * np + i];
  red[ty][tx] = sacc;
  __syncthreads();
  if (ty == 0 && i < np) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += red[g][tx];
    (void)taps; (void)dw_t;
    if (i < C) d_ln_w[i] = v;
    else if (i < 2 * C) d_ln_b[i - C] = v;
    else if (d_bias) d_bias[i - 2 * C] = v;
  }
}

static int dw_bwd_blocks(int64_t rows) {
  int64_t want = ceil_div(rows, DW_WARPS);
  int64_t cap = kNumSMs * 4;
  return (int)(want < cap ? want : cap);
}

}  // namespace vg

using namespace vg;

static int dw_check(const char* fn, int64_t B, int64_t T, int64_t C, int32_t taps, int32_t pad_left, int dtype) {
  VG_REQUIRE(valid_dtype(dtype), -2, "%s: bad dtype", fn);
  VG_REQUIRE(B > 0 && T > 0 && C >= 16 && C % 8 == 0 && C <= 1024, -3, "%s: C=%lld must be a multiple of 8 in [16,1024]",
             fn, (long long)C);
  VG_REQUIRE(taps >= 1 && taps <= DW_MAX_TAPS && pad_left >= 0 && pad_left < taps, -3, "%s: bad taps/padding", fn);
  return 0;
}

extern "C" int vg_dwconv_ln_fwd(const void* x, const float* w_t, const float* bias, const float* t_add,
                                const float* ln_w, const float* ln_b, void* y, int64_t ld_y, float* mean, float* rstd,
                                int64_t B, int64_t T, int64_t C, int32_t taps, int32_t pad_left, float eps, int dtype,
                                vg_stream_t stream) {
  VG_REQUIRE(x && ln_w && ln_b && y && mean && rstd, -1, "vg_dwconv_ln_fwd: null pointer");
  if (int rc = dw_check("vg_dwconv_ln_fwd", B, T, C, taps, pad_left, dtype)) return rc;
  VG_REQUIRE(ld_y >= C && ld_y % 8 == 0 && aligned(x, 16) && aligned(y, 16), -4, "vg_dwconv_ln_fwd: unaligned");
  DwShape s{(int)B, (int)T, (int)C, taps, pad_left, eps};
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div(B * T, DW_WARPS);
  const int iters = (int)ceil_div(C, 256);
#define VG_DW_FWD(TT, I)                                                                                        \
  dwconv_ln_fwd_kernel<TT, I><<<grid, DW_WARPS * 32, 0, st>>>((const TT*)x, w_t, bias, t_add, ln_w, ln_b, (TT*)y, \
                                                             ld_y, mean, rstd, s)
  if (dtype == VG_F32) { if (iters <= 1) VG_DW_FWD(float, 1); else if (iters <= 2) VG_DW_FWD(float, 2); else VG_DW_FWD(float, 4); }
  else { if (iters <= 1) VG_DW_FWD(__nv_bfloat16, 1); else if (iters <= 2) VG_DW_FWD(__nv_bfloat16, 2); else VG_DW_FWD(__nv_bfloat16, 4); }
#undef VG_DW_FWD
  VG_LAUNCH_CHECK("vg_dwconv_ln_fwd");
  return 0;
}

static size_t dw_partial_a_bytes(int64_t B, int64_t T, int64_t C) {
  return align_up((size_t)dw_bwd_blocks(B * T) * 3 * (size_t)C * sizeof(float), 256);
}
extern "C" size_t vg_dwconv_ln_bwd_workspace(int64_t B, int64_t T, int64_t C, int32_t taps) {
  return dw_partial_a_bytes(B, T, C) + (size_t)B * (size_t)ceil_div(T, DW_STRIP) * (size_t)(taps + 1) * (size_t)C * sizeof(float);
}

extern "C" int vg_dwconv_ln_bwd(const void* dy, int64_t ld_dy, const void* x, const float* w_t, const float* bias,
                                const float* t_add, const float* ln_w, const float* mean, const float* rstd, void* dh,
                                void* dx, float* dw_t, float* d_ln_w, float* d_ln_b, float* d_bias, float* d_t_add,
                                void* workspace, size_t workspace_bytes, int64_t B, int64_t T, int64_t C, int32_t taps,
                                int32_t pad_left, int dtype, vg_stream_t stream) {
  VG_REQUIRE(dy && x && ln_w && mean && rstd && dh && dx && d_ln_w && d_ln_b, -1, "vg_dwconv_ln_bwd: null pointer");
  if (int rc = dw_check("vg_dwconv_ln_bwd", B, T, C, taps, pad_left, dtype)) return rc;
  VG_REQUIRE(ld_dy >= C && ld_dy % 8 == 0 && aligned(dy, 16) && aligned(x, 16) && aligned(dh, 16) && aligned(dx, 16),
             -4, "vg_dwconv_ln_bwd: unaligned");
  VG_REQUIRE(workspace && workspace_bytes >= vg_dwconv_ln_bwd_workspace(B, T, C, taps), -5,
             "vg_dwconv_ln_bwd: workspace too small");
  DwShape s{(int)B, (int)T, (int)C, taps, pad_left, 0.f};
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = dw_bwd_blocks(B * T);
  const int np = 3 * (int)C;
  const int smem = DW_WARPS * np * (int)sizeof(float);
  const int iters = (int)ceil_div(C, 256);
  float* partial = (float*)workspace;
#define VG_DW_BWD(TT, I)                                                                                          \
  do {                                                                                                            \
    auto kern = dwconv_ln_bwd_a_kernel<TT, I>;                                                                    \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                             \
                                   DW_WARPS * 3 * 1024 * (int)sizeof(float)));                                    \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    kern<<<nb, DW_WARPS * 32, smem, st>>>((const TT*)dy, ld_dy, (const TT*)x, w_t, bias, t_add, ln_w, mean, rstd,  \
                                          (TT*)dh, partial, s);                                                   \
  } while (0)
  if (dtype == VG_F32) { if (iters <= 1) VG_DW_BWD(float, 1); else if (iters <= 2) VG_DW_BWD(float, 2); else VG_DW_BWD(float, 4); }
  else { if (iters <= 1) VG_DW_BWD(__nv_bfloat16, 1); else if (iters <= 2) VG_DW_BWD(__nv_bfloat16, 2); else VG_DW_BWD(__nv_bfloat16, 4); }
#undef VG_DW_BWD
  VG_LAUNCH_CHECK("vg_dwconv_ln_bwd(a)");
  dw_reduce_partials_kernel<<<(unsigned)ceil_div(np, 32), 256, 0, st>>>(partial, nb, np, taps, (int)C, dw_t, d_ln_w,
                                                                        d_ln_b, d_bias);
  VG_LAUNCH_CHECK("vg_dwconv_ln_bwd(reduce)");
  if (dw_t || d_t_add) {          // tap gradients (need the neighbouring frames) + per-sequence sum of dh
    float* partial_w = (float*)((uint8_t*)workspace + dw_partial_a_bytes(B, T, C));
    const int sps = (int)ceil_div(T, DW_STRIP);
    dim3 gw((unsigned)ceil_div(C / 8, 64), (unsigned)(B * sps));
    if (dtype == VG_F32)
      dwconv_wgrad_kernel<float><<<gw, 64, 0, st>>>((const float*)dh, (const float*)x, partial_w, s, sps);
    else
      dwconv_wgrad_kernel<__nv_bfloat16><<<gw, 64, 0, st>>>((const __nv_bfloat16*)dh, (const __nv_bfloat16*)x, partial_w,
                                                           s, sps);
    VG_LAUNCH_CHECK("vg_dwconv_ln_bwd(wgrad)");
    dw_reduce_wgrad_kernel<<<(unsigned)ceil_div((int64_t)(taps + B) * C, 32), 256, 0, st>>>(
        partial_w, (int)B, sps, taps, (int)C, w_t ? dw_t : nullptr, d_t_add);
    VG_LAUNCH_CHECK("vg_dwconv_ln_bwd(wgrad reduce)");
  }
  const int64_t n = B * T * (C / 8);
  if (dtype == VG_F32)
    dwconv_bwd_dx_kernel<float><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((const float*)dh, w_t, (float*)dx, s);
  else
    dwconv_bwd_dx_kernel<__nv_bfloat16><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((const __nv_bfloat16*)dh, w_t,
                                                                                   (__nv_bfloat16*)dx, s);
  VG_LAUNCH_CHECK("vg_dwconv_ln_bwd(dx)");
  return 0;
}
