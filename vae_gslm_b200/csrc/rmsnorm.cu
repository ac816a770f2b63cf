// rmsnorm.cu — fused RMSNorm (+ row mask) forward/backward.  Reference: modules/norm.py:22-32 and the
// apply_mask that follows norm1 in transformer/layers.py:53-54.
//
// HBM-bound: one warp owns one row; 16-byte vector accesses; the row stays in registers between
// the mean-square reduction and the scale, so x is read once and y written once
// (algorithmic bytes/row = dim*(sizeof(x)+sizeof(y)) + 4).  Backward accumulates dscale in
// registers per thread (fixed columns) over a grid-strided row loop, then reduces deterministically
// in two stages (no float atomics).
#include "common.cuh"

namespace vg {

constexpr int kRmsWarps = 8;          // warps (= rows in flight) per CTA
constexpr int kRmsMaxIters = 8;       // register-resident path: dim <= 8*256 = 2048

template <typename TX, typename TY, int ITERS>
__global__ void __launch_bounds__(kRmsWarps * 32)
rmsnorm_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ scale,
                   const uint8_t* __restrict__ mask, TY* __restrict__ y, float* __restrict__ rstd,
                   int64_t rows, int dim, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRmsWarps + (threadIdx.x >> 5);
  // programmatic dependent launch (vg_set_pdl_mode; no-ops otherwise): the kernel behind may start its prologue, and x
  // is complete once the kernel in front is
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (row >= rows) return;
  const TX* xr = x + row * dim;
  Vec8<TX> v[ITERS];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (it * 32 + lane) * 8;
    if (c < dim) {
      v[it].load(xr + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[it].v[j] * v[it].v[j];
    }
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / (float)dim + eps);
  if (lane == 0) rstd[row] = r;
  const bool keep = mask ? (mask[row] != 0) : true;
  TY* yr = y + row * dim;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (it * 32 + lane) * 8;
    if (c < dim) {
      Vec8<float> s;
      s.load(scale + c);
      Vec8<TY> o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = keep ? s.v[j] * (v[it].v[j] * r) : 0.f;
      o.store(yr + c);
    }
  }
}

template <typename TX, typename TG, int ITERS>
__global__ void __launch_bounds__(kRmsWarps * 32, 2)
rmsnorm_bwd_kernel(const TG* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ scale,
                   const float* __restrict__ rstd, const uint8_t* __restrict__ mask,
                   const TX* __restrict__ dres, TX* __restrict__ dx, float* __restrict__ partial,
                   int64_t rows, int dim) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched kernel behind may start its prologue (vg_set_pdl_mode)
  __shared__ float red[kRmsWarps][32 * 8];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float acc[ITERS][8];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[it][j] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * kRmsWarps + warp; row < rows;
       row += (int64_t)gridDim.x * kRmsWarps) {
    const bool keep = mask ? (mask[row] != 0) : true;
    const float r = rstd[row];
    Vec8<TX> xv[ITERS];
    Vec8<TG> gv[ITERS];
    float dot = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < dim) {
        xv[it].load(x + row * dim + c);
        gv[it].load(dy + row * dim + c);
        Vec8<float> sc;                    // re-read per row (L1-resident): keeping it in registers spilled at dim 1024
        sc.load(scale + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = keep ? gv[it].v[j] : 0.f;
          const float xh = xv[it].v[j] * r;
          acc[it][j] += g * xh;
          gv[it].v[j] = g * sc.v[j];       // dL/d(xhat)
          xv[it].v[j] = xh;
          dot += gv[it].v[j] * xh;
        }
      }
    }
    dot = warp_sum(dot) / (float)dim;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < dim) {
        Vec8<TX> o;
        if (dres) o.load(dres + row * dim + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = r * (gv[it].v[j] - xv[it].v[j] * dot);
          o.v[j] = dres ? o.v[j] + d : d;
        }
        o.store(dx + row * dim + c);
      }
    }
  }
  // stage 1: deterministic reduction over the CTA's warps, one 256-column slab at a time
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[it][j];
    __syncthreads();
    const int c0 = it * 256;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      if (c0 + i < dim) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kRmsWarps; ++w) s += red[w][i];
        partial[(int64_t)blockIdx.x * dim + c0 + i] = s;
      }
    }
  }
}

// out[c] = sum_p partial[p][c], fixed order (deterministic): 32 columns x 8 row-groups per CTA
__global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int dim, float beta) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < dim)
    for (int p = ty; p < nparts; p += 8) s += partial[(int64_t)p * dim + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < dim) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][tx];
    out[c] = beta != 0.f ? fmaf(beta, out[c], t) : t;
  }
}

static int rms_bwd_blocks(int64_t rows) {
  int64_t want = ceil_div(rows, kRmsWarps);
  int64_t cap = (int64_t)kNumSMs * 2;      // 2 CTAs x 8 warps per SM: 16 rows in flight per SM cover HBM latency
  return (int)(want < cap ? want : cap);
}

template <typename TX, typename TY>
static int launch_fwd(const void* x, const float* scale, const uint8_t* mask, void* y, float* rstd,
                      int64_t rows, int dim, float eps, cudaStream_t st) {
  const int iters = (dim + 255) / 256;
  dim3 grid((unsigned)ceil_div(rows, kRmsWarps)), block(kRmsWarps * 32);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl_mode & 1) ? 1 : 0;
#define VG_RMS_FWD(I)                                                                                        \
  VG_CUDA(cudaLaunchKernelEx(&cfg, rmsnorm_fwd_kernel<TX, TY, I>, (const TX*)x, scale, mask, (TY*)y, rstd, \
                             rows, dim, eps))
  if (iters <= 1) VG_RMS_FWD(1);
  else if (iters <= 2) VG_RMS_FWD(2);
  else if (iters <= 4) VG_RMS_FWD(4);
  else VG_RMS_FWD(8);
#undef VG_RMS_FWD
  VG_LAUNCH_CHECK("vg_rmsnorm_fwd");
  return 0;
}

template <typename TX, typename TG>
static int launch_bwd(const void* dy, const void* x, const float* scale, const float* rstd,
                      const uint8_t* mask, const void* dres, void* dx, float* dscale, float dscale_beta,
                      float* partial, int64_t rows, int dim, cudaStream_t st) {
  const int iters = (dim + 255) / 256;
  const int nb = rms_bwd_blocks(rows);
  dim3 grid(nb), block(kRmsWarps * 32);
#define VG_RMS_BWD(I)                                                                       \
  rmsnorm_bwd_kernel<TX, TG, I><<<grid, block, 0, st>>>((const TG*)dy, (const TX*)x, scale, rstd, \
                                                        mask, (const TX*)dres, (TX*)dx, partial,  \
                                                        rows, dim)
  if (iters <= 1) VG_RMS_BWD(1);
  else if (iters <= 2) VG_RMS_BWD(2);
  else if (iters <= 4) VG_RMS_BWD(4);
  else VG_RMS_BWD(8);
#undef VG_RMS_BWD
  VG_LAUNCH_CHECK("vg_rmsnorm_bwd");
  colsum_partials_kernel<<<(dim + 31) / 32, 256, 0, st>>>(partial, dscale, nb, dim, dscale_beta);
  VG_LAUNCH_CHECK("vg_rmsnorm_bwd(reduce)");
  return 0;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_rmsnorm_fwd(const void* x, const float* scale, const uint8_t* row_mask, void* y,
                              float* rstd, int64_t rows, int64_t dim, float eps, int x_dtype,
                              int y_dtype, vg_stream_t stream) {
  VG_REQUIRE(x && scale && y && rstd, -1, "vg_rmsnorm_fwd: null pointer");
  VG_REQUIRE(valid_dtype(x_dtype) && valid_dtype(y_dtype), -2, "vg_rmsnorm_fwd: bad dtype");
  VG_REQUIRE(rows >= 0 && dim > 0 && dim % 8 == 0 && dim <= 256 * kRmsMaxIters, -3,
             "vg_rmsnorm_fwd: dim=%lld must be a multiple of 8 and <= %d", (long long)dim,
             256 * kRmsMaxIters);
  VG_REQUIRE(aligned(x, 16) && aligned(y, 16) && aligned(scale, 16), -4,
             "vg_rmsnorm_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == VG_F32 && y_dtype == VG_F32)
    return launch_fwd<float, float>(x, scale, row_mask, y, rstd, rows, (int)dim, eps, st);
  if (x_dtype == VG_BF16 && y_dtype == VG_BF16)
    return launch_fwd<__nv_bfloat16, __nv_bfloat16>(x, scale, row_mask, y, rstd, rows, (int)dim, eps, st);
  if (x_dtype == VG_BF16 && y_dtype == VG_F32)
    return launch_fwd<__nv_bfloat16, float>(x, scale, row_mask, y, rstd, rows, (int)dim, eps, st);
  return launch_fwd<float, __nv_bfloat16>(x, scale, row_mask, y, rstd, rows, (int)dim, eps, st);
}

extern "C" size_t vg_rmsnorm_bwd_workspace(int64_t rows, int64_t dim) {
  return (size_t)rms_bwd_blocks(rows > 0 ? rows : 1) * (size_t)dim * sizeof(float);
}

extern "C" int vg_rmsnorm_bwd(const void* dy, const void* x, const float* scale, const float* rstd,
                              const uint8_t* row_mask, const void* dres, void* dx, float* dscale,
                              float dscale_beta, void* workspace, size_t workspace_bytes, int64_t rows, int64_t dim,
                              int x_dtype, int dy_dtype, vg_stream_t stream) {
  VG_REQUIRE(dy && x && scale && rstd && dx && dscale, -1, "vg_rmsnorm_bwd: null pointer");
  VG_REQUIRE(valid_dtype(x_dtype) && valid_dtype(dy_dtype), -2, "vg_rmsnorm_bwd: bad dtype");
  VG_REQUIRE(rows > 0 && dim > 0 && dim % 8 == 0 && dim <= 256 * kRmsMaxIters, -3,
             "vg_rmsnorm_bwd: bad shape rows=%lld dim=%lld", (long long)rows, (long long)dim);
  VG_REQUIRE(aligned(x, 16) && aligned(dy, 16) && aligned(dx, 16) && aligned(scale, 16) &&
                 (!dres || aligned(dres, 16)), -4, "vg_rmsnorm_bwd: pointers must be 16-byte aligned");
  VG_REQUIRE(workspace && workspace_bytes >= vg_rmsnorm_bwd_workspace(rows, dim), -5,
             "vg_rmsnorm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  if (x_dtype == VG_F32 && dy_dtype == VG_F32)
    return launch_bwd<float, float>(dy, x, scale, rstd, row_mask, dres, dx, dscale, dscale_beta, partial, rows, (int)dim, st);
  if (x_dtype == VG_BF16 && dy_dtype == VG_BF16)
    return launch_bwd<__nv_bfloat16, __nv_bfloat16>(dy, x, scale, rstd, row_mask, dres, dx, dscale, dscale_beta, partial, rows, (int)dim, st);
  if (x_dtype == VG_BF16 && dy_dtype == VG_F32)
    return launch_bwd<__nv_bfloat16, float>(dy, x, scale, rstd, row_mask, dres, dx, dscale, dscale_beta, partial, rows, (int)dim, st);
  return launch_bwd<float, __nv_bfloat16>(dy, x, scale, rstd, row_mask, dres, dx, dscale, dscale_beta, partial, rows, (int)dim, st);
}
