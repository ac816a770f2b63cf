// attn_simt.cu — causal self-attention with in-kernel ALiBi bias and per-sequence key length, forward
// and backward, on CUDA cores with exact f32 softmax arithmetic.  Reference: attention.py:52-78
// (dense additive mask + F.scaled_dot_product_attention) and position/alibi.py:6-33.
//
// This is the PARITY backend (fp32 mode, 1e-4 tolerance) and the correctness baseline for the
// tensor-core attention.  Flash-style: the [B,H,T,T] mask/bias tensor of the reference is never
// materialised — bias = -slope_h * (i - j) and the validity test (j <= i, j < kv_len[b]) are
// computed in registers; memory is O(T).  Backward = delta pre-pass + a dK/dV kernel (loop over
// query tiles) + a dQ kernel (loop over key tiles); both recompute P from the saved LSE, so no float
// atomics are needed and results are run-to-run deterministic.
#include <math_constants.h>
#include "common.cuh"

namespace vg {

constexpr int AT = 64;        // tile edge (queries and keys per tile)
constexpr int AD = 64;        // head dim
constexpr int APAD = AT + 4;  // smem row pitch (floats)
constexpr int ATHREADS = 256;

typedef float Tile[AT][APAD];

// natural layout: dst[r][c] = src[(row0 + r) * ld + c], rows >= nrows → 0
template <typename T>
__device__ __forceinline__ void load_tile(Tile& dst, const T* __restrict__ src, int64_t ld, int row0, int nrows) {
  const int r = threadIdx.x / 4, c0 = (threadIdx.x % 4) * 16;
  const bool ok = (row0 + r) < nrows;
  const T* p = src + (int64_t)(row0 + r) * ld + c0;
#pragma unroll
  for (int e = 0; e < 16; ++e) dst[r][c0 + e] = ok ? to_f32<T>(p[e]) : 0.f;
}
// transposed layout: dst[c][r] = src[(row0 + r) * ld + c]
template <typename T>
__device__ __forceinline__ void load_tile_t(Tile& dst, const T* __restrict__ src, int64_t ld, int row0, int nrows) {
  const int r = threadIdx.x / 4, c0 = (threadIdx.x % 4) * 16;
  const bool ok = (row0 + r) < nrows;
  const T* p = src + (int64_t)(row0 + r) * ld + c0;
#pragma unroll
  for (int e = 0; e < 16; ++e) dst[c0 + e][r] = ok ? to_f32<T>(p[e]) : 0.f;
}

// out[i][j] = sum_d At[d][i0+i] * Bt[d][j0+j]   (both operands stored transposed: [d][row])
__device__ __forceinline__ void mma_tt(const Tile& At, const Tile& Bt, int i0, int j0, float (&out)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[i][j] = 0.f;
#pragma unroll 8
  for (int d = 0; d < AD; ++d) {
    const float4 a4 = *reinterpret_cast<const float4*>(&At[d][i0]);
    const float4 b4 = *reinterpret_cast<const float4*>(&Bt[d][j0]);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) out[i][j] = fmaf(a[i], b[j], out[i][j]);
  }
}
// acc[i][c] += sum_r A[i0+i][r] * Bn[r][c0+c]    (A natural [i][r], B natural [r][c])
__device__ __forceinline__ void mma_nn_acc(const Tile& A, const Tile& Bn, int i0, int c0, float (&acc)[4][4]) {
#pragma unroll 8
  for (int r = 0; r < AT; ++r) {
    const float4 b4 = *reinterpret_cast<const float4*>(&Bn[r][c0]);
    const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = A[i0 + i][r];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(a, b[c], acc[i][c]);
    }
  }
}
// acc[j][c] += sum_r A[r][j0+j] * Bn[r][c0+c]    (Aᵀ·B with A natural [r][j])
__device__ __forceinline__ void mma_tn_acc(const Tile& A, const Tile& Bn, int j0, int c0, float (&acc)[4][4]) {
#pragma unroll 8
  for (int r = 0; r < AT; ++r) {
    const float4 a4 = *reinterpret_cast<const float4*>(&A[r][j0]);
    const float4 b4 = *reinterpret_cast<const float4*>(&Bn[r][c0]);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = fmaf(a[j], b[c], acc[j][c]);
  }
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct AttnShape {
  int B, H, Tq, Tk, q_offset;
  int64_t ld_q, ld_kv, ld_out;
  float scale;
  int64_t kv_bs, kv_hs;     // K/V batch and head strides (elements): packed [B,T,H*D] or head-major cache
};

// ------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(ATHREADS)
attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ out,
                float* __restrict__ lse, const int32_t* __restrict__ kv_len, const float* __restrict__ slopes,
                AttnShape sh) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile& Qt = *reinterpret_cast<Tile*>(smem_raw);
  Tile& Kt = *reinterpret_cast<Tile*>(smem_raw + sizeof(Tile));
  Tile& Vs = *reinterpret_cast<Tile*>(smem_raw + 2 * sizeof(Tile));
  Tile& Ps = *reinterpret_cast<Tile*>(smem_raw + 3 * sizeof(Tile));

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  const int klen = kv_len ? min(kv_len[b], sh.Tk) : sh.Tk;
  const float slope = slopes ? slopes[h] : 0.f;

  const T* qb = q + (int64_t)b * sh.Tq * sh.ld_q + h * AD;
  const T* kb = k + (int64_t)b * sh.kv_bs + (int64_t)h * sh.kv_hs;
  const T* vb = v + (int64_t)b * sh.kv_bs + (int64_t)h * sh.kv_hs;
  load_tile_t<T>(Qt, qb, sh.ld_q, q0, sh.Tq);

  float m[4], l[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -CUDART_INF_F; l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[i][c] = 0.f;
  }
  const int q_abs_last = sh.q_offset + min(q0 + AT, sh.Tq) - 1;
  const int k_end = min(klen, q_abs_last + 1);
  for (int j0 = 0; j0 < k_end; j0 += AT) {
    __syncthreads();   // previous tile's consumers done (also orders the Qt load on the first pass)
    load_tile_t<T>(Kt, kb, sh.ld_kv, j0, sh.Tk);
    load_tile<T>(Vs, vb, sh.ld_kv, j0, sh.Tk);
    __syncthreads();
    float s[4][4];
    mma_tt(Qt, Kt, ty * 4, tx * 4, s);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ia = sh.q_offset + q0 + ty * 4 + i;
      float rmax = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ja = j0 + tx * 4 + j;
        const bool ok = (ja <= ia) && (ja < klen);
        s[i][j] = ok ? s[i][j] * sh.scale - slope * (float)(ia - ja) : -CUDART_INF_F;
        rmax = fmaxf(rmax, s[i][j]);
      }
      rmax = half_warp_max(rmax);
      const float mnew = fmaxf(m[i], rmax);
      const float corr = (mnew == -CUDART_INF_F) ? 1.f : expf(m[i] - mnew);
      float rsum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (s[i][j] == -CUDART_INF_F) ? 0.f : expf(s[i][j] - mnew);
        Ps[ty * 4 + i][tx * 4 + j] = p;
        rsum += p;
      }
      rsum = half_warp_sum(rsum);
      l[i] = l[i] * corr + rsum;
      m[i] = mnew;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[i][c] *= corr;
    }
    __syncthreads();
    mma_nn_acc(Ps, Vs, ty * 4, tx * 4, o);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int iq = q0 + ty * 4 + i;
    if (iq >= sh.Tq) continue;
    const bool valid = (sh.q_offset + iq) < klen && l[i] > 0.f;
    const float inv = valid ? 1.f / l[i] : 0.f;
    T* op = out + ((int64_t)b * sh.Tq + iq) * sh.ld_out + h * AD + tx * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) op[c] = from_f32<T>(o[i][c] * inv);
    if (tx == 0) lse[((int64_t)b * sh.H + h) * sh.Tq + iq] = valid ? m[i] + logf(l[i]) : 0.f;
  }
}

// ------------------------------------------------------------------------------------ backward
// delta[b,h,i] = sum_d dO[i,d] * O[i,d]
template <typename T>
__global__ void attn_delta_kernel(const T* __restrict__ dout, int64_t ld_dout, const T* __restrict__ out,
                                  int64_t ld_out, float* __restrict__ delta, int B, int H, int Tq) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * H * Tq) return;
  const int i = (int)(w % Tq);
  const int h = (int)((w / Tq) % H);
  const int b = (int)(w / ((int64_t)Tq * H));
  const T* dp = dout + ((int64_t)b * Tq + i) * ld_dout + h * AD;
  const T* op = out + ((int64_t)b * Tq + i) * ld_out + h * AD;
  float s = to_f32<T>(dp[lane]) * to_f32<T>(op[lane]) + to_f32<T>(dp[lane + 32]) * to_f32<T>(op[lane + 32]);
  s = warp_sum(s);
  if (lane == 0) delta[w] = s;
}

struct AttnBwdShape {
  AttnShape s;
  int64_t ld_dout, ld_dq, ld_dkv;
};

// P and dS for one (query tile, key tile) pair; thread owns rows ty*4.., cols tx*4..
// Requires Qt, Kt, dOt, Vt staged.  Writes P into Ps and dS into dSs (natural [i][j]).
__device__ __forceinline__ void attn_bwd_p_ds(const Tile& Qt, const Tile& Kt, const Tile& dOt, const Tile& Vt,
                                              Tile& Ps, Tile& dSs, const float* __restrict__ lse_row,
                                              const float* __restrict__ delta_row, int q0, int j0, int Tq,
                                              int q_offset, int klen, float scale, float slope) {
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  float s[4][4], dp[4][4];
  mma_tt(Qt, Kt, ty * 4, tx * 4, s);
  mma_tt(dOt, Vt, ty * 4, tx * 4, dp);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int iq = q0 + ty * 4 + i;
    const int ia = q_offset + iq;
    const bool row_ok = iq < Tq && ia < klen;
    const float L = row_ok ? lse_row[iq] : 0.f;
    const float dl = row_ok ? delta_row[iq] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ja = j0 + tx * 4 + j;
      const bool ok = row_ok && (ja <= ia) && (ja < klen);
      const float p = ok ? expf(s[i][j] * scale - slope * (float)(ia - ja) - L) : 0.f;
      Ps[ty * 4 + i][tx * 4 + j] = p;
      dSs[ty * 4 + i][tx * 4 + j] = p * (dp[i][j] - dl) * scale;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(ATHREADS)
attn_bwd_dkdv_kernel(const T* __restrict__ dout, const T* __restrict__ q, const T* __restrict__ k,
                     const T* __restrict__ v, const float* __restrict__ lse, const float* __restrict__ delta,
                     T* __restrict__ dk, T* __restrict__ dv, const int32_t* __restrict__ kv_len,
                     const float* __restrict__ slopes, AttnBwdShape bs) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile* tiles = reinterpret_cast<Tile*>(smem_raw);
  Tile &Qt = tiles[0], &Qs = tiles[1], &Kt = tiles[2], &Vt = tiles[3], &dOt = tiles[4], &dOs = tiles[5],
       &Ps = tiles[6], &dSs = tiles[7];
  const AttnShape& sh = bs.s;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * AT;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  const int klen = kv_len ? min(kv_len[b], sh.Tk) : sh.Tk;
  const float slope = slopes ? slopes[h] : 0.f;
  const T* qb = q + (int64_t)b * sh.Tq * sh.ld_q + h * AD;
  const T* kb = k + (int64_t)b * sh.kv_bs + (int64_t)h * sh.kv_hs;
  const T* vb = v + (int64_t)b * sh.kv_bs + (int64_t)h * sh.kv_hs;
  const T* dob = dout + (int64_t)b * sh.Tq * bs.ld_dout + h * AD;
  const float* lse_row = lse + ((int64_t)b * sh.H + h) * sh.Tq;
  const float* delta_row = delta + ((int64_t)b * sh.H + h) * sh.Tq;

  float acc_dk[4][4], acc_dv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) { acc_dk[i][c] = 0.f; acc_dv[i][c] = 0.f; }

  if (j0 < klen) {
    load_tile_t<T>(Kt, kb, sh.ld_kv, j0, sh.Tk);
    load_tile_t<T>(Vt, vb, sh.ld_kv, j0, sh.Tk);
    // first query tile that can see key j0: q_offset + iq >= j0
    int q_start = j0 - sh.q_offset;
    if (q_start < 0) q_start = 0;
    q_start = (q_start / AT) * AT;
    for (int q0 = q_start; q0 < sh.Tq; q0 += AT) {
      __syncthreads();
      load_tile_t<T>(Qt, qb, sh.ld_q, q0, sh.Tq);
      load_tile<T>(Qs, qb, sh.ld_q, q0, sh.Tq);
      load_tile_t<T>(dOt, dob, bs.ld_dout, q0, sh.Tq);
      load_tile<T>(dOs, dob, bs.ld_dout, q0, sh.Tq);
      __syncthreads();
      attn_bwd_p_ds(Qt, Kt, dOt, Vt, Ps, dSs, lse_row, delta_row, q0, j0, sh.Tq, sh.q_offset, klen, sh.scale, slope);
      __syncthreads();
      mma_tn_acc(Ps, dOs, ty * 4, tx * 4, acc_dv);    // dV[j][d] += sum_i P[i][j] dO[i][d]
      mma_tn_acc(dSs, Qs, ty * 4, tx * 4, acc_dk);    // dK[j][d] += sum_i dS[i][j] Q[i][d]
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int jr = j0 + ty * 4 + j;
    if (jr >= sh.Tk) continue;
    T* dkp = dk + ((int64_t)b * sh.Tk + jr) * bs.ld_dkv + h * AD + tx * 4;   // gradients are always packed
    T* dvp = dv + ((int64_t)b * sh.Tk + jr) * bs.ld_dkv + h * AD + tx * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      dkp[c] = from_f32<T>(acc_dk[j][c]);
      dvp[c] = from_f32<T>(acc_dv[j][c]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(ATHREADS)
attn_bwd_dq_kernel(const T* __restrict__ dout, const T* __restrict__ q, const T* __restrict__ k,
                   const T* __restrict__ v, const float* __restrict__ lse, const float* __restrict__ delta,
                   T* __restrict__ dq, const int32_t* __restrict__ kv_len, const float* __restrict__ slopes,
                   AttnBwdShape bs) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tile* tiles = reinterpret_cast<Tile*>(smem_raw);
  Tile &Qt = tiles[0], &Kt = tiles[1], &Ks = tiles[2], &Vt = tiles[3], &dOt = tiles[4], &Ps = tiles[5],
       &dSs = tiles[6];
  const AttnShape& sh = bs.s;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  const int klen = kv_len ? min(kv_len[b], sh.Tk) : sh.Tk;
  const float slope = slopes ? slopes[h] : 0.f;
  const T* qb = q + (int64_t)b * sh.Tq * sh.ld_q + h * AD;
  const T* kb = k + (int64_t)b * sh.kv_bs + (int64_t)h * sh.kv_hs;
  const T* vb = v + (int64_t)b * sh.kv_bs + (int64_t)h * sh.kv_hs;
  const T* dob = dout + (int64_t)b * sh.Tq * bs.ld_dout + h * AD;
  const float* lse_row = lse + ((int64_t)b * sh.H + h) * sh.Tq;
  const float* delta_row = delta + ((int64_t)b * sh.H + h) * sh.Tq;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
  load_tile_t<T>(Qt, qb, sh.ld_q, q0, sh.Tq);
  load_tile_t<T>(dOt, dob, bs.ld_dout, q0, sh.Tq);
  const int q_abs_last = sh.q_offset + min(q0 + AT, sh.Tq) - 1;
  const int k_end = min(klen, q_abs_last + 1);
  for (int j0 = 0; j0 < k_end; j0 += AT) {
    __syncthreads();
    load_tile_t<T>(Kt, kb, sh.ld_kv, j0, sh.Tk);
    load_tile<T>(Ks, kb, sh.ld_kv, j0, sh.Tk);
    load_tile_t<T>(Vt, vb, sh.ld_kv, j0, sh.Tk);
    __syncthreads();
    attn_bwd_p_ds(Qt, Kt, dOt, Vt, Ps, dSs, lse_row, delta_row, q0, j0, sh.Tq, sh.q_offset, klen, sh.scale, slope);
    __syncthreads();
    mma_nn_acc(dSs, Ks, ty * 4, tx * 4, acc);         // dQ[i][d] += sum_j dS[i][j] K[j][d]
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int iq = q0 + ty * 4 + i;
    if (iq >= sh.Tq) continue;
    T* p = dq + ((int64_t)b * sh.Tq + iq) * bs.ld_dq + h * AD + tx * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) p[c] = from_f32<T>(acc[i][c]);
  }
}

template <typename T>
static int attn_fwd_launch(const void* q, const void* k, const void* v, void* out, float* lse,
                           const int32_t* kv_len, const float* slopes, const AttnShape& sh, cudaStream_t st) {
  auto kern = attn_fwd_kernel<T>;
  const int smem = 4 * sizeof(Tile);
  static bool set = false;
  if (!set) { VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
  dim3 grid((unsigned)ceil_div(sh.Tq, AT), sh.H, sh.B);
  kern<<<grid, ATHREADS, smem, st>>>((const T*)q, (const T*)k, (const T*)v, (T*)out, lse, kv_len, slopes, sh);
  VG_LAUNCH_CHECK("vg_attn_fwd");
  return 0;
}

template <typename T>
static int attn_bwd_launch(const void* dout, const void* q, const void* k, const void* v, const void* out,
                           const float* lse, void* dq, void* dk, void* dv, const int32_t* kv_len,
                           const float* slopes, const AttnBwdShape& bs, float* delta, cudaStream_t st) {
  const AttnShape& sh = bs.s;
  const int64_t nrows = (int64_t)sh.B * sh.H * sh.Tq;
  attn_delta_kernel<T><<<(unsigned)ceil_div(nrows * 32, 256), 256, 0, st>>>((const T*)dout, bs.ld_dout, (const T*)out,
                                                                           sh.ld_out, delta, sh.B, sh.H, sh.Tq);
  VG_LAUNCH_CHECK("vg_attn_bwd(delta)");
  {
    auto kern = attn_bwd_dkdv_kernel<T>;
    const int smem = 8 * sizeof(Tile);
    static bool set = false;
    if (!set) { VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    dim3 grid((unsigned)ceil_div(sh.Tk, AT), sh.H, sh.B);
    kern<<<grid, ATHREADS, smem, st>>>((const T*)dout, (const T*)q, (const T*)k, (const T*)v, lse, delta, (T*)dk,
                                       (T*)dv, kv_len, slopes, bs);
    VG_LAUNCH_CHECK("vg_attn_bwd(dkdv)");
  }
  {
    auto kern = attn_bwd_dq_kernel<T>;
    const int smem = 7 * sizeof(Tile);
    static bool set = false;
    if (!set) { VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
    dim3 grid((unsigned)ceil_div(sh.Tq, AT), sh.H, sh.B);
    kern<<<grid, ATHREADS, smem, st>>>((const T*)dout, (const T*)q, (const T*)k, (const T*)v, lse, delta, (T*)dq,
                                       kv_len, slopes, bs);
    VG_LAUNCH_CHECK("vg_attn_bwd(dq)");
  }
  return 0;
}

}  // namespace vg

namespace vg {
void attn_tc_set_trace(void* buf);
bool attn_tc_supported(int dtype, int64_t D, int64_t ld_q, int64_t ld_kv, const void* q, const void* k, const void* v,
                       int64_t kv_batch_stride, int64_t kv_head_stride);
int attn_tc_fwd_launch(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out,
                       int64_t ld_out, float* lse, const int32_t* kv_len, const float* slopes, int B, int H, int Tq,
                       int Tk, int q_offset, float scale, cudaStream_t st);
size_t attn_tc_bwd_workspace(int64_t B, int64_t H, int64_t Tq);
int attn_tc_bwd_launch(const void* dout, int64_t ld_dout, const void* q, const void* k, const void* v, int64_t ld_q,
                       int64_t ld_kv, const void* out, int64_t ld_out, const float* lse, void* dq, void* dk, void* dv,
                       int64_t ld_dq, int64_t ld_dkv, const int32_t* kv_len, const float* slopes, int B, int H, int Tq,
                       int Tk, int q_offset, float scale, void* workspace, cudaStream_t st);
static int g_attn_backend = 0;     // 0 = auto (tcgen05 for bf16), 1 = CUDA-core parity kernels, 2 = tcgen05 required
}  // namespace vg

using namespace vg;

extern "C" int vg_debug_attn_trace(void* buf) {
  attn_tc_set_trace(buf);
  return 0;
}

extern "C" int vg_set_attn_backend(int backend) {
  VG_REQUIRE(backend >= 0 && backend <= 2, -3, "vg_set_attn_backend: backend must be 0 (auto), 1 (simt) or 2 (tcgen05)");
  g_attn_backend = backend;
  return 0;
}

static int check_attn_shape(const char* fn, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D,
                            int64_t q_offset, int dtype) {
  VG_REQUIRE(valid_dtype(dtype), -2, "%s: bad dtype", fn);
  VG_REQUIRE(D == AD, -3, "%s: head dim %lld unsupported (only 64)", fn, (long long)D);
  VG_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0 && B < 65536 && H < 65536, -3, "%s: bad shape", fn);
  VG_REQUIRE(q_offset >= 0 && q_offset + Tq <= Tk, -3, "%s: q_offset + Tq must be <= Tk", fn);
  return 0;
}

extern "C" int vg_attn_fwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv, void* out,
                           int64_t ld_out, float* lse, const int32_t* kv_len, const float* slopes, int64_t B,
                           int64_t H, int64_t Tq, int64_t Tk, int64_t D, int64_t q_offset, int64_t kv_batch_stride,
                           int64_t kv_head_stride, float scale, int dtype, vg_stream_t stream) {
  VG_REQUIRE(q && k && v && out && lse, -1, "vg_attn_fwd: null pointer");
  if (int rc = check_attn_shape("vg_attn_fwd", B, H, Tq, Tk, D, q_offset, dtype)) return rc;
  VG_REQUIRE(ld_q >= H * D && ld_kv >= D && ld_out >= H * D, -3, "vg_attn_fwd: row stride too small");
  AttnShape sh{(int)B, (int)H, (int)Tq, (int)Tk, (int)q_offset, ld_q, ld_kv, ld_out, scale,
               kv_batch_stride > 0 ? kv_batch_stride : Tk * ld_kv, kv_head_stride > 0 ? kv_head_stride : D};
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = attn_tc_supported(dtype, D, ld_q, ld_kv, q, k, v, kv_batch_stride, kv_head_stride) &&
                     ld_out % 8 == 0 && aligned(out, 16);
  VG_REQUIRE(g_attn_backend != 2 || tc_ok, -6, "vg_attn_fwd: tcgen05 backend needs packed bf16 q/k/v with ld %% 8 == 0");
  if (tc_ok && g_attn_backend != 1)
    return attn_tc_fwd_launch(q, k, v, ld_q, ld_kv, out, ld_out, lse, kv_len, slopes, (int)B, (int)H, (int)Tq, (int)Tk,
                              (int)q_offset, scale, st);
  if (dtype == VG_F32) return attn_fwd_launch<float>(q, k, v, out, lse, kv_len, slopes, sh, st);
  return attn_fwd_launch<__nv_bfloat16>(q, k, v, out, lse, kv_len, slopes, sh, st);
}

extern "C" size_t vg_attn_bwd_workspace(int64_t B, int64_t H, int64_t Tq, int64_t, int64_t) {
  return attn_tc_bwd_workspace(B, H, Tq);      // delta [B,H,Tq] (+ the tcgen05 path's fp32 dQ accumulator)
}

extern "C" int vg_attn_bwd(const void* dout, int64_t ld_dout, const void* q, const void* k, const void* v,
                           int64_t ld_q, int64_t ld_kv, const void* out, int64_t ld_out, const float* lse, void* dq,
                           void* dk, void* dv, int64_t ld_dq, int64_t ld_dkv, const int32_t* kv_len,
                           const float* slopes, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D,
                           int64_t q_offset, float scale, int dtype, void* workspace, size_t workspace_bytes,
                           vg_stream_t stream) {
  VG_REQUIRE(dout && q && k && v && out && lse && dq && dk && dv, -1, "vg_attn_bwd: null pointer");
  if (int rc = check_attn_shape("vg_attn_bwd", B, H, Tq, Tk, D, q_offset, dtype)) return rc;
  VG_REQUIRE(workspace && workspace_bytes >= vg_attn_bwd_workspace(B, H, Tq, Tk, D), -5,
             "vg_attn_bwd: workspace too small");
  AttnBwdShape bs{{(int)B, (int)H, (int)Tq, (int)Tk, (int)q_offset, ld_q, ld_kv, ld_out, scale, Tk * ld_kv, D},
                  ld_dout, ld_dq, ld_dkv};
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = attn_tc_supported(dtype, D, ld_q, ld_kv, q, k, v, 0, 0) && ld_dout % 8 == 0 && ld_out % 8 == 0 &&
                     ld_dq % 8 == 0 && ld_dkv % 8 == 0 && aligned(dout, 16) && aligned(out, 16) && aligned(dq, 16) &&
                     aligned(dk, 16) && aligned(dv, 16);
  VG_REQUIRE(g_attn_backend != 2 || tc_ok, -6, "vg_attn_bwd: tcgen05 backend needs packed bf16 tensors with ld %% 8 == 0");
  if (tc_ok && g_attn_backend != 1)
    return attn_tc_bwd_launch(dout, ld_dout, q, k, v, ld_q, ld_kv, out, ld_out, lse, dq, dk, dv, ld_dq, ld_dkv, kv_len,
                              slopes, (int)B, (int)H, (int)Tq, (int)Tk, (int)q_offset, scale, workspace, st);
  if (dtype == VG_F32)
    return attn_bwd_launch<float>(dout, q, k, v, out, lse, dq, dk, dv, kv_len, slopes, bs, (float*)workspace, st);
  return attn_bwd_launch<__nv_bfloat16>(dout, q, k, v, out, lse, dq, dk, dv, kv_len, slopes, bs, (float*)workspace, st);
}
