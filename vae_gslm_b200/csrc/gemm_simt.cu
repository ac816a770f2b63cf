// gemm_simt.cu — CUDA-core GEMM with the full vg_gemm epilogue.  This is the fp32 PARITY backend
// (exact fp32 FMA accumulation, used to hold the 1e-4 tolerance against the fp32 oracle) and the
// fallback for shapes the tcgen05 backend rejects (tiny K/N, unaligned leading dimensions).
// The bf16 throughput backend is gemm_tc.cu.
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace vg {

constexpr int SBM = 64, SBN = 64, SBK = 16, STHREADS = 256;

template <typename TAB, typename TC>
__global__ void __launch_bounds__(STHREADS)
gemm_simt_kernel(const TAB* __restrict__ A, int64_t sam, int64_t sak,
                 const TAB* __restrict__ B, int64_t sbk, int64_t sbn,
                 int M, int N, int K, EpilogueParams ep) {
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Bs[SBK][SBN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int ty = t / 16, tx = t % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kcontig = (sak == 1);
  const bool b_ncontig = (sbn == 1);
  for (int k0 = 0; k0 < K; k0 += SBK) {
    // ---- stage A tile (64 x 16) into As[k][m]
    if (a_kcontig) {
      const int m = t / 4, kk = (t % 4) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gm = m0 + m, gk = k0 + kk + j;
        As[kk + j][m] = (gm < M && gk < K) ? to_f32<TAB>(A[(int64_t)gm * sam + (int64_t)gk * sak]) : 0.f;
      }
    } else {
      const int kk = t / 16, m = (t % 16) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gm = m0 + m + j, gk = k0 + kk;
        As[kk][m + j] = (gm < M && gk < K) ? to_f32<TAB>(A[(int64_t)gm * sam + (int64_t)gk * sak]) : 0.f;
      }
    }
    // ---- stage B tile (16 x 64) into Bs[k][n]
    if (b_ncontig) {
      const int kk = t / 16, n = (t % 16) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gn = n0 + n + j, gk = k0 + kk;
        Bs[kk][n + j] = (gn < N && gk < K) ? to_f32<TAB>(B[(int64_t)gk * sbk + (int64_t)gn * sbn]) : 0.f;
      }
    } else {
      const int n = t / 4, kk = (t % 4) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gn = n0 + n, gk = k0 + kk + j;
        Bs[kk + j][n] = (gn < N && gk < K) ? to_f32<TAB>(B[(int64_t)gk * sbk + (int64_t)gn * sbn]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) epilogue_store<TC>(ep, gm, gn, acc[i][j]);
    }
  }
}

int gemm_simt_launch(const vg_gemm_args* a, cudaStream_t st) {
  const int64_t sam = a->trans_a ? 1 : a->lda, sak = a->trans_a ? a->lda : 1;
  const int64_t sbk = a->trans_b ? 1 : a->ldb, sbn = a->trans_b ? a->ldb : 1;
  EpilogueParams ep = make_epilogue(a);
  dim3 grid((unsigned)ceil_div(a->N, SBN), (unsigned)ceil_div(a->M, SBM)), block(STHREADS);
  const int M = (int)a->M, N = (int)a->N, K = (int)a->K;
  if (a->ab_dtype == VG_F32 && a->c_dtype == VG_F32)
    gemm_simt_kernel<float, float><<<grid, block, 0, st>>>((const float*)a->A, sam, sak, (const float*)a->B, sbk, sbn, M, N, K, ep);
  else if (a->ab_dtype == VG_BF16 && a->c_dtype == VG_BF16)
    gemm_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)a->A, sam, sak, (const __nv_bfloat16*)a->B, sbk, sbn, M, N, K, ep);
  else if (a->ab_dtype == VG_BF16 && a->c_dtype == VG_F32)
    gemm_simt_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>((const __nv_bfloat16*)a->A, sam, sak, (const __nv_bfloat16*)a->B, sbk, sbn, M, N, K, ep);
  else
    gemm_simt_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>((const float*)a->A, sam, sak, (const float*)a->B, sbk, sbn, M, N, K, ep);
  VG_LAUNCH_CHECK("vg_gemm(simt)");
  return 0;
}

// ---- column sums (bias gradients) ----------------------------------------------------------------
// stage 1: each thread owns 8 consecutive columns (one 16-byte load per row) and walks a strip of rows;
// stage 2: fixed-order sum of the strip partials.  HBM-bound: the matrix is read exactly once.
constexpr int kColsumRowsPerBlock = 32;

template <typename T>
__global__ void __launch_bounds__(128)
colsum_stage1(const T* __restrict__ x, int64_t ld, float* __restrict__ partial, int64_t rows, int cols) {
  const int c0 = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (c0 >= cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * kColsumRowsPerBlock;
  const int64_t r1 = r0 + kColsumRowsPerBlock < rows ? r0 + kColsumRowsPerBlock : rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bool vec = (c0 + 8 <= cols) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    // 8 independent 16-byte loads in flight per thread (the row loop is otherwise latency-bound), fixed add order
    int64_t r = r0;
    for (; r + 8 <= r1; r += 8) {
      Vec8<T> v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u].load(x + (r + u) * ld + c0);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u].v[j];
    }
    for (; r < r1; ++r) {
      Vec8<T> v;
      v.load(x + r * ld + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
    }
  } else {
    for (int64_t r = r0; r < r1; ++r)
      for (int j = 0; j < 8 && c0 + j < cols; ++j) acc[j] += to_f32<T>(x[r * ld + c0 + j]);
  }
  for (int j = 0; j < 8 && c0 + j < cols; ++j) partial[(int64_t)blockIdx.y * cols + c0 + j] = acc[j];
}
__global__ void __launch_bounds__(256)
colsum_stage2(const float* __restrict__ partial, float* __restrict__ out, int nparts, int cols, float beta) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cols)
    for (int p = ty; p < nparts; p += 8) s += partial[(int64_t)p * cols + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][tx];
    out[c] = beta != 0.f ? fmaf(beta, out[c], t) : t;
  }
}

}  // namespace vg

using namespace vg;

extern "C" size_t vg_colsum_workspace(int64_t rows, int64_t cols) {
  return (size_t)ceil_div(rows > 0 ? rows : 1, kColsumRowsPerBlock) * (size_t)cols * sizeof(float);
}

extern "C" int vg_colsum(const void* x, int64_t ld, float* out, int64_t rows, int64_t cols,
                         int x_dtype, float beta, void* workspace, size_t workspace_bytes, vg_stream_t stream) {
  VG_REQUIRE(x && out, -1, "vg_colsum: null pointer");
  VG_REQUIRE(valid_dtype(x_dtype), -2, "vg_colsum: bad dtype");
  VG_REQUIRE(rows > 0 && cols > 0 && ld >= cols, -3, "vg_colsum: bad shape");
  VG_REQUIRE(workspace && workspace_bytes >= vg_colsum_workspace(rows, cols), -5,
             "vg_colsum: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nparts = (int)ceil_div(rows, kColsumRowsPerBlock);
  dim3 grid((unsigned)ceil_div(cols, 1024), nparts), block(128);
  if (x_dtype == VG_F32)
    colsum_stage1<float><<<grid, block, 0, st>>>((const float*)x, ld, (float*)workspace, rows, (int)cols);
  else
    colsum_stage1<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, ld, (float*)workspace, rows, (int)cols);
  VG_LAUNCH_CHECK("vg_colsum(stage1)");
  colsum_stage2<<<(unsigned)ceil_div(cols, 32), 256, 0, st>>>((const float*)workspace, out, nparts, (int)cols, beta);
  VG_LAUNCH_CHECK("vg_colsum(stage2)");
  return 0;
}
