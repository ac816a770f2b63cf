// elementwise.cu — the two elementwise passes the backward needs outside a GEMM epilogue.
//   vg_mask_rows : backward of TensorMask.apply_mask (utils/tensormask.py:63-67)
//   vg_act_bwd   : dx = dy * act'(src)  (ReLU from the saved output, GELU from the saved pre-activation)
// HBM-bound, 16-byte vector accesses, grid-stride free (one vector per thread).
#include "common.cuh"

namespace vg {

template <typename T>
__global__ void __launch_bounds__(256)
mask_rows_kernel(const T* __restrict__ x, const uint8_t* __restrict__ mask, T* __restrict__ y, int64_t rows, int cols8) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched kernel behind may start its prologue (vg_set_pdl_mode)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over rows * cols/8
  if (i >= rows * cols8) return;
  const int64_t r = i / cols8;
  Vec8<T> v;
  if (mask[r]) {
    v.load(x + i * 8);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
  }
  v.store(y + i * 8);
}

template <typename T>
__global__ void __launch_bounds__(256)
mask_rows_scalar_kernel(const T* __restrict__ x, const uint8_t* __restrict__ mask, T* __restrict__ y, int64_t rows, int cols) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  y[i] = mask[i / cols] ? x[i] : from_f32<T>(0.f);
}

template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ src, T* __restrict__ dx, int64_t n8, int act) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched kernel behind may start its prologue (vg_set_pdl_mode)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  Vec8<T> g, s;
  g.load(dy + i * 8);
  s.load(src + i * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) g.v[j] *= act_grad(s.v[j], act);
  g.store(dx + i * 8);
}

}  // namespace vg

using namespace vg;

extern "C" int vg_mask_rows(const void* x, const uint8_t* row_mask, void* y, int64_t rows, int64_t cols, int dtype,
                            vg_stream_t stream) {
  VG_REQUIRE(x && row_mask && y, -1, "vg_mask_rows: null pointer");
  VG_REQUIRE(valid_dtype(dtype), -2, "vg_mask_rows: bad dtype");
  VG_REQUIRE(rows > 0 && cols > 0, -3, "vg_mask_rows: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (cols % 8 != 0 || !aligned(x, 16) || !aligned(y, 16)) {      // narrow tensors (e.g. the 4-wide latent heads)
    const int64_t total = rows * cols;
    if (dtype == VG_F32)
      mask_rows_scalar_kernel<float><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>((const float*)x, row_mask, (float*)y, rows, (int)cols);
    else
      mask_rows_scalar_kernel<__nv_bfloat16><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
          (const __nv_bfloat16*)x, row_mask, (__nv_bfloat16*)y, rows, (int)cols);
    VG_LAUNCH_CHECK("vg_mask_rows(scalar)");
    return 0;
  }
  const int64_t n = rows * (cols / 8);
  if (dtype == VG_F32)
    mask_rows_kernel<float><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((const float*)x, row_mask, (float*)y, rows, (int)(cols / 8));
  else
    mask_rows_kernel<__nv_bfloat16><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((const __nv_bfloat16*)x, row_mask,
                                                                               (__nv_bfloat16*)y, rows, (int)(cols / 8));
  VG_LAUNCH_CHECK("vg_mask_rows");
  return 0;
}

extern "C" int vg_act_bwd(const void* dy, const void* src, void* dx, int64_t n, int act, int dtype, vg_stream_t stream) {
  VG_REQUIRE(dy && src && dx, -1, "vg_act_bwd: null pointer");
  VG_REQUIRE(valid_dtype(dtype), -2, "vg_act_bwd: bad dtype");
  VG_REQUIRE(n > 0 && n % 8 == 0, -3, "vg_act_bwd: n must be a positive multiple of 8");
  VG_REQUIRE(act >= VG_ACT_RELU && act <= VG_ACT_MULT, -3, "vg_act_bwd: activation must be ReLU, GELU, SiLU or MULT");
  VG_REQUIRE(aligned(dy, 16) && aligned(src, 16) && aligned(dx, 16), -4, "vg_act_bwd: unaligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VG_F32)
    act_bwd_kernel<float><<<(unsigned)ceil_div(n / 8, 256), 256, 0, st>>>((const float*)dy, (const float*)src, (float*)dx, n / 8, act);
  else
    act_bwd_kernel<__nv_bfloat16><<<(unsigned)ceil_div(n / 8, 256), 256, 0, st>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)src, (__nv_bfloat16*)dx, n / 8, act);
  VG_LAUNCH_CHECK("vg_act_bwd");
  return 0;
}
