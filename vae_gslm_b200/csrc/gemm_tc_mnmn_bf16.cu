// gemm_tc_mnmn_bf16.cu — tcgen05 GEMM instantiations: A MN-major, B MN-major, bf16 C (see gemm_tc_kernel.cuh).
#include "gemm_tc_kernel.cuh"

namespace vg {
int gemm_tc_launch_mnmn_bf16(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st) {
  return launch_tc_layout<true, true, TCM_MNMN, __nv_bfloat16>(bn, tmA, tmB, a, epi, st);
}
}  // namespace vg
