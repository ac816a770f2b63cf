// gemm_tc_mnk_f32.cu — tcgen05 GEMM instantiations: A MN-major, B K-major, f32 C (see gemm_tc_kernel.cuh).
#include "gemm_tc_kernel.cuh"

namespace vg {
int gemm_tc_launch_mnk_f32(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                              const TcEpilogue& epi, cudaStream_t st) {
  return launch_tc_layout<true, false, TCM_MNK, float>(bn, tmA, tmB, a, epi, st);
}
}  // namespace vg
