// gemm_tc_kmn.cu — tcgen05 GEMM instantiations for A K-major / B MN-major operands (see gemm_tc_kernel.cuh).
#include "gemm_tc_kernel.cuh"

namespace vg {
int gemm_tc_launch_kmn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                       const TcEpilogue& epi, cudaStream_t st) {
  return launch_tc_layout<false, true, TCM_KMN>(bn, tmA, tmB, a, epi, st);
}
}  // namespace vg
