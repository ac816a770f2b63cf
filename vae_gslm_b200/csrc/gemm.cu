// gemm.cu — vg_gemm entry point: argument validation and backend selection (include/vgslm.h).
#include "common.cuh"

namespace vg {
int gemm_simt_launch(const vg_gemm_args* a, cudaStream_t st);
int gemm_tc_launch(const vg_gemm_args* a, cudaStream_t st);
bool gemm_tc_supported(const vg_gemm_args* a);
}  // namespace vg

using namespace vg;

extern "C" size_t vg_gemm_workspace(const vg_gemm_args*, int) { return 0; }

extern "C" int vg_gemm(const vg_gemm_args* a, int backend, void* /*workspace*/, size_t /*workspace_bytes*/,
                       vg_stream_t stream) {
  VG_REQUIRE(a != nullptr, -1, "vg_gemm: null args");
  VG_REQUIRE(a->A && a->B && a->C, -1, "vg_gemm: null operand pointer");
  VG_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, -3, "vg_gemm: bad shape M=%lld N=%lld K=%lld",
             (long long)a->M, (long long)a->N, (long long)a->K);
  VG_REQUIRE(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), -3, "vg_gemm: dimension too large");
  VG_REQUIRE(valid_dtype(a->ab_dtype) && valid_dtype(a->c_dtype), -2, "vg_gemm: bad dtype");
  VG_REQUIRE(a->lda >= (a->trans_a ? a->M : a->K), -3, "vg_gemm: lda=%lld too small", (long long)a->lda);
  VG_REQUIRE(a->ldb >= (a->trans_b ? a->K : a->N), -3, "vg_gemm: ldb=%lld too small", (long long)a->ldb);
  VG_REQUIRE(a->ldc >= a->N, -3, "vg_gemm: ldc too small");
  VG_REQUIRE(a->beta == 0.f || a->beta == 1.f, -3, "vg_gemm: beta must be 0 or 1");
  VG_REQUIRE(a->beta == 0.f || a->c_dtype == VG_F32, -3, "vg_gemm: beta=1 requires an f32 C");
  VG_REQUIRE(a->act >= VG_ACT_NONE && a->act <= VG_ACT_SILU && a->dact >= VG_ACT_NONE && a->dact <= VG_ACT_MULT,
             -3, "vg_gemm: bad activation id");
  VG_REQUIRE(!a->preact || a->ld_preact >= a->N, -3, "vg_gemm: ld_preact too small");
  VG_REQUIRE(!a->dact_src || a->ld_dact >= a->N, -3, "vg_gemm: ld_dact too small");
  VG_REQUIRE(!a->residual || a->ld_res >= a->N, -3, "vg_gemm: ld_res too small");
  if (a->M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == VG_GEMM_SIMT) return gemm_simt_launch(a, st);
  if (backend == VG_GEMM_TCGEN05) {
    VG_REQUIRE(gemm_tc_supported(a), -6,
               "vg_gemm: tcgen05 backend needs bf16 operands, 16-byte aligned bases and ld %% 8 == 0");
    return gemm_tc_launch(a, st);
  }
  VG_REQUIRE(backend == VG_GEMM_AUTO, -3, "vg_gemm: unknown backend %d", backend);
  if (gemm_tc_supported(a)) return gemm_tc_launch(a, st);
  return gemm_simt_launch(a, st);
}
