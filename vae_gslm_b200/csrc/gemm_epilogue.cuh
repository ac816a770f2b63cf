// gemm_epilogue.cuh — the vg_gemm epilogue contract (include/vgslm.h) shared by both GEMM backends.
#pragma once
#include "common.cuh"

namespace vg {

struct EpilogueParams {
  void* C; int64_t ldc;
  const float* bias;
  int act, dact;
  void* preact; int64_t ld_preact;
  const void* dact_src; int64_t ld_dact;
  const void* residual; int64_t ld_res;
  const uint8_t* row_mask;
  int mask_first;
  int preact_is_grad;
  float beta;
};

static inline EpilogueParams make_epilogue(const vg_gemm_args* a) {
  EpilogueParams ep;
  ep.C = a->C; ep.ldc = a->ldc;
  ep.bias = a->bias;
  ep.act = a->act; ep.dact = a->dact;
  ep.preact = a->preact; ep.ld_preact = a->ld_preact;
  ep.dact_src = a->dact_src; ep.ld_dact = a->ld_dact;
  ep.residual = a->residual; ep.ld_res = a->ld_res;
  ep.row_mask = a->row_mask;
  ep.mask_first = a->mask_before_residual;
  ep.preact_is_grad = a->preact_is_grad;
  ep.beta = a->beta;
  return ep;
}

// scalar form: one accumulator element → C[m,n]
template <typename TC>
__device__ __forceinline__ void epilogue_store(const EpilogueParams& ep, int m, int n, float v) {
  if (ep.bias) v += ep.bias[n];
  if (ep.preact)
    reinterpret_cast<TC*>(ep.preact)[(int64_t)m * ep.ld_preact + n] =
        from_f32<TC>(ep.preact_is_grad ? act_grad(v, ep.act) : v);
  v = apply_act(v, ep.act);
  if (ep.dact_src)
    v *= act_grad(to_f32<TC>(reinterpret_cast<const TC*>(ep.dact_src)[(int64_t)m * ep.ld_dact + n]), ep.dact);
  const bool drop = ep.row_mask && !ep.row_mask[m];
  if (drop && ep.mask_first) v = 0.f;
  if (ep.residual) v += to_f32<TC>(reinterpret_cast<const TC*>(ep.residual)[(int64_t)m * ep.ld_res + n]);
  if (drop && !ep.mask_first) v = 0.f;
  TC* c = reinterpret_cast<TC*>(ep.C) + (int64_t)m * ep.ldc + n;
  if (ep.beta != 0.f) v += ep.beta * to_f32<TC>(*c);
  *c = from_f32<TC>(v);
}

}  // namespace vg
