// gemm_tc.cu — bf16 GEMM on the 5th-generation tensor cores: TMA → 128B-swizzled shared memory →
// tcgen05.mma (accumulators in TMEM) → tcgen05.ld epilogue with the fused vg_gemm epilogue.
//
// Replaces the cuBLAS calls behind every nn.Linear on the path (attention.py:52,79;
// transformer/layers.py:82,152; lvtr.py:171,172,194,195), their dgrad (dX = dY·W) and wgrad
// (dW = dYᵀ·X).  All four operand storage combinations are native: a K-major operand is loaded as
// one [rows x 64] box, an MN-major operand (dgrad's W, both wgrad operands) as [64 k-rows x 64] boxes,
// and the UMMA descriptors/instruction descriptor carry the major-ness — nothing is transposed in HBM.
//
// Kernel shape: persistent, one CTA per SM, 640 threads:
//   warp 0 lane 0 : TMA producer           (smem ring of kStages {A,B} tiles, full/empty mbarriers)
//   warp 1 lane 0 : tcgen05.mma issuer     (128 x BN x 16 per instruction, 4 per 64-wide k-block)
//   warp 2        : TMEM allocator         (2 accumulator buffers of BN columns → epilogue overlap)
//   warps 4..19   : epilogue               (thread = accumulator row; four warps per TMEM lane quadrant split the
//                                           columns; tcgen05.ld 32 columns at a time; packed-f32x2 erf GELU)
#include "gemm_tc_kernel.cuh"
#include <stdlib.h>

namespace vg {

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map: inner dimension `inner` (contiguous), `outer` rows of pitch ld elements.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld,
                      int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  VG_REQUIRE(fn != nullptr, -20, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_REQUIRE(r == CUDA_SUCCESS, -21, "cuTensorMapEncodeTiled failed (CUresult %d) inner=%lld outer=%lld ld=%lld",
             (int)r, (long long)inner, (long long)outer, (long long)ld);
  return 0;
}

// 3-D bf16 tensor map (e.g. [B, T, H*D] activations): box = box0 x box1 x 1, SWIZZLE_128B.
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                      int64_t stride2, int box0, int box1) {
  EncodeTiledFn fn = get_encode_fn();
  VG_REQUIRE(fn != nullptr, -20, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1 * 2, (cuuint64_t)stride2 * 2};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_REQUIRE(r == CUDA_SUCCESS, -21, "cuTensorMapEncodeTiled(3d) failed (CUresult %d) dims=%lld,%lld,%lld", (int)r,
             (long long)d0, (long long)d1, (long long)d2);
  return 0;
}

// 3-D fp32 tensor map for bulk reduce-add of [box1 x 32] tiles (128-byte rows, SWIZZLE_128B): the fp32 dQ accumulator
// of the attention backward.  Strides in elements.
int make_tmap_f32_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t stride1,
                     int64_t stride2, int box0, int box1) {
  EncodeTiledFn fn = get_encode_fn();
  VG_REQUIRE(fn != nullptr, -20, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1 * 4, (cuuint64_t)stride2 * 4};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_REQUIRE(r == CUDA_SUCCESS, -21, "cuTensorMapEncodeTiled(3d f32) failed (CUresult %d) dims=%lld,%lld,%lld", (int)r,
             (long long)d0, (long long)d1, (long long)d2);
  return 0;
}

bool gemm_tc_supported(const vg_gemm_args* a) {
  if (a->ab_dtype != VG_BF16) return false;
  if (a->M < 1 || a->N < 8 || a->K < 8) return false;
  if (!aligned(a->A, 16) || !aligned(a->B, 16)) return false;
  if (a->lda % 8 != 0 || a->ldb % 8 != 0) return false;
  // the contiguous extent of each operand must cover its logical inner dimension
  return true;
}


static int g_sm_budget = 0;
int gemm_sm_budget() {
  if (g_sm_budget == 0) {
    const char* e = getenv("VG_GEMM_SMS");
    int v = e ? atoi(e) : kNumSMs;
    g_sm_budget = (v >= 16 && v <= kNumSMs) ? (v & ~1) : kNumSMs;
  }
  return g_sm_budget;
}

// ---- tile / split-K selection -----------------------------------------------------------------------
// A work unit is (128 x BN output tile, k-range).  Cost model in microseconds, fitted to per-shape sweeps on B200
// (profiles/r01_gemm_splitk_sweep.md): one 64-deep k-block of a 128x256 tile costs ~0.50 us, of a 128x128 tile
// ~0.36 us (half the MMA work but the same A traffic through shared memory), of a 128x64 tile ~0.30 us; the epilogue
// of a unit hides behind the next unit's main loop (two TMEM accumulators).  Split-K is only available when the
// epilogue is a pure f32 accumulation (wgrad into the gradient arena): partial tiles are reduced with
// red.global.add, which costs ~2.5 us per unit of L2 atomic traffic — so it only pays when the tile count alone
// leaves SMs idle (e.g. 32 tiles of the 1024x1024 out_proj wgrad on 148 SMs).
struct TcPlan { int bn, splits, pair; };

static bool splitk_eligible(const vg_gemm_args* a) {
  return a->c_dtype == VG_F32 && !a->bias && a->act == VG_ACT_NONE && !a->preact && !a->dact_src && !a->residual &&
         !a->row_mask;
}

static TcPlan pick_plan(const vg_gemm_args* a) {
  static const int env_bn = getenv("VG_GEMM_BN") ? atoi(getenv("VG_GEMM_BN")) : 0;
  static const int env_splits = getenv("VG_GEMM_SPLITS") ? atoi(getenv("VG_GEMM_SPLITS")) : 0;
  const int64_t num_kb = ceil_div(a->K, TBK);
  const bool can_split = splitk_eligible(a);
  static const int env_pair = getenv("VG_GEMM_PAIR") ? atoi(getenv("VG_GEMM_PAIR")) : 1;
  if (a->N <= 64) return TcPlan{64, 1, 0};
  if (a->M <= 256 && !env_bn && !env_splits) {
    // Skinny problems (the cached generation step at batch <= 256: M = batch): HBM-bound on the weights, so what matters
    // is that ALL SMs stream a share of W.  128 x 64 tiles give the most units; when the epilogue is a pure f32
    // accumulation (out-proj / FFN2 reducing into the fp32 residual stream) the k-range is split until the grid is full
    // (>= 2 k-blocks per unit).  Measured per shape in profiles/r02_decode.md.
    const int bn = a->N >= 512 ? 64 : 128;
    const int64_t tiles = ceil_div(a->M, TBM) * ceil_div(a->N, bn);
    int s = 1;
    if (can_split)
      while (tiles * s * 2 <= gemm_sm_budget() && num_kb / (s * 2) >= 2) s *= 2;
    return TcPlan{bn, s, 0};
  }
  TcPlan best{128, 1, 0};
  double best_cost = 1e30;
  for (int bn = 128; bn <= 256; bn *= 2) {
    if (bn == 256 && a->N < 256) continue;
    if (env_bn && bn != env_bn && !(env_bn == 256 && a->N < 256)) continue;
    const double per_kb = bn == 256 ? 0.50 : 0.36;
    const int64_t tiles = ceil_div(a->M, TBM) * ceil_div(a->N, bn);
    const int max_splits = can_split ? 16 : 1;
    for (int s = 1; s <= max_splits; ++s) {
      if (env_splits && can_split && s != env_splits) continue;
      const int64_t kbpu = ceil_div(num_kb, s);
      if (s > 1 && kbpu < 8) break;
      const int64_t waves = ceil_div(tiles * s, gemm_sm_budget());
      const double cost = (double)waves * ((double)kbpu * per_kb + (s > 1 ? 2.5 : 0.0));
      if (cost < best_cost * 0.97) { best_cost = cost; best = TcPlan{bn, s, 0}; }   // ties → fewer splits / smaller BN
    }
  }
  // CTA pairs (cta_group::2, 256 x 256 tiles): same tile count per SM as 128 x 256, half the B traffic per SM
  if (env_pair && best.bn == 256 && a->M >= 256) best.pair = 1;
  return best;
}

int gemm_tc_launch(const vg_gemm_args* a, cudaStream_t st) {
  const TcPlan plan = pick_plan(a);
  const int bn = plan.bn;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a->trans_a) rc = make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, TBK, TBM);
  else             rc = make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, 64, TBK);
  if (rc) return rc;
  if (a->trans_b)  rc = make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, TBK, plan.pair ? bn / 2 : bn);
  else             rc = make_tmap_bf16_2d(&tmB, a->B, a->N, a->K, a->ldb, 64, TBK);
  if (rc) return rc;

  TcEpilogue epi;
  epi.p = make_epilogue(a);
  const int64_t vec = (a->c_dtype == VG_BF16) ? 8 : 4;   // elements per 16 B
  bool ok = aligned(a->C, 16) && a->ldc % vec == 0;
  if (a->bias) ok = ok && aligned(a->bias, 16);
  if (a->preact) ok = ok && aligned(a->preact, 16) && a->ld_preact % vec == 0;
  if (a->dact_src) ok = ok && aligned(a->dact_src, 16) && a->ld_dact % vec == 0;
  if (a->residual) ok = ok && aligned(a->residual, 16) && a->ld_res % vec == 0;
  epi.vec_ok = ok ? 1 : 0;
  epi.splits = plan.splits;
  if (plan.splits > 1) {
    epi.variant = TCV_RED;
    if (a->beta == 0.f) {           // partial sums accumulate into C: clear it first (enqueue-only, graph-capturable)
      if (a->ldc == a->N) VG_CUDA(cudaMemsetAsync(a->C, 0, (size_t)a->M * a->N * 4, st));
      else VG_CUDA(cudaMemset2DAsync(a->C, (size_t)a->ldc * 4, 0, (size_t)a->N * 4, (size_t)a->M, st));
    }
  } else if (a->dact_src) {
    epi.variant = (a->dact == VG_ACT_MULT && a->act == VG_ACT_NONE && !a->preact) ? TCV_MULT : TCV_GENERIC;
  } else {
    epi.variant = a->act == VG_ACT_NONE ? TCV_PLAIN : a->act == VG_ACT_RELU ? TCV_RELU
                  : a->act == VG_ACT_GELU ? TCV_GELU : a->act == VG_ACT_SILU ? TCV_SILU : TCV_GENERIC;
  }

  const bool a_mn = a->trans_a != 0;      // stored [K,M] → M contiguous
  const bool b_mn = a->trans_b == 0;      // stored [K,N] → N contiguous
  const int cfg = plan.pair ? 512 : bn;   // 512 selects the CTA-pair instantiation (256 x 256 tiles)
  const bool f32 = a->c_dtype == VG_F32;
  if (!a_mn && !b_mn) return f32 ? gemm_tc_launch_kk_f32(cfg, tmA, tmB, a, epi, st) : gemm_tc_launch_kk_bf16(cfg, tmA, tmB, a, epi, st);
  if (!a_mn && b_mn) return f32 ? gemm_tc_launch_kmn_f32(cfg, tmA, tmB, a, epi, st) : gemm_tc_launch_kmn_bf16(cfg, tmA, tmB, a, epi, st);
  if (a_mn && !b_mn) return f32 ? gemm_tc_launch_mnk_f32(cfg, tmA, tmB, a, epi, st) : gemm_tc_launch_mnk_bf16(cfg, tmA, tmB, a, epi, st);
  return f32 ? gemm_tc_launch_mnmn_f32(cfg, tmA, tmB, a, epi, st) : gemm_tc_launch_mnmn_bf16(cfg, tmA, tmB, a, epi, st);
}

}  // namespace vg

extern "C" int vg_set_gemm_sm_budget(int sms) {
  VG_REQUIRE(sms >= 16 && sms <= vg::kNumSMs, -3, "vg_set_gemm_sm_budget: %d not in [16, %d]", sms, vg::kNumSMs);
  vg::g_sm_budget = sms & ~1;
  return 0;
}
