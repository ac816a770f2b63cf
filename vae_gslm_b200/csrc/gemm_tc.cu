// gemm_tc.cu — bf16 GEMM on the 5th-generation tensor cores: TMA → 128B-swizzled shared memory →
// tcgen05.mma (accumulators in TMEM) → tcgen05.ld epilogue with the fused vg_gemm epilogue.
//
// Replaces the cuBLAS calls behind every nn.Linear on the path (attention.py:52,79;
// transformer/layers.py:82,152; lvtr.py:171,172,194,195), their dgrad (dX = dY·W) and wgrad
// (dW = dYᵀ·X).  All four operand storage combinations are native: a K-major operand is loaded as
// one [rows x 64] box, an MN-major operand (dgrad's W, both wgrad operands) as [64 k-rows x 64] boxes,
// and the UMMA descriptors/instruction descriptor carry the major-ness — nothing is transposed in HBM.
//
// Kernel shape: persistent, one CTA per SM, 256 threads:
//   warp 0 lane 0 : TMA producer           (smem ring of kStages {A,B} tiles, full/empty mbarriers)
//   warp 1 lane 0 : tcgen05.mma issuer     (128 x BN x 16 per instruction, 4 per 64-wide k-block)
//   warp 2        : TMEM allocator         (2 accumulator buffers of BN columns → epilogue overlap)
//   warps 4..11   : epilogue               (thread = accumulator row; two warps per TMEM lane quadrant split the
//                                           columns; tcgen05.ld 32 columns at a time; fast-erf GELU)
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "sm100.cuh"

namespace vg {
using namespace sm100;

constexpr int TBM = 128;            // tile rows  (UMMA M)
constexpr int TBK = 64;             // k-block: 64 bf16 = 128 B = one swizzle atom
constexpr int TC_THREADS = 384;
constexpr int TC_EPI_THREADS = 256;         // warps 4..11
constexpr int A_TILE_BYTES = TBM * TBK * 2;   // 16 KB

template <int BN> struct TcCfg {
  static constexpr int kBBytes = BN * TBK * 2;
  static constexpr int kStageBytes = A_TILE_BYTES + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* align slack */ + 256 /* barriers */;
};

struct TcEpilogue {
  EpilogueParams p;
  int vec_ok;   // every pointer/ld allows 8-wide vector access
};

template <typename TC>
__device__ __forceinline__ void tc_epilogue_chunk(const TcEpilogue& e, int m, int n0, int N, const uint32_t* r) {
  const EpilogueParams& ep = e.p;
  const bool keep = !(ep.row_mask && !ep.row_mask[m]);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n0 + g * 8;
    if (n >= N) break;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
    if (e.vec_ok && n + 8 <= N) {
      if (ep.bias) {
        Vec8<float> b;
        b.load(ep.bias + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += b.v[j];
      }
      if (ep.act != VG_ACT_NONE || ep.preact) {
        Vec8<TC> p;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y, dy;
          act_and_grad_fast(v[j], ep.act, y, dy);
          p.v[j] = ep.preact_is_grad ? dy : v[j];
          v[j] = y;
        }
        if (ep.preact) p.store(reinterpret_cast<TC*>(ep.preact) + (int64_t)m * ep.ld_preact + n);
      }
      if (ep.dact_src) {
        Vec8<TC> d;
        d.load(reinterpret_cast<const TC*>(ep.dact_src) + (int64_t)m * ep.ld_dact + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= act_grad_fast(d.v[j], ep.dact);
      }
      if (!keep && ep.mask_first) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      if (ep.residual) {
        Vec8<TC> d;
        d.load(reinterpret_cast<const TC*>(ep.residual) + (int64_t)m * ep.ld_res + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += d.v[j];
      }
      if (!keep && !ep.mask_first) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      TC* c = reinterpret_cast<TC*>(ep.C) + (int64_t)m * ep.ldc + n;
      Vec8<TC> o;
      if (ep.beta != 0.f) {
        o.load(c);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += ep.beta * o.v[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = v[j];
      o.store(c);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n + j < N) epilogue_store<TC>(ep, m, n + j, v[j]);
    }
  }
}

template <int BN, bool A_MN, bool B_MN, typename TC>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, TcEpilogue epi) {
  using Cfg = TcCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);   // 1024 B alignment for SW128

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], TC_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + TBM - 1) / TBM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + TBK - 1) / TBK;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * TBM;
      const int n0 = (tile / num_m) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
        uint8_t* a_dst = smem + s * Cfg::kStageBytes;
        uint8_t* b_dst = a_dst + A_TILE_BYTES;
        const int k0 = kb * TBK;
        if constexpr (!A_MN) {
          tma_load_2d(a_dst, &tmA, &full_bar[s], k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < TBM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
        }
        if constexpr (!B_MN) {
          tma_load_2d(b_dst, &tmB, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
        }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(TBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < TBK / 16; ++k) {
          // K-major  : rows at 128 B pitch, 8-row groups every 1024 B (SBO); +32 B per 16-wide k step.
          // MN-major : 64-element MN chunks every 8192 B (LBO), 8 k-rows per 1024 B (SBO); +2048 B per k step.
          const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);            // frees the smem slot once these MMAs retire
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      umma_commit(&tmem_full[acc]);            // accumulator complete → epilogue
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int wq = warp & 3;                   // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * TBM;
      const int n0 = (tile / num_m) * BN;
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const int row = m0 + wq * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = (warp >= 8 ? BN / 64 : 0); c < (warp >= 8 ? BN / 32 : BN / 64); ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (row < M && n0 + c * 32 < N) tc_epilogue_chunk<TC>(epi, row, n0 + c * 32, N, r);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

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

bool gemm_tc_supported(const vg_gemm_args* a) {
  if (a->ab_dtype != VG_BF16) return false;
  if (a->M < 1 || a->N < 8 || a->K < 8) return false;
  if (!aligned(a->A, 16) || !aligned(a->B, 16)) return false;
  if (a->lda % 8 != 0 || a->ldb % 8 != 0) return false;
  // the contiguous extent of each operand must cover its logical inner dimension
  return true;
}

static int pick_bn(const vg_gemm_args* a) {
  if (a->N <= 64) return 64;
  const int64_t tiles256 = ceil_div(a->M, TBM) * ceil_div(a->N, 256);
  if (a->N >= 256 && tiles256 >= kNumSMs) return 256;
  return 128;
}

template <int BN, bool A_MN, bool B_MN, typename TC>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a, const TcEpilogue& epi,
                     cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, TC>;
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = ceil_div(a->M, TBM) * ceil_div(a->N, BN);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  kern<<<grid, TC_THREADS, Cfg::kSmemBytes, st>>>(tmA, tmB, (int)a->M, (int)a->N, (int)a->K, epi);
  VG_LAUNCH_CHECK("vg_gemm(tcgen05)");
  return 0;
}

template <int BN, typename TC>
static int dispatch_major(const CUtensorMap& tmA, const CUtensorMap& tmB, const vg_gemm_args* a,
                          const TcEpilogue& epi, cudaStream_t st) {
  const bool a_mn = a->trans_a != 0;      // stored [K,M] → M contiguous
  const bool b_mn = a->trans_b == 0;      // stored [K,N] → N contiguous
  if (!a_mn && !b_mn) return launch_tc<BN, false, false, TC>(tmA, tmB, a, epi, st);
  if (!a_mn && b_mn) return launch_tc<BN, false, true, TC>(tmA, tmB, a, epi, st);
  if (a_mn && !b_mn) return launch_tc<BN, true, false, TC>(tmA, tmB, a, epi, st);
  return launch_tc<BN, true, true, TC>(tmA, tmB, a, epi, st);
}

int gemm_tc_launch(const vg_gemm_args* a, cudaStream_t st) {
  const int bn = pick_bn(a);
  CUtensorMap tmA, tmB;
  int rc;
  if (!a->trans_a) rc = make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, TBK, TBM);
  else             rc = make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, 64, TBK);
  if (rc) return rc;
  if (a->trans_b)  rc = make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, TBK, bn);
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

  if (a->c_dtype == VG_BF16) {
    if (bn == 256) return dispatch_major<256, __nv_bfloat16>(tmA, tmB, a, epi, st);
    if (bn == 128) return dispatch_major<128, __nv_bfloat16>(tmA, tmB, a, epi, st);
    return dispatch_major<64, __nv_bfloat16>(tmA, tmB, a, epi, st);
  }
  if (bn == 256) return dispatch_major<256, float>(tmA, tmB, a, epi, st);
  if (bn == 128) return dispatch_major<128, float>(tmA, tmB, a, epi, st);
  return dispatch_major<64, float>(tmA, tmB, a, epi, st);
}

}  // namespace vg
