// skinny_linear.cu — y[B,N] = epilogue(x[B,K] · W[N,K]ᵀ) for B <= 256 rows: every nn.Linear of the cached generation
// step (LVTR.step on ONE new frame per sequence: attention.py:52,79, transformer/layers.py:82,152, lvtr.py:171,172,194,195).
//
// At B <= 256 the layer is a weight stream (2·N·K bytes read once) with a sliver of tensor work; what decides its time is
// (1) that ALL SMs pull a share of W and (2) how many bytes each SM must pull through its shared memory.  The general
// GEMM (gemm_tc_kernel, 128 x 64 output tiles of x·Wᵀ) makes every tile re-read a [128 x K] slice of the activations from
// L2 (36 MB of L2 → shared-memory traffic for the 6 MB QKV weight at B = 256) and spends 8–10 us per launch; the
// mma.sync weight streamer (decode_linear.cu) 13–22 us at 64 sequences (profiles/r02_decode.md).  Here:
//   * swap-AB: the WEIGHTS are the M = 128 operand of tcgen05.mma (a CTA owns 128 output features), the batch is the N
//     operand (BT = 64 / 128 / 256 columns), so one code path serves every batch size and no tensor lanes idle on padding;
//   * a thread-block CLUSTER of S CTAs splits the k-range of one (feature tile, batch tile): each CTA pulls only
//     (128 + BT) · K / S elements — the host picks (BT, S) per shape so that one wave of clusters covers the SMs with the
//     fewest bytes per SM; the S partial accumulators are reduce-scattered through DISTRIBUTED SHARED MEMORY (each CTA
//     ends up owning BT / S batch columns), so split-K costs no global atomics and no second kernel;
//   * TMA (SWIZZLE_128B boxes, out-of-range rows zero-filled) feeds the ring; launched with the programmatic-dependent-
//     launch attribute: barrier init, TMEM allocation and the W loads of the first ring stages go out BEFORE
//     griddepcontrol.wait — weights do not depend on the kernel in front — and only the activation loads wait;
//   * epilogue (bias, ReLU / GELU, row mask, residual) in registers, one thread per output feature: a warp's store of
//     one batch row is 32 consecutive features = one full 64- or 128-byte segment.
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace vg {
using namespace sm100;

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                      int box_outer);                       // gemm_tc.cu

constexpr int SK_THREADS = 192;            // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue
constexpr int SK_FT = 128;                 // output features per CTA (the M of the MMA)
constexpr int SK_W_BYTES = SK_FT * 64 * 2; // one 64-wide k-block of the weight tile
constexpr int SK_MAX_STAGES = 12;

struct SkParams {
  const float* bias;
  const void* residual;
  const uint8_t* row_mask;
  void* y;
  int64_t ld_res, ldy;
  int B, N, nkb, S, NS, FT;                // rows, features, k-blocks per CTA, cluster size, ring stages, feature tiles
  int act, y_f32, mask_first;
  const float* ss_in;                      // [B] row sums of squares of x: acc *= rsqrt(ss_in[b] * inv_k + eps) (folded RMSNorm)
  float* ss_out;                           // [B] += sum over features of y^2 (the stored, rounded value)
  float* zero_ss;                          // [B] cleared (the accumulator of a later launch)
  float inv_k, eps;
};

struct SkBars {
  uint64_t full[SK_MAX_STAGES], empty[SK_MAX_STAGES], acc_full;
  uint32_t tmem_base;
};

// Epilogue of NC (<= 32) batch rows b_first .. of one feature n held by this thread: every global read (residual, row mask)
// is issued before the first store — y and the residual may alias as far as the compiler knows, so a load placed after a
// store is not hoisted above it, and 32 dependent round trips per chunk were the whole cost of the first version of this
// kernel (profiles/r02_decode.md).
template <int NC>
__device__ __forceinline__ void sk_epilogue(const SkParams& p, int b_first, int n, const float (&acc)[NC], float bias) {
  // Called by whole warps (the row sums of squares are reduced with shuffles); lanes whose feature n is beyond N load from
  // a clamped address and store nothing.  Raw loads only in the first pass (rows clamped instead of branched around, no
  // conversion: a use of the loaded value here would wait for it and serialise the NC round trips again).
  const bool valid = n < p.N;
  const int nc = valid ? n : p.N - 1;
  uint32_t raw[NC];
  float ssv[NC];
  uint32_t keep_bits = 0xffffffffu;
  const int b_last = p.B - 1;
  if (p.residual) {
    if (p.y_f32) {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        raw[j] = __float_as_uint(reinterpret_cast<const float*>(p.residual)[(int64_t)min(b_first + j, b_last) * p.ld_res + nc]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        raw[j] = reinterpret_cast<const unsigned short*>(p.residual)[(int64_t)min(b_first + j, b_last) * p.ld_res + nc];
    }
  }
  if (p.ss_in) {
#pragma unroll
    for (int j = 0; j < NC; ++j) ssv[j] = p.ss_in[min(b_first + j, b_last)];
  }
  uint8_t mk[NC];
  if (p.row_mask) {
#pragma unroll
    for (int j = 0; j < NC; ++j) mk[j] = p.row_mask[min(b_first + j, b_last)];
#pragma unroll
    for (int j = 0; j < NC; ++j) if (!mk[j]) keep_bits &= ~(1u << j);
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int b = b_first + j;
    if (b > b_last) break;                                   // warp-uniform
    const bool keep = (keep_bits >> j) & 1u;
    float v = acc[j];
    if (p.ss_in) v *= rsqrtf(ssv[j] * p.inv_k + p.eps);
    v = apply_act_fast(v + bias, p.act);
    if (!keep && p.mask_first) v = 0.f;
    if (p.residual) v += p.y_f32 ? __uint_as_float(raw[j]) : __uint_as_float(raw[j] << 16);
    if (!keep && !p.mask_first) v = 0.f;
    float stored = v;
    if (p.y_f32) {
      if (valid) reinterpret_cast<float*>(p.y)[(int64_t)b * p.ldy + n] = v;
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      stored = __bfloat162float(h);
      if (valid) reinterpret_cast<__nv_bfloat16*>(p.y)[(int64_t)b * p.ldy + n] = h;
    }
    if (p.ss_out) {
      float sq = valid ? stored * stored : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(p.ss_out + b, sq);
    }
  }
}

template <int BT>
__global__ void __launch_bounds__(SK_THREADS, 1)
skinny_linear_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const SkParams p) {
  constexpr int X_BYTES = BT * 64 * 2;
  constexpr int STAGE = SK_W_BYTES + X_BYTES;
  extern __shared__ uint8_t sk_raw[];
  uint8_t* smem = sk_raw + ((1024u - (smem_u32(sk_raw) & 1023u)) & 1023u);        // SWIZZLE_128B atoms: 1024-byte aligned
  const int ring_bytes = max(p.NS * STAGE, BT * SK_FT * 4);                        // the ring doubles as the reduction buffer
  SkBars& bars = *reinterpret_cast<SkBars*>(smem + ring_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const int rank = S > 1 ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x / S;
  const int n0 = (cid % p.FT) * SK_FT;
  const int b0 = (cid / p.FT) * BT;
  const int kb0 = rank * p.nkb;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmX);
    for (int s = 0; s < p.NS; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars.tmem_base, BT);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== producer: W slabs before the dependency wait, X slabs after =====================
    const int npre = p.nkb < p.NS ? p.nkb : p.NS;
    if (elect_one()) {
      for (int i = 0; i < npre; ++i) {
        mbar_arrive_expect_tx(&bars.full[i], STAGE);
        tma_load_2d(smem + i * STAGE, &tmW, &bars.full[i], (kb0 + i) * 64, n0);
      }
    }
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (elect_one()) {
      for (int i = 0; i < npre; ++i) tma_load_2d(smem + i * STAGE + SK_W_BYTES, &tmX, &bars.full[i], (kb0 + i) * 64, b0);
    }
    __syncwarp();
    for (int i = npre; i < p.nkb; ++i) {
      const int s = i % p.NS;
      mbar_wait(&bars.empty[s], ((i / p.NS) - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars.full[s], STAGE);
        tma_load_2d(smem + s * STAGE, &tmW, &bars.full[s], (kb0 + i) * 64, n0);
        tma_load_2d(smem + s * STAGE + SK_W_BYTES, &tmX, &bars.full[s], (kb0 + i) * 64, b0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: D[128 features x BT rows] += W_slab[128 x 64] · X_slab[BT x 64]ᵀ =====================
    constexpr uint32_t idesc = make_idesc_bf16(SK_FT, BT, 0, 0);
    for (int i = 0; i < p.nkb; ++i) {
      const int s = i % p.NS;
      mbar_wait(&bars.full[s], (i / p.NS) & 1);
      tc_fence_after();
      const uint32_t w_addr = smem_u32(smem + s * STAGE);
      const uint32_t x_addr = w_addr + SK_W_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem, make_smem_desc_sw128(w_addr + k * 32, 16, 1024), make_smem_desc_sw128(x_addr + k * 32, 16, 1024),
                      idesc, (i | k) != 0);
        umma_commit(&bars.empty[s]);
        if (i == p.nkb - 1) umma_commit(&bars.acc_full);
      }
      __syncwarp();
    }
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");       // residual / row mask / row sums come from kernels in front
    if (p.zero_ss && blockIdx.x == 0)
      for (int i = (int)threadIdx.x - 64; i < p.B; i += SK_THREADS - 64) p.zero_ss[i] = 0.f;
  }

  // ===================== epilogue (warps 2..5; warps 0..1 only take part in the cluster barriers) =====================
  const int q = warp & 3;                                     // TMEM lane quadrant this warp may read
  const int f = q * 32 + lane;                                // feature row of this thread
  const int n = n0 + f;
  const bool epi = warp >= 2;
  float bias = 0.f;
  if (epi) {
    if (p.bias && n < p.N) bias = __ldg(p.bias + n);
    mbar_wait(&bars.acc_full, 0);
    tc_fence_after();
  }
  if (S == 1) {
    if (epi) {
#pragma unroll 1
      for (int c0 = 0; c0 < BT; c0 += 32) {
        if (b0 + c0 >= p.B) break;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        sk_epilogue<32>(p, b0 + c0, n, acc, bias);
      }
    }
  } else {
    // reduce-scatter through distributed shared memory: rank r owns batch columns [r CPR, (r + 1) CPR) of the tile.
    // red[((src_rank * CPR + c_local) / 4 * 128 + feature) * 4 + c_local % 4] (fp32) lives in the ring of the OWNER — a
    // thread sends four consecutive columns as one 16-byte store, a warp 512 contiguous bytes; the ring is free once every
    // CTA of the cluster has seen its own accumulator complete — the first cluster barrier.
    const int CPR = BT / S;                                   // a multiple of 8
    cluster_sync_all();
    if (epi) {
      const uint32_t red_addr = smem_u32(smem);
#pragma unroll 1
      for (int c0 = 0; c0 < BT; c0 += 32) {
        if (b0 + c0 >= p.B) break;                           // (CTA-uniform: rows beyond B are never stored)
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = c0 + j;
          const int owner = c / CPR, cl = c - owner * CPR;
          const uint32_t off = (uint32_t)((((rank * CPR + cl) >> 2) * SK_FT + f) * 16);
          if (owner == rank) {
            *reinterpret_cast<uint4*>(smem + off) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            const uint32_t dst = mapa_shared(red_addr + off, (uint32_t)owner);
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]),
                         "r"(v[j + 3]) : "memory");
          }
        }
      }
    }
    cluster_sync_all();
    if (epi) {
      const float4* red = reinterpret_cast<const float4*>(smem);
#pragma unroll 1
      for (int cl = 0; cl < CPR; cl += 8) {
        const int b = b0 + rank * CPR + cl;
        if (b >= p.B) break;
        float acc[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int r = 0; r < S; ++r) {
            const float4 t = red[((r * CPR + cl + 4 * h) >> 2) * SK_FT + f];
            sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
          }
          acc[4 * h] = sum.x; acc[4 * h + 1] = sum.y; acc[4 * h + 2] = sum.z; acc[4 * h + 3] = sum.w;
        }
        sk_epilogue<8>(p, b, n, acc, bias);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BT);
}

}  // namespace vg

using namespace vg;

typedef void (*SkKernel)(const CUtensorMap, const CUtensorMap, const SkParams);

static SkKernel sk_kernel(int bt) {
  return bt == 64 ? skinny_linear_kernel<64> : (bt == 128 ? skinny_linear_kernel<128> : skinny_linear_kernel<256>);
}

// ring depth and dynamic shared memory of one CTA for batch tile `bt` and `nkb` k-blocks per CTA
static void sk_smem(int bt, int nkb, int* ns_out, size_t* smem_out) {
  const int stage = SK_W_BYTES + bt * 128;
  int ns = (200 * 1024) / stage;
  if (ns > SK_MAX_STAGES) ns = SK_MAX_STAGES;
  if (ns > nkb) ns = nkb;
  const int ring = ns * stage > bt * SK_FT * 4 ? ns * stage : bt * SK_FT * 4;
  *ns_out = ns;
  *smem_out = 1024 + (size_t)ring + sizeof(SkBars) + 64;
}

static int sk_set_smem_attr(int bt) {
  static bool attr_set[3] = {false, false, false};
  const int ki = bt == 64 ? 0 : (bt == 128 ? 1 : 2);
  if (!attr_set[ki]) {
    VG_CUDA(cudaFuncSetAttribute(sk_kernel(bt), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[ki] = true;
  }
  return 0;
}

// How many clusters of `s` CTAs of this shape the device holds at once (a cluster needs its s SMs inside one GPC, so this
// is less than SMs / s: e.g. 16 clusters of 8 would need every GPC to give exactly two).  Cached per (bt, s, ring depth).
static int sk_resident_clusters(int bt, int s, int nkb, int sms) {
  static int cache[3][4][SK_MAX_STAGES + 1] = {};
  const int ki = bt == 64 ? 0 : (bt == 128 ? 1 : 2);
  const int si = s == 1 ? 0 : (s == 2 ? 1 : (s == 4 ? 2 : 3));
  int ns; size_t smem;
  sk_smem(bt, nkb, &ns, &smem);
  int& slot = cache[ki][si][ns];
  if (slot) return slot;
  int n = 0;
  if (sk_set_smem_attr(bt) == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(s * sms));
    cfg.blockDim = dim3(SK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)s;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)sk_kernel(bt), &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
  }
  if (n <= 0) n = s == 8 ? 14 : sms / s;                      // the query failed: a conservative guess
  slot = n;
  return n;
}

// (BT, S) for one shape: one wave of clusters over the SMs with the fewest bytes through each SM's shared memory;
// a split costs the distributed-shared-memory exchange of the partial tile.  Above 64 rows the wave count comes from
// the number of clusters the device really holds at once (sk_resident_clusters); the plans up to 64 rows were swept
// shape by shape on the device (profiles/r02_decode.md 3b) and keep the SM-count rule they were fitted with.
static void sk_plan(int64_t B, int64_t N, int64_t K, int sms, bool residency, int* bt_out, int* s_out) {
  const int64_t kb = K / 64;
  const int64_t ft = ceil_div(N, SK_FT);
  double best = 1e30;
  *bt_out = 64; *s_out = 1;
  for (int bt = 64; bt <= 256; bt *= 2) {
    if (bt > 64 && bt / 2 >= B) break;                        // a wider tile than the batch needs only adds padding
    const int64_t nbt = ceil_div(B, bt);
    for (int s = 1; s <= 8; s *= 2) {
      if (kb % s || bt / s < 8) continue;
      const int64_t ctas = ft * nbt * s;
      int64_t waves = ceil_div(ctas, sms);
      if (residency && s > 1) {
        const int64_t w2 = ceil_div(ft * nbt, sk_resident_clusters(bt, s, (int)(kb / s), sms));
        if (w2 > waves) waves = w2;
      }
      const double bytes = (double)(kb / s) * (SK_W_BYTES + bt * 128);
      const double cost = waves * (1.0 + bytes / 150e3) + (s > 1 ? 0.6 + bt * 512.0 / 200e3 : 0.0);     // microseconds
      if (cost < best) { best = cost; *bt_out = bt; *s_out = s; }
    }
  }
}

extern "C" int vg_skinny_linear(const vg_skinny_linear_args* a, vg_stream_t stream) {
  VG_REQUIRE(a && a->x && a->w && a->y, -1, "vg_skinny_linear: null pointer");
  VG_REQUIRE(a->B >= 1 && a->B <= 256 && a->N >= 8 && a->K >= 64 && a->K % 64 == 0, -3,
             "vg_skinny_linear: B=%lld (1..256), N=%lld, K=%lld (multiple of 64)", (long long)a->B, (long long)a->N,
             (long long)a->K);
  VG_REQUIRE(aligned(a->x, 16) && aligned(a->w, 16) && a->ldx % 8 == 0 && a->ldw % 8 == 0, -4, "vg_skinny_linear: unaligned");
  VG_REQUIRE(a->act >= VG_ACT_NONE && a->act <= VG_ACT_SILU, -3, "vg_skinny_linear: bad activation");
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    VG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int bt = 64, S = 1;
  static const int env_res = getenv("VG_SK_RESIDENCY") ? atoi(getenv("VG_SK_RESIDENCY")) : 1;
  sk_plan(a->B, a->N, a->K, sms, env_res && a->B > 64, &bt, &S);
  static const int env_bt = getenv("VG_SK_BT") ? atoi(getenv("VG_SK_BT")) : 0;
  static const int env_s = getenv("VG_SK_S") ? atoi(getenv("VG_SK_S")) : 0;
  if (env_bt) bt = env_bt;
  if (env_s && (a->K / 64) % env_s == 0) S = env_s;
  SkParams p;
  p.bias = a->bias; p.residual = a->residual; p.row_mask = a->row_mask; p.y = a->y;
  p.ld_res = a->ld_res; p.ldy = a->ldy;
  p.B = (int)a->B; p.N = (int)a->N; p.S = S; p.nkb = (int)(a->K / 64 / S);
  p.FT = (int)ceil_div(a->N, SK_FT);
  p.act = a->act; p.y_f32 = a->y_dtype == VG_F32; p.mask_first = a->mask_before_residual;
  p.ss_in = a->row_ss_in; p.ss_out = a->row_ss_out; p.zero_ss = a->zero_ss; p.inv_k = a->ss_inv_k; p.eps = a->ss_eps;
  int ns;
  size_t smem;
  sk_smem(bt, p.nkb, &ns, &smem);
  p.NS = ns;
  static const int env_dbg = getenv("VG_SK_DEBUG") ? atoi(getenv("VG_SK_DEBUG")) : 0;
  if (env_dbg) {
    static int printed = 0;
    if (printed < 64) {
      ++printed;
      fprintf(stderr, "[vg_skinny_linear] B=%lld N=%lld K=%lld -> bt=%d S=%d stages=%d smem=%zu resident clusters=%d (of %lld)\n",
              (long long)a->B, (long long)a->N, (long long)a->K, bt, S, ns, smem,
              S > 1 ? sk_resident_clusters(bt, S, p.nkb, sms) : sms, (long long)(ceil_div(a->N, SK_FT) * ceil_div(a->B, bt)));
    }
  }
  CUtensorMap tmW, tmX;
  int rc = make_tmap_bf16_2d(&tmW, a->w, a->K, a->N, a->ldw, 64, SK_FT);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmX, a->x, a->K, a->B, a->ldx, 64, bt);
  if (rc) return rc;
  SkKernel kern = sk_kernel(bt);
  if (sk_set_smem_attr(bt)) return -2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.FT * ceil_div(a->B, bt) * S));
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (S > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)S;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  VG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmW, tmX, p));
  return 0;
}
