/*
 * vgslm.h — C ABI of libvgslm.so, the B200 (sm_100a) kernel library behind the VAE-GSLM hot path.
 *
 * The reference (b04901014/vae-gslm) has no FFI layer of its own: its hot path is a graph of
 * PyTorch nn.Modules.  Every entry point below therefore replaces the torch-library call sequence
 * of ONE reference call site, cited as `file:line` (paths relative to the reference root).
 *
 * Conventions:
 *   - plain pointers and sizes only; every buffer (inputs, outputs, saved-for-backward, workspace,
 *     KV cache) is owned by the caller; nothing is allocated, freed or retained by the library;
 *   - every function only ENQUEUES work on `stream` (no device/host synchronisation) and is
 *     therefore CUDA-graph capturable;
 *   - return value: 0 = ok, <0 = argument/shape/alignment error detected before launch,
 *     >0 = cudaError_t.  `vg_last_error_string()` holds the message (thread-local);
 *   - there is no CPU fallback: an unsupported shape is an error, never a silent dispatch;
 *   - `dtype` arguments use vg_dtype; "act" tensors are the activation stream (f32 in parity mode,
 *     bf16 in bf16 mode); statistics, losses and parameter gradients are always f32.
 */
#ifndef VGSLM_H_
#define VGSLM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vg_stream_t;

typedef enum { VG_F32 = 0, VG_BF16 = 1 } vg_dtype;
/* VG_ACT_MULT is valid for `dact` only: the saved tensor already holds act'(pre) and is simply multiplied in */
typedef enum { VG_ACT_NONE = 0, VG_ACT_RELU = 1, VG_ACT_GELU = 2, VG_ACT_SILU = 3, VG_ACT_MULT = 4 } vg_act;
typedef enum { VG_GEMM_AUTO = 0, VG_GEMM_SIMT = 1, VG_GEMM_TCGEN05 = 2 } vg_gemm_backend;

#define VG_VERSION 100

/* ---- library ------------------------------------------------------------------------------ */
int         vg_version(void);
const char* vg_last_error_string(void);
/* 1 if the running device is compute capability 10.x (tcgen05/TMEM present), 0 otherwise, <0 error */
int         vg_device_is_sm100(void);

/* ---- RMSNorm: modules/norm.py:22-32 (+ the apply_mask that follows it, transformer/layers.py:53-54)
 * y[r,:] = mask[r] ? scale * x[r,:] * rsqrt(mean(x[r,:]^2) + eps) : 0 ; rstd[r] saved for backward. */
int vg_rmsnorm_fwd(const void* x, const float* scale, const uint8_t* row_mask /* nullable */,
                   void* y, float* rstd, int64_t rows, int64_t dim, float eps,
                   int x_dtype, int y_dtype, vg_stream_t stream);
size_t vg_rmsnorm_bwd_workspace(int64_t rows, int64_t dim);
/* dx = (dres ? dres : 0) + d/dx  (dres: the gradient arriving through the residual branch that bypasses the norm,
 * transformer/layers.py:57,63 — fused here instead of a separate add);
 * dscale[dim] = dscale_beta * dscale + sum_rows(...)  (deterministic two-stage reduction; beta = 1 accumulates
 * straight into a gradient buffer). */
int vg_rmsnorm_bwd(const void* dy, const void* x, const float* scale, const float* rstd,
                   const uint8_t* row_mask /* nullable */, const void* dres /* nullable */,
                   void* dx, float* dscale, float dscale_beta, void* workspace, size_t workspace_bytes,
                   int64_t rows, int64_t dim, int x_dtype, int dy_dtype, vg_stream_t stream);

/* ---- GEMM with fused epilogue: every nn.Linear on the path
 * (attention.py:52,79; transformer/layers.py:82,152; lvtr.py:171,172,194,195; flow FiLM linear/layers.py:286)
 *   acc = op(A)[M,K] · op(B)[K,N]
 *   v   = acc + bias[n]                     (bias nullable)
 *   if preact: preact[m,n] = v              (saved pre-activation, for GELU backward)
 *   v   = act(v)
 *   if dact_src: v *= act'(dact_src[m,n])   (dgrad epilogue: dact_src holds the saved pre-activation)
 *   v  += residual[m,n]                     (nullable)
 *   if row_mask && !row_mask[m]: v = 0      (TensorMask.apply_mask, utils/tensormask.py:63-67; with
 *                                            mask_before_residual=1 the mask is applied BEFORE the
 *                                            residual add: x + mask(y), transformer/layers.py:63)
 *   C[m,n] = beta * C[m,n] + v              (beta in {0,1}; beta=1 needs f32 C: gradient accumulation)
 * Storage: trans_a=0 → A is [M,K] row-major (ld = lda); trans_a=1 → A is stored [K,M] row-major.
 *          trans_b=1 → B is stored [N,K] row-major (an nn.Linear weight); trans_b=0 → stored [K,N].
 */
typedef struct {
  int64_t M, N, K;
  const void* A; int64_t lda; int32_t trans_a;
  const void* B; int64_t ldb; int32_t trans_b;
  void* C; int64_t ldc;
  int32_t ab_dtype;     /* vg_dtype of A and B */
  int32_t c_dtype;      /* vg_dtype of C, preact, residual, dact_src */
  const float* bias;
  int32_t act;          /* vg_act applied after bias */
  int32_t dact;         /* vg_act whose derivative multiplies the result (with dact_src) */
  void* preact; int64_t ld_preact;
  const void* dact_src; int64_t ld_dact;
  const void* residual; int64_t ld_res;
  const uint8_t* row_mask;
  int32_t mask_before_residual;
  int32_t preact_is_grad;   /* 1: `preact` receives act'(v) instead of v, so that backward is a plain multiply
                               (VG_ACT_MULT) — the forward epilogue already has erf/exp of v in registers */
  float beta;
} vg_gemm_args;
size_t vg_gemm_workspace(const vg_gemm_args* a, int backend);
int    vg_gemm(const vg_gemm_args* a, int backend, void* workspace, size_t workspace_bytes,
               vg_stream_t stream);
/* SMs the persistent tcgen05 GEMM grids may occupy (default 148 = all; env VG_GEMM_SMS).  Data-parallel training leaves
 * a few SMs to the NCCL all-reduce kernels that overlap backward (scripts/train.py:93-95 → DDP). */
int    vg_set_gemm_sm_budget(int sms);
/* Programmatic dependent launch for chains of short kernels (the layer-by-layer cached generation step, LVTR.step at
 * batch >= 64: ~7 launches per layer of lvtr.py:253-257 / transformer/layers.py:134-195).  bit 0: vg_gemm (tcgen05
 * path) and vg_rmsnorm_fwd are launched with the programmatic-stream-serialization attribute — their prologue (barrier
 * init, TMEM allocation, tensor-map prefetch) overlaps the tail of the kernel in front and they execute
 * griddepcontrol.wait before touching memory; bit 1: the B operand of vg_gemm is a static weight and its first ring
 * stages are fetched BEFORE the wait.  Process-wide; 0 (default) = plain stream order. */
int    vg_set_pdl_mode(int mode);
/* column sums: out[n] = beta * out[n] + sum_m X[m,n] (bias gradients), deterministic; beta in {0,1}. */
size_t vg_colsum_workspace(int64_t rows, int64_t cols);
int    vg_colsum(const void* x, int64_t ld, float* out, int64_t rows, int64_t cols, int x_dtype, float beta,
                 void* workspace, size_t workspace_bytes, vg_stream_t stream);

/* ---- small elementwise helpers used by the backward passes
 * vg_mask_rows: y[m,:] = mask[m] ? x[m,:] : 0           (backward of apply_mask; in-place allowed)
 * vg_act_bwd  : dx = dy * act'(src) where src is the saved pre-activation (GELU) or the output (ReLU) */
int vg_mask_rows(const void* x, const uint8_t* row_mask, void* y, int64_t rows, int64_t cols, int dtype,
                 vg_stream_t stream);
int vg_act_bwd(const void* dy, const void* src, void* dx, int64_t n, int act, int dtype, vg_stream_t stream);

/* ---- causal self-attention with in-kernel ALiBi and per-sequence kv length:
 * attention.py:52-78 (mask build + F.scaled_dot_product_attention) and position/alibi.py:6-33.
 * q,k,v: [B, T*, H*D] views with row stride ld_q / ld_kv (elements) so they may alias a packed qkv
 * buffer; K/V may instead be a head-major cache [B,H,Tmax,D] (ld_kv = D, kv_head_stride = Tmax*D,
 * kv_batch_stride = H*Tmax*D).  Query row iq sits at absolute position q_offset + iq
 * (q_offset = Tk - Tq with a KV cache).
 * score[b,h,i,j] = scale * q_i·k_j - slopes[h] * (i - j)   for j <= i and j < kv_len[b], else -inf.
 * Rows with iq >= q_len[b] (padded queries) are written as zeros (the reference masks them after
 * out_proj, attention.py:80).  lse[B,H,Tq] is saved for backward.                                   */
int vg_attn_fwd(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv,
                void* out, int64_t ld_out, float* lse,
                const int32_t* kv_len /* [B] nullable */, const float* slopes /* [H] nullable */,
                int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D, int64_t q_offset,
                int64_t kv_batch_stride /* 0 → Tk*ld_kv (packed) */, int64_t kv_head_stride /* 0 → D */,
                float scale, int dtype, vg_stream_t stream);
/* backend selection for vg_attn_fwd/bwd: 0 = auto (tcgen05 kernels for packed bf16, CUDA-core kernels otherwise),
 * 1 = CUDA-core (fp32-exact softmax) kernels only, 2 = tcgen05 required (error if the layout does not qualify) */
int vg_set_attn_backend(int backend);
size_t vg_attn_bwd_workspace(int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D);
int vg_attn_bwd(const void* dout, int64_t ld_dout,
                const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_kv,
                const void* out, int64_t ld_out, const float* lse,
                void* dq, void* dk, void* dv, int64_t ld_dq, int64_t ld_dkv,
                const int32_t* kv_len, const float* slopes,
                int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t D, int64_t q_offset,
                float scale, int dtype, void* workspace, size_t workspace_bytes, vg_stream_t stream);

/* ---- KV-cache decode attention (single query per sequence): attention.py:56-85 in the cached
 * LVTR.step loop (lvtr.py:253-257).  Cache layout: k_cache/v_cache [B, H, Tmax, D] (head-major).
 * The new token's k/v (read from the packed qkv row of this step) are appended at index `pos`
 * and attended together with cache rows [0,pos).                                                  */
size_t vg_attn_decode_workspace(int64_t B, int64_t H, int64_t D, int64_t splits);
int vg_attn_decode(const void* qkv /* [B, 3*H*D] */, void* k_cache, void* v_cache,
                   void* out /* [B, H*D] */, const float* slopes,
                   int64_t B, int64_t H, int64_t D, int64_t Tmax, int64_t pos,
                   const int32_t* pos_dev /* nullable: device-resident position (CUDA-graph replay) */,
                   int64_t splits, float scale, int dtype, void* workspace, size_t workspace_bytes,
                   int32_t* tickets /* nullable: [B*H] counters, zero on entry and left zero → the split-KV partials are
                                       merged by the last CTA of each (b,h) in the SAME launch (no merge kernel) */,
                   vg_stream_t stream);
/* *counter += delta (device-side step counter used with pos_dev) */
int vg_add_i32(int32_t* counter, int32_t delta, vg_stream_t stream);
/* bulk append of a prefill chunk: k,v rows [B, T, H*D] (stride ld) → cache[:, :, pos:pos+T, :] */
int vg_kv_append(const void* k, const void* v, int64_t ld, void* k_cache, void* v_cache,
                 int64_t B, int64_t H, int64_t D, int64_t T, int64_t Tmax, int64_t pos, int dtype,
                 vg_stream_t stream);

/* ---- weight-streaming linear layer for the cached generation step (decode batch B <= 256 rows):
 * every nn.Linear that LVTR.step runs on ONE new frame per sequence — attention.py:52,79, transformer/layers.py:82,152,
 * lvtr.py:171,172,194,195 — with the RMSNorm in front of it (norm.py:28-32) folded in.
 *   v[b,n] = sum_k (x[b,k] * norm_scale[k]) * W[n,k]            (norm_scale nullable: plain x)
 *   v     *= rsqrt(x_ss[b] / K + norm_eps)                       (only with norm_scale; x_ss[b] = sum_k x[b,k]^2)
 *   v      = act(v + bias[n]) + residual[b,n]                    (bias, residual nullable; act: none/ReLU/GELU/SiLU)
 *   y[b,n] = bf16(v)  and/or  y_f32[b,n] = v
 *   y_ss[b] += sum_n bf16(v)^2                                   (nullable; feeds the next layer's x_ss — the caller
 *                                                                 zeroes it, or passes it as `zero_ss` to an EARLIER
 *                                                                 call that does not read it)
 * Each SM streams a disjoint slab of W once (cp.async, issued before the programmatic-dependent-launch wait so that
 * it overlaps the previous kernel when allow_overlap = 1).  `workspace` (vg_decode_linear_workspace bytes) must be
 * ZERO-FILLED once by the caller; the kernel leaves it zeroed.  x / W are bf16; K % 64 == 0; ldx, ldw % 8 == 0.   */
typedef struct {
  int64_t B, N, K;
  const void* x; int64_t ldx;
  const void* W; int64_t ldw;
  const float* norm_scale; const float* x_ss; float norm_eps;
  const float* bias;
  int32_t act;
  const void* residual; int64_t ld_res;   /* may alias y (in-place residual update) */
  void* y; int64_t ldy;
  float* y_f32; int64_t ldy_f32;
  float* y_ss;
  float* zero_ss;                          /* nullable: [B] floats cleared by this call (after its dependency wait) */
  int32_t allow_overlap;                   /* 1: launch with programmatic stream serialization (PDL) */
} vg_decode_linear_args;
size_t vg_decode_linear_workspace(int64_t max_batch, int64_t max_n);
int    vg_decode_linear(const vg_decode_linear_args* a, void* workspace, size_t workspace_bytes, vg_stream_t stream);
/* debug aid: when non-null, CTA 0 of every following vg_decode_linear launch writes 7 clock64 phase stamps to buf */
int    vg_debug_decode_linear_trace(void* buf /* device, 48 x uint64, nullable */);
/* debug: clock64 phase stamps of CTA (0,0,0) of the tcgen05 attention kernels, 8 per loop iteration (tools/attn_trace.py) */
int    vg_debug_attn_trace(void* buf /* device, 256 x uint64, nullable */);

/* ---- fused latent front end: lvtr.py:151-169 (+ linear/layers.py:87-134,150-152, lvtr.py:390-392)
 * per frame: mean/logstd heads (Linear L→L), z = (mean + exp(logstd)·eps·temperature)·mask,
 * log_q = (−logstd − 0.5 − 0.5 ln 2π)·mask, u = tok_emb[id]·mask + ReLU(Wf z + bf),
 * u_shift[b,0] = s0[b]; u_shift[b,t] = u[b,t−1]; masked to the original lengths.                  */
typedef struct {
  int64_t B, T;
  int32_t latent_dim;       /* L (4) */
  int32_t emb_dim;          /* E (64) */
  int32_t vocab;
  const float* h_enc;       /* [B,T,L] encoder output */
  const float* eps;         /* [B,T,L] caller-supplied N(0,1) */
  const int64_t* ids;       /* [B,T] */
  const uint8_t* mask;      /* [B,T] 1 = valid */
  const float* init_state;  /* [B,E] (s0), nullable → zeros */
  const float* w_mean; const float* b_mean;       /* [L,L],[L] */
  const float* w_logstd; const float* b_logstd;   /* [L,L],[L] */
  const float* tok_emb;     /* [vocab,E] */
  const float* w_fuse; const float* b_fuse;       /* [E,L],[E] */
  float temperature;
  float* mean; float* logstd; float* z; float* log_q;    /* [B,T,L] outputs (f32) */
  void* u; void* u_shift;   /* [B,T,E] act dtype */
  int32_t act_dtype;
} vg_latent_front_args;
int vg_latent_front_fwd(const vg_latent_front_args* a, vg_stream_t stream);

typedef struct {
  vg_latent_front_args f;   /* forward arguments (inputs + saved outputs) */
  const float* d_z;         /* [B,T,L] gradient reaching z from the flow (nullable) */
  const float* d_log_q;     /* [B,T,L] nullable */
  const float* d_mean_out;  /* [B,T,L] gradient on the exposed mean (nullable) */
  const float* d_logstd_out;/* [B,T,L] nullable */
  const void* d_u;          /* [B,T,E] act dtype, nullable */
  const void* d_u_shift;    /* [B,T,E] act dtype, nullable */
  float* d_h_enc;           /* [B,T,L] */
  float* d_w_mean; float* d_b_mean; float* d_w_logstd; float* d_b_logstd;
  float* d_tok_emb; float* d_w_fuse; float* d_b_fuse;     /* overwritten */
} vg_latent_front_bwd_args;
size_t vg_latent_front_bwd_workspace(int64_t B, int64_t T, int32_t L, int32_t E, int32_t vocab);
int vg_latent_front_bwd(const vg_latent_front_bwd_args* a, void* workspace, size_t workspace_bytes,
                        vg_stream_t stream);

/* ---- fused latent back end: prior head consumption + conditional coupling flow + log_p + KL:
 * lvtr.py:172-191, flow/layers.py:42-73,225-234, linear/layers.py:278-288, trainers/speech/lvtr.py:122-124.
 * `head` is the [M, head_ld] f32 output of ONE GEMM whose columns are
 *   [0,L) prior mean, [L,2L) prior logstd, then per flow layer l: 2*Hd FiLM columns (gamma | beta).
 * Outputs log_p [M,L] (masked), y (flow output) and per-frame kl = mean_c(log_q − log_p);
 * kl_sum[0] = Σ_valid kl (deterministic).                                                          */
typedef struct {
  int64_t M;                /* B*T frames */
  int32_t latent_dim;       /* L = 4 (flow splits L/2 | L/2) */
  int32_t hidden;           /* Hd = 64 */
  int32_t n_layers;         /* 4 */
  const float* head; int64_t head_ld;
  const float* z;           /* [M,L] masked posterior sample */
  const float* log_q;       /* [M,L] */
  const uint8_t* mask;      /* [M] */
  /* flow parameters, layer-major contiguous: w1 [n,Hd,L/2], b1 [n,Hd], ln_w [n,Hd], ln_b [n,Hd],
   * w2 [n,L,Hd], b2 [n,L] */
  const float* w1; const float* b1; const float* ln_w; const float* ln_b;
  const float* w2; const float* b2;
  float ln_eps; float scale_lo; float scale_hi;  /* σ-range: logs = ln(sigmoid(s)*(lo−hi)+hi) */
  float* log_p;             /* [M,L] */
  float* y;                 /* [M,L] flow output */
  float* kl_frame;          /* [M] */
  float* kl_sum;            /* [1] */
} vg_latent_back_args;
size_t vg_latent_back_workspace(int64_t M, int32_t L, int32_t Hd, int32_t n_layers);
int vg_latent_back_fwd(const vg_latent_back_args* a, void* workspace, size_t workspace_bytes,
                       vg_stream_t stream);
typedef struct {
  vg_latent_back_args f;
  const float* d_log_p;     /* [M,L] upstream gradient on log_p */
  float* d_head;            /* [M, head_ld] */
  float* d_z;               /* [M,L] */
  float* d_w1; float* d_b1; float* d_ln_w; float* d_ln_b; float* d_w2; float* d_b2; /* overwritten */
} vg_latent_back_bwd_args;
int vg_latent_back_bwd(const vg_latent_back_bwd_args* a, void* workspace, size_t workspace_bytes,
                       vg_stream_t stream);
/* decode-time inverse flow + prior sample: lvtr.py:267-275, flow/layers.py:75-99,236-245.
 * z0 = mean_p + exp(logstd_p)·eps·temperature ; z = Flow^{-1}(z0; FiLM columns of head). */
int vg_latent_prior_sample(const vg_latent_back_args* a, const float* eps, float temperature,
                           float* z_out, vg_stream_t stream);

/* ---- conv-stack front half of a ResidualBlock on [B,T,C] rows (the first "next" row):
 * y = LayerNorm_C( depthwise_conv_k(x) + bias [+ t_add[b,:]] ) with zero padding at the array ends
 * (conv/layers.py:13-31,117-135,238-253) and the UNBIASED-variance channel norm of norm.py:43-47.
 * w_t is the depthwise weight transposed to [taps][C]; w_t == NULL means identity (plain channel LayerNorm, used
 * for BottleNeckResNet.final_norm).  mean/rstd [B*T] are saved for backward.  The surrounding 1x1 convolutions are
 * plain vg_gemm calls on the same [B*T, C] buffers.                                                          */
int vg_dwconv_ln_fwd(const void* x, const float* w_t, const float* bias /* nullable */,
                     const float* t_add /* [B,C] nullable */, const float* ln_w, const float* ln_b,
                     void* y, int64_t ld_y, float* mean, float* rstd,
                     int64_t B, int64_t T, int64_t C, int32_t taps, int32_t pad_left, float eps, int dtype,
                     vg_stream_t stream);
size_t vg_dwconv_ln_bwd_workspace(int64_t B, int64_t T, int64_t C, int32_t taps);
/* dh = dL/d(conv output) [B,T,C]; dx = transposed depthwise conv of dh;
 * dw_t [taps][C], d_ln_w, d_ln_b, d_bias [C] and d_t_add [B][C] (= sum_t dh, the gradient of the per-sequence time
 * embedding) are overwritten (deterministic two-stage reductions).                                              */
int vg_dwconv_ln_bwd(const void* dy, int64_t ld_dy, const void* x, const float* w_t, const float* bias,
                     const float* t_add, const float* ln_w, const float* mean, const float* rstd,
                     void* dh, void* dx, float* dw_t /* nullable iff w_t is NULL */, float* d_ln_w, float* d_ln_b,
                     float* d_bias /* nullable */, float* d_t_add /* nullable */, void* workspace, size_t workspace_bytes,
                     int64_t B, int64_t T, int64_t C, int32_t taps, int32_t pad_left, int dtype,
                     vg_stream_t stream);

/* ---- token cross-entropy: losses.py:30-41 (F.cross_entropy, ignore_index −100, reduction=sum)   */
size_t vg_softmax_ce_workspace(int64_t rows);
int vg_softmax_ce_fwd(const void* logits, int64_t ld, const int64_t* targets,
                      const uint8_t* row_mask /* nullable; masked rows are ignored */,
                      float* lse, float* loss_rows, float* loss_sum,
                      int64_t rows, int64_t vocab, int dtype,
                      void* workspace, size_t workspace_bytes, vg_stream_t stream);
int vg_softmax_ce_bwd(const void* logits, int64_t ld, const int64_t* targets,
                      const uint8_t* row_mask, const float* lse, const float* d_loss /* [1] device */,
                      void* d_logits, int64_t ld_d, int64_t rows, int64_t vocab, int dtype,
                      vg_stream_t stream);

/* ---- diffusion loss glue: ddpm.py:328-334 (q_sample), 345-366 + losses.py:9-27,44-57 (masked L1) */
int vg_qsample(const float* x0, const float* noise, const int64_t* t /* [B] */,
               const float* sqrt_ac /* [steps] */, const float* sqrt_1mac, const uint8_t* mask,
               float x0_scale, float* x_t, float* target, int64_t B, int64_t T, int64_t C,
               vg_stream_t stream);
size_t vg_masked_l1_workspace(int64_t B, int64_t T, int64_t C);
int vg_masked_l1_fwd(const float* pred, const float* target, const uint8_t* mask, float* loss_sum,
                     int64_t B, int64_t T, int64_t C, void* workspace, size_t workspace_bytes,
                     vg_stream_t stream);
int vg_masked_l1_bwd(const float* pred, const float* target, const uint8_t* mask,
                     const float* d_loss, float* d_pred, int64_t B, int64_t T, int64_t C,
                     vg_stream_t stream);

/* ---- token sampling: lvtr.py:277-285 (softmax(logits/τ) → multinomial), plus a greedy mode
 * (argmax; lowest index wins ties) used for bit-exact parity.  `u` = caller-supplied U[0,1) [rows].  */
int vg_sample_token(const void* logits, int64_t ld, const float* u /* nullable → greedy */,
                    float temperature, int64_t* out_ids, int64_t rows, int64_t vocab, int dtype,
                    vg_stream_t stream);

/* ---- optimizer: training_lib/optimizer.py:18-25,110-130 (AdamW) fused with the bf16 shadow cast  */
int vg_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  void* shadow_bf16 /* nullable */, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float bias_corr1, float bias_corr2,
                  float grad_scale,
                  const float* hyper_dev /* nullable: device {lr, 1/bc1, 1/sqrt(bc2), 1-lr*wd} overriding the
                                            scalar arguments, so a captured CUDA graph can be replayed */,
                  vg_stream_t stream);
int vg_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vg_stream_t stream);

/* Clear n_seg element ranges [seg_off[i], seg_off[i] + seg_len[i]) of one fp32 buffer in ONE launch (offsets and lengths
 * in elements, multiples of 4; tables in device memory).  The gradient arena (vae_gslm_b200/arena.py) clears only the
 * slices whose gradients are ACCUMULATED by autograd (torch's AccumulateGrad, training_lib/optimizer.py semantics); the
 * big matrices are overwritten by their first weight-gradient GEMM of the step (beta = 0) instead. */
int vg_zero_segments(float* base, const int64_t* seg_off, const int64_t* seg_len, int n_seg, vg_stream_t stream);

/* ---- skinny linear layer for the cached generation step: y[B,N] = epilogue(x[B,K] · W[N,K]^T), B <= 256 rows — every
 * nn.Linear that LVTR.step runs on one new frame per sequence (attention.py:52,79, transformer/layers.py:82,152,
 * lvtr.py:171,172,194,195) when the step runs layer by layer.  Swap-AB tcgen05 GEMM: 128 output features per CTA as the M
 * operand, the batch as the N operand, the k-range split over a thread-block cluster whose partial tiles are
 * reduce-scattered through distributed shared memory; weights prefetched ahead of the programmatic-dependent-launch wait.
 * Epilogue order as vg_gemm: (x row scale 1/rms) → + bias → activation → (row mask if mask_before_residual) → + residual → (row mask otherwise).
 * x, W bf16 with unit inner stride (ld multiples of 8); y / residual bf16 or f32 (y_dtype); K a multiple of 64.            */
typedef struct {
  const void* x; int64_t ldx;
  const void* w; int64_t ldw;
  void* y; int64_t ldy;
  const float* bias;                 /* nullable, [N] */
  const void* residual; int64_t ld_res;   /* nullable, [B, N] in y_dtype */
  const uint8_t* row_mask;           /* nullable, [B] */
  int64_t B, N, K;
  int32_t y_dtype;                   /* VG_F32 | VG_BF16 */
  int32_t act;                       /* vg_act: NONE | RELU | GELU | SILU */
  int32_t mask_before_residual;
  float ss_inv_k, ss_eps;            /* folded RMSNorm (norm.py:28-32): acc *= rsqrt(row_ss_in[b] * ss_inv_k + ss_eps), the norm's */
  const float* row_ss_in;            /* scale vector pre-multiplied into W by the caller; nullable                              */
  float* row_ss_out;                 /* nullable, [B]: += sum over features of the stored y^2 (the next RMSNorm's statistics)    */
  float* zero_ss;                    /* nullable, [B]: cleared (the accumulator of a later launch)                               */
} vg_skinny_linear_args;
int vg_skinny_linear(const vg_skinny_linear_args* a, vg_stream_t stream);

/* ---- strided Conv1d of the utterance encoder as a GEMM on [B,T,C] rows (modules/conv/layers.py:549-593 ConvNormAct,
 * models/speech/lvtr.py:127-136,203-207): the window gather in front of vg_gemm and its adjoint.
 *   a[b, t, c*K + j] = f(x[b, S*t - pad + j, c])  (0 outside [0,T)),  T_out = (T + 2 pad - K) / S + 1,  f = ReLU if relu
 * Column order = memory order of the Conv1d weight [Cout, Cin, K]: the convolution is one GEMM with the weight in place.
 * bwd: dx = adjoint(da), multiplied by [x_relu > 0] when x_relu (the forward input of a relu gather) is given.          */
int vg_im2col_fwd(const void* x, void* a, int64_t B, int64_t T, int64_t C, int64_t K, int64_t S, int64_t pad,
                  int64_t T_out, int relu, int dtype, vg_stream_t stream);
int vg_im2col_bwd(const void* da, const void* x_relu /* nullable */, void* dx, int64_t B, int64_t T, int64_t C, int64_t K,
                  int64_t S, int64_t pad, int64_t T_out, int dtype, vg_stream_t stream);

/* ---- persistent cached-generation step: the transformer part of LVTR.step (models/speech/lvtr.py:253-279 →
 * modules/transformer/layers.py:134-195 with past_kv, modules/attention/attention.py:52-85, modules/norm.py:28-32,
 * lvtr.py:171,172,194,195) as ONE cooperative launch — stack-input linear, L x [RMSNorm1+QKV | cached attention + KV append |
 * out-proj+residual | RMSNorm3+FFN1 | GELU+FFN2+residual], final RMSNorm, q_spliter|token_spliter, prior/FiLM head, logits.
 * The step is a list of NP phases separated by device-wide barriers.  In a GEMM phase every CTA runs at most one unit of
 * the host-built task table `tasks[grid][NP]` (vae_gslm_b200/decode_step.py builds it): `R` (multiple of 16, <= 256) output
 * features x `nkb` 64-wide k-blocks of one linear layer as a swap-AB tcgen05 GEMM whose weight slab comes from the CTA's
 * own packed byte stream (`wstream + wstream_off[cta]`: per unit, per k-block, R/8 SWIZZLE_128B atoms of 8 rows x 64 bf16,
 * in consumption order) and whose X operand the CTA forms from `x` (x_kind 0: bf16 rows; 1: f32 rows * vec[k], optionally
 * accumulating the row sum-of-squares into ss_out; 2: act(f32 rows * rsqrt(ss_in[b]*inv_k + eps) + vec[k]); 3: the
 * attention output merged from the kv-split partials `attn_partial` — see late_merge).  The
 * accumulator rows are reduced (red.global.add.f32) or stored into acc[b*ldacc + n0 + r] (+ bias_out[n0 + r]).
 * phase_kind[p] < 0: GEMM phase; >= 0: attention phase of that layer over the head-major cache
 * [L][2][B][H][Tmax][64] bf16 (reads the QKV sums of the phase before, appends k/v at *pos_dev).
 * aux jobs: 1 = clear aux_i1 float4 at float4 index aux_i0 of aux_ptr0; 2 = write bf16 transformer_latent groups.
 * State carried between launches (caller-allocated, zero-initialised once): bar_flags [grid] u32, epoch u32, tickets.   */
typedef struct {
  const void* x;            /* X source rows */
  float* acc;               /* accumulator [B, ldacc] f32 */
  const float* vec;         /* per-k scale (x_kind 1) or bias (x_kind 2) */
  const float* ss_in;       /* [B] row sum-of-squares for x_kind 2 (nullable: rstd = 1) */
  float* ss_out;            /* [B] accumulated by this unit (x_kind 1; nullable) */
  const float* bias_out;    /* added to this unit's accumulator rows (nullable) */
  void* aux_ptr0; void* aux_ptr1; void* aux_ptr2; void* aux_ptr3;
  int32_t ldx, ldacc;
  int32_t k0, nkb;
  int32_t n0, rows;         /* first feature, number of valid features (<= R) */
  int32_t R;                /* padded rows of the packed slab; 0 = no GEMM unit in this phase */
  int32_t x_kind, act, store;
  float inv_k, eps;
  int32_t aux_kind, aux_i0, aux_i1, aux_i2;
} vg_decode_step_task;

typedef struct {
  const vg_decode_step_task* tasks;      /* [grid][NP] */
  const int32_t* phase_kind;             /* [NP] */
  const uint8_t* wstream; const int64_t* wstream_off;   /* packed weights, per-CTA byte offsets [grid] */
  /* attention phases */
  const float* qkv_acc;                  /* [B, 3*H*64] f32 sums of the QKV phase */
  const float* ss_base;                  /* row sum-of-squares arrays: layer l's RMSNorm1 at ss_base + 2*l*B */
  void* cache;                           /* bf16 [L][2][B][H][Tmax][64] */
  int64_t cache_layer_stride, cache_kv_stride;          /* in elements */
  void* attn_out;                        /* bf16 [B, H*64] */
  const float* slopes;                   /* [H] ALiBi slopes (nullable) */
  float* attn_partial;                   /* [B*H*nsplit*(64+8)] f32 (nsplit > 1) */
  int32_t* tickets;                      /* [B*H], zero on entry and left zero */
  int32_t* pos_dev;                      /* number of cached positions; advanced by the kernel when advance_pos */
  uint32_t* bar_flags; uint32_t* epoch;  /* barrier state, persistent across launches */
  unsigned long long* debug;             /* nullable: [8] words written before a time-out trap */
  long long* trace;                      /* nullable: [grid][NP][8] clock64 stamps (barrier entry / exit, unit stages) */
  float inv_d, eps, scale;               /* 1/d_model, RMSNorm eps, softmax scale */
  int32_t NP, grid, B, Bp, H, Tmax, nsplit, barrier_mode, advance_pos;
  int32_t rep;                           /* copies of the batch rows in the X tile / accumulator: 4 (B <= 32), 2 (<= 64), 1 */
  int32_t attn_coop;                     /* 1: one CTA per attention item (few sequences); 0: one warp per item */
  int32_t late_merge;                    /* 1: attention phases publish (m, l, o) per kv-split only; the out-projection unit (x_kind 3,
                                            x = attn_partial) merges and normalises them while forming its X operand */
} vg_decode_step_args;
size_t vg_decode_step_task_bytes(void);
size_t vg_decode_step_smem_bytes(int32_t n_phases);
int    vg_decode_step(const vg_decode_step_args* a, vg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VGSLM_H_ */
