// latent.cu — the fused latent kernels of the VAE-GSLM step (warp-level primitives, no tensor cores).
//
//  front : lvtr.py:151-169 — posterior heads (linear/layers.py:91-106), reparameterised sample from a
//          caller-supplied eps (:114-128; note the reference parameterises with logstd, not logvar),
//          log_q (:158), token embedding (:150-152), fuse (lvtr.py:390-392) and the BOS shift (:161-169).
//  back  : lvtr.py:172-191 — prior head consumption, 4 conditional affine coupling layers
//          (flow/layers.py:42-73 with FiLM linear/layers.py:278-288), log_p and the per-frame KL
//          (trainers/speech/lvtr.py:122-124), plus the matching backward and the decode-time inverse
//          (flow/layers.py:75-99,236-245).
//
// All of it is HBM/latency-bound per-frame math on tiny matrices (4x4, 64x4, 64x2, 4x64); the reference
// spends ~100 kernel launches on it.  Parameter-gradient reductions are two-stage and deterministic.
#include <math_constants.h>
#include "common.cuh"

namespace vg {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;
constexpr int LAT = 4;     // latent dim (compile-time; validated at the ABI)
constexpr int EMB = 64;    // token embedding dim
constexpr int FH = 64;     // flow hidden dim
constexpr int FL_MAX = 4;  // max flow layers (per-layer backward state is register resident)

// ============================================================================================ front
struct FrontP {
  vg_latent_front_args a;
};

// 16 lanes per frame; lane `sub` owns embedding channels sub*4..sub*4+3.
template <typename T>
__global__ void __launch_bounds__(256)
latent_front_fwd_kernel(const vg_latent_front_args a) {
  const int64_t M = a.B * a.T;
  const int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 16;
  const int sub = threadIdx.x & 15;
  if (f >= M) return;
  const int64_t b = f / a.T, t = f % a.T;
  const bool valid = a.mask[f] != 0;

  float he[LAT], mean[LAT], logstd[LAT], z[LAT];
#pragma unroll
  for (int j = 0; j < LAT; ++j) he[j] = a.h_enc[f * LAT + j];
#pragma unroll
  for (int i = 0; i < LAT; ++i) {
    float mu = a.b_mean[i], ls = a.b_logstd[i];
#pragma unroll
    for (int j = 0; j < LAT; ++j) {
      mu = fmaf(a.w_mean[i * LAT + j], he[j], mu);
      ls = fmaf(a.w_logstd[i * LAT + j], he[j], ls);
    }
    mean[i] = mu; logstd[i] = ls;
    const float zr = mu + expf(ls) * a.eps[f * LAT + i] * a.temperature;
    z[i] = valid ? zr : 0.f;
  }
  if (sub < LAT) {
    a.mean[f * LAT + sub] = mean[sub];
    a.logstd[f * LAT + sub] = logstd[sub];
    a.z[f * LAT + sub] = z[sub];
    a.log_q[f * LAT + sub] = valid ? (-logstd[sub] - 0.5f - kHalfLog2Pi) : 0.f;
  }
  int64_t id = a.ids[f];
  id = id < 0 ? 0 : (id >= a.vocab ? a.vocab - 1 : id);
  T* u = reinterpret_cast<T*>(a.u);
  T* us = reinterpret_cast<T*>(a.u_shift);
  const bool next_valid = (t + 1 < a.T) && a.mask[f + 1] != 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int e = sub * 4 + c;
    float pre = a.b_fuse[e];
#pragma unroll
    for (int j = 0; j < LAT; ++j) pre = fmaf(a.w_fuse[e * LAT + j], z[j], pre);
    const float val = (valid ? a.tok_emb[id * EMB + e] : 0.f) + fmaxf(pre, 0.f);
    u[f * EMB + e] = from_f32<T>(val);
    if (t + 1 < a.T) us[(f + 1) * EMB + e] = from_f32<T>(next_valid ? val : 0.f);
    if (t == 0) {
      const float s0 = a.init_state ? a.init_state[b * EMB + e] : 0.f;
      us[f * EMB + e] = from_f32<T>(valid ? s0 : 0.f);
    }
  }
}

// per-thread partial layout of the small front parameters (floats):
//   [0,16) dWm  [16,20) dbm  [20,36) dWs  [36,40) dbs  [40,296) dWf[e][j]  [296,360) dbf
constexpr int FRONT_NPARAM = 16 + 4 + 16 + 4 + EMB * LAT + EMB;   // 360
constexpr int FRONT_BWD_THREADS = 256;                             // 16 frame slots x 16 lanes

template <typename T>
__global__ void __launch_bounds__(FRONT_BWD_THREADS)
latent_front_bwd_kernel(const vg_latent_front_bwd_args g, float* __restrict__ partial, float* __restrict__ gu_out) {
  __shared__ float red[FRONT_BWD_THREADS / 16][FRONT_NPARAM];
  const vg_latent_front_args& a = g.f;
  const int64_t M = a.B * a.T;
  const int slot = threadIdx.x / 16, sub = threadIdx.x & 15;
  const T* d_u = reinterpret_cast<const T*>(g.d_u);
  const T* d_us = reinterpret_cast<const T*>(g.d_u_shift);

  float acc_wf[4][LAT], acc_bf[4];           // channels sub*4..+3
  float acc_wm[LAT], acc_ws[LAT], acc_bm = 0.f, acc_bs = 0.f;   // row i = sub (sub < LAT)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    acc_bf[c] = 0.f;
#pragma unroll
    for (int j = 0; j < LAT; ++j) acc_wf[c][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < LAT; ++j) { acc_wm[j] = 0.f; acc_ws[j] = 0.f; }

  const int64_t frames_per_iter = (int64_t)gridDim.x * (FRONT_BWD_THREADS / 16);
  // the loop bound is CTA-uniform (shuffles below need whole warps); out-of-range slots are predicated
  for (int64_t base = (int64_t)blockIdx.x * (FRONT_BWD_THREADS / 16); base < M; base += frames_per_iter) {
    const bool active = base + slot < M;
    const int64_t f = active ? base + slot : M - 1;
    const int64_t t = f % a.T;
    const bool valid = active && a.mask[f] != 0;
    const bool next_valid = (t + 1 < a.T) && a.mask[f + 1] != 0;
    float he[LAT], z[LAT];
#pragma unroll
    for (int j = 0; j < LAT; ++j) { he[j] = a.h_enc[f * LAT + j]; z[j] = a.z[f * LAT + j]; }
    float gz[LAT];
#pragma unroll
    for (int j = 0; j < LAT; ++j) gz[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int e = sub * 4 + c;
      float gu = d_u ? to_f32<T>(d_u[f * EMB + e]) : 0.f;
      if (d_us && next_valid) gu += to_f32<T>(d_us[(f + 1) * EMB + e]);
      if (!active) gu = 0.f;
      if (active) gu_out[f * EMB + e] = valid ? gu : 0.f;   // gradient reaching tok_emb[id] (second kernel)
      float pre = a.b_fuse[e];
#pragma unroll
      for (int j = 0; j < LAT; ++j) pre = fmaf(a.w_fuse[e * LAT + j], z[j], pre);
      const float gpre = pre > 0.f ? gu : 0.f;
      acc_bf[c] += gpre;
#pragma unroll
      for (int j = 0; j < LAT; ++j) {
        acc_wf[c][j] = fmaf(gpre, z[j], acc_wf[c][j]);
        gz[j] = fmaf(a.w_fuse[e * LAT + j], gpre, gz[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < LAT; ++j) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) gz[j] += __shfl_xor_sync(0xffffffffu, gz[j], o);
      if (g.d_z) gz[j] += g.d_z[f * LAT + j];
      if (!valid) gz[j] = 0.f;                         // z = where(mask, z_raw, 0)
    }
    float gmean[LAT], glogstd[LAT];
#pragma unroll
    for (int i = 0; i < LAT; ++i) {
      const float std_i = expf(a.logstd[f * LAT + i]);
      gmean[i] = gz[i] + ((active && g.d_mean_out) ? g.d_mean_out[f * LAT + i] : 0.f);
      glogstd[i] = gz[i] * std_i * a.eps[f * LAT + i] * a.temperature
                   - ((valid && g.d_log_q) ? g.d_log_q[f * LAT + i] : 0.f)
                   + ((active && g.d_logstd_out) ? g.d_logstd_out[f * LAT + i] : 0.f);
    }
    if (sub < LAT && active) {
      float dh = 0.f;
#pragma unroll
      for (int i = 0; i < LAT; ++i) dh += a.w_mean[i * LAT + sub] * gmean[i] + a.w_logstd[i * LAT + sub] * glogstd[i];
      g.d_h_enc[f * LAT + sub] = dh;
      // row i = sub of the two head weights
      float gm = 0.f, gs = 0.f;
#pragma unroll
      for (int i = 0; i < LAT; ++i) { if (i == sub) { gm = gmean[i]; gs = glogstd[i]; } }
      acc_bm += gm; acc_bs += gs;
#pragma unroll
      for (int j = 0; j < LAT; ++j) { acc_wm[j] = fmaf(gm, he[j], acc_wm[j]); acc_ws[j] = fmaf(gs, he[j], acc_ws[j]); }
    }
  }
  // ---- stage 1: every (slot, sub) thread owns a disjoint set of parameters; reduce over the 16 slots
  float* my = red[slot];
  if (sub < LAT) {
#pragma unroll
    for (int j = 0; j < LAT; ++j) { my[sub * LAT + j] = acc_wm[j]; my[20 + sub * LAT + j] = acc_ws[j]; }
    my[16 + sub] = acc_bm; my[36 + sub] = acc_bs;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int e = sub * 4 + c;
#pragma unroll
    for (int j = 0; j < LAT; ++j) my[40 + e * LAT + j] = acc_wf[c][j];
    my[296 + e] = acc_bf[c];
  }
  __syncthreads();
  for (int p = threadIdx.x; p < FRONT_NPARAM; p += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int sl = 0; sl < FRONT_BWD_THREADS / 16; ++sl) s += red[sl][p];
    partial[(int64_t)blockIdx.x * FRONT_NPARAM + p] = s;
  }
}

__global__ void latent_front_bwd_reduce_kernel(const float* __restrict__ partial, int nparts,
                                               const vg_latent_front_bwd_args g) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= FRONT_NPARAM) return;
  float s = 0.f;
  for (int i = 0; i < nparts; ++i) s += partial[(int64_t)i * FRONT_NPARAM + p];
  if (p < 16) g.d_w_mean[p] = s;
  else if (p < 20) g.d_b_mean[p - 16] = s;
  else if (p < 36) g.d_w_logstd[p - 20] = s;
  else if (p < 40) g.d_b_logstd[p - 36] = s;
  else if (p < 296) g.d_w_fuse[p - 40] = s;
  else g.d_b_fuse[p - 296] = s;
}

// d_tok_emb[v,:] = sum over frames with id == v of gu[f,:], in frame order (deterministic).
// one CTA per vocabulary entry, 8 warps scan the id list with coalesced loads + ballot.
__global__ void __launch_bounds__(256)
embedding_grad_kernel(const int64_t* __restrict__ ids, const float* __restrict__ gu, float* __restrict__ d_emb,
                      int64_t M, int vocab) {
  __shared__ float red[8][EMB];
  const int v = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a0 = 0.f, a1 = 0.f;
  const int64_t per_warp = (M + 7) / 8;
  const int64_t beg = warp * per_warp, end = (beg + per_warp < M) ? beg + per_warp : M;
  for (int64_t base = beg; base < end; base += 32) {
    const int64_t f = base + lane;
    int64_t id = (f < end) ? ids[f] : -1;
    if (f < end) id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    unsigned hit = __ballot_sync(0xffffffffu, id == v);
    while (hit) {
      const int bit = __ffs(hit) - 1;
      hit &= hit - 1;
      const float* row = gu + (base + bit) * EMB;
      a0 += row[lane];
      a1 += row[lane + 32];
    }
  }
  red[warp][lane] = a0;
  red[warp][lane + 32] = a1;
  __syncthreads();
  if (threadIdx.x < EMB) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    d_emb[(int64_t)v * EMB + threadIdx.x] = s;
  }
}

static int front_bwd_blocks(int64_t M) {
  int64_t want = ceil_div(M, FRONT_BWD_THREADS / 16);
  int64_t cap = kNumSMs * 2;
  return (int)(want < cap ? want : cap);
}

// ============================================================================================= back
// warp per frame; lane owns hidden units j0 = lane and j1 = lane + 32.  The flow parameters (580 floats
// per layer) are staged once per CTA into shared memory in the layout
//   w1[64][2] | b1[64] | lnw[64] | lnb[64] | w2[4][64] | b2[4]
constexpr int FLOW_LAYER_NPARAM = FH * 2 + FH * 3 + LAT * FH + LAT;   // 580

__device__ __forceinline__ void stage_flow_params(float* __restrict__ sw, const vg_latent_back_args& a) {
  for (int i = threadIdx.x; i < a.n_layers * FLOW_LAYER_NPARAM; i += blockDim.x) {
    const int l = i / FLOW_LAYER_NPARAM, r = i % FLOW_LAYER_NPARAM;
    float v;
    if (r < 2 * FH) v = a.w1[l * 2 * FH + r];
    else if (r < 3 * FH) v = a.b1[l * FH + (r - 2 * FH)];
    else if (r < 4 * FH) v = a.ln_w[l * FH + (r - 3 * FH)];
    else if (r < 5 * FH) v = a.ln_b[l * FH + (r - 4 * FH)];
    else if (r < 5 * FH + LAT * FH) v = a.w2[l * LAT * FH + (r - 5 * FH)];
    else v = a.b2[l * LAT + (r - 5 * FH - LAT * FH)];
    sw[i] = v;
  }
}

struct FlowLayerW {           // this lane's slice of one coupling layer's parameters
  float w1[2][2], b1[2], lnw[2], lnb[2], w2[LAT][2], b2[LAT];
};
__device__ __forceinline__ void load_flow_layer(FlowLayerW& w, const float* __restrict__ p, int lane) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = lane + 32 * u;
    w.w1[u][0] = p[j * 2 + 0];
    w.w1[u][1] = p[j * 2 + 1];
    w.b1[u] = p[2 * FH + j];
    w.lnw[u] = p[3 * FH + j];
    w.lnb[u] = p[4 * FH + j];
#pragma unroll
    for (int k = 0; k < LAT; ++k) w.w2[k][u] = p[5 * FH + k * FH + j];
  }
#pragma unroll
  for (int k = 0; k < LAT; ++k) w.b2[k] = p[5 * FH + LAT * FH + k];
}

struct FlowLayerState {       // forward values needed by the backward of one layer
  float x0[2], x1[2];         // conditioning half / transformed half (inputs)
  float shat[2], rs;          // layer-norm normalised value (this lane's units) and rstd
  float nrm[2], f[2];         // LN output, FiLM output (pre-GELU)
  float sg[2], v[2];          // sigmoid(raw log-scale), scale = sg*(lo-hi)+hi
};

// conditioner network of one coupling layer: x0 → (m[2], raw log-scale[2]); fills st when given
__device__ __forceinline__ void flow_conditioner(const FlowLayerW& w, const float* __restrict__ film, int lane,
                                                 float ln_eps, const float x0[2], float out[LAT], FlowLayerState* st) {
  float s[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) s[u] = fmaf(w.w1[u][0], x0[0], fmaf(w.w1[u][1], x0[1], w.b1[u]));
  const float mu = warp_sum(s[0] + s[1]) * (1.f / FH);
  const float d0 = s[0] - mu, d1 = s[1] - mu;
  const float var = warp_sum(d0 * d0 + d1 * d1) * (1.f / FH);
  const float rs = rsqrtf(var + ln_eps);
  float g[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = lane + 32 * u;
    const float sh = (u == 0 ? d0 : d1) * rs;
    const float n = fmaf(sh, w.lnw[u], w.lnb[u]);
    const float f = fmaf(film[j], n, film[FH + j]);      // gamma | beta
    g[u] = gelu_f(f);
    if (st) { st->shat[u] = sh; st->nrm[u] = n; st->f[u] = f; }
  }
  if (st) st->rs = rs;
#pragma unroll
  for (int k = 0; k < LAT; ++k) out[k] = warp_sum(fmaf(w.w2[k][0], g[0], w.w2[k][1] * g[1])) + w.b2[k];
}

__global__ void __launch_bounds__(128)
latent_back_fwd_kernel(const vg_latent_back_args a, float* __restrict__ kl_partial) {
  __shared__ float sw[FL_MAX * FLOW_LAYER_NPARAM];
  __shared__ float kl_red[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_flow_params(sw, a);
  __syncthreads();
  float kl_acc = 0.f;
  for (int64_t f = (int64_t)blockIdx.x * 4 + warp; f < a.M; f += (int64_t)gridDim.x * 4) {
    const bool valid = a.mask[f] != 0;
    const float* head = a.head + f * a.head_ld;
    float x[LAT];
#pragma unroll
    for (int c = 0; c < LAT; ++c) x[c] = a.z[f * LAT + c];
    float ld = 0.f;
    for (int l = 0; l < a.n_layers; ++l) {
      FlowLayerW w;
      load_flow_layer(w, sw + l * FLOW_LAYER_NPARAM, lane);
      const float x0[2] = {x[2], x[3]}, x1[2] = {x[0], x[1]};
      float out[LAT];
      flow_conditioner(w, head + 2 * LAT + l * 2 * FH, lane, a.ln_eps, x0, out, nullptr);
      float nx[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float sg = 1.f / (1.f + expf(-out[2 + c]));
        const float v = sg * (a.scale_lo - a.scale_hi) + a.scale_hi;
        nx[c] = fmaf(x1[c], v, out[c]);
        ld += logf(v);
      }
      x[0] = x0[0]; x[1] = x0[1]; x[2] = nx[0]; x[3] = nx[1];
    }
    float kl = 0.f;
    if (lane < LAT) {
      float xl = x[0];
#pragma unroll
      for (int c = 1; c < LAT; ++c) if (lane == c) xl = x[c];
      const float mu = head[lane], sp = head[LAT + lane];
      const float diff = xl - mu;
      const float lp = ld * (1.f / LAT) - sp - kHalfLog2Pi - 0.5f * expf(-2.f * sp) * diff * diff;
      a.log_p[f * LAT + lane] = valid ? lp : 0.f;
      a.y[f * LAT + lane] = xl;
      kl = valid ? (a.log_q[f * LAT + lane] - lp) * (1.f / LAT) : 0.f;
    }
    kl = warp_sum(kl);
    if (lane == 0) a.kl_frame[f] = kl;
    kl_acc += kl;
  }
  if (lane == 0) kl_red[warp] = kl_acc;
  __syncthreads();
  if (threadIdx.x == 0) kl_partial[blockIdx.x] = kl_red[0] + kl_red[1] + kl_red[2] + kl_red[3];
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  // single warp, fixed order → deterministic
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(128)
latent_back_bwd_kernel(const vg_latent_back_bwd_args g, float* __restrict__ partial) {
  // dynamic smem: staged parameters [nl*580] followed by per-warp gradient accumulators [4][nl*580]
  extern __shared__ float dsm[];
  const vg_latent_back_args& a = g.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nl = a.n_layers;
  const int np = nl * FLOW_LAYER_NPARAM;
  float* sw = dsm;
  float* acc = dsm + np + warp * np;      // each lane only touches its own parameter slots → no races
  stage_flow_params(sw, a);
  for (int i = threadIdx.x; i < 4 * np; i += blockDim.x) dsm[np + i] = 0.f;
  __syncthreads();

  for (int64_t f = (int64_t)blockIdx.x * 4 + warp; f < a.M; f += (int64_t)gridDim.x * 4) {
    const bool valid = a.mask[f] != 0;
    float* dhead = g.d_head + f * a.head_ld;
    if (!valid) {          // masked frame: log_p is the constant 0 → every gradient is 0
      for (int c = lane; c < 2 * LAT + nl * 2 * FH; c += 32) dhead[c] = 0.f;
      if (lane < LAT) g.d_z[f * LAT + lane] = 0.f;
      continue;
    }
    const float* head = a.head + f * a.head_ld;
    // ---- recompute forward, keeping per-layer state
    FlowLayerState st[FL_MAX];
    float x[LAT];
#pragma unroll
    for (int c = 0; c < LAT; ++c) x[c] = a.z[f * LAT + c];
#pragma unroll
    for (int l = 0; l < FL_MAX; ++l) {
      if (l < nl) {
        FlowLayerW w;
        load_flow_layer(w, sw + l * FLOW_LAYER_NPARAM, lane);
        st[l].x0[0] = x[2]; st[l].x0[1] = x[3]; st[l].x1[0] = x[0]; st[l].x1[1] = x[1];
        float out[LAT];
        flow_conditioner(w, head + 2 * LAT + l * 2 * FH, lane, a.ln_eps, st[l].x0, out, &st[l]);
        float nx[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          st[l].sg[c] = 1.f / (1.f + expf(-out[2 + c]));
          st[l].v[c] = st[l].sg[c] * (a.scale_lo - a.scale_hi) + a.scale_hi;
          nx[c] = fmaf(st[l].x1[c], st[l].v[c], out[c]);
        }
        x[0] = st[l].x0[0]; x[1] = st[l].x0[1]; x[2] = nx[0]; x[3] = nx[1];
      }
    }
    // ---- log_p head
    float dx[LAT], dld = 0.f;
#pragma unroll
    for (int c = 0; c < LAT; ++c) {
      const float glp = g.d_log_p[f * LAT + c];
      const float mu = head[c], sp = head[LAT + c];
      const float diff = x[c] - mu;
      const float e2 = expf(-2.f * sp);
      dx[c] = -glp * e2 * diff;
      dld += glp * (1.f / LAT);
      if (lane == c) {
        dhead[c] = glp * e2 * diff;                        // d mean_p
        dhead[LAT + c] = glp * (e2 * diff * diff - 1.f);   // d logstd_p
      }
    }
    // ---- coupling layers in reverse
#pragma unroll
    for (int l = FL_MAX - 1; l >= 0; --l) {
      if (l < nl) {
        const FlowLayerState& s = st[l];
        FlowLayerW w;
        load_flow_layer(w, sw + l * FLOW_LAYER_NPARAM, lane);
        float* ga = acc + l * FLOW_LAYER_NPARAM;
        const float* film = head + 2 * LAT + l * 2 * FH;
        float* dfilm = dhead + 2 * LAT + l * 2 * FH;
        float dout[LAT], dx1[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float dx1p = dx[2 + c];
          dout[c] = dx1p;                                   // d m
          dx1[c] = dx1p * s.v[c];
          const float dv = dx1p * s.x1[c] + dld / s.v[c];   // through x1' and through logdet = log v
          dout[2 + c] = dv * (a.scale_lo - a.scale_hi) * s.sg[c] * (1.f - s.sg[c]);
        }
        float ds[2], dsh[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = lane + 32 * u;
          const float gj = gelu_f(s.f[u]);
          float dg = 0.f;
#pragma unroll
          for (int k = 0; k < LAT; ++k) {
            ga[5 * FH + k * FH + j] += dout[k] * gj;        // d w2[k][j]
            dg = fmaf(w.w2[k][u], dout[k], dg);
          }
          const float df = dg * gelu_grad_f(s.f[u]);
          dfilm[j] = df * s.nrm[u];                         // d gamma
          dfilm[FH + j] = df;                               // d beta
          const float dn = df * film[j];
          ga[3 * FH + j] += dn * s.shat[u];                 // d ln_w
          ga[4 * FH + j] += dn;                             // d ln_b
          dsh[u] = dn * w.lnw[u];
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < LAT; ++k) ga[5 * FH + LAT * FH + k] += dout[k];   // d b2
        }
        const float m1 = warp_sum(dsh[0] + dsh[1]) * (1.f / FH);
        const float m2 = warp_sum(dsh[0] * s.shat[0] + dsh[1] * s.shat[1]) * (1.f / FH);
        float dx0[2] = {0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = lane + 32 * u;
          ds[u] = s.rs * (dsh[u] - m1 - s.shat[u] * m2);
          ga[j * 2 + 0] += ds[u] * s.x0[0];                 // d w1[j][0]
          ga[j * 2 + 1] += ds[u] * s.x0[1];
          ga[2 * FH + j] += ds[u];                          // d b1
          dx0[0] = fmaf(w.w1[u][0], ds[u], dx0[0]);
          dx0[1] = fmaf(w.w1[u][1], ds[u], dx0[1]);
        }
        dx0[0] = warp_sum(dx0[0]) + dx[0];
        dx0[1] = warp_sum(dx0[1]) + dx[1];
        // layer input was (x1, x0)
        dx[0] = dx1[0]; dx[1] = dx1[1]; dx[2] = dx0[0]; dx[3] = dx0[1];
      }
    }
    if (lane < LAT) {
      float d = dx[0];
#pragma unroll
      for (int c = 1; c < LAT; ++c) if (lane == c) d = dx[c];
      g.d_z[f * LAT + lane] = d;
    }
  }
  // ---- stage 1: reduce the 4 warps of this CTA in fixed order
  __syncthreads();
  const float* r0 = dsm + np;
  for (int p = threadIdx.x; p < np; p += blockDim.x)
    partial[(int64_t)blockIdx.x * np + p] = r0[p] + r0[np + p] + r0[2 * np + p] + r0[3 * np + p];
}

__global__ void latent_back_bwd_reduce_kernel(const float* __restrict__ partial, int nparts, int nl,
                                              const vg_latent_back_bwd_args g) {
  const int np = nl * FLOW_LAYER_NPARAM;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  float s = 0.f;
  for (int i = 0; i < nparts; ++i) s += partial[(int64_t)i * np + p];
  const int l = p / FLOW_LAYER_NPARAM, r = p % FLOW_LAYER_NPARAM;
  if (r < 2 * FH) g.d_w1[l * 2 * FH + r] = s;
  else if (r < 3 * FH) g.d_b1[l * FH + (r - 2 * FH)] = s;
  else if (r < 4 * FH) g.d_ln_w[l * FH + (r - 3 * FH)] = s;
  else if (r < 5 * FH) g.d_ln_b[l * FH + (r - 4 * FH)] = s;
  else if (r < 5 * FH + LAT * FH) g.d_w2[l * LAT * FH + (r - 5 * FH)] = s;
  else g.d_b2[l * LAT + (r - 5 * FH - LAT * FH)] = s;
}

// decode: z0 = mean_p + exp(logstd_p)*eps*temperature ; inverse flow, layers reversed
__global__ void __launch_bounds__(128)
latent_prior_sample_kernel(const vg_latent_back_args a, const float* __restrict__ eps, float temperature,
                           float* __restrict__ z_out) {
  __shared__ float sw[FL_MAX * FLOW_LAYER_NPARAM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_flow_params(sw, a);
  __syncthreads();
  const int64_t f = (int64_t)blockIdx.x * 4 + warp;
  if (f >= a.M) return;
  const float* head = a.head + f * a.head_ld;
  float x[LAT];
#pragma unroll
  for (int c = 0; c < LAT; ++c) {
    const float e = eps ? eps[f * LAT + c] : 0.f;
    x[c] = head[c] + expf(head[LAT + c]) * e * temperature;
  }
  for (int l = a.n_layers - 1; l >= 0; --l) {
    FlowLayerW w;
    load_flow_layer(w, sw + l * FLOW_LAYER_NPARAM, lane);
    const float x0[2] = {x[0], x[1]};
    float out[LAT];
    flow_conditioner(w, head + 2 * LAT + l * 2 * FH, lane, a.ln_eps, x0, out, nullptr);
    float x1[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float sg = 1.f / (1.f + expf(-out[2 + c]));
      const float v = sg * (a.scale_lo - a.scale_hi) + a.scale_hi;
      x1[c] = (x[2 + c] - out[c]) / v;
    }
    x[0] = x1[0]; x[1] = x1[1]; x[2] = x0[0]; x[3] = x0[1];   // un-flip
  }
  if (lane < LAT) {
    float xl = x[0];
#pragma unroll
    for (int c = 1; c < LAT; ++c) if (lane == c) xl = x[c];
    z_out[f * LAT + lane] = xl;
  }
}

static int back_blocks(int64_t M) {
  int64_t want = ceil_div(M, 4);
  int64_t cap = kNumSMs * 4;
  return (int)(want < cap ? want : cap);
}

}  // namespace vg

using namespace vg;

static int check_front(const char* fn, const vg_latent_front_args* a) {
  VG_REQUIRE(a, -1, "%s: null args", fn);
  VG_REQUIRE(a->latent_dim == LAT && a->emb_dim == EMB, -3, "%s: only latent_dim=4, emb_dim=64 are built", fn);
  VG_REQUIRE(a->B > 0 && a->T > 0 && a->vocab > 0, -3, "%s: bad shape", fn);
  VG_REQUIRE(valid_dtype(a->act_dtype), -2, "%s: bad dtype", fn);
  VG_REQUIRE(a->h_enc && a->eps && a->ids && a->mask && a->w_mean && a->b_mean && a->w_logstd && a->b_logstd &&
                 a->tok_emb && a->w_fuse && a->b_fuse && a->mean && a->logstd && a->z && a->log_q && a->u && a->u_shift,
             -1, "%s: null pointer", fn);
  return 0;
}

extern "C" int vg_latent_front_fwd(const vg_latent_front_args* a, vg_stream_t stream) {
  if (int rc = check_front("vg_latent_front_fwd", a)) return rc;
  const int64_t M = a->B * a->T;
  const unsigned grid = (unsigned)ceil_div(M * 16, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == VG_F32) latent_front_fwd_kernel<float><<<grid, 256, 0, st>>>(*a);
  else latent_front_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(*a);
  VG_LAUNCH_CHECK("vg_latent_front_fwd");
  return 0;
}

extern "C" size_t vg_latent_front_bwd_workspace(int64_t B, int64_t T, int32_t, int32_t E, int32_t) {
  const int64_t M = B * T;
  return align_up((size_t)front_bwd_blocks(M) * FRONT_NPARAM * sizeof(float), 256) + (size_t)M * E * sizeof(float);
}

extern "C" int vg_latent_front_bwd(const vg_latent_front_bwd_args* g, void* workspace, size_t workspace_bytes,
                                   vg_stream_t stream) {
  VG_REQUIRE(g, -1, "vg_latent_front_bwd: null args");
  if (int rc = check_front("vg_latent_front_bwd", &g->f)) return rc;
  VG_REQUIRE(g->d_h_enc && g->d_w_mean && g->d_b_mean && g->d_w_logstd && g->d_b_logstd && g->d_tok_emb &&
                 g->d_w_fuse && g->d_b_fuse, -1, "vg_latent_front_bwd: null output pointer");
  const vg_latent_front_args& a = g->f;
  const int64_t M = a.B * a.T;
  VG_REQUIRE(workspace && workspace_bytes >= vg_latent_front_bwd_workspace(a.B, a.T, LAT, EMB, a.vocab), -5,
             "vg_latent_front_bwd: workspace too small");
  const int nb = front_bwd_blocks(M);
  float* partial = (float*)workspace;
  float* gu = (float*)((uint8_t*)workspace + align_up((size_t)nb * FRONT_NPARAM * sizeof(float), 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (a.act_dtype == VG_F32) latent_front_bwd_kernel<float><<<nb, FRONT_BWD_THREADS, 0, st>>>(*g, partial, gu);
  else latent_front_bwd_kernel<__nv_bfloat16><<<nb, FRONT_BWD_THREADS, 0, st>>>(*g, partial, gu);
  VG_LAUNCH_CHECK("vg_latent_front_bwd");
  latent_front_bwd_reduce_kernel<<<(FRONT_NPARAM + 127) / 128, 128, 0, st>>>(partial, nb, *g);
  VG_LAUNCH_CHECK("vg_latent_front_bwd(reduce)");
  embedding_grad_kernel<<<a.vocab, 256, 0, st>>>(a.ids, gu, g->d_tok_emb, M, a.vocab);
  VG_LAUNCH_CHECK("vg_latent_front_bwd(embedding)");
  return 0;
}

static int check_back(const char* fn, const vg_latent_back_args* a) {
  VG_REQUIRE(a, -1, "%s: null args", fn);
  VG_REQUIRE(a->latent_dim == LAT && a->hidden == FH, -3, "%s: only latent_dim=4, hidden=64 are built", fn);
  VG_REQUIRE(a->n_layers >= 1 && a->n_layers <= FL_MAX, -3, "%s: n_layers must be in [1,%d]", fn, FL_MAX);
  VG_REQUIRE(a->M > 0 && a->head_ld >= 2 * LAT + a->n_layers * 2 * FH, -3, "%s: bad shape", fn);
  VG_REQUIRE(a->head && a->w1 && a->b1 && a->ln_w && a->ln_b && a->w2 && a->b2, -1, "%s: null pointer", fn);
  return 0;
}

extern "C" size_t vg_latent_back_workspace(int64_t M, int32_t, int32_t, int32_t n_layers) {
  const size_t a = (size_t)back_blocks(M) * sizeof(float);
  const size_t b = (size_t)back_blocks(M) * n_layers * FLOW_LAYER_NPARAM * sizeof(float);
  return a > b ? a : b;
}

extern "C" int vg_latent_back_fwd(const vg_latent_back_args* a, void* workspace, size_t workspace_bytes,
                                  vg_stream_t stream) {
  if (int rc = check_back("vg_latent_back_fwd", a)) return rc;
  VG_REQUIRE(a->z && a->log_q && a->mask && a->log_p && a->y && a->kl_frame && a->kl_sum, -1,
             "vg_latent_back_fwd: null pointer");
  const int nb = back_blocks(a->M);
  VG_REQUIRE(workspace && workspace_bytes >= (size_t)nb * sizeof(float), -5, "vg_latent_back_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  latent_back_fwd_kernel<<<nb, 128, 0, st>>>(*a, (float*)workspace);
  VG_LAUNCH_CHECK("vg_latent_back_fwd");
  sum_partials_kernel<<<1, 32, 0, st>>>((const float*)workspace, nb, a->kl_sum);
  VG_LAUNCH_CHECK("vg_latent_back_fwd(sum)");
  return 0;
}

extern "C" int vg_latent_back_bwd(const vg_latent_back_bwd_args* g, void* workspace, size_t workspace_bytes,
                                  vg_stream_t stream) {
  VG_REQUIRE(g, -1, "vg_latent_back_bwd: null args");
  if (int rc = check_back("vg_latent_back_bwd", &g->f)) return rc;
  const vg_latent_back_args& a = g->f;
  VG_REQUIRE(a.z && a.mask && g->d_log_p && g->d_head && g->d_z && g->d_w1 && g->d_b1 && g->d_ln_w && g->d_ln_b &&
                 g->d_w2 && g->d_b2, -1, "vg_latent_back_bwd: null pointer");
  const int nb = back_blocks(a.M);
  const int np = a.n_layers * FLOW_LAYER_NPARAM;
  VG_REQUIRE(workspace && workspace_bytes >= (size_t)nb * np * sizeof(float), -5,
             "vg_latent_back_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 5 * np * (int)sizeof(float);
  static bool set = false;
  if (!set) {
    VG_CUDA(cudaFuncSetAttribute(latent_back_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 5 * FL_MAX * FLOW_LAYER_NPARAM * (int)sizeof(float)));
    set = true;
  }
  latent_back_bwd_kernel<<<nb, 128, smem, st>>>(*g, (float*)workspace);
  VG_LAUNCH_CHECK("vg_latent_back_bwd");
  latent_back_bwd_reduce_kernel<<<(np + 127) / 128, 128, 0, st>>>((const float*)workspace, nb, a.n_layers, *g);
  VG_LAUNCH_CHECK("vg_latent_back_bwd(reduce)");
  return 0;
}

extern "C" int vg_latent_prior_sample(const vg_latent_back_args* a, const float* eps, float temperature,
                                      float* z_out, vg_stream_t stream) {
  if (int rc = check_back("vg_latent_prior_sample", a)) return rc;
  VG_REQUIRE(z_out, -1, "vg_latent_prior_sample: null output");
  latent_prior_sample_kernel<<<(unsigned)ceil_div(a->M, 4), 128, 0, (cudaStream_t)stream>>>(*a, eps, temperature, z_out);
  VG_LAUNCH_CHECK("vg_latent_prior_sample");
  return 0;
}
