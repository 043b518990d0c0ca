// net_tiles.cuh -- MLP-base forward / backward on one row tile, and the per-row loss math.
// Used by the rollout kernel (policy_step.cu) and the training kernels (update_mlp.cu, update_gru.cu).
#pragma once
#include "common.cuh"

namespace mappo {

// Device view of mappo_batch_t + constants.
struct BatchDev {
  const float *obs, *share_obs, *actions, *old_logp, *value_preds, *returns, *advantages, *masks, *active_masks,
      *avail, *h0_actor, *h0_critic;
  const float* factor;         // [.,1] HAPPO importance factor of the row (separated_buffer.py:62-63), NULL = 1
  const int32_t *rows, *seq_first;
  int n_rows, seq_len, n_seq, act_shape, n_avail;
  // evaluate_actions mode (no gradients): per-position outputs, actor -> log-probs [n_rows, as], critic -> values
  float* eval_out;
  int eval_only;
};

struct LossDev {
  float clip, ent_coef, vl_coef, huber_delta;
  int use_clipped_value_loss, use_huber, use_value_active, use_policy_active, use_valuenorm, update_actor;
  int happo;                   // actor loss of algorithms/happo/happo_trainer.py:129-141 (joint ratio x factor)
};

// gather TR rows of `dim` floats into a transposed tile; rowid < 0 -> zeros
template <int TR>
__device__ __forceinline__ void load_rows_T(const float* __restrict__ src, int dim, const int* __restrict__ rowid,
                                            float* __restrict__ XT, int tid) {
  constexpr int LD = Tile<TR>::LD;
  const int n = TR * dim;
  for (int i = tid; i < n; i += Tile<TR>::NT) {
    const int r = i / dim, c = i - r * dim;
    const int g = rowid[r];
    XT[c * LD + r] = g >= 0 ? __ldg(src + (size_t)g * dim + c) : 0.f;
  }
}

// Pointers of the activation tiles an MLP base needs.  Training keeps every A/Y; the rollout aliases.
template <int TR> struct BaseTiles {
  float* xh0;                       // [in][LD]  normalised input (pre-affine); unused without feature norm
  float* x0;                        // [in][LD]  fc1 input
  float* A[kMaxLayers + 1];         // [H][LD]   act(Linear(.))           (LayerNorm input)
  float* Y[kMaxLayers + 1];         // [H][LD]   LayerNorm output         (next layer input)
  float* mean[kMaxLayers + 2];      // [TR]      index 0 = feature norm, 1.. = layer LNs
  float* rstd[kMaxLayers + 2];
  float* red;                       // [8*TR]
  bool keep_act;                    // training: keep act(.) tiles A[l] for the backward pass
};

// base.feature_norm + base.mlp (mlp.py:26-30, 52-57) on the tile already loaded in t.x0 (raw input).
// NJH = H/16.  Leaves the features in t.Y[layer_n].  Starts and ends with a barrier.
template <int TR, int NJH>
__device__ __forceinline__ void base_forward(const NetDev& n, const SmemW& s, const float* __restrict__ sW,
                                             const BaseTiles<TR>& t, int tid) {
  constexpr int LD = Tile<TR>::LD;
  const int H = n.hid, act = n.use_relu ? ACT_RELU : ACT_TANH;
  __syncthreads();
  if (n.use_fn) {
    tile_layernorm<TR>(t.x0, n.in_dim, nullptr, nullptr, t.xh0, t.mean[0], t.rstd[0], t.red, tid);
    const int cnt = n.in_dim * TR;
    for (int i = tid; i < cnt; i += Tile<TR>::NT) {
      const int c = i / TR, r = i - c * TR;
      t.x0[c * LD + r] = fmaf(t.xh0[c * LD + r], sW[s.fn_w + c], sW[s.fn_b + c]);
    }
    __syncthreads();
  }
  // Linear + act + LayerNorm fused per layer (row statistics by warp shuffles); A / mean / rstd are kept for backward
  // when the caller's tiles do not alias (training), and are simply scratch for the rollout.
  tile_mm_ln<TR, NJH>(t.x0, n.in_dim, sW + s.fc1_w, s.ld1, sW + s.fc1_b, act, sW + s.ln1_w, sW + s.ln1_b,
                      t.keep_act ? t.A[0] : nullptr, t.Y[0], t.mean[1], t.rstd[1], tid);
  __syncthreads();
  for (int l = 0; l < n.layer_n; ++l) {
    tile_mm_ln<TR, NJH>(t.Y[l], H, sW + s.fc2_w[l], s.ldh, sW + s.fc2_b[l], act, sW + s.ln2_w[l], sW + s.ln2_b[l],
                        t.keep_act ? t.A[l + 1] : nullptr, t.Y[l + 1], t.mean[l + 2], t.rstd[l + 2], tid);
    __syncthreads();
  }
}

// Backward of base_forward.  dY (grad w.r.t. t.Y[layer_n]) is in gA on entry; gB is a second scratch
// tile of the same size.  Parameter gradients are accumulated into the CTA's slot `g` (global).
template <int TR, int NJH, int NJIN>
__device__ __forceinline__ void base_backward(const NetDev& n, const SmemW& s, const float* __restrict__ sW,
                                              const BaseTiles<TR>& t, float* gA, float* gB, float* __restrict__ g,
                                              int tid) {
  constexpr int NI = 16 * NJH / Tile<TR>::NTY;
  const int H = n.hid, act = n.use_relu ? ACT_RELU : ACT_TANH;
  float* cur = gA;
  float* oth = gB;
  for (int l = n.layer_n; l >= 1; --l) {
    tile_ln_param_grads<TR>(cur, t.A[l], t.mean[l + 1], t.rstd[l + 1], H, g + n.g.ln2_w[l - 1], g + n.g.ln2_b[l - 1],
                            tid);
    tile_layernorm_bwd<TR>(cur, t.A[l], t.mean[l + 1], t.rstd[l + 1], sW + s.ln2_w[l - 1], H, act, t.red, tid);
    tile_colsum<TR>(cur, H, g + n.g.fc2_b[l - 1], tid);
    tile_dw<TR, NI, NJH>(cur, H, t.Y[l - 1], H, g + n.g.fc2_w[l - 1], H, tid);
    tile_mm<TR, NJH>(cur, H, sW + s.fc2_w[l - 1], 1, s.ldh, H, nullptr, ACT_NONE, oth, tid);
    __syncthreads();
    float* tmp = cur; cur = oth; oth = tmp;
  }
  tile_ln_param_grads<TR>(cur, t.A[0], t.mean[1], t.rstd[1], H, g + n.g.ln1_w, g + n.g.ln1_b, tid);
  tile_layernorm_bwd<TR>(cur, t.A[0], t.mean[1], t.rstd[1], sW + s.ln1_w, H, act, t.red, tid);
  tile_colsum<TR>(cur, H, g + n.g.fc1_b, tid);
  tile_dw<TR, NI, NJIN>(cur, H, t.x0, n.in_dim, g + n.g.fc1_w, n.in_dim, tid);
  if (n.use_fn) {
    // dX0 = dZ1 * W1, then feature_norm.{weight,bias} gradients (xh0 is already normalised: mean 0 / rstd 1)
    tile_mm<TR, NJIN>(cur, H, sW + s.fc1_w, 1, s.ld1, n.in_dim, nullptr, ACT_NONE, oth, tid);
    __syncthreads();
    constexpr int LD = Tile<TR>::LD;
    for (int c = tid; c < n.in_dim; c += Tile<TR>::NT) {
      float sw = 0.f, sb = 0.f;
#pragma unroll 4
      for (int r = 0; r < TR; ++r) {
        const float d = oth[c * LD + r];
        sw = fmaf(d, t.xh0[c * LD + r], sw);
        sb += d;
      }
      g[n.g.fn_w + c] += sw;
      g[n.g.fn_b + c] += sb;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// per-row head math
// ---------------------------------------------------------------------------------------------
// log-softmax pieces of one head held in the logits tile: returns max and log-sum-exp.
template <int LD>
__device__ __forceinline__ void head_lse(const float* __restrict__ lgT, int off, int A, int r,
                                         const float* __restrict__ avail_row, float& lse) {
  float mx = -INFINITY;
  for (int j = 0; j < A; ++j) {
    float l = lgT[(off + j) * LD + r];
    if (avail_row && avail_row[j] == 0.f) l = -1e10f;        // distributions.py:66-67
    mx = fmaxf(mx, l);
  }
  float se = 0.f;
  for (int j = 0; j < A; ++j) {
    float l = lgT[(off + j) * LD + r];
    if (avail_row && avail_row[j] == 0.f) l = -1e10f;
    se += expf(l - mx);
  }
  lse = mx + logf(se);
}

// ---------------------------------------------------------------------------------------------
// loss math of one row (r_mappo.py:52-89, 129-146; act.py:147-176), shared by the MLP and GRU kernels
// ---------------------------------------------------------------------------------------------
struct LossConsts {
  double sum_active, n_rows_d;
  double inv_sum_active, inv_n_rows;      // reciprocals formed once per kernel (a row then needs one DMUL, not a DDIV)
  float adv_mean, adv_inv, vmean, vrs;
};

__device__ __forceinline__ LossConsts make_loss_consts(const NetDev& n, const LossDev& L,
                                                       const double* __restrict__ norm_stats,
                                                       const double* __restrict__ adv_stats,
                                                       const float* __restrict__ vn_state) {
  LossConsts c;
  c.sum_active = norm_stats[0];
  c.n_rows_d = norm_stats[3];
  c.inv_sum_active = 1.0 / c.sum_active;
  c.inv_n_rows = 1.0 / c.n_rows_d;
  c.adv_mean = 0.f;
  c.adv_inv = 1.f;
  if (adv_stats) {                               // r_mappo.py:183-187: stats over active entries, applied to all
    const double cnt = adv_stats[2] > 0.0 ? adv_stats[2] : 1.0;
    const double m = adv_stats[0] / cnt;
    double var = adv_stats[1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    c.adv_mean = (float)m;
    c.adv_inv = (float)(1.0 / (sqrt(var) + 1e-5));
  }
  float vmean = 0.f, vvar = 1.f;
  if (n.is_critic && L.use_valuenorm && vn_state) vn_mean_var(vn_state, vmean, vvar);
  c.vmean = vmean;
  c.vrs = 1.0f / sqrtf(vvar);
  return c;
}

// The per-row scalars the loss needs, loadable EARLY (right after the row id is known) so their global-memory latency
// hides behind the forward pass instead of sitting between the logits and the gradient.
struct RowIn {
  float active, v_old, ret, adv, factor;
  float action[kMaxHeads], old_logp[kMaxHeads];
};
// load that the compiler may not sink towards its first use (so it is issued where it is written)
__device__ __forceinline__ float ldg_pinned(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ RowIn load_row_in(const NetDev& n, const BatchDev& b, int gr) {
  RowIn q;
  q.active = q.v_old = q.ret = q.adv = 0.f;
  q.factor = 1.f;
#pragma unroll
  for (int k = 0; k < kMaxHeads; ++k) q.action[k] = q.old_logp[k] = 0.f;
  if (gr < 0) return q;
  q.active = ldg_pinned(b.active_masks + gr);
  if (n.is_critic) {
    q.v_old = ldg_pinned(b.value_preds + gr);
    q.ret = ldg_pinned(b.returns + gr);
  } else {
    q.adv = ldg_pinned(b.advantages + gr);
    if (b.factor) q.factor = ldg_pinned(b.factor + gr);
#pragma unroll
    for (int k = 0; k < kMaxHeads; ++k)
      if (k < n.n_heads) {
        q.action[k] = ldg_pinned(b.actions + (size_t)gr * b.act_shape + k);
        q.old_logp[k] = ldg_pinned(b.old_logp + (size_t)gr * b.act_shape + k);
      }
  }
  return q;
}

// Row r of the logits tile lgT (storage row gr, minibatch position p): accumulates the loss terms into acc
// (critic: [0] value_loss; actor: [0] policy_loss, [1] entropy, [2] sum of ratios) and overwrites the logits /
// value with d(loss)/d(logit | value).
template <int LD>
__device__ __forceinline__ void row_loss_pre(const NetDev& n, const BatchDev& b, const LossDev& L, const LossConsts& c,
                                             float* __restrict__ lgT, int r, int gr, int p, const RowIn& q,
                                             double (&acc)[3]) {
  const int Atot = n.head_total;
  if (gr < 0) {
    for (int j = 0; j < Atot; ++j) lgT[j * LD + r] = 0.f;
    return;
  }
  if (n.is_critic) {
    const float act = q.active;
    const float w = L.use_value_active ? (float)((double)act * c.inv_sum_active) : (float)c.inv_n_rows;
    const float v = lgT[r], vo = q.v_old;
    const float ret = q.ret;
    const float target = L.use_valuenorm ? (ret - c.vmean) * c.vrs : ret;       // valuenorm.py:57-66
    const float d = v - vo;
    const float vclip = vo + fminf(fmaxf(d, -L.clip), L.clip);                  // r_mappo.py:62-63
    const float eo = target - v, ec = target - vclip;
    float lo, lcl, go, gc;                                                      // loss and d loss / d e
    if (L.use_huber) {                                                          // utils/util.py:23-26
      const float dl = L.huber_delta;
      lo = fabsf(eo) <= dl ? 0.5f * eo * eo : dl * (fabsf(eo) - 0.5f * dl);
      lcl = fabsf(ec) <= dl ? 0.5f * ec * ec : dl * (fabsf(ec) - 0.5f * dl);
      go = fabsf(eo) <= dl ? eo : copysignf(dl, eo);
      gc = fabsf(ec) <= dl ? ec : copysignf(dl, ec);
    } else {                                                                    // utils/util.py:28-29
      lo = 0.5f * eo * eo; lcl = 0.5f * ec * ec; go = eo; gc = ec;
    }
    float l = lo, dv = -go;
    if (L.use_clipped_value_loss) {               // torch.max: gradient to the larger, 1/2 - 1/2 on ties
      const float inclip = (d >= -L.clip && d <= L.clip) ? 1.f : 0.f;
      const float dvc = -gc * inclip;
      if (lcl > lo) { l = lcl; dv = dvc; }
      else if (lcl == lo) { dv = 0.5f * (dv + dvc); }
    }
    acc[0] += (double)(l * w);
    if (b.eval_out) b.eval_out[p] = v;
    lgT[r] = dv * w * L.vl_coef;
    return;
  }
  const float act = q.active;
  const float w = L.use_policy_active ? (float)((double)act * c.inv_sum_active) : (float)c.inv_n_rows;
  const float adv = (q.adv - c.adv_mean) * c.adv_inv;
  const float* av = (b.avail && n.n_heads == 1) ? b.avail + (size_t)gr * b.n_avail : nullptr;
  const float inv_heads = 1.0f / (float)n.n_heads;
  if (L.happo) {
    // HAPPO (happo_trainer.py:129-141): ONE importance weight per row, R = prod_k exp(logp_k - old_logp_k), and the row's
    // factor f multiplies min(R adv, clamp(R) adv).  Pass 1: per-head log-sum-exp, entropy, log-prob of the stored action.
    float lse_k[kMaxHeads], ent_k[kMaxHeads], R = 1.f;
    int off1 = 0;
#pragma unroll
    for (int k = 0; k < kMaxHeads; ++k) {
      if (k < n.n_heads) {
        const int A = n.head_dim[k];
        head_lse<LD>(lgT, off1, A, r, av, lse_k[k]);
        const int a = (int)q.action[k];
        float ent = 0.f, lp_a = 0.f;
        for (int j = 0; j < A; ++j) {
          float lgt = lgT[(off1 + j) * LD + r];
          if (av && av[j] == 0.f) lgt = -1e10f;
          const float lp = lgt - lse_k[k];
          ent = fmaf(-expf(lp), lp, ent);
          if (j == a) lp_a = lp;
        }
        ent_k[k] = ent;
        if (b.eval_out) b.eval_out[(size_t)p * b.act_shape + k] = lp_a;
        R *= expf(lp_a - q.old_logp[k]);
        acc[1] += (double)(ent * w * inv_heads);
        off1 += A;
      }
    }
    const float s1 = R * adv, s2 = fminf(fmaxf(R, 1.f - L.clip), 1.f + L.clip) * adv;
    const float mn = fminf(s1, s2);
    const bool inr = R >= 1.f - L.clip && R <= 1.f + L.clip;
    const float dm = inr ? adv : (s1 < s2 ? adv : (s1 == s2 ? 0.5f * adv : 0.f));
    const float dlp = -w * q.factor * dm * R;               // d loss / d logp_k: the same for every head (dR / dlogp_k = R)
    const float dH = -L.ent_coef * w * inv_heads;
    acc[0] += (double)(-q.factor * mn * w);
    acc[2] += (double)R * (double)n.n_heads;                // reported as imp_weights.mean() over [rows, 1]
    int off2 = 0;
#pragma unroll
    for (int k = 0; k < kMaxHeads; ++k) {
      if (k < n.n_heads) {
        const int A = n.head_dim[k], a = (int)q.action[k];
        for (int j = 0; j < A; ++j) {
          float lgt = lgT[(off2 + j) * LD + r];
          const bool masked = av && av[j] == 0.f;
          if (masked) lgt = -1e10f;
          const float lp = lgt - lse_k[k];
          const float pj = expf(lp);
          float dl = dlp * ((j == a ? 1.f : 0.f) - pj) + dH * (-pj * (lp + ent_k[k]));
          if (masked || !L.update_actor) dl = 0.f;
          lgT[(off2 + j) * LD + r] = dl;
        }
        off2 += A;
      }
    }
    return;
  }
  int off = 0;
  for (int k = 0; k < n.n_heads; ++k) {
    const int A = n.head_dim[k];
    float lse;
    head_lse<LD>(lgT, off, A, r, av, lse);
    const int a = (int)q.action[k];
    float ent = 0.f, lp_a = 0.f;
    for (int j = 0; j < A; ++j) {
      float lgt = lgT[(off + j) * LD + r];
      if (av && av[j] == 0.f) lgt = -1e10f;
      const float lp = lgt - lse;
      const float pj = expf(lp);
      ent = fmaf(-pj, lp, ent);
      if (j == a) lp_a = lp;
    }
    if (b.eval_out) b.eval_out[(size_t)p * b.act_shape + k] = lp_a;
    const float ratio = expf(lp_a - q.old_logp[k]);                                    // r_mappo.py:129
    const float s1 = ratio * adv;
    const float s2 = fminf(fmaxf(ratio, 1.f - L.clip), 1.f + L.clip) * adv;
    const float mn = fminf(s1, s2);
    // d min(s1, s2) / d ratio with torch.min / clamp tie semantics
    const bool inr = ratio >= 1.f - L.clip && ratio <= 1.f + L.clip;
    const float dm = inr ? adv : (s1 < s2 ? adv : (s1 == s2 ? 0.5f * adv : 0.f));
    const float dlp = -w * dm * ratio;
    const float dH = -L.ent_coef * w * inv_heads;
    acc[0] += (double)(-mn * w);
    acc[1] += (double)(ent * w * inv_heads);
    acc[2] += (double)ratio;
    for (int j = 0; j < A; ++j) {
      float lgt = lgT[(off + j) * LD + r];
      const bool masked = av && av[j] == 0.f;
      if (masked) lgt = -1e10f;
      const float lp = lgt - lse;
      const float pj = expf(lp);
      float dl = dlp * ((j == a ? 1.f : 0.f) - pj) + dH * (-pj * (lp + ent));
      if (masked || !L.update_actor) dl = 0.f;
      lgT[(off + j) * LD + r] = dl;
    }
    off += A;
  }
}


template <int LD>
__device__ __forceinline__ void row_loss(const NetDev& n, const BatchDev& b, const LossDev& L, const LossConsts& c,
                                         float* __restrict__ lgT, int r, int gr, int p, double (&acc)[3]) {
  const RowIn q = load_row_in(n, b, gr);
  row_loss_pre<LD>(n, b, L, c, lgT, r, gr, p, q, acc);
}

}  // namespace mappo
