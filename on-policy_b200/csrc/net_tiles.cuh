// net_tiles.cuh -- MLP-base forward / backward on one row tile, and the per-row loss math.
// Used by the rollout kernel (policy_step.cu) and the training kernels (update_mlp.cu, update_gru.cu).
#pragma once
#include "common.cuh"

namespace mappo {

// Device view of mappo_batch_t + constants.
struct BatchDev {
  const float *obs, *share_obs, *actions, *old_logp, *value_preds, *returns, *advantages, *masks, *active_masks,
      *avail, *h0_actor, *h0_critic;
  const int32_t *rows, *seq_first;
  int n_rows, seq_len, n_seq, act_shape, n_avail;
  // evaluate_actions mode (no gradients): per-position outputs, actor -> log-probs [n_rows, as], critic -> values
  float* eval_out;
  int eval_only;
};

struct LossDev {
  float clip, ent_coef, vl_coef, huber_delta;
  int use_clipped_value_loss, use_huber, use_value_active, use_policy_active, use_valuenorm, update_actor;
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
  tile_mm<TR, NJH>(t.x0, n.in_dim, sW + s.fc1_w, s.ld1, 1, H, sW + s.fc1_b, act, t.A[0], tid);
  __syncthreads();
  tile_layernorm<TR>(t.A[0], H, sW + s.ln1_w, sW + s.ln1_b, t.Y[0], t.mean[1], t.rstd[1], t.red, tid);
  for (int l = 0; l < n.layer_n; ++l) {
    tile_mm<TR, NJH>(t.Y[l], H, sW + s.fc2_w[l], s.ldh, 1, H, sW + s.fc2_b[l], act, t.A[l + 1], tid);
    __syncthreads();
    tile_layernorm<TR>(t.A[l + 1], H, sW + s.ln2_w[l], sW + s.ln2_b[l], t.Y[l + 1], t.mean[l + 2], t.rstd[l + 2],
                       t.red, tid);
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

}  // namespace mappo
