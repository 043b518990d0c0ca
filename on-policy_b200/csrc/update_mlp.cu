// update_mlp.cu -- fused training step for feed-forward (MLP) actor / critic nets:
//   index-driven gather -> base forward -> heads -> PPO / value loss -> backward -> per-CTA gradient slot.
// One launch per net per optimiser step; activations never leave the SM.
// Replaces: feed_forward_generator's gather (utils/shared_buffer.py:377-396), policy.evaluate_actions
// (algorithms/r_mappo/algorithm/rMAPPOPolicy.py:88-114), the losses of R_MAPPO.ppo_update / cal_value_loss
// (algorithms/r_mappo/r_mappo.py:52-89, 129-146) and autograd's backward (:146, :160).
#include "net_tiles.cuh"

namespace mappo {

template <int TR>
struct UpdSmem {
  // offsets in floats
  int w, xh0, x0, A[kMaxLayers + 1], Y[kMaxLayers + 1], gA, gB, lg, stats, red, rowid, rowf, sred, total;
};

template <int TR>
__host__ __device__ inline UpdSmem<TR> make_upd_smem(const NetDev& n, const SmemW& s) {
  constexpr int LD = Tile<TR>::LD;
  UpdSmem<TR> u;
  int o = 0;
  const int inT = ((n.in_dim + 3) & ~3) * LD, hT = n.hid * LD;
  const int gT = (n.in_dim > n.hid ? ((n.in_dim + 3) & ~3) : n.hid) * LD;
  u.w = o; o += s.total;
  u.xh0 = o; o += n.use_fn ? inT : 0;
  u.x0 = o; o += inT;
  for (int l = 0; l <= kMaxLayers; ++l) {
    const bool on = l <= n.layer_n;
    u.A[l] = o; o += on ? hT : 0;
    u.Y[l] = o; o += on ? hT : 0;
  }
  u.gA = o; o += gT;
  u.gB = o; o += gT;
  u.lg = o; o += ((n.head_total + 3) & ~3) * LD;
  u.stats = o; o += 2 * (kMaxLayers + 2) * TR;
  u.red = o; o += 8 * TR;
  u.rowid = o; o += TR;
  u.rowf = o; o += 8 * TR;          // per-row scalars: adv, active, v_old, ret, (spare)
  o = (o + 1) & ~1;
  u.sred = o; o += 2 * 6 * 32;      // doubles
  u.total = o;
  return u;
}

template <int TR, int NJH, int NJIN>
__global__ void __launch_bounds__(4 * TR, 1)
update_mlp_kernel(const NetDev n, const float* __restrict__ params, const BatchDev b, const LossDev L,
                  const double* __restrict__ norm_stats, const double* __restrict__ adv_stats,
                  const float* __restrict__ vn_state, float* __restrict__ grad_part, double* __restrict__ loss_out,
                  int n_tiles) {
  constexpr int LD = Tile<TR>::LD;
  constexpr int NT = Tile<TR>::NT;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const SmemW s = make_smem_w(n, false);
  const UpdSmem<TR> u = make_upd_smem<TR>(n, s);
  float* sW = smem + u.w;
  BaseTiles<TR> t;
  t.xh0 = smem + u.xh0;
  t.x0 = smem + u.x0;
  for (int l = 0; l <= kMaxLayers; ++l) { t.A[l] = smem + u.A[l]; t.Y[l] = smem + u.Y[l]; }
  for (int l = 0; l < kMaxLayers + 2; ++l) { t.mean[l] = smem + u.stats + 2 * l * TR; t.rstd[l] = t.mean[l] + TR; }
  t.red = smem + u.red;
  float* gA = smem + u.gA;
  float* gB = smem + u.gB;
  float* lgT = smem + u.lg;
  int* rowid = reinterpret_cast<int*>(smem + u.rowid);
  float* rowf = smem + u.rowf;
  double* sred = reinterpret_cast<double*>(smem + u.sred);

  load_weights(sW, s, n, params, false, tid, NT);
  float* g = b.eval_only ? nullptr : grad_part + (size_t)blockIdx.x * n.g.total;
  if (g) for (int i = tid; i < n.g.total; i += NT) g[i] = 0.f;

  // normalisers (r_mappo.py:83-86, 134-139; act.py:173-176) and advantage statistics (r_mappo.py:183-187)
  const double sum_active = norm_stats[0], n_rows_d = norm_stats[3];
  float adv_mean = 0.f, adv_inv = 1.f;
  if (adv_stats) {
    const double cnt = adv_stats[2] > 0.0 ? adv_stats[2] : 1.0;
    const double m = adv_stats[0] / cnt;
    double var = adv_stats[1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    adv_mean = (float)m;
    adv_inv = (float)(1.0 / (sqrt(var) + 1e-5));
  }
  float vmean = 0.f, vvar = 1.f;
  if (n.is_critic && L.use_valuenorm && vn_state) vn_mean_var(vn_state, vmean, vvar);
  const float vrs = 1.0f / sqrtf(vvar);
  const int H = n.hid;
  const int Atot = n.head_total;
  double acc[3] = {0.0, 0.0, 0.0};      // critic: value_loss | actor: policy_loss, entropy, ratio

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < TR) {
      const int p = tile * TR + tid;
      rowid[tid] = p < b.n_rows ? (b.rows ? b.rows[p] : p) : -1;
    }
    __syncthreads();
    load_rows_T<TR>(n.is_critic ? b.share_obs : b.obs, n.in_dim, rowid, t.x0, tid);
    base_forward<TR, NJH>(n, s, sW, t, tid);
    const float* feat = t.Y[n.layer_n];
    tile_mm<TR, 2>(feat, H, sW + s.head_w, s.ldh, 1, Atot, sW + s.head_b, ACT_NONE, lgT, tid);
    __syncthreads();

    // ---- per-row loss and d(loss)/d(logits | value), in place in lgT ----
    if (tid < TR) {
      const int r = tid, gr = rowid[r];
      if (gr < 0) {
        for (int j = 0; j < Atot; ++j) lgT[j * LD + r] = 0.f;
      } else if (n.is_critic) {
        const float act = b.active_masks[gr];
        const float w = L.use_value_active ? (float)((double)act / sum_active) : (float)(1.0 / n_rows_d);
        const float v = lgT[r], vo = b.value_preds[gr];
        const float ret = b.returns[gr];
        const float target = L.use_valuenorm ? (ret - vmean) * vrs : ret;
        const float d = v - vo;
        const float vclip = vo + fminf(fmaxf(d, -L.clip), L.clip);
        const float eo = target - v, ec = target - vclip;
        float lo, lc, go, gc;                    // loss and d loss / d e
        if (L.use_huber) {                       // utils/util.py:23-26
          const float dl = L.huber_delta;
          lo = fabsf(eo) <= dl ? 0.5f * eo * eo : dl * (fabsf(eo) - 0.5f * dl);
          lc = fabsf(ec) <= dl ? 0.5f * ec * ec : dl * (fabsf(ec) - 0.5f * dl);
          go = fabsf(eo) <= dl ? eo : copysignf(dl, eo);
          gc = fabsf(ec) <= dl ? ec : copysignf(dl, ec);
        } else {                                 // utils/util.py:28-29
          lo = 0.5f * eo * eo; lc = 0.5f * ec * ec; go = eo; gc = ec;
        }
        float l = lo, dv = -go;
        if (L.use_clipped_value_loss) {          // torch.max: gradient to the larger, 1/2 - 1/2 on ties
          const float inclip = (d >= -L.clip && d <= L.clip) ? 1.f : 0.f;
          const float dvc = -gc * inclip;
          if (lc > lo) { l = lc; dv = dvc; }
          else if (lc == lo) { dv = 0.5f * (dv + dvc); }
        }
        acc[0] += (double)(l * w);
        if (b.eval_out) b.eval_out[tile * TR + r] = v;
        lgT[r] = dv * w * L.vl_coef;
      } else {
        const float act = b.active_masks[gr];
        const float w = L.use_policy_active ? (float)((double)act / sum_active) : (float)(1.0 / n_rows_d);
        float adv = b.advantages[gr];
        adv = (adv - adv_mean) * adv_inv;
        const float* av = (b.avail && n.n_heads == 1) ? b.avail + (size_t)gr * b.n_avail : nullptr;
        const float inv_heads = 1.0f / (float)n.n_heads;
        int off = 0;
        for (int k = 0; k < n.n_heads; ++k) {
          const int A = n.head_dim[k];
          float lse;
          head_lse<LD>(lgT, off, A, r, av, lse);
          const int a = (int)b.actions[(size_t)gr * b.act_shape + k];
          float ent = 0.f, lp_a = 0.f;
          for (int j = 0; j < A; ++j) {
            float lgt = lgT[(off + j) * LD + r];
            if (av && av[j] == 0.f) lgt = -1e10f;
            const float lp = lgt - lse;
            const float p = expf(lp);
            ent = fmaf(-p, lp, ent);
            if (j == a) lp_a = lp;
          }
          if (b.eval_out) b.eval_out[(size_t)(tile * TR + r) * b.act_shape + k] = lp_a;
          const float ratio = expf(lp_a - b.old_logp[(size_t)gr * b.act_shape + k]);      // r_mappo.py:129
          const float s1 = ratio * adv;
          const float s2 = fminf(fmaxf(ratio, 1.f - L.clip), 1.f + L.clip) * adv;
          const float mn = fminf(s1, s2);
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
            const float p = expf(lp);
            float dl = dlp * ((j == a ? 1.f : 0.f) - p) + dH * (-p * (lp + ent));
            if (masked || !L.update_actor) dl = 0.f;
            lgT[(off + j) * LD + r] = dl;
          }
          off += A;
        }
      }
    }
    __syncthreads();
    if (b.eval_only) continue;

    // ---- head backward ----
    tile_colsum<TR>(lgT, Atot, g + n.g.head_b, tid);
    tile_dw<TR, 2, NJH>(lgT, Atot, feat, H, g + n.g.head_w, H, tid);        // NI=2: Atot <= 32 with NTY=16
    tile_mm<TR, NJH>(lgT, Atot, sW + s.head_w, 1, s.ldh, H, nullptr, ACT_NONE, gA, tid);
    __syncthreads();
    base_backward<TR, NJH, NJIN>(n, s, sW, t, gA, gB, g, tid);
  }

  // loss scalars: [0] value_loss [1] policy_loss [2] dist_entropy [5] ratio
  double v[3];
  if (n.is_critic) { v[0] = acc[0]; v[1] = 0.0; v[2] = 0.0; }
  else { v[0] = acc[0]; v[1] = acc[1]; v[2] = acc[2] / (n_rows_d * (double)b.act_shape); }
  __syncthreads();
  if (n.is_critic) {
    double one[1] = {v[0]};
    block_accumulate<1>(one, loss_out + 0, sred, tid, NT);
  } else {
    double two[2] = {v[0], v[1]};
    block_accumulate<2>(two, loss_out + 1, sred, tid, NT);
    double rt[1] = {v[2]};
    block_accumulate<1>(rt, loss_out + 5, sred, tid, NT);
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
constexpr int kUpdTR = 64;

static int upd_grid(int n_rows, int sm) {
  const int n_tiles = (n_rows + kUpdTR - 1) / kUpdTR;
  return n_tiles < 2 * sm ? n_tiles : 2 * sm;
}

int update_mlp_slots(const NetDev& n, int n_rows, int sm_count) { return upd_grid(n_rows, sm_count); }

template <int NJIN>
static int launch_upd(const NetDev& n, const float* params, const BatchDev& b, const LossDev& L,
                      const double* norm_stats, const double* adv_stats, const float* vn_state, float* grad_part,
                      int n_slots, double* loss_out, cudaStream_t st) {
  const SmemW s = make_smem_w(n, false);
  const UpdSmem<kUpdTR> u = make_upd_smem<kUpdTR>(n, s);
  const size_t bytes = (size_t)u.total * sizeof(float);
  if (bytes > 227 * 1024) {
    set_error("update_mlp: net needs %zu B of shared memory per CTA (> 227 KB): in_dim=%d hidden=%d layer_N=%d",
              bytes, n.in_dim, n.hid, n.layer_n);
    return MAPPO_ERR_UNSUPPORTED;
  }
  auto kern = update_mlp_kernel<kUpdTR, 4, NJIN>;
  static thread_local size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return check_launch("update_mlp: cudaFuncSetAttribute");
    configured = bytes;
  }
  const int n_tiles = (b.n_rows + kUpdTR - 1) / kUpdTR;
  kern<<<n_slots, 4 * kUpdTR, bytes, st>>>(n, params, b, L, norm_stats, adv_stats, vn_state, grad_part, loss_out,
                                           n_tiles);
  return check_launch("update_mlp_kernel");
}

int update_mlp_launch(const NetDev& n, const float* params, const BatchDev& b, const LossDev& L,
                      const double* norm_stats, const double* adv_stats, const float* vn_state, float* grad_part,
                      int n_slots, double* loss_out, cudaStream_t st) {
  if (n.hid != 64) {
    set_error("update_mlp: hidden_size %d not built in the fused SIMT path (64 only)", n.hid);
    return MAPPO_ERR_UNSUPPORTED;
  }
  if (n.head_total > 32) { set_error("update_mlp: sum(head_dim)=%d > 32", n.head_total); return MAPPO_ERR_UNSUPPORTED; }
  if (n.in_dim <= 64)
    return launch_upd<4>(n, params, b, L, norm_stats, adv_stats, vn_state, grad_part, n_slots, loss_out, st);
  if (n.in_dim <= 128)
    return launch_upd<8>(n, params, b, L, norm_stats, adv_stats, vn_state, grad_part, n_slots, loss_out, st);
  set_error("update_mlp: in_dim %d > 128 not built in the fused SIMT path", n.in_dim);
  return MAPPO_ERR_UNSUPPORTED;
}

}  // namespace mappo
