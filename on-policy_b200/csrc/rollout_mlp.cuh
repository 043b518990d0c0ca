// rollout_mlp.cuh -- rollout inference of feed-forward (MLP) policies: ONE WARP OWNS ONE ROW END TO END.
// (included by policy_step.cu only; the GRU policies keep the tile path there.)
//
// Rows of a rollout step never interact, and a 64-wide MLP row fits one warp: lane l owns hidden columns l and l + 32.
//   * LayerNorm statistics are warp shuffles,
//   * a layer's activations travel through a 64-float scratch that only this warp touches: __syncwarp(), never a CTA
//     barrier; the warp reads them back as broadcast LDS.128 (4 k per load),
//   * the weights are packed as [k][lane][2] so one conflict-free LDS.64 feeds the lane's 2 FMAs of a k
//     (mappo_pack_rollout_weights builds this image for non-recurrent nets; a CTA fetches it with one TMA bulk copy),
//   * every dot product accumulates k = 0, 1, 2, ... sequentially in one register, exactly like tile_mm / tile_mm_ln,
//   * sampling is parallel over the actions of a head (lane off + j owns action j) but adds the softmax denominator in
//     the serial j = 0..A-1 order of the tile path, so the argmax(p / Exp(1)) draw sees the same numbers.
// A CTA = 4 warps = 4 rows, so E rows spread over E/4 CTAs per net (c2: 96 + 96 CTAs on 148 SMs, ~1 warp per scheduler:
// the per-step latency chain of a row is what is left).  In the persistent rollout each warp walks t = 0..T with the NEXT
// step's row already in flight (prefetched into registers), so no global-memory latency sits between two steps.
#pragma once
#include "net_tiles.cuh"
#include "launch_args.h"

namespace mappo {

constexpr int kFR = 4;              // rows (= warps) per CTA
constexpr int kFT = 32 * kFR;       // threads per CTA
constexpr int kFWarpScratch = 128;  // floats per warp: two 64-float activation buffers

struct FastImg {
  int fn_w, fn_b, w1, b1, g1, be1, K1;       // K1 = in_dim padded to a multiple of 4 (zero rows)
  int w2[kMaxLayers], b2[kMaxLayers], g2[kMaxLayers], be2[kMaxLayers];
  int wh, bh, AP, total;
};
__host__ __device__ inline FastImg make_fast_img(const NetDev& n) {
  FastImg f;
  int o = 0;
  f.K1 = (n.in_dim + 3) & ~3;
  f.fn_w = o; o += n.use_fn ? f.K1 : 0;
  f.fn_b = o; o += n.use_fn ? f.K1 : 0;
  f.w1 = o; o += f.K1 * 64;
  f.b1 = o; o += 64; f.g1 = o; o += 64; f.be1 = o; o += 64;
  for (int l = 0; l < kMaxLayers; ++l) {
    f.w2[l] = f.b2[l] = f.g2[l] = f.be2[l] = 0;
    if (l < n.layer_n) {
      f.w2[l] = o; o += 64 * 64;
      f.b2[l] = o; o += 64; f.g2[l] = o; o += 64; f.be2[l] = o; o += 64;
    }
  }
  f.AP = (n.head_total + 3) & ~3;
  f.wh = o; o += 64 * f.AP;           // heads transposed: [k][AP]
  f.bh = o; o += f.AP;
  f.total = o;                        // a multiple of 4 floats (TMA bulk copies move 16-byte units)
  return f;
}
__host__ __device__ inline bool fast_rollout_supported(const NetDev& n) {
  return !n.recurrent && n.hid == 64 && n.in_dim <= 64 && n.head_total <= 32;
}

// element i of the image (also the front part of the recurrent nets' image, rollout_gru.cuh)
__device__ __forceinline__ float pack_fast_element(const NetDev& n, const FastImg& f, const float* __restrict__ p, int i) {
  const mappo_net_layout_t& g = n.g;
  const int in = n.in_dim;
  {
    float v = 0.f;
    if (i < f.w1) {                                       // feature-norm affine (only present when use_fn)
      const int t = i - f.fn_w;
      if (t < f.K1) { if (t < in) v = p[g.fn_w + t]; }
      else if (t - f.K1 < in) v = p[g.fn_b + t - f.K1];
    } else if (i < f.b1) {                                // fc1 as [k][lane][j]: element = W1[lane + 32 j][k]
      const int t = i - f.w1, k = t >> 6, ln = (t >> 1) & 31, j = t & 1;
      if (k < in) v = p[g.fc1_w + (ln + 32 * j) * in + k];
    } else if (i < f.g1) v = p[g.fc1_b + i - f.b1];
    else if (i < f.be1) v = p[g.ln1_w + i - f.g1];
    else if (i < f.be1 + 64) v = p[g.ln1_b + i - f.be1];
    else if (i < f.wh) {
      for (int l = 0; l < n.layer_n; ++l) {
        if (i >= f.w2[l] && i < f.b2[l]) {
          const int t = i - f.w2[l], k = t >> 6, ln = (t >> 1) & 31, j = t & 1;
          v = p[g.fc2_w[l] + (ln + 32 * j) * 64 + k];
        } else if (i >= f.b2[l] && i < f.g2[l]) v = p[g.fc2_b[l] + i - f.b2[l]];
        else if (i >= f.g2[l] && i < f.be2[l]) v = p[g.ln2_w[l] + i - f.g2[l]];
        else if (i >= f.be2[l] && i < f.be2[l] + 64) v = p[g.ln2_b[l] + i - f.be2[l]];
      }
    } else if (i < f.bh) {
      const int t = i - f.wh, k = t / f.AP, a = t - k * f.AP;
      if (a < n.head_total) v = p[g.head_w + a * 64 + k];
    } else {
      const int a = i - f.bh;
      if (a < n.head_total) v = p[g.head_b + a];
    }
    return v;
  }
}
__global__ void __launch_bounds__(256) pack_fast_kernel(const NetDev n, const float* __restrict__ p, float* __restrict__ img) {
  const FastImg f = make_fast_img(n);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f.total; i += gridDim.x * blockDim.x) img[i] = pack_fast_element(n, f, p, i);
}

__device__ __forceinline__ float warp_sum(float s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Y = LayerNorm(act(X W^T + b)) * gamma + beta for the warp's row: X [K4] -> Y [64] in the warp's scratch.
// K4 is a multiple of 4 (X and the weight rows are zero-padded).
__device__ __forceinline__ void fast_layer(const float* __restrict__ X, int K4, const float* __restrict__ Wq,
                                           const float* __restrict__ b, const float* __restrict__ gm,
                                           const float* __restrict__ be, int act, float* __restrict__ Y, int lane) {
  float a0 = 0.f, a1 = 0.f;
  const float2* w2 = reinterpret_cast<const float2*>(Wq) + lane;
  const float4* x4 = reinterpret_cast<const float4*>(X);
#pragma unroll 4
  for (int q = 0; q < (K4 >> 2); ++q) {
    const float4 x = x4[q];
    const float2 wa = w2[(4 * q + 0) * 32], wb = w2[(4 * q + 1) * 32], wc = w2[(4 * q + 2) * 32], wd = w2[(4 * q + 3) * 32];
    a0 = fmaf(x.x, wa.x, a0); a1 = fmaf(x.x, wa.y, a1);
    a0 = fmaf(x.y, wb.x, a0); a1 = fmaf(x.y, wb.y, a1);
    a0 = fmaf(x.z, wc.x, a0); a1 = fmaf(x.z, wc.y, a1);
    a0 = fmaf(x.w, wd.x, a0); a1 = fmaf(x.w, wd.y, a1);
  }
  a0 = act_fwd(a0 + b[lane], act);
  a1 = act_fwd(a1 + b[lane + 32], act);
  const float m = warp_sum(a0 + a1) * (1.0f / 64.f);
  const float d0 = a0 - m, d1 = a1 - m;
  const float rs = 1.0f / sqrtf(warp_sum(fmaf(d1, d1, d0 * d0)) * (1.0f / 64.f) + kLnEps);
  Y[lane] = fmaf(d0 * rs, gm[lane], be[lane]);
  Y[lane + 32] = fmaf(d1 * rs, gm[lane + 32], be[lane + 32]);
  __syncwarp();
}

struct FastCtx {
  const float* sW;      // weight image in shared memory
  FastImg f;
  float* bufA;          // this warp's scratch: 64 floats
  float* bufB;          // 64 floats
};

// One rollout step of net `which` for the row of this warp (storage row g, or -1 past the end).  xin: the lane's input
// features k = lane and lane + 32.  PolStep as in the tile path (policy_step.cu); recurrent fields are unused here.
__device__ __forceinline__ void fast_step(const NetDev& n, int which, const FastCtx& c, const PolStep& p,
                                          const float (&xin)[2], int g, int lane, int n_avail, int deterministic,
                                          uint64_t rng_seed, long long& t_last, int tid) {
  const int in = n.in_dim;
  const FastImg& f = c.f;
  const float* sW = c.sW;
  // ---- the insert of this slot: rows, availability, masks ----
  if (g >= 0) {
    if (p.in_copy) {
      if (lane < in) p.in_copy[(size_t)g * in + lane] = xin[0];
      if (lane + 32 < in) p.in_copy[(size_t)g * in + lane + 32] = xin[1];
    }
    if (which == 0 && p.avail_copy && p.avail)
      for (int k = lane; k < n_avail; k += 32) p.avail_copy[(size_t)g * n_avail + k] = p.avail[(size_t)g * n_avail + k];
    if (which == 0 && p.masks_copy && lane == 0)
      p.masks_copy[g] = p.done_prev ? (p.done_prev[g] != 0.f ? 0.f : 1.f) : p.masks[g];
  }
  if (!p.forward) return;
  POL_T(0);
  // ---- feature LayerNorm (mlp.py:47-56) straight from registers ----
  {
    const bool v0 = lane < in, v1 = lane + 32 < in;
    float y0 = xin[0], y1 = xin[1];
    if (n.use_fn) {
      const float m = warp_sum((v0 ? xin[0] : 0.f) + (v1 ? xin[1] : 0.f)) / (float)in;
      const float d0 = v0 ? xin[0] - m : 0.f, d1 = v1 ? xin[1] - m : 0.f;
      const float rs = 1.0f / sqrtf(warp_sum(fmaf(d1, d1, d0 * d0)) / (float)in + kLnEps);
      y0 = v0 ? fmaf(d0 * rs, sW[f.fn_w + lane], sW[f.fn_b + lane]) : 0.f;
      y1 = v1 ? fmaf(d1 * rs, sW[f.fn_w + lane + 32], sW[f.fn_b + lane + 32]) : 0.f;
    }
    if (lane < f.K1) c.bufA[lane] = v0 ? y0 : 0.f;
    if (lane + 32 < f.K1) c.bufA[lane + 32] = v1 ? y1 : 0.f;
    __syncwarp();
  }
  const int act = n.use_relu ? ACT_RELU : ACT_TANH;
  fast_layer(c.bufA, f.K1, sW + f.w1, sW + f.b1, sW + f.g1, sW + f.be1, act, c.bufB, lane);
  float* X = c.bufB;
  float* Y = c.bufA;
  for (int l = 0; l < n.layer_n; ++l) {
    fast_layer(X, 64, sW + f.w2[l], sW + f.b2[l], sW + f.g2[l], sW + f.be2[l], act, Y, lane);
    float* t = X; X = Y; Y = t;
  }
  POL_T(1);
  // ---- heads: lane a owns output a ----
  const int Atot = n.head_total;
  float lg = 0.f;
  if (lane < Atot) {
    float acc = 0.f;
    const float* w = sW + f.wh + lane;
    const float4* x4 = reinterpret_cast<const float4*>(X);
    const int AP = f.AP;
#pragma unroll 4
    for (int q = 0; q < 16; ++q) {
      const float4 x = x4[q];
      acc = fmaf(x.x, w[(4 * q + 0) * AP], acc);
      acc = fmaf(x.y, w[(4 * q + 1) * AP], acc);
      acc = fmaf(x.z, w[(4 * q + 2) * AP], acc);
      acc = fmaf(x.w, w[(4 * q + 3) * AP], acc);
    }
    lg = acc + sW[f.bh + lane];
  }
  __syncwarp();
  POL_T(3);
  if (which == 1) {
    if (lane == 0 && g >= 0 && p.values) p.values[g] = lg;
    POL_T(4);
    return;
  }
  const float* av = (p.avail && n.n_heads == 1 && g >= 0) ? p.avail + (size_t)g * n_avail : nullptr;
  const uint64_t ctr = p.rng_ctr + (uint64_t)(g < 0 ? 0 : g);
  int off = 0;
  for (int k = 0; k < n.n_heads; ++k) {
    const int A = n.head_dim[k];
    const int j = lane - off;
    const bool valid = j >= 0 && j < A && g >= 0;
    float l = valid ? lg : -INFINITY;
    if (valid && av && av[j] == 0.f) l = -1e10f;                          // distributions.py:66-67
    float mx = l;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = valid ? expf(l - mx) : 0.f;
    float se = 0.f;
    for (int jj = 0; jj < A; ++jj) se += __shfl_sync(0xffffffffu, e, off + jj);   // serial order: head_lse's rounding
    const float lse = mx + logf(se);
    float bestv = -INFINITY, best_lp = 0.f;
    int best = 1 << 30;
    if (valid) {
      const float lp = l - lse;
      const float pr = expf(lp);
      float score = pr;
      if (!deterministic) {
        float q;
        if (p.exp_noise) {
          q = p.exp_noise[(size_t)g * Atot + off + j];
        } else {
          const uint4 rnd = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(k * 64 + (j >> 2)), 0u),
                                          make_uint2((uint32_t)rng_seed, (uint32_t)(rng_seed >> 32)));
          const uint32_t x = (j & 3) == 0 ? rnd.x : ((j & 3) == 1 ? rnd.y : ((j & 3) == 2 ? rnd.z : rnd.w));
          q = -logf(((float)x + 0.5f) * 2.3283064365386963e-10f);
        }
        score = pr / q;                                                   // torch multinomial: argmax(p / Exp(1))
      }
      bestv = score; best = j; best_lp = lp;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                                    // first maximum wins, like the serial scan
      const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best, o);
      const float olp = __shfl_xor_sync(0xffffffffu, best_lp, o);
      if (ov > bestv || (ov == bestv && oj < best)) { bestv = ov; best = oj; best_lp = olp; }
    }
    if (best == (1 << 30)) best = 0;
    if (lane == 0 && g >= 0) {
      const int as = n.n_heads;
      if (p.actions) p.actions[(size_t)g * as + k] = (float)best;
      if (p.actions_i64) p.actions_i64[(size_t)g * as + k] = (int64_t)best;
      if (p.logp) p.logp[(size_t)g * as + k] = best_lp;
    }
    off += A;
  }
  POL_T(4);
}

// carve the CTA's shared memory, fetch the image (one TMA bulk copy, or a plain pack when the caller has none)
__device__ __forceinline__ FastCtx fast_setup(const NetDev& n, float* smem, const float* image, uint64_t* wbar, int tid) {
  FastCtx c;
  c.f = make_fast_img(n);
  c.sW = smem;
  float* ws = smem + c.f.total + (tid >> 5) * kFWarpScratch;
  c.bufA = ws; c.bufB = ws + 64;
  if (tid == 0) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(wbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(c.f.total * 4)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(image), "r"((uint32_t)(c.f.total * 4)), "r"(bar) : "memory");
  }
  __syncthreads();
  {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(wbar);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
  }
  return c;
}

__device__ __forceinline__ void load_row_lane(const float* __restrict__ src, int g, int in, int lane, float (&x)[2]) {
  x[0] = (g >= 0 && lane < in) ? __ldg(src + (size_t)g * in + lane) : 0.f;
  x[1] = (g >= 0 && lane + 32 < in) ? __ldg(src + (size_t)g * in + lane + 32) : 0.f;
}

__global__ void __launch_bounds__(kFT)
policy_step_fast_kernel(const NetDev na, const NetDev nc, const PolArgs a, int first_net) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t wbar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int which = first_net + blockIdx.y;
  const NetDev& n = which == 0 ? na : nc;
  const int row = blockIdx.x * kFR + (tid >> 5);
  const int g = row < a.n_rows ? row : -1;
  float x[2];
  load_row_lane(a.in[which], g, n.in_dim, lane, x);                 // in flight while the weights arrive
  const FastCtx c = fast_setup(n, smem, a.image[which], &wbar, tid);
  PolStep p;
  p.in = nullptr; p.in_copy = nullptr; p.h_in = nullptr; p.masks = a.masks; p.done_prev = nullptr;
  p.masks_copy = nullptr; p.h_out = nullptr; p.done_now = nullptr; p.avail = a.avail; p.avail_copy = nullptr;
  p.exp_noise = a.exp_noise;
  p.rng_ctr = (!a.exp_noise && !a.deterministic && which == 0) ? *a.rng_offset : 0ull;
  p.values = a.values; p.actions = a.actions; p.actions_i64 = a.actions_i64; p.logp = a.logp; p.forward = true;
  long long t_last = clock64();
  fast_step(n, which, c, p, x, g, lane, a.n_avail, a.deterministic, a.rng_seed, t_last, tid);
}

// The T collect steps + inserts of one iteration for feed-forward policies (see rollout_persistent_kernel for the
// contract); the next step's rows are prefetched while the current one is computed.
__global__ void __launch_bounds__(kFT)
rollout_fast_kernel(const NetDev na, const NetDev nc, const RolloutArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t wbar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int which = blockIdx.y;
  const NetDev& n = which == 0 ? na : nc;
  const int E = a.E, T = a.T, in = n.in_dim;
  const int row = blockIdx.x * kFR + (tid >> 5);
  const int g = row < E ? row : -1;
  float* store_in = which == 0 ? a.obs : a.share_obs;
  const float* feed_in = which == 0 ? a.f_obs : a.f_share;
  float x[2];
  load_row_lane(store_in, g, in, lane, x);                          // slot 0
  const FastCtx c = fast_setup(n, smem, a.image[which], &wbar, tid);
  const int Atot = na.head_total;
  const uint64_t rng0 = (!a.exp_noise && which == 0) ? *a.rng_offset : 0ull;
  long long t_last = clock64();
#pragma unroll 1
  for (int t = 0; t <= T; ++t) {
    float xn[2] = {0.f, 0.f};
    if (t < T) {                                                  // rows of step t + 1
      if (which == 1 && a.share_agents > 0)       // centralized V without a staged copy: concat of the thread's agents
        load_row_lane(a.f_obs + (size_t)t * E * na.in_dim, g < 0 ? -1 : g / a.share_agents, in, lane, xn);
      else
        load_row_lane(feed_in + (size_t)t * E * in, g, in, lane, xn);
    }
    PolStep p;
    p.in = nullptr;
    p.in_copy = t == 0 ? nullptr : store_in + (size_t)t * E * in;
    p.h_in = nullptr; p.h_out = nullptr; p.done_now = nullptr;
    p.masks = a.masks;
    p.done_prev = nullptr;
    p.masks_copy = nullptr;                                       // written below, off the critical path
    p.avail = a.avail ? (t == 0 ? a.avail : a.f_avail + (size_t)(t - 1) * E * a.n_avail) : nullptr;
    p.avail_copy = (a.avail && t > 0) ? a.avail + (size_t)t * E * a.n_avail : nullptr;
    p.exp_noise = (a.exp_noise && t < T) ? a.exp_noise + (size_t)t * E * Atot : nullptr;
    p.rng_ctr = rng0 + (uint64_t)t * (uint64_t)E;
    p.values = a.value_preds + (size_t)t * E;
    p.actions = t < T ? a.actions + (size_t)t * E * na.n_heads : nullptr;
    p.actions_i64 = nullptr;
    p.logp = t < T ? a.logp + (size_t)t * E * na.n_heads : nullptr;
    p.forward = (t < T) || which == 1;                            // slot T: only the critic's bootstrap value
    // rewards / masks / active masks of env step t-1 -> slot t-1 / t (the rest of insert), by the actor's lane 0 of the
    // row: the loads are issued here, the dependent stores wait until the step has been computed
    const bool book = which == 0 && t > 0 && lane == 0 && g >= 0;
    float b_rew = 0.f, b_done = 0.f, b_act = 0.f;
    if (book) {
      b_rew = __ldg(a.f_rew + (size_t)(t - 1) * E + g);
      b_done = __ldg(a.f_done + (size_t)(t - 1) * E + g);
      if (a.f_active) b_act = __ldg(a.f_active + (size_t)(t - 1) * E + g);
    }
    fast_step(n, which, c, p, x, g, lane, a.n_avail, 0, a.rng_seed, t_last, tid);
    if (book) {
      a.rewards[(size_t)(t - 1) * E + g] = b_rew;
      a.masks[(size_t)t * E + g] = b_done != 0.f ? 0.f : 1.f;
      if (a.f_active) a.active[(size_t)t * E + g] = b_act;
    }
    x[0] = xn[0]; x[1] = xn[1];
  }
}

inline size_t fast_smem_bytes(const NetDev& n) {
  return (size_t)(make_fast_img(n).total + (kFT / 32) * kFWarpScratch) * sizeof(float);
}

}  // namespace mappo
