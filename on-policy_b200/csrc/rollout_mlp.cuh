// rollout_mlp.cuh -- rollout inference of feed-forward (MLP) policies: ONE WARP OWNS TWO ROWS END TO END.
// (included by policy_step.cu only; the GRU policies keep the tile path there.)
//
// Rows of a rollout step never interact, and a 64-wide MLP row fits one half-warp: lane tx of the 16 lanes of a row
// owns hidden columns tx, tx+16, tx+32, tx+48.  Then
//   * LayerNorm statistics are 16-lane shuffles (the same xor 8,4,2,1 tree as tile_mm_ln: identical rounding),
//   * a layer's activations travel through a 2-row scratch that only this warp touches: __syncwarp(), never a CTA barrier,
//   * the weights are packed as [k][tx][4] so one LDS.128 feeds the 4 FMAs of a k (mappo_pack_rollout_weights builds this
//     image for non-recurrent nets; a CTA fetches it with one TMA bulk copy),
//   * sampling is parallel over the actions of a head (lane j owns action j) but adds the softmax denominator in the
//     serial j = 0..A-1 order of the tile path, so log-probs and the argmax(p / Exp(1)) draw stay bit-identical to it.
// A CTA = 8 warps = 16 rows.  In the persistent rollout each warp walks t = 0..T with the NEXT step's rows already in
// flight (prefetched into registers), so no global-memory latency sits between two steps.
#pragma once
#include "net_tiles.cuh"
#include "launch_args.h"

namespace mappo {

constexpr int kFR = 16;             // rows per CTA
constexpr int kFT = 256;            // threads per CTA
constexpr int kFWarpScratch = 320;  // floats per warp: two [64][2] activation buffers + [2][32] logits

struct FastImg {
  int fn_w, fn_b, w1, b1, g1, be1;
  int w2[kMaxLayers], b2[kMaxLayers], g2[kMaxLayers], be2[kMaxLayers];
  int wh, bh, AP, total;
};
__host__ __device__ inline FastImg make_fast_img(const NetDev& n) {
  FastImg f;
  int o = 0;
  const int inp = (n.in_dim + 3) & ~3;
  f.fn_w = o; o += n.use_fn ? inp : 0;
  f.fn_b = o; o += n.use_fn ? inp : 0;
  f.w1 = o; o += n.in_dim * 64;
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

__global__ void __launch_bounds__(256) pack_fast_kernel(const NetDev n, const float* __restrict__ p, float* __restrict__ img) {
  const FastImg f = make_fast_img(n);
  const mappo_net_layout_t& g = n.g;
  const int in = n.in_dim, inp = (in + 3) & ~3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f.total; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < f.w1) {                                       // feature-norm affine (only present when use_fn)
      const int t = i - f.fn_w;
      if (t < inp) { if (t < in) v = p[g.fn_w + t]; }
      else if (t - inp < in) v = p[g.fn_b + t - inp];
    } else if (i < f.b1) {                                // fc1 as [k][tx][j]: element = W1[tx + 16 j][k]
      const int t = i - f.w1, k = t >> 6, tx = (t >> 2) & 15, j = t & 3;
      v = p[g.fc1_w + (tx + 16 * j) * in + k];
    } else if (i < f.g1) v = p[g.fc1_b + i - f.b1];
    else if (i < f.be1) v = p[g.ln1_w + i - f.g1];
    else if (i < f.be1 + 64) v = p[g.ln1_b + i - f.be1];
    else if (i < f.wh) {
      for (int l = 0; l < n.layer_n; ++l) {
        if (i >= f.w2[l] && i < f.b2[l]) {
          const int t = i - f.w2[l], k = t >> 6, tx = (t >> 2) & 15, j = t & 3;
          v = p[g.fc2_w[l] + (tx + 16 * j) * 64 + k];
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
    img[i] = v;
  }
}

__device__ __forceinline__ float sum16(float s) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Y = LayerNorm(act(X W^T + b)) * gamma + beta for this lane's row: X [K][2] -> Y [64][2] in the warp's scratch.
// Same arithmetic, in the same order, as tile_mm_ln (common.cuh).
__device__ __forceinline__ void fast_layer(const float* __restrict__ X, int K, const float* __restrict__ Wq,
                                           const float* __restrict__ b, const float* __restrict__ gm,
                                           const float* __restrict__ be, int act, float* __restrict__ Y, int tx, int rr) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* w4 = reinterpret_cast<const float4*>(Wq) + tx;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const float a = X[k * 2 + rr];
    const float4 w = w4[k * 16];
    acc[0] = fmaf(a, w.x, acc[0]);
    acc[1] = fmaf(a, w.y, acc[1]);
    acc[2] = fmaf(a, w.z, acc[2]);
    acc[3] = fmaf(a, w.w, acc[3]);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { acc[j] = act_fwd(acc[j] + b[tx + 16 * j], act); s += acc[j]; }
  const float m = sum16(s) * (1.0f / 64.f);
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float d = acc[j] - m; v = fmaf(d, d, v); }
  const float rs = 1.0f / sqrtf(sum16(v) * (1.0f / 64.f) + kLnEps);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tx + 16 * j;
    Y[c * 2 + rr] = fmaf((acc[j] - m) * rs, gm[c], be[c]);
  }
  __syncwarp();
}

struct FastCtx {
  const float* sW;      // weight image in shared memory
  FastImg f;
  float* bufA;          // this warp's scratch: [64][2]
  float* bufB;          // [64][2]
  float* lgs;           // [2][32] logits of the two rows
};

// One rollout step of net `which` for the row of this half-warp (storage row g, or -1 past the end).  xin: the lane's
// input features k = tx + 16 i.  PolStep as in the tile path (policy_step.cu); recurrent fields are unused here.
__device__ __forceinline__ void fast_step(const NetDev& n, int which, const FastCtx& c, const PolStep& p,
                                          const float (&xin)[4], int g, int tx, int rr, int lane, int n_avail,
                                          int deterministic, uint64_t rng_seed, long long& t_last, int tid) {
  const int in = n.in_dim;
  const FastImg& f = c.f;
  const float* sW = c.sW;
  // ---- the insert of this slot: rows, availability, masks ----
  if (g >= 0) {
    if (p.in_copy) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int k = tx + 16 * i; if (k < in) p.in_copy[(size_t)g * in + k] = xin[i]; }
    }
    if (which == 0 && p.avail_copy && p.avail)
      for (int k = tx; k < n_avail; k += 16) p.avail_copy[(size_t)g * n_avail + k] = p.avail[(size_t)g * n_avail + k];
    if (which == 0 && p.masks_copy && tx == 0)
      p.masks_copy[g] = p.done_prev ? (p.done_prev[g] != 0.f ? 0.f : 1.f) : p.masks[g];
  }
  if (!p.forward) return;
  POL_T(0);
  // ---- feature LayerNorm (mlp.py:47-56) straight from registers ----
  {
    float y[4];
    if (n.use_fn) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) s += (tx + 16 * i < in) ? xin[i] : 0.f;
      const float m = sum16(s) / (float)in;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float d = xin[i] - m; if (tx + 16 * i < in) v = fmaf(d, d, v); }
      const float rs = 1.0f / sqrtf(sum16(v) / (float)in + kLnEps);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = tx + 16 * i;
        y[i] = k < in ? fmaf((xin[i] - m) * rs, sW[f.fn_w + k], sW[f.fn_b + k]) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = xin[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int k = tx + 16 * i; if (k < in) c.bufA[k * 2 + rr] = y[i]; }
    __syncwarp();
  }
  const int act = n.use_relu ? ACT_RELU : ACT_TANH;
  fast_layer(c.bufA, in, sW + f.w1, sW + f.b1, sW + f.g1, sW + f.be1, act, c.bufB, tx, rr);
  float* X = c.bufB;
  float* Y = c.bufA;
  for (int l = 0; l < n.layer_n; ++l) {
    fast_layer(X, 64, sW + f.w2[l], sW + f.b2[l], sW + f.g2[l], sW + f.be2[l], act, Y, tx, rr);
    float* t = X; X = Y; Y = t;
  }
  POL_T(1);
  // ---- heads: lane tx owns outputs tx and tx + 16 (k-sequential accumulation like tile_mm) ----
  const int Atot = n.head_total;
  float lg[2] = {0.f, 0.f};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int a = tx + 16 * s;
    if (a < Atot) {
      float acc = 0.f;
      const float* w = sW + f.wh + a;
#pragma unroll 8
      for (int k = 0; k < 64; ++k) acc = fmaf(X[k * 2 + rr], w[k * f.AP], acc);
      lg[s] = acc + sW[f.bh + a];
    }
  }
  POL_T(3);
  if (which == 1) {
    if (tx == 0 && g >= 0 && p.values) p.values[g] = lg[0];
    POL_T(4);
    return;
  }
  c.lgs[rr * 32 + tx] = lg[0];
  c.lgs[rr * 32 + tx + 16] = lg[1];
  __syncwarp();
  const float* av = (p.avail && n.n_heads == 1 && g >= 0) ? p.avail + (size_t)g * n_avail : nullptr;
  const uint64_t ctr = p.rng_ctr + (uint64_t)(g < 0 ? 0 : g);
  const int half = lane & 16;
  int off = 0;
  for (int k = 0; k < n.n_heads; ++k) {
    const int A = n.head_dim[k];
    float l[2], e[2];
    bool valid[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = tx + 16 * s;
      valid[s] = j < A && g >= 0;
      l[s] = valid[s] ? c.lgs[rr * 32 + off + j] : -INFINITY;
      if (valid[s] && av && av[j] == 0.f) l[s] = -1e10f;                  // distributions.py:66-67
    }
    float mx = fmaxf(l[0], l[1]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    e[0] = valid[0] ? expf(l[0] - mx) : 0.f;
    e[1] = valid[1] ? expf(l[1] - mx) : 0.f;
    float se = 0.f;
    for (int j = 0; j < A; ++j) {                                         // serial order: same rounding as head_lse
      const float e0 = __shfl_sync(0xffffffffu, e[0], half | (j & 15));
      const float e1 = __shfl_sync(0xffffffffu, e[1], half | (j & 15));
      se += (j < 16) ? e0 : e1;
    }
    const float lse = mx + logf(se);
    float bestv = -INFINITY, best_lp = 0.f;
    int best = 1 << 30;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (valid[s]) {
        const int j = tx + 16 * s;
        const float lp = l[s] - lse;
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
          score = pr / q;                                                 // torch multinomial: argmax(p / Exp(1))
        }
        if (score > bestv) { bestv = score; best = j; best_lp = lp; }
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {                                     // first maximum wins, like the serial scan
      const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best, o);
      const float olp = __shfl_xor_sync(0xffffffffu, best_lp, o);
      if (ov > bestv || (ov == bestv && oj < best)) { bestv = ov; best = oj; best_lp = olp; }
    }
    if (best == (1 << 30)) best = 0;
    if (tx == 0 && g >= 0) {
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
  c.bufA = ws; c.bufB = ws + 128; c.lgs = ws + 256;
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

__device__ __forceinline__ void load_row_lane(const float* __restrict__ src, int g, int in, int tx, float (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = tx + 16 * i;
    x[i] = (g >= 0 && k < in) ? __ldg(src + (size_t)g * in + k) : 0.f;
  }
}

__global__ void __launch_bounds__(kFT)
policy_step_fast_kernel(const NetDev na, const NetDev nc, const PolArgs a, int first_net) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t wbar;
  const int tid = threadIdx.x, lane = tid & 31, tx = lane & 15, rr = lane >> 4;
  const int which = first_net + blockIdx.y;
  const NetDev& n = which == 0 ? na : nc;
  const int row = blockIdx.x * kFR + (tid >> 5) * 2 + rr;
  const int g = row < a.n_rows ? row : -1;
  float x[4];
  load_row_lane(a.in[which], g, n.in_dim, tx, x);                 // in flight while the weights arrive
  const FastCtx c = fast_setup(n, smem, a.image[which], &wbar, tid);
  PolStep p;
  p.in = nullptr; p.in_copy = nullptr; p.h_in = nullptr; p.masks = a.masks; p.done_prev = nullptr;
  p.masks_copy = nullptr; p.h_out = nullptr; p.done_now = nullptr; p.avail = a.avail; p.avail_copy = nullptr;
  p.exp_noise = a.exp_noise;
  p.rng_ctr = (!a.exp_noise && !a.deterministic && which == 0) ? *a.rng_offset : 0ull;
  p.values = a.values; p.actions = a.actions; p.actions_i64 = a.actions_i64; p.logp = a.logp; p.forward = true;
  long long t_last = clock64();
  fast_step(n, which, c, p, x, g, tx, rr, lane, a.n_avail, a.deterministic, a.rng_seed, t_last, tid);
}

// The T collect steps + inserts of one iteration for feed-forward policies (see rollout_persistent_kernel for the
// contract); the next step's rows are prefetched while the current one is computed.
__global__ void __launch_bounds__(kFT)
rollout_fast_kernel(const NetDev na, const NetDev nc, const RolloutArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t wbar;
  const int tid = threadIdx.x, lane = tid & 31, tx = lane & 15, rr = lane >> 4;
  const int which = blockIdx.y;
  const NetDev& n = which == 0 ? na : nc;
  const int E = a.E, T = a.T, in = n.in_dim;
  const int row = blockIdx.x * kFR + (tid >> 5) * 2 + rr;
  const int g = row < E ? row : -1;
  float* store_in = which == 0 ? a.obs : a.share_obs;
  const float* feed_in = which == 0 ? a.f_obs : a.f_share;
  float x[4];
  load_row_lane(store_in, g, in, tx, x);                          // slot 0
  const FastCtx c = fast_setup(n, smem, a.image[which], &wbar, tid);
  const int Atot = na.head_total;
  const uint64_t rng0 = (!a.exp_noise && which == 0) ? *a.rng_offset : 0ull;
  long long t_last = clock64();
#pragma unroll 1
  for (int t = 0; t <= T; ++t) {
    float xn[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < T) load_row_lane(feed_in + (size_t)t * E * in, g, in, tx, xn);       // rows of step t + 1
    PolStep p;
    p.in = nullptr;
    p.in_copy = t == 0 ? nullptr : store_in + (size_t)t * E * in;
    p.h_in = nullptr; p.h_out = nullptr; p.done_now = nullptr;
    p.masks = a.masks;
    p.done_prev = t == 0 ? nullptr : a.f_done + (size_t)(t - 1) * E;
    p.masks_copy = t == 0 ? nullptr : a.masks + (size_t)t * E;
    p.avail = a.avail ? (t == 0 ? a.avail : a.f_avail + (size_t)(t - 1) * E * a.n_avail) : nullptr;
    p.avail_copy = (a.avail && t > 0) ? a.avail + (size_t)t * E * a.n_avail : nullptr;
    p.exp_noise = (a.exp_noise && t < T) ? a.exp_noise + (size_t)t * E * Atot : nullptr;
    p.rng_ctr = rng0 + (uint64_t)t * (uint64_t)E;
    p.values = a.value_preds + (size_t)t * E;
    p.actions = t < T ? a.actions + (size_t)t * E * na.n_heads : nullptr;
    p.actions_i64 = nullptr;
    p.logp = t < T ? a.logp + (size_t)t * E * na.n_heads : nullptr;
    p.forward = (t < T) || which == 1;                            // slot T: only the critic's bootstrap value
    // rewards / active masks of env step t-1 -> slot t-1 / t (the rest of insert), by the actor's lane 0 of the row
    if (which == 0 && t > 0 && tx == 0 && g >= 0) {
      a.rewards[(size_t)(t - 1) * E + g] = a.f_rew[(size_t)(t - 1) * E + g];
      if (a.f_active) a.active[(size_t)t * E + g] = a.f_active[(size_t)(t - 1) * E + g];
    }
    fast_step(n, which, c, p, x, g, tx, rr, lane, a.n_avail, 0, a.rng_seed, t_last, tid);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = xn[i];
  }
}

inline size_t fast_smem_bytes(const NetDev& n) {
  return (size_t)(make_fast_img(n).total + (kFT / 32) * kFWarpScratch) * sizeof(float);
}

}  // namespace mappo
