// rollout_gru.cuh -- rollout inference of recurrent (GRU) hidden-64 policies: ONE WARP OWNS TWO ROWS END TO END.
// (included by policy_step.cu; the same design as rollout_mlp.cuh, extended by the GRU cell of algorithms/utils/rnn.py:24-79.)
//
// The 32-row tile path (pol_step) runs one warp per scheduler and spends 47 k cycles per step in the six 64 x 64 gate GEMMs
// (profiles/r2_summary.md 6.7).  Rows of a rollout step never interact, so here a warp keeps two rows: lane l owns hidden columns
// l and l + 32 of both; the recurrent state of a row lives in two registers of each lane across all T steps.
//   * every weight is read as one conflict-free LDS.64 per k ([k][lane][2] packing, like the feed-forward image) and feeds FOUR
//     FMAs (two columns x two rows): half the shared-memory traffic per FMA of the one-row kernel, 16 warps per CTA hide latency,
//   * activations travel through per-row scratch that only this warp touches (__syncwarp, no CTA barrier inside a step),
//   * every dot product accumulates k = 0, 1, 2, ... sequentially in one register and the gate pre-activations are added in the
//     order of the tile path -- ((h part + (x part + b_ih)) + b_hh) -- so the numbers a step produces are those of pol_step up
//     to the LayerNorm statistics (warp-shuffle sums instead of the tile path's 16-lane sums): fp32, no tensor cores (the sampled
//     integer actions must match the reference bit for bit).
#pragma once
#include "rollout_mlp.cuh"

namespace mappo {

constexpr int kGW = 16;                  // warps per CTA at most (the launch picks 2..16 so that the CTAs of both nets cover the SMs)
constexpr int kGT = 32 * kGW;            // threads per CTA at most
constexpr int kGWarpScratch = 2 * 192;   // floats per warp: per row two 64-float activation buffers + the masked state

struct GruFastImg {
  FastImg f;                             // base MLP + heads, as for feed-forward nets
  int wih, whh;                          // [gate][k][lane][2]
  int bih, bhh, rg, rb, total;           // biases [192], rnn.norm affine [64]
};
__host__ __device__ inline GruFastImg make_gru_fast_img(const NetDev& n) {
  GruFastImg m;
  m.f = make_fast_img(n);
  int o = m.f.total;
  m.wih = o; o += 3 * 64 * 64;
  m.whh = o; o += 3 * 64 * 64;
  m.bih = o; o += 192;
  m.bhh = o; o += 192;
  m.rg = o; o += 64;
  m.rb = o; o += 64;
  m.total = o;
  return m;
}
__host__ __device__ inline bool gru_fast_supported(const NetDev& n) {
  return n.recurrent && n.hid == 64 && n.in_dim <= 64 && n.head_total <= 32;
}

__global__ void __launch_bounds__(256) pack_gru_fast_kernel(const NetDev n, const float* __restrict__ p, float* __restrict__ img) {
  const GruFastImg m = make_gru_fast_img(n);
  const mappo_net_layout_t& g = n.g;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.total; i += gridDim.x * blockDim.x) {
    float v;
    if (i < m.f.total) v = pack_fast_element(n, m.f, p, i);
    else if (i < m.bih) {                                   // gate matrices as [gate][k][lane][j]: element = W[gate * 64 + lane + 32 j][k]
      const bool hh = i >= m.whh;
      const int t = i - (hh ? m.whh : m.wih), gate = t >> 12, u = t & 4095, k = u >> 6, ln = (u >> 1) & 31, j = u & 1;
      v = p[(hh ? g.gru_whh : g.gru_wih) + (gate * 64 + ln + 32 * j) * 64 + k];
    } else if (i < m.bhh) v = p[g.gru_bih + i - m.bih];
    else if (i < m.rg) v = p[g.gru_bhh + i - m.bhh];
    else if (i < m.rb) v = p[g.rnn_ln_w + i - m.rg];
    else v = p[g.rnn_ln_b + i - m.rb];
    img[i] = v;
  }
}

struct GruFastCtx {
  const float* sW;
  GruFastImg m;
  float* buf[2][3];        // per row of this warp: activation buffers A, B and the masked recurrent state
};

// Y = LayerNorm(act(X W^T + b)) * gamma + beta for the warp's two rows (the arithmetic of fast_layer, one weight load for both rows)
__device__ __forceinline__ void fast_layer2(const float* __restrict__ X0, const float* __restrict__ X1, int K4,
                                            const float* __restrict__ Wq, const float* __restrict__ b, const float* __restrict__ gm,
                                            const float* __restrict__ be, int act, float* __restrict__ Y0, float* __restrict__ Y1,
                                            int lane) {
  float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
  const float2* w2 = reinterpret_cast<const float2*>(Wq) + lane;
  const float4* x4 = reinterpret_cast<const float4*>(X0);
  const float4* y4 = reinterpret_cast<const float4*>(X1);
#pragma unroll 4
  for (int q = 0; q < (K4 >> 2); ++q) {
    const float4 x = x4[q], y = y4[q];
    const float2 wa = w2[(4 * q + 0) * 32], wb = w2[(4 * q + 1) * 32], wc = w2[(4 * q + 2) * 32], wd = w2[(4 * q + 3) * 32];
    a0 = fmaf(x.x, wa.x, a0); a1 = fmaf(x.x, wa.y, a1); c0 = fmaf(y.x, wa.x, c0); c1 = fmaf(y.x, wa.y, c1);
    a0 = fmaf(x.y, wb.x, a0); a1 = fmaf(x.y, wb.y, a1); c0 = fmaf(y.y, wb.x, c0); c1 = fmaf(y.y, wb.y, c1);
    a0 = fmaf(x.z, wc.x, a0); a1 = fmaf(x.z, wc.y, a1); c0 = fmaf(y.z, wc.x, c0); c1 = fmaf(y.z, wc.y, c1);
    a0 = fmaf(x.w, wd.x, a0); a1 = fmaf(x.w, wd.y, a1); c0 = fmaf(y.w, wd.x, c0); c1 = fmaf(y.w, wd.y, c1);
  }
  const float b0 = b[lane], b1 = b[lane + 32];
  a0 = act_fwd(a0 + b0, act); a1 = act_fwd(a1 + b1, act);
  c0 = act_fwd(c0 + b0, act); c1 = act_fwd(c1 + b1, act);
  {
    const float m = warp_sum(a0 + a1) * (1.0f / 64.f);
    const float d0 = a0 - m, d1 = a1 - m;
    const float rs = 1.0f / sqrtf(warp_sum(fmaf(d1, d1, d0 * d0)) * (1.0f / 64.f) + kLnEps);
    Y0[lane] = fmaf(d0 * rs, gm[lane], be[lane]);
    Y0[lane + 32] = fmaf(d1 * rs, gm[lane + 32], be[lane + 32]);
  }
  {
    const float m = warp_sum(c0 + c1) * (1.0f / 64.f);
    const float d0 = c0 - m, d1 = c1 - m;
    const float rs = 1.0f / sqrtf(warp_sum(fmaf(d1, d1, d0 * d0)) * (1.0f / 64.f) + kLnEps);
    Y1[lane] = fmaf(d0 * rs, gm[lane], be[lane]);
    Y1[lane + 32] = fmaf(d1 * rs, gm[lane + 32], be[lane + 32]);
  }
  __syncwarp();
}

// r and z gates of both rows in ONE k loop (the activations are read once for two gates, 16 independent accumulators):
// acc[gate][row][part (0 = x, 1 = h)][column half]; the gate matrices are 4096 floats apart in the image
__device__ __forceinline__ void gru_gates_rz(const float* __restrict__ F0, const float* __restrict__ F1, const float* __restrict__ H0,
                                             const float* __restrict__ H1, const float* __restrict__ Wi, const float* __restrict__ Wh,
                                             int lane, float (&acc)[2][2][2][2]) {
#pragma unroll
  for (int gt = 0; gt < 2; ++gt)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int s = 0; s < 2; ++s) acc[gt][r][s][0] = acc[gt][r][s][1] = 0.f;
  const float2* wi = reinterpret_cast<const float2*>(Wi) + lane;
  const float2* wh = reinterpret_cast<const float2*>(Wh) + lane;
  const float4* f0 = reinterpret_cast<const float4*>(F0);
  const float4* f1 = reinterpret_cast<const float4*>(F1);
  const float4* h0 = reinterpret_cast<const float4*>(H0);
  const float4* h1 = reinterpret_cast<const float4*>(H1);
#pragma unroll 1
  for (int q = 0; q < 16; ++q) {
    const float4 x0 = f0[q], x1 = f1[q], y0 = h0[q], y1 = h1[q];
    const float xs0[4] = {x0.x, x0.y, x0.z, x0.w}, xs1[4] = {x1.x, x1.y, x1.z, x1.w};
    const float ys0[4] = {y0.x, y0.y, y0.z, y0.w}, ys1[4] = {y1.x, y1.y, y1.z, y1.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
#pragma unroll
      for (int gt = 0; gt < 2; ++gt) {
        const float2 u = wi[gt * 2048 + (4 * q + t) * 32], v = wh[gt * 2048 + (4 * q + t) * 32];
        acc[gt][0][0][0] = fmaf(xs0[t], u.x, acc[gt][0][0][0]); acc[gt][0][0][1] = fmaf(xs0[t], u.y, acc[gt][0][0][1]);
        acc[gt][1][0][0] = fmaf(xs1[t], u.x, acc[gt][1][0][0]); acc[gt][1][0][1] = fmaf(xs1[t], u.y, acc[gt][1][0][1]);
        acc[gt][0][1][0] = fmaf(ys0[t], v.x, acc[gt][0][1][0]); acc[gt][0][1][1] = fmaf(ys0[t], v.y, acc[gt][0][1][1]);
        acc[gt][1][1][0] = fmaf(ys1[t], v.x, acc[gt][1][1][0]); acc[gt][1][1][1] = fmaf(ys1[t], v.y, acc[gt][1][1][1]);
      }
    }
  }
}

// x part and state part of one gate for both rows: acc[row][part (0 = x, 1 = h)][column half]
__device__ __forceinline__ void gru_gate2(const float* __restrict__ F0, const float* __restrict__ F1, const float* __restrict__ H0,
                                          const float* __restrict__ H1, const float* __restrict__ Wi, const float* __restrict__ Wh,
                                          int lane, float (&acc)[2][2][2]) {
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int s = 0; s < 2; ++s) acc[r][s][0] = acc[r][s][1] = 0.f;
  const float2* wi = reinterpret_cast<const float2*>(Wi) + lane;
  const float2* wh = reinterpret_cast<const float2*>(Wh) + lane;
  const float4* f0 = reinterpret_cast<const float4*>(F0);
  const float4* f1 = reinterpret_cast<const float4*>(F1);
  const float4* h0 = reinterpret_cast<const float4*>(H0);
  const float4* h1 = reinterpret_cast<const float4*>(H1);
#pragma unroll 2
  for (int q = 0; q < 16; ++q) {
    const float4 x0 = f0[q], x1 = f1[q], y0 = h0[q], y1 = h1[q];
    const float xs0[4] = {x0.x, x0.y, x0.z, x0.w}, xs1[4] = {x1.x, x1.y, x1.z, x1.w};
    const float ys0[4] = {y0.x, y0.y, y0.z, y0.w}, ys1[4] = {y1.x, y1.y, y1.z, y1.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 u = wi[(4 * q + t) * 32], v = wh[(4 * q + t) * 32];
      acc[0][0][0] = fmaf(xs0[t], u.x, acc[0][0][0]); acc[0][0][1] = fmaf(xs0[t], u.y, acc[0][0][1]);
      acc[1][0][0] = fmaf(xs1[t], u.x, acc[1][0][0]); acc[1][0][1] = fmaf(xs1[t], u.y, acc[1][0][1]);
      acc[0][1][0] = fmaf(ys0[t], v.x, acc[0][1][0]); acc[0][1][1] = fmaf(ys0[t], v.y, acc[0][1][1]);
      acc[1][1][0] = fmaf(ys1[t], v.x, acc[1][1][0]); acc[1][1][1] = fmaf(ys1[t], v.y, acc[1][1][1]);
    }
  }
}

// heads + sampling of one row (the tail of fast_step): X = the row's 64 features, g = storage row or -1
__device__ __forceinline__ void fast_heads_row(const NetDev& n, int which, const float* __restrict__ sW, const FastImg& f,
                                               const PolStep& p, const float* __restrict__ X, int g, int lane, int n_avail,
                                               int deterministic, uint64_t rng_seed) {
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
  if (which == 1) {
    if (lane == 0 && g >= 0 && p.values) p.values[g] = lg;
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
}

// One rollout step of net `which` for the two rows of this warp (storage rows g[0], g[1], or -1 past the end).
// xin[r]: the lane's input features k = lane and lane + 32 of row r; h[r]: the lane's two columns of the row's recurrent state
// (updated in place: the new state, zeroed where the environment reported done -- mpe_runner.py:128-131); mask[r]: the step's mask.
__device__ __forceinline__ void gru_fast_step(const NetDev& n, int which, const GruFastCtx& c, const PolStep& p,
                                              const float (&xin)[2][2], const int (&g)[2], float (&h)[2][2], const float (&mask)[2],
                                              int lane, int n_avail, int deterministic, uint64_t rng_seed) {
  const int in = n.in_dim;
  const FastImg& f = c.m.f;
  const float* sW = c.sW;
  // ---- the insert of this slot: rows, availability, masks ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (g[r] < 0) continue;
    if (p.in_copy) {
      if (lane < in) p.in_copy[(size_t)g[r] * in + lane] = xin[r][0];
      if (lane + 32 < in) p.in_copy[(size_t)g[r] * in + lane + 32] = xin[r][1];
    }
    if (which == 0 && p.avail_copy && p.avail)
      for (int k = lane; k < n_avail; k += 32) p.avail_copy[(size_t)g[r] * n_avail + k] = p.avail[(size_t)g[r] * n_avail + k];
    if (which == 0 && p.masks_copy && lane == 0) p.masks_copy[g[r]] = mask[r];
  }
  if (!p.forward) return;
  // ---- feature LayerNorm (mlp.py:47-56) straight from registers ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const bool v0 = lane < in, v1 = lane + 32 < in;
    float y0 = xin[r][0], y1 = xin[r][1];
    if (n.use_fn) {
      const float m = warp_sum((v0 ? xin[r][0] : 0.f) + (v1 ? xin[r][1] : 0.f)) / (float)in;
      const float d0 = v0 ? xin[r][0] - m : 0.f, d1 = v1 ? xin[r][1] - m : 0.f;
      const float rs = 1.0f / sqrtf(warp_sum(fmaf(d1, d1, d0 * d0)) / (float)in + kLnEps);
      y0 = v0 ? fmaf(d0 * rs, sW[f.fn_w + lane], sW[f.fn_b + lane]) : 0.f;
      y1 = v1 ? fmaf(d1 * rs, sW[f.fn_w + lane + 32], sW[f.fn_b + lane + 32]) : 0.f;
    }
    if (lane < f.K1) c.buf[r][0][lane] = v0 ? y0 : 0.f;
    if (lane + 32 < f.K1) c.buf[r][0][lane + 32] = v1 ? y1 : 0.f;
  }
  __syncwarp();
  const int act = n.use_relu ? ACT_RELU : ACT_TANH;
  fast_layer2(c.buf[0][0], c.buf[1][0], f.K1, sW + f.w1, sW + f.b1, sW + f.g1, sW + f.be1, act, c.buf[0][1], c.buf[1][1], lane);
  int xi = 1;                                                // index of the buffer that holds the current features
  for (int l = 0; l < n.layer_n; ++l) {
    fast_layer2(c.buf[0][xi], c.buf[1][xi], 64, sW + f.w2[l], sW + f.b2[l], sW + f.g2[l], sW + f.be2[l], act, c.buf[0][xi ^ 1],
                c.buf[1][xi ^ 1], lane);
    xi ^= 1;
  }
  // ---- h <- h * mask (rnn.py:27), one GRU step (torch gate order r, z, n), LayerNorm (rnn.py:79) ----
  float hm[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    hm[r][0] = h[r][0] * mask[r];
    hm[r][1] = h[r][1] * mask[r];
    c.buf[r][2][lane] = hm[r][0];
    c.buf[r][2][lane + 32] = hm[r][1];
  }
  __syncwarp();
  const float* F0 = c.buf[0][xi];
  const float* F1 = c.buf[1][xi];
  float rgate[2][2], zgate[2][2];
  {
    float acc2[2][2][2][2];
    gru_gates_rz(F0, F1, c.buf[0][2], c.buf[1][2], sW + c.m.wih, sW + c.m.whh, lane, acc2);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = lane + 32 * j;
        const float gir = acc2[0][r][0][j] + sW[c.m.bih + col];
        rgate[r][j] = sigmoidf_((acc2[0][r][1][j] + gir) + sW[c.m.bhh + col]);
        const float giz = acc2[1][r][0][j] + sW[c.m.bih + 64 + col];
        zgate[r][j] = sigmoidf_((acc2[1][r][1][j] + giz) + sW[c.m.bhh + 64 + col]);
      }
  }
  {
    float acc[2][2][2];
    gru_gate2(F0, F1, c.buf[0][2], c.buf[1][2], sW + c.m.wih + 8192, sW + c.m.whh + 8192, lane, acc);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = lane + 32 * j;
        const float gin = acc[r][0][j] + sW[c.m.bih + 128 + col];
        const float ghn = acc[r][1][j] + sW[c.m.bhh + 128 + col];
        const float rg = rgate[r][j], zg = zgate[r][j];
        const float ng = tanhf(gin + rg * ghn);
        h[r][j] = (1.f - zg) * ng + zg * hm[r][j];
      }
  }
  __syncwarp();                                              // every lane has read the feature / state buffers
  float* Y0 = c.buf[0][xi ^ 1];
  float* Y1 = c.buf[1][xi ^ 1];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    // LayerNorm of the new state BEFORE the done reset (the tile path normalises hn and resets only what it stores / carries)
    const float a0 = h[r][0], a1 = h[r][1];
    const float m = warp_sum(a0 + a1) * (1.0f / 64.f);
    const float d0 = a0 - m, d1 = a1 - m;
    const float rs = 1.0f / sqrtf(warp_sum(fmaf(d1, d1, d0 * d0)) * (1.0f / 64.f) + kLnEps);
    float* Y = r == 0 ? Y0 : Y1;
    Y[lane] = fmaf(d0 * rs, sW[c.m.rg + lane], sW[c.m.rb + lane]);
    Y[lane + 32] = fmaf(d1 * rs, sW[c.m.rg + lane + 32], sW[c.m.rb + lane + 32]);
    if (g[r] >= 0 && p.done_now && p.done_now[g[r]] != 0.f) { h[r][0] = 0.f; h[r][1] = 0.f; }   // env done: next episode starts from zeros
    if (g[r] >= 0 && p.h_out) {
      p.h_out[(size_t)g[r] * 64 + lane] = h[r][0];
      p.h_out[(size_t)g[r] * 64 + lane + 32] = h[r][1];
    }
  }
  __syncwarp();
  fast_heads_row(n, which, sW, f, p, Y0, g[0], lane, n_avail, deterministic, rng_seed);
  fast_heads_row(n, which, sW, f, p, Y1, g[1], lane, n_avail, deterministic, rng_seed);
}

// carve the CTA's shared memory, fetch the image by TMA bulk copies
__device__ __forceinline__ GruFastCtx gru_fast_setup(const NetDev& n, float* smem, const float* image, uint64_t* wbar, int tid) {
  GruFastCtx c;
  c.m = make_gru_fast_img(n);
  c.sW = smem;
  float* ws = smem + c.m.total + (tid >> 5) * kGWarpScratch;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int b = 0; b < 3; ++b) c.buf[r][b] = ws + (r * 3 + b) * 64;
  if (tid == 0) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(wbar);
    const uint32_t bytes = (uint32_t)(c.m.total * 4), half = (uint32_t)(c.m.whh * 4);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(image), "r"(half), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem + c.m.whh)), "l"(image + c.m.whh), "r"(bytes - half), "r"(bar) : "memory");
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

inline size_t gru_fast_smem_bytes(const NetDev& n, int warps) {
  return (size_t)(make_gru_fast_img(n).total + warps * kGWarpScratch) * sizeof(float);
}
// warps per CTA (2 rows each): as few as cover the rows with one CTA per SM for both nets -- rows never interact and a warp's step is one
// long dependent chain, so spreading the warps over the SMs beats stacking them on a scheduler
inline int gru_fast_warps(int n_rows, int n_nets, int sm_count) {
  const int ctas = sm_count / (n_nets > 0 ? n_nets : 1) > 0 ? sm_count / (n_nets > 0 ? n_nets : 1) : 1;
  int w = (n_rows + 2 * ctas - 1) / (2 * ctas);
  if (w < 2) w = 2;
  if (w > kGW) w = kGW;
  return w;
}

// one step (mappo_policy_step) of recurrent nets: state in from a.h_in, out to a.h_out (no done reset here: the caller's insert does it)
__global__ void __launch_bounds__(kGT)
policy_step_gru_fast_kernel(const NetDev na, const NetDev nc, const PolArgs a, int first_net) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t wbar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int which = first_net + blockIdx.y;
  const NetDev& n = which == 0 ? na : nc;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (tid >> 5)) * 2;
  int g[2];
  float x[2][2], h[2][2], mask[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    g[r] = row0 + r < a.n_rows ? row0 + r : -1;
    load_row_lane(a.in[which], g[r], n.in_dim, lane, x[r]);           // in flight while the weights arrive
    h[r][0] = g[r] >= 0 ? a.h_in[which][(size_t)g[r] * 64 + lane] : 0.f;
    h[r][1] = g[r] >= 0 ? a.h_in[which][(size_t)g[r] * 64 + lane + 32] : 0.f;
    mask[r] = g[r] >= 0 ? a.masks[g[r]] : 0.f;
  }
  const GruFastCtx c = gru_fast_setup(n, smem, a.image[which], &wbar, tid);
  PolStep p;
  p.in = nullptr; p.in_copy = nullptr; p.h_in = nullptr; p.masks = a.masks; p.done_prev = nullptr;
  p.masks_copy = nullptr; p.h_out = a.h_out[which]; p.done_now = nullptr; p.avail = a.avail; p.avail_copy = nullptr;
  p.exp_noise = a.exp_noise;
  p.rng_ctr = (!a.exp_noise && !a.deterministic && which == 0) ? *a.rng_offset : 0ull;
  p.values = a.values; p.actions = a.actions; p.actions_i64 = a.actions_i64; p.logp = a.logp; p.forward = true;
  gru_fast_step(n, which, c, p, x, g, h, mask, lane, a.n_avail, a.deterministic, a.rng_seed);
}

// The T collect steps + inserts of one iteration for recurrent policies (the contract of rollout_persistent_kernel): the state is
// carried in registers, the next step's rows, done flags and bookkeeping scalars are prefetched while the current step is computed.
__global__ void __launch_bounds__(kGT)
rollout_gru_fast_kernel(const NetDev na, const NetDev nc, const RolloutArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t wbar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int which = blockIdx.y;
  const NetDev& n = which == 0 ? na : nc;
  const int E = a.E, T = a.T, in = n.in_dim;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (tid >> 5)) * 2;
  float* store_in = which == 0 ? a.obs : a.share_obs;
  const float* feed_in = which == 0 ? a.f_obs : a.f_share;
  float* h_store = which == 0 ? a.h_actor : a.h_critic;
  int g[2];
  float x[2][2], h[2][2], mask[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    g[r] = row0 + r < E ? row0 + r : -1;
    load_row_lane(store_in, g[r], in, lane, x[r]);                    // slot 0
    h[r][0] = g[r] >= 0 ? h_store[(size_t)g[r] * 64 + lane] : 0.f;
    h[r][1] = g[r] >= 0 ? h_store[(size_t)g[r] * 64 + lane + 32] : 0.f;
    mask[r] = g[r] >= 0 ? a.masks[g[r]] : 0.f;                        // slot 0
  }
  const GruFastCtx c = gru_fast_setup(n, smem, a.image[which], &wbar, tid);
  const int Atot = na.head_total;
  const uint64_t rng0 = (!a.exp_noise && which == 0) ? *a.rng_offset : 0ull;
#pragma unroll 1
  for (int t = 0; t <= T; ++t) {
    float xn[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    float mask_n[2] = {0.f, 0.f};
    if (t < T) {                                                    // rows and masks of step t + 1 (mask = 1 - done of env step t)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (which == 1 && a.share_agents > 0)
          load_row_lane(a.f_obs + (size_t)t * E * na.in_dim, g[r] < 0 ? -1 : g[r] / a.share_agents, in, lane, xn[r]);
        else
          load_row_lane(feed_in + (size_t)t * E * in, g[r], in, lane, xn[r]);
        if (g[r] >= 0) mask_n[r] = __ldg(a.f_done + (size_t)t * E + g[r]) != 0.f ? 0.f : 1.f;
      }
    }
    PolStep p;
    p.in = nullptr;
    p.in_copy = t == 0 ? nullptr : store_in + (size_t)t * E * in;
    p.h_in = nullptr;
    p.masks = a.masks;
    p.done_prev = nullptr;
    p.masks_copy = t == 0 ? nullptr : a.masks + (size_t)t * E;
    p.h_out = t < T ? h_store + (size_t)(t + 1) * E * 64 : nullptr;
    p.done_now = t < T ? a.f_done + (size_t)t * E : nullptr;
    p.avail = a.avail ? (t == 0 ? a.avail : a.f_avail + (size_t)(t - 1) * E * a.n_avail) : nullptr;
    p.avail_copy = (a.avail && t > 0) ? a.avail + (size_t)t * E * a.n_avail : nullptr;
    p.exp_noise = (a.exp_noise && t < T) ? a.exp_noise + (size_t)t * E * Atot : nullptr;
    p.rng_ctr = rng0 + (uint64_t)t * (uint64_t)E;
    p.values = a.value_preds + (size_t)t * E;
    p.actions = t < T ? a.actions + (size_t)t * E * na.n_heads : nullptr;
    p.actions_i64 = nullptr;
    p.logp = t < T ? a.logp + (size_t)t * E * na.n_heads : nullptr;
    p.forward = (t < T) || which == 1;                              // slot T: only the critic's bootstrap value
    // rewards / active masks of env step t-1 -> slot t-1 / t (the rest of insert), by the actor's lane 0 of each row
    if (which == 0 && t > 0 && lane == 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
        if (g[r] >= 0) {
          a.rewards[(size_t)(t - 1) * E + g[r]] = __ldg(a.f_rew + (size_t)(t - 1) * E + g[r]);
          if (a.f_active) a.active[(size_t)t * E + g[r]] = __ldg(a.f_active + (size_t)(t - 1) * E + g[r]);
        }
    }
    gru_fast_step(n, which, c, p, x, g, h, mask, lane, a.n_avail, 0, a.rng_seed);
#pragma unroll
    for (int r = 0; r < 2; ++r) { x[r][0] = xn[r][0]; x[r][1] = xn[r][1]; mask[r] = mask_n[r]; }
  }
}

}  // namespace mappo
