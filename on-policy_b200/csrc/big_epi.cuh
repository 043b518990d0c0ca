// big_epi.cuh -- row epilogues of the layer-by-layer GEMM pipeline for hidden >= 128 MLP nets.
//
// One thread owns one row of a 128-row tile and walks its accumulator 32 columns at a time.  The SAME functors run
// behind the tcgen05 kernels (accumulator chunks from TMEM, big_gemm.cu) and behind the exact-fp32 FFMA path
// (accumulator chunks from a global scratch matrix, big_ref.cu), so the algebra below is validated independently of
// the tensor-core main loop.  The algebra (restated and checked against autograd in tests/test_bignet_algebra.py):
//
//   LayerNorm outputs are never materialised.  A layer stores a_l = act(z_l) (tf32-rounded in tf32 mode) plus the row
//   scalars (mu_l, rs_l); the NEXT GEMM runs on a_l and its epilogue applies
//       z_{l+1} = rs_l (a_l W'^T - mu_l s) + b',     W' = W diag(gamma_l),  b' = b + W beta_l,  s = rowsum(W').
//   Backward stores P_l = dZ_l rs_{l-1} (row scaled).  With acc = P_{l+1} W'_{l+1} (= rs_l dxhat_l):
//       dA_l = acc - (m1 + xhat_l m2),  xhat_l = (a_l - mu_l) rs_l,
//   where the two LayerNorm-backward row means come for free from the epilogue that produced P_{l+1}:
//       m1 = sum_o P_{l+1}[o] s[o] / H,     m2 = sum_o P_{l+1}[o] (z_{l+1}[o] - b'[o]) / H        (z = act^-1(a)).
//   Weight gradients: G = P_l^T [a_{l-1} | mu_{l-1} | sigma_{l-1} | 0...]  (one MN-major GEMM over the rows), then
//       dW' = G[:, :H] - G[:, H],  db' = G[:, H+1]   (unfolded by big_grad_finish_kernel).
// Reference semantics: algorithms/utils/mlp.py:6-57, act.py:115-178, r_mappo.py:52-169.
#pragma once
#include "net_tiles.cuh"
#include "rng.cuh"

namespace mappo {
namespace big {

constexpr int kChunk = 32;            // accumulator columns per epilogue step (= one 128-byte swizzled store tile)
constexpr int kExt = 64;              // extra columns of a stored activation row: [H] = mu, [H+1] = sigma, rest 0 (64: the
                                      // last 256 + 64 column tile of the pair weight-gradient GEMM splits into whole 32-column groups)
constexpr int kTileRows = 128;
constexpr int kLgLd = kTileRows + 4;  // leading dimension of the transposed logits scratch lgT[j][kLgLd]

__device__ __forceinline__ float round_op(float x, bool round_tf32) {
  if (!round_tf32) return x;
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float act_apply(float z, int act) { return act == ACT_RELU ? fmaxf(z, 0.f) : tanhf(z); }
__device__ __forceinline__ float act_inverse(float a, int act) {          // pre-activation from the stored output
  if (act == ACT_RELU) return a;                                          // only used where act'(a) != 0, i.e. z = a > 0
  const float c = fminf(fmaxf(a, -0.99999994f), 0.99999994f);
  return 0.5f * (log1pf(c) - log1pf(-c));
}

// ------------------------------------------------------------------------------------------------------------
// forward hidden layer:  acc = A W'^T  ->  a = act(rs_in (acc - mu_in s) + b'),  row statistics of a
// ------------------------------------------------------------------------------------------------------------
struct EpiFwd {
  static constexpr bool kHasAin = false, kStoresOut = true, kNeedsScratch = false;
  struct Args {
    const float* colvec;        // [2][N] global: s[o], b'[o]
    const float2* stats_in;     // (mu, rs) of the input activation per row; NULL = explicit (already normalised) input
    float2* stats_out;          // (mu, rs) of the output activation per row
    float* out;                 // output activation matrix [rows][ld_out] (mu / sigma go to columns N, N + 1)
    int ld_out, N, n_rows, act, round_tf32;
  };
  struct Row { float rs_in, c_in, sum, sumsq; };
  struct Thread {};
  // partial row sums handed from one epilogue group to the other when the N tiles of a row are split between them
  __device__ static void get_part(const Row& w, float* p) { p[0] = w.sum; p[1] = w.sumsq; }
  __device__ static void add_part(Row& w, const float* p) { w.sum += p[0]; w.sumsq += p[1]; }
  __device__ static void init_thread(Thread&) {}
  __device__ static void finish_thread(const Args&, Thread&, double*, int, int) {}
  __device__ static void begin_row(const Args& a, Row& w, int grow) {
    w.rs_in = 1.f; w.c_in = 0.f;
    if (a.stats_in && grow < a.n_rows) { const float2 st = a.stats_in[grow]; w.rs_in = st.y; w.c_in = st.x * st.y; }
    w.sum = 0.f; w.sumsq = 0.f;
  }
  // The activation and the tf32 rounding are compile-time inside the element loop (a run-time `act` makes the compiler evaluate
  // tanhf for every element and select afterwards: 4x the instructions of the ReLU case).
  template <int ACT, bool RT>
  __device__ static __forceinline__ void chunk_t(Row& w, const float* acc, float* out, const float* s, const float* b) {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      const float t = fmaf(-w.c_in, s[j], b[j]);
      const float z = fmaf(w.rs_in, acc[j], t);
      const float v = round_op(ACT == ACT_RELU ? fmaxf(z, 0.f) : tanhf(z), RT);
      w.sum += v;
      w.sumsq = fmaf(v, v, w.sumsq);
      out[j] = v;
    }
  }
  __device__ static void chunk(const Args& a, Thread& th, Row& w, const float* acc, const float* /*ain*/, float* out, int col0,
                               const float* cv, float* /*scratch*/, int /*r*/, int /*grow*/) {
    const float* s = cv + col0;
    const float* b = cv + a.N + col0;
    if (a.act == ACT_RELU) { if (a.round_tf32) chunk_t<ACT_RELU, true>(w, acc, out, s, b); else chunk_t<ACT_RELU, false>(w, acc, out, s, b); }
    else { if (a.round_tf32) chunk_t<ACT_TANH, true>(w, acc, out, s, b); else chunk_t<ACT_TANH, false>(w, acc, out, s, b); }
  }
  __device__ static void end_row(const Args& a, Row& w, int grow) {
    if (grow >= a.n_rows) return;
    const float inv = 1.0f / (float)a.N;
    const bool rt = a.round_tf32 != 0;
    const float m = w.sum * inv;
    const float var = fmaxf(fmaf(-m, m, w.sumsq * inv), 0.f);             // E[a^2] - mean^2 (one pass over the row)
    const float mu = round_op(m, rt);
    const float sigma = round_op(sqrtf(var + kLnEps), rt);
    a.stats_out[grow] = make_float2(mu, 1.0f / sigma);
    float4* e = reinterpret_cast<float4*>(a.out + (size_t)grow * a.ld_out + a.N);      // the kExt extra columns of the row
    e[0] = make_float4(mu, sigma, 0.f, 0.f);
#pragma unroll
    for (int q = 1; q < kExt / 4; ++q) e[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
};

// ------------------------------------------------------------------------------------------------------------
// backward hidden layer:  acc = P_{l+1} W'_{l+1}  ->  P_l = (acc - (m1 + xhat m2)) act'(a) rs_{l-1};  next (m1, m2)
// ------------------------------------------------------------------------------------------------------------
struct EpiBwd {
  static constexpr bool kHasAin = true, kStoresOut = true, kNeedsScratch = false;
  struct Args {
    const float* colvec;        // [2][N] global: s[o], b'[o] of the layer that PRODUCED a (for the next m1, m2)
    const float2* stats;        // (mu, rs) of a per row
    const float2* mprime;       // (m1, m2) of a per row
    const float2* stats_prev;   // (mu, rs) of the layer below (its rs scales P); NULL = explicit input (rs = 1)
    float2* mprime_out;         // (m1, m2) for the layer below; NULL when that is the explicit input
    int N, n_rows, act, round_tf32;
  };
  struct Row { float mu, rs, m1, m2, rs_prev, S1, S2; };
  struct Thread {};
  __device__ static void get_part(const Row& w, float* p) { p[0] = w.S1; p[1] = w.S2; }
  __device__ static void add_part(Row& w, const float* p) { w.S1 += p[0]; w.S2 += p[1]; }
  __device__ static void init_thread(Thread&) {}
  __device__ static void finish_thread(const Args&, Thread&, double*, int, int) {}
  __device__ static void begin_row(const Args& a, Row& w, int grow) {
    w.mu = 0.f; w.rs = 0.f; w.m1 = 0.f; w.m2 = 0.f; w.rs_prev = 0.f;
    if (grow < a.n_rows) {
      const float2 st = a.stats[grow], m = a.mprime[grow];
      w.mu = st.x; w.rs = st.y; w.m1 = m.x; w.m2 = m.y;
      w.rs_prev = a.stats_prev ? a.stats_prev[grow].y : 1.f;
    }
    w.S1 = 0.f; w.S2 = 0.f;
  }
  template <int ACT, bool RT>
  __device__ static __forceinline__ void chunk_t(Row& w, const float* acc, const float* ain, float* out, const float* s, const float* b) {
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      const float av = ain[j];
      const float xh = (av - w.mu) * w.rs;
      const float dA = acc[j] - fmaf(xh, w.m2, w.m1);
      const float p = round_op(dA * act_bwd(av, ACT) * w.rs_prev, RT);
      w.S1 = fmaf(p, s[j], w.S1);
      w.S2 = fmaf(p, act_inverse(av, ACT) - b[j], w.S2);
      out[j] = p;
    }
  }
  __device__ static void chunk(const Args& a, Thread& th, Row& w, const float* acc, const float* ain, float* out, int col0,
                               const float* cv, float* /*scratch*/, int /*r*/, int /*grow*/) {
    const float* s = cv + col0;
    const float* b = cv + a.N + col0;
    if (a.act == ACT_RELU) { if (a.round_tf32) chunk_t<ACT_RELU, true>(w, acc, ain, out, s, b); else chunk_t<ACT_RELU, false>(w, acc, ain, out, s, b); }
    else { if (a.round_tf32) chunk_t<ACT_TANH, true>(w, acc, ain, out, s, b); else chunk_t<ACT_TANH, false>(w, acc, ain, out, s, b); }
  }
  __device__ static void end_row(const Args& a, Row& w, int grow) {
    if (grow >= a.n_rows || !a.mprime_out) return;
    const float inv = 1.0f / (float)a.N;
    a.mprime_out[grow] = make_float2(w.S1 * inv, w.S2 * inv);
  }
};

// ------------------------------------------------------------------------------------------------------------
// heads + losses:  acc = a_L Wh'^T (N = 32 padded)  ->  logits, PPO / value loss, P_h = dlogits rs_L, (m1, m2) of a_L
// ------------------------------------------------------------------------------------------------------------
struct EpiHead {
  static constexpr bool kHasAin = false, kStoresOut = true, kNeedsScratch = true;
  struct Args {
    const float* colvec;        // [2][32] global: s_h[j], b_h'[j]
    const float2* stats;        // (mu, rs) of a_L
    float2* mprime_out;         // (m1, m2) of a_L
    NetDev n;
    BatchDev b;
    LossDev L;
    const double* norm_stats;
    const double* adv_stats;
    const float* vn_state;
    double* loss_out;
    int H, n_rows, round_tf32;
  };
  struct Row { float rs, c, S1, S2; RowIn rin; int gr; };
  struct Thread { double acc[3]; };
  __device__ static void get_part(const Row&, float*) {}          // (N = 32: one tile per row)
  __device__ static void add_part(Row&, const float*) {}
  __device__ static void init_thread(Thread& t) { t.acc[0] = t.acc[1] = t.acc[2] = 0.0; }
  // every thread of the CTA calls this once at the end (non-epilogue threads carry zeros); sred: [2 * 32] doubles
  __device__ static void finish_thread(const Args& a, Thread& t, double* sred, int tid, int nthreads) {
    if (a.n.is_critic) {
      double one[1] = {t.acc[0]};
      block_accumulate<1>(one, a.loss_out + 0, sred, tid, nthreads);
    } else {
      double two[2] = {t.acc[0], t.acc[1]};
      block_accumulate<2>(two, a.loss_out + 1, sred, tid, nthreads);
      double rt[1] = {t.acc[2] / (a.norm_stats[3] * (double)a.b.act_shape)};
      block_accumulate<1>(rt, a.loss_out + 5, sred, tid, nthreads);
    }
  }
  __device__ static void begin_row(const Args& a, Row& w, int grow) {
    w.rs = 0.f; w.c = 0.f; w.S1 = 0.f; w.S2 = 0.f;
    w.gr = -1;
    if (grow < a.n_rows) {
      const float2 st = a.stats[grow];
      w.rs = st.y; w.c = st.x * st.y;
      w.gr = a.b.rows ? a.b.rows[grow] : grow;
    }
    w.rin = load_row_in(a.n, a.b, w.gr);
  }
  // acc: the first kChunk columns of the head accumulator (col0 == 0); scratch = lgT[32][kLgLd]
  __device__ static void chunk(const Args& a, Thread& th, Row& w, const float* acc, const float* /*ain*/, float* out, int /*col0*/,
                               const float* cv, float* lgT, int r, int grow) {
    const int Atot = a.n.head_total;
    const LossConsts lc = make_loss_consts(a.n, a.L, a.norm_stats, a.adv_stats, a.vn_state);
    float z[kChunk];
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      z[j] = fmaf(w.rs, acc[j], fmaf(-w.c, cv[j], cv[kChunk + j]));
      if (j < Atot) lgT[j * kLgLd + r] = z[j];
    }
    row_loss_pre<kLgLd>(a.n, a.b, a.L, lc, lgT, r, w.gr, grow, w.rin, th.acc);      // logits -> d loss / d logits
    const bool rt = a.round_tf32 != 0;
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      float p = 0.f;
      if (j < Atot && w.gr >= 0) {
        p = round_op(lgT[j * kLgLd + r] * w.rs, rt);
        w.S1 = fmaf(p, cv[j], w.S1);
        w.S2 = fmaf(p, z[j] - cv[kChunk + j], w.S2);
      }
      out[j] = p;
    }
  }
  __device__ static void end_row(const Args& a, Row& w, int grow) {
    if (grow >= a.n_rows) return;
    const float inv = 1.0f / (float)a.H;
    a.mprime_out[grow] = make_float2(w.S1 * inv, w.S2 * inv);
  }
};

// ------------------------------------------------------------------------------------------------------------
// rollout heads:  logits -> Categorical sample (argmax p / Exp(1)) + log-prob  |  value      (act.py:44-113)
// ------------------------------------------------------------------------------------------------------------
struct EpiSample {
  static constexpr bool kHasAin = false, kStoresOut = false, kNeedsScratch = true;
  struct Args {
    const float* colvec;        // [2][32]
    const float2* stats;
    NetDev n;
    const float* avail;         // [rows][n_avail] or NULL
    const float* exp_noise;     // [rows][sum A] or NULL (Philox)
    uint64_t rng_seed;
    const uint64_t* rng_offset; // device counter (Philox), NULL with exp_noise / deterministic / critic
    int deterministic, n_rows, n_avail;
    float* values;              // critic
    float* actions;             // actor [rows][n_heads] (float, as stored by the buffer)
    int64_t* actions_i64;
    float* logp;                // [rows][n_heads]
  };
  struct Row { float rs, c; };
  struct Thread {};
  __device__ static void get_part(const Row&, float*) {}
  __device__ static void add_part(Row&, const float*) {}
  __device__ static void init_thread(Thread&) {}
  __device__ static void finish_thread(const Args&, Thread&, double*, int, int) {}
  __device__ static void begin_row(const Args& a, Row& w, int grow) {
    w.rs = 0.f; w.c = 0.f;
    if (grow < a.n_rows) { const float2 st = a.stats[grow]; w.rs = st.y; w.c = st.x * st.y; }
  }
  __device__ static void chunk(const Args& a, Thread& th, Row& w, const float* acc, const float* /*ain*/, float* /*out*/, int /*col0*/,
                               const float* cv, float* lgT, int r, int grow) {
    if (grow >= a.n_rows) return;
    const int Atot = a.n.head_total;
#pragma unroll
    for (int j = 0; j < kChunk; ++j)
      if (j < Atot) lgT[j * kLgLd + r] = fmaf(w.rs, acc[j], fmaf(-w.c, cv[j], cv[kChunk + j]));
    if (a.n.is_critic) {
      if (a.values) a.values[grow] = lgT[r];
      return;
    }
    const float* av = (a.avail && a.n.n_heads == 1) ? a.avail + (size_t)grow * a.n_avail : nullptr;
    const int as = a.n.n_heads;
    const uint64_t ctr = (a.rng_offset ? *a.rng_offset : 0ull) + (uint64_t)grow;
    int off = 0;
    for (int k = 0; k < as; ++k) {
      const int A = a.n.head_dim[k];
      float lse;
      head_lse<kLgLd>(lgT, off, A, r, av, lse);
      int best = 0;
      float bestv = -INFINITY, best_lp = 0.f;
      uint4 rnd = make_uint4(0, 0, 0, 0);
      for (int j = 0; j < A; ++j) {
        float lgt = lgT[(off + j) * kLgLd + r];
        if (av && av[j] == 0.f) lgt = -1e10f;
        const float lp = lgt - lse;
        const float pr = expf(lp);
        float score = pr;
        if (!a.deterministic) {
          float q;
          if (a.exp_noise) {
            q = a.exp_noise[(size_t)grow * Atot + off + j];
          } else {
            if ((j & 3) == 0)
              rnd = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(k * 64 + (j >> 2)), 0u),
                                  make_uint2((uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32)));
            const uint32_t x = (j & 3) == 0 ? rnd.x : ((j & 3) == 1 ? rnd.y : ((j & 3) == 2 ? rnd.z : rnd.w));
            q = -logf(((float)x + 0.5f) * 2.3283064365386963e-10f);
          }
          score = pr / q;                                  // torch multinomial: argmax(p / Exp(1))
        }
        if (score > bestv) { bestv = score; best = j; best_lp = lp; }
      }
      if (a.actions) a.actions[(size_t)grow * as + k] = (float)best;
      if (a.actions_i64) a.actions_i64[(size_t)grow * as + k] = (int64_t)best;
      if (a.logp) a.logp[(size_t)grow * as + k] = best_lp;
      off += A;
    }
  }
  __device__ static void end_row(const Args&, Row&, int) {}
};

}  // namespace big
}  // namespace mappo
