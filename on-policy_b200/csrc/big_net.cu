// big_net.cu -- MLP actor / critic nets with hidden_size >= 128 (BASELINE config c5: hidden 512, layer_N 2): the
// layer-by-layer GEMM pipeline behind mappo_update_fwd_bwd / mappo_policy_step / mappo_evaluate_actions.
// Reference: algorithms/utils/mlp.py:6-57, act.py:44-178, r_actor_critic.py:44-117,156-175, r_mappo.py:91-169.
//
// Weights no longer fit next to the activations in shared memory, so every Linear is its own GEMM launch with the
// LayerNorm / activation / loss work fused into the epilogues (big_epi.cuh).  Per net and optimiser step:
//     pack  ->  [feature norm]  ->  (layer_N + 1) x fwd  ->  head + losses
//           ->  head grad GEMM  ->  (layer_N + 1) x (input-grad GEMM, weight-grad GEMM)  ->  reduce + unfold
// MAPPO_GEMM_TF32: tcgen05 kernels (big_gemm.cu); MAPPO_GEMM_FP32: FFMA kernels behind the same epilogues (big_ref.cu).
//
// Workspace (one flat fp32 buffer, offsets from make_plan):
//   packed weights  W'_i [H][Kp_i] (K-major, tf32-rounded in tf32 mode), W'_i^T [H][H] for i >= 1, (s_i, b'_i) [2][H],
//                   heads Wh' [32][H], Wh'^T [H][32], (s_h, b_h') [2][32]
//   x0   [rows][K0p]   normalised (pre-affine) input + constant-1 column at in_dim, zero padded to a multiple of 32
//   a_l  [rows][H+32]  l = 1..layer_N+1: act(z_l); columns H, H+1 = mu_l, sigma_l (what the weight-gradient GEMM needs)
//   stats_l, mprime_l  float2 per row;   P ping-pong [rows][H];   Ph [rows][32]
//   partial [splits][M][ldq], gsum [M][ldq];  fp32 mode: one accumulator scratch [rows][H]
#include <cstdlib>
#include <cstring>
#include <vector>
#include "big_net.h"

namespace mappo {

int grad_reduce_launch(const float*, int, int, float*, float*, int*, cudaStream_t);

namespace big {

constexpr int kMaxMat = kMaxLayers + 1;      // hidden matrices: fc1, fc2[0..layer_N)
constexpr int kUnfY = 16;                    // output-row groups of the unfold grid

struct Plan {
  int H, Hx, Lh, in_dim, K0p, rows, n_heads_pad;
  size_t wf[kMaxMat], wft[kMaxMat], cv[kMaxMat], whf, whft, cvh, pack_total;
  size_t x0, act[kMaxMat + 1], stats[kMaxMat + 1], mprime[kMaxMat + 1], P[2], Ph, partial, gsum, scratch, total;
  size_t unf_part, unf_ticket;           // big_unfold_kernel: [2][kUnfY][1024] column partials, 64 int tickets (zeroed by the pack kernel)
  size_t partial_floats;
};

static inline size_t align64(size_t x) { return (x + 63) & ~(size_t)63; }

// diagnostic (bench.py's roofline leg, eager passes only -- events are not recorded while a stream is capturing):
// CUDA-event timing of every launch family of the pipeline
enum { T_PACK = 0, T_FEATNORM, T_FWD, T_HEAD, T_BWD, T_GRAD, T_FINISH, T_N };
struct TimedLaunch { cudaEvent_t a, b; int cat; };
static bool g_timing = false;
static std::vector<TimedLaunch> g_timed;
struct Timed {
  cudaStream_t st; cudaEvent_t a; int cat; bool on;
  Timed(int c, cudaStream_t s) : st(s), a(nullptr), cat(c), on(false) {
    if (!g_timing) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    on = cudaEventCreate(&a) == cudaSuccess && cudaEventRecord(a, s) == cudaSuccess;
  }
  ~Timed() {
    if (!on) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) == cudaSuccess && cudaEventRecord(b, st) == cudaSuccess) g_timed.push_back({a, b, cat});
  }
};
int debug_timing(int enable, double* ms_out, long long* n_out) {
  for (int i = 0; i < T_N; ++i) { if (ms_out) ms_out[i] = 0.0; if (n_out) n_out[i] = 0; }
  for (const TimedLaunch& t : g_timed) {
    cudaEventSynchronize(t.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.a, t.b);
    if (ms_out) ms_out[t.cat] += ms;
    if (n_out) n_out[t.cat] += 1;
    cudaEventDestroy(t.a); cudaEventDestroy(t.b);
  }
  g_timed.clear();
  g_timing = enable != 0;
  return MAPPO_OK;
}

bool supported(const NetDev& n) {
  return !n.recurrent && n.hid >= 128 && n.hid <= 1024 && n.hid % 128 == 0 && n.head_total <= 32 && n.in_dim <= 1023;
}

static int grad_splits(int tiles, int rows, int sm) {
  int s = sm / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  const int cap = (rows + 255) / 256;
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}

// tiling of one weight-gradient GEMM G[M][Qw] = P^T Q: Q tiles of 256 columns, the last one takes what is left (<= 320), the rows of
// the batch split so that about one CTA (pair) per SM (pair) is busy.  pair: CTA pairs own 256-column tiles of P (big_grad_pair.cu),
// possible when P spans whole 256-column tiles and every Q tile is a multiple of 64 columns; otherwise 128-column tiles, one CTA each.
static bool grad_pair_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MAPPO_B200_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
static GradShape make_grad_shape(int rows, int M, int Pw, int Qw, int sm, bool* pair_out = nullptr) {
  GradShape g;
  memset(&g, 0, sizeof(g));
  g.rows = rows; g.M = M; g.Pw = Pw; g.Qw = Qw; g.ldq = Qw;
  int nt = 0, q = 0;
  while (q < Qw && nt < 4) {
    int w = Qw - q;
    if (w > 320) w = 256;
    g.q0[nt] = q; g.qw[nt] = w; q += w; ++nt;
  }
  g.n_tiles = nt;
  const bool pair = grad_pair_enabled() && Pw % 256 == 0 && Qw % 64 == 0 && sm >= 2;
  g.m_tiles = pair ? Pw / 256 : (Pw + 127) / 128;
  g.splits = grad_splits(g.m_tiles * g.n_tiles, rows, pair ? sm / 2 : sm);
  g.rows_per_split = (((rows + g.splits - 1) / g.splits) + 31) & ~31;
  if (pair_out) *pair_out = pair;
  return g;
}

static Plan make_plan(const NetDev& n, int rows, int sm) {
  Plan p;
  memset(&p, 0, sizeof(p));
  p.H = n.hid; p.Hx = n.hid + kExt; p.Lh = n.layer_n + 1; p.in_dim = n.in_dim; p.rows = rows;
  p.K0p = (n.in_dim + 1 + 63) & ~63;          // multiple of 64: pair gradient tiles split into whole 32-column groups
  size_t o = 0;
  for (int i = 0; i < p.Lh; ++i) {
    const int Kp = i == 0 ? p.K0p : p.H;
    p.wf[i] = o; o = align64(o + (size_t)p.H * Kp);
    p.wft[i] = o; if (i > 0) o = align64(o + (size_t)p.H * p.H);
    p.cv[i] = o; o = align64(o + 2 * (size_t)p.H);
  }
  p.whf = o; o = align64(o + 32 * (size_t)p.H);
  p.whft = o; o = align64(o + 32 * (size_t)p.H);
  p.cvh = o; o = align64(o + 64);
  p.unf_ticket = o; o = align64(o + 64);
  p.pack_total = o;
  const size_t R = (size_t)((rows + 127) / 128) * 128;
  p.x0 = o; o = align64(o + R * p.K0p);
  for (int l = 1; l <= p.Lh; ++l) {
    p.act[l] = o; o = align64(o + R * p.Hx);
    p.stats[l] = o; o = align64(o + 2 * R);
    p.mprime[l] = o; o = align64(o + 2 * R);
  }
  p.P[0] = o; o = align64(o + R * p.H);
  p.P[1] = o; o = align64(o + R * p.H);
  p.Ph = o; o = align64(o + R * 32);
  // gradient partials / slot sums: the largest of the three GEMM families (first matrix, hidden matrices, heads)
  const GradShape g0 = make_grad_shape(rows, p.H, p.H, p.K0p, sm), gh = make_grad_shape(rows, p.H, p.H, p.Hx, sm),
                  gq = make_grad_shape(rows, p.Hx, p.Hx, 32, sm);
  size_t pf = (size_t)g0.splits * g0.M * g0.ldq, gs = (size_t)g0.M * g0.ldq;
  if ((size_t)gh.splits * gh.M * gh.ldq > pf) pf = (size_t)gh.splits * gh.M * gh.ldq;
  if ((size_t)gq.splits * gq.M * gq.ldq > pf) pf = (size_t)gq.splits * gq.M * gq.ldq;
  if ((size_t)gh.M * gh.ldq > gs) gs = (size_t)gh.M * gh.ldq;
  if ((size_t)gq.M * gq.ldq > gs) gs = (size_t)gq.M * gq.ldq;
  p.partial_floats = pf;
  p.partial = o; o = align64(o + pf);
  p.gsum = o; o = align64(o + gs);
  p.unf_part = o; o = align64(o + 2 * kUnfY * 1024);
  p.scratch = o; o = align64(o + R * p.H);
  p.total = o;
  return p;
}

int64_t workspace_floats(const NetDev& n, int rows, int sm) { return (int64_t)make_plan(n, rows, sm).total; }

// diagnostic: float offsets of the workspace regions (scripts/diag_big.py compares the stored intermediates with float64 algebra)
//   out[0..7] = H, Hx, Lh, K0p, x0, P0, P1, Ph;  out[8] = partial, out[9] = gsum, out[10] = whf, out[11] = whft, out[12] = cvh;
//   out[16 + 4 l + {0,1,2}] = act[l], stats[l], mprime[l] (l = 1..Lh);  out[40 + 3 i + {0,1,2}] = wf[i], wft[i], cv[i]
int debug_plan(const NetDev& n, int rows, int sm, long long* out) {
  const Plan p = make_plan(n, rows, sm);
  for (int i = 0; i < 64; ++i) out[i] = -1;
  out[0] = p.H; out[1] = p.Hx; out[2] = p.Lh; out[3] = p.K0p; out[4] = (long long)p.x0; out[5] = (long long)p.P[0]; out[6] = (long long)p.P[1];
  out[7] = (long long)p.Ph; out[8] = (long long)p.partial; out[9] = (long long)p.gsum; out[10] = (long long)p.whf; out[11] = (long long)p.whft;
  out[12] = (long long)p.cvh;
  for (int l = 1; l <= p.Lh; ++l) { out[16 + 4 * l] = (long long)p.act[l]; out[17 + 4 * l] = (long long)p.stats[l]; out[18 + 4 * l] = (long long)p.mprime[l]; }
  for (int i = 0; i < p.Lh; ++i) { out[40 + 3 * i] = (long long)p.wf[i]; out[41 + 3 * i] = (long long)p.wft[i]; out[42 + 3 * i] = (long long)p.cv[i]; }
  return MAPPO_OK;
}

// ------------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------------
struct PackArgs {
  int H, Lh, in_dim, K0p, Atot, round_tf32;
  int w_off[kMaxMat], b_off[kMaxMat], gam_off[kMaxMat + 1], bet_off[kMaxMat + 1];   // flat-parameter offsets (-1 = identity LN)
  int hw_off, hb_off;
  long long wf[kMaxMat], wft[kMaxMat], cv[kMaxMat], whf, whft, cvh, ticket;
};

// one warp per output row of each folded matrix: W' = W diag(gamma_in), b' = b + W beta_in, s = rowsum(W')
__global__ void __launch_bounds__(256) big_pack_kernel(const PackArgs a, const float* __restrict__ p, float* __restrict__ ws) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_rows_total = a.Lh * a.H + 32;
  if (blockIdx.x == 0 && threadIdx.x < 64) reinterpret_cast<int*>(ws + a.ticket)[threadIdx.x] = 0;     // big_unfold_kernel's tickets
  if (gw >= n_rows_total) return;
  const bool rt = a.round_tf32 != 0;
  if (gw < a.Lh * a.H) {
    const int i = gw / a.H, o = gw % a.H;
    const int K = i == 0 ? a.in_dim : a.H, Kp = i == 0 ? a.K0p : a.H;
    const float* W = p + a.w_off[i] + (size_t)o * K;
    const float* gam = a.gam_off[i] >= 0 ? p + a.gam_off[i] : nullptr;
    const float* bet = a.bet_off[i] >= 0 ? p + a.bet_off[i] : nullptr;
    float* wf = ws + a.wf[i] + (size_t)o * Kp;
    float s = 0.f, bs = 0.f;
    for (int k = lane; k < Kp; k += 32) {
      float w = 0.f;
      if (k < K) {
        const float w0 = W[k];
        w = round_op(gam ? w0 * gam[k] : w0, rt);
        if (bet) bs = fmaf(w0, bet[k], bs);
      }
      wf[k] = w;
      if (i > 0 && k < K) ws[a.wft[i] + (size_t)k * a.H + o] = w;
      s += w;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); bs += __shfl_xor_sync(0xffffffffu, bs, d); }
    if (lane == 0) {
      ws[a.cv[i] + o] = i == 0 ? 0.f : s;                       // layer 0 reads an explicitly normalised input (no mean term)
      ws[a.cv[i] + a.H + o] = p[a.b_off[i] + o] + bs;
    }
  } else {
    const int j = gw - a.Lh * a.H;                               // head row (zero rows beyond the real outputs)
    const int L = a.Lh;
    const float* gam = a.gam_off[L] >= 0 ? p + a.gam_off[L] : nullptr;
    const float* bet = a.bet_off[L] >= 0 ? p + a.bet_off[L] : nullptr;
    float s = 0.f, bs = 0.f;
    for (int k = lane; k < a.H; k += 32) {
      float w = 0.f;
      if (j < a.Atot) {
        const float w0 = p[a.hw_off + (size_t)j * a.H + k];
        w = round_op(gam ? w0 * gam[k] : w0, rt);
        if (bet) bs = fmaf(w0, bet[k], bs);
      }
      ws[a.whf + (size_t)j * a.H + k] = w;
      ws[a.whft + (size_t)k * 32 + j] = w;
      s += w;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); bs += __shfl_xor_sync(0xffffffffu, bs, d); }
    if (lane == 0) {
      ws[a.cvh + j] = s;
      ws[a.cvh + 32 + j] = j < a.Atot ? p[a.hb_off + j] + bs : 0.f;
    }
  }
}

// x0[p] = [ LayerNorm(x[rows[p]]) (pre-affine) | 1 | 0... ]   one warp per row (mlp.py:52-54 feature_norm)
__global__ void __launch_bounds__(256) big_featnorm_kernel(const float* __restrict__ x, const int32_t* __restrict__ rows, int n_rows,
                                                           int in_dim, int K0p, int use_fn, int round_tf32, float* __restrict__ x0) {
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= n_rows) return;
  const int g = rows ? rows[p] : p;
  const float* src = x + (size_t)g * in_dim;
  float v[32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + 32 * i;
    v[i] = k < in_dim ? __ldg(src + k) : 0.f;
    s += v[i];
  }
  float mu = 0.f, rs = 1.f;
  if (use_fn) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    mu = s / (float)in_dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = lane + 32 * i;
      const float d = v[i] - mu;
      if (k < in_dim) q = fmaf(d, d, q);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
    rs = 1.0f / sqrtf(q / (float)in_dim + kLnEps);
  }
  float* dst = x0 + (size_t)p * K0p;
  const bool rt = round_tf32 != 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + 32 * i;
    if (k < K0p) dst[k] = k < in_dim ? round_op((v[i] - mu) * rs, rt) : (k == in_dim ? 1.f : 0.f);
  }
}

// hidden matrix i: the row-split partials of G[o][ldq] -> dW, db, d gamma_in, d beta_in   (chain rule of the folding, header of
// big_epi.cuh).   first (explicit input): dW' = G[:, :K], db' = G[:, K];  else dW' = G[:, :H] - G[:, H], db' = G[:, H + 1].
// Grid (K / 32, kUnfY): a CTA owns 32 input features x H / kUnfY output rows and sums the `splits` partials itself (in split
// order) -- no separate slot-sum launch, no summed copy of G.  The LayerNorm column sums cross the kUnfY row groups through
// `part`; the last CTA of a column block to arrive (ticket) adds them in row-group order, so the result is reproducible.
__global__ void __launch_bounds__(256) big_unfold_kernel(const float* __restrict__ partial, int splits, int ldq, const float* __restrict__ p,
                                                         float* __restrict__ g, int H, int K, int first, int w_off, int b_off, int gam_off,
                                                         int bet_off, float* part, int* tickets) {
  __shared__ float pg[8][33], pb[8][33];
  __shared__ int s_last;
  const int kx = threadIdx.x & 31, og = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx, y = blockIdx.y;
  const int rows_y = H / kUnfY, o0 = y * rows_y;
  const int ucol = first ? -1 : H, bcol = first ? K : H + 1;
  const size_t slot = (size_t)H * ldq;
  const float gam = (gam_off >= 0 && k < K) ? p[gam_off + k] : 1.f, bet = (bet_off >= 0 && k < K) ? p[bet_off + k] : 0.f;
  float sg = 0.f, sb = 0.f;
  for (int o = o0 + og; o < o0 + rows_y; o += 8) {
    const float* G = partial + (size_t)o * ldq;
    // four independent partial sums over the splits (12 loads in flight), combined in a fixed order
    float d4[4] = {0.f, 0.f, 0.f, 0.f}, u4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
    int s = 0;
    for (; s + 3 < splits; s += 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k < K) d4[q] += G[(s + q) * slot + k];
        if (ucol >= 0) u4[q] += G[(s + q) * slot + ucol];
        b4[q] += G[(s + q) * slot + bcol];
      }
    }
    for (; s < splits; ++s) {
      if (k < K) d4[0] += G[s * slot + k];
      if (ucol >= 0) u4[0] += G[s * slot + ucol];
      b4[0] += G[s * slot + bcol];
    }
    const float dbp = (b4[0] + b4[1]) + (b4[2] + b4[3]);
    const float dwf = ((d4[0] + d4[1]) + (d4[2] + d4[3])) - ((u4[0] + u4[1]) + (u4[2] + u4[3]));
    if (k < K) {
      const float w = p[w_off + (size_t)o * K + k];
      g[w_off + (size_t)o * K + k] = fmaf(dbp, bet, dwf * gam);
      sg = fmaf(dwf, w, sg);
      sb = fmaf(dbp, w, sb);
    }
    if (blockIdx.x == 0 && kx == 0) g[b_off + o] = dbp;
  }
  if (gam_off < 0) return;                                // no LayerNorm in front of this matrix (uniform over the grid)
  pg[og][kx] = sg; pb[og][kx] = sb;
  __syncthreads();
  if (og == 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { a += pg[q][kx]; c += pb[q][kx]; }
    part[(size_t)y * 1024 + k] = a;
    part[(size_t)(kUnfY + y) * 1024 + k] = c;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(tickets + blockIdx.x, 1) == kUnfY - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (og == 0 && k < K) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int q = 0; q < kUnfY; ++q) { a += __ldcg(part + (size_t)q * 1024 + k); c += __ldcg(part + (size_t)(kUnfY + q) * 1024 + k); }
    g[gam_off + k] = a;
    g[bet_off + k] = c;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0;          // ready for the next launch
}

// heads: gsum[k][32] (k over the extended activation row) -> dWh, dbh, d gamma_L, d beta_L
__global__ void __launch_bounds__(256) big_unfold_head_kernel(const float* __restrict__ gsum, const float* __restrict__ p, float* __restrict__ g,
                                                              int H, int Atot, int hw_off, int hb_off, int gam_off, int bet_off) {
  __shared__ float pg[8][33], pb[8][33];
  const int kx = threadIdx.x & 31, og = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + kx;
  const float gam = (gam_off >= 0 && k < H) ? p[gam_off + k] : 1.f, bet = (bet_off >= 0 && k < H) ? p[bet_off + k] : 0.f;
  float sg = 0.f, sb = 0.f;
  if (k < H) {
    for (int j = og; j < Atot; j += 8) {
      const float dbp = gsum[(size_t)(H + 1) * 32 + j];
      const float dwf = gsum[(size_t)k * 32 + j] - gsum[(size_t)H * 32 + j];
      const float w = p[hw_off + (size_t)j * H + k];
      g[hw_off + (size_t)j * H + k] = fmaf(dbp, bet, dwf * gam);
      sg = fmaf(dwf, w, sg);
      sb = fmaf(dbp, w, sb);
    }
  }
  pg[og][kx] = sg; pb[og][kx] = sb;
  __syncthreads();
  if (og == 0 && k < H && gam_off >= 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { a += pg[q][kx]; c += pb[q][kx]; }
    g[gam_off + k] = a;
    g[bet_off + k] = c;
  }
  if (blockIdx.x == 0 && threadIdx.x < Atot) g[hb_off + threadIdx.x] = gsum[(size_t)(H + 1) * 32 + threadIdx.x];
}

// ------------------------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------------------------
static PackArgs make_pack_args(const NetDev& n, const Plan& pl, bool round_tf32) {
  PackArgs a;
  memset(&a, 0, sizeof(a));
  a.H = pl.H; a.Lh = pl.Lh; a.in_dim = n.in_dim; a.K0p = pl.K0p; a.Atot = n.head_total; a.round_tf32 = round_tf32 ? 1 : 0;
  for (int i = 0; i < pl.Lh; ++i) {
    a.w_off[i] = i == 0 ? n.g.fc1_w : n.g.fc2_w[i - 1];
    a.b_off[i] = i == 0 ? n.g.fc1_b : n.g.fc2_b[i - 1];
    a.wf[i] = (long long)pl.wf[i]; a.wft[i] = (long long)pl.wft[i]; a.cv[i] = (long long)pl.cv[i];
  }
  // input LayerNorm of matrix i (i = Lh: the heads): feature_norm, ln1, ln2[0], ln2[1], ...
  for (int i = 0; i <= pl.Lh; ++i) {
    a.gam_off[i] = i == 0 ? (n.use_fn ? n.g.fn_w : -1) : (i == 1 ? n.g.ln1_w : n.g.ln2_w[i - 2]);
    a.bet_off[i] = i == 0 ? (n.use_fn ? n.g.fn_b : -1) : (i == 1 ? n.g.ln1_b : n.g.ln2_b[i - 2]);
  }
  a.hw_off = n.g.head_w; a.hb_off = n.g.head_b;
  a.whf = (long long)pl.whf; a.whft = (long long)pl.whft; a.cvh = (long long)pl.cvh; a.ticket = (long long)pl.unf_ticket;
  return a;
}

int pack_launch(const NetDev& n, const float* params, float* ws, int rows, bool round_tf32, int sm, cudaStream_t st) {
  const Plan pl = make_plan(n, rows, sm);
  const PackArgs a = make_pack_args(n, pl, round_tf32);
  const int warps = pl.Lh * pl.H + 32;
  Timed tm(T_PACK, st);
  big_pack_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(a, params, ws);
  return check_launch("big_pack_kernel");
}

static int run_forward(const NetDev& n, const Plan& pl, float* ws, const float* input, const int32_t* rows, int n_rows, bool tf32,
                       int sm, cudaStream_t st, bool inputs_prepared = false) {
  const int act = n.use_relu ? ACT_RELU : ACT_TANH;
  int rc = MAPPO_OK;
  if (!inputs_prepared) {
    Timed tm(T_FEATNORM, st);
    big_featnorm_kernel<<<(n_rows * 32 + 255) / 256, 256, 0, st>>>(input, rows, n_rows, n.in_dim, pl.K0p, n.use_fn, tf32 ? 1 : 0, ws + pl.x0);
    rc = check_launch("big_featnorm_kernel");
  }
  if (rc) return rc;
  for (int i = 0; i < pl.Lh; ++i) {
    LinOperands o;
    memset(&o, 0, sizeof(o));
    o.A = i == 0 ? ws + pl.x0 : ws + pl.act[i]; o.lda = i == 0 ? pl.K0p : pl.Hx;
    o.W = ws + pl.wf[i]; o.ldw = i == 0 ? pl.K0p : pl.H;
    o.out = ws + pl.act[i + 1]; o.ldo = pl.Hx; o.sm_count = sm;
    EpiFwd::Args ea;
    ea.colvec = ws + pl.cv[i];
    ea.stats_in = i == 0 ? nullptr : reinterpret_cast<const float2*>(ws + pl.stats[i]);
    ea.stats_out = reinterpret_cast<float2*>(ws + pl.stats[i + 1]);
    ea.out = ws + pl.act[i + 1]; ea.ld_out = pl.Hx; ea.N = pl.H; ea.n_rows = n_rows; ea.act = act; ea.round_tf32 = tf32 ? 1 : 0;
    LinShape sh;
    memset(&sh, 0, sizeof(sh));
    sh.n_rows = n_rows; sh.K = i == 0 ? pl.K0p : pl.H; sh.N = pl.H; sh.BN = pl.H % 256 == 0 ? 256 : 128; sh.store_out = 1;
    {
      Timed tm(T_FWD, st);
      rc = tf32 ? lin_fwd_launch(o, ea, sh, st) : ref_lin_fwd_launch(o, ea, sh, ws + pl.scratch, st);
    }
    if (rc) return rc;
  }
  return MAPPO_OK;
}

static int run_grad(const Plan& pl, float* ws, const float* P, int ldp, int Pw, int M, const float* Q, int ldq_in, int Qw, int rows, bool tf32,
                    int sm, cudaStream_t st, int* splits_out = nullptr) {
  if (Qw > 3 * 256 + 320) { set_error("big net: gradient GEMM operand %d columns wide", Qw); return MAPPO_ERR_UNSUPPORTED; }
  bool pair = false;
  const GradShape g = make_grad_shape(rows, M, Pw, Qw, sm, &pair);
  if ((size_t)g.splits * M * g.ldq > pl.partial_floats) { set_error("big net: gradient partial buffer too small"); return MAPPO_ERR_INVALID; }
  int rc;
  {
    Timed tm(T_GRAD, st);
    rc = !tf32 ? ref_grad_gemm_launch(P, ldp, Q, ldq_in, ws + pl.partial, g, st)
               : (pair ? grad_gemm_pair_launch(P, ldp, Q, ldq_in, ws + pl.partial, g, st) : grad_gemm_launch(P, ldp, Q, ldq_in, ws + pl.partial, g, st));
  }
  if (rc) return rc;
  if (splits_out) { *splits_out = g.splits; return MAPPO_OK; }     // the consumer sums the partials itself
  Timed tm(T_FINISH, st);
  return grad_reduce_launch(ws + pl.partial, g.splits, M * g.ldq, ws + pl.gsum, nullptr, nullptr, st);
}

// forward + loss + backward of one net; the complete flat gradient goes to `grad`, loss sums to loss_out
int update_launch(const NetDev& n, const float* params, const BatchDev& b, const LossDev& L, const double* norm_stats,
                  const double* adv_stats, const float* vn_state, float* grad, double* loss_out, float* ws, bool tf32, int sm,
                  cudaStream_t st, bool inputs_prepared) {
  if (!supported(n)) { set_error("big net path: unsupported configuration (hidden %d, heads %d, in_dim %d)", n.hid, n.head_total, n.in_dim); return MAPPO_ERR_UNSUPPORTED; }
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) { set_error("big net path: workspace NULL or not 256-byte aligned"); return MAPPO_ERR_INVALID; }
  const int rows = b.n_rows;
  const Plan pl = make_plan(n, rows, sm);
  const int act = n.use_relu ? ACT_RELU : ACT_TANH;
  int rc = pack_launch(n, params, ws, rows, tf32, sm, st);
  if (rc) return rc;
  rc = run_forward(n, pl, ws, n.is_critic ? b.share_obs : b.obs, b.rows, rows, tf32, sm, st, inputs_prepared);
  if (rc) return rc;
  const int Lh = pl.Lh;
  {   // heads + losses
    LinOperands o;
    memset(&o, 0, sizeof(o));
    o.A = ws + pl.act[Lh]; o.lda = pl.Hx; o.W = ws + pl.whf; o.ldw = pl.H; o.out = ws + pl.Ph; o.ldo = 32; o.sm_count = sm;
    EpiHead::Args ea;
    ea.colvec = ws + pl.cvh; ea.stats = reinterpret_cast<const float2*>(ws + pl.stats[Lh]);
    ea.mprime_out = reinterpret_cast<float2*>(ws + pl.mprime[Lh]);
    ea.n = n; ea.b = b; ea.L = L; ea.norm_stats = norm_stats; ea.adv_stats = adv_stats; ea.vn_state = vn_state; ea.loss_out = loss_out;
    ea.H = pl.H; ea.n_rows = rows; ea.round_tf32 = tf32 ? 1 : 0;
    LinShape sh;
    memset(&sh, 0, sizeof(sh));
    sh.n_rows = rows; sh.K = pl.H; sh.N = 32; sh.BN = 32; sh.store_out = b.eval_only ? 0 : 1;
    {
      Timed tm(T_HEAD, st);
      rc = tf32 ? lin_head_launch(o, ea, sh, st) : ref_lin_head_launch(o, ea, sh, ws + pl.scratch, st);
    }
    if (rc) return rc;
  }
  if (b.eval_only) return MAPPO_OK;
  if (!grad) { set_error("big net path: gradient output is NULL"); return MAPPO_ERR_INVALID; }
  const PackArgs pa = make_pack_args(n, pl, tf32);
  // head weight gradient: G_h[k][j] = sum_rows a_L_ext[row][k] Ph[row][j]
  rc = run_grad(pl, ws, ws + pl.act[Lh], pl.Hx, pl.Hx, pl.Hx, ws + pl.Ph, 32, 32, rows, tf32, sm, st);
  if (rc) return rc;
  {
    Timed tm(T_FINISH, st);
    big_unfold_head_kernel<<<(pl.H + 31) / 32, 256, 0, st>>>(ws + pl.gsum, params, grad, pl.H, n.head_total, n.g.head_w, n.g.head_b,
                                                            pa.gam_off[Lh], pa.bet_off[Lh]);
    rc = check_launch("big_unfold_head_kernel");
  }
  if (rc) return rc;
  // walk down: l = Lh .. 1 produces P_l (gradient w.r.t. the pre-activation that made a_l) and the gradient of matrix l - 1
  const float* Pup = ws + pl.Ph;
  int ld_up = 32, K_up = 32;
  for (int l = Lh; l >= 1; --l) {
    float* Pl = ws + pl.P[l & 1];
    LinOperands o;
    memset(&o, 0, sizeof(o));
    o.A = Pup; o.lda = ld_up;
    o.W = l == Lh ? ws + pl.whft : ws + pl.wft[l]; o.ldw = l == Lh ? 32 : pl.H;
    o.out = Pl; o.ldo = pl.H; o.ain = ws + pl.act[l]; o.ldain = pl.Hx; o.sm_count = sm;
    EpiBwd::Args ea;
    ea.colvec = ws + pl.cv[l - 1];
    ea.stats = reinterpret_cast<const float2*>(ws + pl.stats[l]);
    ea.mprime = reinterpret_cast<const float2*>(ws + pl.mprime[l]);
    ea.stats_prev = l > 1 ? reinterpret_cast<const float2*>(ws + pl.stats[l - 1]) : nullptr;
    ea.mprime_out = l > 1 ? reinterpret_cast<float2*>(ws + pl.mprime[l - 1]) : nullptr;
    ea.N = pl.H; ea.n_rows = rows; ea.act = act; ea.round_tf32 = tf32 ? 1 : 0;
    LinShape sh;
    memset(&sh, 0, sizeof(sh));
    sh.n_rows = rows; sh.K = K_up; sh.N = pl.H; sh.BN = pl.H % 256 == 0 ? 256 : 128; sh.store_out = 1;
    {
      Timed tm(T_BWD, st);
      rc = tf32 ? lin_bwd_launch(o, ea, sh, st) : ref_lin_bwd_launch(o, ea, sh, ws + pl.scratch, st);
    }
    if (rc) return rc;
    // weight gradient of matrix l - 1: G[o][k] = sum_rows P_l[row][o] Q[row][k],  Q = x0 (l == 1) or the extended a_{l-1}
    const int i = l - 1;
    const float* Q = i == 0 ? ws + pl.x0 : ws + pl.act[i];
    const int ldq = i == 0 ? pl.K0p : pl.Hx;
    int splits = 1;
    rc = run_grad(pl, ws, Pl, pl.H, pl.H, pl.H, Q, ldq, ldq, rows, tf32, sm, st, &splits);
    if (rc) return rc;
    const int K = i == 0 ? n.in_dim : pl.H;
    {
      Timed tm(T_FINISH, st);
      big_unfold_kernel<<<dim3((K + 31) / 32, kUnfY), 256, 0, st>>>(ws + pl.partial, splits, ldq, params, grad, pl.H, K, i == 0 ? 1 : 0,
                                                                   pa.w_off[i], pa.b_off[i], pa.gam_off[i], pa.bet_off[i],
                                                                   ws + pl.unf_part, reinterpret_cast<int*>(ws + pl.unf_ticket));
      rc = check_launch("big_unfold_kernel");
    }
    if (rc) return rc;
    Pup = Pl; ld_up = pl.H; K_up = pl.H;
  }
  return MAPPO_OK;
}

// ---- kernel-level test entries (tests/test_gpu_bignet.py): the GEMM kernels in isolation against torch.matmul ----
// out[rows, N] (ld N + 32) = relu(A[rows, K] W[N, K]^T) through big_lin_kernel<EpiFwd> (or the FFMA build), + row stats
int debug_lin(const float* A, int lda, const float* W, int ldw, float* out, float* stats, const float* colvec, float* scratch,
              int rows, int K, int N, bool tf32, int sm, cudaStream_t st) {
  LinOperands o;
  memset(&o, 0, sizeof(o));
  o.A = A; o.lda = lda; o.W = W; o.ldw = ldw; o.out = out; o.ldo = N + kExt; o.sm_count = sm;
  EpiFwd::Args ea;
  ea.colvec = colvec; ea.stats_in = nullptr; ea.stats_out = reinterpret_cast<float2*>(stats); ea.out = out; ea.ld_out = N + kExt;
  ea.N = N; ea.n_rows = rows; ea.act = ACT_RELU; ea.round_tf32 = 0;
  LinShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.n_rows = rows; sh.K = K; sh.N = N; sh.BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 32); sh.store_out = 1;
  return tf32 ? lin_fwd_launch(o, ea, sh, st) : ref_lin_fwd_launch(o, ea, sh, scratch, st);
}
// gsum[M, Qw] = P[rows, :M]^T Q[rows, :Qw] through big_grad_kernel (or the FFMA build) + the slot reduction
int debug_grad(const float* P, int ldp, int Pw, int M, const float* Q, int ldq, int Qw, int rows, float* partial, float* gsum,
               bool tf32, int sm, cudaStream_t st) {
  bool pair = false;
  const GradShape g = make_grad_shape(rows, M, Pw, Qw, sm, &pair);
  int rc = !tf32 ? ref_grad_gemm_launch(P, ldp, Q, ldq, partial, g, st)
                 : (pair ? grad_gemm_pair_launch(P, ldp, Q, ldq, partial, g, st) : grad_gemm_launch(P, ldp, Q, ldq, partial, g, st));
  if (rc) return rc;
  return grad_reduce_launch(partial, g.splits, M * g.ldq, gsum, nullptr, nullptr, st);
}
int debug_grad_splits(int rows, int M, int Pw, int Qw, int sm) { return make_grad_shape(rows, M, Pw, Qw, sm).splits; }

// rollout inference of one net on n_rows rows: packed weights must already sit at the front of `ws` (pack_launch)
int policy_launch(const NetDev& n, float* ws, const float* input, int n_rows, const EpiSample::Args& sample_in, bool tf32, int sm,
                  cudaStream_t st) {
  if (!supported(n)) { set_error("big net path: unsupported configuration"); return MAPPO_ERR_UNSUPPORTED; }
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) { set_error("big net path: workspace (weight image) NULL or not 256-byte aligned"); return MAPPO_ERR_INVALID; }
  const Plan pl = make_plan(n, n_rows, sm);
  int rc = run_forward(n, pl, ws, input, nullptr, n_rows, tf32, sm, st);
  if (rc) return rc;
  LinOperands o;
  memset(&o, 0, sizeof(o));
  o.A = ws + pl.act[pl.Lh]; o.lda = pl.Hx; o.W = ws + pl.whf; o.ldw = pl.H; o.sm_count = sm;
  EpiSample::Args ea = sample_in;
  ea.colvec = ws + pl.cvh; ea.stats = reinterpret_cast<const float2*>(ws + pl.stats[pl.Lh]); ea.n = n; ea.n_rows = n_rows;
  LinShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.n_rows = n_rows; sh.K = pl.H; sh.N = 32; sh.BN = 32; sh.store_out = 0;
  return tf32 ? lin_sample_launch(o, ea, sh, st) : ref_lin_sample_launch(o, ea, sh, ws + pl.scratch, st);
}

}  // namespace big
}  // namespace mappo
