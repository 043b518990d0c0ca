// policy_step.cu -- rollout inference for one env step over E = N*M rows (a8 of SURVEY.md section 8).
// Replaces R_MAPPOPolicy.get_actions / get_values / act (algorithms/r_mappo/algorithm/rMAPPOPolicy.py:48-127):
// actor = feature LN -> MLP -> [GRU step + LN] -> categorical heads -> sample/mode + log-prob,
// critic = the same trunk -> value.  blockIdx.y picks the net, blockIdx.x the 32-row tile; results are
// written straight into the rollout-storage slots handed in by the caller.
#include "net_tiles.cuh"
#include "launch_args.h"

namespace mappo {

constexpr int kPolTR = 32;

// Philox4x32-10 (Salmon et al. 2011), counter-based: no state to keep between launches.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// ---- packed weight image: the shared-memory weight layout (odd leading dimensions), pre-built in global memory so a
//      CTA fetches it with ONE TMA bulk copy instead of ~70 address-computing cp.async per thread ----
__device__ __forceinline__ int simt_image_offset(const NetDev& n, const SmemW& s, int i) {
  const int H = n.hid, I = n.in_dim;
  auto mat = [&](int base, int rows_cols, int img, int ld, int cols) { const int t = i - base; const int r = t / cols; return img + r * ld + (t - r * cols); };
  (void)mat;
  const mappo_net_layout_t& g = n.g;
  if (n.use_fn) {
    if (i >= g.fn_w && i < g.fn_w + I) return s.fn_w + (i - g.fn_w);
    if (i >= g.fn_b && i < g.fn_b + I) return s.fn_b + (i - g.fn_b);
  }
  if (i >= g.fc1_w && i < g.fc1_w + H * I) { const int t = i - g.fc1_w, r = t / I; return s.fc1_w + r * s.ld1 + (t - r * I); }
  if (i >= g.fc1_b && i < g.fc1_b + H) return s.fc1_b + (i - g.fc1_b);
  if (i >= g.ln1_w && i < g.ln1_w + H) return s.ln1_w + (i - g.ln1_w);
  if (i >= g.ln1_b && i < g.ln1_b + H) return s.ln1_b + (i - g.ln1_b);
  for (int l = 0; l < n.layer_n; ++l) {
    if (i >= g.fc2_w[l] && i < g.fc2_w[l] + H * H) { const int t = i - g.fc2_w[l], r = t / H; return s.fc2_w[l] + r * s.ldh + (t - r * H); }
    if (i >= g.fc2_b[l] && i < g.fc2_b[l] + H) return s.fc2_b[l] + (i - g.fc2_b[l]);
    if (i >= g.ln2_w[l] && i < g.ln2_w[l] + H) return s.ln2_w[l] + (i - g.ln2_w[l]);
    if (i >= g.ln2_b[l] && i < g.ln2_b[l] + H) return s.ln2_b[l] + (i - g.ln2_b[l]);
  }
  if (n.recurrent) {
    if (i >= g.gru_wih && i < g.gru_wih + 3 * H * H) { const int t = i - g.gru_wih, r = t / H; return s.wih + r * s.ldh + (t - r * H); }
    if (i >= g.gru_whh && i < g.gru_whh + 3 * H * H) { const int t = i - g.gru_whh, r = t / H; return s.whh + r * s.ldh + (t - r * H); }
    if (i >= g.gru_bih && i < g.gru_bih + 3 * H) return s.bih + (i - g.gru_bih);
    if (i >= g.gru_bhh && i < g.gru_bhh + 3 * H) return s.bhh + (i - g.gru_bhh);
    if (i >= g.rnn_ln_w && i < g.rnn_ln_w + H) return s.rln_w + (i - g.rnn_ln_w);
    if (i >= g.rnn_ln_b && i < g.rnn_ln_b + H) return s.rln_b + (i - g.rnn_ln_b);
  }
  if (i >= g.head_w && i < g.head_w + n.head_total * H) { const int t = i - g.head_w, r = t / H; return s.head_w + r * s.ldh + (t - r * H); }
  if (i >= g.head_b && i < g.head_b + n.head_total) return s.head_b + (i - g.head_b);
  return -1;
}

__global__ void __launch_bounds__(256) pack_rollout_kernel(const NetDev n, const float* __restrict__ p, float* __restrict__ img) {
  const SmemW s = make_smem_w(n, true);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n.g.total; i += gridDim.x * blockDim.x) {
    const int o = simt_image_offset(n, s, i);
    if (o >= 0) img[o] = p[i];
  }
}

struct PolSmem { int w, x0, xh0, s0, s1, h, gi, gh, stats, red, rowid, total; };

__host__ __device__ inline PolSmem make_pol_smem(const NetDev& n, const SmemW& s) {
  constexpr int LD = Tile<kPolTR>::LD;
  PolSmem u;
  int o = 0;
  const int inT = ((n.in_dim + 3) & ~3) * LD, hT = n.hid * LD;
  u.w = o; o += s.total;
  u.x0 = o; o += inT;
  u.xh0 = o; o += n.use_fn ? inT : 0;
  u.s0 = o; o += hT;
  u.s1 = o; o += hT;
  u.h = o; o += n.recurrent ? hT : 0;
  u.gi = o; o += n.recurrent ? 3 * hT : 0;      // r, z pre-activations (input + hidden parts summed), n input part
  u.gh = o; o += n.recurrent ? hT : 0;          // n hidden part (multiplied by r before the tanh)
  const int lg = ((n.head_total + 3) & ~3) * LD;     // logits share the gi area when recurrent
  if (!n.recurrent) { u.gi = o; o += lg; }
  u.stats = o; o += 2 * (kMaxLayers + 3) * kPolTR;
  u.red = o; o += 8 * kPolTR;
  u.rowid = o; o += kPolTR;
  u.total = o;
  return u;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int NJH>
__global__ void __launch_bounds__(4 * kPolTR)
policy_step_kernel(const NetDev na, const NetDev nc, const PolArgs a, int first_net) {
  constexpr int TR = kPolTR;
  constexpr int LD = Tile<TR>::LD;
  constexpr int NT = Tile<TR>::NT;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int which = first_net + blockIdx.y;           // 0 actor, 1 critic
  const NetDev& n = which == 0 ? na : nc;
  const SmemW s = make_smem_w(n, true);
  const PolSmem u = make_pol_smem(n, s);
  float* sW = smem + u.w;
  BaseTiles<TR> t;
  t.xh0 = smem + u.xh0;
  t.x0 = smem + u.x0;
  // fused layers read Y[l-1] and write Y[l]: ping-pong between the two scratch tiles
  for (int l = 0; l <= kMaxLayers; ++l) { t.A[l] = nullptr; t.Y[l] = smem + ((l & 1) ? u.s0 : u.s1); }
  t.keep_act = false;
  for (int l = 0; l < kMaxLayers + 2; ++l) { t.mean[l] = smem + u.stats + 2 * l * TR; t.rstd[l] = t.mean[l] + TR; }
  t.red = smem + u.red;
  int* rowid = reinterpret_cast<int*>(smem + u.rowid);
  const int H = n.hid;

  __shared__ uint64_t wbar;
  const float* image = a.image[which];
  if (image) {                                        // one TMA bulk copy of the pre-packed image (UBLKCP)
    if (tid == 0) {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&wbar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(s.total * 4)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((uint32_t)__cvta_generic_to_shared(sW)), "l"(image), "r"((uint32_t)(s.total * 4)), "r"(bar) : "memory");
    }
  } else {
    load_weights(sW, s, n, a.params[which], true, tid, NT);
  }
  const int row0 = blockIdx.x * TR;
  if (tid < TR) rowid[tid] = row0 + tid < a.n_rows ? row0 + tid : -1;
  __syncthreads();
  load_rows_T<TR>(a.in[which], n.in_dim, rowid, t.x0, tid);
  if (image) {                                        // every thread waits for the image (barrier init is ordered by the
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&wbar);   // __syncthreads() above)
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
  }
  base_forward<TR, NJH>(n, s, sW, t, tid);
  const float* feat = t.Y[n.layer_n];

  if (n.recurrent) {
    // h <- h * mask (rnn.py:27), one GRU step (torch gate order r,z,n), LN (rnn.py:79)
    float* hT = smem + u.h;
    float* gi = smem + u.gi;
    float* gh = smem + u.gh;
    for (int i = tid; i < TR * H; i += NT) {
      const int r = i / H, c = i - r * H;
      const int g = rowid[r];
      hT[c * LD + r] = g >= 0 ? a.h_in[which][(size_t)g * H + c] * a.masks[g] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int gate = 0; gate < 3; ++gate)
      tile_mm<TR, NJH>(feat, H, sW + s.wih + gate * H * s.ldh, s.ldh, 1, H, sW + s.bih + gate * H, ACT_NONE,
                       gi + gate * H * LD, tid);
    // each thread re-reads exactly the outputs it wrote, so no barrier is needed before accumulating
#pragma unroll 1
    for (int gate = 0; gate < 2; ++gate)
      tile_mm<TR, NJH>(hT, H, sW + s.whh + gate * H * s.ldh, s.ldh, 1, H, sW + s.bhh + gate * H, ACT_NONE,
                       gi + gate * H * LD, tid, true);
    tile_mm<TR, NJH>(hT, H, sW + s.whh + 2 * H * s.ldh, s.ldh, 1, H, sW + s.bhh + 2 * H, ACT_NONE, gh, tid);
    __syncthreads();
    float* hn = (feat == smem + u.s0) ? smem + u.s1 : smem + u.s0;      // new hidden state (pre-LN): the free tile
    for (int i = tid; i < TR * H; i += NT) {
      const int c = i / TR, r = i - c * TR;
      const int o = c * LD + r;
      const float rg = sigmoidf_(gi[o]);
      const float zg = sigmoidf_(gi[H * LD + o]);
      const float ng = tanhf(gi[2 * H * LD + o] + rg * gh[o]);
      hn[o] = (1.f - zg) * ng + zg * hT[o];
    }
    __syncthreads();
    for (int i = tid; i < TR * H; i += NT) {
      const int r = i / H, c = i - r * H;
      const int g = rowid[r];
      if (g >= 0 && a.h_out[which]) a.h_out[which][(size_t)g * H + c] = hn[c * LD + r];
    }
    float* ln_out = (hn == smem + u.s0) ? smem + u.s1 : smem + u.s0;    // old feat tile: gates are done with it
    tile_layernorm<TR>(hn, H, sW + s.rln_w, sW + s.rln_b, ln_out, t.mean[kMaxLayers + 1] + 2 * TR,
                       t.rstd[kMaxLayers + 1] + 2 * TR, t.red, tid);
    feat = ln_out;
  }

  float* lgT = smem + u.gi;
  const int Atot = n.head_total;
  __syncthreads();
  tile_mm<TR, 2>(feat, H, sW + s.head_w, s.ldh, 1, Atot, sW + s.head_b, ACT_NONE, lgT, tid);
  __syncthreads();
  if (tid < TR && rowid[tid] >= 0) {
    const int r = tid, g = rowid[r];
    if (which == 1) {
      if (a.values) a.values[g] = lgT[r];
    } else {
      const float* av = (a.avail && n.n_heads == 1) ? a.avail + (size_t)g * a.n_avail : nullptr;
      const int as = n.n_heads;
      uint64_t ctr = 0;
      if (!a.exp_noise && !a.deterministic) ctr = *a.rng_offset + (uint64_t)g;
      int off = 0;
      for (int k = 0; k < as; ++k) {
        const int A = n.head_dim[k];
        float lse;
        head_lse<LD>(lgT, off, A, r, av, lse);
        int best = 0;
        float bestv = -INFINITY, best_lp = 0.f;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        for (int j = 0; j < A; ++j) {
          float lgt = lgT[(off + j) * LD + r];
          if (av && av[j] == 0.f) lgt = -1e10f;
          const float lp = lgt - lse;
          const float p = expf(lp);
          float score = p;
          if (!a.deterministic) {
            float q;
            if (a.exp_noise) {
              q = a.exp_noise[(size_t)g * Atot + off + j];
            } else {
              if ((j & 3) == 0)
                rnd = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(k * 64 + (j >> 2)), 0u),
                                    make_uint2((uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32)));
              const uint32_t x = (j & 3) == 0 ? rnd.x : ((j & 3) == 1 ? rnd.y : ((j & 3) == 2 ? rnd.z : rnd.w));
              q = -logf(((float)x + 0.5f) * 2.3283064365386963e-10f);
            }
            score = p / q;                                   // torch multinomial: argmax(p / Exp(1))
          }
          if (score > bestv) { bestv = score; best = j; best_lp = lp; }
        }
        if (a.actions) a.actions[(size_t)g * as + k] = (float)best;
        if (a.actions_i64) a.actions_i64[(size_t)g * as + k] = (int64_t)best;
        if (a.logp) a.logp[(size_t)g * as + k] = best_lp;
        off += A;
      }
    }
  }
}

__global__ void counter_add_kernel(uint64_t* c, uint64_t inc) { *c += inc; }

int policy_step_launch(const NetDev* na, const NetDev* nc, const PolArgs& a, cudaStream_t st) {
  const NetDev& ref = na ? *na : *nc;
  size_t bytes = 0;
  for (const NetDev* n : {na, nc}) {
    if (!n) continue;
    if (n->hid != 64) { set_error("policy_step: hidden_size %d not built in the fused SIMT path (64 only)", n->hid); return MAPPO_ERR_UNSUPPORTED; }
    if (n->head_total > 32) { set_error("policy_step: sum(head_dim) > 32"); return MAPPO_ERR_UNSUPPORTED; }
    const SmemW s = make_smem_w(*n, true);
    const size_t b = (size_t)make_pol_smem(*n, s).total * sizeof(float);
    bytes = b > bytes ? b : bytes;
  }
  if (bytes > 227 * 1024) { set_error("policy_step: %zu B shared memory per CTA > 227 KB (in_dim too large)", bytes); return MAPPO_ERR_UNSUPPORTED; }
  auto kern = policy_step_kernel<4>;
  static thread_local size_t configured = 0;
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return check_launch("policy_step: cudaFuncSetAttribute");
    configured = bytes;
  }
  const dim3 grid((a.n_rows + kPolTR - 1) / kPolTR, (na && nc) ? 2 : 1);
  kern<<<grid, 4 * kPolTR, bytes, st>>>(na ? *na : ref, nc ? *nc : ref, a, na ? 0 : 1);
  return check_launch("policy_step_kernel");
}

int pack_rollout_launch(const NetDev& n, const float* params, float* image, cudaStream_t st) {
  pack_rollout_kernel<<<(n.g.total + 255) / 256, 256, 0, st>>>(n, params, image);
  return check_launch("pack_rollout_kernel");
}
int rollout_image_floats(const NetDev& n) { return make_smem_w(n, true).total; }

int counter_add_launch(uint64_t* c, uint64_t inc, cudaStream_t st) {
  counter_add_kernel<<<1, 1, 0, st>>>(c, inc);
  return check_launch("counter_add_kernel");
}

}  // namespace mappo
