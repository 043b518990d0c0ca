// rollout_closed.cuh -- the CLOSED rollout loop of one iteration as ONE launch: policy step -> MPE simple_spread world step
// -> insert, T times, plus the bootstrap value (SURVEY.md section 8(f) row f1 on top of a8 / a2).
// (included by policy_step.cu only, after rollout_mlp.cuh)
//
// Unlike rollout_fast_kernel there is no staged feed: the observation of step t + 1 is produced from the action of step t
// inside the kernel, so the sequential dependence of on-policy rollouts is real here.  A CTA owns kCG worlds: for every
// world M actor warps and M critic warps (one row each, the warp-per-row path of rollout_mlp.cuh, both weight images in
// shared memory) and ONE environment thread that keeps the float64 world in registers (mpe_world.cuh).  Per step:
//   actor / critic warps: forward, sample, write values / actions / log-probs (and the row into its storage slot)
//   __syncthreads
//   environment thread: read the world's M actions, step the physics, reward / done -> storage, reset if the episode ended,
//                       new observations -> shared memory
//   __syncthreads
//   every warp picks its next row from shared memory (the critic's share_obs row is the world's M observations
//   back to back, mpe_runner.py:133-135).
#pragma once
#include "rollout_mlp.cuh"
#include "mpe_world.cuh"

namespace mappo {

constexpr int kCG = 2;               // worlds per CTA


// MT / LT: compile-time agent / landmark counts (0 = runtime, launch bound for 8 agents)
template <int MT, int LT>
__global__ void __launch_bounds__(64 * kCG * (MT ? MT : kMpeMaxAgents))
rollout_closed_kernel(const NetDev na, const NetDev nc, const ClosedArgs ca) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t wbar;
  const RolloutArgs& a = ca.r;
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const int M = MT ? MT : ca.M, L = LT ? LT : ca.L, E = a.E, T = a.T, N = E / M;
  const int rows = kCG * M;                                   // actor warps [0, rows), critic warps [rows, 2 rows)
  const int which = wq >= rows ? 1 : 0;
  const int rl = which ? wq - rows : wq, env_local = rl / M, m = rl - env_local * M;
  const int env = blockIdx.x * kCG + env_local;
  const int g = env < N ? env * M + m : -1;
  const NetDev& n = which == 0 ? na : nc;
  const FastImg fa = make_fast_img(na), fc = make_fast_img(nc);
  const int D = na.in_dim;                                    // 4 + 2 L + 4 (M - 1)
  float* obs_s = smem + fa.total + fc.total + 2 * rows * kFWarpScratch;     // [kCG][M][D]

  // ---- both weight images by TMA, one mbarrier ----
  if (tid == 0) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&wbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)((fa.total + fc.total) * 4)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(a.image[0]), "r"((uint32_t)(fa.total * 4)), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem + fa.total)), "l"(a.image[1]), "r"((uint32_t)(fc.total * 4)), "r"(bar) : "memory");
  }
  FastCtx c;
  c.f = which == 0 ? fa : fc;
  c.sW = smem + (which == 0 ? 0 : fa.total);
  c.bufA = smem + fa.total + fc.total + wq * kFWarpScratch;
  c.bufB = c.bufA + 64;
  // slot 0 of the storage holds the observations the env produced last (warm-up / previous iteration)
  float* store_in = which == 0 ? a.obs : a.share_obs;
  const int in = n.in_dim;
  float x[2];
  load_row_lane(store_in, g, in, lane, x);
  // the environment thread of a world: lane 0 of its first actor warp
  const bool env_thread = which == 0 && m == 0 && lane == 0 && env < N;
  MpeWorld w;
  if (env_thread) mpe_world_load<MT, LT>(w, M, L, ca.apos, ca.avel, ca.lpos, ca.step_count, env);
  const uint64_t env_ctr0 = (env_thread && !ca.reset_states) ? *ca.env_counter : 0ull;
  __syncthreads();
  {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&wbar);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
  }
  const int Atot = na.head_total, as = na.n_heads;
  const uint64_t rng0 = (!a.exp_noise && which == 0) ? *a.rng_offset : 0ull;
  long long t_last = clock64();
#pragma unroll 1
  for (int t = 0; t <= T; ++t) {
    PolStep p;
    p.in = nullptr;
    p.in_copy = t == 0 ? nullptr : store_in + (size_t)t * E * in;
    p.h_in = nullptr; p.h_out = nullptr; p.done_now = nullptr; p.done_prev = nullptr;
    p.masks = a.masks; p.masks_copy = nullptr;
    p.avail = nullptr; p.avail_copy = nullptr;
    p.exp_noise = (a.exp_noise && t < T) ? a.exp_noise + (size_t)t * E * Atot : nullptr;
    p.rng_ctr = rng0 + (uint64_t)t * (uint64_t)E;
    p.values = a.value_preds + (size_t)t * E;
    p.actions = t < T ? a.actions + (size_t)t * E * as : nullptr;
    p.actions_i64 = nullptr;
    p.logp = t < T ? a.logp + (size_t)t * E * as : nullptr;
    p.forward = (t < T) || which == 1;                            // slot T: only the critic's bootstrap value
    fast_step(n, which, c, p, x, g, lane, 0, 0, a.rng_seed, t_last, tid);
    if (t == T) break;
    __syncthreads();                                              // the world's actions of step t are visible
    if (env_thread) {
      int act[kMpeMaxAgents];
      for (int q = 0; q < M; ++q) act[q] = (int)a.actions[((size_t)t * E + (size_t)env * M + q) * as];
      bool done;
      const double reward = mpe_world_step<MT, LT>(w, M, L, act, ca.episode_length, &done);
      if (done)                                                   // env_wrappers.py:146-152
        mpe_world_reset<MT, LT>(w, M, L, ca.reset_states ? ca.reset_states + ((size_t)t * N + env) * 2 * (M + L) : nullptr,
                        ca.env_seed, env_ctr0 + (uint64_t)t * N + env);
      for (int q = 0; q < M; ++q) {
        a.rewards[(size_t)t * E + (size_t)env * M + q] = (float)reward;                   // insert: rewards of slot t,
        a.masks[(size_t)(t + 1) * E + (size_t)env * M + q] = done ? 0.f : 1.f;            // masks of slot t + 1
        mpe_world_obs<MT, LT>(w, M, L, q, obs_s + ((size_t)env_local * M + q) * D);
      }
    }
    __syncthreads();                                              // the next observations are in shared memory
    {
      const float* src = obs_s + (size_t)env_local * M * D + (which == 0 ? m * D : 0);   // critic: the world's M rows
      x[0] = (g >= 0 && lane < in) ? src[lane] : 0.f;
      x[1] = (g >= 0 && lane + 32 < in) ? src[lane + 32] : 0.f;
    }
  }
  if (env_thread) mpe_world_store<MT, LT>(w, M, L, ca.apos, ca.avel, ca.lpos, ca.step_count, env);
}

inline size_t closed_smem_bytes(const NetDev& na, const NetDev& nc, int M) {
  return (size_t)(make_fast_img(na).total + make_fast_img(nc).total + 2 * kCG * M * kFWarpScratch + kCG * M * na.in_dim + 4) *
         sizeof(float);
}

}  // namespace mappo
