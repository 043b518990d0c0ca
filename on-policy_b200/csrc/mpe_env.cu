// mpe_env.cu -- N vectorised MPE `simple_spread` worlds stepped on the device.
//
// Replaces, for this scenario, SubprocVecEnv.step -> MultiAgentEnv.step -> World.step + the scenario callbacks
// (envs/env_wrappers.py:140-154, envs/mpe/environment.py:115-146, envs/mpe/core.py:207-323,
// envs/mpe/scenarios/simple_spread.py:32-103): action decoding, action + contact forces, damping / integration, reward
// (minimum agent distance per landmark, collision penalties incl. the reference's self-"collision", summed over the
// agents of the world), observation, done = step >= episode_length, auto-reset with the reset observation replacing the
// terminal one.  The state is float64 and every expression keeps the reference's order of operations, so a trajectory
// follows the NumPy one to the last bits of exp / log1p (the only non-IEEE-exact operations involved).
// One thread per world: a world is 6 entities, the work per step is a few hundred flops -- pure latency; what this buys
// is a rollout without a host round trip (the reference pays a pipe round trip + NumPy physics per env step).
#include "launch_args.h"
#include "rng.cuh"

namespace mappo {

constexpr int kMpeMaxAgents = 8, kMpeMaxLandmarks = 8;
constexpr double kAgentSize = 0.15;      // simple_spread.py:22
constexpr double kContactForce = 1e2;    // core.py:128
constexpr double kContactMargin = 1e-3;  // core.py:129
constexpr double kDamping = 0.25;        // core.py:126
constexpr double kDt = 0.1;              // core.py:124
constexpr double kSensitivity = 5.0;     // environment.py:243


// np.logaddexp(0, y) (numpy/core/src/npymath/npy_math_internal.h.src: npy_logaddexp)
__device__ __forceinline__ double logaddexp0(double y) {
  if (y == 0.0) return 0.6931471805599453094172321214581766;
  const double tmp = 0.0 - y;
  if (tmp > 0) return 0.0 + log1p(exp(-tmp));
  if (tmp <= 0) return y + log1p(exp(tmp));
  return tmp;
}

__device__ __forceinline__ void mpe_reset_world(const MpeArgs& a, int e, double (*ap)[2], double (*av)[2], double (*lp)[2]) {
  const int M = a.M, L = a.L;
  if (a.reset_states) {
    const double* s = a.reset_states + (size_t)e * 2 * (M + L);
    for (int m = 0; m < M; ++m) { ap[m][0] = s[2 * m]; ap[m][1] = s[2 * m + 1]; }
    for (int l = 0; l < L; ++l) { lp[l][0] = s[2 * (M + l)]; lp[l][1] = s[2 * (M + l) + 1]; }
  } else {                               // uniform(-1, 1) agents, 0.8 * uniform(-1, 1) landmarks (simple_spread.py:39-45)
    const uint64_t ctr = *a.rng_counter + (uint64_t)e;
    for (int q = 0; q < (2 * (M + L) + 3) / 4; ++q) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x4d504500u + q, 0u),
                                    make_uint2((uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32)));
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        if (i >= 2 * (M + L)) break;
        const double u = -1.0 + 2.0 * (((double)w[j] + 0.5) * 2.3283064365386962890625e-10);
        if (i < 2 * M) ap[i >> 1][i & 1] = u;
        else lp[(i - 2 * M) >> 1][i & 1] = 0.8 * u;
      }
    }
  }
  for (int m = 0; m < M; ++m) av[m][0] = av[m][1] = 0.0;
}

__global__ void __launch_bounds__(128) mpe_spread_kernel(const MpeArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.N) return;
  const int M = a.M, L = a.L;
  double ap[kMpeMaxAgents][2], av[kMpeMaxAgents][2], lp[kMpeMaxLandmarks][2];
  for (int m = 0; m < M; ++m)
    for (int d = 0; d < 2; ++d) { ap[m][d] = a.apos[((size_t)e * M + m) * 2 + d]; av[m][d] = a.avel[((size_t)e * M + m) * 2 + d]; }
  for (int l = 0; l < L; ++l)
    for (int d = 0; d < 2; ++d) lp[l][d] = a.lpos[((size_t)e * L + l) * 2 + d];
  int step = a.step_count[e];
  double reward = 0.0;
  bool done = false;
  if (!a.actions) {                      // reset()
    mpe_reset_world(a, e, ap, av, lp);
    step = 0;
  } else {
    // ---- forces: action (environment.py:232-246, core.py:229-238), then contacts between agents (core.py:241-323) ----
    double f[kMpeMaxAgents][2];
    for (int m = 0; m < M; ++m) {
      const int act = (int)a.actions[(size_t)e * M + m];
      double u0 = 0.0, u1 = 0.0;
      u0 += (act == 1 ? 1.0 : 0.0) - (act == 2 ? 1.0 : 0.0);
      u1 += (act == 3 ? 1.0 : 0.0) - (act == 4 ? 1.0 : 0.0);
      u0 *= kSensitivity; u1 *= kSensitivity;
      f[m][0] = 1.0 * u0 + 0.0;
      f[m][1] = 1.0 * u1 + 0.0;
    }
    for (int ia = 0; ia < M; ++ia)
      for (int ib = ia + 1; ib < M; ++ib) {
        const double dx = ap[ia][0] - ap[ib][0], dy = ap[ia][1] - ap[ib][1];
        const double dist = sqrt(dx * dx + dy * dy);
        const double k = kContactMargin;
        const double pen = logaddexp0(-(dist - (kAgentSize + kAgentSize)) / k) * k;
        const double fx = kContactForce * dx / dist * pen, fy = kContactForce * dy / dist * pen;
        f[ia][0] = fx + f[ia][0]; f[ia][1] = fy + f[ia][1];
        f[ib][0] = -fx + f[ib][0]; f[ib][1] = -fy + f[ib][1];
      }
    // ---- integrate (core.py:267-281) ----
    for (int m = 0; m < M; ++m)
      for (int d = 0; d < 2; ++d) {
        double v = av[m][d] * (1 - kDamping);
        v += (f[m][d] / 1.0) * kDt;
        av[m][d] = v;
        ap[m][d] += v * kDt;
      }
    step += 1;
    // ---- reward (simple_spread.py:72-85), shared = sum over agents (environment.py:139-142) ----
    for (int m = 0; m < M; ++m) {
      double rew = 0.0;
      for (int l = 0; l < L; ++l) {
        double mn = 0.0;
        for (int q = 0; q < M; ++q) {
          const double dx = ap[q][0] - lp[l][0], dy = ap[q][1] - lp[l][1];
          const double d = sqrt(dx * dx + dy * dy);
          mn = (q == 0 || d < mn) ? d : mn;
        }
        rew -= mn;
      }
      for (int q = 0; q < M; ++q) {        // q == m included: an agent "collides" with itself in the reference
        const double dx = ap[q][0] - ap[m][0], dy = ap[q][1] - ap[m][1];
        if (sqrt(dx * dx + dy * dy) < kAgentSize + kAgentSize) rew -= 1;
      }
      reward = m == 0 ? rew : reward + rew;
    }
    done = step >= a.episode_length;
    if (done) {                            // env_wrappers.py:146-152: the reset observation replaces the terminal one
      mpe_reset_world(a, e, ap, av, lp);
      step = 0;
    }
  }
  // ---- state back, observations (simple_spread.py:87-103), share_obs = all agents' obs (mpe_runner.py:133-135) ----
  for (int m = 0; m < M; ++m)
    for (int d = 0; d < 2; ++d) { a.apos[((size_t)e * M + m) * 2 + d] = ap[m][d]; a.avel[((size_t)e * M + m) * 2 + d] = av[m][d]; }
  if (!a.actions || done)
    for (int l = 0; l < L; ++l)
      for (int d = 0; d < 2; ++d) a.lpos[((size_t)e * L + l) * 2 + d] = lp[l][d];
  a.step_count[e] = step;
  const int D = 4 + 2 * L + 4 * (M - 1);
  for (int m = 0; m < M; ++m) {
    float* o = a.obs + ((size_t)e * M + m) * D;
    int c = 0;
    o[c++] = (float)av[m][0]; o[c++] = (float)av[m][1];
    o[c++] = (float)ap[m][0]; o[c++] = (float)ap[m][1];
    for (int l = 0; l < L; ++l) { o[c++] = (float)(lp[l][0] - ap[m][0]); o[c++] = (float)(lp[l][1] - ap[m][1]); }
    for (int q = 0; q < M; ++q)
      if (q != m) { o[c++] = (float)(ap[q][0] - ap[m][0]); o[c++] = (float)(ap[q][1] - ap[m][1]); }
    for (int q = 0; q < M; ++q)
      if (q != m) { o[c++] = 0.f; o[c++] = 0.f; }            // silent agents: communication state is zero
    if (a.actions) {
      if (a.rewards) a.rewards[(size_t)e * M + m] = (float)reward;
      if (a.dones) a.dones[(size_t)e * M + m] = done ? 1.f : 0.f;
    }
  }
  if (a.share_obs) {
    const float* src = a.obs + (size_t)e * M * D;            // written above by this thread
    for (int m = 0; m < M; ++m) {
      float* s = a.share_obs + ((size_t)e * M + m) * (size_t)(M * D);
      for (int i = 0; i < M * D; ++i) s[i] = src[i];
    }
  }
}

int mpe_spread_launch(const MpeArgs& a, cudaStream_t st) {
  if (a.M < 1 || a.M > kMpeMaxAgents || a.L < 1 || a.L > kMpeMaxLandmarks) { set_error("mpe_spread: %d agents / %d landmarks outside [1,8]", a.M, a.L); return MAPPO_ERR_UNSUPPORTED; }
  mpe_spread_kernel<<<(a.N + 127) / 128, 128, 0, st>>>(a);
  return check_launch("mpe_spread_kernel");
}

}  // namespace mappo
