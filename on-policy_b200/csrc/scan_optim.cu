// scan_optim.cu -- the HBM-bound kernels of the path: GAE scan, statistics, gathers, insert,
// gradient reduction and clip + Adam.  All coalesced / vectorised elementwise work with warp shuffles.
#include "common.cuh"
#include "launch_args.h"

namespace mappo {

// -------------------------------------------------------------------------------------------------
// a3 + a4 + a5: compute_returns (utils/shared_buffer.py:179-262) as one backward scan.
// One thread per lane e = (n, m); at step t the warp reads rewards[t*E + e .. +31]: fully coalesced in the
// reference's own [T, N, M, 1] layout.  Algorithmic bytes: 4*E*(4T+1) + 4*E*T for the
// emitted advantages.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ vpred, const float* __restrict__ masks,
                const float* __restrict__ bad, const float* __restrict__ active, const float* __restrict__ vn_state,
                int T, int E, float gamma, float lam, int use_gae, int use_ptl, float* __restrict__ returns,
                float* __restrict__ adv_out, double* __restrict__ adv_stats) {
  __shared__ double sred[3 * 32];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float mean = 0.f, sd = 1.f;
  if (vn_state) {
    float var;
    vn_mean_var(vn_state, mean, var);
    sd = sqrtf(var);
  }
  double st[3] = {0.0, 0.0, 0.0};
  if (e < E) {
    if (use_gae) {
      float gae = 0.f;
      float v_next = fmaf(vpred[(size_t)T * E + e], sd, mean);          // denormalize (valuenorm.py:68-79)
#pragma unroll 4
      for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * E + e, in = i + E;
        const float m_next = masks[in];
        const float v_t = fmaf(vpred[i], sd, mean);
        const float delta = rewards[i] + gamma * v_next * m_next - v_t;  // shared_buffer.py:236-238
        gae = delta + gamma * lam * m_next * gae;                         // :239
        if (use_ptl) gae *= bad[in];                                      // :194
        const float ret = gae + v_t;                                      // :240
        returns[i] = ret;
        const float adv = ret - v_t;                                      // r_mappo.py:179-182
        if (adv_out) adv_out[i] = adv;
        if (active[i] != 0.f) { st[0] += adv; st[1] += (double)adv * adv; st[2] += 1.0; }
        v_next = v_t;
      }
    } else {
      float ret = vpred[(size_t)T * E + e];                               // returns[-1] = next_value (:204, :260)
      returns[(size_t)T * E + e] = ret;
      for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * E + e, in = i + E;
        const float v_t = fmaf(vpred[i], sd, mean);
        if (use_ptl) {                                                    // :206-215
          const float b = bad[in];
          ret = (ret * gamma * masks[in] + rewards[i]) * b + (1.f - b) * v_t;
        } else {
          ret = ret * gamma * masks[in] + rewards[i];                     // :261-262
        }
        returns[i] = ret;
        const float adv = ret - v_t;
        if (adv_out) adv_out[i] = adv;
        if (active[i] != 0.f) { st[0] += adv; st[1] += (double)adv * adv; st[2] += 1.0; }
      }
    }
  }
  if (adv_stats) block_accumulate<3>(st, adv_stats, sred, threadIdx.x, blockDim.x);
}

int gae_launch(const float* rewards, const float* vpred, const float* masks, const float* bad, const float* active,
               const float* vn_state, int T, int E, float gamma, float lam, int use_gae, int use_ptl, float* returns,
               float* adv, double* adv_stats, cudaStream_t st) {
  const int nt = 128;
  gae_scan_kernel<<<(E + nt - 1) / nt, nt, 0, st>>>(rewards, vpred, masks, bad, active, vn_state, T, E, gamma, lam,
                                                    use_gae, use_ptl, returns, adv, adv_stats);
  return check_launch("gae_scan_kernel");
}

// a5 stand-alone: advantages = returns - denorm(value_preds) over n entries + masked statistics
// (r_mappo.py:179-187) for callers that filled `returns` themselves.
__global__ void __launch_bounds__(256)
advantages_kernel(const float* __restrict__ returns, const float* __restrict__ vpred, const float* __restrict__ active,
                  const float* __restrict__ vn_state, int n, float* __restrict__ adv_out, double* __restrict__ stats) {
  __shared__ double sred[3 * 32];
  float mean = 0.f, sd = 1.f;
  if (vn_state) { float var; vn_mean_var(vn_state, mean, var); sd = sqrtf(var); }
  double st[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float adv = returns[i] - fmaf(vpred[i], sd, mean);
    adv_out[i] = adv;
    if (active[i] != 0.f) { st[0] += adv; st[1] += (double)adv * adv; st[2] += 1.0; }
  }
  block_accumulate<3>(st, stats, sred, threadIdx.x, blockDim.x);
}

int advantages_launch(const float* returns, const float* vpred, const float* active, const float* vn_state, int n,
                      float* adv, double* stats, cudaStream_t st) {
  int blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  advantages_kernel<<<blocks, 256, 0, st>>>(returns, vpred, active, vn_state, n, adv, stats);
  return check_launch("advantages_kernel");
}

// -------------------------------------------------------------------------------------------------
// a4: minibatch statistics + ValueNorm.update (utils/valuenorm.py:38-55)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
minibatch_stats_kernel(const float* __restrict__ returns, const float* __restrict__ active,
                       const int32_t* __restrict__ rows, int n, double* __restrict__ stats) {
  __shared__ double sred[4 * 32];
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int g = rows ? rows[p] : p;
    const float r = returns[g];
    v[0] += active[g];
    v[1] += r;
    v[2] += (double)r * r;
    v[3] += 1.0;
  }
  block_accumulate<4>(v, stats, sred, threadIdx.x, blockDim.x);
}

// All minibatches of a train() call at once: blockIdx.y = update index u, rows of update u start at rows + u*stride.
__global__ void __launch_bounds__(256)
minibatch_stats_batch_kernel(const float* __restrict__ returns, const float* __restrict__ active,
                             const int32_t* __restrict__ rows, long long stride, int n, double* __restrict__ stats) {
  __shared__ double sred[4 * 32];
  const int32_t* r = rows + (size_t)blockIdx.y * stride;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int g = r[p];
    const float x = returns[g];
    v[0] += active[g];
    v[1] += x;
    v[2] += (double)x * x;
    v[3] += 1.0;
  }
  block_accumulate<4>(v, stats + 4 * blockIdx.y, sred, threadIdx.x, blockDim.x);
}

int minibatch_stats_batch_launch(const float* returns, const float* active, const int32_t* rows, long long stride, int n,
                                 int n_batches, double* stats, cudaStream_t st) {
  int bx = (n + 255) / 256;
  if (bx > 32) bx = 32;
  if (bx < 1) bx = 1;
  minibatch_stats_batch_kernel<<<dim3(bx, n_batches), 256, 0, st>>>(returns, active, rows, stride, n, stats);
  return check_launch("minibatch_stats_batch_kernel");
}

int minibatch_stats_launch(const float* returns, const float* active, const int32_t* rows, int n, double* stats,
                           cudaStream_t st) {
  int blocks = (n + 255) / 256;
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  minibatch_stats_kernel<<<blocks, 256, 0, st>>>(returns, active, rows, n, stats);
  return check_launch("minibatch_stats_kernel");
}

__global__ void valuenorm_update_kernel(float* __restrict__ vn, const double* __restrict__ stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double n = stats[3] > 0.0 ? stats[3] : 1.0;
    const float bm = (float)(stats[1] / n), bsq = (float)(stats[2] / n);
    const float w = 0.99999f, om = (float)(1.0 - 0.99999);     // the reference forms (1 - beta) in double
    vn[0] = vn[0] * w + bm * om;
    vn[1] = vn[1] * w + bsq * om;
    vn[2] = vn[2] * w + 1.0f * om;
  }
}

int valuenorm_update_launch(float* vn, const double* stats, cudaStream_t st) {
  valuenorm_update_kernel<<<1, 32, 0, st>>>(vn, stats);
  return check_launch("valuenorm_update_kernel");
}

// -------------------------------------------------------------------------------------------------
// a6 / a7: gathers
// -------------------------------------------------------------------------------------------------
// dst[p, :] = src[rows[p], :]; a warp walks one row with coalesced loads.  Bytes: 2 * 4 * n_rows * dim.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ rows, int n_rows, int dim,
                   float* __restrict__ dst) {
  const size_t total = (size_t)n_rows * dim;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / dim), c = (int)(i - (size_t)p * dim);
    dst[i] = __ldg(src + (size_t)rows[p] * dim + c);
  }
}

int gather_rows_launch(const float* src, const int32_t* rows, int n_rows, int dim, float* dst, cudaStream_t st) {
  const size_t total = (size_t)n_rows * dim;
  if (total == 0) return MAPPO_OK;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_rows_kernel<<<(int)blocks, 256, 0, st>>>(src, rows, n_rows, dim, dst);
  return check_launch("gather_rows_kernel");
}

// recurrent_generator row arithmetic (shared_buffer.py:505-569): (n,m,t)-ordered position j = chunk*L + l
// -> t = j % T, lane = j / T, storage row = t*E + lane.
__global__ void chunk_rows_kernel(const int32_t* __restrict__ chunks, int n_chunks, int L, int T, int E,
                                  int32_t* __restrict__ rows, int32_t* __restrict__ first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chunks * L) return;
  const int l = i / n_chunks, c = i - l * n_chunks;
  const long long j = (long long)chunks[c] * L + l;
  const int t = (int)(j % T), lane = (int)(j / T);
  const int row = t * E + lane;
  rows[i] = row;
  if (l == 0 && first) first[c] = row;
}

int chunk_rows_launch(const int32_t* chunks, int n_chunks, int L, int T, int E, int32_t* rows, int32_t* first,
                      cudaStream_t st) {
  const int n = n_chunks * L;
  if (n == 0) return MAPPO_OK;
  chunk_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(chunks, n_chunks, L, T, E, rows, first);
  return check_launch("chunk_rows_kernel");
}

// Keyed 4-round Feistel permutation on the smallest even-bit domain >= n, cycle-walked into [0, n).
__device__ __forceinline__ uint32_t mix32(uint32_t x, uint32_t k) {
  x ^= k; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// blockIdx.y = which permutation of a batch (independent keys), out + y*n
__global__ void randperm_kernel(int n, uint64_t seed, const uint64_t* __restrict__ counter, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  seed += 0x9E3779B97F4A7C15ull * (uint64_t)(blockIdx.y + 1) * (uint64_t)(gridDim.y > 1);
  out += (size_t)blockIdx.y * n;
  int bits = 2;
  while ((1u << bits) < (uint32_t)n) bits += 2;
  const int hb = bits / 2;
  const uint32_t hm = (1u << hb) - 1u;
  const uint64_t key = seed ^ ((counter ? *counter : 0ull) * 0x9E3779B97F4A7C15ull);
  uint32_t x = (uint32_t)i;
  do {
    uint32_t l = x >> hb, r = x & hm;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint32_t f = mix32(r, (uint32_t)(key >> (16 * round)) + 0x9E3779B9u * (uint32_t)round) & hm;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = (l << hb) | r;
  } while (x >= (uint32_t)n);
  out[i] = (int32_t)x;
}

int randperm_launch(int n, uint64_t seed, const uint64_t* counter, int32_t* out, cudaStream_t st, int n_perms) {
  if (n <= 0 || n_perms <= 0) return MAPPO_OK;
  randperm_kernel<<<dim3((n + 255) / 256, n_perms), 256, 0, st>>>(n, seed, counter, out);
  return check_launch("randperm_kernel");
}

// -------------------------------------------------------------------------------------------------
// a2: insert of one env step (shared_buffer.py:90-123 + mpe_runner.py:125-139)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) env_insert_kernel(const InsertArgs a) {
  const int seg0 = a.next_obs ? a.E * a.Do : 0;
  const int seg1 = seg0 + (a.next_share ? a.E * a.Ds : 0);
  const int seg2 = seg1 + (a.next_avail ? a.E * a.A : 0);
  const int seg3 = seg2 + ((a.ha && a.dones) ? a.E * a.H : 0);
  const int seg4 = seg3 + ((a.hc && a.dones) ? a.E * a.H : 0);
  const int seg5 = seg4 + a.E;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < seg5; i += gridDim.x * blockDim.x) {
    if (i < seg0) a.obs[i] = a.next_obs[i];
    else if (i < seg1) a.share[i - seg0] = a.next_share[i - seg0];
    else if (i < seg2) a.avail[i - seg1] = a.next_avail[i - seg1];
    else if (i < seg3) { const int j = i - seg2; if (a.dones[j / a.H] != 0.f) a.ha[j] = 0.f; }
    else if (i < seg4) { const int j = i - seg3; if (a.dones[j / a.H] != 0.f) a.hc[j] = 0.f; }
    else {
      const int e = i - seg4;
      if (a.rewards) a.rew[e] = a.rewards[e];
      if (a.dones && a.masks) a.masks[e] = a.dones[e] != 0.f ? 0.f : 1.f;
      if (a.next_active && a.active) a.active[e] = a.next_active[e];
    }
  }
  // optional: advance the sampling counter the preceding policy_step launch consumed (saves a launch per step)
  if (a.rng_counter && blockIdx.x == 0 && threadIdx.x == 0) *a.rng_counter += a.rng_inc;
}

int env_insert_launch(const InsertArgs& a, cudaStream_t st) {
  const int total = a.E * (a.Do + a.Ds + a.A + 2 * a.H + 1);
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  env_insert_kernel<<<blocks, 256, 0, st>>>(a);
  return check_launch("env_insert_kernel");
}

// -------------------------------------------------------------------------------------------------
// a13: slot reduction, clip_grad_norm_ and Adam.  Algorithmic bytes per optimiser step:
//   reduce: 4*n_slots*P read + 4*P written;  clip+Adam: 16*P read + 12*P written  (32 B/param)
// -------------------------------------------------------------------------------------------------
// Block = 32 parameters x 8 slot groups: warp sg sums slots sg, sg+8, ... (4 independent loads in flight) for 32
// consecutive parameters (one 128-byte line per slot), then the 8 partial sums are combined in a FIXED order through
// shared memory -> bit-reproducible, and ~8x more memory-level parallelism than one thread per parameter.
__global__ void __launch_bounds__(256)
grad_reduce_kernel(const float* __restrict__ part, int n_slots, int P, float* __restrict__ grad,
                   float* __restrict__ sumsq_part) {
  __shared__ float sacc[8][33];
  pdl_prologue();
  const int pi = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + pi;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
  if (i < P) {
    int s = sg;
    for (; s + 24 < n_slots; s += 32) {
      g0 += part[(size_t)s * P + i];
      g1 += part[(size_t)(s + 8) * P + i];
      g2 += part[(size_t)(s + 16) * P + i];
      g3 += part[(size_t)(s + 24) * P + i];
    }
    for (; s < n_slots; s += 8) g0 += part[(size_t)s * P + i];
  }
  sacc[sg][pi] = (g0 + g1) + (g2 + g3);
  __syncthreads();
  if (sg == 0) {
    float g = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) g += sacc[k][pi];
    if (i < P) grad[i] = g; else g = 0.f;
    float q = g * g;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (pi == 0 && sumsq_part) sumsq_part[blockIdx.x] = q;
  }
}

int grad_reduce_launch(const float* part, int n_slots, int P, float* grad, float* sumsq_part, int* n_blocks_out,
                       cudaStream_t st) {
  const int blocks = (P + 31) / 32;
  if (n_blocks_out) *n_blocks_out = blocks;
  launch_pdl(grad_reduce_kernel, dim3(blocks), dim3(256), 0, st, part, n_slots, P, grad, sumsq_part);
  return check_launch("grad_reduce_kernel");
}

// Per-block sums of squares of an (all-reduced) gradient vector: the multi-GPU caller re-derives the global
// norm after the collective.
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ grad, int P, float* __restrict__ sumsq_part) {
  __shared__ float sred[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float g = i < P ? grad[i] : 0.f;
  float q = g * g;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x < 32) {
    float x = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) sumsq_part[blockIdx.x] = x;
  }
}

int sumsq_launch(const float* grad, int P, float* sumsq_part, int* n_blocks_out, cudaStream_t st) {
  const int blocks = (P + 255) / 256;
  if (n_blocks_out) *n_blocks_out = blocks;
  sumsq_kernel<<<blocks, 256, 0, st>>>(grad, P, sumsq_part);
  return check_launch("sumsq_kernel");
}

// VEC = 1: one element per thread (the small nets: 6 - 33 K parameters, latency bound).  VEC = 4: four consecutive elements
// per thread as 128-bit loads / stores (the hidden >= 128 nets: ~0.9 M parameters, 28 B/param of HBM traffic).
template <int VEC>
__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                 int P, const float* __restrict__ sumsq_part, int n_part, const float* __restrict__ lr_dev,
                 int* __restrict__ step_dev, float eps, float max_norm, int use_clip, double* __restrict__ norm_out,
                 double* __restrict__ beta_pow) {
  __shared__ float s_total, s_coef, s_step_size, s_bc2_sqrt;
  __shared__ int s_step;
  __shared__ double s_p1, s_p2;
  pdl_prologue();
  // this thread's operands first: their latency overlaps the scalar prologue below instead of following it
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  float g_in[VEC], m_in[VEC], v_in[VEC], p_in[VEC];
  const bool full = i + VEC <= P;
  if (VEC == 4 && full) {
    const float4 a = *reinterpret_cast<const float4*>(grad + i), b = *reinterpret_cast<const float4*>(m + i);
    const float4 c = *reinterpret_cast<const float4*>(v + i), d = *reinterpret_cast<const float4*>(p + i);
    g_in[0] = a.x; g_in[VEC > 1 ? 1 : 0] = a.y; g_in[VEC > 2 ? 2 : 0] = a.z; g_in[VEC > 3 ? 3 : 0] = a.w;
    m_in[0] = b.x; m_in[VEC > 1 ? 1 : 0] = b.y; m_in[VEC > 2 ? 2 : 0] = b.z; m_in[VEC > 3 ? 3 : 0] = b.w;
    v_in[0] = c.x; v_in[VEC > 1 ? 1 : 0] = c.y; v_in[VEC > 2 ? 2 : 0] = c.z; v_in[VEC > 3 ? 3 : 0] = c.w;
    p_in[0] = d.x; p_in[VEC > 1 ? 1 : 0] = d.y; p_in[VEC > 2 ? 2 : 0] = d.z; p_in[VEC > 3 ? 3 : 0] = d.w;
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const bool ok = i + k < P;
      g_in[k] = ok ? grad[i + k] : 0.f; m_in[k] = ok ? m[i + k] : 0.f; v_in[k] = ok ? v[i + k] : 0.f; p_in[k] = ok ? p[i + k] : 0.f;
    }
  }
  // every block re-derives the global norm from the per-block partials (n_part is small) in a fixed order; the
  // scalar prologue (norm, clip coefficient, Adam bias corrections in fp64) runs in ONE warp and is broadcast
  if (threadIdx.x < 32) {
    double x = 0.0;
    for (int i = threadIdx.x; i < n_part; i += 32) x += (double)sumsq_part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) {
      const float tot = (float)sqrt(x);
      const int st = *step_dev + 1;                                      // 1-based Adam step
      // beta^t: carried from the previous step when the cache is current ({0.9^t, 0.999^t, t} as doubles; two fp64 pow()
      // calls cost ~1.5 us of single-thread latency on every launch), recomputed otherwise (first step, reloaded state)
      double p1, p2;
      if (beta_pow && beta_pow[2] == (double)(st - 1)) { p1 = beta_pow[0] * 0.9; p2 = beta_pow[1] * 0.999; }
      else { p1 = pow(0.9, (double)st); p2 = pow(0.999, (double)st); }
      s_p1 = p1; s_p2 = p2;
      const double bc1 = 1.0 - p1, bc2 = 1.0 - p2;
      s_total = tot;
      s_coef = use_clip ? fminf(max_norm / (tot + 1e-6f), 1.0f) : 1.f;   // clip_grad_norm_
      s_step_size = (float)((double)lr_dev[0] / bc1);
      s_bc2_sqrt = (float)sqrt(bc2);
      s_step = st;
    }
  }
  __syncthreads();
  const float total = s_total, coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const int step = s_step;
  float po[VEC], mo[VEC], vo[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float g = g_in[k] * coef;
    mo[k] = m_in[k] + (g - m_in[k]) * (float)(1.0 - 0.9);                     // torch: exp_avg.lerp_(grad, 1 - beta1)
    vo[k] = v_in[k] * 0.999f + (float)(1.0 - 0.999) * g * g;                  // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    const float denom = sqrtf(vo[k]) / bc2_sqrt + eps;
    po[k] = p_in[k] - step_size * (mo[k] / denom);
  }
  if (VEC == 4 && full) {
    *reinterpret_cast<float4*>(p + i) = make_float4(po[0], po[VEC > 1 ? 1 : 0], po[VEC > 2 ? 2 : 0], po[VEC > 3 ? 3 : 0]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mo[0], mo[VEC > 1 ? 1 : 0], mo[VEC > 2 ? 2 : 0], mo[VEC > 3 ? 3 : 0]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vo[0], vo[VEC > 1 ? 1 : 0], vo[VEC > 2 ? 2 : 0], vo[VEC > 3 ? 3 : 0]);
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (i + k < P) { p[i + k] = po[k]; m[i + k] = mo[k]; v[i + k] = vo[k]; }
  }
  // the last block to finish publishes the new step count (step_dev[1] is its ticket counter): every block has
  // read the old value by then, and no extra launch is needed
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int ticket = atomicAdd(step_dev + 1, 1);
    if (ticket == (int)gridDim.x - 1) {
      step_dev[0] = step;
      step_dev[1] = 0;
      if (beta_pow) { beta_pow[0] = s_p1; beta_pow[1] = s_p2; beta_pow[2] = (double)step; }
      if (norm_out) *norm_out += (double)total;
    }
  }
}

int clip_adam_launch(float* p, const float* grad, float* m, float* v, int P, const float* sumsq_part, int n_part,
                     const float* lr_dev, int* step_dev, float eps, float max_norm, int use_clip, double* norm_out,
                     double* beta_pow, cudaStream_t st) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (P >= 65536 && aligned)
    clip_adam_kernel<4><<<(P + 1023) / 1024, 256, 0, st>>>(p, grad, m, v, P, sumsq_part, n_part, lr_dev, step_dev, eps,
                                                           max_norm, use_clip, norm_out, beta_pow);
  else
    launch_pdl(clip_adam_kernel<1>, dim3((P + 255) / 256), dim3(256), 0, st, p, grad, m, v, P, sumsq_part, n_part, lr_dev, step_dev, eps,
               max_norm, use_clip, norm_out, beta_pow);
  return check_launch("clip_adam_kernel");
}

// dst = src with one partial sum of squares per 1024-element block (128-bit accesses): the single-slot case of the
// gradient reduction for the hidden >= 128 nets (their GEMM pipeline leaves one complete flat gradient)
__global__ void __launch_bounds__(256)
copy_sumsq4_kernel(const float* __restrict__ src, float* __restrict__ dst, int P, float* __restrict__ sumsq_part) {
  __shared__ float sred[8];
  const int i = (blockIdx.x * 256 + threadIdx.x) * 4;
  float q = 0.f;
  if (i + 4 <= P) {
    const float4 a = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<float4*>(dst + i) = a;
    q = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, a.w * a.w)));
  } else {
    for (int k = i; k < P; ++k) { const float a = src[k]; dst[k] = a; q = fmaf(a, a, q); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sred[k];
    if (sumsq_part) sumsq_part[blockIdx.x] = t;
  }
}

int copy_sumsq_launch(const float* src, float* dst, int P, float* sumsq_part, int* n_blocks_out, cudaStream_t st) {
  const int blocks = (P + 1023) / 1024;
  if (n_blocks_out) *n_blocks_out = blocks;
  copy_sumsq4_kernel<<<blocks, 256, 0, st>>>(src, dst, P, sumsq_part);
  return check_launch("copy_sumsq4_kernel");
}

}  // namespace mappo
