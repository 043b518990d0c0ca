// p2p_allreduce.cu -- one-shot all-reduce over NVLink / NVSwitch peer memory, as ONE graph-capturable kernel.
//
// Every rank owns a "symmetric" buffer (same layout on every GPU, peer-mapped into every process, e.g. through
// torch.distributed._symmetric_memory) and a tiny signal pad uint32[world].  One launch per rank:
//   1. arrive : rank r stores the round number into slot r of EVERY peer's signal pad (st.release.sys)
//   2. wait   : spin until all world slots of my own pad have reached the round (ld.acquire.sys)
//   3. reduce : out[i] = sum_{p = 0..world-1} peer_p[offset + i], read straight through NVLink (L1-bypassing loads);
//               fixed order, so every rank ends up with bit-identical sums and the replicas never drift
//   4. the last CTA publishes the completed round in device memory (read by the next launch; nothing on the host).
// round_dev: uint32[4] = {completed round, CTA ticket, error flag (0 = ok, 1 + peer whose arrival timed out), unused}.
// Reuse rule: a region may be rewritten as soon as a LATER round on any region has completed locally (each round is
// a full barrier) -- the gradient all-reduce therefore alternates between two halves of the buffer.
// The reference has no collective (single process); this is the multi-GPU exchange of SURVEY section 8e, C1.
#include "p2p.cuh"

namespace mappo {

template <typename T>
__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(const P2PArgs a, long long offset_bytes, int n, T* __restrict__ out, uint32_t* __restrict__ round_dev,
                     float* __restrict__ sumsq_part) {
  __shared__ uint32_t s_round;
  __shared__ float s_sq[8];
  float sq = 0.f;
  if (threadIdx.x == 0) s_round = round_dev[0] + 1;
  __syncthreads();
  const uint32_t round = s_round;
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();                           // my earlier writes to the symmetric buffer are visible to peers
    st_release_sys(a.sig[threadIdx.x] + a.rank, round);
  }
  if (threadIdx.x < a.world) {
    // bounded wait (~4 s at 2 GHz): a dead peer, or two reducers whose kernels cannot be co-resident on some rank (a
    // serialising tool), raise round_dev[2] instead of hanging the box; the caller reads it with mappo_p2p_error().
    const uint32_t* mine = a.sig[a.rank] + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - round) < 0) {
      if (clock64() - t0 > 8000000000LL) { atomicExch(round_dev + 2, 1u + (uint32_t)threadIdx.x); break; }
    }
  }
  __syncthreads();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (sizeof(T) == 4 && (n & 3) == 0 && (offset_bytes & 15) == 0) {
    for (int i = tid; i < n / 4; i += nt) {
      float4 v[kMaxPeers];
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)               // every peer's load in flight at once: ONE NVLink round trip
        if (p < a.world)
          v[p] = ld_peer4(reinterpret_cast<const float*>(static_cast<const char*>(a.buf[p]) + offset_bytes) + 4 * i);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)               // summed in rank order: bit-identical on every rank
        if (p < a.world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
      reinterpret_cast<float4*>(out)[i] = s;
      sq = fmaf(s.x, s.x, fmaf(s.y, s.y, fmaf(s.z, s.z, fmaf(s.w, s.w, sq))));
    }
  } else {
    for (int i = tid; i < n; i += nt) {
      T v[kMaxPeers];
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < a.world) v[p] = ld_peer<T>(reinterpret_cast<const T*>(static_cast<const char*>(a.buf[p]) + offset_bytes) + i);
      T s = (T)0;
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < a.world) s += v[p];
      out[i] = s;
      sq = fmaf((float)s, (float)s, sq);
    }
  }
  if (sumsq_part) {                                     // per-CTA sum of squares of the reduced vector (for clip_grad_norm_)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) s_sq[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += s_sq[q];
      sumsq_part[blockIdx.x] = t;
    }
  }
  // completion: the last CTA to finish publishes the round (round_dev[1] is its ticket counter)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(round_dev + 1, 1u);
    if (t == gridDim.x - 1) { round_dev[0] = round; round_dev[1] = 0; }
  }
}

template <typename T>
static int launch_p2p(const void* const* bufs, void* const* sigs, int world, int rank, long long offset_bytes, int n, T* out,
                      uint32_t* round_dev, float* sumsq_part, int* n_blocks_out, cudaStream_t st) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) { set_error("p2p_allreduce: world %d / rank %d", world, rank); return MAPPO_ERR_INVALID; }
  if (!bufs || !sigs || !out || !round_dev || n <= 0) { set_error("p2p_allreduce: NULL / empty"); return MAPPO_ERR_INVALID; }
  P2PArgs a;
  for (int p = 0; p < kMaxPeers; ++p) { a.buf[p] = p < world ? bufs[p] : nullptr; a.sig[p] = p < world ? static_cast<uint32_t*>(sigs[p]) : nullptr; }
  a.world = world; a.rank = rank;
  int blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  if (n_blocks_out) *n_blocks_out = blocks;
  p2p_allreduce_kernel<T><<<blocks, 256, 0, st>>>(a, offset_bytes, n, out, round_dev, sumsq_part);
  return check_launch("p2p_allreduce_kernel");
}

int p2p_allreduce_f32_launch(const void* const* bufs, void* const* sigs, int world, int rank, long long off, int n, float* out,
                             uint32_t* round_dev, float* sumsq_part, int* n_blocks_out, cudaStream_t st) {
  return launch_p2p<float>(bufs, sigs, world, rank, off, n, out, round_dev, sumsq_part, n_blocks_out, st);
}
int p2p_allreduce_f64_launch(const void* const* bufs, void* const* sigs, int world, int rank, long long off, int n, double* out,
                             uint32_t* round_dev, cudaStream_t st) {
  return launch_p2p<double>(bufs, sigs, world, rank, off, n, out, round_dev, nullptr, nullptr, st);
}

}  // namespace mappo
