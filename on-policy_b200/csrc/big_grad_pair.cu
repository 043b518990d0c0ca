// big_grad_pair.cu -- weight-gradient GEMM on CTA pairs (cta_group::2):  G[m, q] = sum_{row in split} P[row, m] Q[row, q].
//
// Same contract as big_grad_kernel (big_gemm.cu) with a 256 x (up to 320) output tile per PAIR of CTAs: CTA r of the pair
// stages 128 of the 256 P columns and HALF of the Q columns of every 32-row k-step, the leader issues M = 256 MMAs for
// both SMs, each CTA ends up with its 128 output rows x all tile columns in its own TMEM.  Operand bytes per SM and k-step drop
// from 52 KB (128 x 288 tile) to 36 KB at a larger tile -- 45 -> 73 FLOP per operand byte -- which is what bounds the single-
// CTA kernel (L2 -> shared-memory traffic ~10 TB/s at 47 % tensor-pipe activity, profiles/r2b_ncu_big_grad_full.csv).
// Q tiles are multiples of 64 columns (half per CTA = whole 32-column groups): 256, or 256 + 64 = 320 for the last tile of
// an extended activation row (H + 64: the mean / sigma columns).
#include <cstring>
#include "big_tc.cuh"
#include "big_net.h"

namespace mappo {
namespace big {

constexpr int kGPThreads = 192;
constexpr int kGPKR = 32;                               // rows per stage
constexpr int kGPGroup = kGPKR * 128;                   // one 32-column group of a stage: 4 KB
constexpr int kGPStages = 6;
constexpr int kGPStageBytes = (4 + 5) * kGPGroup;       // per CTA: 4 P groups + up to 5 Q groups (160 of 320 columns)

__global__ void __launch_bounds__(kGPThreads, 1)
big_grad_pair_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ4, const __grid_constant__ CUtensorMap mapQa,
                     const __grid_constant__ CUtensorMap mapQb, float* __restrict__ partial, const GradShape sh) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full[kGPStages], empty[kGPStages], done;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int unit = (int)blockIdx.x >> 1;
  const int split = unit / (sh.m_tiles * sh.n_tiles), rem = unit % (sh.m_tiles * sh.n_tiles);
  const int mt = rem / sh.n_tiles, nt = rem % sh.n_tiles;                 // mt counts 256-column tiles of P here
  const int q0 = sh.q0[nt], qw = sh.qw[nt];
  const int n1 = qw > 256 ? 256 : qw, n2 = qw - n1;                       // two MMAs per k-step: N = n1 and N = n2 (0 or 64)
  const int g1 = n1 / 64, g2 = n2 / 64;                                   // 32-column groups of each part held by ONE CTA
  const int r0 = split * sh.rows_per_split, r1 = min(sh.rows, r0 + sh.rows_per_split);
  const int n_kb = r1 > r0 ? (r1 - r0 + kGPKR - 1) / kGPKR : 0;
  if (tid == 0) {
    for (int i = 0; i < kGPStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(&done, 1);
    mbar_fence_init();
    tma_prefetch_desc(&mapP); tma_prefetch_desc(&mapQ4); tma_prefetch_desc(&mapQa); tma_prefetch_desc(&mapQb);
  }
  if (warp == 1) tmem_alloc_pair(&tmem_slot, 512);
  tc_fence_before();
  cluster_sync();
  __syncthreads();             // (the cluster barrier already orders the allocator's write of tmem_slot; racecheck only models bar.sync)
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t stage_tx = (uint32_t)((4 + g1 + g2) * kGPGroup);         // bytes ONE CTA loads per stage
  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(empty + stage, phase ^ 1);
      uint8_t* sP = smem + stage * kGPStageBytes;
      uint8_t* sQ = sP + 4 * kGPGroup;
      if (rank == 0) mbar_expect_tx(full + stage, 2u * stage_tx);         // the pair's bytes land on the leader's barrier
      const uint32_t lb = leader_addr(full + stage);
      const int row = r0 + kb * kGPKR;
      // one 3-D box per operand part ([groups][32 rows][32 columns]): this CTA's 4 groups of P, its half of the N = n1 part of Q
      // (4 groups through mapQ4, fewer through mapQa) and its half of the N = 64 part (mapQb, one group)
      tma_load_3d_pair(sP, &mapP, 0, row, mt * 8 + (int)rank * 4, lb);
      tma_load_3d_pair(sQ, g1 == 4 ? &mapQ4 : &mapQa, 0, row, q0 / 32 + (int)rank * g1, lb);
      if (g2 > 0) tma_load_3d_pair(sQ + g1 * kGPGroup, &mapQb, 0, row, (q0 + n1) / 32 + (int)rank * g2, lb);
      if (++stage == kGPStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t id1 = make_idesc(256, n1, 1, 1), id2 = make_idesc(256, n2 > 0 ? n2 : 64, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(full + stage, phase);
      tc_fence_after();
      const uint32_t p0 = smem_u32(smem + stage * kGPStageBytes), qb = p0 + 4 * kGPGroup;
#pragma unroll
      for (int k = 0; k < kGPKR / 8; ++k) {
        const uint64_t pd = make_desc(p0 + k * 1024, kGPGroup, 512, 1);
        umma_tf32_pair(tmem, pd, make_desc(qb + k * 1024, kGPGroup, 512, 1), id1, (kb | k) ? 1u : 0u);
        if (n2 > 0)
          umma_tf32_pair(tmem + 256, pd, make_desc(qb + g1 * kGPGroup + k * 1024, kGPGroup, 512, 1), id2, (kb | k) ? 1u : 0u);
      }
      umma_commit_pair(empty + stage);
      if (++stage == kGPStages) { stage = 0; phase ^= 1; }
    }
    umma_commit_pair(&done);
  } else if (warp >= 2) {
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const int m = mt * 256 + (int)rank * 128 + r;
    if (n_kb > 0) { mbar_wait(&done, 0); tc_fence_after(); }
    float* dst = partial + ((size_t)split * sh.M + (size_t)min(m, sh.M - 1)) * sh.ldq + q0;
    for (int c = 0; c < qw / 32; ++c) {
      float v[32];
      if (n_kb > 0) { tmem_ld32(tmem + lane_base + (uint32_t)(c * 32), v); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (m < sh.M) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<float4*>(dst + c * 32 + 4 * ch) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
      }
    }
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 1) tmem_dealloc_pair(tmem, 512);
}

int make_map3(CUtensorMap* m, const float* base, long long width, long long rows, long long ld, int box_rows, int box_groups);

// sh: m_tiles counts 256-column tiles of P; every qw is a multiple of 64 (<= 320; only the last tile may differ from 256)
int grad_gemm_pair_launch(const float* P, int ldp, const float* Q, int ldq_in, float* partial, GradShape sh, cudaStream_t st) {
  for (int i = 0; i < sh.n_tiles; ++i)
    if (sh.qw[i] % 64 || sh.qw[i] > 320 || (i + 1 < sh.n_tiles && sh.qw[i] != 256)) { set_error("grad_gemm_pair: tile of %d columns", sh.qw[i]); return MAPPO_ERR_INVALID; }
  const int lw = sh.qw[sh.n_tiles - 1], l1 = (lw > 256 ? 256 : lw) / 64;              // groups per CTA of the last tile's first part
  CUtensorMap mP, mQ4, mQa, mQb;
  int rc = make_map3(&mP, P, sh.Pw, sh.rows, ldp, kGPKR, 4);
  if (rc) return rc;
  rc = make_map3(&mQ4, Q, sh.Qw, sh.rows, ldq_in, kGPKR, 4);
  if (rc) return rc;
  rc = make_map3(&mQa, Q, sh.Qw, sh.rows, ldq_in, kGPKR, l1);
  if (rc) return rc;
  rc = make_map3(&mQb, Q, sh.Qw, sh.rows, ldq_in, kGPKR, 1);
  if (rc) return rc;
  const size_t bytes = (size_t)kGPStages * kGPStageBytes + 1024;
  if (cudaFuncSetAttribute(big_grad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return check_launch("big_grad_pair_kernel: cudaFuncSetAttribute");
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * sh.splits * sh.m_tiles * sh.n_tiles);
  cfg.blockDim = dim3(kGPThreads);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, big_grad_pair_kernel, mP, mQ4, mQa, mQb, partial, sh);
  return check_launch("big_grad_pair_kernel");
}

}  // namespace big
}  // namespace mappo
