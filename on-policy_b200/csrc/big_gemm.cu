// big_gemm.cu -- the tensor-core kernels of the hidden >= 128 MLP path: TMA-fed, warp-specialised tcgen05 GEMMs.
//
//   big_lin_kernel<Epi>   out = epilogue(A[rows, K] . W[N, K]^T)   K-major operands (activations x weights), used for
//                         forward layers (EpiFwd), the head + losses (EpiHead / EpiSample) and input gradients (EpiBwd)
//   big_grad_kernel       G[m, q] = sum_rows P[row, m] Q[row, q]   MN-major operands (contraction over the rows of two
//                         row-major matrices): every weight gradient
//
// big_lin_kernel: persistent CTAs (one per SM), 6 warps:
//   warp 0   TMA producer: A box {32 K x 128 rows} + W box {32 K x BN rows} per stage, SWIZZLE_128B, 3-stage mbarrier ring
//   warp 1   MMA issuer (one lane): 4 x tcgen05.mma kind::tf32 (M 128, N BN, K 8) per stage into one of TWO TMEM accumulator
//            stages (2 x BN columns), tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5  epilogue: thread = row (TMEM lane), 32 columns per step through an Epi functor (big_epi.cuh), results staged
//            in a 128B-swizzled [128 x 32] tile and written with TMA stores; EpiBwd also TMA-loads the stored activation
//            tile of the same coordinates into the staging slot first (in-place transform)
// The two N tiles of a 512-wide row go to the SAME CTA back to back, so the row statistics stay in registers, and the MMA
// of tile i + 1 overlaps the epilogue of tile i (double-buffered accumulators).
#include <cstdlib>
#include <cstring>
#include "big_tc.cuh"
#include "big_epi.cuh"
#include "big_net.h"

namespace mappo {
namespace big {

constexpr int kLinThreads = 192;
constexpr int kNS = 4;                    // staging slots (16 KB each)
constexpr int kAinDepth = 2;              // activation tiles in flight ahead of the epilogue (EpiBwd)
constexpr int kABytes = 128 * 128;        // A stage: 128 rows x 32 tf32

struct LinSmem { int stage_bytes, stages_off, slots_off, colvec_off, scratch_off, rowpart_off, bars_off, total; };
constexpr int kPairGroups = 1;            // epilogue groups of the pair kernel (2 = one per accumulator stage; measured: no gain, the
                                          // kernel is bound by operand latency, not by the epilogue -- profiles/r2_summary.md)
constexpr int kNSPair = 3;                // staging slots of each epilogue group of the pair kernel
constexpr int kPairStages = 5;            // 5 x 32 KB of operands in flight per CTA (the k-step rate is latency / stages)
// pair: each CTA of a cta_group::2 pair stages its 128 rows of A and HALF of the B tile (BN / 2 rows of W)
__host__ __device__ inline LinSmem make_lin_smem(int BN, int n_stages, int N_cv, bool scratch, bool pair = false) {
  LinSmem s;
  s.stage_bytes = kABytes + (pair ? BN / 2 : BN) * 128;
  s.stages_off = 0;
  s.slots_off = n_stages * s.stage_bytes;
  s.colvec_off = s.slots_off + (pair ? kPairGroups * kNSPair : kNS) * 16384;
  s.scratch_off = s.colvec_off + ((2 * N_cv * 4 + 127) & ~127);
  s.rowpart_off = s.scratch_off + (scratch ? 32 * kLgLd * 4 : 0);
  s.bars_off = s.rowpart_off + (pair ? 3 * 128 * 2 * 4 : 0);
  s.total = s.bars_off + 256;
  return s;
}

// PAIR: two epilogue groups of four warps (warps 2-5 and 6-9): group q drains accumulator stage q, i.e. every other tile, so every
// SM sub-partition has two epilogue warps to interleave (a single warp per scheduler left the epilogue latency bound: 0.19 IPC,
// tensor pipe 47 % busy -- profiles/r2b_ncu_big_lin_full.csv).  The two N tiles of a 512-wide row then belong to different groups:
// group 0 hands its partial row sums to group 1 through shared memory (rowpart, triple buffered) and group 1 closes the row.
template <class Epi, bool PAIR>
__global__ void __launch_bounds__(PAIR ? 64 + 128 * kPairGroups : kLinThreads, 1)
big_lin_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapAin,
               const typename Epi::Args ea, const LinShape sh) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ double sred[2 * 32];
  // 1024-byte alignment by pointer arithmetic on the shared array (an integer round trip would make every later access a
  // generic LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = sh.BN, NT = sh.N / sh.BN, KB = (sh.K + 31) / 32, NST = sh.n_stages;
  const LinSmem L = make_lin_smem(BN, NST, sh.N, Epi::kNeedsScratch, PAIR);
  // PAIR: CTAs 2 p and 2 p + 1 form a cta_group::2 pair on one 256-row block; rank 0 (the leader) issues the MMAs for both
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  const int unit = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x, n_units = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
  const int n_ublocks = PAIR ? (sh.n_rowblocks + 1) >> 1 : sh.n_rowblocks;        // row blocks of a unit (256 or 128 rows)
  float* cv = reinterpret_cast<float*>(smem + L.colvec_off);
  float* scratch = reinterpret_cast<float*>(smem + L.scratch_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars_off);
  uint64_t* full = bars;                       // [NST <= 6]
  uint64_t* empty = bars + 6;                  // [NST <= 6]
  uint64_t* tfull = bars + 12;                 // [2]
  uint64_t* tempty = bars + 14;                // [2]
  uint64_t* ainfull = bars + 16;               // [2 groups][4]
  uint64_t* rowbar = bars + 24;                // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  float* rowpart = reinterpret_cast<float*>(smem + L.rowpart_off);
  const int n_threads = PAIR ? 64 + 128 * kPairGroups : kLinThreads;
  const uint32_t tmem_cols = (2 * BN <= 32) ? 32u : (2 * BN <= 64 ? 64u : (2 * BN <= 128 ? 128u : (2 * BN <= 256 ? 256u : 512u)));

  for (int i = tid; i < 2 * sh.N; i += n_threads) cv[i] = ea.colvec[i];
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, PAIR ? 8 : 128); }     // pair: 4 warps x 2 CTAs
    for (int i = 0; i < 8; ++i) mbar_init(ainfull + i, 1);
    for (int i = 0; i < 3; ++i) mbar_init(rowbar + i, 128);
    mbar_fence_init();
    tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB);
    if (Epi::kStoresOut) tma_prefetch_desc(&mapOut);
    if (Epi::kHasAin) tma_prefetch_desc(&mapAin);
  }
  if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, tmem_cols); else tmem_alloc(tmem_slot, tmem_cols); }
  tc_fence_before();
  if (PAIR) cluster_sync();                                      // (pair: the peer's barriers are initialised before anything signals them)
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  typename Epi::Thread th;
  Epi::init_thread(th);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ub = unit; ub < n_ublocks; ub += n_units) {
        const int rb = PAIR ? 2 * ub + (int)rank : ub;
        for (int nt = 0; nt < NT; ++nt)
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(empty + stage, phase ^ 1);
            uint8_t* sA = smem + L.stages_off + stage * L.stage_bytes;
            if (PAIR) {
              // both CTAs' bytes are accounted on the LEADER's barrier (armed by the leader for the whole pair)
              if (rank == 0) mbar_expect_tx(full + stage, 2u * (uint32_t)L.stage_bytes);
              const uint32_t lb = leader_addr(full + stage);
              tma_load_2d_pair(sA, &mapA, kb * 32, rb * 128, lb);
              tma_load_2d_pair(sA + kABytes, &mapB, kb * 32, nt * BN + (int)rank * (BN / 2), lb);
            } else {
              mbar_expect_tx(full + stage, (uint32_t)L.stage_bytes);
              tma_load_2d(sA, &mapA, kb * 32, rb * 128, full + stage);
              tma_load_2d(sA + kABytes, &mapB, kb * 32, nt * BN, full + stage);
            }
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: the leader only) =====================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(PAIR ? 256 : 128, BN, 0, 0);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int ub = unit; ub < n_ublocks; ub += n_units)
        for (int nt = 0; nt < NT; ++nt) {
          mbar_wait(tempty + as, aphase ^ 1);             // the epilogue(s) have drained this accumulator stage
          tc_fence_after();
          const uint32_t d = tmem + (uint32_t)(as * BN);
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            const uint32_t a0 = smem_u32(smem + L.stages_off + stage * L.stage_bytes), b0 = a0 + kABytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_desc(a0 + k * 32, 16, 1024, 2), bd = make_desc(b0 + k * 32, 16, 1024, 2);
              if (PAIR) umma_tf32_pair(d, ad, bd, idesc, (kb | k) ? 1u : 0u);
              else umma_tf32(d, ad, bd, idesc, (kb | k) ? 1u : 0u);
            }
            if (PAIR) umma_commit_pair(empty + stage); else umma_commit(empty + stage);     // frees the smem stage (in both CTAs)
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
          if (PAIR) umma_commit_pair(tfull + as); else umma_commit(tfull + as);             // accumulator complete
          as ^= 1;
          if (as == 0) aphase ^= 1;
        }
    }
  } else {
    // ===================== epilogue (one or two groups of 128 threads) =====================
    constexpr int NG = PAIR ? kPairGroups : 1;
    constexpr int NSg = PAIR ? kNSPair : kNS;
    const int grp = NG > 1 ? (warp - 2) >> 2 : 0;
    const int et = (tid - 64) & 127;                      // 0..127 within the group
    const int r = (warp & 3) * 32 + lane;                 // TMEM lane = row of the tile (a warp may only touch lanes 32 (warp % 4)..)
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    uint8_t* slots = smem + L.slots_off + grp * NSg * 16384;
    uint64_t* ainf = ainfull + grp * 4;
    const int barid = 1 + grp;
    const int CPT = BN / kChunk;                          // chunks per tile
    const int my_rbs = unit < n_ublocks ? (n_ublocks - unit + n_units - 1) / n_units : 0;
    const int my_tiles = my_rbs * NT;                     // tiles of this CTA; group q takes tiles q, q + NG, ...
    const int my_gtiles = my_tiles > grp ? (my_tiles - grp + NG - 1) / NG : 0;
    const long long total_chunks = (long long)my_gtiles * CPT;
    auto row_of = [&](int rbl) { const int ubl = unit + rbl * n_units; return (PAIR ? 2 * ubl + (int)rank : ubl) * 128; };
    auto issue_ain = [&](long long g) {                   // thread et == 0 of the group only; g = chunk index within the group
      if (g >= total_chunks) return;
      const int tile = (int)(g / CPT) * NG + grp, c = (int)(g % CPT);
      const int slot = (int)(g % NSg);
      mbar_expect_tx(ainf + slot, 16384u);
      tma_load_2d(slots + slot * 16384, &mapAin, (tile % NT) * BN + c * kChunk, row_of(tile / NT), ainf + slot);
    };
    const bool direct = sh.direct != 0;
    if (Epi::kHasAin && et == 0 && !direct)
      for (int i = 0; i < kAinDepth; ++i) issue_ain(i);
    long long g = 0;                                      // running chunk index of this group
    typename Epi::Row row;
    int cur_rbl = -1, rb128 = 0, grow = 0;
    for (int tile = grp; tile < my_tiles; tile += NG) {
      const int rbl = tile / NT, nt = tile % NT;
      const int as = tile & 1;
      const uint32_t aphase = (uint32_t)((tile >> 1) & 1);
      if (rbl != cur_rbl) {
        cur_rbl = rbl;
        rb128 = row_of(rbl);
        grow = rb128 + r;
        Epi::begin_row(ea, row, grow);
      }
      mbar_wait(tfull + as, aphase);
      tc_fence_after();
      for (int c = 0; c < CPT; ++c, ++g) {
        const int col0 = nt * BN + c * kChunk;
        float acc[kChunk], ain[kChunk], out[kChunk];
        tmem_ld32(tmem + lane_base + (uint32_t)(as * BN + c * kChunk), acc);
        tmem_ld_wait();
        if (c == CPT - 1) {                               // last read of this accumulator stage: hand it back to the MMA warp
          tc_fence_before();
          if (PAIR) {
            // one remote arrive per WARP: a release at cluster scope costs a full memory barrier (ncu: MEMBAR.ALL.GPU + ERRBAR
            // were the hottest epilogue instructions with one arrive per thread)
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty + as);
          } else {
            mbar_arrive(tempty + as);
          }
        }
        const int slot = (int)(g % NSg);
        uint8_t* sl = slots + slot * 16384;
        if (direct) {
          // the staging tile (16 KB written + 16 KB read per chunk, twice with the saved activation) shares the shared-memory port with
          // the tensor core's operand reads and the TMA operand writes; here every thread moves its own 128-byte row segment instead
          if (Epi::kHasAin) {
            if (grow < sh.n_rows) {
              const float4* ap = reinterpret_cast<const float4*>(sh.ain_ptr + (size_t)grow * sh.ain_ld + col0);
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                const float4 v = __ldg(ap + ch);
                ain[4 * ch] = v.x; ain[4 * ch + 1] = v.y; ain[4 * ch + 2] = v.z; ain[4 * ch + 3] = v.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < kChunk; ++j) ain[j] = 0.f;
            }
          }
          Epi::chunk(ea, th, row, acc, ain, out, col0, cv, scratch, r, grow);
          if (Epi::kStoresOut && sh.store_out && grow < sh.n_rows) {
            float4* op = reinterpret_cast<float4*>(sh.out_ptr + (size_t)grow * sh.out_ld + col0);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) op[ch] = make_float4(out[4 * ch], out[4 * ch + 1], out[4 * ch + 2], out[4 * ch + 3]);
          }
          continue;
        }
        if (Epi::kHasAin) {
          mbar_wait(ainf + slot, (uint32_t)((g / NSg) & 1));
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const float4 v = *reinterpret_cast<const float4*>(sl + sw128_off(r, ch));
            ain[4 * ch] = v.x; ain[4 * ch + 1] = v.y; ain[4 * ch + 2] = v.z; ain[4 * ch + 3] = v.w;
          }
        } else if (Epi::kStoresOut && sh.store_out) {
          if (et == 0) tma_store_wait_read<NSg - 1>();    // the store that last used this slot has read it
          named_bar_sync(barid, 128);
        }
        Epi::chunk(ea, th, row, acc, ain, out, col0, cv, scratch, r, grow);
        if (Epi::kStoresOut && sh.store_out) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<float4*>(sl + sw128_off(r, ch)) = make_float4(out[4 * ch], out[4 * ch + 1], out[4 * ch + 2], out[4 * ch + 3]);
          fence_async_smem();
          named_bar_sync(barid, 128);
          if (et == 0) {
            tma_store_2d(&mapOut, col0, rb128, sl);
            tma_store_commit();
            if (Epi::kHasAin) {                            // refill: the slot of chunk g + depth was last stored by chunk g + depth - NSg
              tma_store_wait_read<NSg - kAinDepth>();
              issue_ain(g + kAinDepth);
            }
          }
        } else if (Epi::kHasAin) {
          named_bar_sync(barid, 128);
          if (et == 0) issue_ain(g + kAinDepth);
        }
      }
      if (nt + NG >= NT) {                                 // this group's last tile of the row
        if (NG == 2 && NT > 1) {                           // (NT even: group 0 owns the even tiles, group 1 the odd ones and the end of the row)
          float* rp = rowpart + ((rbl % 3) * 128 + r) * 2;
          if (grp == 0) { Epi::get_part(row, rp); mbar_arrive(rowbar + rbl % 3); }
          else { mbar_wait(rowbar + rbl % 3, (uint32_t)((rbl / 3) & 1)); Epi::add_part(row, rp); Epi::end_row(ea, row, grow); }
        } else {
          Epi::end_row(ea, row, grow);
        }
      }
    }
    if (et == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  if (PAIR) cluster_sync(); else __syncthreads();                // (pair: nobody leaves while the peer may still signal its barriers / read its tiles)
  Epi::finish_thread(ea, th, sred, tid, n_threads);
  if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem, tmem_cols); else tmem_dealloc(tmem, tmem_cols); }
}

// ------------------------------------------------------------------------------------------------------------
// weight-gradient GEMM: G[m, q] = sum_{row in split} P[row, m] Q[row, q]; one output tile (128 m x up to 288 q) per CTA,
// the rows of the batch split over `splits` CTAs per tile; partial[split][m][q] written with plain stores.
// Operands are MN-major tf32: TMA boxes {32 columns x 32 rows} with the 128B-swizzle / 32B-atom mode land as
// [column group][row][32] and are consumed with UMMA layout type SWIZZLE_128B_BASE32B (SBO = 4 rows, LBO = one group).
// ------------------------------------------------------------------------------------------------------------
constexpr int kGradKR = 32;               // rows per stage
constexpr int kGradGroupBytes = kGradKR * 128;
constexpr int kGradStages = 4;
constexpr int kGradStageBytes = (4 + 10) * kGradGroupBytes;     // P: 4 groups, Q: up to 10 groups (320 columns)

__global__ void __launch_bounds__(kLinThreads, 1)
big_grad_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ8, const __grid_constant__ CUtensorMap mapQr,
                float* __restrict__ partial, const GradShape sh) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full[kGradStages], empty[kGradStages], done;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = blockIdx.x;
  const int split = unit / (sh.m_tiles * sh.n_tiles), rem = unit % (sh.m_tiles * sh.n_tiles);
  const int mt = rem / sh.n_tiles, nt = rem % sh.n_tiles;
  const int q0 = sh.q0[nt], qw = sh.qw[nt], qgroups = qw / 32;
  const int r0 = split * sh.rows_per_split, r1 = min(sh.rows, r0 + sh.rows_per_split);
  const int n_kb = r1 > r0 ? (r1 - r0 + kGradKR - 1) / kGradKR : 0;
  if (tid == 0) {
    for (int i = 0; i < kGradStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(&done, 1);
    mbar_fence_init();
    tma_prefetch_desc(&mapP); tma_prefetch_desc(&mapQ8); tma_prefetch_desc(&mapQr);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(empty + stage, phase ^ 1);
      uint8_t* sP = smem + stage * kGradStageBytes;
      uint8_t* sQ = sP + 4 * kGradGroupBytes;
      mbar_expect_tx(full + stage, (uint32_t)((4 + qgroups) * kGradGroupBytes));
      const int row = r0 + kb * kGradKR;
      // rows beyond r1 belong to the next split: they must not be counted twice -> the row coordinate is clamped by
      // rows_per_split being a multiple of kGradKR (host), only the LAST split can run past `rows` (TMA zero fill)
      // one 3-D box per operand part: [groups][32 rows][32 columns]
      tma_load_3d(sP, &mapP, 0, row, mt * 4, full + stage);
      const int g8 = qgroups < 8 ? 0 : 8;                 // a whole 256-column part through the 8-group map
      if (g8) tma_load_3d(sQ, &mapQ8, 0, row, q0 / 32, full + stage);
      if (qgroups - g8 > 0) tma_load_3d(sQ + g8 * kGradGroupBytes, &mapQr, 0, row, q0 / 32 + g8, full + stage);
      if (++stage == kGradStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    const int n1 = qw > 256 ? 256 : qw, n2 = qw - n1;
    const uint32_t id1 = make_idesc(128, n1, 1, 1), id2 = make_idesc(128, n2 > 0 ? n2 : 32, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(full + stage, phase);
      tc_fence_after();
      const uint32_t p0 = smem_u32(smem + stage * kGradStageBytes), qb = p0 + 4 * kGradGroupBytes;
#pragma unroll
      for (int k = 0; k < kGradKR / 8; ++k) {
        const uint64_t pd = make_desc(p0 + k * 1024, kGradGroupBytes, 512, 1);
        umma_tf32(tmem, pd, make_desc(qb + k * 1024, kGradGroupBytes, 512, 1), id1, (kb | k) ? 1u : 0u);
        if (n2 > 0)
          umma_tf32(tmem + 256, pd, make_desc(qb + 8 * kGradGroupBytes + k * 1024, kGradGroupBytes, 512, 1), id2, (kb | k) ? 1u : 0u);
      }
      umma_commit(empty + stage);
      if (++stage == kGradStages) { stage = 0; phase ^= 1; }
    }
    umma_commit(&done);
  } else if (warp >= 2) {
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const int m = mt * 128 + r;
    if (n_kb > 0) { mbar_wait(&done, 0); tc_fence_after(); }
    float* dst = partial + ((size_t)split * sh.M + (size_t)min(m, sh.M - 1)) * sh.ldq + q0;
    for (int c = 0; c < qgroups; ++c) {
      float v[kChunk];
      if (n_kb > 0) { tmem_ld32(tmem + lane_base + (uint32_t)(c * kChunk), v); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] = 0.f;
      }
      if (m < sh.M) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<float4*>(dst + c * kChunk + 4 * ch) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------
// host side: tensor maps + launches
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int ensure_encode() {
  if (g_encode) return MAPPO_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    set_error("big net path: cuTensorMapEncodeTiled is not available from the driver");
    return MAPPO_ERR_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return MAPPO_OK;
}

// 2-D fp32 map over a row-major matrix [rows][width] with leading dimension ld (floats), box {box_w, box_h}
int make_map(CUtensorMap* m, const float* base, long long width, long long rows, long long ld, int box_w, int box_h, int swizzle) {
  int rc = ensure_encode();
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base %p width %lld rows %lld ld %lld box %dx%d swizzle %d", (int)r, (const void*)base,
              width, rows, ld, box_w, box_h, swizzle);
    return MAPPO_ERR_CUDA;
  }
  return MAPPO_OK;
}

// 3-D fp32 map over a row-major matrix [rows][width] (leading dimension ld) seen as [width / 32 column groups][rows][32]:
// ONE box {32 columns, box_rows, box_groups} lands in shared memory as [group][row][32] -- the MN-major operand layout of the
// weight-gradient GEMMs -- so a stage is filled by one TMA instruction per operand instead of one per 32-column group
int make_map3(CUtensorMap* m, const float* base, long long width, long long rows, long long ld, int box_rows, int box_groups) {
  int rc = ensure_encode();
  if (rc) return rc;
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)((width + 31) / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_groups};
  cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed (%d): base %p width %lld rows %lld ld %lld box %d rows x %d groups", (int)r, (const void*)base,
              width, rows, ld, box_rows, box_groups);
    return MAPPO_ERR_CUDA;
  }
  return MAPPO_OK;
}

// pair mode (cta_group::2, 256-row tiles, each CTA stages half of the B tile): implemented, parity-tested and MEASURED -- it moves
// a third less operand data per SM (ncu: 672 MB instead of 1008 MB through the crossbar per 512 x 512 layer at c5) and is not
// faster (22.5 vs 22.9 ms of forward GEMMs per c5 iteration; 3, 4 or 5 operand stages, one or two epilogue groups all land within
// 5 %): the producer sits waiting for free stages, i.e. the MMA stream itself is the limit (tests/cuda/umma_rate.cu measures the
// instruction's own rate).  Opt-in with MAPPO_B200_PAIR_LIN=1; the weight-gradient GEMM does use pairs (big_grad_pair.cu, + 5 %).
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MAPPO_B200_PAIR_LIN"); v = (e && e[0] == '1') ? 1 : 0; }
  return v != 0;
}

template <class Epi>
static int lin_launch_t(const LinOperands& o, const typename Epi::Args& ea, LinShape sh, const char* name, cudaStream_t st) {
  sh.n_rowblocks = (sh.n_rows + 127) / 128;
  {
    static int direct = -1;
    if (direct < 0) { const char* e = getenv("MAPPO_B200_DIRECT_EPI"); direct = (e && e[0] == '1') ? 1 : 0; }
    sh.direct = (direct && sh.BN >= 128 && ((reinterpret_cast<uintptr_t>(o.out) | reinterpret_cast<uintptr_t>(o.ain)) & 15) == 0 &&
                 o.ldo % 4 == 0 && o.ldain % 4 == 0) ? 1 : 0;
    sh.out_ptr = o.out; sh.out_ld = o.ldo; sh.ain_ptr = o.ain; sh.ain_ld = o.ldain;
  }
  const int NT = sh.N / sh.BN;
  const bool pair = pair_enabled() && sh.BN == 256 && (NT == 1 || NT % 2 == 0) && sh.n_rowblocks >= 2 && o.sm_count >= 2;
  CUtensorMap mA, mB, mO, mI;
  int rc = make_map(&mA, o.A, sh.K, sh.n_rows, o.lda, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_map(&mB, o.W, sh.K, sh.N, o.ldw, 32, pair ? sh.BN / 2 : sh.BN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  mO = mA; mI = mA;
  if (Epi::kStoresOut && sh.store_out) { rc = make_map(&mO, o.out, sh.N, sh.n_rows, o.ldo, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc; }
  if (Epi::kHasAin) { rc = make_map(&mI, o.ain, sh.N, sh.n_rows, o.ldain, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc; }
  sh.n_stages = pair ? kPairStages : (sh.BN >= 256 ? 3 : 4);
  const LinSmem L = make_lin_smem(sh.BN, sh.n_stages, sh.N, Epi::kNeedsScratch, pair);
  const size_t bytes = (size_t)L.total + 1024;
  if (bytes > 227 * 1024) { set_error("%s: %zu B shared memory > 227 KB (N = %d)", name, bytes, sh.N); return MAPPO_ERR_UNSUPPORTED; }
  if (pair) {
    if (cudaFuncSetAttribute(big_lin_kernel<Epi, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return check_launch("big_lin_kernel: cudaFuncSetAttribute");
    const int n_ublocks = (sh.n_rowblocks + 1) / 2, max_pairs = o.sm_count / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (n_ublocks < max_pairs ? n_ublocks : max_pairs));
    cfg.blockDim = dim3(64 + 128 * kPairGroups);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, big_lin_kernel<Epi, true>, mA, mB, mO, mI, ea, sh) != cudaSuccess) return check_launch(name);
    return check_launch(name);
  }
  if (cudaFuncSetAttribute(big_lin_kernel<Epi, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return check_launch("big_lin_kernel: cudaFuncSetAttribute");
  const int grid = sh.n_rowblocks < o.sm_count ? sh.n_rowblocks : o.sm_count;
  big_lin_kernel<Epi, false><<<grid, kLinThreads, bytes, st>>>(mA, mB, mO, mI, ea, sh);
  return check_launch(name);
}

int lin_fwd_launch(const LinOperands& o, const EpiFwd::Args& ea, const LinShape& sh, cudaStream_t st) { return lin_launch_t<EpiFwd>(o, ea, sh, "big_lin_kernel<EpiFwd>", st); }
int lin_bwd_launch(const LinOperands& o, const EpiBwd::Args& ea, const LinShape& sh, cudaStream_t st) { return lin_launch_t<EpiBwd>(o, ea, sh, "big_lin_kernel<EpiBwd>", st); }
int lin_head_launch(const LinOperands& o, const EpiHead::Args& ea, const LinShape& sh, cudaStream_t st) { return lin_launch_t<EpiHead>(o, ea, sh, "big_lin_kernel<EpiHead>", st); }
int lin_sample_launch(const LinOperands& o, const EpiSample::Args& ea, const LinShape& sh, cudaStream_t st) { return lin_launch_t<EpiSample>(o, ea, sh, "big_lin_kernel<EpiSample>", st); }

int grad_gemm_launch(const float* P, int ldp, const float* Q, int ldq_in, float* partial, GradShape sh, cudaStream_t st) {
  CUtensorMap mP, mQ8, mQr;
  const int last = sh.qw[sh.n_tiles - 1] / 32, rem = last >= 8 ? last - 8 : last;       // groups the last tile loads through mapQr
  int rc = make_map3(&mP, P, sh.Pw, sh.rows, ldp, kGradKR, 4);
  if (rc) return rc;
  rc = make_map3(&mQ8, Q, sh.Qw, sh.rows, ldq_in, kGradKR, 8);
  if (rc) return rc;
  rc = make_map3(&mQr, Q, sh.Qw, sh.rows, ldq_in, kGradKR, rem > 0 ? rem : 1);
  if (rc) return rc;
  const size_t bytes = (size_t)kGradStages * kGradStageBytes + 1024;
  if (cudaFuncSetAttribute(big_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return check_launch("big_grad_kernel: cudaFuncSetAttribute");
  big_grad_kernel<<<sh.splits * sh.m_tiles * sh.n_tiles, kLinThreads, bytes, st>>>(mP, mQ8, mQr, partial, sh);
  return check_launch("big_grad_kernel");
}

}  // namespace big
}  // namespace mappo
