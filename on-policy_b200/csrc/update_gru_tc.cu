// update_gru_tc.cu -- the recurrent (GRU) training step on Blackwell tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in
// TMEM).  Same contract as update_gru_launch (update_gru.cu): a time-major [L, Nc] chunked minibatch (utils/shared_buffer.py:
// 557-604), RNNLayer semantics h <- h * mask_t before every cell (rnn.py:27, :43-77), LayerNorm after the GRU (rnn.py:79).
//
// Only the recurrence itself is sequential, so only it runs as sequence kernels; everything else is position-parallel:
//   1. update_mlp_tc_kernel<TC_BASE_FWD>   base MLP of every position                         -> X  (xhat2: pre-affine LN output)
//   2. gru_tc_fwd_kernel                   128-chunk tile, l = 0..L-1: gates by UMMA, cell math out of TMEM -> R, Z, N, GHN, H
//   3. update_mlp_tc_kernel<TC_HEAD>       every position: LayerNorm(h) -> heads -> loss -> dL/dh of the head path -> DHH
//   4. gru_tc_bwd_kernel                   128-chunk tile, l = L-1..0: BPTT, dh_{l-1} = dgh W_hh by UMMA -> DR, DZ, DN
//   5. gru_tc_grad_kernel                  every position: dL/dxhat2 = dgi W_ih', dW_ih' / dW_hh / db accumulated in TMEM -> DFEAT
//   6. update_mlp_tc_kernel<TC_BASE_BWD>   base MLP backward of every position (forward recomputed on chip)
//   7. slot sums + unfold of the LayerNorm folding -> the complete flat gradient
// Workspace planes are [position][64] fp32.  All UMMA operands use the K-major no-swizzle layout of tc64.cuh; the LayerNorm
// affine in front of the GRU (base.mlp.fc2[0] LayerNorm) is folded into W_ih' = W_ih diag(gamma), b' = b_ih + W_ih beta
// (+ b_hh for the r and z gates, whose pre-activations are plain sums), the rnn.norm affine into the heads.
#include <cstdlib>
#include <vector>
#include "tc64.cuh"

namespace mappo {

int grad_reduce_launch(const float*, int, int, float*, float*, int*, cudaStream_t);
int update_mlp_tc_mode_launch(int mode, const NetDev& n, const float* params, const float* image, const BatchDev& b, const LossDev& L,
                              const double* norm_stats, const double* adv_stats, const float* vn_state, float* grad_part, int n_ctas,
                              double* loss_out, const float* plane_in, float* plane_out, cudaStream_t st);
int update_mlp_tc_pack_launch(const NetDev& n, const float* params, float* image, cudaStream_t st);
int update_mlp_tc_unfold_launch(const NetDev& n, const float* params, const float* raw_sum, float* grad, float* sumsq_part, cudaStream_t st);
int64_t update_mlp_tc_workspace_floats(const NetDev& n);
int update_mlp_tc_slot_floats(const NetDev& n);
int64_t update_gru_workspace_floats(const NetDev& n, int n_rows);

constexpr int kG3 = 192;                         // 3 gates x 64
constexpr int kXS = 145;                         // row stride of the combined [x | 1 | hm | 1] transposed tile (144 features, odd pad)
constexpr int kGcat = 3 * 64 * 144;              // raw GRU gradient slot: [gate][64 outputs][x 0..63, db_x 64, pad, hm 72..135, db_h 136, pad]
constexpr uint32_t kRowB = kTM * 16;             // chunk stride (bytes) of a 128-row K-major tile

// folded GRU weight images (global, rebuilt every optimiser step)
//   wih  [18][192][4]  forward B operand, K = 64 features + constant-1 (bias column) + pad
//   whh  [16][192][4]  forward B operand
//   bhn  [64]          b_hh of the n gate (added in the epilogue: it sits inside r * (.))
//   whht [48][64][4]   BPTT B operand: rows = hidden index k, K = gate outputs o           (dh = dgh W_hh)
//   wiht [48][64][4]   input-gradient B operand: W_ih'[o][k]                                 (dxhat2 = dgi W_ih')
struct GruImage { int wih, whh, bhn, whht, wiht, fwd_floats, total; };
__host__ __device__ inline GruImage make_gru_image() {
  GruImage m;
  m.wih = 0;
  m.whh = m.wih + kHC * kG3 * 4;
  m.bhn = m.whh + 16 * kG3 * 4;
  m.fwd_floats = m.bhn + 64;
  m.whht = m.fwd_floats;
  m.wiht = m.whht + 48 * 64 * 4;
  m.total = m.wiht + 48 * 64 * 4;
  return m;
}

// Workspace planes hold one 64-wide fp32 row per minibatch position p in the K-major tile layout of the UMMA operands:
// [p / 128][16 chunks of 4 columns][128 rows][4] -- a warp (32 consecutive positions) then reads / writes 512 contiguous bytes per
// 16-byte access instead of 32 rows 256 bytes apart (pl_off / ld_pl16 / st_pl16 in tc64.cuh).
struct GruPlanes { float *X, *R, *Z, *N, *GHN, *H, *DHH, *DR, *DZ, *DN, *DFEAT; };

__global__ void __launch_bounds__(256) gru_pack_kernel(const NetDev n, const float* __restrict__ p, float* __restrict__ img) {
  const GruImage m = make_gru_image();
  const float* Wih = p + n.g.gru_wih;
  const float* Whh = p + n.g.gru_whh;
  const float* gam = p + n.g.ln2_w[0];
  const float* bet = p + n.g.ln2_b[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.total; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    bool round = true;
    if (i < m.whh) {
      const int kc = i / (kG3 * 4), o = (i >> 2) % kG3, k = kc * 4 + (i & 3);
      if (k < 64) v = Wih[o * 64 + k] * gam[k];
      else if (k == kOne) {
        v = p[n.g.gru_bih + o] + (o < 128 ? p[n.g.gru_bhh + o] : 0.f);
        for (int j0 = 0; j0 < 64; ++j0) { const int j = (j0 + o) & 63; v = fmaf(Wih[o * 64 + j], bet[j], v); }
      }
    } else if (i < m.bhn) {
      const int t = i - m.whh, kc = t / (kG3 * 4), o = (t >> 2) % kG3, k = kc * 4 + (t & 3);
      v = Whh[o * 64 + k];
    } else if (i < m.fwd_floats) {
      v = p[n.g.gru_bhh + 128 + (i - m.bhn)];
      round = false;
    } else if (i < m.wiht) {
      const int t = i - m.whht, oc = t / 256, k = (t >> 2) & 63, o = oc * 4 + (t & 3);
      v = Whh[o * 64 + k];
    } else {
      const int t = i - m.wiht, oc = t / 256, k = (t >> 2) & 63, o = oc * 4 + (t & 3);
      v = Wih[o * 64 + k] * gam[k];
    }
    img[i] = round ? to_tf32(v) : v;
  }
}

__device__ __forceinline__ void ld_half16(const float* __restrict__ src, float* v) {      // 16 consecutive floats (64-byte aligned)
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) { const float4 t = __ldg(s4 + q); v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
}
__device__ __forceinline__ void st_half16(float* __restrict__ dst, const float* v) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
// 16 values = chunks [c4, c4 + 4) of row r of a K-major [chunks][128][4] tile
__device__ __forceinline__ void put_kmajor16(float* T, int c4, int r, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(T)[(c4 + q) * kTM + r] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// -------------------------------------------------------------------------------------------------------------------
// 2. sequence forward.  512 threads = FOUR threads per chunk row: thread (row r, quarter q) owns hidden columns [16 q, 16 q + 16) of
// every gate; warps w, w + 4, w + 8, w + 12 address the same 32 TMEM lanes.  The state h stays in fp32 registers across the L steps
// (only the UMMA operand copy is rounded to tf32).
// TMEM: two accumulator sets of 256 columns (step parity): [0,64) r, [64,128) z (x part, then the h part accumulated on top),
// [128,192) W_hn hm + b_hn (b_hn preset with tcgen05.st, the h part accumulated on top), [192,256) W_in x + b_in.
// The x part of step l + 1 does not depend on the state: it is issued right behind the h part of step l (into the other set, from
// the other XA buffer) and runs under the cell math of step l, so the critical path of a step is the cell math, one barrier and
// the 8 state MMAs (M 128, N 192, K 8 each).
// -------------------------------------------------------------------------------------------------------------------
constexpr int kSeqThreads = 512;

// clock64 stamps of CTA 0 / thread 0 inside step 2 of its first tile (mappo_debug_gru_cycles): [0] top of the step, [1] state MMAs of
// the previous step done, [2] cell math + plane stores issued, [3] operand tiles written, [4] barrier passed, [5] MMAs issued,
// [6] prefetch loads issued; [8..] the same points of the BPTT kernel ([8] top, [9] loads landed + gate math + stores, [10] operand
// written, [11] barrier, [12] MMAs issued, [13] MMAs done, [14] dh updated)
__device__ long long g_gru_cycles[16];
#define GRU_STAMP(i, cond) do { if (blockIdx.x == 0 && tid == 0 && (cond)) g_gru_cycles[i] = clock64(); } while (0)

__device__ __forceinline__ float tanh_ap(float x) {            // MUFU.TANH, relative error 2^-11
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm_gate(float x) { return fmaf(0.5f, tanh_ap(0.5f * x), 0.5f); }      // absolute error <= 2.5e-4
__device__ __forceinline__ float tanh_cell(float x) {          // 1 - 2 / (1 + e^(2x)): absolute error ~ 2e-7
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  return 1.f - __fdividef(2.f, 1.f + e);
}
// 16 values = features [c0, c0 + 16) of column r of a transposed tile [32][S][4]
__device__ __forceinline__ void put_transposed16(float* T, int S, int r, int c0, const float* v) {
  float* base = T + ((r >> 2) * S + c0) * 4 + (r & 3);
#pragma unroll
  for (int f = 0; f < 16; ++f) base[f * 4] = v[f];
}

__global__ void __launch_bounds__(kSeqThreads, 1)
gru_tc_fwd_kernel(const NetDev n, const float* __restrict__ gimg, const BatchDev b, const GruPlanes ws, int n_seq_tiles) {
  extern __shared__ __align__(1024) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & (kTM - 1), q = tid >> 7, c0 = q * 16;
  const GruImage im = make_gru_image();
  float* sImg = smem;
  float* XA0 = sImg + im.fwd_floats;                // 2 x [18][128][4]
  float* HA = XA0 + 2 * kHC * kTM * 4;              // [16][128][4]
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(HA + 16 * kTM * 4);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int Nc = b.n_seq, Lsteps = b.seq_len;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // constant-1 feature (chunk 16) and pad (chunk 17) of both XA buffers: 2 x 2 x 128 float4, one per thread
  reinterpret_cast<float4*>(XA0 + (tid >> 8) * kHC * kTM * 4)[(16 + ((tid >> 7) & 1)) * kTM + r] =
      make_float4(((tid >> 7) & 1) == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(im.fwd_floats * sizeof(float)), half = (uint32_t)(im.whh * sizeof(float));
    mbar_expect_tx(bar_w, bytes);
    tma_bulk_g2s(sImg, gimg, half, bar_w);
    tma_bulk_g2s(sImg + im.whh, gimg + im.whh, bytes - half, bar_w);
  }
  const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
  const uint32_t aXA = smem_u32(XA0), aHA = smem_u32(HA), aWih = smem_u32(sImg + im.wih), aWhh = smem_u32(sImg + im.whh);
  constexpr uint32_t kXAB = kHC * kTM * 16;         // bytes of one XA buffer
  float bhn[16];                                    // b_hh of the n gate for this thread's columns
  ld_half16(gimg + im.bhn + c0, bhn);
  const float* h0 = n.is_critic ? b.h0_critic : b.h0_actor;
  uint32_t phase = 0;
  bool first = true;

  // x part of a step: D[0,128) = xaug W_ih'[r, z]^T, D[192,256) = xaug W_ih'[n]^T
  auto issue_x = [&](int step) {
    const uint32_t d = tmem + 256u * (step & 1), a = aXA + (step & 1) * kXAB;
    umma_seq(d, a, 2 * kRowB, kRowB, aWih, 2 * kG3 * 16, kG3 * 16, make_idesc(128, 128, 0, 0), kHF / 8, false);
    umma_seq(d + 192, a, 2 * kRowB, kRowB, aWih + 128 * 16, 2 * kG3 * 16, kG3 * 16, make_idesc(128, 64, 0, 0), kHF / 8, false);
  };
  auto preset_bhn = [&](int step) {                 // D[128 + c0, +16) of this row <- b_hn
    tmem_st8(tmem + 256u * (step & 1) + lane_base + 128 + c0, bhn);
    tmem_st8(tmem + 256u * (step & 1) + lane_base + 128 + c0 + 8, bhn + 8);
  };

  for (int st = blockIdx.x; st < n_seq_tiles; st += gridDim.x) {
    const int c = st * kTM + r;
    const bool valid = c < Nc;
    float h[16], xq[16];
    float mq = 0.f;
    int grq = -1;                                   // storage row of the step after the one mq belongs to
    auto load_x = [&](int l) {                       // xq <- X row of step l (zeros past the end)
      if (valid && l < Lsteps) ld_pl16(ws.X, (size_t)l * Nc + c, q * 4, xq);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) xq[i] = 0.f;
      }
    };
    // mask of step l through the storage-row index fetched one step earlier, then the index of step l + 1: two independent loads
    // (rows[p] -> masks[gr] back to back would park the warp on the first one: in-order issue)
    auto load_gr = [&](int l) {
      grq = -1;
      if (valid && l < Lsteps) { const size_t p = (size_t)l * Nc + c; grq = b.rows ? b.rows[p] : (int)p; }
    };
    auto load_m = [&](int l) {
      mq = grq >= 0 ? b.masks[grq] : 0.f;
      load_gr(l + 1);
    };
    if (valid) ld_half16(h0 + (size_t)(b.seq_first ? b.seq_first[c] : c) * 64 + c0, h);
    else {
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = 0.f;
    }
    load_x(0);
    load_gr(0);
    load_m(0);
    put_kmajor16(XA0, q * 4, r, xq);
    preset_bhn(0);
    load_x(1);
    fence_async_smem();
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      if (first) mbar_wait(bar_w, 0);
      issue_x(0);
    }
    first = false;

    // cell math of step l out of accumulator set l & 1: h (= hm_l, masked state of that step) -> h_l; planes of step l
    auto cell = [&](int l) {
      const uint32_t d = tmem + 256u * (l & 1) + lane_base + c0;
      const size_t pp = (size_t)l * Nc + c;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float ar[8], az[8], ah[8], an[8];
        tmem_ld8(d + hh * 8, ar);
        tmem_ld8(d + 64 + hh * 8, az);
        tmem_ld8(d + 128 + hh * 8, ah);
        tmem_ld8(d + 192 + hh * 8, an);
        tmem_ld_wait();
        float hn[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {                // torch GRU cell, gate order (r, z, n)
          const float rg = sigm_gate(ar[j]);
          const float zg = sigm_gate(az[j]);
          const float ng = tanh_cell(fmaf(rg, ah[j], an[j]));
          hn[j] = fmaf(zg, h[hh * 8 + j] - ng, ng);  // (1 - z) n + z hm
          ar[j] = rg; az[j] = zg; an[j] = ng;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) h[hh * 8 + j] = hn[j];
        if (valid) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const size_t o = pl_off(pp, q * 4 + hh * 2 + t);
            *reinterpret_cast<float4*>(ws.R + o) = make_float4(ar[4 * t], ar[4 * t + 1], ar[4 * t + 2], ar[4 * t + 3]);
            *reinterpret_cast<float4*>(ws.Z + o) = make_float4(az[4 * t], az[4 * t + 1], az[4 * t + 2], az[4 * t + 3]);
            *reinterpret_cast<float4*>(ws.N + o) = make_float4(an[4 * t], an[4 * t + 1], an[4 * t + 2], an[4 * t + 3]);
            *reinterpret_cast<float4*>(ws.GHN + o) = make_float4(ah[4 * t], ah[4 * t + 1], ah[4 * t + 2], ah[4 * t + 3]);
            *reinterpret_cast<float4*>(ws.H + o) = make_float4(hn[4 * t], hn[4 * t + 1], hn[4 * t + 2], hn[4 * t + 3]);
          }
        }
      }
    };

    for (int l = 0; l < Lsteps; ++l) {
      const bool stamp = l == 2 && st == (int)blockIdx.x;
      GRU_STAMP(0, stamp);
      if (l > 0) {
        mbar_wait(bar_m, phase); phase ^= 1;         // state MMAs of step l - 1 done (and, in order before them, its x part)
        tc_fence_after();
        GRU_STAMP(1, stamp);
        cell(l - 1);
      }
      GRU_STAMP(2, stamp);
      {
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { h[i] *= mq; t[i] = to_tf32(h[i]); }      // h <- h * mask_l (rnn.py:27, :67)
        put_kmajor16(HA, q * 4, r, t);
      }
      if (l + 1 < Lsteps) {
        put_kmajor16(XA0 + ((l + 1) & 1) * kHC * kTM * 4, q * 4, r, xq);
        preset_bhn(l + 1);                           // (this thread has read its columns of that set in cell(l - 1) above)
      }
      fence_async_smem();
      tmem_st_wait();
      tc_fence_before();
      GRU_STAMP(3, stamp);
      __syncthreads();
      GRU_STAMP(4, stamp);
      if (tid == 0) {
        tc_fence_after();
        umma_seq(tmem + 256u * (l & 1), aHA, 2 * kRowB, kRowB, aWhh, 2 * kG3 * 16, kG3 * 16, make_idesc(128, 192, 0, 0), 8, true);
        umma_commit(bar_m);
        if (l + 1 < Lsteps) issue_x(l + 1);
      }
      GRU_STAMP(5, stamp);
      load_x(l + 2);                                 // in flight under the MMAs and the cell math
      load_m(l + 1);
      GRU_STAMP(6, stamp);
    }
    mbar_wait(bar_m, phase); phase ^= 1;
    tc_fence_after();
    cell(Lsteps - 1);
    tc_fence_before();
    __syncthreads();                                 // every read of both accumulator sets is done before the next tile presets them
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// -------------------------------------------------------------------------------------------------------------------
// 4. sequence backward (BPTT), four threads per chunk row like the forward.  dh carried in fp32 registers; per step the gate
// gradients go to the DR / DZ / DN planes (fp32) and, tf32-rounded, into the K-major operand of dh_{l-1} += dgh W_hh (K = 192 gate
// outputs, 24 MMAs M 128 x N 64).
// -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSeqThreads, 1)
gru_tc_bwd_kernel(const NetDev n, const float* __restrict__ gimg, const BatchDev b, const GruPlanes ws, int n_seq_tiles) {
  extern __shared__ __align__(1024) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & (kTM - 1), q = tid >> 7, c0 = q * 16;
  const GruImage im = make_gru_image();
  float* sW = smem;                                 // whht [48][64][4]
  float* DG = sW + 48 * 64 * 4;                     // [48][128][4]: (dr, dz, dn * r)
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(DG + 48 * kTM * 4);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int Nc = b.n_seq, Lsteps = b.seq_len;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t bytes = 48 * 64 * 4 * sizeof(float);
    mbar_expect_tx(bar_w, bytes);
    tma_bulk_g2s(sW, gimg + im.whht, bytes, bar_w);
  }
  const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
  const uint32_t aDG = smem_u32(DG), aW = smem_u32(sW);
  const float* h0 = n.is_critic ? b.h0_critic : b.h0_actor;
  uint32_t phase = 0;
  bool first = true;
  for (int st = blockIdx.x; st < n_seq_tiles; st += gridDim.x) {
    const int c = st * kTM + r;
    const bool valid = c < Nc;
    const int src = valid ? (b.seq_first ? b.seq_first[c] : c) : 0;
    float dh[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dh[i] = 0.f;
    // mask of step l through the row index fetched one step earlier (no rows[p] -> masks[gr] chain inside a step)
    auto row_of = [&](int l) { const size_t pp = (size_t)l * Nc + c; return (valid && l >= 0) ? (b.rows ? b.rows[pp] : (int)pp) : -1; };
    int grq = row_of(Lsteps - 1);
    float mq = grq >= 0 ? b.masks[grq] : 0.f;
    grq = row_of(Lsteps - 2);
    for (int l = Lsteps - 1; l >= 0; --l) {
      const size_t p = (size_t)l * Nc + c;
      const float m = mq;
      const bool stamp = l == Lsteps - 3 && st == (int)blockIdx.x;
      GRU_STAMP(8, stamp);
      float rg[16], zg[16], ng[16], gh[16], dd[16], hp[16];
      if (valid) {
        ld_pl16(ws.R, p, q * 4, rg);
        ld_pl16(ws.Z, p, q * 4, zg);
        ld_pl16(ws.N, p, q * 4, ng);
        ld_pl16(ws.GHN, p, q * 4, gh);
        ld_pl16(ws.DHH, p, q * 4, dd);
        if (l > 0) ld_pl16(ws.H, p - (size_t)Nc, q * 4, hp);
        else ld_half16(h0 + (size_t)src * 64 + c0, hp);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { rg[j] = zg[j] = ng[j] = gh[j] = dd[j] = hp[j] = 0.f; }
      }
      mq = grq >= 0 ? b.masks[grq] : 0.f;          // step l - 1 (in flight under this step)
      grq = row_of(l - 2);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float d = dh[j] + dd[j];                                          // dL/dh_l: future steps + head path
        const float hm = hp[j] * m;
        const float dn_pre = d * (1.f - zg[j]) * (1.f - ng[j] * ng[j]);
        const float dz_pre = d * (hm - ng[j]) * zg[j] * (1.f - zg[j]);
        const float dr_pre = dn_pre * gh[j] * rg[j] * (1.f - rg[j]);
        dh[j] = d * zg[j];                                                      // direct path h' = ... + z * hm
        dd[j] = dr_pre; hp[j] = dz_pre; gh[j] = dn_pre;
        ng[j] = to_tf32(dn_pre * rg[j]);
      }
      if (valid) {
        st_pl16(ws.DR, p, q * 4, dd);
        st_pl16(ws.DZ, p, q * 4, hp);
        st_pl16(ws.DN, p, q * 4, gh);
      }
      if (l == 0) break;                                                        // h0 is data: no gradient beyond the first step
      GRU_STAMP(9, stamp);
#pragma unroll
      for (int j = 0; j < 16; ++j) { dd[j] = to_tf32(dd[j]); hp[j] = to_tf32(hp[j]); }
      put_kmajor16(DG, 0 + q * 4, r, dd);
      put_kmajor16(DG, 16 + q * 4, r, hp);
      put_kmajor16(DG, 32 + q * 4, r, ng);
      fence_async_smem();
      tc_fence_before();
      GRU_STAMP(10, stamp);
      __syncthreads();
      GRU_STAMP(11, stamp);
      if (tid == 0) {
        tc_fence_after();
        if (first) mbar_wait(bar_w, 0);
        umma_seq(tmem, aDG, 2 * kRowB, kRowB, aW, 2 * 1024, 1024, make_idesc(128, 64, 0, 0), 24, false);
        umma_commit(bar_m);
      }
      first = false;
      GRU_STAMP(12, stamp);
      mbar_wait(bar_m, phase); phase ^= 1;
      tc_fence_after();
      GRU_STAMP(13, stamp);
      {
        float t[16];
        tmem_ld16(tmem + lane_base + c0, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) dh[i] = (dh[i] + t[i]) * m;               // hm = h_{l-1} * mask_l
      }
      GRU_STAMP(14, stamp);
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// -------------------------------------------------------------------------------------------------------------------
// 5. gate gradients of every position: dL/dxhat2 = dgi W_ih' (K = 192, three gate passes accumulate into D) and the weight-gradient
// GEMMs over the rows of the tile, Gcat_g[o][*] += dg_g^T [x | 1 | hm | 1] (M = 64 gate outputs, N = 144, K = 128 rows), persistent
// in TMEM across the tiles of a CTA.  The n gate's hidden-side gradient is dn * r, so its two halves use separate A tiles.
// TMEM: [0,64) D, [64 + 144 g, 64 + 144 g + 144) Gcat_g.
// -------------------------------------------------------------------------------------------------------------------
struct GradSmem { int w, xht, dgk, dgt, dht, misc, total; };
__host__ __device__ inline GradSmem make_grad_smem() {
  GradSmem s;
  int o = 0;
  s.w = o; o += 48 * 64 * 4;
  s.xht = o; o += 32 * kXS * 4;
  s.dgk = o; o += 16 * kTM * 4;
  s.dgt = o; o += 32 * kS65 * 4;
  s.dht = o; o += 32 * kS65 * 4;
  s.misc = o; o += 16;
  s.total = o;
  return s;
}

__global__ void __launch_bounds__(kSeqThreads, 1)
gru_tc_grad_kernel(const NetDev n, const float* __restrict__ gimg, const BatchDev b, const GruPlanes ws, float* __restrict__ slots,
                   int n_tiles) {
  extern __shared__ __align__(1024) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & (kTM - 1), q = tid >> 7, c0 = q * 16;     // four threads per row, 16 columns each
  const GruImage im = make_gru_image();
  const GradSmem sm = make_grad_smem();
  float* sW = smem + sm.w;
  float* XHT = smem + sm.xht;
  float* DGk = smem + sm.dgk;
  float* DGT = smem + sm.dgt;
  float* DHT = smem + sm.dht;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + sm.misc);
  uint64_t* bar_m = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int Nc = b.n_seq, P = b.n_rows;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (q == 0) {                                      // constant-1 features 64 / 136 and the zero pads of both halves
    float* base = XHT + (r >> 2) * kXS * 4 + (r & 3);
#pragma unroll
    for (int f = 64; f < 72; ++f) { base[f * 4] = (f == 64) ? 1.f : 0.f; base[(f + 72) * 4] = (f == 64) ? 1.f : 0.f; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t bytes = 48 * 64 * 4 * sizeof(float);
    mbar_expect_tx(bar_w, bytes);
    tma_bulk_g2s(sW, gimg + im.wiht, bytes, bar_w);
  }
  const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
  const uint32_t aW = smem_u32(sW), aXHT = smem_u32(XHT), aDGk = smem_u32(DGk), aDGT = smem_u32(DGT), aDHT = smem_u32(DHT);
  const float* h0 = n.is_critic ? b.h0_critic : b.h0_actor;
  uint32_t phase = 0;
  bool first_tile = true;
  // The six plane rows of a tile travel in registers (pinned loads): x, hm are consumed at the top of the tile, dg3 / rg by the three gate
  // passes -- all dead once the last pass has written its operand tiles, so the NEXT tile's rows are fetched there and land under the
  // last pass's MMAs and the result read-out (one exposed round trip per CTA instead of one per tile).
  float x[16], hm[16], rg[16], dg3[3][16];
  float mk = 0.f;
  auto fetch_tile = [&](int tile) {
    const int p = tile * kTM + r;
    if (tile < n_tiles && p < P) {
      const int l = p / Nc, c = p - l * Nc;
      const int gr = b.rows ? b.rows[p] : p;           // index first, mask (dependent) behind the plane loads
      ld_pl16_pinned(ws.X, (size_t)p, q * 4, x);
      if (l > 0) ld_pl16_pinned(ws.H, (size_t)(p - Nc), q * 4, hm);
      else ld_half16(h0 + (size_t)(b.seq_first ? b.seq_first[c] : c) * 64 + c0, hm);
      ld_pl16_pinned(ws.R, (size_t)p, q * 4, rg);
      ld_pl16_pinned(ws.DR, (size_t)p, q * 4, dg3[0]);
      ld_pl16_pinned(ws.DZ, (size_t)p, q * 4, dg3[1]);
      ld_pl16_pinned(ws.DN, (size_t)p, q * 4, dg3[2]);
      mk = b.masks[gr];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) { x[i] = 0.f; hm[i] = 0.f; rg[i] = 0.f; dg3[0][i] = dg3[1][i] = dg3[2][i] = 0.f; }
      mk = 0.f;
    }
  };
  fetch_tile(blockIdx.x);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int p = tile * kTM + r;
    const bool valid = p < P;
    {
      float t[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = to_tf32(hm[i] * mk);
      put_transposed16(XHT, kXS, r, c0, x);
      put_transposed16(XHT, kXS, r, 72 + c0, t);
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float dg[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) dg[i] = to_tf32(dg3[g][i]);
      put_kmajor16(DGk, q * 4, r, dg);
      put_transposed16(DGT, kS65, r, c0, dg);
      if (g == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dg[i] = to_tf32(dg[i] * rg[i]);
        put_transposed16(DHT, kS65, r, c0, dg);
        fetch_tile(tile + gridDim.x);                  // every register of this tile's rows is dead: next tile's rows in flight
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        if (first_tile && g == 0) mbar_wait(bar_w, 0);
        umma_seq(tmem, aDGk, 2 * kRowB, kRowB, aW + g * 16 * 1024, 2 * 1024, 1024, make_idesc(128, 64, 0, 0), 8, g > 0);
        const uint32_t cG = 64 + g * 144;
        if (g < 2) {
          umma_seq(tmem + cG, aDGT, 2 * kS65 * 16, kS65 * 16, aXHT, 2 * kXS * 16, kXS * 16, make_idesc(64, 144, 0, 0), kTM / 8, !first_tile);
        } else {
          umma_seq(tmem + cG, aDGT, 2 * kS65 * 16, kS65 * 16, aXHT, 2 * kXS * 16, kXS * 16, make_idesc(64, 72, 0, 0), kTM / 8, !first_tile);
          umma_seq(tmem + cG + 72, aDHT, 2 * kS65 * 16, kS65 * 16, aXHT + 72 * 16, 2 * kXS * 16, kXS * 16, make_idesc(64, 72, 0, 0), kTM / 8,
                   !first_tile);
        }
        umma_commit(bar_m);
      }
      mbar_wait(bar_m, phase); phase ^= 1;             // the pass's operand tiles are free again
      tc_fence_after();
    }
    {
      float d[16];
      tmem_ld16(tmem + lane_base + c0, d);
      tmem_ld_wait();
      if (valid) st_pl16(ws.DFEAT, (size_t)p, q * 4, d);
    }
    first_tile = false;
    tc_fence_before();
  }
  // raw accumulators -> this CTA's slot [3][64][144]; warps 0..7: warpgroup w = q dumps columns [72 w, 72 w + 72) of every gate
  if (q < 2) {
    const int wg = q;
    float* gslot = slots + (size_t)blockIdx.x * kGcat;
    const bool has_tile = !first_tile;
    const int o = (warp & 3) * 16 + lane;              // accumulator row of this thread in the M = 64 layout
    const bool own = lane < 16;
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
      float v[72];
      const uint32_t cG = 64 + g * 144 + wg * 72;
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld16(tmem + lane_base + cG + q * 16, v + q * 16);
      tmem_ld8(tmem + lane_base + cG + 64, v + 64);
      tmem_ld_wait();
      if (own) {
        float4* dst = reinterpret_cast<float4*>(gslot + ((size_t)g * 64 + o) * 144 + wg * 72);
#pragma unroll
        for (int q = 0; q < 18; ++q)
          dst[q] = has_tile ? make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// slot-summed raw GRU accumulators G [192][144] -> flat gradient entries of gru.{weight,bias}_{ih,hh} and of the LayerNorm in front of
// the GRU (chain rule of the folding: dW = dW' diag(gamma) + db' beta^T, dgamma = colsum(dW' .* W), dbeta = W^T db').
// Grid 4 blocks of 16 hidden columns; 256 threads = 16 columns x 16 groups of 12 gate outputs.
__global__ void __launch_bounds__(256)
gru_unfold_kernel(const NetDev n, const float* __restrict__ p, const float* __restrict__ G, float* __restrict__ g) {
  __shared__ float part_g[16][17], part_b[16][17];
  const int tid = threadIdx.x, kx = tid & 15, og = tid >> 4, k = blockIdx.x * 16 + kx;
  const float gam = p[n.g.ln2_w[0] + k], bet = p[n.g.ln2_b[0] + k];
  float sg = 0.f, sb = 0.f;
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    const int o = og * 12 + j;
    const float dw = G[o * 144 + k], w = p[n.g.gru_wih + o * 64 + k], dbo = G[o * 144 + 64];
    g[n.g.gru_wih + o * 64 + k] = fmaf(dbo, bet, dw * gam);
    g[n.g.gru_whh + o * 64 + k] = G[o * 144 + 72 + k];
    sg = fmaf(dw, w, sg);
    sb = fmaf(dbo, w, sb);
  }
  part_g[og][kx] = sg; part_b[og][kx] = sb;
  __syncthreads();
  if (og == 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) { a += part_g[q][kx]; c += part_b[q][kx]; }
    g[n.g.ln2_w[0] + k] = a;
    g[n.g.ln2_b[0] + k] = c;
  }
  if (blockIdx.x == 0 && tid < kG3) {
    g[n.g.gru_bih + tid] = G[tid * 144 + 64];
    g[n.g.gru_bhh + tid] = G[tid * 144 + 136];
  }
}

// -------------------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------------------
// diagnostic (bench.py's roofline leg, eager passes only -- no events while a stream is capturing): CUDA-event time of each kernel
// family of the pipeline
enum { TG_PACK = 0, TG_BASE_FWD, TG_SEQ_FWD, TG_HEAD, TG_BPTT, TG_GATE_GRAD, TG_BASE_BWD, TG_FINISH, TG_N };
struct GruTimedLaunch { cudaEvent_t a, b; int cat; };
static bool g_gru_timing = false;
static std::vector<GruTimedLaunch> g_gru_timed;
struct GruTimed {
  cudaStream_t st; cudaEvent_t a; int cat; bool on;
  GruTimed(int c, cudaStream_t s) : st(s), a(nullptr), cat(c), on(false) {
    if (!g_gru_timing) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    on = cudaEventCreate(&a) == cudaSuccess && cudaEventRecord(a, s) == cudaSuccess;
  }
  ~GruTimed() {
    if (!on) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) == cudaSuccess && cudaEventRecord(b, st) == cudaSuccess) g_gru_timed.push_back({a, b, cat});
  }
};
int debug_gru_cycles(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_gru_cycles, sizeof(long long) * 16) == cudaSuccess ? 0 : MAPPO_ERR_CUDA;
}
int debug_gru_timing(int enable, double* ms_out, long long* n_out) {
  for (int i = 0; i < TG_N; ++i) { if (ms_out) ms_out[i] = 0.0; if (n_out) n_out[i] = 0; }
  for (const GruTimedLaunch& t : g_gru_timed) {
    cudaEventSynchronize(t.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.a, t.b);
    if (ms_out) ms_out[t.cat] += ms;
    if (n_out) n_out[t.cat] += 1;
    cudaEventDestroy(t.a); cudaEventDestroy(t.b);
  }
  g_gru_timed.clear();
  g_gru_timing = enable != 0;
  return MAPPO_OK;
}

bool update_gru_tc_supported(const NetDev& n) {
  return n.recurrent && n.hid == 64 && n.layer_n == 1 && n.in_dim <= 63 && n.head_total <= 32;
}

// MAPPO_B200_GRU_CTAS caps the persistent grids (tests: several tiles per CTA at small sizes)
static int cta_cap(int sm_count) {
  const char* e = getenv("MAPPO_B200_GRU_CTAS");
  if (e) { const int v = atoi(e); if (v > 0 && v < sm_count) return v; }
  return sm_count;
}

static int pos_ctas(int n_rows, int sm_count) {
  const int tiles = (n_rows + kTM - 1) / kTM;
  return tiles < sm_count ? tiles : sm_count;
}

struct GruTcWs { int64_t img_base, img_gru, raw_base, raw_gru, scratch, slots_tc, slots_gru, planes, total; };
static GruTcWs make_gru_tc_ws(const NetDev& n, int n_rows, int sm_count) {
  GruTcWs w;
  auto up = [](int64_t v) { return (v + 31) & ~(int64_t)31; };
  const int ctas = pos_ctas(n_rows, sm_count);
  int64_t o = 0;
  w.img_base = o; o += up(update_mlp_tc_workspace_floats(n));
  w.img_gru = o; o += up(make_gru_image().total);
  w.raw_base = o; o += up(update_mlp_tc_slot_floats(n));
  w.raw_gru = o; o += up(kGcat);
  w.scratch = o; o += 64;
  w.slots_tc = o; o += up((int64_t)2 * ctas * update_mlp_tc_slot_floats(n));
  w.slots_gru = o; o += up((int64_t)ctas * kGcat);
  w.planes = o; o += (int64_t)10 * plane_floats(n_rows);
  w.total = o;
  return w;
}

int64_t update_gru_tc_workspace_floats(const NetDev& n, int n_rows, int sm_count) {
  const int64_t a = make_gru_tc_ws(n, n_rows, sm_count).total, f = update_gru_workspace_floats(n, n_rows);
  return a > f ? a : f;                       // evaluate_actions runs the exact-fp32 kernels of update_gru.cu in the same workspace
}

template <typename K>
static int set_smem(K kern, size_t bytes, SmemConfig& cfg, const char* what) {
  size_t& configured = cfg.slot();
  if (bytes > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return check_launch(what);
    configured = bytes;
  }
  return MAPPO_OK;
}

// leaves the complete flat gradient in grad_out (the caller's slot 0)
int update_gru_tc_launch(const NetDev& n, const float* params, const BatchDev& b, const LossDev& L, const double* norm_stats,
                         const double* adv_stats, const float* vn_state, float* grad_out, double* loss_out, float* workspace,
                         int sm_count, cudaStream_t st) {
  if (!update_gru_tc_supported(n)) { set_error("update_gru_tc: configuration not built for the tcgen05 path"); return MAPPO_ERR_UNSUPPORTED; }
  const float* h0 = n.is_critic ? b.h0_critic : b.h0_actor;
  if (!h0 || !b.masks) { set_error("update_gru_tc: h0 / masks missing"); return MAPPO_ERR_INVALID; }
  if (b.eval_only) { set_error("update_gru_tc: evaluation runs the fp32 kernels"); return MAPPO_ERR_INVALID; }
  if ((reinterpret_cast<uintptr_t>(workspace) & 127) != 0) { set_error("update_gru_tc: workspace must be 128-byte aligned"); return MAPPO_ERR_INVALID; }
  if ((int64_t)b.n_seq * b.seq_len != b.n_rows) { set_error("update_gru_tc: n_rows != n_seq * seq_len"); return MAPPO_ERR_INVALID; }
  const GruTcWs w = make_gru_tc_ws(n, b.n_rows, sm_count);      // (offsets from the uncapped SM count, like the allocation)
  sm_count = cta_cap(sm_count);
  const GruImage im = make_gru_image();
  float* img_base = workspace + w.img_base;
  float* img_gru = workspace + w.img_gru;
  const size_t plane = (size_t)plane_floats(b.n_rows);
  float* pl = workspace + w.planes;
  GruPlanes ws;
  ws.X = pl; ws.R = pl + plane; ws.Z = pl + 2 * plane; ws.N = pl + 3 * plane; ws.GHN = pl + 4 * plane; ws.H = pl + 5 * plane;
  ws.DHH = pl + 6 * plane; ws.DR = pl + 7 * plane; ws.DZ = pl + 8 * plane; ws.DN = pl + 9 * plane;
  ws.DFEAT = ws.DHH;                               // dead after the BPTT kernel
  const int ctas = pos_ctas(b.n_rows, sm_count);
  const int n_tiles = (b.n_rows + kTM - 1) / kTM;
  const int n_seq_tiles = (b.n_seq + kTM - 1) / kTM;
  const int seq_ctas = n_seq_tiles < sm_count ? n_seq_tiles : sm_count;
  const int R = update_mlp_tc_slot_floats(n);
  float* slots_tc = workspace + w.slots_tc;
  float* slots_gru = workspace + w.slots_gru;
  int rc;
  // weight images of the current parameters
  {
    GruTimed t(TG_PACK, st);
    if ((rc = update_mlp_tc_pack_launch(n, params, img_base, st))) return rc;
    gru_pack_kernel<<<(im.total + 255) / 256, 256, 0, st>>>(n, params, img_gru);
    if ((rc = check_launch("gru_pack_kernel"))) return rc;
  }
  // 1. base forward
  {
    GruTimed t(TG_BASE_FWD, st);
    if ((rc = update_mlp_tc_mode_launch(1, n, params, img_base, b, L, norm_stats, adv_stats, vn_state, nullptr, ctas, loss_out, nullptr, ws.X, st)))
      return rc;
  }
  // 2. sequence forward
  {
    GruTimed t(TG_SEQ_FWD, st);
    static thread_local SmemConfig cfg = {};
    const size_t bytes = (size_t)(im.fwd_floats + 2 * kHC * kTM * 4 + 16 * kTM * 4 + 16) * sizeof(float) + 1024;
    if ((rc = set_smem(gru_tc_fwd_kernel, bytes, cfg, "gru_tc_fwd: cudaFuncSetAttribute"))) return rc;
    gru_tc_fwd_kernel<<<seq_ctas, kSeqThreads, bytes, st>>>(n, img_gru, b, ws, n_seq_tiles);
    if ((rc = check_launch("gru_tc_fwd_kernel"))) return rc;
  }
  // 3. heads + loss of every position
  {
    GruTimed t(TG_HEAD, st);
    if ((rc = update_mlp_tc_mode_launch(3, n, params, img_base, b, L, norm_stats, adv_stats, vn_state, slots_tc, ctas, loss_out, ws.H, ws.DHH, st)))
      return rc;
  }
  // 4. BPTT
  {
    GruTimed t(TG_BPTT, st);
    static thread_local SmemConfig cfg = {};
    const size_t bytes = (size_t)(48 * 64 * 4 + 48 * kTM * 4 + 16) * sizeof(float) + 1024;
    if ((rc = set_smem(gru_tc_bwd_kernel, bytes, cfg, "gru_tc_bwd: cudaFuncSetAttribute"))) return rc;
    gru_tc_bwd_kernel<<<seq_ctas, kSeqThreads, bytes, st>>>(n, img_gru, b, ws, n_seq_tiles);
    if ((rc = check_launch("gru_tc_bwd_kernel"))) return rc;
  }
  // 5. gate gradients
  {
    GruTimed t(TG_GATE_GRAD, st);
    static thread_local SmemConfig cfg = {};
    const size_t bytes = (size_t)make_grad_smem().total * sizeof(float) + 1024;
    if ((rc = set_smem(gru_tc_grad_kernel, bytes, cfg, "gru_tc_grad: cudaFuncSetAttribute"))) return rc;
    gru_tc_grad_kernel<<<ctas, kSeqThreads, bytes, st>>>(n, img_gru, b, ws, slots_gru, n_tiles);
    if ((rc = check_launch("gru_tc_grad_kernel"))) return rc;
  }
  // 6. base backward
  {
    GruTimed t(TG_BASE_BWD, st);
    if ((rc = update_mlp_tc_mode_launch(2, n, params, img_base, b, L, norm_stats, adv_stats, vn_state, slots_tc + (size_t)ctas * R, ctas, loss_out,
                                        ws.DFEAT, nullptr, st)))
      return rc;
  }
  // 7. slot sums + unfold
  GruTimed t_fin(TG_FINISH, st);
  float* raw_base = workspace + w.raw_base;
  float* raw_gru = workspace + w.raw_gru;
  if ((rc = grad_reduce_launch(slots_tc, 2 * ctas, R, raw_base, nullptr, nullptr, st))) return rc;
  if ((rc = update_mlp_tc_unfold_launch(n, params, raw_base, grad_out, workspace + w.scratch, st))) return rc;
  if ((rc = grad_reduce_launch(slots_gru, ctas, kGcat, raw_gru, nullptr, nullptr, st))) return rc;
  gru_unfold_kernel<<<4, 256, 0, st>>>(n, params, raw_gru, grad_out);
  return check_launch("gru_unfold_kernel");
}

}  // namespace mappo
