// tc64.cuh -- PTX wrappers, tile layouts and the folded weight image shared by the hidden-64 tcgen05 kernels
// (update_mlp_tc.cu: fused MLP step; update_gru_tc.cu: GRU sequence kernels).  Encodings as probed on the B200 by
// tests/cuda/tc_probe.cu.
#pragma once
#include "net_tiles.cuh"

namespace mappo {

constexpr int kTM = 128;                 // rows per tile, threads per CTA
constexpr int kHF = 72;                  // hidden features incl. the constant-1 feature, padded to a multiple of 8
constexpr int kHC = kHF / 4;             // 18 chunks
constexpr int kOne = 64;                 // index of the constant-1 feature in hidden tiles

// ------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  return ok != 0;
}
// a pipeline bug must surface as a launch failure, not as a hung GPU: the wait traps after ~2 s of spinning
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  if (mbar_try(a, parity)) return;
  const long long t0 = clock64();
  for (uint32_t it = 1;; ++it) {
    if (mbar_try(a, parity)) return;
    if ((it & 0xFFFu) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {     // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {       // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tanh on the SFU (MUFU.TANH): max relative error 2^-11, the same class as the tf32 rounding of the GEMM inputs.
__device__ __forceinline__ float act_fwd_tc(float z, int act) {
  if (act == ACT_RELU) return fmaxf(z, 0.f);
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(z));
  return y;
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// n MMAs over consecutive K-slices of two operands: the slice-to-slice advance only changes the 14-bit start-address
// field (bits [0,14), in 16-byte units), so the descriptors are built once and bumped by a constant (shared memory ends
// below 2^18 bytes: the field cannot overflow into the next one)
__device__ __forceinline__ void umma_seq(uint32_t d_tmem, uint32_t a_addr, uint32_t a_step, uint32_t a_lbo, uint32_t b_addr,
                                         uint32_t b_step, uint32_t b_lbo, uint32_t idesc, int n, bool accumulate_first) {
  uint64_t ad = make_desc(a_addr, a_lbo, 128), bd = make_desc(b_addr, b_lbo, 128);
  const uint64_t da = a_step >> 4, db = b_step >> 4;
  for (int s = 0; s < n; ++s) {
    umma_tf32(d_tmem, ad, bd, idesc, (accumulate_first || s > 0) ? 1u : 0u);
    ad += da; bd += db;
  }
}

// instruction descriptor: D fp32, A/B tf32 (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------
// folded weight image (global), built once per optimiser step by pack_tc_kernel
// ------------------------------------------------------------------------------------------------------------
struct TcImage {
  int inF;      // input features incl. constant-1, padded to a multiple of 8
  int NH;       // head outputs padded to a multiple of 16
  // float offsets.  forward (K = in-features):  w1 [inF/4][64][4], w2 [18][64][4], wh [18][NH][4]
  //                 dX      (K = out-features): w2t [16][64][4] (rows = in-feature k), wht [NH/4][64][4]
  int w1, w2, wh, w2t, wht, total;
};
__host__ __device__ inline TcImage make_tc_image(const NetDev& n) {
  TcImage m;
  m.inF = (n.in_dim + 1 + 7) & ~7;
  m.NH = (n.head_total + 15) & ~15;
  m.w1 = 0;
  m.w2 = m.w1 + m.inF * 64;
  m.wh = m.w2 + kHF * 64;
  m.w2t = m.wh + kHF * m.NH;
  m.wht = m.w2t + 64 * 64;
  m.total = m.wht + m.NH * 64;
  return m;
}

// Raw per-CTA gradient slot of the tcgen05 kernel: the folded accumulators exactly as they sit in TMEM.
//   g2 [64][72] = dW2'  (column 64 = db2'),  g1 [64][inF] = dW1' (column in = db1'),  gh [64][NH] = dWh'^T (row = feature),
//   dbh [NH].  mappo_update_finish sums the slots and unfolds ONCE (tc_unfold_kernel) instead of once per CTA.
struct TcRaw { int g2, g1, gh, dbh, total; };
__host__ __device__ inline TcRaw make_tc_raw(const TcImage& m) {
  TcRaw r;
  r.g2 = 0;
  r.g1 = r.g2 + 64 * kHF;
  r.gh = r.g1 + 64 * m.inF;
  r.dbh = r.gh + 64 * m.NH;
  r.total = (r.dbh + m.NH + 3) & ~3;
  return r;
}

constexpr int kS65 = 65, kS73 = 73;       // padded row strides of the transposed tiles (odd -> conflict-free scatter)
constexpr int kTCThreads = 2 * kTM;       // two threads per row: warpgroup g owns hidden columns [32 g, 32 g + 32)

// The two threads of a row live in warps w and w + 4 (same TMEM lane window).  They meet on named barrier 1 + (w & 3)
// (64 threads) and swap two partial sums through shared memory; both get bit-identical totals (a + b == b + a).
// Two slots alternate so that a fast pair cannot overwrite values its partner has not read yet.
struct PairXch {
  float2* buf;       // [slot][warpgroup][128]
  int wg, r, bar;
  int slot;
  __device__ __forceinline__ float2 sum(float a, float b) {
    buf[(slot * 2 + wg) * kTM + r] = make_float2(a, b);
    asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
    const float2 o = buf[(slot * 2 + (wg ^ 1)) * kTM + r];
    slot ^= 1;
    return make_float2(a + o.x, b + o.y);
  }
};

// LayerNorm statistics of a 64-wide row held as 2 x 32 register values (two-pass like torch)
__device__ __forceinline__ void ln_stats_pair(const float* a, PairXch& px, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
  mean = px.sum(s, 0.f).x * (1.f / 64.f);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { const float d = a[i] - mean; v = fmaf(d, d, v); }
  rstd = 1.0f / sqrtf(px.sum(v, 0.f).x * (1.f / 64.f) + kLnEps);
}

// thread (row r, warpgroup g): write its 32 values as chunks [8 g, 8 g + 8) of row r of a K-major staging tile
// [16 (+2 aug)][128][4] ...
__device__ __forceinline__ void put_kmajor32(float* P, int r, int wg, const float* v, bool aug) {
#pragma unroll
  for (int kc = 0; kc < 8; ++kc)
    reinterpret_cast<float4*>(P)[(wg * 8 + kc) * kTM + r] = make_float4(v[4 * kc], v[4 * kc + 1], v[4 * kc + 2], v[4 * kc + 3]);
  if (aug) reinterpret_cast<float4*>(P)[(16 + wg) * kTM + r] = make_float4(wg == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
}
// ... and as features [32 g, 32 g + 32) of column r of a transposed tile [32][S][4] (element (feature f, row r) at
// ((r/4)*S + f)*4 + r%4)
__device__ __forceinline__ void put_transposed32(float* T, int S, int r, int wg, const float* v) {
  float* base = T + ((r >> 2) * S + wg * 32) * 4 + (r & 3);
#pragma unroll
  for (int f = 0; f < 32; ++f) base[f * 4] = v[f];
}

// ---- [position][64] fp32 workspace planes of the recurrent pipeline, tiled like the K-major operands (update_gru_tc.cu) ----
__host__ __device__ inline int64_t plane_floats(int64_t n_rows) { return ((n_rows + kTM - 1) / kTM) * (int64_t)(16 * kTM * 4); }
// float offset of the 4-column chunk `chunk` (0..15) of position p
__device__ __forceinline__ size_t pl_off(size_t p, int chunk) { return ((p >> 7) * 16 + (size_t)chunk) * (kTM * 4) + (p & (kTM - 1)) * 4; }
__device__ __forceinline__ void ld_pl16(const float* __restrict__ plane, size_t p, int chunk0, float* v) {      // 4 chunks = 16 columns
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(plane + pl_off(p, chunk0 + j)));
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
// the same load, issued where it is written (asm volatile: the compiler may not sink it towards its first use -- a prefetch)
__device__ __forceinline__ void ld_pl16_pinned(const float* __restrict__ plane, size_t p, int chunk0, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3]) : "l"(plane + pl_off(p, chunk0 + j)));
}
__device__ __forceinline__ void st_pl16(float* __restrict__ plane, size_t p, int chunk0, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(plane + pl_off(p, chunk0 + j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

}  // namespace mappo
