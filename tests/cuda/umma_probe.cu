// umma_probe.cu -- stand-alone hardware probe for the TMA-fed tcgen05 GEMM pipeline of csrc/big_*.cu (hidden >= 128 nets).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tests/cuda/umma_probe tests/cuda/umma_probe.cu
// Run on a B200: prints one line per case, "PASS"/"FAIL" with the max abs error against a CPU product.
//
//   T1  K-major tf32 operands, TMA SWIZZLE_128B boxes {32 x rows}, UMMA SW128 descriptors, M=128 N=256 K=64
//   T2  MN-major tf32 operands (contraction over the ROWS of row-major matrices: the weight-gradient GEMM), TMA
//       128B-swizzle-with-32B-atom boxes, UMMA layout type SWIZZLE_128B_BASE32B; a few LBO/SBO/swizzle variants
//   T3  epilogue staging: threads write a [128 x 32] tile with the 128B-swizzle formula, TMA store -> global
//   T4  cta_group::2 (CTA pair, M=256): 2SM TMA loads signalling the leader's barrier, one MMA for both SMs
// Every wait is bounded (a failed handshake reports TIMEOUT instead of hanging the box).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled g_encode = nullptr;

static CUtensorMap make_map_2d(const float* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems, uint32_t box_inner,
                               uint32_t box_outer, CUtensorMapSwizzle swz) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (swizzle %d)\n", (int)r, (int)swz); memset(&m, 0, sizeof(m)); }
  return m;
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ int g_timeout;

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return true;
  }
  g_timeout = 1;
  return false;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start [0,14) lbo [16,30) sbo [32,46) version=1 @46, layout [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(layout & 7u) << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                 "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
                 "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
                 "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// T1 / T2: one CTA, 128 threads.  a_boxes x [a_box_bytes] and b_boxes x [b_box_bytes] are loaded per K block.
struct GemmCase {
  int n_kblocks;             // TMA K blocks
  int mma_per_kblock;        // MMAs (K = 8) per K block
  int a_boxes, b_boxes;      // TMA boxes per K block and operand
  int a_box_bytes, b_box_bytes;
  int a_c0_step, a_c1_step, a_c0_kstep, a_c1_kstep;     // box coordinates: (i*c0_step + kb*c0_kstep, i*c1_step + kb*c1_kstep)
  int b_c0_step, b_c1_step, b_c0_kstep, b_c1_kstep;
  uint32_t a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv, layout;       // descriptor fields (bytes), per-MMA start advance
  int a_mn, b_mn, N;
};

__global__ void __launch_bounds__(128) gemm_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                  const GemmCase c, float* __restrict__ D /* [128][256] */) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tslot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int a_kb_bytes = c.a_boxes * c.a_box_bytes, b_kb_bytes = c.b_boxes * c.b_box_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + c.n_kblocks * a_kb_bytes;
  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  bool ok = true;
  if (tid == 0) {
    mbar_expect_tx(&bar_full, (uint32_t)(c.n_kblocks * (a_kb_bytes + b_kb_bytes)));
    for (int kb = 0; kb < c.n_kblocks; ++kb) {
      for (int i = 0; i < c.a_boxes; ++i)
        tma_load_2d(sA + kb * a_kb_bytes + i * c.a_box_bytes, &mapA, i * c.a_c0_step + kb * c.a_c0_kstep,
                    i * c.a_c1_step + kb * c.a_c1_kstep, &bar_full);
      for (int i = 0; i < c.b_boxes; ++i)
        tma_load_2d(sB + kb * b_kb_bytes + i * c.b_box_bytes, &mapB, i * c.b_c0_step + kb * c.b_c0_kstep,
                    i * c.b_c1_step + kb * c.b_c1_kstep, &bar_full);
    }
    ok = mbar_wait(&bar_full, 0);
    tc_fence_after();
    if (ok) {
      const uint32_t id = make_idesc(128, c.N, c.a_mn, c.b_mn);
      int n = 0;
      for (int kb = 0; kb < c.n_kblocks; ++kb)
        for (int k = 0; k < c.mma_per_kblock; ++k, ++n) {
          const uint64_t ad = make_desc(smem_u32(sA + kb * a_kb_bytes) + k * c.a_kadv, c.a_lbo, c.a_sbo, c.layout);
          const uint64_t bd = make_desc(smem_u32(sB + kb * b_kb_bytes) + k * c.b_kadv, c.b_lbo, c.b_sbo, c.layout);
          umma_tf32(tmem, ad, bd, id, n > 0 ? 1u : 0u);
        }
      umma_commit(&bar_mma);
    }
  }
  __syncthreads();
  if (!g_timeout) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    if (!g_timeout) {
      for (int c0 = 0; c0 < c.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32; ++j) D[tid * 256 + c0 + j] = v[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// T3: swizzled staging + TMA store, and TMA load + swizzled read
__global__ void __launch_bounds__(128) store_probe(const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapIn,
                                                   float* __restrict__ readback /* [128][32] */) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tile = reinterpret_cast<float*>(smem);                 // [128 rows][32 floats], 128B-swizzled
  float* tile2 = reinterpret_cast<float*>(smem + 16384);
  const int r = threadIdx.x;
  for (int ch = 0; ch < 8; ++ch) {
    float4 v = make_float4(r * 100.f + ch * 4 + 0, r * 100.f + ch * 4 + 1, r * 100.f + ch * 4 + 2, r * 100.f + ch * 4 + 3);
    *reinterpret_cast<float4*>(smem + r * 128 + ((ch ^ (r & 7)) << 4)) = v;
  }
  fence_async_smem();
  if (r == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (r == 0) {
    tma_store_2d(&mapOut, 32, 0, tile);                          // columns [32, 64) of the [128][64] output
    tma_store_commit_wait();
    mbar_expect_tx(&bar, 16384);
    tma_load_2d(tile2, &mapIn, 32, 0, &bar);                     // columns [32, 64) of the input
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  for (int ch = 0; ch < 8; ++ch) {
    const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<uint8_t*>(tile2) + r * 128 + ((ch ^ (r & 7)) << 4));
    readback[r * 32 + ch * 4 + 0] = v.x; readback[r * 32 + ch * 4 + 1] = v.y;
    readback[r * 32 + ch * 4 + 2] = v.z; readback[r * 32 + ch * 4 + 3] = v.w;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// T4: CTA pair.  D[256][256] = A[256][64] B[256][64]^T; CTA r owns A rows [128 r, 128 r + 128) and B rows (= N) likewise.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
pair_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ D /* [256][256] */) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tslot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint8_t* sA = smem;              // 2 k-blocks x 16 KB
  uint8_t* sB = smem + 32768;      // 2 k-blocks x 16 KB
  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (tid == 0) {
    if (rank == 0) mbar_expect_tx(&bar_full, 2 * 2 * 32768);                       // both CTAs' bytes land on the leader's barrier
    const uint32_t bar_leader = smem_u32(&bar_full) & 0xFEFFFFFFu;                 // peer bit cleared -> CTA 0 of the pair
    for (int kb = 0; kb < 2; ++kb) {
      asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(sA + kb * 16384)), "l"(reinterpret_cast<uint64_t>(&mapA)), "r"(kb * 32), "r"((int)rank * 128),
                     "r"(bar_leader) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(sB + kb * 16384)), "l"(reinterpret_cast<uint64_t>(&mapB)), "r"(kb * 32), "r"((int)rank * 128),
                     "r"(bar_leader) : "memory");
    }
    if (rank == 0) {
      const bool ok = mbar_wait(&bar_full, 0);
      tc_fence_after();
      if (ok) {
        const uint32_t id = make_idesc(256, 256, 0, 0);
        int n = 0;
        for (int kb = 0; kb < 2; ++kb)
          for (int k = 0; k < 4; ++k, ++n) {
            const uint64_t ad = make_desc(smem_u32(sA + kb * 16384) + k * 32, 16, 1024, 2);
            const uint64_t bd = make_desc(smem_u32(sB + kb * 16384) + k * 32, 16, 1024, 2);
            umma_tf32_2cta(tmem, ad, bd, id, n > 0 ? 1u : 0u);
          }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar_mma)), "h"((uint16_t)3) : "memory");
      }
    }
  }
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  if (!g_timeout) {
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) D[(rank * 128 + tid) * 256 + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
static float rnd_tf32() { return (float)((rand() % 33) - 16) / 8.0f; }        // exactly representable, small

static int reset_timeout() { int z = 0; CK(cudaMemcpyToSymbol(g_timeout, &z, sizeof(int))); return 0; }
static int read_timeout() { int z = 0; CK(cudaMemcpyFromSymbol(&z, g_timeout, sizeof(int))); return z; }

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
  g_encode = (EncodeTiled)fn;
  srand(7);
  int fails = 0;

  // ---------------- T1: K-major ----------------
  {
    const int M = 128, N = 256, K = 64;
    std::vector<float> A(M * K), B(N * K), Dh(128 * 256), ref(M * N);
    for (auto& x : A) x = rnd_tf32();
    for (auto& x : B) x = rnd_tf32();
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = (float)s; }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, Dh.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, Dh.size() * 4));
    CUtensorMap mA = make_map_2d(dA, K, M, K, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mB = make_map_2d(dB, K, N, K, 32, 256, CU_TENSOR_MAP_SWIZZLE_128B);
    GemmCase c{};
    c.n_kblocks = 2; c.mma_per_kblock = 4; c.a_boxes = 1; c.b_boxes = 1; c.a_box_bytes = 128 * 128; c.b_box_bytes = 256 * 128;
    c.a_c0_kstep = 32; c.b_c0_kstep = 32;
    c.a_lbo = 16; c.a_sbo = 1024; c.a_kadv = 32; c.b_lbo = 16; c.b_sbo = 1024; c.b_kadv = 32; c.layout = 2; c.N = N;
    const int smem = 2 * (c.a_box_bytes + c.b_box_bytes) + 1024;
    CK(cudaFuncSetAttribute(gemm_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    reset_timeout();
    gemm_probe<<<1, 128, smem>>>(mA, mB, c, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0; for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) err = fmax(err, fabs((double)Dh[m * 256 + n] - ref[m * N + n]));
    const int to = read_timeout();
    printf("T1 K-major SW128 TMA+UMMA M128 N256 K64: %s max_err %.3g%s\n", (err < 1e-3 && !to) ? "PASS" : "FAIL", err, to ? " TIMEOUT" : "");
    fails += !(err < 1e-3 && !to);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  // ---------------- T2: MN-major variants ----------------
  {
    const int M = 128, N = 256, K = 64;
    std::vector<float> P(K * M), Q(K * N), Dh(128 * 256), ref(M * N);
    for (auto& x : P) x = rnd_tf32();
    for (auto& x : Q) x = rnd_tf32();
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)P[k * M + m] * Q[k * N + n]; ref[m * N + n] = (float)s; }
    float *dP, *dQ, *dD;
    CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dQ, Q.size() * 4)); CK(cudaMalloc(&dD, Dh.size() * 4));
    CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice));
    struct Var { const char* name; CUtensorMapSwizzle swz; uint32_t layout; int krows_per_box; uint32_t sbo, lbo_is_box; uint32_t kadv; int swap; };
    // boxes are {32 MN-elements (128 B), K rows}: one box per 32-wide MN group, the whole K extent (64 rows) per box
    const Var vars[] = {
      {"ATOM_32B  layout1 SBO=512(4 k-rows) LBO=box", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1, 64, 512, 1, 1024, 0},
      {"ATOM_32B  layout1 swapped LBO/SBO", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1, 64, 512, 1, 1024, 1},
      {"ATOM_32B_FLIP_8B layout1", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B, 1, 64, 512, 1, 1024, 0},
      {"ATOM_32B_FLIP_8B layout1 swapped", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B, 1, 64, 512, 1, 1024, 1},
      {"SWIZZLE_128B layout2 SBO=1024(8 k-rows) LBO=box", CU_TENSOR_MAP_SWIZZLE_128B, 2, 64, 1024, 1, 1024, 0},
      {"SWIZZLE_128B layout2 swapped", CU_TENSOR_MAP_SWIZZLE_128B, 2, 64, 1024, 1, 1024, 1},
    };
    for (const Var& v : vars) {
      CUtensorMap mP = make_map_2d(dP, M, K, M, 32, 64, v.swz);
      CUtensorMap mQ = make_map_2d(dQ, N, K, N, 32, 64, v.swz);
      GemmCase c{};
      c.n_kblocks = 1; c.mma_per_kblock = 8; c.a_boxes = 4; c.b_boxes = 8; c.a_box_bytes = 64 * 128; c.b_box_bytes = 64 * 128;
      c.a_c0_step = 32; c.b_c0_step = 32;
      const uint32_t box = 64 * 128;
      c.a_lbo = v.swap ? v.sbo : box; c.a_sbo = v.swap ? box : v.sbo; c.b_lbo = c.a_lbo; c.b_sbo = c.a_sbo;
      c.a_kadv = v.kadv; c.b_kadv = v.kadv; c.layout = v.layout; c.a_mn = 1; c.b_mn = 1; c.N = N;
      const int smem = 4 * box + 8 * box + 1024;
      CK(cudaFuncSetAttribute(gemm_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaMemset(dD, 0, Dh.size() * 4));
      reset_timeout();
      gemm_probe<<<1, 128, smem>>>(mP, mQ, c, dD);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("T2 %s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 3; }
      CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
      double err = 0; int bad = 0;
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { const double d = fabs((double)Dh[m * 256 + n] - ref[m * N + n]); err = fmax(err, d); bad += d > 1e-3; }
      const int to = read_timeout();
      printf("T2 MN-major %-52s: %s max_err %.3g bad %d/%d%s\n", v.name, (err < 1e-3 && !to) ? "PASS" : "fail", err, bad, M * N, to ? " TIMEOUT" : "");
    }
    cudaFree(dP); cudaFree(dQ); cudaFree(dD);
  }

  // ---------------- T3: staging + TMA store / load ----------------
  {
    std::vector<float> in(128 * 64), out(128 * 64, -1.f), rb(128 * 32);
    for (int i = 0; i < 128 * 64; ++i) in[i] = (float)i;
    float *dIn, *dOut, *dRb;
    CK(cudaMalloc(&dIn, in.size() * 4)); CK(cudaMalloc(&dOut, out.size() * 4)); CK(cudaMalloc(&dRb, rb.size() * 4));
    CK(cudaMemcpy(dIn, in.data(), in.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dOut, out.data(), out.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap mO = make_map_2d(dOut, 64, 128, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mI = make_map_2d(dIn, 64, 128, 64, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CK(cudaFuncSetAttribute(store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
    reset_timeout();
    store_probe<<<1, 128, 34 * 1024>>>(mO, mI, dRb);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rb.data(), dRb, rb.size() * 4, cudaMemcpyDeviceToHost));
    int bad_s = 0, bad_l = 0;
    for (int r = 0; r < 128; ++r) for (int cc = 0; cc < 32; ++cc) {
      bad_s += out[r * 64 + 32 + cc] != r * 100.f + cc;
      bad_s += out[r * 64 + cc] != -1.f;
      bad_l += rb[r * 32 + cc] != in[r * 64 + 32 + cc];
    }
    printf("T3 swizzled staging -> TMA store: %s (bad %d);  TMA load -> swizzled read: %s (bad %d)%s\n", bad_s ? "FAIL" : "PASS", bad_s,
           bad_l ? "FAIL" : "PASS", bad_l, read_timeout() ? " TIMEOUT" : "");
    fails += (bad_s || bad_l);
    cudaFree(dIn); cudaFree(dOut); cudaFree(dRb);
  }

  // ---------------- T4: CTA pair ----------------
  {
    const int M = 256, N = 256, K = 64;
    std::vector<float> A(M * K), B(N * K), Dh(M * N), ref(M * N);
    for (auto& x : A) x = rnd_tf32();
    for (auto& x : B) x = rnd_tf32();
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = (float)s; }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, Dh.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, Dh.size() * 4));
    CUtensorMap mA = make_map_2d(dA, K, M, K, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mB = make_map_2d(dB, K, N, K, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    const int smem = 65536 + 1024;
    CK(cudaFuncSetAttribute(pair_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    reset_timeout();
    pair_probe<<<2, 128, smem>>>(mA, mB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T4 CTA pair: CUDA error %s\n", cudaGetErrorString(e)); return 4; }
    CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0; for (int i = 0; i < M * N; ++i) err = fmax(err, fabs((double)Dh[i] - ref[i]));
    const int to = read_timeout();
    printf("T4 cta_group::2 M256 N256 K64: %s max_err %.3g%s\n", (err < 1e-3 && !to) ? "PASS" : "fail", err, to ? " TIMEOUT" : "");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  printf("umma_probe done, required failures: %d\n", fails);
  return fails ? 1 : 0;
}
