// tc_probe.cu -- stand-alone probe of the tcgen05 kind::tf32 descriptor encodings used by update_mlp_tc.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/tc_probe tests/cuda/tc_probe.cu
// Each case runs ONE GEMM D[M,N] = A[M,K] B[N,K]^T through tcgen05.mma from "chunked" shared-memory tiles
// ([feature/4][row][4] floats) and compares with a CPU product.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

struct Case {
  int M, N, ksteps;          // ksteps MMAs of K = 8
  int a_mn, b_mn;            // 0 = K-major, 1 = MN-major
  uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step;   // bytes
  int a_floats, b_floats;    // sizes of the operand images
};

__global__ void __launch_bounds__(128) probe(const Case c, const float* __restrict__ A, const float* __restrict__ B,
                                             float* __restrict__ D /* [128 lanes][256 cols] raw TMEM dump */) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  float* sA = smem;
  float* sB = smem + c.a_floats;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < c.a_floats; i += 128) sA[i] = A[i];
  for (int i = tid; i < c.b_floats; i += 128) sB[i] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tslot;
  // zero the accumulator region first so "never written" is visible
  {
    uint32_t z = 0;
    for (int col = 0; col < 256; col += 1)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + col), "r"(z));
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    for (int s = 0; s < c.ksteps; ++s) {
      const uint64_t da = make_desc(smem_u32(sA) + s * c.a_step, c.a_lbo, c.a_sbo);
      const uint64_t db = make_desc(smem_u32(sB) + s * c.b_step, c.b_lbo, c.b_sbo);
      const uint32_t accum = s > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int col = 0; col < 256; col += 8) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + col));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) D[tid * 256 + col + j] = __uint_as_float(u[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static float tf32r(float x) {   // round-to-nearest-even to 10 mantissa bits (what cvt.rna does up to ties)
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xFFFFE000u;
  float y; memcpy(&y, &u, 4); return y;
}

// chunked image of an R x F matrix X[r][f]: ((f/4)*R + r)*4 + f%4
static std::vector<float> chunked(const std::vector<float>& X, int R, int F) {
  std::vector<float> o((size_t)R * F, 0.f);
  for (int r = 0; r < R; ++r) for (int f = 0; f < F; ++f) o[((size_t)(f / 4) * R + r) * 4 + f % 4] = X[(size_t)r * F + f];
  return o;
}

int main() {
  const int ROWB = 128 * 16;
  struct Named { const char* name; Case c; int Ka; /* contraction length */ bool a_feat_is_k, b_feat_is_k; int Ra, Fa, Rb, Fb; };
  std::vector<Named> cases;
  // 1. forward: A = X[128 rows][64 feats] K-major, B = W[64 out][64 in] K-major, K = 64
  cases.push_back({"M128 N64  A K-major  B K-major (forward)", {128, 64, 8, 0, 0, (uint32_t)ROWB, 128, 2u * ROWB, 1024, 128, 2048, 128 * 64, 64 * 64}, 64, true, true, 128, 64, 64, 64});
  // 2. dX: A = dY[128][64] K-major (K = o), B = W[o=64 rows][k=64 feats] used MN-major (N = k, K = o)
  cases.push_back({"M128 N64  A K-major  B MN-major lbo=128 sbo=R*16 (dX, current)", {128, 64, 8, 0, 1, (uint32_t)ROWB, 128, 2u * ROWB, 128, 1024, 128, 128 * 64, 64 * 64}, 64, true, false, 128, 64, 64, 64});
  cases.push_back({"M128 N64  A K-major  B MN-major lbo=R*16 sbo=128 (swapped)", {128, 64, 8, 0, 1, (uint32_t)ROWB, 128, 2u * ROWB, 1024, 128, 128, 128 * 64, 64 * 64}, 64, true, false, 128, 64, 64, 64});
  // 3. dW: A = dY[128 rows][64 feats] MN-major (M = feat, K = rows), B = X[128 rows][72 feats] MN-major (N = feat)
  cases.push_back({"M64  N72  A MN-major B MN-major lbo=128 sbo=ROWB (dW, current)", {64, 72, 16, 1, 1, 128, (uint32_t)ROWB, 128, 128, (uint32_t)ROWB, 128, 128 * 64, 128 * 72}, 128, false, false, 128, 64, 128, 72});
  cases.push_back({"M64  N72  A MN-major B MN-major lbo=ROWB sbo=128 (swapped)", {64, 72, 16, 1, 1, (uint32_t)ROWB, 128, 128, (uint32_t)ROWB, 128, 128, 128 * 64, 128 * 72}, 128, false, false, 128, 64, 128, 72});
  // 4. M = 64, K-major both (accumulator layout check)
  cases.push_back({"M64  N64  A K-major  B K-major", {64, 64, 8, 0, 0, 1024, 128, 2048, 1024, 128, 2048, 64 * 64, 64 * 64}, 64, true, true, 64, 64, 64, 64});
  // 5. M = 128 with MN-major A and B (dW with zero-padded M)
  cases.push_back({"M128 N72  A MN-major B MN-major lbo=128 sbo=ROWB (128-feature A)", {128, 72, 16, 1, 1, 128, (uint32_t)ROWB, 128, 128, (uint32_t)ROWB, 128, 128 * 128, 128 * 72}, 128, false, false, 128, 128, 128, 72});

  // 6. dW through K-major operands: A = dZ^T (M = 64 feats, K = 128 rows) stored [r/4][65 rows][4] (padded chunk
  //    stride -> conflict-free transposed stores), B = X^T (N = 72 feats) stored [r/4][73][4]
  cases.push_back({"M64  N72  K-major transposed tiles, padded LBO 1040/1168 (dW, planned)", {64, 72, 16, 0, 0, 65 * 16, 128, 2 * 65 * 16, 73 * 16, 128, 2 * 73 * 16, 32 * 65 * 4, 32 * 73 * 4}, 128, true, true, 65, 128, 73, 128});
  float *dA, *dB, *dD;
  cudaMalloc(&dA, 1 << 20); cudaMalloc(&dB, 1 << 20); cudaMalloc(&dD, 128 * 256 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (auto& nc : cases) {
    const Case& c = nc.c;
    // logical operands: Aop[M][Kc], Bop[N][Kc]
    const int Kc = nc.Ka;
    std::vector<float> Aop((size_t)c.M * Kc), Bop((size_t)c.N * Kc);
    srand(7);
    for (auto& v : Aop) v = tf32r((float)rand() / RAND_MAX - 0.5f);
    for (auto& v : Bop) v = tf32r((float)rand() / RAND_MAX - 0.5f);
    // physical matrices X[R][F] in "rows x features" form
    std::vector<float> XA((size_t)nc.Ra * nc.Fa, 0.f), XB((size_t)nc.Rb * nc.Fb, 0.f);
    if (nc.a_feat_is_k) { for (int m = 0; m < c.M; ++m) for (int k = 0; k < Kc; ++k) XA[(size_t)m * nc.Fa + k] = Aop[(size_t)m * Kc + k]; }
    else                { for (int m = 0; m < c.M; ++m) for (int k = 0; k < Kc; ++k) XA[(size_t)k * nc.Fa + m] = Aop[(size_t)m * Kc + k]; }
    if (nc.b_feat_is_k) { for (int n = 0; n < c.N; ++n) for (int k = 0; k < Kc; ++k) XB[(size_t)n * nc.Fb + k] = Bop[(size_t)n * Kc + k]; }
    else                { for (int n = 0; n < c.N; ++n) for (int k = 0; k < Kc; ++k) XB[(size_t)k * nc.Fb + n] = Bop[(size_t)n * Kc + k]; }
    auto ia = chunked(XA, nc.Ra, nc.Fa), ib = chunked(XB, nc.Rb, nc.Fb);
    cudaMemcpy(dA, ia.data(), ia.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, ib.data(), ib.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 256 * 4);
    Case cc = c; cc.a_floats = (int)ia.size(); cc.b_floats = (int)ib.size();
    probe<<<1, 128, (ia.size() + ib.size()) * 4 + 1024>>>(cc, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-70s CUDA ERROR %s\n", nc.name, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * 256);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    // reference and error under the two candidate accumulator layouts
    double err_plain = 0, err_16x4 = 0, maxref = 0, nz = 0;
    for (int m = 0; m < c.M; ++m) for (int n = 0; n < c.N; ++n) {
      double r = 0; for (int k = 0; k < Kc; ++k) r += (double)Aop[(size_t)m * Kc + k] * Bop[(size_t)n * Kc + k];
      maxref = fmax(maxref, fabs(r));
      err_plain = fmax(err_plain, fabs(D[(size_t)m * 256 + n] - r));
      const int lane = (m % 16) + 32 * (m / 16);
      if (c.M == 64) err_16x4 = fmax(err_16x4, fabs(D[(size_t)lane * 256 + n] - r));
    }
    for (float v : D) nz += v != 0.f;
    printf("%-70s max|ref| %.3f  err(lane=m) %.3e  err(lane=16x4) %.3e  nonzero %d\n", nc.name, maxref, err_plain, err_16x4, (int)nz);
  }
  return 0;
}
