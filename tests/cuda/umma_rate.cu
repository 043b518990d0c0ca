// umma_rate.cu -- raw issue / execution rate of tcgen05.mma on the B200, operands resident in shared memory (no TMA in the loop).
// Answers one question for profiles/r2_summary.md: how many cycles does ONE 128 x N x 8 kind::tf32 MMA take when issued back to
// back, for the operand layouts the GEMM kernels use -- i.e. is ~450 TFLOP/s (what big_lin / big_grad reach) the instruction's
// own ceiling or the pipeline's?   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a umma_rate.cu -o umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)(layout & 7u) << 61);
}
// kind: 0 tf32 (A/B format 2), 1 bf16 (format 1); D fp32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Case { int N, mn, kind, iters; };

template <int KIND>
__global__ void __launch_bounds__(128) rate_kernel(Case c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 49152 * 3 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;      // 3 stages x (16 KB A + 32 KB B)
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, c.N, c.mn, c.mn, KIND == 0 ? 2 : 1);
    const long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      const uint32_t a0 = smem_u32(smem + (it % 3) * 49152), b0 = a0 + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t ad, bd;
        if (c.mn == 0) { ad = make_desc(a0 + k * 32, 16, 1024, 2); bd = make_desc(b0 + k * 32, 16, 1024, 2); }     // K-major, SWIZZLE_128B
        else { ad = make_desc(a0 + k * 1024, 4096, 512, 1); bd = make_desc(b0 + k * 1024, 4096, 512, 1); }          // MN-major, 128B / 32B atoms
        umma<KIND>(tmem + (uint32_t)((it & 1) * 256), ad, bd, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const size_t bytes = 3 * 49152 + 1024;
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  const Case cases[] = {{256, 0, 0, 2048}, {128, 0, 0, 2048}, {64, 0, 0, 2048}, {256, 1, 0, 2048}, {256, 0, 1, 2048}, {128, 0, 1, 2048}};
  for (int grid : {1, p.multiProcessorCount}) {
    for (const Case& c : cases) {
      for (int rep = 0; rep < 2; ++rep) {
        if (c.kind == 0) rate_kernel<0><<<grid, 128, bytes>>>(c, out); else rate_kernel<1><<<grid, 128, bytes>>>(c, out);
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2] = {0, 0};
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const double n = 4.0 * c.iters, kdim = c.kind == 0 ? 8 : 16;
      const double cyc = (double)h[1] / n, flop = 2.0 * 128 * c.N * kdim;
      printf("grid %3d  %s M128 N%-3d %s: %8.1f cycles / MMA (issue %6.1f)  -> %7.1f FLOP/clk/SM = %6.1f TFLOP/s on %d SMs at %.3f GHz  %s\n", grid,
             c.kind == 0 ? "tf32 K8 " : "bf16 K16", c.N, c.mn ? "MN-major" : "K-major ", cyc, (double)h[0] / n, flop / cyc,
             flop / cyc * p.multiProcessorCount * (p.clockRate * 1e-6) * 1e-3, p.multiProcessorCount, p.clockRate * 1e-6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
