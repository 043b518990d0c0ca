// big_ref.cu -- exact-fp32 (FFMA) build of the hidden >= 128 MLP path: MAPPO_GEMM_FP32.
// Plain shared-memory-tiled SGEMMs write the raw accumulator matrix to a scratch buffer; a row kernel then runs the SAME
// epilogue functors (big_epi.cuh) as the tcgen05 kernels, one thread per row, 32 columns at a time, with tf32 rounding off.
// This is the tight-parity / debugging mode (all arithmetic fp32, sums only re-associated), not the fast path.
#include "big_net.h"

namespace mappo {
namespace big {

// C[r][n] = sum_k A[r][k] W[n][k]        (64 x 64 tile, 256 threads, 4 x 4 per thread)
__global__ void __launch_bounds__(256) ref_gemm_nt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W,
                                                          long long ldw, float* __restrict__ C, long long ldc, int rows, int N, int K) {
  __shared__ float sA[16][64 + 4], sW[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = tid; i < 64 * 16; i += 256) {
      const int rr = i >> 4, kk = i & 15;
      const int r = r0 + rr, n = n0 + rr, k = k0 + kk;
      sA[kk][rr] = (r < rows && k < K) ? A[(size_t)r * lda + k] : 0.f;
      sW[kk][rr] = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&sW[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[(size_t)r * ldc + n] = acc[i][j];
    }
  }
}

// partial[split][m][q] = sum_{r in split} P[r][m] Q[r][q]
__global__ void __launch_bounds__(256) ref_gemm_tn_kernel(const float* __restrict__ P, long long ldp, const float* __restrict__ Q,
                                                          long long ldq_in, float* __restrict__ partial, GradShape sh) {
  __shared__ float sP[16][64 + 4], sQ[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, q0 = blockIdx.x * 64, split = blockIdx.z;
  const int ra = split * sh.rows_per_split, rb = min(sh.rows, ra + sh.rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = ra; k0 < rb; k0 += 16) {
    for (int i = tid; i < 64 * 16; i += 256) {
      const int kk = i >> 6, cc = i & 63;
      const int r = k0 + kk;
      sP[kk][cc] = (r < rb && m0 + cc < sh.Pw) ? P[(size_t)r * ldp + m0 + cc] : 0.f;
      sQ[kk][cc] = (r < rb && q0 + cc < sh.Qw) ? Q[(size_t)r * ldq_in + q0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sP[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&sQ[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= sh.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + tx * 4 + j;
      if (q < sh.Qw) partial[((size_t)split * sh.M + m) * sh.ldq + q] = acc[i][j];
    }
  }
}

// one thread per row: accumulator chunks from the scratch matrix through the epilogue functor
template <class Epi>
__global__ void __launch_bounds__(kTileRows) ref_epilogue_kernel(const float* __restrict__ C, long long ldc, const float* __restrict__ ain_m,
                                                                 long long ldain, float* __restrict__ out_m, long long ldo,
                                                                 const typename Epi::Args ea, const LinShape sh) {
  extern __shared__ float sm[];
  __shared__ double sred[2 * 32];
  float* cv = sm;
  float* scratch = sm + 2 * sh.N;
  const int r = threadIdx.x, grow = blockIdx.x * kTileRows + r;
  for (int i = r; i < 2 * sh.N; i += kTileRows) cv[i] = ea.colvec[i];
  __syncthreads();
  typename Epi::Thread th;
  Epi::init_thread(th);
  typename Epi::Row row;
  Epi::begin_row(ea, row, grow);
  for (int col0 = 0; col0 < sh.N; col0 += kChunk) {
    float acc[kChunk], ain[kChunk], out[kChunk];
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      acc[j] = grow < sh.n_rows ? C[(size_t)grow * ldc + col0 + j] : 0.f;
      ain[j] = (Epi::kHasAin && grow < sh.n_rows) ? ain_m[(size_t)grow * ldain + col0 + j] : 0.f;
    }
    Epi::chunk(ea, th, row, acc, ain, out, col0, cv, scratch, r, grow);
    if (Epi::kStoresOut && sh.store_out && grow < sh.n_rows) {
#pragma unroll
      for (int j = 0; j < kChunk; ++j) out_m[(size_t)grow * ldo + col0 + j] = out[j];
    }
  }
  Epi::end_row(ea, row, grow);
  __syncthreads();
  Epi::finish_thread(ea, th, sred, r, kTileRows);
}

template <class Epi>
static int ref_lin_t(const LinOperands& o, const typename Epi::Args& ea, LinShape sh, float* scratch, const char* name, cudaStream_t st) {
  if (!scratch) { set_error("%s: fp32 path needs the accumulator scratch", name); return MAPPO_ERR_INVALID; }
  sh.n_rowblocks = (sh.n_rows + kTileRows - 1) / kTileRows;
  ref_gemm_nt_kernel<<<dim3((sh.N + 63) / 64, (sh.n_rows + 63) / 64), 256, 0, st>>>(o.A, o.lda, o.W, o.ldw, scratch, sh.N, sh.n_rows, sh.N, sh.K);
  int rc = check_launch("ref_gemm_nt_kernel");
  if (rc) return rc;
  const size_t bytes = (size_t)(2 * sh.N + (Epi::kNeedsScratch ? 32 * kLgLd : 0)) * sizeof(float);
  ref_epilogue_kernel<Epi><<<sh.n_rowblocks, kTileRows, bytes, st>>>(scratch, sh.N, o.ain, o.ldain, o.out, o.ldo, ea, sh);
  return check_launch(name);
}

int ref_lin_fwd_launch(const LinOperands& o, const EpiFwd::Args& ea, const LinShape& sh, float* s, cudaStream_t st) { return ref_lin_t<EpiFwd>(o, ea, sh, s, "ref_epilogue_kernel<EpiFwd>", st); }
int ref_lin_bwd_launch(const LinOperands& o, const EpiBwd::Args& ea, const LinShape& sh, float* s, cudaStream_t st) { return ref_lin_t<EpiBwd>(o, ea, sh, s, "ref_epilogue_kernel<EpiBwd>", st); }
int ref_lin_head_launch(const LinOperands& o, const EpiHead::Args& ea, const LinShape& sh, float* s, cudaStream_t st) { return ref_lin_t<EpiHead>(o, ea, sh, s, "ref_epilogue_kernel<EpiHead>", st); }
int ref_lin_sample_launch(const LinOperands& o, const EpiSample::Args& ea, const LinShape& sh, float* s, cudaStream_t st) { return ref_lin_t<EpiSample>(o, ea, sh, s, "ref_epilogue_kernel<EpiSample>", st); }

int ref_grad_gemm_launch(const float* P, int ldp, const float* Q, int ldq_in, float* partial, GradShape sh, cudaStream_t st) {
  ref_gemm_tn_kernel<<<dim3((sh.Qw + 63) / 64, (sh.M + 63) / 64, sh.splits), 256, 0, st>>>(P, ldp, Q, ldq_in, partial, sh);
  return check_launch("ref_gemm_tn_kernel");
}

}  // namespace big
}  // namespace mappo
