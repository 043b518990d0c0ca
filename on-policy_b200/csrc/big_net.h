// big_net.h -- host-side declarations of the hidden >= 128 MLP path (big_net.cu orchestrates, big_gemm.cu holds the
// tcgen05 kernels, big_ref.cu the exact-fp32 FFMA kernels running the same epilogues).
#pragma once
#include "big_epi.cuh"

namespace mappo {
namespace big {

struct LinShape {
  int n_rows, K, N, BN, store_out, n_rowblocks, n_stages;
  int direct;                           // epilogue tiles bypass shared memory: result registers -> global, saved activation global -> registers
  float* out_ptr; long long out_ld;     // (direct) result matrix
  const float* ain_ptr; long long ain_ld;   // (direct) saved activation read by EpiBwd
};
struct LinOperands {
  const float* A; long long lda;        // [rows][K]   activations / gradients, K-major
  const float* W; long long ldw;        // [N][K]      packed weights, K-major
  float* out; long long ldo;            // [rows][N]   epilogue output (when the epilogue stores a tile)
  const float* ain; long long ldain;    // [rows][N]   stored activation read by EpiBwd
  int sm_count;
};
struct GradShape {
  int rows, rows_per_split, splits, m_tiles, n_tiles;
  int M;                                // valid output rows (columns of P)
  int Pw, Qw, ldq;                      // widths of P and Q (multiples of 32); ldq = leading dimension of a partial row
  int q0[4], qw[4];                     // column range of every Q tile (qw <= 320, multiple of 32)
};

// tcgen05 path (big_gemm.cu)
int lin_fwd_launch(const LinOperands&, const EpiFwd::Args&, const LinShape&, cudaStream_t);
int lin_bwd_launch(const LinOperands&, const EpiBwd::Args&, const LinShape&, cudaStream_t);
int lin_head_launch(const LinOperands&, const EpiHead::Args&, const LinShape&, cudaStream_t);
int lin_sample_launch(const LinOperands&, const EpiSample::Args&, const LinShape&, cudaStream_t);
int grad_gemm_launch(const float* P, int ldp, const float* Q, int ldq_in, float* partial, GradShape sh, cudaStream_t st);
int grad_gemm_pair_launch(const float* P, int ldp, const float* Q, int ldq_in, float* partial, GradShape sh, cudaStream_t st);   // big_grad_pair.cu

// exact fp32 path (big_ref.cu): same contracts, FFMA main loops; `scratch` holds one [rows][N] accumulator matrix
int ref_lin_fwd_launch(const LinOperands&, const EpiFwd::Args&, const LinShape&, float* scratch, cudaStream_t);
int ref_lin_bwd_launch(const LinOperands&, const EpiBwd::Args&, const LinShape&, float* scratch, cudaStream_t);
int ref_lin_head_launch(const LinOperands&, const EpiHead::Args&, const LinShape&, float* scratch, cudaStream_t);
int ref_lin_sample_launch(const LinOperands&, const EpiSample::Args&, const LinShape&, float* scratch, cudaStream_t);
int ref_grad_gemm_launch(const float* P, int ldp, const float* Q, int ldq_in, float* partial, GradShape sh, cudaStream_t st);

}  // namespace big
}  // namespace mappo
