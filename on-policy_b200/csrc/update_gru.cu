// update_gru.cu -- training step for recurrent (GRU) actor / critic nets on chunked sequences.
//
// A recurrent minibatch is time-major [L, Nc] (position p = l*Nc + c): L = data_chunk_length steps of Nc chunks,
// each with the stored hidden state of its first row (utils/shared_buffer.py:557-604, rnn.py:30-79).
// The update runs as four launches per net; activations that cross a launch live in an HBM workspace laid out
// [position][H] (L2-resident at the BASELINE sizes):
//   1. update_mlp_kernel(feat_out)  base MLP forward of every position                  -> FEAT       (tile-parallel)
//   2. gru_seq_fwd_kernel           per 32-chunk tile, l = 0..L-1: h *= mask, GRU cell, LayerNorm, heads, loss,
//                                   head + LN backward                                   -> HM,R,Z,N,GHN,DHH
//   3. gru_seq_bwd_kernel           per 32-chunk tile, l = L-1..0: BPTT through the cell, dW_ih / dW_hh
//                                                                                        -> DFEAT
//   4. update_mlp_kernel(dfeat_in)  base MLP backward (forward recomputed in shared memory)
// Equivalent to RNNLayer.forward's segment loop (rnn.py:43-77): inside a segment all masks are 1 and a segment
// starts where some mask is 0 -- i.e. per step h <- h * mask_t before the cell.
#include "net_tiles.cuh"

namespace mappo {

constexpr int kSeqTR = 32;

int update_mlp_launch(const NetDev&, const float*, const BatchDev&, const LossDev&, const double*, const double*,
                      const float*, float*, int, double*, cudaStream_t, float* feat_out, const float* dfeat_in);
int update_mlp_slots(const NetDev& n, int n_rows, int sm_count);

// workspace arrays, each [n_rows][H]
enum { WS_FEAT = 0, WS_HM, WS_R, WS_Z, WS_N, WS_GHN, WS_DHH, WS_DFEAT, WS_COUNT };

struct GruW { int wih, whh, bih, bhh, rln_w, rln_b, head_w, head_b, ldh, total; };

__host__ __device__ inline GruW make_gru_w(const NetDev& n, bool with_head) {
  GruW s;
  const int H = n.hid;
  int o = 0;
  s.ldh = H | 1;
  s.wih = o; o += 3 * H * s.ldh;
  s.whh = o; o += 3 * H * s.ldh;
  s.bih = o; o += 3 * H;
  s.bhh = o; o += 3 * H;
  s.rln_w = o; o += with_head ? H : 0;
  s.rln_b = o; o += with_head ? H : 0;
  s.head_w = o; o += with_head ? n.head_total * s.ldh : 0;
  s.head_b = o; o += with_head ? n.head_total : 0;
  s.total = (o + 3) & ~3;
  return s;
}

__device__ inline void load_gru_w(float* sW, const GruW& s, const NetDev& n, const float* __restrict__ p,
                                  bool with_head, int tid, int nt) {
  const int H = n.hid;
  copy_mat(sW + s.wih, s.ldh, p + n.g.gru_wih, 3 * H, H, tid, nt);
  copy_mat(sW + s.whh, s.ldh, p + n.g.gru_whh, 3 * H, H, tid, nt);
  copy_vec(sW + s.bih, p + n.g.gru_bih, 3 * H, tid, nt);
  copy_vec(sW + s.bhh, p + n.g.gru_bhh, 3 * H, tid, nt);
  if (with_head) {
    copy_vec(sW + s.rln_w, p + n.g.rnn_ln_w, H, tid, nt);
    copy_vec(sW + s.rln_b, p + n.g.rnn_ln_b, H, tid, nt);
    copy_mat(sW + s.head_w, s.ldh, p + n.g.head_w, n.head_total, H, tid, nt);
    copy_vec(sW + s.head_b, p + n.g.head_b, n.head_total, tid, nt);
  }
  cp_async_wait_all();
}

// workspace [p][H] <-> transposed tile
template <int TR>
__device__ __forceinline__ void ws_load_T(const float* __restrict__ ws, size_t p0, int nvalid, int H,
                                          float* __restrict__ T, int tid) {
  constexpr int LD = Tile<TR>::LD;
  for (int i = tid; i < TR * H; i += Tile<TR>::NT) {
    const int r = i / H, c = i - r * H;
    T[c * LD + r] = r < nvalid ? ws[(p0 + r) * H + c] : 0.f;
  }
}
template <int TR>
__device__ __forceinline__ void ws_store_T(float* __restrict__ ws, size_t p0, int nvalid, int H,
                                           const float* __restrict__ T, int tid) {
  constexpr int LD = Tile<TR>::LD;
  for (int i = tid; i < TR * H; i += Tile<TR>::NT) {
    const int r = i / H, c = i - r * H;
    if (r < nvalid) ws[(p0 + r) * H + c] = T[c * LD + r];
  }
}

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// -------------------------------------------------------------------------------------------------
// 2. sequence forward + heads + loss
// -------------------------------------------------------------------------------------------------
template <int NJH>
__global__ void __launch_bounds__(4 * kSeqTR, 1)
gru_seq_fwd_kernel(const NetDev n, const float* __restrict__ params, const BatchDev b, const LossDev L,
                   const double* __restrict__ norm_stats, const double* __restrict__ adv_stats,
                   const float* __restrict__ vn_state, float* __restrict__ grad_part, double* __restrict__ loss_out,
                   float* __restrict__ ws, int n_seq_tiles) {
  constexpr int TR = kSeqTR;
  constexpr int LD = Tile<TR>::LD;
  constexpr int NT = Tile<TR>::NT;
  constexpr int NI = 32 / Tile<TR>::NTY;             // head gradient: No <= 32
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int H = n.hid, Atot = n.head_total, Nc = b.n_seq, Lsteps = b.seq_len;
  const GruW s = make_gru_w(n, true);
  const int hT_sz = H * LD;
  float* sW = smem;
  float* featT = sW + s.total;                       // later: LayerNorm output y
  float* hT = featT + hT_sz;
  float* gi = hT + hT_sz;                            // 3 tiles; gi[0] later holds dy / dh_head
  float* gh = gi + 3 * hT_sz;                        // 3 tiles
  float* hnew = gh + 3 * hT_sz;
  float* lgT = hnew + hT_sz;
  float* mean = lgT + ((Atot + 3) & ~3) * LD;
  float* rstd = mean + TR;
  float* red = rstd + TR;                            // 8*TR
  float* maskr = red + 8 * TR;                       // TR
  int* rowid = reinterpret_cast<int*>(maskr + TR);   // TR
  __shared__ double sred[2 * 32];

  load_gru_w(sW, s, n, params, true, tid, NT);
  float* g = b.eval_only ? nullptr : grad_part + (size_t)blockIdx.x * n.g.total;
  const LossConsts lc = make_loss_consts(n, L, norm_stats, adv_stats, vn_state);
  const size_t plane = (size_t)b.n_rows * H;
  const float* h0 = n.is_critic ? b.h0_critic : b.h0_actor;
  double acc[3] = {0.0, 0.0, 0.0};

  for (int st = blockIdx.x; st < n_seq_tiles; st += gridDim.x) {
    const int c0 = st * TR;
    const int nvalid = min(TR, Nc - c0);
    __syncthreads();
    // initial hidden state of each chunk (shared_buffer.py:568-569)
    for (int i = tid; i < TR * H; i += NT) {
      const int r = i / H, c = i - r * H;
      float v = 0.f;
      if (r < nvalid) {
        const int src = b.seq_first ? b.seq_first[c0 + r] : (c0 + r);
        v = h0[(size_t)src * H + c];
      }
      hT[c * LD + r] = v;
    }
    float* hcur = hT;
    float* hnxt = hnew;
    for (int l = 0; l < Lsteps; ++l) {
      const size_t p0 = (size_t)l * Nc + c0;
      __syncthreads();
      if (tid < TR) {
        const int gr = tid < nvalid ? (b.rows ? b.rows[p0 + tid] : (int)(p0 + tid)) : -1;
        rowid[tid] = gr;
        maskr[tid] = gr >= 0 ? b.masks[gr] : 0.f;
      }
      ws_load_T<TR>(ws + WS_FEAT * plane, p0, nvalid, H, featT, tid);
      __syncthreads();
      for (int i = tid; i < TR * H; i += NT) {                 // h <- h * mask (rnn.py:27, :67)
        const int c = i / TR, r = i - c * TR;
        hcur[c * LD + r] *= maskr[r];
      }
      __syncthreads();
      if (!b.eval_only) ws_store_T<TR>(ws + WS_HM * plane, p0, nvalid, H, hcur, tid);
#pragma unroll 1
      for (int gate = 0; gate < 3; ++gate) {
        tile_mm<TR, NJH>(featT, H, sW + s.wih + gate * H * s.ldh, s.ldh, 1, H, sW + s.bih + gate * H, ACT_NONE,
                         gi + gate * hT_sz, tid);
        tile_mm<TR, NJH>(hcur, H, sW + s.whh + gate * H * s.ldh, s.ldh, 1, H, sW + s.bhh + gate * H, ACT_NONE,
                         gh + gate * hT_sz, tid);
      }
      __syncthreads();
      for (int i = tid; i < TR * H; i += NT) {                 // torch GRU cell, gate order (r, z, n)
        const int c = i / TR, r = i - c * TR;
        const int o = c * LD + r;
        const float rg = sigm(gi[o] + gh[o]);
        const float zg = sigm(gi[hT_sz + o] + gh[hT_sz + o]);
        const float ghn = gh[2 * hT_sz + o];
        const float ng = tanhf(gi[2 * hT_sz + o] + rg * ghn);
        hnxt[o] = (1.f - zg) * ng + zg * hcur[o];
        gi[o] = rg; gi[hT_sz + o] = zg; gi[2 * hT_sz + o] = ng;   // keep (r, z, n) for the workspace
      }
      __syncthreads();
      if (!b.eval_only) {
        ws_store_T<TR>(ws + WS_R * plane, p0, nvalid, H, gi, tid);
        ws_store_T<TR>(ws + WS_Z * plane, p0, nvalid, H, gi + hT_sz, tid);
        ws_store_T<TR>(ws + WS_N * plane, p0, nvalid, H, gi + 2 * hT_sz, tid);
        ws_store_T<TR>(ws + WS_GHN * plane, p0, nvalid, H, gh + 2 * hT_sz, tid);
      }
      // LayerNorm of the new state (rnn.py:79) -> y (reuses featT), heads, loss
      float* y = featT;
      tile_layernorm<TR>(hnxt, H, sW + s.rln_w, sW + s.rln_b, y, mean, rstd, red, tid);
      tile_mm<TR, 2>(y, H, sW + s.head_w, s.ldh, 1, Atot, sW + s.head_b, ACT_NONE, lgT, tid);
      __syncthreads();
      if (tid < TR) row_loss<LD>(n, b, L, lc, lgT, tid, rowid[tid], (int)(p0 + tid), acc);
      __syncthreads();
      if (!b.eval_only) {
        float* dy = gi;                                        // (r,z,n) already stored
        tile_colsum<TR>(lgT, Atot, g + n.g.head_b, tid);
        tile_dw<TR, NI, NJH>(lgT, Atot, y, H, g + n.g.head_w, H, tid);
        tile_mm<TR, NJH>(lgT, Atot, sW + s.head_w, 1, s.ldh, H, nullptr, ACT_NONE, dy, tid);
        __syncthreads();
        tile_ln_param_grads<TR>(dy, hnxt, mean, rstd, H, g + n.g.rnn_ln_w, g + n.g.rnn_ln_b, tid);
        tile_layernorm_bwd<TR>(dy, hnxt, mean, rstd, sW + s.rln_w, H, ACT_NONE, red, tid);
        ws_store_T<TR>(ws + WS_DHH * plane, p0, nvalid, H, dy, tid);
      }
      float* tmp = hcur; hcur = hnxt; hnxt = tmp;
    }
  }
  __syncthreads();
  if (n.is_critic) {
    double one[1] = {acc[0]};
    block_accumulate<1>(one, loss_out + 0, sred, tid, NT);
  } else {
    double two[2] = {acc[0], acc[1]};
    block_accumulate<2>(two, loss_out + 1, sred, tid, NT);
    double rt[1] = {acc[2] / (lc.n_rows_d * (double)b.act_shape)};
    block_accumulate<1>(rt, loss_out + 5, sred, tid, NT);
  }
}

// -------------------------------------------------------------------------------------------------
// 3. sequence backward (BPTT)
// -------------------------------------------------------------------------------------------------
template <int NJH>
__global__ void __launch_bounds__(4 * kSeqTR, 1)
gru_seq_bwd_kernel(const NetDev n, const float* __restrict__ params, const BatchDev b, float* __restrict__ grad_part,
                   float* __restrict__ ws, int n_seq_tiles) {
  constexpr int TR = kSeqTR;
  constexpr int LD = Tile<TR>::LD;
  constexpr int NT = Tile<TR>::NT;
  constexpr int NI = 16 * NJH / Tile<TR>::NTY;       // H rows of one gate per tile_dw call
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int H = n.hid, Nc = b.n_seq, Lsteps = b.seq_len;
  const GruW s = make_gru_w(n, false);
  const int hT_sz = H * LD;
  float* sW = smem;
  float* dh = sW + s.total;
  float* dgi = dh + hT_sz;                           // 3 tiles: (r, z, n) on load, (dr, dz, dn) pre-activations after
  float* dgh = dgi + 3 * hT_sz;                      // 3 tiles
  float* ghn = dgh + 3 * hT_sz;
  float* hm = ghn + hT_sz;
  float* feat = hm + hT_sz;
  float* dfe = feat + hT_sz;
  float* maskr = dfe + hT_sz;
  float* g = grad_part + (size_t)blockIdx.x * n.g.total;
  const size_t plane = (size_t)b.n_rows * H;

  load_gru_w(sW, s, n, params, false, tid, NT);

  for (int st = blockIdx.x; st < n_seq_tiles; st += gridDim.x) {
    const int c0 = st * TR;
    const int nvalid = min(TR, Nc - c0);
    __syncthreads();
    for (int i = tid; i < hT_sz; i += NT) dh[i] = 0.f;
    for (int l = Lsteps - 1; l >= 0; --l) {
      const size_t p0 = (size_t)l * Nc + c0;
      __syncthreads();
      if (tid < TR) {
        const int gr = tid < nvalid ? (b.rows ? b.rows[p0 + tid] : (int)(p0 + tid)) : -1;
        maskr[tid] = gr >= 0 ? b.masks[gr] : 0.f;
      }
      ws_load_T<TR>(ws + WS_R * plane, p0, nvalid, H, dgi, tid);
      ws_load_T<TR>(ws + WS_Z * plane, p0, nvalid, H, dgi + hT_sz, tid);
      ws_load_T<TR>(ws + WS_N * plane, p0, nvalid, H, dgi + 2 * hT_sz, tid);
      ws_load_T<TR>(ws + WS_GHN * plane, p0, nvalid, H, ghn, tid);
      ws_load_T<TR>(ws + WS_HM * plane, p0, nvalid, H, hm, tid);
      ws_load_T<TR>(ws + WS_FEAT * plane, p0, nvalid, H, feat, tid);
      ws_load_T<TR>(ws + WS_DHH * plane, p0, nvalid, H, dfe, tid);          // dfe used as staging for dh_head
      __syncthreads();
      for (int i = tid; i < TR * H; i += NT) {
        const int c = i / TR, r = i - c * TR;
        const int o = c * LD + r;
        const float d = dh[o] + dfe[o];                                      // dL/dh_l: future steps + head path
        const float rg = dgi[o], zg = dgi[hT_sz + o], ng = dgi[2 * hT_sz + o];
        const float dn_pre = d * (1.f - zg) * (1.f - ng * ng);
        const float dz_pre = d * (hm[o] - ng) * zg * (1.f - zg);
        const float dr_pre = dn_pre * ghn[o] * rg * (1.f - rg);
        dgi[o] = dr_pre; dgi[hT_sz + o] = dz_pre; dgi[2 * hT_sz + o] = dn_pre;
        dgh[o] = dr_pre; dgh[hT_sz + o] = dz_pre; dgh[2 * hT_sz + o] = dn_pre * rg;
        dh[o] = d * zg;                                                      // direct path h' = ... + z * hm
      }
      __syncthreads();
      tile_colsum<TR>(dgi, 3 * H, g + n.g.gru_bih, tid);
      tile_colsum<TR>(dgh, 3 * H, g + n.g.gru_bhh, tid);
#pragma unroll 1
      for (int gate = 0; gate < 3; ++gate) {
        tile_dw<TR, NI, NJH>(dgi + gate * hT_sz, H, feat, H, g + n.g.gru_wih + gate * H * H, H, tid);
        tile_dw<TR, NI, NJH>(dgh + gate * hT_sz, H, hm, H, g + n.g.gru_whh + gate * H * H, H, tid);
      }
      __syncthreads();                                                       // feat / dfe free from here
      tile_mm<TR, NJH>(dgi, 3 * H, sW + s.wih, 1, s.ldh, H, nullptr, ACT_NONE, dfe, tid);     // dL/dfeat_l
      tile_mm<TR, NJH>(dgh, 3 * H, sW + s.whh, 1, s.ldh, H, nullptr, ACT_NONE, feat, tid);    // dL/dhm via W_hh
      __syncthreads();
      ws_store_T<TR>(ws + WS_DFEAT * plane, p0, nvalid, H, dfe, tid);
      for (int i = tid; i < TR * H; i += NT) {
        const int c = i / TR, r = i - c * TR;
        const int o = c * LD + r;
        dh[o] = (dh[o] + feat[o]) * maskr[r];                                // hm = h_{l-1} * mask_l
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
static size_t fwd_smem_bytes(const NetDev& n) {
  const GruW s = make_gru_w(n, true);
  constexpr int LD = Tile<kSeqTR>::LD;
  size_t f = s.total + (size_t)9 * n.hid * LD + ((n.head_total + 3) & ~3) * LD + 2 * kSeqTR + 8 * kSeqTR + 2 * kSeqTR + 2;
  return f * sizeof(float);
}
static size_t bwd_smem_bytes(const NetDev& n) {
  const GruW s = make_gru_w(n, false);
  constexpr int LD = Tile<kSeqTR>::LD;
  return ((size_t)s.total + (size_t)11 * n.hid * LD + kSeqTR) * sizeof(float);
}

int update_gru_slots(const NetDev& n, int n_rows, int, int sm_count) { return update_mlp_slots(n, n_rows, sm_count); }

int64_t update_gru_workspace_floats(const NetDev& n, int n_rows) { return (int64_t)WS_COUNT * n_rows * n.hid; }

int update_gru_launch(const NetDev& n, const float* params, const BatchDev& b, const LossDev& L,
                      const double* norm_stats, const double* adv_stats, const float* vn_state, float* grad_part,
                      int n_slots, double* loss_out, float* ws, cudaStream_t st) {
  if (n.hid != 64) { set_error("update_gru: hidden_size %d not built (64 only)", n.hid); return MAPPO_ERR_UNSUPPORTED; }
  if (n.head_total > 32) { set_error("update_gru: sum(head_dim) > 32"); return MAPPO_ERR_UNSUPPORTED; }
  const float* h0 = n.is_critic ? b.h0_critic : b.h0_actor;
  if (!h0 || !b.masks) { set_error("update_gru: h0 / masks missing"); return MAPPO_ERR_INVALID; }
  const size_t plane = (size_t)b.n_rows * n.hid;
  // 1. base forward of every position (also zeroes every gradient slot)
  int rc = update_mlp_launch(n, params, b, L, norm_stats, adv_stats, vn_state, grad_part, n_slots, loss_out, st,
                             ws + WS_FEAT * plane, nullptr);
  if (rc) return rc;
  const int n_seq_tiles = (b.n_seq + kSeqTR - 1) / kSeqTR;
  const int grid = n_seq_tiles < n_slots ? n_seq_tiles : n_slots;
  {
    const size_t bytes = fwd_smem_bytes(n);
    if (bytes > 227 * 1024) { set_error("gru_seq_fwd: %zu B shared memory > 227 KB", bytes); return MAPPO_ERR_UNSUPPORTED; }
    auto kern = gru_seq_fwd_kernel<4>;
    static thread_local SmemConfig configured_dev = {};
  size_t& configured = configured_dev.slot();
    if (bytes > configured) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        return check_launch("gru_seq_fwd: cudaFuncSetAttribute");
      configured = bytes;
    }
    kern<<<grid, 4 * kSeqTR, bytes, st>>>(n, params, b, L, norm_stats, adv_stats, vn_state, grad_part, loss_out, ws,
                                          n_seq_tiles);
    rc = check_launch("gru_seq_fwd_kernel");
    if (rc) return rc;
  }
  if (b.eval_only) return MAPPO_OK;
  {
    const size_t bytes = bwd_smem_bytes(n);
    if (bytes > 227 * 1024) { set_error("gru_seq_bwd: %zu B shared memory > 227 KB", bytes); return MAPPO_ERR_UNSUPPORTED; }
    auto kern = gru_seq_bwd_kernel<4>;
    static thread_local SmemConfig configured_dev = {};
  size_t& configured = configured_dev.slot();
    if (bytes > configured) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        return check_launch("gru_seq_bwd: cudaFuncSetAttribute");
      configured = bytes;
    }
    kern<<<grid, 4 * kSeqTR, bytes, st>>>(n, params, b, grad_part, ws, n_seq_tiles);
    rc = check_launch("gru_seq_bwd_kernel");
    if (rc) return rc;
  }
  // 4. base backward with dL/dfeat from the sequence pass
  return update_mlp_launch(n, params, b, L, norm_stats, adv_stats, vn_state, grad_part, n_slots, loss_out, st, nullptr,
                           ws + WS_DFEAT * plane);
}

}  // namespace mappo
