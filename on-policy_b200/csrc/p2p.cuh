// p2p.cuh -- peer-memory primitives shared by the stand-alone all-reduce (p2p_allreduce.cu) and the fused optimiser tail
// (update_mlp_tc.cu): signal stores / loads at system scope and L1-bypassing loads through NVLink.
#pragma once
#include "common.cuh"

namespace mappo {

constexpr int kMaxPeers = 8;
struct P2PArgs {
  const void* buf[kMaxPeers];
  uint32_t* sig[kMaxPeers];
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <typename T> __device__ __forceinline__ T ld_peer(const T* p);
template <> __device__ __forceinline__ float ld_peer<float>(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <> __device__ __forceinline__ double ld_peer<double>(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

}  // namespace mappo
