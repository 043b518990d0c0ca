// rng.cuh -- counter-based random numbers shared by the rollout sampler and the device environments.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mappo {

// Philox4x32-10 (Salmon et al. 2011), counter-based: no state to keep between launches.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

}  // namespace mappo
