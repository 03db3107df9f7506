// philox.cuh -- Philox4x32-10, identical to oracle/philox.py (counter = (pair, step, tensor_id, 0),
// key = seed lo/hi; uniforms ((bits>>8)+0.5)*2^-24).  Replaces TF's op-seeded Philox stream behind
// tfp RelaxedOneHotCategorical.sample() (reference sga.py:97,120), which is not reproducible outside TF.
#pragma once
#include <stdint.h>

namespace sga {

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float bits_to_uniform(uint32_t b) {
  return (static_cast<float>(b >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 2^-24
}

}  // namespace sga
