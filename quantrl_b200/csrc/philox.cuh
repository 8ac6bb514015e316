// Philox4x32-10 counter RNG + FP64 Box–Muller: the in-kernel noise source of the stochastic path.
// Stream definition: include/quantrl_b200.h ("RNG stream"); CPU restatement: oracle/em_oracle.py:philox_normals.
// Replaces the per-episode host draws of the reference (quantrl/envs/stochastic.py:162-170, quantrl/envs/base.py:408-424),
// whose O(shape_T * E * n) noise tensors never exist here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qrl {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k0, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k1, (uint32_t)p0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// 52 random mantissa bits -> double in [1,2)
__device__ __forceinline__ double unit_12(uint32_t hi, uint32_t lo) {
    return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
}

// Two independent N(0,1) for (tau, episode, env_global, comp).
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint32_t tau, uint32_t episode, uint32_t env,
                                               uint32_t comp, double& z0, double& z1) {
    const uint4 x = philox4x32_10(make_uint4(tau, episode, env, comp), (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = 2.0 - unit_12(x.y, x.x);   // (0,1]
    const double u2 = unit_12(x.w, x.z) - 1.0;   // [0,1)
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = r * s;
    z1 = r * c;
}

}  // namespace qrl
