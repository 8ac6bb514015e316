// Philox4x32-10 counter RNG + lean FP64 Box–Muller: the in-kernel noise source of the stochastic path.
// Stream definition: include/quantrl_b200.h ("RNG stream"); CPU restatement: oracle/em_oracle.py:philox_normals.
// Replaces the per-episode host draws of the reference (quantrl/envs/stochastic.py:162-170, quantrl/envs/base.py:408-424),
// whose O(shape_T * E * n) noise tensors never exist here.
//
// The generator is written for issue-slot economy (ncu r1a/r1b: the libm-based version spent ~350 issue slots per
// pair of normals, 2/3 of them integer/uniform-move overhead):
//   * the ten round keys are precomputed on the host and live in the kernel-parameter constant bank, so a round is
//     2 IMAD.WIDE + 2 LOP3;
//   * ln, sqrt and sin/cos are branch-free polynomial kernels (fdlibm's minimax coefficients, |err| < 2e-16) whose
//     coefficients sit in __constant__ memory and are consumed as constant-bank operands of DFMA;
//   * 1/x and 1/sqrt(x) start from the hardware approximations (MUFU.RCP64H / MUFU.RSQ64H) + Newton steps;
//   * the angle uses 49 bits for alpha in [0, pi/4) and 3 bits for the octant symmetry (swap / two sign flips), which
//     is measure-preserving on the circle and needs no range reduction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qrl_r1 {

struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};

inline PhiloxKeys make_philox_keys(uint64_t seed) {
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = a;
        k.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.k0[r], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.k1[r], (uint32_t)p0);
    }
    return c;
}

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

// polynomial coefficients (fdlibm e_log.c / k_sin.c / k_cos.c)
static __constant__ double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                     2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                     1.479819860511658591e-01};
static __constant__ double kSin[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                      2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double kCos[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                      -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
static __constant__ double kMisc[2] = {-1.3862943611198906188,   // -2 ln 2
                                       0.78539816339744830962};  // pi / 4

// Two independent N(0,1) from 128 random bits (x.y:x.x -> radius, x.w:x.z -> direction).
__device__ __forceinline__ void normal_pair(const uint4 x, double& z0, double& z1) {
    // ---- u1 in (0,1): d1 = bits(0x3FF<<52 | x1<<20 | x0>>12 | 1) in (1,2), u1 = 2 - d1 ----
    const double d1 = __hiloint2double((int)(0x3FF00000u | (x.y >> 12)), (int)((x.y << 20) | (x.x >> 12) | 1u));
    const double u1 = 2.0 - d1;
    // ---- a = -2 ln(u1): u1 = 2^k * m, m in [sqrt(1/2), sqrt(2)); ln m = f - f^2/2 + s (f^2/2 + R(s^2)), s = f/(2+f) ----
    int hx = __double2hiint(u1);
    int k = (hx >> 20) - 1023;
    hx &= 0x000FFFFF;
    const int i = (hx + 0x95F64) & 0x100000;
    hx |= (i ^ 0x3FF00000);
    k += (i >> 20);
    const double m = __hiloint2double(hx, __double2loint(u1));
    const double f = m - 1.0;
    const double den = 2.0 + f;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    double s = f * r;
    s = fma(fma(-den, s, f), r, s);
    const double zz = s * s;
    double R = kLg[6];
    R = fma(R, zz, kLg[5]);
    R = fma(R, zz, kLg[4]);
    R = fma(R, zz, kLg[3]);
    R = fma(R, zz, kLg[2]);
    R = fma(R, zz, kLg[1]);
    R = fma(R, zz, kLg[0]);
    R *= zz;
    const double hfsq = 0.5 * f * f;
    const double lnm = f - (hfsq - s * (hfsq + R));
    const double a = fma((double)k, kMisc[0], -2.0 * lnm);   // > 0 because u1 < 1
    // ---- radius = sqrt(a): Goldschmidt from the hardware rsqrt approximation ----
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    // ---- direction: B = x3<<20 | x2>>12 (52 bits); alpha = (low 49 bits / 2^49) * pi/4; top 3 bits = symmetry ----
    const uint32_t mant_hi = (x.w << 3) >> 12;
    const uint32_t mant_lo = (x.w << 23) | ((x.z >> 9) & ~7u);
    const double phi = __hiloint2double((int)(0x3FF00000u | mant_hi), (int)mant_lo) - 1.0;
    const double al = phi * kMisc[1];
    const double z = al * al;
    double ps = kSin[5];
    ps = fma(ps, z, kSin[4]);
    ps = fma(ps, z, kSin[3]);
    ps = fma(ps, z, kSin[2]);
    ps = fma(ps, z, kSin[1]);
    ps = fma(ps, z, kSin[0]);
    const double sn = fma(al * z, ps, al);
    double pc = kCos[5];
    pc = fma(pc, z, kCos[4]);
    pc = fma(pc, z, kCos[3]);
    pc = fma(pc, z, kCos[2]);
    pc = fma(pc, z, kCos[1]);
    pc = fma(pc, z, kCos[0]);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    const bool swap = (x.w >> 29) & 1u;                 // bit 49 of B
    double s0 = swap ? cs : sn, c0 = swap ? sn : cs;
    // bit 50 flips the sign of the first component, bit 51 of the second
    s0 = __hiloint2double(__double2hiint(s0) ^ (int)((x.w << 1) & 0x80000000u), __double2loint(s0));
    c0 = __hiloint2double(__double2hiint(c0) ^ (int)(x.w & 0x80000000u), __double2loint(c0));
    z0 = g * s0;
    z1 = g * c0;
}

// Two N(0,1) for the counter (tau, episode, env_global, comp).
__device__ __forceinline__ void philox_normal2(const PhiloxKeys& keys, uint32_t tau, uint32_t episode, uint32_t env,
                                               uint32_t comp, double& z0, double& z1) {
    normal_pair(philox4x32_10(make_uint4(tau, episode, env, comp), keys), z0, z1);
}

__device__ __forceinline__ void philox_normal2(uint64_t seed, uint32_t tau, uint32_t episode, uint32_t env,
                                               uint32_t comp, double& z0, double& z1) {
    normal_pair(philox4x32_10(make_uint4(tau, episode, env, comp), (uint32_t)seed, (uint32_t)(seed >> 32)), z0, z1);
}

}  // namespace qrl_r1
