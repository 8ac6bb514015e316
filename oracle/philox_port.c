/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the in-kernel noise generator, never linked into the product.
 *
 * Restates quantrl_b200/csrc/philox.cuh operation by operation: Philox4x32-10 (Salmon et al., SC'11; stream definition in
 * include/quantrl_b200.h, "RNG stream") and the lean FP64 Box–Muller with its polynomial ln / sin / cos kernels (fdlibm
 * coefficients) and Goldschmidt square root.  Two differences, both confined to the last rounding: the hardware seeds
 * rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64 are replaced by the correctly rounded 1/x and 1/sqrt(x) (the Newton steps
 * that follow are kept), so kernel and port agree to an ulp or two, not bit for bit.
 * Used by tests/test_oracle_c_port.py to bound the polynomial kernels against libm (oracle/em_oracle.py:philox_normals)
 * on the CPU — the accuracy claim of philox.cuh without a GPU.
 * Build: gcc -O2 -march=x86-64-v3 -ffp-contract=off -shared -fPIC oracle/philox_port.c -o oracle/_build/libphilox_port.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                              2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                              1.479819860511658591e-01};
static const double kSin[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                               2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static const double kCos[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                               -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
static const double kM2Ln2 = -1.3862943611198906188, kPi4 = 0.78539816339744830962;

static inline double hilo(uint32_t hi, uint32_t lo) {
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline uint32_t hi_of(double d) { uint64_t b; memcpy(&b, &d, 8); return (uint32_t)(b >> 32); }
static inline uint32_t lo_of(double d) { uint64_t b; memcpy(&b, &d, 8); return (uint32_t)b; }

void philox_port_4x32_10(const uint32_t c_in[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c[4] = {c_in[0], c_in[1], c_in[2], c_in[3]};
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

/* philox.cuh:normal_pair */
static void normal_pair(const uint32_t x[4], double* z0, double* z1) {
    const double d1 = hilo(0x3FF00000u | (x[1] >> 12), (x[1] << 20) | (x[0] >> 12) | 1u);
    const double u1 = 2.0 - d1;
    int32_t hx = (int32_t)hi_of(u1);
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000FFFFF;
    const int32_t i = (hx + 0x95F64) & 0x100000;
    hx |= (i ^ 0x3FF00000);
    k += (i >> 20);
    const double m = hilo((uint32_t)hx, lo_of(u1));
    const double f = m - 1.0;
    const double den = 2.0 + f;
    double r = 1.0 / den;
    r = fma(r, fma(-den, r, 1.0), r);
    double s = f * r;
    s = fma(fma(-den, s, f), r, s);
    const double zz = s * s;
    double R = kLg[6];
    for (int j = 5; j >= 0; --j) R = fma(R, zz, kLg[j]);
    R *= zz;
    const double hfsq = 0.5 * f * f;
    const double lnm = f - (hfsq - s * (hfsq + R));
    const double a = fma((double)k, kM2Ln2, -2.0 * lnm);
    double y = 1.0 / sqrt(a);
    double g = a * y, h = 0.5 * y;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    const uint32_t mant_hi = (x[3] << 3) >> 12;
    const uint32_t mant_lo = (x[3] << 23) | ((x[2] >> 9) & ~7u);
    const double phi = hilo(0x3FF00000u | mant_hi, mant_lo) - 1.0;
    const double al = phi * kPi4;
    const double z = al * al;
    double ps = kSin[5], pc = kCos[5];
    for (int j = 4; j >= 0; --j) { ps = fma(ps, z, kSin[j]); pc = fma(pc, z, kCos[j]); }
    const double sn = fma(al * z, ps, al);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int swap = (x[3] >> 29) & 1u;
    double s0 = swap ? cs : sn, c0 = swap ? sn : cs;
    s0 = hilo(hi_of(s0) ^ ((x[3] << 1) & 0x80000000u), lo_of(s0));
    c0 = hilo(hi_of(c0) ^ (x[3] & 0x80000000u), lo_of(c0));
    *z0 = g * s0;
    *z1 = g * c0;
}

/* z0, z1 (count) for the counters (tau[i], episode, env[i], comp[i]) under the 64-bit seed */
void philox_port_normals(uint64_t seed, uint32_t episode, const uint32_t* tau, const uint32_t* env, const uint32_t* comp,
                         int count, double* z0, double* z1) {
    for (int i = 0; i < count; ++i) {
        const uint32_t c[4] = {tau[i], episode, env[i], comp[i]};
        uint32_t x[4];
        philox_port_4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32), x);
        normal_pair(x, &z0[i], &z1[i]);
    }
}
