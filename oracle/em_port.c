/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the stochastic hot path, never linked into the product.
 *
 * Restates, for E independent environments, the Euler–Maruyama interval of the reference
 *   quantrl/envs/stochastic.py:178-242   Y[i+1] = (I + A dt) Y[i] + n * Ws[t_idx + i]
 *   quantrl/envs/base.py:462             Observations = States + Observation_noises
 * (the same recurrence as oracle/em_oracle.py:em_interval_vec, which is pinned against the reference's fixtures).
 * Two entry points:
 *   em_port_fed   noise supplied by the caller — second implementation of the oracle, compared with the NumPy one and
 *                 with fixture G3 in tests/test_oracle_c_port.py;
 *   em_port_rng   noise drawn in place (xoshiro256++ per env + libm Box–Muller), OpenMP over environments — the
 *                 "what a compiled, threaded CPU implementation achieves" comparator of tools/cpu_compiled_em.py.
 *                 Its random stream is its own (neither NumPy's PCG64 nor the kernel's Philox): throughput only.
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC oracle/em_port.c -o oracle/_build/libem_port.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define MAXN 16

/* x0 (E,n); A (E,n,n); nz (E,n); W (K,E,n) already scaled by sqrt(dt); ON (K+1,E,n) or NULL; w (n) reward weights.
 * Outputs: States, Obs (K+1,E,n); Reward (K+1,E) = -sum_c w_c Obs_c^2 (row 0 included, as the fused functor does). */
void em_port_fed(int E, int n, int K, double dt, const double* x0, const double* A, const double* nz, const double* W,
                 const double* ON, const double* w, double* States, double* Obs, double* Reward) {
    const size_t row = (size_t)E * n;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        double M[MAXN][MAXN], x[MAXN], y[MAXN];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) M[i][j] = (i == j ? 1.0 : 0.0) + A[((size_t)e * n + i) * n + j] * dt;
        for (int c = 0; c < n; ++c) x[c] = x0[(size_t)e * n + c];
        for (int r = 0; r <= K; ++r) {
            if (r > 0) {
                for (int i = 0; i < n; ++i) {
                    double acc = M[i][0] * x[0];
                    for (int j = 1; j < n; ++j) acc = fma(M[i][j], x[j], acc);      /* the kernels' FMA chain over columns */
                    y[i] = fma(nz[(size_t)e * n + i], W[(size_t)(r - 1) * row + (size_t)e * n + i], acc);
                }
                for (int c = 0; c < n; ++c) x[c] = y[c];
            }
            double rew = 0.0;
            for (int c = 0; c < n; ++c) {
                const size_t k = (size_t)r * row + (size_t)e * n + c;
                const double o = x[c] + (ON ? ON[k] : 0.0);
                States[k] = x[c];
                Obs[k] = o;
                rew = fma(w ? w[c] : 1.0, o * o, rew);
            }
            Reward[(size_t)r * E + e] = -rew;
        }
    }
}

static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xoshiro(uint64_t s[4]) {
    const uint64_t r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
static inline uint64_t splitmix(uint64_t* z) {
    uint64_t x = (*z += 0x9E3779B97F4A7C15ull);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline void normal2(uint64_t s[4], double* z0, double* z1) {
    const double u1 = ((xoshiro(s) >> 11) + 1) * (1.0 / 9007199254740992.0);   /* (0, 1] */
    const double u2 = (xoshiro(s) >> 11) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586 * u2;
    *z0 = r * cos(a);
    *z1 = r * sin(a);
}

/* One action interval with in-place noise: same outputs as em_port_fed; x_last (E,n) receives the last States row. */
void em_port_rng(int E, int n, int K, double dt, uint64_t seed, uint64_t interval, const double* x0, const double* A,
                 const double* nz, const double* stds, const double* w, double* States, double* Obs, double* Reward,
                 double* x_last) {
    const size_t row = (size_t)E * n;
    const double sq = sqrt(dt);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        uint64_t z = seed ^ (0xD1342543DE82EF95ull * (uint64_t)(e + 1)) ^ (interval << 32), s[4];
        for (int i = 0; i < 4; ++i) s[i] = splitmix(&z);
        double M[MAXN][MAXN], x[MAXN], y[MAXN], g[2 * MAXN + 2];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) M[i][j] = (i == j ? 1.0 : 0.0) + A[((size_t)e * n + i) * n + j] * dt;
        for (int c = 0; c < n; ++c) x[c] = x0[(size_t)e * n + c];
        for (int r = 0; r <= K; ++r) {
            for (int c = 0; c < 2 * n; c += 2) normal2(s, &g[c], &g[c + 1]);     /* n increments + n measurement noises */
            if (r > 0) {
                for (int i = 0; i < n; ++i) {
                    double acc = M[i][0] * x[0];
                    for (int j = 1; j < n; ++j) acc = fma(M[i][j], x[j], acc);
                    y[i] = fma(nz[(size_t)e * n + i], sq * g[i], acc);
                }
                for (int c = 0; c < n; ++c) x[c] = y[c];
            }
            double rew = 0.0;
            for (int c = 0; c < n; ++c) {
                const size_t k = (size_t)r * row + (size_t)e * n + c;
                const double o = fma(stds[c], g[n + c], x[c]);
                States[k] = x[c];
                Obs[k] = o;
                rew = fma(w ? w[c] : 1.0, o * o, rew);
            }
            Reward[(size_t)r * E + e] = -rew;
        }
        for (int c = 0; c < n; ++c) x_last[(size_t)e * n + c] = x[c];
    }
}
