/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the deterministic hot path for the optomechanical system, never
 * linked into the product.
 *
 * Restates  quantrl/envs/deterministic.py:653-746  (RHS: y = [Re a | Im a | vec(V)], dV/dt = A V + V A^T + D with the
 * hooks of oracle/systems.py: optomech_mode_rates / optomech_A / optomech_D, SURVEY.md §8c "canonical system") under
 * the fixed-step classic RK4 the north_star defines (oracle/lyap_oracle.py:rk4_interval, pinned by fixture G2: RK4 on
 * the reference's own env.func).  Second implementation of the oracle (tests/test_oracle_c_port.py) and the compiled,
 * threaded CPU comparator of BASELINE config 2 (tools/cpu_compiled_em.py --lyap).
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC oracle/lyap_port.c -o oracle/_build/liblyap_port.so -lm
 */
#include <stddef.h>

#define D_ 20

/* p = [kappa, gamma, omega_m, Delta0, g0, n_th, A_l];  y = [Re a, Re b, Im a, Im b, V(4x4 row-major)] */
static void rhs(const double* p, double u, const double* y, double* dy) {
    const double kappa = p[0], gamma = p[1], om = p[2], D0 = p[3], g0 = p[4], nth = p[5], Al = p[6];
    const double ar = y[0], br = y[1], ai = y[2], bi = y[3];
    /* da = -(kappa/2 - i D0) a + i g0 a (b + b*) + A_l (1 + u);  db = -(gamma/2 + i om) b + i g0 |a|^2 */
    const double s = 2.0 * br;
    dy[0] = -(kappa / 2) * ar - D0 * ai - g0 * ai * s + Al * (1.0 + u);
    dy[2] = -(kappa / 2) * ai + D0 * ar + g0 * ar * s;
    dy[1] = -(gamma / 2) * br + om * bi;
    dy[3] = -(gamma / 2) * bi - om * br + g0 * (ar * ar + ai * ai);
    const double Delta = D0 + 2 * g0 * br, GR = 2 * g0 * ar, GI = 2 * g0 * ai;
    const double A[4][4] = {{-kappa / 2, -Delta, -GI, 0.0}, {Delta, -kappa / 2, GR, 0.0},
                            {0.0, 0.0, -gamma / 2, om}, {GR, GI, -om, -gamma / 2}};
    const double Dd[4] = {kappa / 2, kappa / 2, gamma * (nth + 0.5), gamma * (nth + 0.5)};
    const double* V = y + 4;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double av = 0.0, va = 0.0;
            for (int k = 0; k < 4; ++k) {
                av += A[i][k] * V[k * 4 + j];
                va += V[i * 4 + k] * A[j][k];
            }
            dy[4 + i * 4 + j] = (av + va) + (i == j ? Dd[i] : 0.0);
        }
}

/* One action interval for E envs: params (E,7), actions (E,) held constant, y0 (E,20) -> Y (K+1,E,20), row 0 = y0;
 * K grid steps of size t_ssz, each cut into `substeps` RK4 steps. */
void lyap_port_rk4(int E, int K, int substeps, double t_ssz, const double* params, const double* actions,
                   const double* y0, double* Y) {
    const size_t row = (size_t)E * D_;
    const double h = t_ssz / substeps;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        double y[D_], k1[D_], k2[D_], k3[D_], k4[D_], w[D_];
        const double* p = params + (size_t)e * 7;
        const double u = actions[e];
        for (int c = 0; c < D_; ++c) Y[(size_t)e * D_ + c] = y[c] = y0[(size_t)e * D_ + c];
        for (int r = 1; r <= K; ++r) {
            for (int s = 0; s < substeps; ++s) {
                rhs(p, u, y, k1);
                for (int c = 0; c < D_; ++c) w[c] = y[c] + (0.5 * h) * k1[c];
                rhs(p, u, w, k2);
                for (int c = 0; c < D_; ++c) w[c] = y[c] + (0.5 * h) * k2[c];
                rhs(p, u, w, k3);
                for (int c = 0; c < D_; ++c) w[c] = y[c] + h * k3[c];
                rhs(p, u, w, k4);
                for (int c = 0; c < D_; ++c) y[c] = y[c] + (h / 6.0) * (k1[c] + 2.0 * k2[c] + 2.0 * k3[c] + k4[c]);
            }
            for (int c = 0; c < D_; ++c) Y[(size_t)r * row + (size_t)e * D_ + c] = y[c];
        }
    }
}
