// Generic explicit Runge–Kutta engine of qrl_ode_step: one kernel template for every tableau besides the two
// specialised kernels (classic RK4 with staged row stores, DOPRI5 with dense output).
//
// Covers the method names of the reference's solver plugins that are explicit RK schemes
//   torchdiffeq (quantrl/solvers/torch.py:28):  'euler', 'midpoint', 'adaptive_heun', 'fehlberg2', 'bosh3', 'dopri8'
//   diffrax     (quantrl/solvers/jax.py:29):    'euler', 'heun', 'midpoint', 'bosh3', 'tsit5', 'dopri8'
//   SciPy       (quantrl/solvers/numpy.py:31):  'dop853' / 'DOP853', 'RK23' (= Bogacki–Shampine)
// with ONE step-size controller PER ENVIRONMENT, like dopri5_kernel.
//
// Form: S stages k_0..k_{S-1}, y_new = y + h sum_j B_j k_j, then k_S = f(t + h, y_new), which is both the last entry
// of the error estimate  err = h sum_{j<=S} E_j k_j  and k_0 of the next step (for FSAL pairs this IS the last stage,
// for the others it is the evaluation the next step needs anyway: S right-hand sides per step in every case).
// Grid output: a step never crosses a grid point (it is clipped to it, like the reference's SciPy path, which
// integrates to every T_step[i], solvers/numpy.py:154-161), so rows are exact step results — no interpolant needed
// and the result is independent of the method's dense-output formula.  A clipped step does not shrink the carried
// step size.  Controller constants as in dopri5_kernel / SciPy (safety 0.9, factor in [0.2, 10], no growth right after
// a rejection), RMS error norm with scale atol + rtol * max(|y|, |y_new|).
// CPU twin: oracle/lyap_oracle.py:erk_interval.
#pragma once
#include "dop853_tableau.cuh"

namespace qrl {
namespace erk {

// ---- tableaus ---------------------------------------------------------------------------------------------------
// Bogacki–Shampine 3(2) ('bosh3', SciPy 'RK23')
__device__ constexpr double BOSH3_A[3][3] = {{0, 0, 0}, {1.0 / 2, 0, 0}, {0, 3.0 / 4, 0}};
__device__ constexpr double BOSH3_B[3] = {2.0 / 9, 1.0 / 3, 4.0 / 9};
__device__ constexpr double BOSH3_C[3] = {0, 1.0 / 2, 3.0 / 4};
__device__ constexpr double BOSH3_E[4] = {2.0 / 9 - 7.0 / 24, 1.0 / 3 - 1.0 / 4, 4.0 / 9 - 1.0 / 3, -1.0 / 8};
// Tsitouras 5(4) ('tsit5'; Ch. Tsitouras, Comput. Math. Appl. 62 (2011) 770)
__device__ constexpr double TSIT5_A[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0}};
__device__ constexpr double TSIT5_B[6] = {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
                                          -3.290069515436081, 2.324710524099774};
__device__ constexpr double TSIT5_C[6] = {0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0};
__device__ constexpr double TSIT5_E[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                                          -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 1.0 / 66};
// Heun 2(1) ('adaptive_heun' / 'heun'): trapezoidal rule with the Euler step as the embedded solution
__device__ constexpr double HEUN_A[2][2] = {{0, 0}, {1.0, 0}};
__device__ constexpr double HEUN_B[2] = {0.5, 0.5};
__device__ constexpr double HEUN_C[2] = {0, 1.0};
__device__ constexpr double HEUN_E[3] = {-0.5, 0.5, 0};
// Fehlberg 2(1) ('fehlberg2')
__device__ constexpr double FEHL2_A[3][3] = {{0, 0, 0}, {1.0 / 2, 0, 0}, {1.0 / 256, 255.0 / 256, 0}};
__device__ constexpr double FEHL2_B[3] = {1.0 / 512, 255.0 / 256, 1.0 / 512};
__device__ constexpr double FEHL2_C[3] = {0, 1.0 / 2, 1.0};
__device__ constexpr double FEHL2_E[4] = {-1.0 / 512, 0, 1.0 / 512, 0};
// fixed step: explicit Euler and the explicit midpoint rule
__device__ constexpr double EULER_A[1][1] = {{0}};
__device__ constexpr double EULER_B[1] = {1.0};
__device__ constexpr double EULER_C[1] = {0};
__device__ constexpr double MIDP_A[2][2] = {{0, 0}, {0.5, 0}};
__device__ constexpr double MIDP_B[2] = {0, 1.0};
__device__ constexpr double MIDP_C[2] = {0, 0.5};

#define QRL_ERK_TABLEAU(NAME, PFX, STAGES, IS_ADAPTIVE, EXPO)                                              \
    struct NAME {                                                                                          \
        static constexpr int S = STAGES;                                                                   \
        static constexpr bool ADAPTIVE = IS_ADAPTIVE, DOP853 = false;                                      \
        static constexpr double ERR_EXP = EXPO;                                                            \
        __device__ static constexpr double A(int s, int j) { return PFX##_A[s][j]; }                       \
        __device__ static constexpr double B(int j) { return PFX##_B[j]; }                                 \
        __device__ static constexpr double C(int s) { return PFX##_C[s]; }                                 \
        __device__ static constexpr double E(int j) { return PFX##_E[j]; }                                 \
    };
#define EULER_E EULER_B   /* unused (fixed step) */
#define MIDP_E MIDP_B
QRL_ERK_TABLEAU(Bosh3, BOSH3, 3, true, -1.0 / 3)
QRL_ERK_TABLEAU(Tsit5, TSIT5, 6, true, -1.0 / 5)
QRL_ERK_TABLEAU(Heun21, HEUN, 2, true, -1.0 / 2)
QRL_ERK_TABLEAU(Fehlberg2, FEHL2, 3, true, -1.0 / 2)
QRL_ERK_TABLEAU(Euler, EULER, 1, false, 0.0)
QRL_ERK_TABLEAU(Midpoint, MIDP, 2, false, 0.0)
#undef EULER_E
#undef MIDP_E
struct Dop853 {   // two error estimators (orders 5 and 3), Hairer's combination
    static constexpr int S = 12;
    static constexpr bool ADAPTIVE = true, DOP853 = true;
    static constexpr double ERR_EXP = -1.0 / 8;
    __device__ static constexpr double A(int s, int j) { return DOP853_A[s][j]; }
    __device__ static constexpr double B(int j) { return DOP853_B[j]; }
    __device__ static constexpr double C(int s) { return DOP853_C[s]; }
    __device__ static constexpr double E(int j) { return DOP853_E5[j]; }
    __device__ static constexpr double E3(int j) { return DOP853_E3[j]; }
};

constexpr double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
constexpr int MAX_STEPS = 4000000;
}  // namespace erk

template <class Sys, class TB, int BLOCK>
__global__ void __launch_bounds__(BLOCK) erk_kernel(const __grid_constant__ qrl_ode_args a) {
    constexpr int D = Dim<Sys>::D, S = TB::S;
    extern __shared__ double ks[];  // [S + 1][D][BLOCK]: stage derivatives k_0..k_{S-1}, one column per thread; column S
                                    // stages the row being written (k_S = f(t + h, y_new) stays in registers)
    const int E = a.n_envs, K = a.K;
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    const bool live = e < E;
    const int el = live ? e : E - 1;
#define KS(s, d) ks[((s) * D + (d)) * BLOCK + threadIdx.x]

    Sys s;
    load_system(s, a, el);
    double y[D];
#pragma unroll
    for (int d = 0; d < D; ++d) y[d] = a.y0[(int64_t)el * D + d];

    double lo = INFINITY, hi = -INFINITY;
    // rows go out through the shared, non-inlined emitter (ode_impl.cuh:emit_row_thread) from a scratch column
    auto emit = [&](int row, const double* v) {
        if (!live) return;
#pragma unroll
        for (int d = 0; d < D; ++d) KS(S, d) = v[d];
        const double2 mm = emit_row_thread<Sys::M, Sys::Q>(a, row, e, &KS(S, 0), BLOCK, lo, hi);
        lo = mm.x;
        hi = mm.y;
    };
    emit(0, y);

    const double t_end = a.t0 + (double)K * a.t_ssz;
    const double h_fixed = a.t_ssz / (double)(a.substeps > 0 ? a.substeps : 1);
    double t = a.t0;
    double h = TB::ADAPTIVE ? ((a.h_state && a.h_state[el] > 0.0) ? a.h_state[el] : a.t_ssz) : h_fixed;
    int n_acc = 0, n_rej = 0, guard = 0;
    bool rejected = false, failed = false;
    {
        double k0[D];
        rhs(s, t, y, k0);
#pragma unroll
        for (int d = 0; d < D; ++d) KS(0, d) = k0[d];
    }
    for (int row = 1; row <= K && live; ++row) {
        const double t_row = (row == K) ? t_end : a.t0 + (double)row * a.t_ssz;
        int sub = 0;
        while (true) {
            if (guard++ >= erk::MAX_STEPS) { failed = true; break; }
            // ---- step size: never across the grid point ----
            const double remaining = t_row - t;
            bool last;
            double ht;
            if (TB::ADAPTIVE) {
                last = h >= remaining * (1.0 - 1e-12);
                ht = last ? remaining : h;
            } else {
                last = (++sub >= a.substeps);
                ht = last ? remaining : h_fixed;
            }
            double yt[D], kk[D];
            // ---- stages 1 .. S-1 ----
#pragma unroll
            for (int st = 1; st < S; ++st) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < st; ++j)
                        if (TB::A(st, j) != 0.0) acc = fma(TB::A(st, j), KS(j, d), acc);
                    yt[d] = fma(ht, acc, y[d]);
                }
                rhs(s, t + TB::C(st) * ht, yt, kk);
#pragma unroll
                for (int d = 0; d < D; ++d) KS(st, d) = kk[d];
            }
            // ---- new solution and its derivative ----
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < S; ++j)
                    if (TB::B(j) != 0.0) acc = fma(TB::B(j), KS(j, d), acc);
                yt[d] = fma(ht, acc, y[d]);
            }
            const double t_new = last ? t_row : t + ht;
            rhs(s, t_new, yt, kk);
            bool accept = true;
            double fac = 1.0;
            if (TB::ADAPTIVE) {
                double err;
                if constexpr (TB::DOP853) {
                    double s5 = 0.0, s3 = 0.0;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        double e5 = TB::E(S) * kk[d], e3 = TB::E3(S) * kk[d];
#pragma unroll
                        for (int j = 0; j < S; ++j) {
                            if (TB::E(j) != 0.0) e5 = fma(TB::E(j), KS(j, d), e5);
                            if (TB::E3(j) != 0.0) e3 = fma(TB::E3(j), KS(j, d), e3);
                        }
                        const double sc = fma(a.rtol, fmax(fabs(y[d]), fabs(yt[d])), a.atol);
                        e5 /= sc;
                        e3 /= sc;
                        s5 = fma(e5, e5, s5);
                        s3 = fma(e3, e3, s3);
                    }
                    const double den = s5 + 0.01 * s3;
                    err = (den > 0.0) ? fabs(ht) * s5 / sqrt(den * (double)D) : 0.0;
                } else {
                    double sq = 0.0;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        double ev = TB::E(S) * kk[d];
#pragma unroll
                        for (int j = 0; j < S; ++j)
                            if (TB::E(j) != 0.0) ev = fma(TB::E(j), KS(j, d), ev);
                        const double sc = fma(a.rtol, fmax(fabs(y[d]), fabs(yt[d])), a.atol);
                        const double r = ht * ev / sc;
                        sq = fma(r, r, sq);
                    }
                    err = sqrt(sq / (double)D);
                }
                accept = err <= 1.0;
                if (accept) {
                    fac = (err == 0.0) ? erk::MAX_FACTOR
                                       : fmin(erk::MAX_FACTOR, fmax(erk::MIN_FACTOR, erk::SAFETY * pow(err, TB::ERR_EXP)));
                    if (rejected) fac = fmin(1.0, fac);
                } else {
                    fac = fmax(erk::MIN_FACTOR, erk::SAFETY * pow(err, TB::ERR_EXP));
                }
                if (!(err == err)) { failed = true; break; }   // NaN: give up, the flag is raised below
            }
            if (accept) {
                rejected = false;
                ++n_acc;
                t = t_new;
#pragma unroll
                for (int d = 0; d < D; ++d) { y[d] = yt[d]; KS(0, d) = kk[d]; }
                // a step that was clipped to the grid point does not shrink the carried step size
                if (TB::ADAPTIVE) h = (last && fac >= 1.0) ? fmax(h, ht * fac) : ht * fac;
                if (last) break;
            } else {
                h = ht * fac;
                rejected = true;
                ++n_rej;
            }
        }
        if (failed) break;
        emit(row, y);
    }
#undef KS
    if (live) {
        if (a.y_last)
#pragma unroll
            for (int d = 0; d < D; ++d) a.y_last[(int64_t)e * D + d] = y[d];
        if (TB::ADAPTIVE && a.h_state) a.h_state[e] = h;
        if (a.n_steps) { a.n_steps[2 * e] += n_acc; a.n_steps[2 * e + 1] += n_rej; }
        bool bad = failed;
#pragma unroll
        for (int d = 0; d < D; ++d) bad |= !isfinite(y[d]);
        if (a.nan_flag && bad) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

// The same engine with q*q lanes per environment (q = 4, LaneEnv in ode_impl.cuh): stage derivatives in registers
// (S + 1 for the lane's covariance element, (S + 1) * 2M for the replicated modes), rows written by the env's own lanes,
// per-env controller uniform over a half-warp, error norm reduced across the env's lanes.  Same tableaus, controller
// and order of operations per element as erk_kernel.
template <class Sys, class TB, int BLOCK>
__global__ void __launch_bounds__(BLOCK) erk_lanes_kernel(const __grid_constant__ qrl_ode_args a) {
    using Env = LaneEnv<Sys, BLOCK / 16>;
    constexpr int M = Env::M, M2 = Env::M2, D = Env::D, S = TB::S;
    __shared__ typename Env::Smem smem;
    Env env(a, smem);
    const int K = a.K, el = env.el;
    double ym[M2];
#pragma unroll
    for (int c = 0; c < 2 * M; ++c) ym[c] = a.y0[(int64_t)el * D + c];
    double v = a.y0[env.col_v];
    if (env.live) {
        env.emit(0, v, env.pick_mode(ym));
        const double t_end = a.t0 + (double)K * a.t_ssz;
        const double h_fixed = a.t_ssz / (double)(a.substeps > 0 ? a.substeps : 1);
        double t = a.t0;
        double h = TB::ADAPTIVE ? ((a.h_state && a.h_state[el] > 0.0) ? a.h_state[el] : a.t_ssz) : h_fixed;
        int n_acc = 0, n_rej = 0, guard = 0;
        bool rejected = false, failed = false;
        double kv[S + 1], km[S + 1][M2];
        kv[0] = env.rhs(t, ym, v, km[0]);
        for (int row = 1; row <= K; ++row) {
            const double t_row = (row == K) ? t_end : a.t0 + (double)row * a.t_ssz;
            int sub = 0;
            while (true) {
                if (guard++ >= erk::MAX_STEPS) { failed = true; break; }
                // ---- step size: never across the grid point ----
                const double remaining = t_row - t;
                bool last;
                double ht;
                if (TB::ADAPTIVE) {
                    last = h >= remaining * (1.0 - 1e-12);
                    ht = last ? remaining : h;
                } else {
                    last = (++sub >= a.substeps);
                    ht = last ? remaining : h_fixed;
                }
                double ymt[M2], vt;
                // ---- stages 1 .. S-1 ----
#pragma unroll
                for (int st = 1; st < S; ++st) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < st; ++j)
                        if (TB::A(st, j) != 0.0) acc = fma(TB::A(st, j), kv[j], acc);
                    vt = fma(ht, acc, v);
#pragma unroll
                    for (int c = 0; c < 2 * M; ++c) {
                        double am = 0.0;
#pragma unroll
                        for (int j = 0; j < st; ++j)
                            if (TB::A(st, j) != 0.0) am = fma(TB::A(st, j), km[j][c], am);
                        ymt[c] = fma(ht, am, ym[c]);
                    }
                    kv[st] = env.rhs(t + TB::C(st) * ht, ymt, vt, km[st]);
                }
                // ---- new solution and its derivative ----
                {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < S; ++j)
                        if (TB::B(j) != 0.0) acc = fma(TB::B(j), kv[j], acc);
                    vt = fma(ht, acc, v);
#pragma unroll
                    for (int c = 0; c < 2 * M; ++c) {
                        double am = 0.0;
#pragma unroll
                        for (int j = 0; j < S; ++j)
                            if (TB::B(j) != 0.0) am = fma(TB::B(j), km[j][c], am);
                        ymt[c] = fma(ht, am, ym[c]);
                    }
                }
                const double t_new = last ? t_row : t + ht;
                kv[S] = env.rhs(t_new, ymt, vt, km[S]);
                bool accept = true;
                double fac = 1.0;
                if (TB::ADAPTIVE) {
                    // my covariance element, plus (mode lanes) my mode amplitude
                    double kme[S + 1];
#pragma unroll
                    for (int j = 0; j <= S; ++j) kme[j] = env.pick_mode(km[j]);
                    const double ymo = env.pick_mode(ym), ymn = env.pick_mode(ymt);
                    double err;
                    if constexpr (TB::DOP853) {
                        auto terms = [&](const double* k, double yo, double yn, double& t5, double& t3) {
                            double e5 = TB::E(S) * k[S], e3 = TB::E3(S) * k[S];
#pragma unroll
                            for (int j = 0; j < S; ++j) {
                                if (TB::E(j) != 0.0) e5 = fma(TB::E(j), k[j], e5);
                                if (TB::E3(j) != 0.0) e3 = fma(TB::E3(j), k[j], e3);
                            }
                            const double sc = fma(a.rtol, fmax(fabs(yo), fabs(yn)), a.atol);
                            e5 /= sc;
                            e3 /= sc;
                            t5 = e5 * e5;
                            t3 = e3 * e3;
                        };
                        double s5, s3;
                        terms(kv, v, vt, s5, s3);
                        if (env.mode_lane) {
                            double m5, m3;
                            terms(kme, ymo, ymn, m5, m3);
                            s5 += m5;
                            s3 += m3;
                        }
                        s5 = env.env_sum(s5);
                        s3 = env.env_sum(s3);
                        const double den = s5 + 0.01 * s3;
                        err = (den > 0.0) ? fabs(ht) * s5 / sqrt(den * (double)D) : 0.0;
                    } else {
                        auto term = [&](const double* k, double yo, double yn) {
                            double ev = TB::E(S) * k[S];
#pragma unroll
                            for (int j = 0; j < S; ++j)
                                if (TB::E(j) != 0.0) ev = fma(TB::E(j), k[j], ev);
                            const double sc = fma(a.rtol, fmax(fabs(yo), fabs(yn)), a.atol);
                            const double r = ht * ev / sc;
                            return r * r;
                        };
                        double sq = term(kv, v, vt);
                        if (env.mode_lane) sq += term(kme, ymo, ymn);
                        err = sqrt(env.env_sum(sq) / (double)D);
                    }
                    accept = err <= 1.0;
                    if (accept) {
                        fac = (err == 0.0) ? erk::MAX_FACTOR
                                           : fmin(erk::MAX_FACTOR, fmax(erk::MIN_FACTOR, erk::SAFETY * pow(err, TB::ERR_EXP)));
                        if (rejected) fac = fmin(1.0, fac);
                    } else {
                        fac = fmax(erk::MIN_FACTOR, erk::SAFETY * pow(err, TB::ERR_EXP));
                    }
                    if (!(err == err)) { failed = true; break; }   // NaN: give up, the flag is raised below
                }
                if (accept) {
                    rejected = false;
                    ++n_acc;
                    t = t_new;
                    v = vt;
                    kv[0] = kv[S];
#pragma unroll
                    for (int c = 0; c < 2 * M; ++c) { ym[c] = ymt[c]; km[0][c] = km[S][c]; }
                    // a step that was clipped to the grid point does not shrink the carried step size
                    if (TB::ADAPTIVE) h = (last && fac >= 1.0) ? fmax(h, ht * fac) : ht * fac;
                    if (last) break;
                } else {
                    h = ht * fac;
                    rejected = true;
                    ++n_rej;
                }
            }
            if (failed) break;
            env.emit(row, v, env.pick_mode(ym));
        }
        env.finish(v, ym, h, TB::ADAPTIVE, n_acc, n_rej, failed);
    }
    if (a.obs_minmax) block_minmax_commit(env.lo, env.hi, a.obs_minmax);
}

template <class Sys, class TB>
int launch_erk(const qrl_ode_args& a, cudaStream_t st, const char* name) {
    if constexpr (Sys::Q == 4) {
        // q = 4: 16 lanes per env for the same batch sizes as dopri5_lanes_kernel (ode_impl.cuh:launch_ode).  B200, E = 1024,
        // 100 grid rows, lanes vs thread per env: tsit5 0.21 vs 1.06 ms, bosh3 0.35 vs 4.7 ms, dop853 0.38 vs 4.5 ms,
        // adaptive_heun 6.3 vs 129 ms; tsit5 at E = 8192: 0.81 vs 1.16 ms, at 16384: 1.50 vs 1.24 ms
        const bool lanes = (a.n_envs <= 12288 || (a.flags & QRL_ODE_FORCE_LANES)) && !(a.flags & QRL_ODE_FORCE_THREAD);
        if (lanes) {
            constexpr int BL = 64;
            erk_lanes_kernel<Sys, TB, BL><<<(a.n_envs + BL / 16 - 1) / (BL / 16), BL, 0, st>>>(a);
            return after_launch(name);
        }
    }
    constexpr int D = Dim<Sys>::D, BLOCK = 32;
    constexpr size_t smem = (size_t)(TB::S + 1) * D * BLOCK * sizeof(double);
    if constexpr (smem > 220 * 1024) {
        return fail(QRL_E_UNSUPPORTED, "%s", "qrl_ode_step: this method's stage storage does not fit for this state size");
    } else {
        static OncePerDevice attr;
        if (smem > 48 * 1024 && attr.first())
            QRL_CUDA(cudaFuncSetAttribute(erk_kernel<Sys, TB, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        erk_kernel<Sys, TB, BLOCK><<<(a.n_envs + BLOCK - 1) / BLOCK, BLOCK, smem, st>>>(a);
        return after_launch(name);
    }
}

}  // namespace qrl
