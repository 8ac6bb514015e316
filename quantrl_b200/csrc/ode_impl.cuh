// Deterministic action interval (qrl_ode_step): classical modes + Lyapunov covariance, fused over all grid points —
// sm_100a.
//
// Replaces solver.step -> integrate -> K x (adaptive integrator + Python RHS callback) of the reference
// (quantrl/envs/deterministic.py:646-651, quantrl/solvers/base.py:168-204, quantrl/solvers/numpy.py:87-178) and the RHS
// `func` with its hooks (deterministic.py:653-746, 820-867):
//     y = [Re a | Im a | vec(V)],   d a/dt = mode_rates(t, a, u),   dV/dt = A V + V A^T + D
// One thread integrates one environment and keeps its state, the RK stage vectors and the drift matrix in registers;
// the system hooks are device functors (systems.cuh).
//   RK4   : fixed step h = t_ssz / substeps; every grid row is staged in shared memory and written by the whole block
//           as one contiguous, coalesced run of the (K+1, E, d) block (+ observation noise, + min/max).
//   DOPRI5: Dormand–Prince 5(4), one step-size controller PER ENVIRONMENT (the reference shares one controller across
//           the flattened batch), free-running steps with the 4th-order dense output evaluated on the
//           grid, the interval end is hit exactly (actions change there).  Stage vectors live in shared memory.
//           Controller = oracle/lyap_oracle.py:dopri5_dense_interval.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "systems.cuh"
#include "measures.cuh"

namespace qrl {

template <class Sys>
struct Dim {
    static constexpr int M = Sys::M, Q = Sys::Q, D = 2 * Sys::M + Sys::Q * Sys::Q;
};

template <class Sys>
__device__ __forceinline__ void rhs(const Sys& s, double t, const double* y, double* dy) {
    constexpr int M = Sys::M, Q = Sys::Q;
    double A[Q][Q];
    s.eval(t, y, dy, A);
    const double* V = y + 2 * M;
    double* dV = dy + 2 * M;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double av = 0.0, va = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                av = fma(A[i][k], V[k * Q + j], av);   // (A V)_ij
                va = fma(V[i * Q + k], A[j][k], va);   // (V A^T)_ij
            }
            dV[i * Q + j] = (av + va) + s.D(i, j);     // same association as deterministic.py:713-732
        }
    }
}

template <class Sys>
__device__ __forceinline__ void load_system(Sys& s, const qrl_ode_args& a, int e) {
    if constexpr (is_const_ad<Sys>::value) {
        s.load_AD(a.A + (int64_t)e * a.A_stride_e, a.D + (int64_t)e * a.D_stride_e);
    } else {
        s.load(a.params + (int64_t)e * a.params_stride_e, a.actions + (int64_t)e * a.n_actions);
    }
}

// measurement noise of one element (row tau, env, component d): fed tensor or the Philox stream
// counter = (tau, episode, env_global, 0x10000 + d/2), z[d & 1]  (include/quantrl_b200.h)
__device__ __forceinline__ double obs_noise_at(const qrl_ode_args& a, int row, int e, int d, int Dd) {
    if (a.obs_noise) return a.obs_noise[((int64_t)row * a.n_envs + e) * Dd + d];
    if (a.obs_std) {
        double z0, z1;
        philox_normal2(a.seed, (uint32_t)(a.t_idx + row), a.episode, (uint32_t)(a.env_offset + e),
                       0x10000u + (uint32_t)(d >> 1), z0, z1);
        return a.obs_std[(int64_t)e * a.obs_std_stride_e + d] * ((d & 1) ? z1 : z0);
    }
    return 0.0;
}

// fused reward of (row, env e) from the env's state vector y (covariance block at y + 2M): log-negativity of the States
// or of the Observations (= states + this row's measurement noise); updates reward / reward_last / rewards_cum
template <int M, int Q>
__device__ __forceinline__ void ode_reward_row(const qrl_ode_args& a, int row, int e, const double* y) {
    if constexpr (Q == 4) {
        if (a.reward_kind == QRL_REWARD_NONE) return;
        constexpr int D = 2 * M + Q * Q;
        // odd kinds are evaluated on the Observations (= states + this row's measurement noise), even ones on the States
        const bool on_obs = (a.reward_kind & 1) != 0;
        double V[16];
#pragma unroll
        for (int c = 0; c < 16; ++c)
            V[c] = y[2 * M + c] + (on_obs ? obs_noise_at(a, row, e, 2 * M + c, D) : 0.0);
        double r;
        if (a.reward_kind <= QRL_REWARD_ENTAN_LN_STATE) r = log_negativity_2mode(V);
        else if (a.reward_kind <= QRL_REWARD_SYNC_C_STATE) r = sync_complete_2mode(V);
        else {
            double md[4] = {1.0, 0.0, 1.0, 0.0};    // re_0, im_0, re_1, im_1
            if constexpr (M >= 2) {
                md[0] = y[0] + (on_obs ? obs_noise_at(a, row, e, 0, D) : 0.0);
                md[1] = y[M] + (on_obs ? obs_noise_at(a, row, e, M, D) : 0.0);
                md[2] = y[1] + (on_obs ? obs_noise_at(a, row, e, 1, D) : 0.0);
                md[3] = y[M + 1] + (on_obs ? obs_noise_at(a, row, e, M + 1, D) : 0.0);
            }
            r = sync_phase_2mode(md[0], md[1], md[2], md[3], V);
        }
        if (a.reward) a.reward[(int64_t)row * a.n_envs + e] = r;
        if (row == a.K) {
            if (a.reward_last) a.reward_last[e] = r;
            if (a.rewards_cum) a.rewards_cum[e] += r;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// RK4
// ------------------------------------------------------------------------------------------------------------------
template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) rk4_kernel(const qrl_ode_args a) {
    constexpr int D = Dim<Sys>::D;
    constexpr int DP = D | 1;  // odd row pitch: conflict-free staging
    extern __shared__ double stage[];  // [BLOCK][DP]
    const int E = a.n_envs, K = a.K;
    const int e0 = blockIdx.x * BLOCK;
    const int e = e0 + threadIdx.x;
    const bool live = e < E;
    const int n_live = min(BLOCK, E - e0);
    const int el = live ? e : E - 1;

    Sys s;
    load_system(s, a, el);

    // coalesced load of y0 through the staging tile
    for (int idx = threadIdx.x; idx < n_live * D; idx += BLOCK)
        stage[(idx / D) * DP + (idx % D)] = a.y0[(int64_t)e0 * D + idx];
    __syncthreads();
    double y[D];
#pragma unroll
    for (int d = 0; d < D; ++d) y[d] = live ? stage[threadIdx.x * DP + d] : 0.0;
    __syncthreads();

    double lo = INFINITY, hi = -INFINITY;
    const int nsub = a.substeps;
    const double h = a.t_ssz / (double)nsub;
    const bool want_rows = a.states || a.observations || a.obs_minmax || (a.reward_kind != QRL_REWARD_NONE && a.reward);

    for (int row = 0; row <= K; ++row) {
        if (row > 0) {
            const double t_row = a.t0 + (double)(row - 1) * a.t_ssz;
            for (int sub = 0; sub < nsub; ++sub) {
                const double t = t_row + (double)sub * h;
                double k[D], yt[D], acc[D];
                rhs(s, t, y, k);
#pragma unroll
                for (int d = 0; d < D; ++d) { acc[d] = k[d]; yt[d] = fma(0.5 * h, k[d], y[d]); }
                rhs(s, t + 0.5 * h, yt, k);
#pragma unroll
                for (int d = 0; d < D; ++d) { acc[d] = fma(2.0, k[d], acc[d]); yt[d] = fma(0.5 * h, k[d], y[d]); }
                rhs(s, t + 0.5 * h, yt, k);
#pragma unroll
                for (int d = 0; d < D; ++d) { acc[d] = fma(2.0, k[d], acc[d]); yt[d] = fma(h, k[d], y[d]); }
                rhs(s, t + h, yt, k);
#pragma unroll
                for (int d = 0; d < D; ++d) y[d] = fma(h / 6.0, acc[d] + k[d], y[d]);
            }
        }
        if (want_rows || row == K) {
            // stage the row, then the whole block writes n_live*D contiguous doubles of the (K+1,E,D) block
#pragma unroll
            for (int d = 0; d < D; ++d) stage[threadIdx.x * DP + d] = y[d];
            __syncthreads();
            const int64_t base = ((int64_t)row * E + e0) * D;
            for (int idx = threadIdx.x; idx < n_live * D; idx += BLOCK) {
                const int elx = idx / D, d = idx % D;
                const double v = stage[elx * DP + d];
                const double ob = v + obs_noise_at(a, row, e0 + elx, d, D);
                if (a.states) a.states[base + idx] = v;
                if (a.observations) a.observations[base + idx] = ob;
                if (row == K) {
                    if (a.y_last) a.y_last[(int64_t)e0 * D + idx] = v;
                    if (a.obs_last) a.obs_last[(int64_t)e0 * D + idx] = ob;
                    if (a.nan_flag && !isfinite(v)) *a.nan_flag = 1;
                }
                lo = fmin(lo, ob);
                hi = fmax(hi, ob);
            }
            if (live) ode_reward_row<Sys::M, Sys::Q>(a, row, e, y);
            __syncthreads();
        }
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

// ------------------------------------------------------------------------------------------------------------------
// DOPRI5 (per-env controller, dense output)
// ------------------------------------------------------------------------------------------------------------------
namespace dp {
__device__ constexpr double C2 = 1.0 / 5, C3 = 3.0 / 10, C4 = 4.0 / 5, C5 = 8.0 / 9;
__device__ constexpr double A21 = 1.0 / 5;
__device__ constexpr double A31 = 3.0 / 40, A32 = 9.0 / 40;
__device__ constexpr double A41 = 44.0 / 45, A42 = -56.0 / 15, A43 = 32.0 / 9;
__device__ constexpr double A51 = 19372.0 / 6561, A52 = -25360.0 / 2187, A53 = 64448.0 / 6561, A54 = -212.0 / 729;
__device__ constexpr double A61 = 9017.0 / 3168, A62 = -355.0 / 33, A63 = 46732.0 / 5247, A64 = 49.0 / 176,
                            A65 = -5103.0 / 18656;
__device__ constexpr double B1 = 35.0 / 384, B3 = 500.0 / 1113, B4 = 125.0 / 192, B5 = -2187.0 / 6784, B6 = 11.0 / 84;
__device__ constexpr double E1 = 71.0 / 57600, E3 = -71.0 / 16695, E4 = 71.0 / 1920, E5 = -17253.0 / 339200,
                            E6 = 22.0 / 525, E7 = -1.0 / 40;
// Shampine's 4th-order dense output (the interpolant of SciPy's RK45): y(t + x h) = y + h x sum_s k_s (P_s . [1,x,x^2,x^3])
__device__ constexpr double P[7][4] = {
    {1.0, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
    {0.0, 0.0, 0.0, 0.0},
    {0.0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
    {0.0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
    {0.0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
    {0.0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
    {0.0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};
constexpr double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
constexpr int MAX_STEPS = 1000000;
}  // namespace dp

template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) dopri5_kernel(const qrl_ode_args a) {
    constexpr int D = Dim<Sys>::D;
    extern __shared__ double ks[];  // [7][D][BLOCK]: stage derivatives, one column per thread (conflict-free)
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
    auto emit = [&](int row, const double* v) {
        if (!live) return;
        const int64_t base = ((int64_t)row * E + e) * D;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double ob = v[d] + obs_noise_at(a, row, e, d, D);
            if (a.states) a.states[base + d] = v[d];
            if (a.observations) a.observations[base + d] = ob;
            if (row == K && a.obs_last) a.obs_last[(int64_t)e * D + d] = ob;
            lo = fmin(lo, ob);
            hi = fmax(hi, ob);
        }
        ode_reward_row<Sys::M, Sys::Q>(a, row, e, v);
    };
    emit(0, y);

    const double t_end = a.t0 + (double)K * a.t_ssz;
    double t = a.t0;
    double h = (a.h_state && a.h_state[el] > 0.0) ? a.h_state[el] : a.t_ssz;
    int next_row = 1, n_acc = 0, n_rej = 0;
    bool rejected = false;
    {
        double k1[D];
        rhs(s, t, y, k1);
#pragma unroll
        for (int d = 0; d < D; ++d) KS(0, d) = k1[d];
    }
    int guard = 0;
    while (live && next_row <= K && guard++ < dp::MAX_STEPS) {
        const double remaining = t_end - t;
        const bool last = h >= remaining * (1.0 - 1e-12);
        const double ht = last ? remaining : h;
        double yt[D], kk[D];
        // stage 2
#pragma unroll
        for (int d = 0; d < D; ++d) yt[d] = fma(ht * dp::A21, KS(0, d), y[d]);
        rhs(s, t + dp::C2 * ht, yt, kk);
#pragma unroll
        for (int d = 0; d < D; ++d) KS(1, d) = kk[d];
        // stage 3
#pragma unroll
        for (int d = 0; d < D; ++d) yt[d] = fma(ht, fma(dp::A32, KS(1, d), dp::A31 * KS(0, d)), y[d]);
        rhs(s, t + dp::C3 * ht, yt, kk);
#pragma unroll
        for (int d = 0; d < D; ++d) KS(2, d) = kk[d];
        // stage 4
#pragma unroll
        for (int d = 0; d < D; ++d)
            yt[d] = fma(ht, fma(dp::A43, KS(2, d), fma(dp::A42, KS(1, d), dp::A41 * KS(0, d))), y[d]);
        rhs(s, t + dp::C4 * ht, yt, kk);
#pragma unroll
        for (int d = 0; d < D; ++d) KS(3, d) = kk[d];
        // stage 5
#pragma unroll
        for (int d = 0; d < D; ++d)
            yt[d] = fma(ht, fma(dp::A54, KS(3, d), fma(dp::A53, KS(2, d), fma(dp::A52, KS(1, d), dp::A51 * KS(0, d)))), y[d]);
        rhs(s, t + dp::C5 * ht, yt, kk);
#pragma unroll
        for (int d = 0; d < D; ++d) KS(4, d) = kk[d];
        // stage 6
#pragma unroll
        for (int d = 0; d < D; ++d)
            yt[d] = fma(ht, fma(dp::A65, KS(4, d), fma(dp::A64, KS(3, d), fma(dp::A63, KS(2, d),
                        fma(dp::A62, KS(1, d), dp::A61 * KS(0, d))))), y[d]);
        rhs(s, t + ht, yt, kk);
#pragma unroll
        for (int d = 0; d < D; ++d) KS(5, d) = kk[d];
        // 5th-order solution (= stage-7 argument, FSAL)
#pragma unroll
        for (int d = 0; d < D; ++d)
            yt[d] = fma(ht, fma(dp::B6, KS(5, d), fma(dp::B5, KS(4, d), fma(dp::B4, KS(3, d),
                        fma(dp::B3, KS(2, d), dp::B1 * KS(0, d))))), y[d]);
        rhs(s, t + ht, yt, kk);
        // error estimate, RMS norm with scale atol + rtol * max(|y|, |y_new|)
        double sq = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            KS(6, d) = kk[d];
            const double ev = ht * fma(dp::E7, kk[d], fma(dp::E6, KS(5, d), fma(dp::E5, KS(4, d),
                                   fma(dp::E4, KS(3, d), fma(dp::E3, KS(2, d), dp::E1 * KS(0, d))))));
            const double sc = fma(a.rtol, fmax(fabs(y[d]), fabs(yt[d])), a.atol);
            const double r = ev / sc;
            sq = fma(r, r, sq);
        }
        const double err = sqrt(sq / (double)D);
        if (err <= 1.0) {
            double fac = (err == 0.0) ? dp::MAX_FACTOR : fmin(dp::MAX_FACTOR, fmax(dp::MIN_FACTOR, dp::SAFETY * pow(err, -0.2)));
            if (rejected) fac = fmin(1.0, fac);
            rejected = false;
            ++n_acc;
            const double t_new = last ? t_end : t + ht;
            // dense output on every grid point inside (t, t_new]
            while (next_row <= K) {
                const double tg = (next_row == K) ? t_end : a.t0 + (double)next_row * a.t_ssz;
                if (tg > t_new) break;
                if (tg == t_new) {
                    emit(next_row, yt);
                } else {
                    const double x = (tg - t) / ht;
                    double yo[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        double q = 0.0;
#pragma unroll
                        for (int st = 0; st < 7; ++st) {
                            if (st == 1) continue;
                            const double p = fma(fma(fma(dp::P[st][3], x, dp::P[st][2]), x, dp::P[st][1]), x, dp::P[st][0]);
                            q = fma(KS(st, d), p, q);
                        }
                        yo[d] = fma(ht * x, q, y[d]);
                    }
                    emit(next_row, yo);
                }
                ++next_row;
            }
            t = t_new;
#pragma unroll
            for (int d = 0; d < D; ++d) { y[d] = yt[d]; KS(0, d) = kk[d]; }
            // a step that was clipped to the interval end does not shrink the carried step size
            h = (last && fac >= 1.0) ? fmax(h, ht * fac) : ht * fac;
        } else {
            h = ht * fmax(dp::MIN_FACTOR, dp::SAFETY * pow(err, -0.2));
            rejected = true;
            ++n_rej;
        }
    }
#undef KS
    if (live) {
        if (a.y_last)
#pragma unroll
            for (int d = 0; d < D; ++d) a.y_last[(int64_t)e * D + d] = y[d];
        if (a.h_state) a.h_state[e] = h;
        if (a.n_steps) { a.n_steps[2 * e] += n_acc; a.n_steps[2 * e + 1] += n_rej; }
        bool bad = next_row <= K;  // step budget exhausted
#pragma unroll
        for (int d = 0; d < D; ++d) bad |= !isfinite(y[d]);
        if (a.nan_flag && bad) *a.nan_flag = 1;
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}


// ------------------------------------------------------------------------------------------------------------------
// RK4, q*q lanes per environment (q = 4: half a warp) — for batches too small to fill the chip with one thread per
// env (BASELINE config 2: E = 1024 is 32 warps on 592 schedulers; measured 300 us per 100-sub-step interval with the
// thread-per-env kernel above).  Lane (i,j) owns V_ij; the classical mode amplitudes are replicated in every lane
// (a few dozen flops); per RK stage the group publishes V and the drift matrix A in shared memory and every lane
// accumulates its 2q products  (A V)_ij + (V A^T)_ij.  Rows of the (K+1,E,d) blocks are written directly: the lanes
// of one env cover its d contiguous doubles.
// ------------------------------------------------------------------------------------------------------------------
template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) rk4_lanes_kernel(const qrl_ode_args a) {
    constexpr int M = Sys::M, Q = Sys::Q, D = Dim<Sys>::D, L = Q * Q;   // L lanes per env
    static_assert(L == 16, "rk4_lanes_kernel is written for q = 4");
    constexpr int EPB = BLOCK / L;                                    // envs per block
    __shared__ double sA[EPB][Q][Q], sV[EPB][Q][Q];
    const int E = a.n_envs, K = a.K;
    const int g = threadIdx.x / L, l = threadIdx.x % L;
    const int i = l / Q, j = l % Q;
    const int e = blockIdx.x * EPB + g;
    const bool live = e < E;
    const int el = live ? e : E - 1;

    Sys s;
    load_system(s, a, el);
    double ym[2 * M > 0 ? 2 * M : 1];                                 // modes, replicated
#pragma unroll
    for (int c = 0; c < 2 * M; ++c) ym[c] = a.y0[(int64_t)el * D + c];
    double v = a.y0[(int64_t)el * D + 2 * M + l];                     // my covariance element
    const double dij = s.D(i, j);

    double lo = INFINITY, hi = -INFINITY;
    auto emit = [&](int row) {
        if (!live) return;
        const int64_t base = ((int64_t)row * E + e) * D;
        // lane l writes V element l; lanes 0 .. 2M-1 also write the modes: one env = D contiguous doubles
        auto put = [&](int d, double val) {
            const double ob = val + obs_noise_at(a, row, e, d, D);
            if (a.states) a.states[base + d] = val;
            if (a.observations) a.observations[base + d] = ob;
            if (row == K) {
                if (a.y_last) a.y_last[(int64_t)e * D + d] = val;
                if (a.obs_last) a.obs_last[(int64_t)e * D + d] = ob;
                if (a.nan_flag && !isfinite(val)) *a.nan_flag = 1;
            }
            lo = fmin(lo, ob);
            hi = fmax(hi, ob);
        };
        put(2 * M + l, v);
        if (l < 2 * M) {
            double mv = 0.0;
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) mv = (l == c) ? ym[c] : mv;
            put(l, mv);
        }
    };
    // fused reward: the group publishes its covariance elements, lane 0 evaluates the measure
    auto reward_row = [&](int row) {
        if (a.reward_kind == QRL_REWARD_NONE) return;
        __syncwarp();
        sV[g][i][j] = v;
        __syncwarp();
        if (live && l == 0) {
            double yfull[D];
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) yfull[c] = ym[c];
#pragma unroll
            for (int c = 0; c < L; ++c) yfull[2 * M + c] = (&sV[g][0][0])[c];
            ode_reward_row<M, Q>(a, row, e, yfull);
        }
        __syncwarp();
    };
    // one RHS evaluation at (modes ymt, my element vt): returns dV_ij, fills dm
    auto rhs_lane = [&](double t, const double* ymt, double vt, double* dm) -> double {
        double yfull[2 * M > 0 ? 2 * M : 1];
#pragma unroll
        for (int c = 0; c < 2 * M; ++c) yfull[c] = ymt[c];
        double A[Q][Q];
        s.eval(t, yfull, dm, A);
        __syncwarp();                     // previous stage's reads of sA / sV are done
        if (l == 0) {
#pragma unroll
            for (int r = 0; r < Q; ++r)
#pragma unroll
                for (int c = 0; c < Q; ++c) sA[g][r][c] = A[r][c];
        }
        sV[g][i][j] = vt;
        __syncwarp();
        double av = 0.0, va = 0.0;
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            av = fma(sA[g][i][k], sV[g][k][j], av);   // (A V)_ij
            va = fma(sV[g][i][k], sA[g][j][k], va);   // (V A^T)_ij
        }
        return (av + va) + dij;
    };

    emit(0);
    reward_row(0);
    const int nsub = a.substeps;
    const double h = a.t_ssz / (double)nsub;
    for (int row = 1; row <= K; ++row) {
        const double t_row = a.t0 + (double)(row - 1) * a.t_ssz;
        for (int sub = 0; sub < nsub; ++sub) {
            const double t = t_row + (double)sub * h;
            double dm[2 * M > 0 ? 2 * M : 1], ymt[2 * M > 0 ? 2 * M : 1], accm[2 * M > 0 ? 2 * M : 1];
            double k = rhs_lane(t, ym, v, dm);
            double accv = k, vt = fma(0.5 * h, k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = dm[c]; ymt[c] = fma(0.5 * h, dm[c], ym[c]); }
            k = rhs_lane(t + 0.5 * h, ymt, vt, dm);
            accv = fma(2.0, k, accv); vt = fma(0.5 * h, k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(0.5 * h, dm[c], ym[c]); }
            k = rhs_lane(t + 0.5 * h, ymt, vt, dm);
            accv = fma(2.0, k, accv); vt = fma(h, k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(h, dm[c], ym[c]); }
            k = rhs_lane(t + h, ymt, vt, dm);
            v = fma(h / 6.0, accv + k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) ym[c] = fma(h / 6.0, accm[c] + dm[c], ym[c]);
        }
        emit(row);
        reward_row(row);
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

}  // namespace qrl
#include "ode_erk.cuh"
namespace qrl {

template <class Sys>
int launch_ode(const qrl_ode_args& a, cudaStream_t st) {
    constexpr int D = Dim<Sys>::D;
    if (a.method == QRL_ODE_RK4) {
        if constexpr (Sys::Q == 4) {
            // small batches: 16 lanes per env (measured crossover with one thread per env at E ~ 8000 on B200:
            // 95 vs 300 us at E = 1024, 175 vs 301 us at 4096, 317 vs 301 us at 8192; tools/quick_ode_bench.py)
            if ((a.n_envs <= 6144 || (a.flags & QRL_ODE_FORCE_LANES)) && !(a.flags & QRL_ODE_FORCE_THREAD)) {
                constexpr int BL = 64;
                rk4_lanes_kernel<Sys, BL><<<(a.n_envs + BL / 16 - 1) / (BL / 16), BL, 0, st>>>(a);
                return after_launch("rk4_lanes_kernel");
            }
        }
        constexpr int BLOCK = 32;
        const size_t smem = (size_t)BLOCK * (D | 1) * sizeof(double);
        const int blocks = (a.n_envs + BLOCK - 1) / BLOCK;
        rk4_kernel<Sys, BLOCK><<<blocks, BLOCK, smem, st>>>(a);
        return after_launch("rk4_kernel");
    }
    if (a.method > QRL_ODE_DOPRI5) {
        // the generic explicit Runge–Kutta engine (ode_erk.cuh) is instantiated for num_quads <= 4 (the BASELINE configs);
        // larger systems have RK4 and DOPRI5 (seven more kernels per system at q = 8 triple the build time)
        if constexpr (Sys::Q > 4)
            return fail(QRL_E_UNSUPPORTED, "%s", "qrl_ode_step: this method is built for num_quads <= 4 (use RK4 or DOPRI5)");
    }
    if constexpr (Sys::Q <= 4) switch (a.method) {
        case QRL_ODE_BOSH3: return launch_erk<Sys, erk::Bosh3>(a, st, "erk_kernel<bosh3>");
        case QRL_ODE_TSIT5: return launch_erk<Sys, erk::Tsit5>(a, st, "erk_kernel<tsit5>");
        case QRL_ODE_HEUN21: return launch_erk<Sys, erk::Heun21>(a, st, "erk_kernel<heun21>");
        case QRL_ODE_FEHLBERG2: return launch_erk<Sys, erk::Fehlberg2>(a, st, "erk_kernel<fehlberg2>");
        case QRL_ODE_DOP853: return launch_erk<Sys, erk::Dop853>(a, st, "erk_kernel<dop853>");
        case QRL_ODE_EULER: return launch_erk<Sys, erk::Euler>(a, st, "erk_kernel<euler>");
        case QRL_ODE_MIDPOINT: return launch_erk<Sys, erk::Midpoint>(a, st, "erk_kernel<midpoint>");
        default: break;
    }
    constexpr int BLOCK = 32;
    const size_t smem = (size_t)7 * D * BLOCK * sizeof(double);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        QRL_CUDA(cudaFuncSetAttribute(dopri5_kernel<Sys, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int blocks = (a.n_envs + BLOCK - 1) / BLOCK;
    dopri5_kernel<Sys, BLOCK><<<blocks, BLOCK, smem, st>>>(a);
    return after_launch("dopri5_kernel");
}


// one translation unit per system family keeps the build parallel (the q=8 kernels dominate compile time)
#define QRL_ODE_INSTANTIATE(NAME, ...) \
    int NAME(const qrl_ode_args& a, cudaStream_t st) { return launch_ode<__VA_ARGS__>(a, st); }

}  // namespace qrl
