// Deterministic action interval (qrl_ode_step): classical modes + Lyapunov covariance, fused over all grid points —
// sm_100a.
//
// Replaces solver.step -> integrate -> K x (adaptive integrator + Python RHS callback) of the reference
// (quantrl/envs/deterministic.py:646-651, quantrl/solvers/base.py:168-204, quantrl/solvers/numpy.py:87-178) and the RHS
// `func` with its hooks (deterministic.py:653-746, 820-867):
//     y = [Re a | Im a | vec(V)],   d a/dt = mode_rates(t, a, u),   dV/dt = A V + V A^T + D
// The system hooks are device functors (systems.cuh).  Two layouts, chosen by batch size (launch_ode at the end of this
// file; measured crossovers in DESIGN.md 4.2), all performing the same operations in the same order:
//   * one THREAD per environment (large batches; any num_quads): state, RK stage vectors and drift matrix in registers
//       rk4_kernel     fixed step h = t_ssz / substeps; every grid row leaves through a shared-memory tile of 16-byte
//                      pieces as one contiguous run of the (K+1, E, d) block (+ observation noise, + min/max)
//       dopri5_kernel  Dormand–Prince 5(4), one step-size controller PER ENVIRONMENT (the reference shares one controller
//                      across the flattened batch), free-running steps with the 4th-order dense output
//                      evaluated on the grid, the interval end hit exactly (actions change there); stage vectors in
//                      shared-memory columns.  Controller = oracle/lyap_oracle.py:dopri5_dense_interval.
//   * q*q = 16 LANES per environment (small batches, num_quads = 4): lane (i,j) owns V_ij
//       rk4_split_kernel      warp-specialised: mode producer / covariance lanes / row writers over shared-memory rings
//       rk4_lanes_kernel      single role, functor replicated in the lanes
//       dopri5_lanes_kernel   stage derivatives in registers, per-env controller uniform over a half-warp (LaneEnv)
//   The other explicit Runge–Kutta methods (ode_erk.cuh, included below) exist in both layouts as well.
#pragma once
#include <stdlib.h>
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
                // structural zeros of the functor's drift matrix are skipped: x + 0*v == x for finite v, so the result
                // is the one of the dense product (an Inf/NaN in V still reaches the entries it is multiplied into)
                if (Sys::nz(i, k)) av = fma(A[i][k], V[k * Q + j], av);   // (A V)_ij
                if (Sys::nz(j, k)) va = fma(V[i * Q + k], A[j][k], va);   // (V A^T)_ij
            }
            // same association as deterministic.py:713-732
            dV[i * Q + j] = Sys::d_nz(i, j) ? (av + va) + s.D(i, j) : (av + va);
        }
    }
}

template <class Sys>
__device__ __forceinline__ void load_system(Sys& s, const qrl_ode_args& a, int e) {
    if constexpr (is_const_ad<Sys>::value) {
        s.load_AD(a.A + (int64_t)e * a.A_stride_e, a.D + (int64_t)e * a.D_stride_e);
    } else {
        const double* u = a.actions + (int64_t)e * a.n_actions;
        if (a.action_scale) {          // raw actions (possibly in pinned host memory): scale a private copy
            double us[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) us[k] = k < a.n_actions ? u[k] * a.action_scale[k] : 0.0;
            s.load(a.params + (int64_t)e * a.params_stride_e, us);
        } else {
            s.load(a.params + (int64_t)e * a.params_stride_e, u);
        }
    }
}

// The measurement noise of the deterministic path is a cold path (observation_stds is None by default in the reference,
// envs/base.py:417-424) that costs ~150 instructions per pair of normals: one out-of-line copy per kernel instead of one
// inlined copy per call site (the adaptive kernels were 18k-28k instructions and stalled on instruction fetch, ncu r1m).
// Arguments by value: a by-reference qrl_ode_args would turn every field access into a generic load.
static __device__ __noinline__ double2 philox_normal2_call(uint64_t seed, uint32_t tau, uint32_t episode, uint32_t env, uint32_t comp) {
    double z0, z1;
    philox_normal2(seed, tau, episode, env, comp, z0, z1);
    return make_double2(z0, z1);
}

// measurement noise of one element (row tau, env, component d): fed tensor or the Philox stream
// counter = (tau, episode, env_global, 0x10000 + d/2), z[d & 1]  (include/quantrl_b200.h)
__device__ __forceinline__ double obs_noise_at(const qrl_ode_args& a, int row, int e, int d, int Dd) {
    if (a.obs_noise) return a.obs_noise[((int64_t)row * a.n_envs + e) * Dd + d];
    if (a.obs_std) {
        const double2 z = philox_normal2_call(a.seed, (uint32_t)(a.t_idx + row), a.episode, (uint32_t)(a.env_offset + e),
                                              0x10000u + (uint32_t)(d >> 1));
        // explicitly rounded product: `state + noise` must not become a fused multiply-add in one kernel and a separate
        // multiply and add in another (the kernel variants are held to each other bit for bit)
        return __dmul_rn(a.obs_std[(int64_t)e * a.obs_std_stride_e + d], (d & 1) ? z.y : z.x);
    }
    return 0.0;
}

// fused reward of one row of one env from the vector yv the measure is defined on (the States row, or the Observations
// row = states + measurement noise, for the odd reward kinds); covariance block at yv + 2M
template <int M, int Q>
__device__ __forceinline__ double ode_reward_value(int reward_kind, const double* yv) {
    if constexpr (Q == 4) {
        const double* V = yv + 2 * M;
        if (reward_kind <= QRL_REWARD_ENTAN_LN_STATE) return log_negativity_2mode(V);
        if (reward_kind <= QRL_REWARD_SYNC_C_STATE) return sync_complete_2mode(V);
        if constexpr (M >= 2) return sync_phase_2mode(yv[0], yv[M], yv[1], yv[M + 1], V);   // re_0, im_0, re_1, im_1
        else return sync_phase_2mode(1.0, 0.0, 1.0, 0.0, V);
    } else {
        return 0.0;
    }
}
__device__ __forceinline__ void ode_reward_store(const qrl_ode_args& a, int row, int e, double r) {
    if (a.reward) a.reward[(int64_t)row * a.n_envs + e] = r;
    if (row == a.K) {
        if (a.reward_last) a.reward_last[e] = r;
        if (a.rewards_cum) a.rewards_cum[e] += r;
    }
}
// from the env's state vector y in registers (RK4 kernels)
template <int M, int Q>
__device__ __forceinline__ void ode_reward_row(const qrl_ode_args& a, int row, int e, const double* y) {
    if constexpr (Q == 4) {
        if (a.reward_kind == QRL_REWARD_NONE) return;
        constexpr int D = 2 * M + Q * Q;
        // odd kinds are evaluated on the Observations (= states + this row's measurement noise), even ones on the States
        const bool on_obs = (a.reward_kind & 1) != 0;
        double yv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) yv[c] = y[c];
        if (on_obs && (a.obs_noise || a.obs_std)) {
            // the measures read the covariance block and (phase synchronization) the amplitudes of modes 0 and 1
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const bool is_amp = M >= 2 && (c == 0 || c == 1 || c == M || c == M + 1);
                if (c >= 2 * M || (is_amp && a.reward_kind > QRL_REWARD_SYNC_C_STATE))
                    yv[c] = y[c] + obs_noise_at(a, row, e, c, D);
            }
        }
        ode_reward_store(a, row, e, ode_reward_value<M, Q>(a.reward_kind, yv));
    }
}

// the same from a row in memory, OUT OF LINE (a cold path: the hot loops of the callers stay compact; the by-reference
// argument block costs generic loads here, which is why the hot row writers do not go through such a call)
template <int M, int Q>
__device__ __noinline__ void ode_reward_row_mem(const qrl_ode_args& a, int row, int e, const double* yrow) {
    constexpr int D = 2 * M + Q * Q;
    double y[D];
#pragma unroll
    for (int c = 0; c < D; ++c) y[c] = yrow[c];
    ode_reward_row<M, Q>(a, row, e, y);
}

// One grid row of ONE environment, written by its own thread — the adaptive kernels, where the lanes of a warp reach
// a row at different times.  The row sits in a shared-memory column (element d at col[d * stride]).  A compact loop
// (not unrolled) over PAIRS of components: one Philox call (out of line), one 16-byte store per tensor.  Unrolled over
// the D components with the Philox code inlined at every call site, the adaptive kernels were 18k-28k instructions each
// and stalled on instruction fetch (ncu r1m, dopri5: 6.9 no-instruction stall cycles per issued instruction).
// Returns the updated (min, max) of the observations.
template <int M, int Q>
__device__ __forceinline__ double2 emit_row_thread(const qrl_ode_args& a, int row, int e, const double* col, int stride,
                                                   double lo, double hi) {
    constexpr int D = 2 * M + Q * Q, D2 = D / 2;
    static_assert(D % 2 == 0, "even state size");
    const int64_t base2 = ((int64_t)row * a.n_envs + e) * D2;
    double2* st = a.states ? reinterpret_cast<double2*>(a.states) + base2 : nullptr;
    double2* ob = a.observations ? reinterpret_cast<double2*>(a.observations) + base2 : nullptr;
    double2* ol = (row == a.K && a.obs_last) ? reinterpret_cast<double2*>(a.obs_last) + (int64_t)e * D2 : nullptr;
    const double2* fed = a.obs_noise ? reinterpret_cast<const double2*>(a.obs_noise) + base2 : nullptr;
    const double* sd = (!fed && a.obs_std) ? a.obs_std + (int64_t)e * a.obs_std_stride_e : nullptr;
    const bool want_reward = Q == 4 && a.reward_kind != QRL_REWARD_NONE;
    double yv[D];   // the States row for the fused reward (local memory: indexed by the loop counter)
#pragma unroll 1
    for (int c = 0; c < D2; ++c) {
        const double2 v = make_double2(col[(2 * c) * stride], col[(2 * c + 1) * stride]);
        double2 o = v;
        if (fed) {
            const double2 z = fed[c];
            o.x = v.x + z.x;
            o.y = v.y + z.y;
        } else if (sd) {
            const double2 z = philox_normal2_call(a.seed, (uint32_t)(a.t_idx + row), a.episode, (uint32_t)(a.env_offset + e),
                                                  0x10000u + (uint32_t)c);
            o.x = __dadd_rn(v.x, __dmul_rn(sd[2 * c], z.x));      // (explicit roundings: see obs_noise_at)
            o.y = __dadd_rn(v.y, __dmul_rn(sd[2 * c + 1], z.y));
        }
        if (st) st[c] = v;
        if (ob) ob[c] = o;
        if (ol) ol[c] = o;
        lo = fmin(lo, fmin(o.x, o.y));
        hi = fmax(hi, fmax(o.x, o.y));
        if (want_reward) {
            yv[2 * c] = v.x;
            yv[2 * c + 1] = v.y;
        }
    }
    if (want_reward) ode_reward_row_mem<M, Q>(a, row, e, yv);
    return make_double2(lo, hi);
}

// ------------------------------------------------------------------------------------------------------------------
// RK4
// ------------------------------------------------------------------------------------------------------------------
// One grid row of a block: the threads' state vectors are transposed through a shared-memory tile of 16-byte pieces
// (pitch P2 odd: the STS.128 of 8 consecutive threads fall into 8 different bank groups) and leave as ONE contiguous run
// of n_live*D doubles of the (K+1, E, D) block, two doubles (one STG.128) per item.  NOISE: 0 none, 1 fed tensor,
// 2 Philox stream (one Philox call per 16-byte piece: counter word 0x10000 + d/2 holds the normals of d and d+1).
template <int D, int BLOCK, int NOISE>
__device__ __forceinline__ void rk4_row_out(const qrl_ode_args& a, const double2* stage2, int row, int e0, int n_live,
                                            double& lo, double& hi) {
    constexpr int D2 = D / 2, P2 = D2 | 1;
    const int n2 = n_live * D2;
    const int64_t base2 = ((int64_t)row * a.n_envs + e0) * D2;
    double2* st = a.states ? reinterpret_cast<double2*>(a.states) + base2 : nullptr;
    double2* ob = a.observations ? reinterpret_cast<double2*>(a.observations) + base2 : nullptr;
    const double2* fed = NOISE == 1 ? reinterpret_cast<const double2*>(a.obs_noise) + base2 : nullptr;
    const bool last = row == a.K;
    double2* yl = (last && a.y_last) ? reinterpret_cast<double2*>(a.y_last) + (int64_t)e0 * D2 : nullptr;
    double2* ol = (last && a.obs_last) ? reinterpret_cast<double2*>(a.obs_last) + (int64_t)e0 * D2 : nullptr;
#pragma unroll 2
    for (int idx2 = threadIdx.x; idx2 < n2; idx2 += BLOCK) {
        const int elx = idx2 / D2, c = idx2 - elx * D2;
        const double2 v = stage2[elx * P2 + c];
        double2 o = v;
        if constexpr (NOISE == 1) {
            const double2 z = fed[idx2];
            o.x = v.x + z.x;
            o.y = v.y + z.y;
        } else if constexpr (NOISE == 2) {
            const double2 z = philox_normal2_call(a.seed, (uint32_t)(a.t_idx + row), a.episode,
                                                  (uint32_t)(a.env_offset + e0 + elx), 0x10000u + (uint32_t)c);
            const double* sd = a.obs_std + (int64_t)(e0 + elx) * a.obs_std_stride_e + 2 * c;
            o.x = __dadd_rn(v.x, __dmul_rn(sd[0], z.x));
            o.y = __dadd_rn(v.y, __dmul_rn(sd[1], z.y));
        }
        if (st) st[idx2] = v;
        if (ob) ob[idx2] = o;
        if (last) {
            if (yl) yl[idx2] = v;
            if (ol) ol[idx2] = o;
            if (a.nan_flag && !(isfinite(v.x) && isfinite(v.y))) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
        }
        lo = fmin(lo, fmin(o.x, o.y));
        hi = fmax(hi, fmax(o.x, o.y));
    }
}

template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) rk4_kernel(const __grid_constant__ qrl_ode_args a) {
    constexpr int D = Dim<Sys>::D;
    static_assert(D % 2 == 0, "the state vector [Re a | Im a | vec(V)] has an even number of components");
    constexpr int D2 = D / 2, P2 = D2 | 1;
    extern __shared__ double2 stage2[];  // [BLOCK][P2] 16-byte pieces
    const int E = a.n_envs, K = a.K;
    const int e0 = blockIdx.x * BLOCK;
    const int e = e0 + threadIdx.x;
    const bool live = e < E;
    const int n_live = min(BLOCK, E - e0);
    const int el = live ? e : E - 1;

    Sys s;
    load_system(s, a, el);

    // coalesced load of y0 through the staging tile
    {
        const double2* y0v = reinterpret_cast<const double2*>(a.y0) + (int64_t)e0 * D2;
        for (int idx2 = threadIdx.x; idx2 < n_live * D2; idx2 += BLOCK) {
            const int elx = idx2 / D2;
            stage2[elx * P2 + (idx2 - elx * D2)] = y0v[idx2];
        }
    }
    __syncthreads();
    double y[D];
#pragma unroll
    for (int c = 0; c < D2; ++c) {
        const double2 v = live ? stage2[threadIdx.x * P2 + c] : make_double2(0.0, 0.0);
        y[2 * c] = v.x;
        y[2 * c + 1] = v.y;
    }
    __syncthreads();

    double lo = INFINITY, hi = -INFINITY;
    const int nsub = a.substeps;
    const double h = a.t_ssz / (double)nsub;
    const bool want_rows = a.states || a.observations || a.obs_minmax || (a.reward_kind != QRL_REWARD_NONE && a.reward);
    const int noise = a.obs_noise ? 1 : (a.obs_std ? 2 : 0);

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
#pragma unroll
            for (int c = 0; c < D2; ++c) stage2[threadIdx.x * P2 + c] = make_double2(y[2 * c], y[2 * c + 1]);
            __syncthreads();
            if (noise == 0) rk4_row_out<D, BLOCK, 0>(a, stage2, row, e0, n_live, lo, hi);
            else if (noise == 1) rk4_row_out<D, BLOCK, 1>(a, stage2, row, e0, n_live, lo, hi);
            else rk4_row_out<D, BLOCK, 2>(a, stage2, row, e0, n_live, lo, hi);
            if (Sys::Q == 4 && a.reward_kind != QRL_REWARD_NONE && live)
                ode_reward_row_mem<Sys::M, Sys::Q>(a, row, e, reinterpret_cast<const double*>(stage2 + threadIdx.x * P2));
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
// the same coefficients as rows (stage arguments 2..6, then the 5th-order solution) for kernels that LOOP over the stages
// (dopri5_rows_kernel: the unrolled step did not fit the instruction cache)
static __constant__ double AS[6][6] = {{A21, 0, 0, 0, 0, 0}, {A31, A32, 0, 0, 0, 0}, {A41, A42, A43, 0, 0, 0},
                                       {A51, A52, A53, A54, 0, 0}, {A61, A62, A63, A64, A65, 0}, {B1, 0, B3, B4, B5, B6}};
static __constant__ double CS[6] = {C2, C3, C4, C5, 1.0, 1.0};
constexpr double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
constexpr int MAX_STEPS = 1000000;
}  // namespace dp

template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) dopri5_kernel(const __grid_constant__ qrl_ode_args a) {
    constexpr int D = Dim<Sys>::D;
    extern __shared__ double ks[];  // [7][D][BLOCK]: stage derivatives, one column per thread.  Column 1 (k2: zero weight in
                                    // the solution, the error estimate and the interpolant) doubles as the staging column of
                                    // the row being written: k2 is dead once stage 6 is done
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
    // the row staged in this thread's scratch column goes out through the shared (non-inlined) emitter
    auto emit_staged = [&](int row) {
        const double2 mm = emit_row_thread<Sys::M, Sys::Q>(a, row, e, &KS(1, 0), BLOCK, lo, hi);
        lo = mm.x;
        hi = mm.y;
    };
    auto emit = [&](int row, const double* v) {
        if (!live) return;
#pragma unroll
        for (int d = 0; d < D; ++d) KS(1, d) = v[d];
        emit_staged(row);
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
                    double pw[7];
#pragma unroll
                    for (int st = 0; st < 7; ++st)
                        pw[st] = fma(fma(fma(dp::P[st][3], x, dp::P[st][2]), x, dp::P[st][1]), x, dp::P[st][0]);
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        double q = 0.0;
#pragma unroll
                        for (int st = 0; st < 7; ++st) {
                            if (st == 1) continue;
                            q = fma(KS(st, d), pw[st], q);
                        }
                        KS(1, d) = fma(ht * x, q, y[d]);
                    }
                    emit_staged(next_row);
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
        if (a.nan_flag && bad) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
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
__global__ void __launch_bounds__(BLOCK) rk4_lanes_kernel(const __grid_constant__ qrl_ode_args a) {
    constexpr int M = Sys::M, Q = Sys::Q, D = Dim<Sys>::D, L = Q * Q;   // L lanes per env
    constexpr int M2 = 2 * M > 0 ? 2 * M : 1;
    static_assert(L == 16, "rk4_lanes_kernel is written for q = 4");
    constexpr int EPB = BLOCK / L;                                    // envs per block
    // exchange tiles, double buffered over the RK stages: stage s writes buffer s & 1, and the __syncwarp of stage s+1
    // orders every lane's reads of stage s before the writes of stage s+2 -> one warp barrier per stage
    __shared__ __align__(16) double sA[2][EPB][L], sV[2][EPB][L];
    const int E = a.n_envs, K = a.K;
    const int g = threadIdx.x / L, l = threadIdx.x % L;
    const int i = l / Q, j = l % Q;
    const int e = blockIdx.x * EPB + g;
    const bool live = e < E;
    const int el = live ? e : E - 1;

    Sys s;
    load_system(s, a, el);
    double ym[M2];                                                    // modes, replicated in the env's lanes
#pragma unroll
    for (int c = 0; c < 2 * M; ++c) ym[c] = a.y0[(int64_t)el * D + c];
    double v = a.y0[(int64_t)el * D + 2 * M + l];                     // my covariance element
    const double dij = s.D(i, j);
    const int noise = a.obs_noise ? 1 : (a.obs_std ? 2 : 0);
    const bool mode_lane = l < 2 * M;                                 // lanes 0 .. 2M-1 also write the modes
    // this lane's elements of a row: V element l at column 2M + l, and (mode lanes) mode l at column l; one env = D
    // contiguous doubles, rows are E * D doubles apart
    const int64_t col_v = (int64_t)el * D + 2 * M + l, col_m = (int64_t)el * D + l, row_pitch = (int64_t)E * D;

    double lo = INFINITY, hi = -INFINITY;
    auto my_mode = [&]() {
        double mv = 0.0;
#pragma unroll
        for (int c = 0; c < 2 * M; ++c) mv = (l == c) ? ym[c] : mv;
        return mv;
    };
    auto emit = [&](int row) {
        const double mv = my_mode();
        double ov = v, om = mv;
        if (noise) {
            ov = v + obs_noise_at(a, row, el, 2 * M + l, D);
            if (mode_lane) om = mv + obs_noise_at(a, row, el, l, D);
        }
        if (live) {
            const int64_t r0 = (int64_t)row * row_pitch;
            if (a.states) {
                a.states[r0 + col_v] = v;
                if (mode_lane) a.states[r0 + col_m] = mv;
            }
            if (a.observations) {
                a.observations[r0 + col_v] = ov;
                if (mode_lane) a.observations[r0 + col_m] = om;
            }
            lo = fmin(lo, ov);
            hi = fmax(hi, ov);
            if (mode_lane) { lo = fmin(lo, om); hi = fmax(hi, om); }
        }
    };
    // fused reward: the group publishes its covariance elements, lane 0 evaluates the measure
    auto reward_row = [&](int row) {
        if (a.reward_kind == QRL_REWARD_NONE) return;
        __syncwarp();
        sV[0][g][l] = v;
        __syncwarp();
        if (live && l == 0) {
            double yfull[D];
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) yfull[c] = ym[c];
#pragma unroll
            for (int c = 0; c < L; ++c) yfull[2 * M + c] = sV[0][g][c];
            ode_reward_row_mem<M, Q>(a, row, e, yfull);
        }
        __syncwarp();
    };
    // one RHS evaluation at (modes ymt, my element vt) through exchange buffer `buf`: returns dV_ij, fills dm
    auto rhs_lane = [&](int buf, double t, const double* ymt, double vt, double* dm) -> double {
        double A[Q][Q];
        s.eval(t, ymt, dm, A);
        sV[buf][g][l] = vt;
        if (l == 0) {
            double2* dst = reinterpret_cast<double2*>(sA[buf][g]);
#pragma unroll
            for (int r = 0; r < Q; ++r)
#pragma unroll
                for (int c = 0; c < Q; c += 2) dst[(r * Q + c) / 2] = make_double2(A[r][c], A[r][c + 1]);
        }
        __syncwarp();
        const double2* Ai = reinterpret_cast<const double2*>(sA[buf][g] + i * Q);
        const double2* Aj = reinterpret_cast<const double2*>(sA[buf][g] + j * Q);
        const double2* Vi = reinterpret_cast<const double2*>(sV[buf][g] + i * Q);
        const double2 ai0 = Ai[0], ai1 = Ai[1], aj0 = Aj[0], aj1 = Aj[1], vi0 = Vi[0], vi1 = Vi[1];
        const double* Vc = sV[buf][g] + j;
        const double v0 = Vc[0], v1 = Vc[Q], v2 = Vc[2 * Q], v3 = Vc[3 * Q];
        // same order of operations as rhs(): (A V)_ij and (V A^T)_ij as two chains over k, then (av + va) + D_ij
        const double av = fma(ai1.y, v3, fma(ai1.x, v2, fma(ai0.y, v1, fma(ai0.x, v0, 0.0))));
        const double va = fma(vi1.y, aj1.y, fma(vi1.x, aj1.x, fma(vi0.y, aj0.y, fma(vi0.x, aj0.x, 0.0))));
        return (av + va) + dij;
    };

    emit(0);
    reward_row(0);
    const int nsub = a.substeps;
    const double h = a.t_ssz / (double)nsub, hh = 0.5 * h, h6 = h / 6.0;
    for (int row = 1; row <= K; ++row) {
        const double t_row = a.t0 + (double)(row - 1) * a.t_ssz;
        for (int sub = 0; sub < nsub; ++sub) {
            const double t = t_row + (double)sub * h;
            double dm[M2], ymt[M2], accm[M2];
            double k = rhs_lane(0, t, ym, v, dm);
            double accv = k, vt = fma(hh, k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = dm[c]; ymt[c] = fma(hh, dm[c], ym[c]); }
            k = rhs_lane(1, t + hh, ymt, vt, dm);
            accv = fma(2.0, k, accv); vt = fma(hh, k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(hh, dm[c], ym[c]); }
            k = rhs_lane(0, t + hh, ymt, vt, dm);
            accv = fma(2.0, k, accv); vt = fma(h, k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(h, dm[c], ym[c]); }
            k = rhs_lane(1, t + h, ymt, vt, dm);
            v = fma(h6, accv + k, v);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) ym[c] = fma(h6, accm[c] + dm[c], ym[c]);
        }
        emit(row);
        reward_row(row);
    }
    // carry-over of the last row
    {
        const double mv = my_mode();
        if (live) {
            if (a.y_last) {
                a.y_last[col_v] = v;
                if (mode_lane) a.y_last[col_m] = mv;
            }
            if (a.obs_last) {
                a.obs_last[col_v] = noise ? v + obs_noise_at(a, K, el, 2 * M + l, D) : v;
                if (mode_lane) a.obs_last[col_m] = noise ? mv + obs_noise_at(a, K, el, l, D) : mv;
            }
            if (a.nan_flag && !(isfinite(v) && isfinite(mv))) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
        }
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

// ------------------------------------------------------------------------------------------------------------------
// RK4 with q LANES per environment (num_quads 6 and 8: the "warp for larger mode counts" of the north_star).
// One thread per environment needs 2 q^2 + 5 (2m + q^2) doubles of state at q = 8 — 255 registers and 8-22 KB of
// spills (profiles/r01r_ptxas_resources.md: 6.4e8 sub-steps/s at E = 16384).  Here lane i of an environment owns ROW i
// of the covariance V and of the drift matrix A (RowSys, systems.cuh) and the replicated mode amplitudes: per RK stage
// the lanes publish their rows of V and A in shared memory (double buffered over the stages: one __syncwarp each), then
//     (A V)_ij   = sum_k A_ik V_kj   own A row from registers, row k of V as LDS.128 broadcasts
//     (V A^T)_ij = sum_k V_ik A_jk   own V row from registers, row j of A as LDS.128 broadcasts
// — 2 q^2 FMAs in 2q independent chains per lane and stage, same order of operations as rhs() (k ascending, then
// (av + va) + D_ij), so the trajectories are bit-identical to the thread-per-env kernel (the structural zeros it skips
// contribute exact zeros).  32 / q environments per warp; a row of the (K+1, E, d) blocks leaves as q-double runs.
// ------------------------------------------------------------------------------------------------------------------
template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) rk4_rows_kernel(const __grid_constant__ qrl_ode_args a) {
    constexpr int M = Sys::M, Q = Sys::Q, D = Dim<Sys>::D;
    constexpr int M2 = 2 * M > 0 ? 2 * M : 1;
    static_assert(Q % 2 == 0 && Q <= 8 && BLOCK % 32 == 0, "row lanes: even q <= 8");
    constexpr int EPW = 32 / Q, EPB = (BLOCK / 32) * EPW;              // envs per warp / per block
    // exchange tiles, padded: row pitch Q + 2 doubles (the 8 lanes of an environment store their 16-byte pieces into 8
    // distinct bank groups) and environment pitch + 2 (the 32 / Q environments of a warp read their pieces of the same
    // row from different banks) — with dense [Q][Q] tiles 58 % of the shared-memory wavefronts were conflicts (ncu r02j)
    constexpr int RPITCH = Q + 2, EPITCH = Q * RPITCH + 2;
    __shared__ __align__(16) double sA[2][EPB * EPITCH], sV[2][EPB * EPITCH];
    const int E = a.n_envs, K = a.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = lane / Q, i = lane - gw * Q;                        // env within the warp, my row
    const bool lane_used = gw < EPW;
    const int g = warp * EPW + (lane_used ? gw : 0);                   // env slot within the block
    const int e = blockIdx.x * EPB + g;
    const bool live = lane_used && e < E;
    const int el = e < E ? e : E - 1;

    RowSys<Sys> s;
    {
        const double* prm = a.params ? a.params + (int64_t)el * a.params_stride_e : nullptr;
        double us[4] = {0.0, 0.0, 0.0, 0.0};
        if (a.actions) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < a.n_actions) us[k] = a.actions[(int64_t)el * a.n_actions + k] * (a.action_scale ? a.action_scale[k] : 1.0);
        }
        s.load(prm, us, a.A ? a.A + (int64_t)el * a.A_stride_e : nullptr, a.D ? a.D + (int64_t)el * a.D_stride_e : nullptr, i);
    }
    double ym[M2], v[Q], drow[Q];
#pragma unroll
    for (int c = 0; c < 2 * M; ++c) ym[c] = a.y0[(int64_t)el * D + c];
#pragma unroll
    for (int j = 0; j < Q; ++j) { v[j] = a.y0[(int64_t)el * D + 2 * M + i * Q + j]; drow[j] = s.Drow(j); }
    const int noise = a.obs_noise ? 1 : (a.obs_std ? 2 : 0);
    // my elements of a row: V row i at columns 2M + i*Q .. +Q-1; lanes i < 2M also write mode component i
    const bool mode_lane = i < 2 * M;
    const int64_t col_v = (int64_t)el * D + 2 * M + i * Q, col_m = (int64_t)el * D + i, row_pitch = (int64_t)E * D;
    double lo = INFINITY, hi = -INFINITY;
    auto my_mode = [&]() {
        double mv = 0.0;
#pragma unroll
        for (int c = 0; c < 2 * M; ++c) mv = (i == c) ? ym[c] : mv;
        return mv;
    };
    auto emit = [&](int row, bool last) {
        const double mv = my_mode();
        double ov[Q], om = mv;
#pragma unroll
        for (int j = 0; j < Q; ++j) ov[j] = noise ? v[j] + obs_noise_at(a, row, el, 2 * M + i * Q + j, D) : v[j];
        if (noise && mode_lane) om = mv + obs_noise_at(a, row, el, i, D);
        if (live) {
            const int64_t r0 = (int64_t)row * row_pitch;
            if (a.states) {
#pragma unroll
                for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.states + r0 + col_v + j) = make_double2(v[j], v[j + 1]);
                if (mode_lane) a.states[r0 + col_m] = mv;
            }
            if (a.observations) {
#pragma unroll
                for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.observations + r0 + col_v + j) = make_double2(ov[j], ov[j + 1]);
                if (mode_lane) a.observations[r0 + col_m] = om;
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) { lo = fmin(lo, ov[j]); hi = fmax(hi, ov[j]); }
            if (mode_lane) { lo = fmin(lo, om); hi = fmax(hi, om); }
            if (last) {
                if (a.y_last) {
#pragma unroll
                    for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.y_last + col_v + j) = make_double2(v[j], v[j + 1]);
                    if (mode_lane) a.y_last[col_m] = mv;
                }
                if (a.obs_last) {
#pragma unroll
                    for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.obs_last + col_v + j) = make_double2(ov[j], ov[j + 1]);
                    if (mode_lane) a.obs_last[col_m] = om;
                }
                bool bad = !isfinite(mv);
#pragma unroll
                for (int j = 0; j < Q; ++j) bad |= !isfinite(v[j]);
                if (a.nan_flag && bad) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
            }
        }
    };
    // one right-hand side at (modes ymt, my row vt) through exchange buffer `buf`: fills dv (my row of dV/dt) and dm
    auto rhs_row = [&](int buf, double t, const double* ymt, const double (&vt)[Q], double (&dv)[Q], double* dm) {
        double Arow[Q];
        s.rates(t, ymt, dm);
        s.row(t, ymt, Arow);
        if (lane_used) {
#pragma unroll
            for (int j = 0; j < Q; j += 2) {
                *reinterpret_cast<double2*>(&sV[buf][g * EPITCH + i * RPITCH + j]) = make_double2(vt[j], vt[j + 1]);
                *reinterpret_cast<double2*>(&sA[buf][g * EPITCH + i * RPITCH + j]) = make_double2(Arow[j], Arow[j + 1]);
            }
        }
        __syncwarp();
        double av[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) av[j] = 0.0;
#pragma unroll
        for (int k = 0; k < Q; ++k) {          // (A V)_ij: row k of V, broadcast to the lanes of the env
            double vk[Q];
#pragma unroll
            for (int j = 0; j < Q; j += 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(&sV[buf][g * EPITCH + k * RPITCH + j]);
                vk[j] = t2.x; vk[j + 1] = t2.y;
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) av[j] = fma(Arow[k], vk[j], av[j]);
        }
#pragma unroll
        for (int j = 0; j < Q; ++j) {          // (V A^T)_ij: row j of A
            double aj[Q];
#pragma unroll
            for (int k = 0; k < Q; k += 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(&sA[buf][g * EPITCH + j * RPITCH + k]);
                aj[k] = t2.x; aj[k + 1] = t2.y;
            }
            double va = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) va = fma(vt[k], aj[k], va);
            dv[j] = (av[j] + va) + drow[j];    // same association as deterministic.py:713-732
        }
    };

    emit(0, K == 0);
    const int nsub = a.substeps;
    const double h = a.t_ssz / (double)nsub, hh = 0.5 * h, h6 = h / 6.0;
    for (int row = 1; row <= K; ++row) {
        const double t_row = a.t0 + (double)(row - 1) * a.t_ssz;
        for (int sub = 0; sub < nsub; ++sub) {
            const double t = t_row + (double)sub * h;
            double dm[M2], ymt[M2], accm[M2], k[Q], accv[Q], vt[Q];
            rhs_row(0, t, ym, v, k, dm);
#pragma unroll
            for (int j = 0; j < Q; ++j) { accv[j] = k[j]; vt[j] = fma(hh, k[j], v[j]); }
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = dm[c]; ymt[c] = fma(hh, dm[c], ym[c]); }
            rhs_row(1, t + hh, ymt, vt, k, dm);
#pragma unroll
            for (int j = 0; j < Q; ++j) { accv[j] = fma(2.0, k[j], accv[j]); vt[j] = fma(hh, k[j], v[j]); }
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(hh, dm[c], ym[c]); }
            rhs_row(0, t + hh, ymt, vt, k, dm);
#pragma unroll
            for (int j = 0; j < Q; ++j) { accv[j] = fma(2.0, k[j], accv[j]); vt[j] = fma(h, k[j], v[j]); }
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(h, dm[c], ym[c]); }
            rhs_row(1, t + h, ymt, vt, k, dm);
#pragma unroll
            for (int j = 0; j < Q; ++j) v[j] = fma(h6, accv[j] + k[j], v[j]);
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) ym[c] = fma(h6, accm[c] + dm[c], ym[c]);
        }
        emit(row, row == K);
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

// ------------------------------------------------------------------------------------------------------------------
// DOPRI5 with q LANES per environment (num_quads 6 and 8): the layout of rk4_rows_kernel — lane i of an environment owns
// row i of the covariance V and of the drift matrix A, the rows meet in padded shared-memory tiles once per right-hand
// side — with the method, controller and dense output of dopri5_kernel, element by element in the same order of
// operations.  One thread per environment keeps 7 d stage derivatives in shared memory (129 KB per warp at q = 8: ONE warp
// per SM) next to 255 registers and 1.2 KB of spills; here a lane keeps the seven derivatives of its own row (7 q shared-
// memory columns) and of ONE mode component (7 registers: lane c < 2m owns mode component c; the stage arguments of the
// modes are computed by their owners and gathered with 2m shuffles), 99 KB per 128-thread block = 8 warps per SM, nothing
// spills.  The per-env controller is uniform over the q lanes of an environment; the 32 / q environments of a warp
// diverge only against each other, and every warp-level primitive carries the environment's lane mask.  The error norm is
// the one reduction across the lanes (q shuffles, the same order in every lane; the thread kernel sums over d
// sequentially, so the two agree to rounding, not bit for bit — like dopri5_lanes_kernel).
// ------------------------------------------------------------------------------------------------------------------
template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) dopri5_rows_kernel(const __grid_constant__ qrl_ode_args a) {
    constexpr int M = Sys::M, Q = Sys::Q, D = Dim<Sys>::D;
    constexpr int M2 = 2 * M > 0 ? 2 * M : 1;
    static_assert(Q % 2 == 0 && Q <= 8 && BLOCK % 32 == 0 && 2 * M <= Q, "row lanes: even q <= 8, one mode component per lane");
    constexpr int EPW = 32 / Q, EPB = (BLOCK / 32) * EPW;              // envs per warp / per block
    constexpr int RPITCH = Q + 2, EPITCH = Q * RPITCH + 2;             // padded exchange tiles (see rk4_rows_kernel)
    __shared__ __align__(16) double sA[2][EPB * EPITCH], sV[2][EPB * EPITCH];
    extern __shared__ double ks_rows[];                                // [7][Q + 1][BLOCK]: stage derivatives of my row (+ my mode component)
#define KS(s, j) ks_rows[((s) * Q + (j)) * BLOCK + threadIdx.x]
    const int E = a.n_envs, K = a.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = lane / Q, i = lane - gw * Q;                        // env within the warp, my row
    const bool lane_used = gw < EPW;
    const int g = warp * EPW + (lane_used ? gw : 0);                   // env slot within the block
    const int e = blockIdx.x * EPB + g;
    const bool live = lane_used && e < E;
    const int lane0 = gw * Q;                                          // first lane of my environment
    const unsigned emask = lane_used ? (((1u << Q) - 1u) << lane0) : 0u;
    double lo = INFINITY, hi = -INFINITY;

    if (live) {
        RowSys<Sys> s;
        {
            const double* prm = a.params ? a.params + (int64_t)e * a.params_stride_e : nullptr;
            double us[4] = {0.0, 0.0, 0.0, 0.0};
            if (a.actions) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < a.n_actions) us[k] = a.actions[(int64_t)e * a.n_actions + k] * (a.action_scale ? a.action_scale[k] : 1.0);
            }
            s.load(prm, us, a.A ? a.A + (int64_t)e * a.A_stride_e : nullptr, a.D ? a.D + (int64_t)e * a.D_stride_e : nullptr, i);
        }
        double ym[M2], v[Q], drow[Q];
#pragma unroll
        for (int c = 0; c < M2; ++c) ym[c] = 2 * M > 0 ? a.y0[(int64_t)e * D + c] : 0.0;
#pragma unroll
        for (int j = 0; j < Q; ++j) { v[j] = a.y0[(int64_t)e * D + 2 * M + i * Q + j]; drow[j] = s.Drow(j); }
        const int noise = a.obs_noise ? 1 : (a.obs_std ? 2 : 0);
        const bool mode_lane = i < 2 * M;
        const int64_t col_v = (int64_t)e * D + 2 * M + i * Q, col_m = (int64_t)e * D + i, row_pitch = (int64_t)E * D;
        auto my_mode = [&](const double* m) {
            double mv = 0.0;
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) mv = (i == c) ? m[c] : mv;
            return mv;
        };
        // the mode amplitudes of a stage argument: every owner lane holds its component, all lanes of the env need all
        auto gather_modes = [&](double own, double* out) {
#pragma unroll
            for (int c = 0; c < 2 * M; ++c) out[c] = __shfl_sync(emask, own, lane0 + c);
        };
        auto emit = [&](int row, const double (&vv)[Q], double mv) {
            double ov[Q], om = mv;
#pragma unroll
            for (int j = 0; j < Q; ++j) ov[j] = noise ? vv[j] + obs_noise_at(a, row, e, 2 * M + i * Q + j, D) : vv[j];
            if (noise && mode_lane) om = mv + obs_noise_at(a, row, e, i, D);
            const int64_t r0 = (int64_t)row * row_pitch;
            if (a.states) {
#pragma unroll
                for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.states + r0 + col_v + j) = make_double2(vv[j], vv[j + 1]);
                if (mode_lane) a.states[r0 + col_m] = mv;
            }
            if (a.observations) {
#pragma unroll
                for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.observations + r0 + col_v + j) = make_double2(ov[j], ov[j + 1]);
                if (mode_lane) a.observations[r0 + col_m] = om;
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) { lo = fmin(lo, ov[j]); hi = fmax(hi, ov[j]); }
            if (mode_lane) { lo = fmin(lo, om); hi = fmax(hi, om); }
            if (row == K && a.obs_last) {
#pragma unroll
                for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.obs_last + col_v + j) = make_double2(ov[j], ov[j + 1]);
                if (mode_lane) a.obs_last[col_m] = om;
            }
        };
        // one right-hand side at (modes ymt, my row vt) through exchange buffer `buf`: dv = my row of dV/dt, dm = d(modes)/dt
        int buf = 0;
        auto rhs_row = [&](double t, const double* ymt, const double (&vt)[Q], double (&dv)[Q], double* dm) {
            double Arow[Q];
            s.rates(t, ymt, dm);
            s.row(t, ymt, Arow);
#pragma unroll
            for (int j = 0; j < Q; j += 2) {
                *reinterpret_cast<double2*>(&sV[buf][g * EPITCH + i * RPITCH + j]) = make_double2(vt[j], vt[j + 1]);
                *reinterpret_cast<double2*>(&sA[buf][g * EPITCH + i * RPITCH + j]) = make_double2(Arow[j], Arow[j + 1]);
            }
            __syncwarp(emask);
            double av[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) av[j] = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) {          // (A V)_ij: row k of V, broadcast to the lanes of the env
                double vk[Q];
#pragma unroll
                for (int j = 0; j < Q; j += 2) {
                    const double2 t2 = *reinterpret_cast<const double2*>(&sV[buf][g * EPITCH + k * RPITCH + j]);
                    vk[j] = t2.x; vk[j + 1] = t2.y;
                }
#pragma unroll
                for (int j = 0; j < Q; ++j) av[j] = fma(Arow[k], vk[j], av[j]);
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) {          // (V A^T)_ij: row j of A
                double aj[Q];
#pragma unroll
                for (int k = 0; k < Q; k += 2) {
                    const double2 t2 = *reinterpret_cast<const double2*>(&sA[buf][g * EPITCH + j * RPITCH + k]);
                    aj[k] = t2.x; aj[k + 1] = t2.y;
                }
                double va = 0.0;
#pragma unroll
                for (int k = 0; k < Q; ++k) va = fma(vt[k], aj[k], va);
                dv[j] = (av[j] + va) + drow[j];    // same association as deterministic.py:713-732
            }
            buf ^= 1;     // the next right-hand side writes the other pair of tiles: its __syncwarp orders this one's reads
        };

        emit(0, v, my_mode(ym));
        const double t_end = a.t0 + (double)K * a.t_ssz;
        double t = a.t0;
        double h = (a.h_state && a.h_state[e] > 0.0) ? a.h_state[e] : a.t_ssz;
        int next_row = 1, n_acc = 0, n_rej = 0, guard = 0;
        bool rejected = false;
        double ymo = my_mode(ym);              // MY mode component (lanes i < 2m) ...
#define KMO(s) ks_rows[((7 * Q) + (s)) * BLOCK + threadIdx.x]   /* ... and its stage derivatives */
        int st0 = 0;                           // the first step of the launch also evaluates k_1 = f(t, y) (stage "0")
#define KV(s) KS(s, j)
#define DP_ERR(KK) (ht * fma(dp::E7, KK(6), fma(dp::E6, KK(5), fma(dp::E5, KK(4), fma(dp::E4, KK(3), fma(dp::E3, KK(2), dp::E1 * KK(0)))))))
        while (next_row <= K && guard++ < dp::MAX_STEPS) {
            const double remaining = t_end - t;
            const bool last = h >= remaining * (1.0 - 1e-12);
            const double ht = last ? remaining : h;
            double vt[Q], ymt[M2];
            // Stage arguments 2 .. 6 and the 5th-order solution (= stage-7 argument, FSAL) as ONE loop body over the rows of
            // the tableau: y + ht (((a_s0 k_0) + a_s1 k_1) + ...) accumulated in the order of dopri5_kernel's nested FMAs
            // (zero coefficients add exact zeros; stage 2 keeps its own form fma(ht a_21, k_0, y)).  Fully unrolled, the step
            // was 11 400 instructions (183 KB) and the warps spent 3.2 stall cycles per issue waiting for instruction fetch
            // (profiles/r02z_ode_ncu_summary.md); the loop is a third of that.
#pragma unroll 1
            for (int st = st0; st <= 6; ++st) {
                double accm = ymo, ts = t;
                if (st == 0) {                 // (one inlined copy of the right-hand side for all seven evaluations)
#pragma unroll
                    for (int j = 0; j < Q; ++j) vt[j] = v[j];
                } else {
                    double acc[Q];
                    accm = dp::AS[st - 1][0] * KMO(0);
#pragma unroll
                    for (int j = 0; j < Q; ++j) acc[j] = dp::AS[st - 1][0] * KS(0, j);
#pragma unroll 1
                    for (int r = 1; r < st; ++r) {
                        const double c = dp::AS[st - 1][r];
#pragma unroll
                        for (int j = 0; j < Q; ++j) acc[j] = fma(c, KS(r, j), acc[j]);
                        accm = fma(c, KMO(r), accm);
                    }
                    if (st == 1) {
#pragma unroll
                        for (int j = 0; j < Q; ++j) vt[j] = fma(ht * dp::A21, KS(0, j), v[j]);
                        accm = fma(ht * dp::A21, KMO(0), ymo);
                    } else {
#pragma unroll
                        for (int j = 0; j < Q; ++j) vt[j] = fma(ht, acc[j], v[j]);
                        accm = fma(ht, accm, ymo);
                    }
                    ts = t + dp::CS[st - 1] * ht;
                }
                if (2 * M > 0) gather_modes(accm, ymt);
                double dv[Q], dm[M2];
                rhs_row(ts, ymt, vt, dv, dm);
#pragma unroll
                for (int j = 0; j < Q; ++j) KS(st, j) = dv[j];
                KMO(st) = my_mode(dm);
            }
            st0 = 1;
            const double ymo_new = my_mode(ymt);
            // error estimate, RMS norm with scale atol + rtol * max(|y|, |y_new|): my row (+ my mode component), then the
            // sum over the env's lanes in lane order
            double sq = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const double ev = DP_ERR(KV);
                const double sc = fma(a.rtol, fmax(fabs(v[j]), fabs(vt[j])), a.atol);
                const double r = ev / sc;
                sq = fma(r, r, sq);
            }
            if (mode_lane) {
                const double ev = DP_ERR(KMO);
                const double sc = fma(a.rtol, fmax(fabs(ymo), fabs(ymo_new)), a.atol);
                const double r = ev / sc;
                sq = fma(r, r, sq);
            }
            double tot = 0.0;
#pragma unroll
            for (int c = 0; c < Q; ++c) tot += __shfl_sync(emask, sq, lane0 + c);
            const double err = sqrt(tot / (double)D);
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
                    double vo[Q], mo = ymo_new;
#pragma unroll
                    for (int j = 0; j < Q; ++j) vo[j] = vt[j];
                    if (tg != t_new) {
                        const double x = (tg - t) / ht;
                        double pw[7];
#pragma unroll
                        for (int st = 0; st < 7; ++st)
                            pw[st] = fma(fma(fma(dp::P[st][3], x, dp::P[st][2]), x, dp::P[st][1]), x, dp::P[st][0]);
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            double q = 0.0;
#pragma unroll
                            for (int st = 0; st < 7; ++st) {
                                if (st == 1) continue;
                                q = fma(KS(st, j), pw[st], q);
                            }
                            vo[j] = fma(ht * x, q, v[j]);
                        }
                        if (mode_lane) {
                            double qm = 0.0;
#pragma unroll
                            for (int st = 0; st < 7; ++st) {
                                if (st == 1) continue;
                                qm = fma(KMO(st), pw[st], qm);
                            }
                            mo = fma(ht * x, qm, ymo);
                        }
                    }
                    emit(next_row, vo, mo);
                    ++next_row;
                }
                t = t_new;
#pragma unroll
                for (int j = 0; j < Q; ++j) { v[j] = vt[j]; KS(0, j) = KS(6, j); }
#pragma unroll
                for (int c = 0; c < 2 * M; ++c) ym[c] = ymt[c];
                ymo = ymo_new;
                KMO(0) = KMO(6);
                // a step that was clipped to the interval end does not shrink the carried step size
                h = (last && fac >= 1.0) ? fmax(h, ht * fac) : ht * fac;
            } else {
                h = ht * fmax(dp::MIN_FACTOR, dp::SAFETY * pow(err, -0.2));
                rejected = true;
                ++n_rej;
            }
        }
#undef DP_ERR
#undef KV
#undef KMO
        if (a.y_last) {
#pragma unroll
            for (int j = 0; j < Q; j += 2) *reinterpret_cast<double2*>(a.y_last + col_v + j) = make_double2(v[j], v[j + 1]);
            if (mode_lane) a.y_last[col_m] = ymo;
        }
        if (i == 0) {
            if (a.h_state) a.h_state[e] = h;
            if (a.n_steps) { a.n_steps[2 * e] += n_acc; a.n_steps[2 * e + 1] += n_rej; }
        }
        bool bad = next_row <= K || !isfinite(ymo);   // step budget exhausted / non-finite state
#pragma unroll
        for (int j = 0; j < Q; ++j) bad |= !isfinite(v[j]);
        if (a.nan_flag && bad) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
    }
#undef KS
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

// ------------------------------------------------------------------------------------------------------------------
// RK4, warp-specialised small-batch kernel (q = 4).  With E = 1024 there is at most one warp per scheduler, so every
// instruction of a lane's loop is on the critical path.  The classical mode amplitudes do not depend on the
// covariance, and writing a row does not feed back into the integration, so both are taken OUT of the covariance
// lanes' dependency chain.  Per CTA of EPB environments, time cut into chunks of C sub-steps, three roles in a 3-stage
// pipeline with one __syncthreads per chunk (like the tiled EM kernel):
//   iteration it:  producer warp   chunk it      one lane per env integrates the modes with RK4 and publishes, per RK
//                                                stage, the drift matrix A (8 pieces of 16 bytes), and per sub-step the
//                                                new amplitudes (into the row ring)
//                  consumer warps  chunk it-1    EPB/2 warps, q*q = 16 lanes per env: lane (i,j) owns V_ij; a stage is
//                                                STS V_ij -> __syncwarp -> LDS row i / column j -> two 4-FMA chains ->
//                                                RK update (the A rows are loaded before the warp barrier); after a
//                                                sub-step the lane drops V_ij into the row ring: ONE store
//                  4 writer warps  chunk it-2    rows [modes | V] of the CTA's envs leave the ring as contiguous runs of
//                                                16-byte pieces (+ measurement noise, min/max, fused reward, carry-over)
// Same operations in the same order as rk4_lanes_kernel / rk4_kernel: results are bit-identical.
// ------------------------------------------------------------------------------------------------------------------
template <class Sys, int EPB>
__global__ void __launch_bounds__(EPB * 16 + 32 + 128) rk4_split_kernel(const __grid_constant__ qrl_ode_args a, const int C) {
    constexpr int M = Sys::M, Q = Sys::Q, D = Dim<Sys>::D, L = Q * Q, D2 = D / 2;
    constexpr int M2 = 2 * M > 0 ? 2 * M : 1;
    static_assert(L == 16 && EPB % 2 == 0 && EPB <= 32 && D % 2 == 0, "q = 4, an even number of envs per CTA");
    constexpr int NV = EPB * L;              // consumer threads
    constexpr int PP = EPB + 1;              // piece pitch of the A ring in 16-byte units: conflict-free writes and reads
    constexpr int RW = EPB * D;              // doubles per ring row: the CTA's envs, [modes | V] each
    extern __shared__ double2 ring2[];
    // A ring: [2 buffers][C sub-steps * 4 stages][8 pieces][PP]   (producer -> consumers)
    // row ring: [3 buffers][C sub-steps][EPB][D]                   (producer: modes, consumers: V  -> writer)
    double2* Ar = ring2;
    double* Rr = reinterpret_cast<double*>(ring2 + (size_t)2 * C * 4 * 8 * PP);
    __shared__ __align__(16) double sV[2][EPB][L];   // covariance exchange, double buffered over the stages
    const int E = a.n_envs, K = a.K;
    const int tid = threadIdx.x;
    const int role = tid < NV ? 0 : (tid < NV + 32 ? 1 : 2);   // 0 consumer, 1 producer, 2 writer
    const int e0 = blockIdx.x * EPB;
    const int n_live = min(EPB, E - e0);
    const int nsub = a.substeps;
    const int NS = K * nsub;                 // sub-steps of the interval
    const int NCH = (NS + C - 1) / C;
    const double h = a.t_ssz / (double)nsub, hh = 0.5 * h, h6 = h / 6.0;
    double lo = INFINITY, hi = -INFINITY;

    // ---- producer state: one lane per env ----
    const int pg = tid - NV;
    const bool p_on = role == 1 && pg < EPB;
    Sys s;
    double ym[M2];
    if (p_on) {
        const int el = min(e0 + pg, E - 1);
        load_system(s, a, el);
#pragma unroll
        for (int c = 0; c < 2 * M; ++c) ym[c] = a.y0[(int64_t)el * D + c];
    }
    int p_row = 1, p_sub = 0;                // row being integrated towards, sub-step inside it
    // ---- consumer state: lane (i, j) of env g ----
    const int g = tid / L, l = tid % L;
    const int i = l / Q, j = l % Q;
    double v = 0.0, dij = 0.0;               // my covariance element, my D_ij
    if (role == 0) {
        const int el = min(e0 + g, E - 1);
        Sys sc;
        load_system(sc, a, el);
        dij = sc.D(i, j);
        v = a.y0[(int64_t)el * D + 2 * M + l];
    }
    const int oi0 = (2 * i) * PP + g, oi1 = (2 * i + 1) * PP + g, oj0 = (2 * j) * PP + g, oj1 = (2 * j + 1) * PP + g;
    int par = 0;                             // exchange-buffer parity of the next stage
    auto stage = [&](const double2* A2, double vt) -> double {
        sV[par][g][l] = vt;
        const double2 ai0 = A2[oi0], ai1 = A2[oi1], aj0 = A2[oj0], aj1 = A2[oj1];   // independent of the exchange
        __syncwarp();
        const double2* Vi = reinterpret_cast<const double2*>(sV[par][g] + i * Q);
        const double2 vi0 = Vi[0], vi1 = Vi[1];
        const double* Vc = sV[par][g] + j;
        const double v0 = Vc[0], v1 = Vc[Q], v2 = Vc[2 * Q], v3 = Vc[3 * Q];
        par ^= 1;
        // same order of operations as rhs(): (A V)_ij and (V A^T)_ij as two chains over k, then (av + va) + D_ij
        const double av = fma(ai1.y, v3, fma(ai1.x, v2, fma(ai0.y, v1, fma(ai0.x, v0, 0.0))));
        const double va = fma(vi1.y, aj1.y, fma(vi1.x, aj1.x, fma(vi0.y, aj0.y, fma(vi0.x, aj0.x, 0.0))));
        return (av + va) + dij;
    };
    // ---- writer: n_rows grid rows (first one = row_first, in ring slot slot_first, then every nsub-th slot) --------
    // One flat walk over the 16-byte pieces of all rows of a chunk (a row = n_live * D2 contiguous pieces in the ring
    // and in the (K+1, E, D) blocks), one piece per lane and trip; NaN-ignoring min/max by compare+select.
    constexpr int NW = 128;                  // writer threads (4 warps: one warp needs ~210 cycles per trip of this walk)
    const int lane = tid - NV - 32;          // writer thread index
    const int RL2 = n_live * D2;             // live 16-byte pieces per row
    const int noise = a.obs_noise ? 1 : (a.obs_std ? 2 : 0);
    const int64_t row_pitch2 = (int64_t)E * D2;
    auto write_rows = [&](const double* Rslot, int row_first, int n_rows) {
        const double2* R2 = reinterpret_cast<const double2*>(Rslot);
        const int ring_pitch2 = nsub * (RW / 2);                 // ring pieces between two rows
        int r = 0, p = lane;                                     // row within the call, piece within the row
        while (p >= RL2) { p -= RL2; ++r; }
        int64_t g2 = ((int64_t)row_first * E + e0) * D2 + (int64_t)r * row_pitch2 + p;   // piece index in the global blocks
        int s2 = r * ring_pitch2 + p;                            // piece index in the ring
        while (r < n_rows) {
            const double2 x = R2[s2];
            const int row = row_first + r;
            double2 o = x;
            if (noise == 1) {
                const double2 z = reinterpret_cast<const double2*>(a.obs_noise)[g2];
                o.x = x.x + z.x;
                o.y = x.y + z.y;
            } else if (noise == 2) {
                const int elx = p / D2, c = p - elx * D2;
                const double2 z = philox_normal2_call(a.seed, (uint32_t)(a.t_idx + row), a.episode,
                                                      (uint32_t)(a.env_offset + e0 + elx), 0x10000u + (uint32_t)c);
                const double* sd = a.obs_std + (int64_t)(e0 + elx) * a.obs_std_stride_e + 2 * c;
                o.x = __dadd_rn(x.x, __dmul_rn(sd[0], z.x));
                o.y = __dadd_rn(x.y, __dmul_rn(sd[1], z.y));
            }
            if (a.states) reinterpret_cast<double2*>(a.states)[g2] = x;
            if (a.observations) reinterpret_cast<double2*>(a.observations)[g2] = o;
            if (row == K) {
                const int64_t c2 = (int64_t)e0 * D2 + p;
                if (a.y_last) reinterpret_cast<double2*>(a.y_last)[c2] = x;
                if (a.obs_last) reinterpret_cast<double2*>(a.obs_last)[c2] = o;
                if (a.nan_flag && !(isfinite(x.x) && isfinite(x.y))) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
            }
            lo = (o.x < lo) ? o.x : lo;
            lo = (o.y < lo) ? o.y : lo;
            hi = (o.x > hi) ? o.x : hi;
            hi = (o.y > hi) ? o.y : hi;
            p += NW; g2 += NW; s2 += NW;
            while (p >= RL2 && r < n_rows) { p -= RL2; ++r; g2 += row_pitch2 - RL2; s2 += ring_pitch2 - RL2; }
        }
        // fused reward: one (row, env) evaluation per writer thread and trip
        if (a.reward_kind != QRL_REWARD_NONE)
            for (int idx = lane; idx < n_rows * n_live; idx += NW) {
                const int rr = idx / n_live, elx = idx - rr * n_live;
                ode_reward_row_mem<M, Q>(a, row_first + rr, e0 + elx, Rslot + (size_t)rr * nsub * RW + elx * D);
            }
    };

    // ---- row 0: the initial state goes through slot 0 of row buffer 2 (free until iteration 2) ----
    {
        double* R0 = Rr + (size_t)2 * C * RW;
        for (int idx = tid; idx < n_live * D; idx += EPB * 16 + 32 + NW) R0[idx] = a.y0[(int64_t)e0 * D + idx];
        __syncthreads();
        if (role == 2) write_rows(R0, 0, 1);
        __syncthreads();
    }
    int w_row = 0, w_until = nsub;           // writer: last row written, sub-steps until the next one
    for (int it = 0; it < NCH + 2; ++it) {
        if (role == 1) {
            if (p_on && it < NCH && !(a.flags & 8)) {
                // ---------------- producer: chunk `it` -> A buffer it & 1, modes into row buffer it % 3 ----------------
                const int gs0 = it * C, gs1 = min(gs0 + C, NS);
                double2* Ab = Ar + (size_t)(it & 1) * C * 4 * 8 * PP;
                double* Rb = Rr + (size_t)(it % 3) * C * RW + pg * D;
                auto publish = [&](int slot, const double (&A)[Q][Q]) {
                    double2* dst = Ab + (size_t)slot * 8 * PP + pg;
#pragma unroll
                    for (int r = 0; r < Q; ++r)
#pragma unroll
                        for (int c = 0; c < Q; c += 2) dst[((r * Q + c) / 2) * PP] = make_double2(A[r][c], A[r][c + 1]);
                };
                for (int gs = gs0; gs < gs1; ++gs) {
                    const double t = (a.t0 + (double)(p_row - 1) * a.t_ssz) + (double)p_sub * h;
                    if (++p_sub == nsub) { p_sub = 0; ++p_row; }
                    const int slot = (gs - gs0) * 4;
                    double A[Q][Q], dm[M2], ymt[M2], accm[M2];
                    s.eval(t, ym, dm, A);
                    publish(slot, A);
#pragma unroll
                    for (int c = 0; c < 2 * M; ++c) { accm[c] = dm[c]; ymt[c] = fma(hh, dm[c], ym[c]); }
                    s.eval(t + hh, ymt, dm, A);
                    publish(slot + 1, A);
#pragma unroll
                    for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(hh, dm[c], ym[c]); }
                    s.eval(t + hh, ymt, dm, A);
                    publish(slot + 2, A);
#pragma unroll
                    for (int c = 0; c < 2 * M; ++c) { accm[c] = fma(2.0, dm[c], accm[c]); ymt[c] = fma(h, dm[c], ym[c]); }
                    s.eval(t + h, ymt, dm, A);
                    publish(slot + 3, A);
#pragma unroll
                    for (int c = 0; c < 2 * M; ++c) {
                        ym[c] = fma(h6, accm[c] + dm[c], ym[c]);
                        Rb[(gs - gs0) * RW + c] = ym[c];
                    }
                }
            }
        } else if (role == 0) {
            if (it >= 1 && it <= NCH && !(a.flags & 16)) {
                // ---------------- consumers: chunk `it - 1` ----------------
                const int cc = it - 1;
                const int ns = min(C, NS - cc * C);
                const double2* A2 = Ar + (size_t)(cc & 1) * C * 4 * 8 * PP;
                double* Rv = Rr + (size_t)(cc % 3) * C * RW + g * D + 2 * M + l;
                for (int ss = 0; ss < ns; ++ss, A2 += 32 * PP, Rv += RW) {
                    double k = stage(A2, v);
                    double accv = k;
                    k = stage(A2 + 8 * PP, fma(hh, k, v));
                    accv = fma(2.0, k, accv);
                    k = stage(A2 + 16 * PP, fma(hh, k, v));
                    accv = fma(2.0, k, accv);
                    k = stage(A2 + 24 * PP, fma(h, k, v));
                    v = fma(h6, accv + k, v);
                    *Rv = v;
                }
            }
        } else if (it >= 2 && !(a.flags & 32)) {
            // ---------------- writer: chunk `it - 2` ----------------
            const int cw = it - 2;
            const int ns = min(C, NS - cw * C);
            const double* Rb = Rr + (size_t)(cw % 3) * C * RW;
            // rows of this chunk: the sub-steps ss with (cw * C + ss + 1) % nsub == 0
            if (w_until <= ns) {
                const int n_rows = (ns - w_until) / nsub + 1;
                write_rows(Rb + (size_t)(w_until - 1) * RW, w_row + 1, n_rows);
                w_row += n_rows;
                w_until += n_rows * nsub;
            }
            w_until -= ns;
        }
        __syncthreads();
    }
    if (a.obs_minmax) block_minmax_commit(lo, hi, a.obs_minmax);
}

// ------------------------------------------------------------------------------------------------------------------
// The 16 lanes of one environment (q = 4): shared plumbing of the adaptive lanes kernels (dopri5_lanes_kernel below,
// erk_lanes_kernel in ode_erk.cuh).  Lane (i,j) owns V_ij and a replica of the classical modes.  Every warp-level
// primitive carries the env's 16-lane mask: the two envs of a warp run their own controllers and may diverge.
// ------------------------------------------------------------------------------------------------------------------
template <class Sys, int EPB>
struct LaneEnv {
    static constexpr int M = Sys::M, Q = Sys::Q, D = 2 * Sys::M + Sys::Q * Sys::Q, L = Sys::Q * Sys::Q;
    static constexpr int M2 = 2 * M > 0 ? 2 * M : 1;
    static_assert(L == 16, "the lanes kernels are written for q = 4");
    static constexpr int RB = L;                         // rows buffered per env for the fused reward
    struct Smem {
        alignas(16) double A[2][EPB][L], V[2][EPB][L];   // exchange tiles, double buffered over the RHS evaluations
        alignas(16) double rows[EPB][RB][D];             // rows [modes | V] waiting for their fused reward
        int ridx[EPB][RB];                               // ... and their grid-row indices
    };
    const qrl_ode_args& a;
    Smem& sm;
    Sys s;
    int g, l, i, j, e, el, par = 0, noise, nbuf = 0;
    unsigned gmask;
    bool live, mode_lane;
    int64_t col_v, row_pitch;
    double dij, lo = INFINITY, hi = -INFINITY;

    __device__ __forceinline__ LaneEnv(const qrl_ode_args& args, Smem& smem) : a(args), sm(smem) {
        g = threadIdx.x / L; l = threadIdx.x % L; i = l / Q; j = l % Q;
        e = blockIdx.x * EPB + g;
        live = e < a.n_envs;
        el = live ? e : a.n_envs - 1;
        gmask = 0xFFFFu << (threadIdx.x & 16);
        mode_lane = l < 2 * M;
        load_system(s, a, el);
        dij = s.D(i, j);
        noise = a.obs_noise ? 1 : (a.obs_std ? 2 : 0);
        col_v = (int64_t)el * D + 2 * M + l;
        row_pitch = (int64_t)a.n_envs * D;
    }
    // my own element of a replicated mode vector (lanes < 2M)
    __device__ __forceinline__ double pick_mode(const double* m) const {
        double mv = 0.0;
#pragma unroll
        for (int c = 0; c < 2 * M; ++c) mv = (l == c) ? m[c] : mv;
        return mv;
    }
    // row `row` of this env: my covariance element vv and (mode lanes) my mode amplitude mv
    __device__ __forceinline__ void emit(int row, double vv, double mv) {
        double ov = vv, om = mv;
        if (noise) {
            ov = vv + obs_noise_at(a, row, el, 2 * M + l, D);
            if (mode_lane) om = mv + obs_noise_at(a, row, el, l, D);
        }
        const int64_t r0 = (int64_t)row * row_pitch + col_v;
        if (a.states) {
            a.states[r0] = vv;
            if (mode_lane) a.states[r0 - 2 * M] = mv;
        }
        if (a.observations) {
            a.observations[r0] = ov;
            if (mode_lane) a.observations[r0 - 2 * M] = om;
        }
        if (row == a.K && a.obs_last) {
            a.obs_last[col_v] = ov;
            if (mode_lane) a.obs_last[col_v - 2 * M] = om;
        }
        lo = fmin(lo, ov);
        hi = fmax(hi, ov);
        if (mode_lane) { lo = fmin(lo, om); hi = fmax(hi, om); }
        // fused reward: rows are collected and evaluated RB at a time, one row per lane (one lane evaluating every
        // row as it appears put ~400 cycles of measure arithmetic per row on the env's critical path)
        if (a.reward_kind != QRL_REWARD_NONE) {
            sm.rows[g][nbuf][2 * M + l] = vv;
            if (mode_lane) sm.rows[g][nbuf][l] = mv;
            if (l == 0) sm.ridx[g][nbuf] = row;
            if (++nbuf == RB) flush_rewards();
        }
    }
    __device__ __forceinline__ void flush_rewards() {
        __syncwarp(gmask);
        if (l < nbuf) ode_reward_row_mem<M, Q>(a, sm.ridx[g][l], e, sm.rows[g][l]);
        __syncwarp(gmask);
        nbuf = 0;
    }
    // one RHS evaluation at (modes ymt, my element vt): returns dV_ij, fills dm  (cf. rk4_lanes_kernel; same order of
    // operations as rhs())
    __device__ __forceinline__ double rhs(double t, const double* ymt, double vt, double* dm) {
        double A[Q][Q];
        s.eval(t, ymt, dm, A);
        sm.V[par][g][l] = vt;
        if (l == 0) {
            double2* dst = reinterpret_cast<double2*>(sm.A[par][g]);
#pragma unroll
            for (int r = 0; r < Q; ++r)
#pragma unroll
                for (int c = 0; c < Q; c += 2) dst[(r * Q + c) / 2] = make_double2(A[r][c], A[r][c + 1]);
        }
        __syncwarp(gmask);
        const double2* Ai = reinterpret_cast<const double2*>(sm.A[par][g] + i * Q);
        const double2* Aj = reinterpret_cast<const double2*>(sm.A[par][g] + j * Q);
        const double2* Vi = reinterpret_cast<const double2*>(sm.V[par][g] + i * Q);
        const double2 ai0 = Ai[0], ai1 = Ai[1], aj0 = Aj[0], aj1 = Aj[1], vi0 = Vi[0], vi1 = Vi[1];
        const double* Vc = sm.V[par][g] + j;
        const double v0 = Vc[0], v1 = Vc[Q], v2 = Vc[2 * Q], v3 = Vc[3 * Q];
        par ^= 1;
        const double av = fma(ai1.y, v3, fma(ai1.x, v2, fma(ai0.y, v1, fma(ai0.x, v0, 0.0))));
        const double va = fma(vi1.y, aj1.y, fma(vi1.x, aj1.x, fma(vi0.y, aj0.y, fma(vi0.x, aj0.x, 0.0))));
        return (av + va) + dij;
    }
    // sum over the env's lanes, the same value in every lane (the controller must not diverge inside an env)
    __device__ __forceinline__ double env_sum(double x) const {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) x += __shfl_xor_sync(gmask, x, o, 16);
        return __shfl_sync(gmask, x, 0, 16);
    }
    // end of the launch: carry-over of the last state, controller state, failure flag
    __device__ __forceinline__ void finish(double v, const double* ym, double h, bool has_h, int n_acc, int n_rej, bool failed) {
        if (nbuf > 0) flush_rewards();
        const double mv = pick_mode(ym);
        if (a.y_last) {
            a.y_last[col_v] = v;
            if (mode_lane) a.y_last[col_v - 2 * M] = mv;
        }
        if (l == 0) {
            if (has_h && a.h_state) a.h_state[e] = h;
            if (a.n_steps) { a.n_steps[2 * e] += n_acc; a.n_steps[2 * e + 1] += n_rej; }
        }
        if (a.nan_flag && (failed || !isfinite(v) || !isfinite(mv))) atomicOr(a.nan_flag, QRL_STATUS_NONFINITE);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// DOPRI5, q*q lanes per environment (q = 4).  Same method, controller, dense output and order of operations as
// dopri5_kernel, in the layout of rk4_lanes_kernel: lane (i,j) owns V_ij (and the classical modes, replicated), so the
// seven stage derivatives of a step are 7 + 7*2M REGISTERS per lane instead of 7*d shared-memory columns per env (41 KB
// per warp = 5 warps per SM in dopri5_kernel), a row is d contiguous doubles written by the env's own lanes, and the
// per-env controller is uniform over a half-warp: the two envs of a warp diverge only against each other.  The error
// norm is the one reduction across the env's lanes (4 xor-shuffle steps; the thread kernel sums sequentially over d, so
// the two agree to rounding, not bit for bit).  Every warp-level primitive carries the env's 16-lane mask.
// ------------------------------------------------------------------------------------------------------------------
template <class Sys, int BLOCK>
__global__ void __launch_bounds__(BLOCK) dopri5_lanes_kernel(const __grid_constant__ qrl_ode_args a) {
    using Env = LaneEnv<Sys, BLOCK / 16>;
    constexpr int M = Env::M, M2 = Env::M2, D = Env::D;
    __shared__ typename Env::Smem smem;
    Env env(a, smem);
    const int K = a.K, l = env.l, el = env.el;
    const bool live = env.live, mode_lane = env.mode_lane;
    double ym[M2];
#pragma unroll
    for (int c = 0; c < 2 * M; ++c) ym[c] = a.y0[(int64_t)el * D + c];
    double v = a.y0[env.col_v];
    auto pick_mode = [&](const double* m) { return env.pick_mode(m); };
    auto emit = [&](int row, double vv, double mv) { env.emit(row, vv, mv); };
    auto rhs_lane = [&](double t, const double* ymt, double vt, double* dm) { return env.rhs(t, ymt, vt, dm); };

    if (live) {
        emit(0, v, pick_mode(ym));
        const double t_end = a.t0 + (double)K * a.t_ssz;
        double t = a.t0;
        double h = (a.h_state && a.h_state[el] > 0.0) ? a.h_state[el] : a.t_ssz;
        int next_row = 1, n_acc = 0, n_rej = 0, guard = 0;
        bool rejected = false;
        double kv[7], km[7][M2];        // stage derivatives of my element and of the modes
        kv[0] = rhs_lane(t, ym, v, km[0]);
        while (next_row <= K && guard++ < dp::MAX_STEPS) {
            const double remaining = t_end - t;
            const bool last = h >= remaining * (1.0 - 1e-12);
            const double ht = last ? remaining : h;
            double ymt[M2], vt;
            // the stage arguments, element by element exactly as in dopri5_kernel
#define QRL_DP_ARG(EXPR_V, EXPR_M)                                       \
            vt = EXPR_V;                                                 \
            _Pragma("unroll") for (int c = 0; c < 2 * M; ++c) ymt[c] = EXPR_M;
            QRL_DP_ARG(fma(ht * dp::A21, kv[0], v), fma(ht * dp::A21, km[0][c], ym[c]))
            kv[1] = rhs_lane(t + dp::C2 * ht, ymt, vt, km[1]);
            QRL_DP_ARG(fma(ht, fma(dp::A32, kv[1], dp::A31 * kv[0]), v),
                       fma(ht, fma(dp::A32, km[1][c], dp::A31 * km[0][c]), ym[c]))
            kv[2] = rhs_lane(t + dp::C3 * ht, ymt, vt, km[2]);
            QRL_DP_ARG(fma(ht, fma(dp::A43, kv[2], fma(dp::A42, kv[1], dp::A41 * kv[0])), v),
                       fma(ht, fma(dp::A43, km[2][c], fma(dp::A42, km[1][c], dp::A41 * km[0][c])), ym[c]))
            kv[3] = rhs_lane(t + dp::C4 * ht, ymt, vt, km[3]);
            QRL_DP_ARG(fma(ht, fma(dp::A54, kv[3], fma(dp::A53, kv[2], fma(dp::A52, kv[1], dp::A51 * kv[0]))), v),
                       fma(ht, fma(dp::A54, km[3][c], fma(dp::A53, km[2][c], fma(dp::A52, km[1][c], dp::A51 * km[0][c]))), ym[c]))
            kv[4] = rhs_lane(t + dp::C5 * ht, ymt, vt, km[4]);
            QRL_DP_ARG(fma(ht, fma(dp::A65, kv[4], fma(dp::A64, kv[3], fma(dp::A63, kv[2], fma(dp::A62, kv[1], dp::A61 * kv[0])))), v),
                       fma(ht, fma(dp::A65, km[4][c], fma(dp::A64, km[3][c], fma(dp::A63, km[2][c],
                                   fma(dp::A62, km[1][c], dp::A61 * km[0][c])))), ym[c]))
            kv[5] = rhs_lane(t + ht, ymt, vt, km[5]);
            // 5th-order solution (= stage-7 argument, FSAL)
            QRL_DP_ARG(fma(ht, fma(dp::B6, kv[5], fma(dp::B5, kv[4], fma(dp::B4, kv[3], fma(dp::B3, kv[2], dp::B1 * kv[0])))), v),
                       fma(ht, fma(dp::B6, km[5][c], fma(dp::B5, km[4][c], fma(dp::B4, km[3][c],
                                   fma(dp::B3, km[2][c], dp::B1 * km[0][c])))), ym[c]))
#undef QRL_DP_ARG
            kv[6] = rhs_lane(t + ht, ymt, vt, km[6]);
            // error estimate: my element (+ my mode amplitude), then the sum over the env's lanes
            auto err_term = [&](double k1, double k3, double k4, double k5, double k6, double k7, double yo, double yn) {
                const double ev = ht * fma(dp::E7, k7, fma(dp::E6, k6, fma(dp::E5, k5, fma(dp::E4, k4, fma(dp::E3, k3, dp::E1 * k1)))));
                const double sc = fma(a.rtol, fmax(fabs(yo), fabs(yn)), a.atol);
                const double r = ev / sc;
                return r * r;
            };
            double sq = err_term(kv[0], kv[2], kv[3], kv[4], kv[5], kv[6], v, vt);
            if (mode_lane) {
                double k1 = 0, k3 = 0, k4 = 0, k5 = 0, k6 = 0, k7 = 0;
#pragma unroll
                for (int c = 0; c < 2 * M; ++c)
                    if (l == c) { k1 = km[0][c]; k3 = km[2][c]; k4 = km[3][c]; k5 = km[4][c]; k6 = km[5][c]; k7 = km[6][c]; }
                sq += err_term(k1, k3, k4, k5, k6, k7, pick_mode(ym), pick_mode(ymt));
            }
            sq = env.env_sum(sq);
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
                        emit(next_row, vt, pick_mode(ymt));
                    } else {
                        const double x = (tg - t) / ht;
                        double pw[7];
#pragma unroll
                        for (int st = 0; st < 7; ++st)
                            pw[st] = fma(fma(fma(dp::P[st][3], x, dp::P[st][2]), x, dp::P[st][1]), x, dp::P[st][0]);
                        double q = 0.0;
#pragma unroll
                        for (int st = 0; st < 7; ++st) {
                            if (st == 1) continue;
                            q = fma(kv[st], pw[st], q);
                        }
                        const double vo = fma(ht * x, q, v);
                        double mo = 0.0;
                        if (mode_lane) {
                            double qm = 0.0;
#pragma unroll
                            for (int st = 0; st < 7; ++st) {
                                if (st == 1) continue;
                                qm = fma(pick_mode(km[st]), pw[st], qm);
                            }
                            mo = fma(ht * x, qm, pick_mode(ym));
                        }
                        emit(next_row, vo, mo);
                    }
                    ++next_row;
                }
                t = t_new;
                v = vt;
                kv[0] = kv[6];
#pragma unroll
                for (int c = 0; c < 2 * M; ++c) { ym[c] = ymt[c]; km[0][c] = km[6][c]; }
                // a step that was clipped to the interval end does not shrink the carried step size
                h = (last && fac >= 1.0) ? fmax(h, ht * fac) : ht * fac;
            } else {
                h = ht * fmax(dp::MIN_FACTOR, dp::SAFETY * pow(err, -0.2));
                rejected = true;
                ++n_rej;
            }
        }
        env.finish(v, ym, h, true, n_acc, n_rej, next_row <= K /* step budget exhausted */);
    }
    if (a.obs_minmax) block_minmax_commit(env.lo, env.hi, a.obs_minmax);
}

}  // namespace qrl
#include "ode_erk.cuh"
namespace qrl {

template <class Sys>
int launch_ode(const qrl_ode_args& a, cudaStream_t st) {
    constexpr int D = Dim<Sys>::D;
    // rows move as 16-byte pieces (LDG/STG.128): torch allocations are 256-byte aligned, slices at odd offsets are not
    const uintptr_t align = (uintptr_t)a.y0 | (uintptr_t)a.states | (uintptr_t)a.observations | (uintptr_t)a.y_last |
                            (uintptr_t)a.obs_last | (uintptr_t)a.obs_noise;
    if (align & 15)
        return fail(QRL_E_INVALID, "%s", "qrl_ode_step: y0, states, observations, y_last, obs_last and obs_noise must be 16-byte aligned");
    if (a.method == QRL_ODE_RK4) {
        if constexpr (Sys::Q == 4) {
            // small batches, 16 lanes per env (B200, 100 sub-steps, tools/quick_ode_bench.py):
            //   E      split (3 roles)   lanes (1 role)   thread per env
            //   1024   54 us             74 us            195 us
            //   2048   107 us            98 us            164 us        (two split CTAs per SM: shared-memory bound)
            //   4096   206 us            171 us           165 us
            // FORCE_LANES picks the split kernel, FORCE_LANES | QRL_ODE_LANES_SINGLE the single-role lanes kernel
            const bool force_lanes = (a.flags & QRL_ODE_FORCE_LANES) != 0;
            if ((a.n_envs <= 3072 || force_lanes) && !(a.flags & QRL_ODE_FORCE_THREAD)) {
                if ((a.flags & QRL_ODE_LANES_SINGLE) || (a.n_envs > 1536 && !force_lanes)) {   // functor inside the lanes' chain
                    constexpr int BL = 64;
                    rk4_lanes_kernel<Sys, BL><<<(a.n_envs + BL / 16 - 1) / (BL / 16), BL, 0, st>>>(a);
                    return after_launch("rk4_lanes_kernel");
                }
                constexpr int EPB = 8;
                const int NS = a.K * a.substeps;
                static const char* c_env = getenv("QRL_ODE_SPLIT_C");   // tuning knob: sub-steps per chunk
                const int c_want = c_env ? atoi(c_env) : 8;
                const int c_max = c_want >= 1 && c_want <= 16 ? c_want : 8;
                const int C = NS < c_max ? (NS > 0 ? NS : 1) : c_max;
                const size_t smem = (size_t)2 * C * 4 * 8 * (EPB + 1) * sizeof(double2) + (size_t)3 * C * EPB * D * sizeof(double);
                static OncePerDevice attr;
                if (attr.first())
                    QRL_CUDA(cudaFuncSetAttribute(rk4_split_kernel<Sys, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                rk4_split_kernel<Sys, EPB><<<(a.n_envs + EPB - 1) / EPB, EPB * 16 + 32 + 128, smem, st>>>(a, C);
                return after_launch("rk4_split_kernel");
            }
        }
        if constexpr (Sys::Q > 4 && Sys::Q % 2 == 0) {
            // q = 6, 8: q lanes per environment.  Measured on B200 (tools/prof_ode.py, kerr_chain, 100 sub-steps, sub-steps/s,
            // row lanes vs one thread per env): q = 8 (the thread kernel spills 8-22 KB): E = 1024 3.5e8 vs 6.6e7, 4096 9.8e8 vs
            // 2.5e8, 16384 1.12e9 vs 7.0e8, 65536 1.17e9 vs 8.8e8; q = 6 (no spills): E = 16384 1.9e9 vs 3.2e9 — row lanes
            // for every batch at q = 8, for small batches at q = 6
            const bool rows = Sys::Q == 8 || a.n_envs <= 6144 || (a.flags & QRL_ODE_FORCE_LANES);
            if (rows && !(a.flags & QRL_ODE_FORCE_THREAD)) {
                constexpr int BR = 128, EPB = (BR / 32) * (32 / Sys::Q);
                rk4_rows_kernel<Sys, BR><<<(a.n_envs + EPB - 1) / EPB, BR, 0, st>>>(a);
                return after_launch("rk4_rows_kernel");
            }
        }
        constexpr int BLOCK = 32;
        const size_t smem = (size_t)BLOCK * ((D / 2) | 1) * sizeof(double2);
        const int blocks = (a.n_envs + BLOCK - 1) / BLOCK;
        // 204 registers at q = 4 = 2 warps per scheduler; capping at 168 (3 warps, 4 bytes of spill) measured slower at
        // every batch size (0.67 vs 0.58 ms at E = 65536)
        // (fused reward: evaluated in the kernel, out of line, +0.18 ms per interval at E = 65536 with the log-negativity on
        // all 101 rows; a separate launch over the block just written measured the same, 0.708 vs 0.703 ms)
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
    if constexpr (Sys::Q == 4) {
        // q = 4: 16 lanes per env, stage derivatives in registers.  B200, 100 grid rows, default tolerances (8.5 steps per
        // interval), lanes vs thread per env: E = 64: 89 vs 477 us, 1024: 98 vs 534 us, 8192: 377 vs 583 us,
        // 16384: 675 vs 629 us, 65536: 2.57 vs 2.12 ms (tools/quick_ode_bench.py)
        const bool lanes = (a.n_envs <= 12288 || (a.flags & QRL_ODE_FORCE_LANES)) && !(a.flags & QRL_ODE_FORCE_THREAD);
        if (lanes) {
            constexpr int BL = 64;
            dopri5_lanes_kernel<Sys, BL><<<(a.n_envs + BL / 16 - 1) / (BL / 16), BL, 0, st>>>(a);
            return after_launch("dopri5_lanes_kernel");
        }
    }
    if constexpr (Sys::Q > 4 && Sys::Q % 2 == 0) {
        // q = 6, 8: q lanes per environment (dopri5_rows_kernel), every batch size.  Measured on B200 (tools/prof_ode.py,
        // kerr_chain, 100 grid rows, default tolerances = 5.4 steps per interval; ms per interval, row lanes vs one thread per
        // env, which keeps 7 d stage derivatives per env in shared memory = 1 (q = 8) / 3 (q = 6) warps per SM):
        //   q = 8: E = 1024 0.21 vs 3.46, 4096 0.37 vs 3.77, 16384 1.38 vs 12.97 (1.19e9 sub-steps/s)
        //   q = 6: E = 1024 0.18 vs 2.30, 4096 0.24 vs 2.51, 16384 0.68 vs 4.46 (2.41e9 sub-steps/s)
        if (!(a.flags & QRL_ODE_FORCE_THREAD)) {
            constexpr int BR = 128, EPB = (BR / 32) * (32 / Sys::Q);
            const size_t smem_rows = (size_t)7 * (Sys::Q + 1) * BR * sizeof(double);
            static OncePerDevice attr_rows;
            if (attr_rows.first())
                QRL_CUDA(cudaFuncSetAttribute(dopri5_rows_kernel<Sys, BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
            dopri5_rows_kernel<Sys, BR><<<(a.n_envs + EPB - 1) / EPB, BR, smem_rows, st>>>(a);
            return after_launch("dopri5_rows_kernel");
        }
    }
    constexpr int BLOCK = 32;
    const size_t smem = (size_t)7 * D * BLOCK * sizeof(double);
    static OncePerDevice attr;
    if (smem > 48 * 1024 && attr.first())
        QRL_CUDA(cudaFuncSetAttribute(dopri5_kernel<Sys, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (a.n_envs + BLOCK - 1) / BLOCK;
    dopri5_kernel<Sys, BLOCK><<<blocks, BLOCK, smem, st>>>(a);
    return after_launch("dopri5_kernel");
}


// one translation unit per system family keeps the build parallel (the q=8 kernels dominate compile time)
#define QRL_ODE_INSTANTIATE(NAME, ...) \
    int NAME(const qrl_ode_args& a, cudaStream_t st) { return launch_ode<__VA_ARGS__>(a, st); }

}  // namespace qrl
