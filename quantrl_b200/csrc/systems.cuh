// Registry of deterministic system functors (device code) — the shipped "user hooks" of the Lyapunov path.
//
// Each functor replaces the three Python hooks the reference calls from its RHS (quantrl/envs/deterministic.py:698-709,
// 849-853): get_mode_rates(t, modes, args), get_A(t, modes, args), get_D(t, modes, args).  Formulas and parameter-row
// layouts are those of oracle/systems.py (NumPy twin).  A functor provides
//     static constexpr int M, Q, NPARAMS
//     void load(const double* params_row, const double* actions_row)
//     void eval(double t, const double* y, double* dmodes /*2M*/, double (&A)[Q][Q]) const     // mode rates + drift
//     double D(int i, int j) const                                                              // diffusion matrix
//     static constexpr bool nz(int i, int k) / d_nz(int i, int j)   // false = A_ik / D_ij is a structural zero: the
//                                                                   // right-hand side skips that product / addition
// y = [Re a (M) | Im a (M) | vec(V) (Q*Q)]  (deterministic.py:667).
#pragma once
#include <cuda_runtime.h>

namespace qrl {

// optomech: m=2 (a optical, b mechanical), q=4.  p = [kappa, gamma, omega_m, Delta0, g0, n_th, A_l]
struct SysOptomech {
    static constexpr int M = 2, Q = 4, NPARAMS = 7;
    double hk, hg, om, D0, g0, dth, drive;  // kappa/2, gamma/2, omega_m, Delta0, g0, gamma(n_th+1/2), A_l(1+u)
    __device__ __forceinline__ void load(const double* p, const double* u) {
        hk = 0.5 * p[0]; hg = 0.5 * p[1]; om = p[2]; D0 = p[3]; g0 = p[4];
        dth = p[1] * (p[5] + 0.5);
        drive = p[6] * (1.0 + u[0]);
    }
    __device__ __forceinline__ void eval(double, const double* y, double* dm, double (&A)[4][4]) const {
        const double ar = y[0], br = y[1], ai = y[2], bi = y[3];
        // da = -(kappa/2 - i Delta0) a + i g0 a (b + b*) + A_l (1 + u);  (b + b*) = 2 Re b
        const double c = g0 * (2.0 * br);
        dm[0] = -hk * ar - D0 * ai - c * ai + drive;      // Re da
        dm[2] = -hk * ai + D0 * ar + c * ar;              // Im da
        // db = -(gamma/2 + i omega_m) b + i g0 |a|^2
        dm[1] = -hg * br + om * bi;                       // Re db
        dm[3] = -hg * bi - om * br + g0 * (ar * ar + ai * ai);   // Im db
        const double Delta = D0 + 2.0 * g0 * br, GR = 2.0 * g0 * ar, GI = 2.0 * g0 * ai;
        A[0][0] = -hk;  A[0][1] = -Delta; A[0][2] = -GI;  A[0][3] = 0.0;
        A[1][0] = Delta; A[1][1] = -hk;   A[1][2] = GR;   A[1][3] = 0.0;
        A[2][0] = 0.0;  A[2][1] = 0.0;    A[2][2] = -hg;  A[2][3] = om;
        A[3][0] = GR;   A[3][1] = GI;     A[3][2] = -om;  A[3][3] = -hg;
    }
    __device__ __forceinline__ double D(int i, int j) const { return i != j ? 0.0 : (i < 2 ? hk : dth); }
    __host__ __device__ static constexpr bool nz(int i, int k) {
        return !((i < 2 && k == 3) || (i == 2 && k < 2));
    }
    __host__ __device__ static constexpr bool d_nz(int i, int j) { return i == j; }
};

// kerr_chain: m=N, q=2N.  p = [kappa, chi, J, n_th, eps, Delta_0..Delta_{N-1}]
template <int N>
struct SysKerrChain {
    static constexpr int M = N, Q = 2 * N, NPARAMS = 5 + N;
    double hk, chi, J, dd, drive, Dl[N];
    __device__ __forceinline__ void load(const double* p, const double* u) {
        hk = 0.5 * p[0]; chi = p[1]; J = p[2]; dd = p[0] * (p[3] + 0.5);
        drive = p[4] * (1.0 + u[0]);
#pragma unroll
        for (int j = 0; j < N; ++j) Dl[j] = p[5 + j];
    }
    __device__ __forceinline__ void eval(double, const double* y, double* dm, double (&A)[2 * N][2 * N]) const {
#pragma unroll
        for (int i = 0; i < 2 * N; ++i)
#pragma unroll
            for (int j = 0; j < 2 * N; ++j) A[i][j] = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double ar = y[j], ai = y[N + j];
            const double n2 = ar * ar + ai * ai;
            double nr = 0.0, ni = 0.0;
            if (j > 0) { nr += y[j - 1]; ni += y[N + j - 1]; }
            if (j < N - 1) { nr += y[j + 1]; ni += y[N + j + 1]; }
            // da_j = -(kappa/2 + i Delta_j) a_j - i chi |a_j|^2 a_j - i J (a_{j-1} + a_{j+1}) + [j==0] eps (1+u)
            const double w = Dl[j] + chi * n2;
            dm[j] = -hk * ar + w * ai + J * ni + (j == 0 ? drive : 0.0);
            dm[N + j] = -hk * ai - w * ar - J * nr;
            const double Dp = Dl[j] + 2.0 * chi * n2;
            const double GR = chi * (ar * ar - ai * ai), GI = chi * (2.0 * ar * ai);
            A[2 * j][2 * j] = -hk + GI;      A[2 * j][2 * j + 1] = Dp - GR;
            A[2 * j + 1][2 * j] = -Dp - GR;  A[2 * j + 1][2 * j + 1] = -hk - GI;
            if (j > 0) { A[2 * j][2 * (j - 1) + 1] = J; A[2 * j + 1][2 * (j - 1)] = -J; }
            if (j < N - 1) { A[2 * j][2 * (j + 1) + 1] = J; A[2 * j + 1][2 * (j + 1)] = -J; }
        }
    }
    __device__ __forceinline__ double D(int i, int j) const { return i == j ? dd : 0.0; }
    // 2x2 blocks on the diagonal, and the two hopping entries (different parity) of the neighbouring blocks
    __host__ __device__ static constexpr bool nz(int i, int k) {
        return (i / 2 == k / 2) || (((i / 2 - k / 2) == 1 || (k / 2 - i / 2) == 1) && (i % 2 != k % 2));
    }
    __host__ __device__ static constexpr bool d_nz(int i, int j) { return i == j; }
};

// const_AD: no modes; A, D given per env and constant over the interval (is_A_constant / is_D_constant,
// deterministic.py:698,705)
template <int Q_>
struct SysConstAD {
    static constexpr int M = 0, Q = Q_, NPARAMS = 0;
    double Ac[Q_][Q_], Dc[Q_][Q_];
    __device__ __forceinline__ void load_AD(const double* A, const double* D) {
#pragma unroll
        for (int i = 0; i < Q_; ++i)
#pragma unroll
            for (int j = 0; j < Q_; ++j) { Ac[i][j] = A[i * Q_ + j]; Dc[i][j] = D[i * Q_ + j]; }
    }
    __device__ __forceinline__ void load(const double*, const double*) {}
    __device__ __forceinline__ void eval(double, const double*, double*, double (&A)[Q_][Q_]) const {
#pragma unroll
        for (int i = 0; i < Q_; ++i)
#pragma unroll
            for (int j = 0; j < Q_; ++j) A[i][j] = Ac[i][j];
    }
    __device__ __forceinline__ double D(int i, int j) const { return Dc[i][j]; }
    __host__ __device__ static constexpr bool nz(int, int) { return true; }
    __host__ __device__ static constexpr bool d_nz(int, int) { return true; }
};

template <class S> struct is_const_ad { static constexpr bool value = false; };
template <int Q_> struct is_const_ad<SysConstAD<Q_>> { static constexpr bool value = true; };

}  // namespace qrl
