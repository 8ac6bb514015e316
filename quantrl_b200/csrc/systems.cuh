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
// The Kerr-chain functor is evaluated in two shapes (full matrix per thread; one row per lane, RowSys below).  Both call
// these helpers, written with explicit roundings (no compiler-chosen fused contractions), so that a row of the drift
// matrix and the mode rates are bit-identical however the surrounding code is compiled.
struct KerrBlock { double e0, e1, o0, o1; };    // A[2j][2j], A[2j][2j+1], A[2j+1][2j], A[2j+1][2j+1]
__device__ __forceinline__ KerrBlock kerr_block(double ar, double ai, double dl, double chi, double hk) {
    const double ar2 = __dmul_rn(ar, ar), ai2 = __dmul_rn(ai, ai), arai = __dmul_rn(ar, ai);
    const double n2 = __dadd_rn(ar2, ai2);
    const double Dp = fma(__dmul_rn(2.0, chi), n2, dl);
    const double GR = __dmul_rn(chi, __dsub_rn(ar2, ai2)), GI = __dmul_rn(chi, __dmul_rn(2.0, arai));
    KerrBlock b;
    b.e0 = __dadd_rn(-hk, GI); b.e1 = __dsub_rn(Dp, GR);
    b.o0 = __dsub_rn(-Dp, GR); b.o1 = __dsub_rn(-hk, GI);
    return b;
}
// d a_j = -(kappa/2 + i Delta_j) a_j - i chi |a_j|^2 a_j - i J (a_{j-1} + a_{j+1}) + [j==0] eps (1+u)
__device__ __forceinline__ void kerr_rate(double ar, double ai, double nr, double ni, double dl, double chi, double hk, double J,
                                          double drive0, double& dre, double& dim) {
    const double n2 = __dadd_rn(__dmul_rn(ar, ar), __dmul_rn(ai, ai));
    const double w = fma(chi, n2, dl);
    dre = __dadd_rn(fma(J, ni, fma(w, ai, __dmul_rn(-hk, ar))), drive0);
    dim = fma(-J, nr, fma(-w, ar, __dmul_rn(-hk, ai)));
}

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
            double nr = 0.0, ni = 0.0;
            if (j > 0) { nr += y[j - 1]; ni += y[N + j - 1]; }
            if (j < N - 1) { nr += y[j + 1]; ni += y[N + j + 1]; }
            kerr_rate(ar, ai, nr, ni, Dl[j], chi, hk, J, j == 0 ? drive : 0.0, dm[j], dm[N + j]);
            const KerrBlock b = kerr_block(ar, ai, Dl[j], chi, hk);
            A[2 * j][2 * j] = b.e0;      A[2 * j][2 * j + 1] = b.e1;
            A[2 * j + 1][2 * j] = b.o0;  A[2 * j + 1][2 * j + 1] = b.o1;
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

// ---- "row" views of the functors for the q-lanes-per-environment kernels (rk4_rows_kernel, ode_impl.cuh): lane i of an
// environment evaluates the mode rates (replicated, they are cheap) and only ROW i of the drift matrix and of the
// diffusion matrix.  Same expressions as eval() above, so the rows are bit-identical to the full matrices.
//     void load(const double* params_row, const double* actions_row, const double* A_e, const double* D_e, int i)
//     void rates(double t, const double* y, double* dm /*2M*/) const
//     void row(double t, const double* y, double (&Arow)[Q]) const
//     double Drow(int j) const
template <class Sys> struct RowSys;

template <int N>
struct RowSys<SysKerrChain<N>> {
    static constexpr int M = N, Q = 2 * N;
    double hk, chi, J, dd, drive, Dl[N];
    int i;
    __device__ __forceinline__ void load(const double* p, const double* u, const double*, const double*, int row) {
        hk = 0.5 * p[0]; chi = p[1]; J = p[2]; dd = p[0] * (p[3] + 0.5);
        drive = p[4] * (1.0 + u[0]);
#pragma unroll
        for (int j = 0; j < N; ++j) Dl[j] = p[5 + j];
        i = row;
    }
    __device__ __forceinline__ void rates(double, const double* y, double* dm) const {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double ar = y[j], ai = y[N + j];
            double nr = 0.0, ni = 0.0;
            if (j > 0) { nr += y[j - 1]; ni += y[N + j - 1]; }
            if (j < N - 1) { nr += y[j + 1]; ni += y[N + j + 1]; }
            kerr_rate(ar, ai, nr, ni, Dl[j], chi, hk, J, j == 0 ? drive : 0.0, dm[j], dm[N + j]);
        }
    }
    __device__ __forceinline__ void row(double, const double* y, double (&Arow)[2 * N]) const {
        const int j = i >> 1;
        const bool odd = i & 1;
        double ar = 0.0, ai = 0.0, dl = 0.0;
#pragma unroll
        for (int jj = 0; jj < N; ++jj)
            if (jj == j) { ar = y[jj]; ai = y[N + jj]; dl = Dl[jj]; }     // predicated selects: no dynamic register indexing
        const KerrBlock b = kerr_block(ar, ai, dl, chi, hk);
        // row 2j:   [2j] = -hk + GI, [2j+1] = Dp - GR, hopping J at columns 2(j-1)+1 and 2(j+1)+1
        // row 2j+1: [2j] = -Dp - GR, [2j+1] = -hk - GI, hopping -J at columns 2(j-1) and 2(j+1)
        const double d0 = odd ? b.o0 : b.e0, d1 = odd ? b.o1 : b.e1, hop = odd ? -J : J;
        const int c_lo = 2 * (j - 1) + (odd ? 0 : 1), c_hi = 2 * (j + 1) + (odd ? 0 : 1);
#pragma unroll
        for (int c = 0; c < 2 * N; ++c)
            Arow[c] = (c == 2 * j) ? d0 : (c == 2 * j + 1) ? d1 : ((j > 0 && c == c_lo) || (j < N - 1 && c == c_hi)) ? hop : 0.0;
    }
    __device__ __forceinline__ double Drow(int j) const { return i == j ? dd : 0.0; }
};

template <int Q_>
struct RowSys<SysConstAD<Q_>> {
    static constexpr int M = 0, Q = Q_;
    double Ar[Q_], Dr[Q_];
    __device__ __forceinline__ void load(const double*, const double*, const double* A, const double* D, int row) {
#pragma unroll
        for (int j = 0; j < Q_; ++j) { Ar[j] = A[row * Q_ + j]; Dr[j] = D[row * Q_ + j]; }
    }
    __device__ __forceinline__ void rates(double, const double*, double*) const {}
    __device__ __forceinline__ void row(double, const double*, double (&Arow)[Q_]) const {
#pragma unroll
        for (int j = 0; j < Q_; ++j) Arow[j] = Ar[j];
    }
    __device__ __forceinline__ double Drow(int j) const {
        double d = 0.0;
#pragma unroll
        for (int c = 0; c < Q_; ++c) d = (c == j) ? Dr[c] : d;
        return d;
    }
};

}  // namespace qrl
