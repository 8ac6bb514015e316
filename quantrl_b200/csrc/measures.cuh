// Fused reward functors of the deterministic path: quantum-correlation measures evaluated in the kernel epilogue instead
// of a (K+1, E, q, q) round trip through Python.  Restates QCMSolver.get_entanglement_logarithmic_negativity_2 and its
// helpers (quantrl/solvers/measure.py:184-254, 401-440), get_synchronization_complete (:442-468) and
// get_synchronization_phase (:470-514); CPU twin: oracle/measure_oracle.py, pinned by tests/golden/measures_optomech.npz.
#pragma once
#include <cuda_runtime.h>

namespace qrl {

// Analytic logarithmic negativity of modes (0,1) from a 4x4 quadrature covariance matrix V (row-major, 16 doubles):
//   A = V[0:2,0:2], B = V[2:4,2:4], C = V[0:2,2:4];  sigma = det A + det B - 2 det C;  I4 = det [[A, C], [C^T, B]]
//   E_N = max(0, -ln(sqrt(2) * sqrt(sigma - sqrt(sigma^2 - 4 I4))))   if sigma^2 - 4 I4 >= 0 and I4 >= 0, else 0
__device__ __forceinline__ double log_negativity_2mode(const double* V) {
    const double a00 = V[0], a01 = V[1], a10 = V[4], a11 = V[5];
    const double c00 = V[2], c01 = V[3], c10 = V[6], c11 = V[7];
    const double b00 = V[10], b01 = V[11], b10 = V[14], b11 = V[15];
    const double I1 = a00 * a11 - a01 * a10, I2 = b00 * b11 - b01 * b10, I3 = c00 * c11 - c01 * c10;
    // 4x4 determinant of [[A, C], [C^T, B]] by 2x2 minors of the top and bottom row pairs (the reference rebuilds the
    // matrix from A, B, C and C^T, measure.py:212-216)
    const double m00 = a00, m01 = a01, m02 = c00, m03 = c01;
    const double m10 = a10, m11 = a11, m12 = c10, m13 = c11;
    const double m20 = c00, m21 = c10, m22 = b00, m23 = b01;
    const double m30 = c01, m31 = c11, m32 = b10, m33 = b11;
    const double s0 = m00 * m11 - m10 * m01, s1 = m00 * m12 - m10 * m02, s2 = m00 * m13 - m10 * m03;
    const double s3 = m01 * m12 - m11 * m02, s4 = m01 * m13 - m11 * m03, s5 = m02 * m13 - m12 * m03;
    const double k5 = m22 * m33 - m32 * m23, k4 = m21 * m33 - m31 * m23, k3 = m21 * m32 - m31 * m22;
    const double k2 = m20 * m33 - m30 * m23, k1 = m20 * m32 - m30 * m22, k0 = m20 * m31 - m30 * m21;
    const double I4 = s0 * k5 - s1 * k4 + s2 * k3 + s3 * k2 - s4 * k1 + s5 * k0;
    const double sigma = I1 + I2 - 2.0 * I3;
    const double disc = sigma * sigma - 4.0 * I4;
    if (!(disc >= 0.0) || !(I4 >= 0.0)) return 0.0;
    const double en = -log(1.4142135623730951 * sqrt(sigma - sqrt(disc)));
    return (en > 0.0 && isfinite(en)) ? en : 0.0;
}

// Complete quantum synchronization of modes (0,1), QCMSolver.get_synchronization_complete (measure.py:442-468):
//   S_c = 1 / (<q_-^2> + <p_-^2>),  <q_-^2> = (V00 + V22 - 2 V02) / 2,  <p_-^2> = (V11 + V33 - 2 V13) / 2
__device__ __forceinline__ double sync_complete_2mode(const double* V) {
    const double q2 = 0.5 * (V[0] + V[10] - 2.0 * V[2]);
    const double p2 = 0.5 * (V[5] + V[15] - 2.0 * V[7]);
    return 1.0 / (q2 + p2);
}

// Quantum phase synchronization of modes (0,1), QCMSolver.get_synchronization_phase (measure.py:470-514): the momentum
// quadratures are rotated by the phases of the mode amplitudes a_i = re_i + i im_i (cos/sin of np.angle = re/|a|, im/|a|;
// np.angle(0) = 0), S_p = 1 / (2 <p'_-^2>).
__device__ __forceinline__ double sync_phase_2mode(double re_i, double im_i, double re_j, double im_j, const double* V) {
    const double ri = hypot(re_i, im_i), rj = hypot(re_j, im_j);
    const double ci = ri > 0.0 ? re_i / ri : 1.0, si = ri > 0.0 ? im_i / ri : 0.0;
    const double cj = rj > 0.0 ? re_j / rj : 1.0, sj = rj > 0.0 ? im_j / rj : 0.0;
    const double pi2 = si * si * V[0] - si * ci * V[1] - ci * si * V[4] + ci * ci * V[5];
    const double pj2 = sj * sj * V[10] - sj * cj * V[11] - cj * sj * V[14] + cj * cj * V[15];
    const double pij = si * sj * V[2] - si * cj * V[3] - ci * sj * V[6] + ci * cj * V[7];
    return 0.5 / (0.5 * (pi2 + pj2 - 2.0 * pij));
}

}  // namespace qrl
