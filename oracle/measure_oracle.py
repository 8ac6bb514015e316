"""TEST INFRASTRUCTURE ONLY — CPU restatement of the quantum-correlation measures used as fused reward functors.

Restates ``QCMSolver.get_entanglement_logarithmic_negativity_2`` (quantrl/solvers/measure.py:401-440) with its helpers
``get_submatrices`` (:184-222) and ``get_invariants`` (:224-254): analytic logarithmic negativity of two modes from
the 2x2 blocks of the quadrature covariance matrix.  Pinned by ``tests/golden/lyap_optomech.npz:E5_entan_ln_ref``,
which ``oracle/make_golden.py`` computes with the reference's own ``QCMSolver``.
"""
import numpy as np


def entan_ln_2(Corrs, pos_i=0, pos_j=2):
    """Corrs (..., q, q) -> (...,) logarithmic negativity of the modes whose quadratures start at pos_i / pos_j."""
    V = np.asarray(Corrs, dtype=np.float64)
    A = V[..., pos_i:pos_i + 2, pos_i:pos_i + 2]
    B = V[..., pos_j:pos_j + 2, pos_j:pos_j + 2]
    Cm = V[..., pos_i:pos_i + 2, pos_j:pos_j + 2]
    # the reference rebuilds the 4x4 matrix from A, B, C and C^T (measure.py:212-216)
    top = np.concatenate((A, Cm), axis=-1)
    bot = np.concatenate((np.swapaxes(Cm, -1, -2), B), axis=-1)
    full = np.concatenate((top, bot), axis=-2)
    I1, I2, I3, I4 = np.linalg.det(A), np.linalg.det(B), np.linalg.det(Cm), np.linalg.det(full)
    sigma = I1 + I2 - 2 * I3
    disc = sigma ** 2 - 4 * I4
    ok = np.logical_and(disc >= 0.0, I4 >= 0.0)
    out = np.zeros(sigma.shape)
    with np.errstate(invalid="ignore", divide="ignore"):
        val = -np.log(2 / np.sqrt(2) * np.sqrt(sigma - np.sqrt(np.where(ok, disc, 0.0))))
    out[ok] = val[ok]
    out[~np.isfinite(out)] = 0.0
    out[out < 0.0] = 0.0
    return out
