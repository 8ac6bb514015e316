"""TEST INFRASTRUCTURE ONLY — CPU restatement of the quantum-correlation measures used as fused reward functors.

Restates ``QCMSolver.get_entanglement_logarithmic_negativity_2`` (quantrl/solvers/measure.py:401-440) with its helpers
``get_submatrices`` (:184-222) and ``get_invariants`` (:224-254): analytic logarithmic negativity of two modes from
the 2x2 blocks of the quadrature covariance matrix; ``get_synchronization_complete`` (:442-468) and
``get_synchronization_phase`` (:470-514).  Pinned by ``tests/golden/lyap_optomech.npz:E5_entan_ln_ref`` and
``tests/golden/measures_optomech.npz`` (smooth and noisy inputs), which ``oracle/make_golden.py`` /
``oracle/make_golden_measures.py`` compute with the reference's own ``QCMSolver``.
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


def sync_complete(Corrs, pos_i=0, pos_j=2):
    """Complete quantum synchronization (measure.py:442-468): ``1 / (<q_-^2> + <p_-^2>)``."""
    V = np.asarray(Corrs, dtype=np.float64)
    q2 = 0.5 * (V[..., pos_i, pos_i] + V[..., pos_j, pos_j] - 2 * V[..., pos_i, pos_j])
    p2 = 0.5 * (V[..., pos_i + 1, pos_i + 1] + V[..., pos_j + 1, pos_j + 1] - 2 * V[..., pos_i + 1, pos_j + 1])
    return 1.0 / (q2 + p2)


def sync_phase(Modes, Corrs, pos_i=0, pos_j=2):
    """Quantum phase synchronization (measure.py:470-514): momentum quadratures rotated by the mode phases."""
    V = np.asarray(Corrs, dtype=np.float64)
    M = np.asarray(Modes)
    ai, aj = np.angle(M[..., pos_i // 2]), np.angle(M[..., pos_j // 2])
    ci, cj, si, sj = np.cos(ai), np.cos(aj), np.sin(ai), np.sin(aj)
    pi2 = si ** 2 * V[..., pos_i, pos_i] - si * ci * V[..., pos_i, pos_i + 1] - ci * si * V[..., pos_i + 1, pos_i] \
        + ci ** 2 * V[..., pos_i + 1, pos_i + 1]
    pj2 = sj ** 2 * V[..., pos_j, pos_j] - sj * cj * V[..., pos_j, pos_j + 1] - cj * sj * V[..., pos_j + 1, pos_j] \
        + cj ** 2 * V[..., pos_j + 1, pos_j + 1]
    pij = si * sj * V[..., pos_i, pos_j] - si * cj * V[..., pos_i, pos_j + 1] - ci * sj * V[..., pos_i + 1, pos_j] \
        + ci * cj * V[..., pos_i + 1, pos_j + 1]
    return 0.5 / (0.5 * (pi2 + pj2 - 2 * pij))


def measures_of_states(States, m=2, q=4):
    """All three measures of a ``(..., 2m + q*q)`` state block ``[Re a | Im a | vec(V)]``."""
    S = np.asarray(States, dtype=np.float64)
    V = S[..., 2 * m:].reshape(S.shape[:-1] + (q, q))
    Modes = S[..., :m] + 1j * S[..., m:2 * m]
    return {"entan_ln": entan_ln_2(V), "sync_c": sync_complete(V), "sync_p": sync_phase(Modes, V)}
