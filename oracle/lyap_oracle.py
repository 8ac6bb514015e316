"""TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy/SciPy) of the reference's deterministic hot path.

Never imported by the product package.  Restates, per reference file:line:

* RHS ``func`` ............ quantrl/envs/deterministic.py:653-746 (vec) / 225-316 (single):
  ``y = [Re a (m) | Im a (m) | vec(V) (q*q)]``; ``d a = get_mode_rates``; ``dV = A V + V A^T + D``
* complex<->real packing .. quantrl/envs/deterministic.py:820-867
* how ``T_step`` is stepped  quantrl/solvers/numpy.py:87-178: the *whole batch flattened to one vector* is
  integrated from ``T_step[0]`` to *every* ``T_step[i]`` with SciPy's ``ode`` (vode/dopri5/dop853/lsoda) or
  ``solve_ivp(t_eval=T_step)``; ``args = [params, func_controls, func_delay]`` (:108)
* interval bookkeeping .... quantrl/envs/base.py:440-483 (shared with em_oracle)

Third-party arithmetic: the adaptive integrators are SciPy's (``scipy`` unpinned in the reference's
pyproject.toml:26; 1.18.1 in this image).  They are *called*, not restated.

Fixed-step RK4 and the per-env dopri5 do not exist in the reference; they are defined by the
north_star.  ``rk4_interval`` is the classic tableau applied to the restated RHS; ``dopri5_interval`` is
Dormand–Prince 5(4) (Hairer, Nørsett & Wanner, *Solving ODEs I*, II.5; same tableau/controller family as SciPy's
``RK45``/``dopri5``) with ONE controller PER ENV.  Both are pinned two ways in ``tests/test_oracle_golden.py``:
(1) RK4 driven by the *reference's own* ``env.func`` (fixture G2) must equal ``rk4_interval`` to ~1e-13, and
(2) tight-tolerance reference SciPy trajectories (fixture G1) bound the truncation error of both.
"""
import numpy as np
import scipy.integrate as si


def make_func(num_modes, num_quads, mode_rates, get_A, get_D, E):
    """Return ``func(t, y, args)`` with the reference's semantics; hooks take (t, modes, args)."""
    m, q = num_modes, num_quads

    def func(t, y, args):
        out = np.empty((E, 2 * m + q * q))
        modes = None
        if m:
            modes = y[:, :m] + 1.0j * y[:, m:2 * m]
            r = mode_rates(t, modes, args)
            out[:, :m] = r.real
            out[:, m:2 * m] = r.imag
        if q:
            V = y[:, 2 * m:].reshape(E, q, q)
            A = get_A(t, modes, args)
            D = get_D(t, modes, args)
            dV = (np.matmul(A, V) + np.matmul(V, np.transpose(A, (0, 2, 1)))) + D
            out[:, 2 * m:] = dV.reshape(E, q * q)
        return out

    return func


def rk4_interval(func, T_step, y0, args, substeps=1):
    """Classic RK4 from T_step[i] to T_step[i+1] with ``substeps`` equal sub-divisions.  Returns (len(T_step),E,d)."""
    Y = np.empty((len(T_step),) + y0.shape)
    Y[0] = y0
    y = y0.copy()
    for i in range(len(T_step) - 1):
        h = (T_step[i + 1] - T_step[i]) / substeps
        t = T_step[i]
        for s in range(substeps):
            k1 = func(t, y, args)
            k2 = func(t + 0.5 * h, y + (0.5 * h) * k1, args)
            k3 = func(t + 0.5 * h, y + (0.5 * h) * k2, args)
            k4 = func(t + h, y + h * k3, args)
            y = y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
            t = T_step[i] + (s + 1) * h
        Y[i + 1] = y
    return Y


def scipy_interval(func, T_step, y0, args, method="vode", atol=1e-9, rtol=1e-6):
    """solvers/numpy.py:87-178 — flattened batch, one shared step controller."""
    shape = y0.shape
    flat = lambda t, y, a: func(t, y.reshape(shape), a).ravel()  # noqa: E731
    Y = np.empty((len(T_step), y0.size))
    Y[0] = y0.ravel()
    if method in ("dop853", "dopri5", "lsoda", "vode", "zvode"):
        integ = si.ode(flat)
        integ.set_integrator(name=method, atol=atol, rtol=rtol, method="adams")
        integ.set_initial_value(y=y0.ravel(), t=T_step[0])
        integ.set_f_params(args)
        for i in range(1, len(T_step)):
            Y[i] = integ.integrate(T_step[i])
    else:
        Y[:] = si.solve_ivp(fun=flat, y0=y0.ravel(), t_span=[T_step[0], T_step[-1]], t_eval=T_step, method=method,
                            atol=atol, rtol=rtol, args=(args,)).y.T
    return Y.reshape((len(T_step),) + shape)


# Dormand–Prince 5(4) tableau
_C = np.array([0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0])
_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_B5 = np.array([35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0])
_E = np.array([71 / 57600, 0.0, -71 / 16695, 71 / 1920, -17253 / 339200, 22 / 525, -1 / 40])

DOPRI5_SAFETY = 0.9
DOPRI5_MIN_FACTOR = 0.2
DOPRI5_MAX_FACTOR = 10.0
DOPRI5_MAX_STEPS = 100000


def dopri5_interval(func_single, T_step, y0, args_per_env, atol, rtol, h0=None):
    """Per-env adaptive DP5(4): each env integrates from T_step[i] to T_step[i+1] *landing on the grid*
    (the step is clipped to the grid point, as SciPy's old ``ode`` interface does when asked to integrate to
    ``T_step[i]``), carrying its own step size across grid points.

    func_single(t, y (d,), args, e) -> (d,) evaluates env ``e`` alone.  Returns Y (len(T_step),E,d),
    n_accepted (E,), n_rejected (E,), h_last (E,).

    Controller (what the CUDA kernel implements; Hairer's DOPRI5 without PI stabilisation / SciPy ``RK45``):
      err = sqrt(mean(((y5 - y4) / (atol + rtol * max(|y|, |y5|)))^2))
      accept iff err <= 1;  h_new = h * clip(0.9 * err^(-1/5), 0.2, 10)   (err == 0 -> factor 10)
      after a rejection the growth factor is capped at 1 for the next step.
    """
    E, d = y0.shape
    Y = np.empty((len(T_step), E, d))
    Y[0] = y0
    n_acc = np.zeros(E, dtype=np.int64)
    n_rej = np.zeros(E, dtype=np.int64)
    h_last = np.zeros(E)
    for e in range(E):
        y = y0[e].copy()
        args = args_per_env(e)
        h = (T_step[1] - T_step[0]) if h0 is None else float(h0[e])
        k = [None] * 7
        k[0] = func_single(T_step[0], y, args, e)
        rejected = False
        for i in range(len(T_step) - 1):
            t, t_end = T_step[i], T_step[i + 1]
            steps = 0
            while t < t_end:
                steps += 1
                assert steps < DOPRI5_MAX_STEPS
                h_try = h
                last = False
                if t + h_try >= t_end or (t_end - (t + h_try)) < 1e-14 * abs(t_end):
                    h_try = t_end - t
                    last = True
                for s in range(1, 7):
                    ys = y.copy()
                    for j, a in enumerate(_A[s]):
                        if a != 0.0:
                            ys = ys + (h_try * a) * k[j]
                    k[s] = func_single(t + _C[s] * h_try, ys, args, e)
                y5 = ys  # stage 7 argument is the 5th-order solution (FSAL)
                errv = h_try * sum(_E[s] * k[s] for s in range(7))
                sc = atol + rtol * np.maximum(np.abs(y), np.abs(y5))
                err = np.sqrt(np.mean((errv / sc) ** 2))
                if err <= 1.0:
                    fac = DOPRI5_MAX_FACTOR if err == 0.0 else min(DOPRI5_MAX_FACTOR,
                                                                  max(DOPRI5_MIN_FACTOR,
                                                                      DOPRI5_SAFETY * err ** (-0.2)))
                    if rejected:
                        fac = min(1.0, fac)
                    rejected = False
                    n_acc[e] += 1
                    t = t_end if last else t + h_try
                    y = y5
                    k[0] = k[6]
                    if not last or fac < 1.0:
                        h = h_try * fac
                    else:
                        h = max(h, h_try * fac) if h_try * fac > h else h
                else:
                    fac = max(DOPRI5_MIN_FACTOR, DOPRI5_SAFETY * err ** (-0.2))
                    h = h_try * fac
                    rejected = True
                    n_rej[e] += 1
            Y[i + 1, e] = y
        h_last[e] = h
    return Y, n_acc, n_rej, h_last
