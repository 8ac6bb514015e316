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
north_star.  ``rk4_interval`` is the classic tableau applied to the restated RHS; ``dopri5_dense_interval`` is
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


# Dormand–Prince 5(4) tableau + Shampine's 4th-order dense output (the interpolant behind SciPy's ``RK45``)
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
_E = np.array([71 / 57600, 0.0, -71 / 16695, 71 / 1920, -17253 / 339200, 22 / 525, -1 / 40])
_P = np.array([
    [1.0, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432],
    [0.0, 0.0, 0.0, 0.0],
    [0.0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
    [0.0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
    [0.0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
    [0.0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
    [0.0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423]])

DOPRI5_SAFETY = 0.9
DOPRI5_MIN_FACTOR = 0.2
DOPRI5_MAX_FACTOR = 10.0
DOPRI5_MAX_STEPS = 1000000


def dopri5_dense_interval(func_single, t0, t_ssz, K, y0, args_per_env, atol, rtol, h0=None):
    """Per-env adaptive DP5(4) over one action interval with dense output on the grid ``t0 + i * t_ssz``.

    Each env runs its own controller; steps are free-running (not clipped to grid points), grid rows inside an
    accepted step come from the 4th-order interpolant, the interval end ``t0 + K * t_ssz`` is hit exactly (the action
    changes there) and the step size is carried to the next interval through ``h0``/``h_last``.

    func_single(t, y (d,), args, e) -> (d,).  Returns Y (K+1,E,d), n_accepted (E,), n_rejected (E,), h_last (E,).

    Controller (= the CUDA kernel; SciPy ``RK45``'s rule):
      err = sqrt(mean(((y5 - y4) / (atol + rtol * max(|y|, |y5|)))^2));  accept iff err <= 1
      accepted: h <- h * min(10, max(0.2, 0.9 err^(-1/5)))  (err == 0 -> 10; capped at 1 right after a rejection)
      rejected: h <- h * max(0.2, 0.9 err^(-1/5))
      a step clipped to the interval end never shrinks the carried step size.
    """
    E, d = y0.shape
    Y = np.empty((K + 1, E, d))
    Y[0] = y0
    n_acc = np.zeros(E, dtype=np.int64)
    n_rej = np.zeros(E, dtype=np.int64)
    h_last = np.zeros(E)
    t_end = t0 + K * t_ssz
    for e in range(E):
        y = y0[e].copy()
        args = args_per_env(e)
        t = t0
        h = float(h0[e]) if (h0 is not None and h0[e] > 0.0) else t_ssz
        k = [None] * 7
        k[0] = func_single(t, y, args, e)
        rejected = False
        next_row = 1
        steps = 0
        while next_row <= K:
            steps += 1
            assert steps < DOPRI5_MAX_STEPS
            remaining = t_end - t
            last = h >= remaining * (1.0 - 1e-12)
            ht = remaining if last else h
            for s in range(1, 7):
                ys = y + ht * sum(a * k[j] for j, a in enumerate(_A[s]))
                if s < 6:
                    k[s] = func_single(t + _C[s] * ht, ys, args, e)
            y5 = ys
            k[6] = func_single(t + ht, y5, args, e)
            errv = ht * sum(_E[s] * k[s] for s in range(7))
            sc = atol + rtol * np.maximum(np.abs(y), np.abs(y5))
            err = np.sqrt(np.mean((errv / sc) ** 2))
            if err <= 1.0:
                fac = DOPRI5_MAX_FACTOR if err == 0.0 else min(DOPRI5_MAX_FACTOR,
                                                              max(DOPRI5_MIN_FACTOR, DOPRI5_SAFETY * err ** (-0.2)))
                if rejected:
                    fac = min(1.0, fac)
                rejected = False
                n_acc[e] += 1
                t_new = t_end if last else t + ht
                while next_row <= K:
                    tg = t_end if next_row == K else t0 + next_row * t_ssz
                    if tg > t_new:
                        break
                    if tg == t_new:
                        Y[next_row, e] = y5
                    else:
                        x = (tg - t) / ht
                        pw = np.array([1.0, x, x * x, x * x * x])
                        Y[next_row, e] = y + ht * x * sum(k[s] * (_P[s] @ pw) for s in range(7))
                    next_row += 1
                t = t_new
                y = y5
                k[0] = k[6]
                h = max(h, ht * fac) if (last and fac >= 1.0) else ht * fac
            else:
                h = ht * max(DOPRI5_MIN_FACTOR, DOPRI5_SAFETY * err ** (-0.2))
                rejected = True
                n_rej[e] += 1
        h_last[e] = h
    return Y, n_acc, n_rej, h_last
