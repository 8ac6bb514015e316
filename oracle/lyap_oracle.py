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


# ---------------------------------------------------------------------------------------------------------------------
# Generic explicit Runge–Kutta engine (CPU twin of csrc/ode_erk.cuh): the explicit-RK method names of the reference's
# solver plugins (torchdiffeq quantrl/solvers/torch.py:28, diffrax quantrl/solvers/jax.py:29, SciPy
# quantrl/solvers/numpy.py:31).  Tableaus are the published ones, written here independently of the CUDA header
# (``tests/test_oracle_golden.py`` checks their order conditions); DOP853's come from SciPy's coefficient module.
# Form: S stages, y_new = y + h sum B_j k_j, k_S = f(t + h, y_new) closes the error estimate and starts the next step.
# ---------------------------------------------------------------------------------------------------------------------
def _tableau(name):
    name = {"RK23": "bosh3", "heun": "adaptive_heun", "DOP853": "dop853", "dopri8": "dop853"}.get(name, name)
    if name == "bosh3":
        A = [[0, 0, 0], [1 / 2, 0, 0], [0, 3 / 4, 0]]
        B = [2 / 9, 1 / 3, 4 / 9]
        Ev = [2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8]
        return dict(A=A, B=B, E=Ev, order=3, err_order=2, adaptive=True)
    if name == "tsit5":
        A = np.zeros((6, 6))
        A[1, 0] = 0.161
        A[2, :2] = [-0.008480655492356989, 0.335480655492357]
        A[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
        A[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
        A[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]
        B = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]
        Ev = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
              0.5823571654525552, -0.45808210592918697, 1 / 66]
        return dict(A=A, B=B, E=Ev, order=5, err_order=4, adaptive=True)
    if name == "adaptive_heun":
        return dict(A=[[0, 0], [1.0, 0]], B=[0.5, 0.5], E=[-0.5, 0.5, 0.0], order=2, err_order=1, adaptive=True)
    if name == "fehlberg2":
        return dict(A=[[0, 0, 0], [1 / 2, 0, 0], [1 / 256, 255 / 256, 0]], B=[1 / 512, 255 / 256, 1 / 512],
                    E=[-1 / 512, 0.0, 1 / 512, 0.0], order=2, err_order=1, adaptive=True)
    if name == "euler":
        return dict(A=[[0]], B=[1.0], E=None, order=1, adaptive=False)
    if name == "midpoint":
        return dict(A=[[0, 0], [0.5, 0]], B=[0.0, 1.0], E=None, order=2, adaptive=False)
    if name == "dop853":
        from scipy.integrate._ivp import dop853_coefficients as dc
        S = dc.N_STAGES
        return dict(A=np.asarray(dc.A[:S, :S]), B=np.asarray(dc.B), E=np.asarray(dc.E5), E3=np.asarray(dc.E3), order=8,
                    err_order=7, adaptive=True)
    raise ValueError(name)


def erk_tableau(name):
    """(A (S,S), B (S,), C (S,), E (S+1,) or None, meta) of an explicit RK method of the generic engine."""
    tb = _tableau(name)
    A = np.asarray(tb["A"], dtype=np.float64)
    tb = dict(tb, A=A, B=np.asarray(tb["B"], dtype=np.float64), C=A.sum(axis=1))
    return tb


def erk_interval(method, func_single, t0, t_ssz, K, y0, args_per_env, atol=1e-9, rtol=1e-6, h0=None, substeps=1):
    """One action interval with the generic explicit RK engine, one controller per env, steps clipped to the grid points
    ``t0 + i * t_ssz`` (rows are step results, no interpolation).  Same controller as ``dopri5_dense_interval`` with the
    exponent ``-1 / (err_order + 1)``; DOP853 combines its two estimators as in SciPy (``|h| e5^2 / sqrt((e5^2 + 0.01
    e3^2) d)``).  Fixed-step methods take ``substeps`` equal steps per grid interval.
    Returns Y (K+1,E,d), n_accepted (E,), n_rejected (E,), h_last (E,)."""
    tb = erk_tableau(method)
    A, B, C, Ev = tb["A"], tb["B"], tb["C"], tb["E"]
    S = len(B)
    E_, d = y0.shape
    Y = np.empty((K + 1, E_, d))
    Y[0] = y0
    n_acc, n_rej, h_last = np.zeros(E_, dtype=np.int64), np.zeros(E_, dtype=np.int64), np.zeros(E_)
    t_end = t0 + K * t_ssz
    expo = -1.0 / (tb["err_order"] + 1) if tb["adaptive"] else 0.0
    for e in range(E_):
        y = y0[e].copy()
        args = args_per_env(e)
        t = t0
        h = (float(h0[e]) if (h0 is not None and h0[e] > 0.0) else t_ssz) if tb["adaptive"] else t_ssz / substeps
        k = [None] * (S + 1)
        k[0] = func_single(t, y, args, e)
        rejected = False
        for row in range(1, K + 1):
            t_row = t_end if row == K else t0 + row * t_ssz
            sub = 0
            while True:
                remaining = t_row - t
                if tb["adaptive"]:
                    last = h >= remaining * (1.0 - 1e-12)
                    ht = remaining if last else h
                else:
                    sub += 1
                    last = sub >= substeps
                    ht = remaining if last else t_ssz / substeps
                for s in range(1, S):
                    ys = y + ht * sum(A[s, j] * k[j] for j in range(s) if A[s, j] != 0.0)
                    k[s] = func_single(t + C[s] * ht, ys, args, e)
                y_new = y + ht * sum(B[j] * k[j] for j in range(S) if B[j] != 0.0)
                t_new = t_row if last else t + ht
                k[S] = func_single(t_new, y_new, args, e)
                accept, fac = True, 1.0
                if tb["adaptive"]:
                    sc = atol + rtol * np.maximum(np.abs(y), np.abs(y_new))
                    if method in ("dop853", "DOP853", "dopri8"):
                        e5 = sum(Ev[j] * k[j] for j in range(S + 1)) / sc
                        e3 = sum(tb["E3"][j] * k[j] for j in range(S + 1)) / sc
                        s5, s3 = float(e5 @ e5), float(e3 @ e3)
                        den = s5 + 0.01 * s3
                        err = abs(ht) * s5 / np.sqrt(den * d) if den > 0.0 else 0.0
                    else:
                        err = float(np.sqrt(np.mean((ht * sum(Ev[j] * k[j] for j in range(S + 1)) / sc) ** 2)))
                    accept = err <= 1.0
                    if accept:
                        fac = DOPRI5_MAX_FACTOR if err == 0.0 else min(DOPRI5_MAX_FACTOR, max(DOPRI5_MIN_FACTOR, DOPRI5_SAFETY * err ** expo))
                        if rejected:
                            fac = min(1.0, fac)
                    else:
                        fac = max(DOPRI5_MIN_FACTOR, DOPRI5_SAFETY * err ** expo)
                if accept:
                    rejected = False
                    n_acc[e] += 1
                    t, y, k[0] = t_new, y_new, k[S]
                    if tb["adaptive"]:
                        h = max(h, ht * fac) if (last and fac >= 1.0) else ht * fac
                    if last:
                        break
                else:
                    h = ht * fac
                    rejected = True
                    n_rej[e] += 1
            Y[row, e] = y
        h_last[e] = h
    return Y, n_acc, n_rej, h_last


# ------------------------------------------------------------------------------------------------------------
# lock-step protocol of the vectorized env (quantrl/envs/base.py:440-483, 485-499, 1233-1293)
# ------------------------------------------------------------------------------------------------------------
def vec_env_steps(integrate, y0, actions, K, S, T, reward_fn, obs_range=(-1e12, 1e12)):
    """Restates ``BaseSB3Env.step_wait`` for a noise-free deterministic vec env: per call ``update`` (integrate the next
    ``T[t_idx : t_idx+K+1]`` from ``States[-1]``, reward on all rows, ``t_idx += dim-1``, ``terminated = not t_idx+1 <
    shape_T``), ``rewards += Reward[dim-1]``, global truncation ``max(Obs) > hi or min(Obs) < lo`` over the whole block,
    and on ``terminated or truncated``: ``data_rewards.append(rewards)`` then reset — the RESET observation is returned.

    ``integrate(T_step, y, actions_step) -> (dim, E, d)``; ``actions (n_steps, E, n_act)``.  Returns a dict of per-call
    records with the keys of fixture G9."""
    E = y0.shape[0]
    rec = {k: [] for k in ("obs", "reward", "done", "t_idx", "batch_idx", "n_data_rewards", "rewards_cum")}
    y, t_idx, batch_idx, rewards, data_rewards = y0.copy(), 0, 0, np.zeros(E), []
    for a in range(actions.shape[0]):
        T_step = T[t_idx:t_idx + K + 1] if t_idx + K < S + 1 else T[t_idx:]
        dim = len(T_step)
        Y = integrate(T_step, y, actions[a])
        Reward = reward_fn(Y)
        t_idx += dim - 1
        terminated = not t_idx + 1 < S + 1
        rewards = rewards + Reward[dim - 1]
        truncated = bool(Y.max() > obs_range[1] or Y.min() < obs_range[0])
        obs, y = Y[dim - 1], Y[dim - 1]
        if terminated or truncated:
            data_rewards.append(rewards)
            y, t_idx, batch_idx, rewards = y0.copy(), 0, batch_idx + 1, np.zeros(E)
            obs = y0.copy()
        for k, v in (("obs", obs), ("reward", Reward[dim - 1]), ("done", terminated or truncated), ("t_idx", t_idx),
                     ("batch_idx", batch_idx), ("n_data_rewards", len(data_rewards)), ("rewards_cum", rewards)):
            rec[k].append(np.array(v))
    out = {k: np.array(v) for k, v in rec.items()}
    out["data_rewards"] = np.array(data_rewards)
    return out
