"""CPU: host logic of ``B200IVPSolver`` that needs no kernel — the stage-callback mode on CPU tensors (plain torch)
and the delay-system protocol of the reference's solver plugin (quantrl/solvers/base.py:96-112, 168-256)."""
import os

import numpy as np
import pytest
import torch

from quantrl_b200.solvers import B200IVPSolver


def test_interpolate_equals_the_reference_spline():
    """``interpolate`` == ``splrep``/``splev`` per component (quantrl/solvers/numpy.py:180-188), inside the window and
    extrapolated, for a batched ``(dim, E, d)`` block."""
    from scipy.interpolate import splev, splrep
    rng = np.random.default_rng(0)
    T = 0.3 + 0.05 * np.arange(11)
    Y = rng.standard_normal((11, 5, 3))
    s = B200IVPSolver(func=None, y_0=torch.zeros((5, 3), dtype=torch.float64), T=torch.arange(101) * 0.05,
                      solver_params={"step_interval": 10}, has_delay=True, func_delay=lambda t: None)
    f = s.interpolate(torch.as_tensor(T), torch.as_tensor(Y))
    for t in (0.3, 0.3173, 0.61, 0.8, 0.83, 0.28):
        ref = np.array([[splev(t, splrep(T, Y[:, e, j])) for j in range(3)] for e in range(5)])
        got = f(t)
        assert tuple(got.shape) == (5, 3)
        np.testing.assert_allclose(got.numpy(), ref, rtol=1e-12, atol=1e-13)
    # short windows (a tail interval): the order drops to what the rows determine
    f2 = s.interpolate(torch.as_tensor(T[:2]), torch.as_tensor(Y[:2]))
    np.testing.assert_allclose(f2(T[0] + 0.025).numpy(), 0.5 * (Y[0] + Y[1]), rtol=1e-13, atol=1e-15)


def test_delay_system_matches_the_reference_solver_plugin(golden_dir):
    """Fixture G6 (reference ``SciPyIVPSolver.solve_ivp``, ``has_delay=True``, vode at 1e-13/1e-12): the stage-callback
    RK4 with the refitted history function reproduces it; ``delay_interval`` overrides ``step_interval``."""
    g = np.load(os.path.join(golden_dir, "dde_scipy.npz"))
    a, b, tau, K = torch.as_tensor(g["a"]), torch.as_tensor(g["b"]), float(g["tau"]), int(g["K"])
    y0, T = torch.as_tensor(g["y0"]), torch.as_tensor(g["T"])
    calls = []

    def func(t, y, args):
        calls.append(t)
        return -a * args[2](t - tau) + b * np.sin(t) * args[0]

    s = B200IVPSolver(func=func, y_0=y0, T=T, solver_params={"method": "rk4", "step_interval": 3, "substeps": 4},
                      has_delay=True, func_delay=lambda t: y0, delay_interval=K)
    assert s.step_interval == K
    Y = s.solve_ivp(y_0=y0, params=float(g["params"]))
    assert tuple(Y.shape) == tuple(g["Y"].shape) and len(calls) == 4 * 4 * (len(g["T"]) - 1)
    np.testing.assert_allclose(Y.numpy(), g["Y"], rtol=1e-7, atol=1e-9)
    # without the refit (history frozen at y_0) the trajectory is a different one
    s2 = B200IVPSolver(func=func, y_0=y0, T=T, solver_params={"method": "rk4", "step_interval": K, "substeps": 4},
                       has_delay=False, func_delay=lambda t: y0)
    assert np.max(np.abs(s2.solve_ivp(y_0=y0, params=float(g["params"])).numpy() - g["Y"])) > 1e-2


def test_delay_validation_follows_the_reference():
    y0, T = torch.zeros(3, dtype=torch.float64), torch.arange(11) * 0.1
    with pytest.raises(AssertionError, match="delay function cannot be"):
        B200IVPSolver(func=None, y_0=y0, T=T, solver_params={"step_interval": 5}, has_delay=True, func_delay=None)
    with pytest.raises(AssertionError, match="stage-callback"):
        B200IVPSolver(func=None, y_0=y0, T=T, solver_params={"step_interval": 5}, has_delay=True,
                      func_delay=lambda t: y0, system={"system": "optomech"})
    with pytest.raises(AssertionError, match="step_interval"):
        B200IVPSolver(func=None, y_0=y0, T=T, solver_params={"step_interval": 5}, has_delay=True,
                      func_delay=lambda t: y0, delay_interval=11)


@pytest.mark.parametrize("substeps,key", [(1, "rk4"), (2, "rk4x2")])
def test_stage_callback_rk4_matches_rk4_on_the_reference_func(golden_dir, substeps, key):
    """Fixture G2 through the solver's stage-callback integrator (the host loop that serves envs without a device
    functor): ``step`` over every action interval == classic RK4 driven by the reference's own ``env.func``."""
    import lyap_oracle as lo
    import systems as sy
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    tag = "E5"
    p, y0, acts, T, K = g[f"{tag}_p"], g[f"{tag}_y0"], g[f"{tag}_actions"], g[f"{tag}_T"], int(g[f"{tag}_K"])
    rhs = lo.make_func(2, 4, lambda t, mo, u: sy.optomech_mode_rates(p, t, mo, u), lambda t, mo, u: sy.optomech_A(p, t, mo, u),
                       lambda t, mo, u: sy.optomech_D(p, t, mo, u), p.shape[0])
    # hooks in NumPy behind a torch signature (the integrator under test is the torch host loop, not the hooks)
    func = lambda t, y, args: torch.as_tensor(rhs(float(t), y.numpy(), args[0].numpy()[:, 0]))  # noqa: E731
    s = B200IVPSolver(func=func, y_0=torch.as_tensor(y0), T=torch.as_tensor(T),
                      solver_params={"method": "rk4", "step_interval": K, "substeps": substeps})
    y, rows = torch.as_tensor(y0), [torch.as_tensor(y0)[None]]
    for a in range(acts.shape[0]):
        Y = s.step(T_step=torch.as_tensor(T[a * K:(a + 1) * K + 1]), y_0=y, params=torch.as_tensor(acts[a]))
        assert tuple(Y.shape) == (K + 1, *y0.shape) and torch.equal(Y[0], y)
        rows.append(Y[1:])
        y = Y[-1]
    np.testing.assert_allclose(torch.cat(rows).numpy(), g[f"{tag}_States_{key}_reffunc"], rtol=1e-12, atol=1e-13)
