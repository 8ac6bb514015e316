"""GPU parity of the fused deterministic interval (qrl_ode_step through the C ABI): RK4 against the fixture made by
running classic RK4 on the reference's own ``env.func`` (G2), dopri5 against the oracle controller and against the
reference's tight-tolerance SciPy trajectories (G1).  Tolerance: FP64, rtol 1e-10 (north_star)."""
import os

import numpy as np
import pytest
import torch

import lyap_oracle as lo
import systems as sy
from helpers import dev, ode_step

pytestmark = pytest.mark.gpu

CASES = [("lyap_optomech.npz", "E5", "optomech", "optomech"), ("lyap_kerr.npz", "N2", "kerr", "kerr_chain"),
         ("lyap_kerr.npz", "N4", "kerr", "kerr_chain")]


def _load(golden_dir, fname, tag):
    g = np.load(os.path.join(golden_dir, fname))
    return g, g[f"{tag}_p"], g[f"{tag}_y0"], g[f"{tag}_actions"], g[f"{tag}_T"], int(g[f"{tag}_K"]), int(g[f"{tag}_m"]), int(g[f"{tag}_q"])


def _run_episode(system, m, q, p, y0, acts, T, K, **kw):
    y, rows, h, steps = y0, [y0[None].copy()], None, 0
    for a in range(acts.shape[0]):
        out = ode_step(system, m, q, y, K, float(T[a * K]), float(T[1] - T[0]), params=p, actions=acts[a][:, 0], h_state=h, **kw)
        rows.append(out["states"].cpu().numpy()[1:])
        np.testing.assert_array_equal(out["y_last"].cpu().numpy(), out["states"][-1].cpu().numpy())
        y, h = out["y_last"].cpu().numpy(), out["h_state"]
        steps += int(out["n_steps"].sum().item())
        assert int(out["nan"].item()) == 0
    return np.concatenate(rows, 0), steps


@pytest.mark.parametrize("fname,tag,osys,system", CASES)
def test_rk4_matches_rk4_on_reference_func(native, cuda, golden_dir, fname, tag, osys, system):
    g, p, y0, acts, T, K, m, q = _load(golden_dir, fname, tag)
    for sub, key in ((1, "rk4"), (2, "rk4x2")):
        St, _ = _run_episode(system, m, q, p, y0, acts, T, K, method="rk4", substeps=sub)
        ref = g[f"{tag}_States_{key}_reffunc"]
        np.testing.assert_allclose(St, ref, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("fname,tag,osys,system", CASES)
def test_dopri5_tight_matches_reference_scipy(native, cuda, golden_dir, fname, tag, osys, system):
    """Per-env dopri5 vs the reference's LinearizedHOVecEnv + SciPy dop853, both at atol 1e-14 / rtol 1e-13
    (fixture G1): both are converged, so the shared-vs-per-env controller no longer matters.  (The reference's own
    two tight solvers, vode and dop853, agree only to 2e-9 on the N=4 chain.)"""
    g, p, y0, acts, T, K, m, q = _load(golden_dir, fname, tag)
    St, steps = _run_episode(system, m, q, p, y0, acts, T, K, method="dopri5", atol=1e-14, rtol=1e-13)
    ref = g[f"{tag}_States_dop853_tight"]
    scale = np.abs(ref).max(axis=(0, 1), keepdims=True) + 1e-300
    err = np.max(np.abs(St - ref) / scale)
    assert err < 1e-10, err
    assert steps > acts.shape[0] * p.shape[0]


def _single_rhs(osys, p, m, q):
    mr, fA, fD = (getattr(sy, f"{osys}_{k}") for k in ("mode_rates", "A", "D"))

    def f(t, y, u, e):
        pe = p[e:e + 1]
        mo = (y[:m] + 1j * y[m:2 * m])[None]
        out = np.empty_like(y)
        r = mr(pe, t, mo, np.array([u]))
        out[:m], out[m:2 * m] = r.real[0], r.imag[0]
        A, D, V = fA(pe, t, mo, None)[0], fD(pe, t, mo, None)[0], y[2 * m:].reshape(q, q)
        out[2 * m:] = (A @ V + V @ A.T + D).ravel()
        return out
    return f


@pytest.mark.parametrize("flags", [0, 1])   # 0: default (q = 4, small batch: 16 lanes per env), 1: one thread per env
@pytest.mark.parametrize("fname,tag,osys,system", CASES)    # (N4: q = 8 -> the row-lanes kernel by default)
def test_dopri5_default_tolerance_follows_oracle_controller(native, cuda, golden_dir, fname, tag, osys, system, flags):
    """Same controller, same accept/reject sequence: both dopri5 kernels == oracle/lyap_oracle.py:dopri5_dense_interval
    at the reference's default tolerances (atol 1e-9, rtol 1e-6), including the carried step size and the step counts."""
    g, p, y0, acts, T, K, m, q = _load(golden_dir, fname, tag)
    f = _single_rhs(osys, p, m, q)
    y, h, yk, hk = y0, None, y0, None
    for a in range(4):
        Y, nacc, nrej, h = lo.dopri5_dense_interval(f, float(T[a * K]), 0.01, K, y, lambda e: acts[a][e, 0], 1e-9, 1e-6, h)
        out = ode_step(system, m, q, yk, K, float(T[a * K]), 0.01, params=p, actions=acts[a][:, 0], method="dopri5",
                       atol=1e-9, rtol=1e-6, h_state=hk, flags=flags)
        np.testing.assert_allclose(out["states"].cpu().numpy(), Y, rtol=1e-9, atol=1e-11)
        np.testing.assert_array_equal(out["n_steps"].cpu().numpy(), np.stack([nacc, nrej], 1))
        np.testing.assert_allclose(out["h_state"].cpu().numpy(), h, rtol=1e-6)
        y, yk, hk = Y[-1], out["y_last"].cpu().numpy(), out["h_state"]


def test_const_AD_lyapunov_fixed_point_and_oracle(native, cuda):
    from scipy.linalg import solve_continuous_lyapunov
    rng = np.random.default_rng(0)
    for q in (2, 4, 6, 8):
        E = 7
        A = -np.eye(q) * 1.5 + 0.2 * rng.standard_normal((E, q, q))
        Dm = rng.standard_normal((E, q, q)); Dm = Dm @ np.transpose(Dm, (0, 2, 1)) + 0.1 * np.eye(q)
        V0 = np.broadcast_to(np.eye(q).reshape(1, -1), (E, q * q)).copy()
        func = lo.make_func(0, q, None, lambda t, m, u: A, lambda t, m, u: Dm, E)
        ref = lo.rk4_interval(func, np.arange(21) * 0.05, V0, None)
        out = ode_step("const_AD", 0, q, V0, 20, 0.0, 0.05, A=A, D=Dm)
        np.testing.assert_allclose(out["states"].cpu().numpy(), ref, rtol=1e-11, atol=1e-13)
        # long dopri5 run converges to the algebraic Lyapunov solution
        y = V0
        for _ in range(6):
            o = ode_step("const_AD", 0, q, y, 100, 0.0, 0.1, A=A, D=Dm, method="dopri5", atol=1e-12, rtol=1e-10)
            y = o["y_last"].cpu().numpy()
        for e in range(E):
            Vs = solve_continuous_lyapunov(A[e], -Dm[e])
            np.testing.assert_allclose(y[e].reshape(q, q), Vs, rtol=1e-7, atol=1e-8)


def test_config2_size_properties(native, cuda):
    """BASELINE config 2 size (E=1024, RK4, K=100): envs are independent (a sub-batch reproduces its rows bitwise),
    V stays symmetric, identical envs give identical rows, fed observation noise is added on top of the states."""
    E, K, dt = 1024, 100, 0.01
    p = sy.optomech_default_params(E, seed=2)
    y0 = sy.optomech_y0(p)
    u = np.random.default_rng(3).uniform(-1, 1, E)
    noise = 1e-3 * np.random.default_rng(4).standard_normal((K + 1, E, 20))
    out = ode_step("optomech", 2, 4, y0, K, 0.0, dt, params=p, actions=u, obs_noise=noise)
    S = out["states"]
    sub = ode_step("optomech", 2, 4, y0[100:164], K, 0.0, dt, params=p[100:164], actions=u[100:164])
    assert torch.equal(S[:, 100:164], sub["states"])
    V = S[:, :, 4:].reshape(K + 1, E, 4, 4)
    assert (V - V.transpose(2, 3)).abs().max().item() < 1e-12
    torch.testing.assert_close(out["observations"], S + dev(noise), rtol=0, atol=1e-15)
    mm = out["minmax"].cpu().numpy()
    ob = out["observations"].cpu().numpy()
    assert mm[0] == ob.min() and mm[1] == ob.max()
    p2 = np.broadcast_to(p[:1], (33, 7)).copy()
    same = ode_step("optomech", 2, 4, sy.optomech_y0(p2), 20, 0.0, dt, params=p2, actions=np.full(33, 0.3))["states"]
    assert torch.equal(same[:, 0], same[:, 32])
    # RK4 error scales as h^4: halving the step cuts the distance to a converged solution by ~16
    ref = ode_step("optomech", 2, 4, y0[:64], 20, 0.0, dt, params=p[:64], actions=u[:64], method="dopri5", atol=1e-14, rtol=1e-13)["y_last"]
    e1 = (ode_step("optomech", 2, 4, y0[:64], 20, 0.0, dt, params=p[:64], actions=u[:64], substeps=1)["y_last"] - ref).abs().max().item()
    e2 = (ode_step("optomech", 2, 4, y0[:64], 20, 0.0, dt, params=p[:64], actions=u[:64], substeps=2)["y_last"] - ref).abs().max().item()
    assert 10 < e1 / e2 < 24


def test_philox_measurement_noise_moments(native, cuda):
    E, K = 512, 20
    p = sy.optomech_default_params(E, seed=2)
    out = ode_step("optomech", 2, 4, sy.optomech_y0(p), K, 0.0, 0.01, params=p, actions=np.zeros(E),
                   philox=dict(seed=11, episode=1, t_idx=40), obs_std=np.full(20, 0.5))
    z = ((out["observations"] - out["states"]) / 0.5).flatten()
    m = z.numel()
    assert abs(z.mean().item()) < 5 / np.sqrt(m) and abs(z.var().item() - 1) < 5 * np.sqrt(2 / m)
    assert abs((z ** 4).mean().item() - 3) < 5 * np.sqrt(96 / m)


def test_ode_edge_cases_and_errors(native, cuda):
    from quantrl_b200 import _native as N
    p = sy.optomech_default_params(3, seed=2)
    y0 = sy.optomech_y0(p)
    out = ode_step("optomech", 2, 4, y0, 0, 0.0, 0.01, params=p, actions=np.zeros(3))      # K = 0
    np.testing.assert_array_equal(out["obs_last"].cpu().numpy(), y0)
    a = N.OdeArgs(); a.system, a.num_modes, a.num_quads, a.n_envs, a.K, a.method, a.substeps, a.t_ssz = 1, 2, 5, 1, 1, 0, 1, 0.1
    yk = dev(y0)
    a.y0 = N.ptr(yk)
    with pytest.raises(N.NativeError, match="num_modes=2, num_quads=4"):
        N.ode_step(a)
    a.system = 77
    with pytest.raises(N.NativeError, match="unknown system"):
        N.ode_step(a)
    a.system, a.num_modes, a.num_quads = N.SYS_KERR_CHAIN, 5, 10
    a.params = a.actions = N.ptr(yk); a.n_actions = 1
    with pytest.raises(N.NativeError, match="num_modes 1..4"):
        N.ode_step(a)
    a.method = 99
    with pytest.raises(N.NativeError, match="unknown method"):
        N.ode_step(a)
    a.method = N.ODE_TSIT5     # adaptive methods need tolerances
    with pytest.raises(N.NativeError, match="atol and rtol must be positive"):
        N.ode_step(a)
    # rows move as 16-byte pieces: a state tensor at an odd 8-byte offset is refused, not mis-written
    a.system, a.num_modes, a.num_quads, a.method = N.SYS_OPTOMECH, 2, 4, N.ODE_RK4
    pp = dev(sy.optomech_default_params(2, seed=2)); uu = dev(np.zeros((2, 1)))
    a.params, a.params_stride_e, a.actions, a.n_actions = N.ptr(pp), pp.shape[1], N.ptr(uu), 1
    buf = torch.zeros(2 * 20 + 1, dtype=torch.float64, device="cuda")
    a.n_envs, a.y0 = 2, N.ptr(buf[1:])
    with pytest.raises(N.NativeError, match="16-byte aligned"):
        N.ode_step(a)


@pytest.mark.parametrize("lanes_flags", [2, 6])   # 2: warp-specialised split kernel, 6: single-role lanes kernel
@pytest.mark.parametrize("E", [1, 3, 17, 1024])
def test_rk4_lanes_kernel_equals_thread_kernel(native, cuda, E, lanes_flags):
    """q*q lanes per env (small batches; split = mode producer / covariance lanes / row writers, or single-role) vs
    one thread per env: same RK4, same order of operations; fed noise, min/max, last rows, sub-steps that are not
    rows, ragged last block and an interval shorter than one chunk included."""
    from quantrl_b200 import _native as N
    rng = np.random.default_rng(E)
    for system, m in (("optomech", 2), ("kerr_chain", 2)):
        if system == "optomech":
            p = sy.optomech_default_params(max(E, 2), seed=2)[:E]; y0 = sy.optomech_y0(p)
        else:
            p = sy.kerr_default_params(E, 2, seed=4); y0 = sy.kerr_y0(p, 2); y0[:, :2] = 0.3
        u = rng.uniform(-1, 1, E)
        noise = 1e-3 * rng.standard_normal((31, E, 20))
        kw = dict(params=p, actions=u, obs_noise=noise, substeps=2)
        a = ode_step(system, m, 4, y0, 30, 0.5, 0.01, flags=lanes_flags, **kw)
        b = ode_step(system, m, 4, y0, 30, 0.5, 0.01, flags=N.ODE_FORCE_THREAD, **kw)
        for key in ("states", "observations", "y_last", "obs_last"):
            torch.testing.assert_close(a[key], b[key], rtol=1e-13, atol=1e-15)
        torch.testing.assert_close(a["minmax"], b["minmax"], rtol=1e-13, atol=1e-15)
        # Philox measurement noise (same stream in every kernel), 3 sub-steps per row, 2 rows (< one chunk)
        kw = dict(params=p, actions=u, substeps=3, philox=dict(seed=11, episode=2, t_idx=40, env_offset=5),
                  obs_std=np.full((E, 20), 1e-3))
        a = ode_step(system, m, 4, y0, 2, 0.5, 0.01, flags=lanes_flags, **kw)
        b = ode_step(system, m, 4, y0, 2, 0.5, 0.01, flags=N.ODE_FORCE_THREAD, **kw)
        for key in ("states", "observations", "y_last", "obs_last", "minmax"):
            torch.testing.assert_close(a[key], b[key], rtol=1e-13, atol=1e-15)
    # const_AD (no modes) through the lanes kernel
    A = -np.eye(4) * 1.5 + 0.2 * rng.standard_normal((E, 4, 4))
    Dm = rng.standard_normal((E, 4, 4))
    V0 = rng.standard_normal((E, 16))
    a = ode_step("const_AD", 0, 4, V0, 12, 0.0, 0.05, A=A, D=Dm, flags=lanes_flags)
    b = ode_step("const_AD", 0, 4, V0, 12, 0.0, 0.05, A=A, D=Dm, flags=N.ODE_FORCE_THREAD)
    torch.testing.assert_close(a["states"], b["states"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("E", [1, 5, 64])
def test_dopri5_lanes_kernel_equals_thread_kernel(native, cuda, E):
    """dopri5 with q*q lanes per env (stage derivatives in registers, error norm reduced across the env's lanes) vs the
    thread-per-env kernel: same controller decisions (step counts), trajectories equal to rounding; carried step
    sizes, fed noise, min/max and last rows included; default and tight tolerances."""
    from quantrl_b200 import _native as N
    rng = np.random.default_rng(100 + E)
    p = sy.optomech_default_params(max(E, 2), seed=2)[:E]; y0 = sy.optomech_y0(p)
    u = rng.uniform(-1, 1, E)
    noise = 1e-3 * rng.standard_normal((41, E, 20))
    for atol, rtol in ((1e-9, 1e-6), (1e-13, 1e-11)):
        kw = dict(params=p, actions=u, obs_noise=noise, method="dopri5", atol=atol, rtol=rtol)
        a = ode_step("optomech", 2, 4, y0, 40, 0.5, 0.01, flags=0, **kw)
        b = ode_step("optomech", 2, 4, y0, 40, 0.5, 0.01, flags=N.ODE_FORCE_THREAD, **kw)
        np.testing.assert_array_equal(a["n_steps"].cpu().numpy(), b["n_steps"].cpu().numpy())
        for key in ("states", "observations", "y_last", "obs_last", "minmax"):
            torch.testing.assert_close(a[key], b[key], rtol=1e-11, atol=1e-14)
        # the error estimate is a difference of O(1) stage derivatives: rounding-level differences are amplified by
        # 1 / atol in err and hence in the carried step size
        torch.testing.assert_close(a["h_state"], b["h_state"], rtol=1e-4, atol=0)
        # second interval: carried step sizes
        a2 = ode_step("optomech", 2, 4, a["y_last"].cpu().numpy(), 40, 0.9, 0.01, flags=0, h_state=a["h_state"].cpu().numpy(), **kw)
        b2 = ode_step("optomech", 2, 4, b["y_last"].cpu().numpy(), 40, 0.9, 0.01, flags=N.ODE_FORCE_THREAD, h_state=b["h_state"].cpu().numpy(), **kw)
        torch.testing.assert_close(a2["states"], b2["states"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("method,flags", [("rk4", 1), ("rk4", 2), ("dopri5", 0), ("dopri5", 1)])
def test_fused_log_negativity_reward_matches_reference_qcm(native, cuda, golden_dir, method, flags):
    """Reward functor ENTAN_LN: (1) on the reference's tight trajectory it must reproduce the reference's own
    ``QCMSolver.get_entanglement_logarithmic_negativity_2`` (fixture ``entan_ln_ref``, measure.py:401-440) — checked
    through the oracle on the kernel's States; (2) the OBS variant uses the noisy covariance; (3) cumulative reward."""
    import measure_oracle as mo
    from quantrl_b200 import _native as N
    g, p, y0, acts, T, K, m, q = _load(golden_dir, "lyap_optomech.npz", "E5")
    kw = dict(method=method, flags=flags, atol=1e-14, rtol=1e-13)
    y, rows_r, cum = y0, [], np.zeros(5)
    for a in range(acts.shape[0]):
        out = ode_step("optomech", m, q, y, K, float(T[a * K]), 0.01, params=p, actions=acts[a][:, 0],
                       reward_kind=N.REWARD_ENTAN_LN_STATE, rewards_cum=cum, **kw)
        S = out["states"].cpu().numpy()
        ref = mo.entan_ln_2(S[:, :, 4:].reshape(K + 1, 5, 4, 4))
        np.testing.assert_allclose(out["reward"].cpu().numpy(), ref, rtol=0, atol=1e-12)
        np.testing.assert_array_equal(out["reward_last"].cpu().numpy(), out["reward"][-1].cpu().numpy())
        cum = cum + ref[-1]
        np.testing.assert_allclose(out["rewards_cum"].cpu().numpy(), cum, rtol=0, atol=1e-11)
        rows_r.append(out["reward"].cpu().numpy()[1 if a else 0:])
        y = out["y_last"].cpu().numpy()
    if method == "dopri5":   # converged trajectory -> the reference's own numbers
        np.testing.assert_allclose(np.concatenate(rows_r, 0), g["E5_entan_ln_ref"], rtol=0, atol=1e-10)
    noise = 1e-2 * np.random.default_rng(0).standard_normal((K + 1, 5, 20))
    out = ode_step("optomech", m, q, y0, K, 0.0, 0.01, params=p, actions=acts[0][:, 0], obs_noise=noise,
                   reward_kind=N.REWARD_ENTAN_LN_OBS, **kw)
    O = out["observations"].cpu().numpy()
    np.testing.assert_allclose(out["reward"].cpu().numpy(), mo.entan_ln_2(O[:, :, 4:].reshape(K + 1, 5, 4, 4)), rtol=0, atol=1e-12)


@pytest.mark.parametrize("method,flags", [("rk4", 1), ("rk4", 2), ("dopri5", 0)])
@pytest.mark.parametrize("measure", ["sync_c", "sync_p", "entan_ln"])
def test_fused_synchronization_and_entanglement_rewards_match_reference(native, cuda, golden_dir, method, flags, measure):
    """Reward functors SYNC_C / SYNC_P / ENTAN_LN of all three ODE kernels vs the reference's ``QCMSolver`` numbers
    (fixture measures_optomech.npz): STATE kinds on the integrated trajectory (tight dopri5 -> the reference's own
    values), OBS kinds on states + fed measurement noise (oracle restatement, itself bit-exact on the fixture)."""
    import measure_oracle as mo
    from quantrl_b200 import _native as N
    kinds = {"sync_c": (N.REWARD_SYNC_C_OBS, N.REWARD_SYNC_C_STATE), "sync_p": (N.REWARD_SYNC_P_OBS, N.REWARD_SYNC_P_STATE),
             "entan_ln": (N.REWARD_ENTAN_LN_OBS, N.REWARD_ENTAN_LN_STATE)}[measure]
    g, p, y0, acts, T, K, m, q = _load(golden_dir, "lyap_optomech.npz", "E5")
    gm = np.load(os.path.join(golden_dir, "measures_optomech.npz"))
    kw = dict(method=method, flags=flags, atol=1e-14, rtol=1e-13)
    y, rows = y0, []
    for a in range(acts.shape[0]):
        out = ode_step("optomech", m, q, y, K, float(T[a * K]), 0.01, params=p, actions=acts[a][:, 0],
                       reward_kind=kinds[1], **kw)
        ref = mo.measures_of_states(out["states"].cpu().numpy())[measure]
        np.testing.assert_allclose(out["reward"].cpu().numpy(), ref, rtol=1e-12, atol=1e-13)
        rows.append(out["reward"].cpu().numpy()[1 if a else 0:])
        y = out["y_last"].cpu().numpy()
    if method == "dopri5":
        np.testing.assert_allclose(np.concatenate(rows, 0), gm[measure + "_ref"], rtol=1e-8, atol=1e-10)
    noise = 1e-2 * np.random.default_rng(0).standard_normal((K + 1, 5, 20))
    out = ode_step("optomech", m, q, y0, K, 0.0, 0.01, params=p, actions=acts[0][:, 0], obs_noise=noise,
                   reward_kind=kinds[0], **kw)
    ref = mo.measures_of_states(out["observations"].cpu().numpy())[measure]
    np.testing.assert_allclose(out["reward"].cpu().numpy(), ref, rtol=1e-12, atol=1e-13)


ERK_ADAPTIVE = ["bosh3", "tsit5", "adaptive_heun", "fehlberg2", "dop853"]


@pytest.mark.parametrize("flags", [0, 1])   # 0: default (q = 4, small batch: 16 lanes per env), 1: one thread per env
@pytest.mark.parametrize("method", ERK_ADAPTIVE)
@pytest.mark.parametrize("fname,tag,osys,system", CASES[:2])
def test_generic_erk_engine_follows_oracle_controller(native, cuda, golden_dir, fname, tag, osys, system, method, flags):
    """Generic explicit RK engine (csrc/ode_erk.cuh), both layouts == oracle/lyap_oracle.py:erk_interval for every
    adaptive tableau: same states, the same accept/reject counts and the same carried step size over consecutive
    intervals."""
    g, p, y0, acts, T, K, m, q = _load(golden_dir, fname, tag)
    f = _single_rhs(osys, p, m, q)
    atol, rtol = (1e-9, 1e-6) if method != "dop853" else (1e-12, 1e-10)
    y, h, yk, hk = y0, None, y0, None
    for a in range(3):
        Y, nacc, nrej, h = lo.erk_interval(method, f, float(T[a * K]), 0.01, K, y, lambda e: acts[a][e, 0], atol, rtol, h)
        out = ode_step(system, m, q, yk, K, float(T[a * K]), 0.01, params=p, actions=acts[a][:, 0], method=method,
                       atol=atol, rtol=rtol, h_state=hk, flags=flags)
        np.testing.assert_allclose(out["states"].cpu().numpy(), Y, rtol=1e-9, atol=1e-11)
        np.testing.assert_array_equal(out["n_steps"].cpu().numpy(), np.stack([nacc, nrej], 1))
        # (DOP853 at 1e-12: the two error estimates are differences at rounding level, the next step size follows them)
        np.testing.assert_allclose(out["h_state"].cpu().numpy(), h, rtol=1e-6 if method != "dop853" else 1e-3)
        assert int(out["nan"].item()) == 0
        y, yk, hk = Y[-1], out["y_last"].cpu().numpy(), out["h_state"]


@pytest.mark.parametrize("method,substeps,order", [("euler", 4, 1), ("midpoint", 2, 2)])
def test_generic_erk_fixed_step_methods(native, cuda, golden_dir, method, substeps, order):
    """Fixed-step Euler / midpoint: kernel == oracle, and halving the step reduces the error vs the reference's tight
    trajectory by ~2^order."""
    g, p, y0, acts, T, K, m, q = _load(golden_dir, "lyap_optomech.npz", "E5")
    f = _single_rhs("optomech", p, m, q)
    Y, _, _, _ = lo.erk_interval(method, f, 0.0, 0.01, K, y0, lambda e: acts[0][e, 0], substeps=substeps)
    for flags in (0, 1):   # lanes layout / one thread per env
        out = ode_step("optomech", m, q, y0, K, 0.0, 0.01, params=p, actions=acts[0][:, 0], method=method, substeps=substeps,
                       flags=flags)
        np.testing.assert_allclose(out["states"].cpu().numpy(), Y, rtol=1e-11, atol=1e-13)
    ref = g["E5_States_dop853_tight"][:K + 1]
    errs = []
    for ss in (substeps, 2 * substeps):
        o = ode_step("optomech", m, q, y0, K, 0.0, 0.01, params=p, actions=acts[0][:, 0], method=method, substeps=ss)
        errs.append(np.max(np.abs(o["states"].cpu().numpy() - ref)))
    assert 0.7 * 2 ** order < errs[0] / errs[1] < 1.4 * 2 ** order


@pytest.mark.parametrize("method,tol", [("tsit5", 1e-9), ("dop853", 1e-10), ("bosh3", 1e-6)])
def test_generic_erk_tight_tolerance_matches_reference_scipy(native, cuda, golden_dir, method, tol):
    """Converged runs of the new methods reproduce the reference's own tight SciPy trajectory (fixture G1) over the
    whole episode, and the env-level plugin accepts the reference's method names."""
    g, p, y0, acts, T, K, m, q = _load(golden_dir, "lyap_optomech.npz", "E5")
    atol, rtol = (1e-14, 1e-13) if method != "bosh3" else (1e-11, 1e-10)
    got, _ = _run_episode("optomech", m, q, p, y0, acts, T, K, method=method, atol=atol, rtol=rtol)
    ref = g["E5_States_dop853_tight"]
    assert np.max(np.abs(got - ref) / (1.0 + np.abs(ref))) < tol


def test_generic_erk_rejects_unsupported_sizes(native, cuda):
    from quantrl_b200._native import NativeError
    V0 = np.broadcast_to(np.eye(8).reshape(1, -1), (3, 64)).copy()
    A = -np.broadcast_to(np.eye(8), (3, 8, 8)).copy()
    with pytest.raises(NativeError, match="num_quads <= 4"):
        ode_step("const_AD", 0, 8, V0, 4, 0.0, 0.05, A=A, D=-A, method="tsit5")
