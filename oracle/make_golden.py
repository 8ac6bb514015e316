"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):  ``python oracle/make_golden.py``.
The reference has no tests/golden vectors of its own, so these fixtures *are* the pin: every array
below is produced by the reference's own classes (``LinearEnv``, ``LinearizedHOEnv``, ``LinearizedHOVecEnv``,
``SciPyIVPSolver``) driven by the NumPy system hooks of ``oracle/systems.py``.

Fixtures
  em_linear_env.npz   G3: ``LinearEnv`` (quantrl/envs/stochastic.py) for 6 envs, n=4, S=200, K=20 and 3 envs, n=8:
                      the exact ``Ws`` and ``Observation_noises`` it drew, States/Observations per sub-step,
                      rewards, plus ``env.data`` rows.
  em_linear_env_tdep.npz  G3b: the same env with hooks that depend on the time index.
  lyap_optomech.npz   G1/G2/G4: ``LinearizedHOVecEnv`` + SciPy at tight tolerance (dop853 and vode), default-tolerance
                      vode, RK4 driven by the reference's own ``env.func``, the single-env twin, and the packed
                      ``env.data`` cache of ``BaseSB3Env.update_data``.
  lyap_optomech_props.npz  G8: vec env with ``n_properties = 2``: the full packed ``data`` cache incl. the property columns.
  lyap_protocol.npz   G9: the vec env's step protocol across terminations and a truncation (returns, counters, resets).
  lyap_config1.npz    G7: BASELINE config 1 exactly — one ``LinearizedHOEnv``, 1000 sub-steps (vode default, dop853 tight).
  dde_scipy.npz       G6: a delay system through ``SciPyIVPSolver.solve_ivp`` with ``has_delay=True`` (vode, tight).
  lyap_kerr.npz       same for the kerr_chain system (N=2 -> q=4, N=4 -> q=8).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import systems  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
TMP = "/tmp/qrl_golden_scratch"

quantrl = ref_shim.load()
from quantrl.envs.stochastic import LinearEnv  # noqa: E402
from quantrl.envs.deterministic import LinearizedHOEnv, LinearizedHOVecEnv  # noqa: E402
from quantrl.backends import context_manager as bcm  # noqa: E402

COMMON = dict(plot=False, cache_all_data=False, dir_prefix=TMP)


def fresh_backend():
    """The reference keeps one backend singleton (and one SeedSequence) per process; reset it so that ``seed``
    is honoured for each env we build (quantrl/backends/context_manager.py:14,38; backends/base.py:101-120)."""
    bcm.BACKEND_INSTANCES.clear()


# ---------------------------------------------------------------------------------------------------------
# G3: stochastic LinearEnv
# ---------------------------------------------------------------------------------------------------------
def make_linear_env(n, A0, Au, nz, x0, weights, seed, K, S, ssz, stds):
    class Env(LinearEnv):
        default_params = {}

        def reset_states(self):
            return np.array(x0, dtype=np.float64)

        def get_A(self, t_idx, args):
            return A0 + np.einsum("a,aij->ij", np.asarray(args[0]), Au)

        def get_noise_prefixes(self, t_idx, args):
            return nz

        def get_reward(self):
            return -np.sum(weights * self.Observations * self.Observations, axis=-1)

    fresh_backend()
    return Env(name="lin", desc="golden", params={}, t_norm_max=S * ssz + ssz / 2, t_norm_ssz=ssz, t_norm_mul=1.0,
               n_observations=n, n_properties=0, n_actions=Au.shape[0], action_maximums=[1.0] * Au.shape[0],
               action_interval=K, data_idxs=list(range(1 + Au.shape[0] + n + 1)), backend_library="numpy",
               seed=seed, observation_stds=list(stds), **COMMON)


def golden_em():
    out = {}
    for tag, n, n_env, n_act in (("n4", 4, 6, 1), ("n8", 8, 3, 2)):
        K, S, ssz = 20, 200, 0.01
        A0, Au = systems.affine_matrices(n, n_act, seed=0 if n == 4 else 4)
        rng = np.random.default_rng(11 + n)
        nz = np.full(n, 0.05) * (1.0 + 0.1 * np.arange(n))
        stds = 0.01 * (1.0 + np.arange(n))
        weights = 1.0 / (1.0 + np.arange(n))
        n_steps = S // K
        actions = rng.uniform(-1, 1, (n_env, n_steps, n_act))
        x0 = 1.0 + 0.1 * rng.standard_normal((n_env, n))
        Ws, ON, St, Ob, Rw, data = [], [], [], [], [], []
        for e in range(n_env):
            env = make_linear_env(n, A0, Au, nz, x0[e], weights, seed=100 + e, K=K, S=S, ssz=ssz, stds=stds)
            assert env.shape_T[0] == S + 1, env.shape_T
            env.reset()
            States = [env.States[-1].copy()]
            Obs = [env.Observations[-1].copy()]
            Rew = [0.0]
            for a in range(n_steps):
                env.step(actions[e, a])
                States.extend(env.States[1:].copy())
                Obs.extend(env.Observations[1:].copy())
                Rew.extend(np.asarray(env.Reward)[1:].copy())
            Ws.append(env.Ws.copy())
            ON.append(env.Observation_noises.copy())
            St.append(np.array(States))
            Ob.append(np.array(Obs))
            Rw.append(np.array(Rew))
            data.append(np.array(env.data))
        out.update({f"{tag}_A0": A0, f"{tag}_Au": Au, f"{tag}_nz": nz, f"{tag}_stds": stds, f"{tag}_weights": weights,
                    f"{tag}_actions": actions, f"{tag}_x0": x0, f"{tag}_K": K, f"{tag}_S": S, f"{tag}_ssz": ssz,
                    f"{tag}_Ws": np.array(Ws), f"{tag}_obs_noises": np.array(ON), f"{tag}_States": np.array(St),
                    f"{tag}_Observations": np.array(Ob), f"{tag}_Reward": np.array(Rw), f"{tag}_data": np.array(data)})
    np.savez_compressed(os.path.join(GOLDEN, "em_linear_env.npz"), **out)
    print("em_linear_env.npz", {k: np.shape(v) for k, v in out.items()})


def tdep_hooks(A0, Au, B, nz):
    """Time-INDEX dependent hooks of the G3b fixture (``get_A(t_idx, args)``, ``get_noise_prefixes(t_idx, args)`` receive the
    integer index, stochastic.py:218-242); shared by the generator and the tests."""
    def get_A(t_idx, args):
        return A0 + np.einsum("a,aij->ij", np.asarray(args[0]), Au) + 0.2 * np.sin(0.3 * t_idx) * B

    def get_nz(t_idx, args):
        return nz * (1.0 + 0.5 * np.cos(0.1 * t_idx))
    return get_A, get_nz


def golden_em_time_dependent():
    """G3b: ``LinearEnv`` with hooks that depend on the time index.  (No ragged last interval: with
    ``(shape_T - 1) % action_interval != 0`` the reference's ``LinearEnv`` raises in ``BaseEnv.update`` — its States block
    keeps K+1 rows while the noise slice has ``dim`` rows, base.py:462 — so there is nothing to pin for that case.)"""
    n, n_env, n_act, K, S, ssz = 4, 4, 1, 20, 200, 0.01
    A0, Au = systems.affine_matrices(n, n_act, seed=0)
    rng = np.random.default_rng(77)
    B = rng.standard_normal((n, n))
    nz = np.full(n, 0.05) * (1.0 + 0.1 * np.arange(n))
    stds = 0.01 * (1.0 + np.arange(n))
    weights = 1.0 / (1.0 + np.arange(n))
    get_A, get_nz = tdep_hooks(A0, Au, B, nz)
    n_steps = -(-S // K)
    actions = rng.uniform(-1, 1, (n_env, n_steps, n_act))
    x0 = 1.0 + 0.1 * rng.standard_normal((n_env, n))
    res = {k: [] for k in ("Ws", "obs_noises", "States", "Observations", "Reward", "data", "dims")}
    for e in range(n_env):
        env = make_linear_env(n, A0, Au, nz, x0[e], weights, seed=300 + e, K=K, S=S, ssz=ssz, stds=stds)
        env.get_A, env.get_noise_prefixes = get_A, get_nz
        assert env.shape_T[0] == S + 1 and env.action_steps == n_steps, (env.shape_T, env.action_steps)
        env.reset()
        States, Obs, Rew, dims = [env.States[-1].copy()], [env.Observations[-1].copy()], [0.0], []
        done = False
        for a in range(n_steps):
            assert not done
            _, _, done, _, _ = env.step(actions[e, a])
            dim = S + 1 - len(States) + 1 if a == n_steps - 1 else K + 1
            dims.append(dim)
            States.extend(env.States[1:dim].copy())
            Obs.extend(env.Observations[1:dim].copy())
            Rew.extend(np.asarray(env.Reward)[1:dim].copy())
        assert done and len(States) == S + 1
        for k, v in (("Ws", env.Ws), ("obs_noises", env.Observation_noises), ("States", States), ("Observations", Obs),
                     ("Reward", Rew), ("data", env.data), ("dims", dims)):
            res[k].append(np.array(v))
    out = {"A0": A0, "Au": Au, "B": B, "nz": nz, "stds": stds, "weights": weights, "actions": actions, "x0": x0,
           "K": K, "S": S, "ssz": ssz}
    out.update({k: np.array(v) for k, v in res.items()})
    np.savez_compressed(os.path.join(GOLDEN, "em_linear_env_tdep.npz"), **out)
    print("em_linear_env_tdep.npz", {k: np.shape(v) for k, v in out.items()})


# ---------------------------------------------------------------------------------------------------------
# G1/G2/G4: deterministic LinearizedHO(Vec)Env
# ---------------------------------------------------------------------------------------------------------
def make_vec_env(system, p, m, q, y0, K, S, ssz, method, atol, rtol, cls=LinearizedHOVecEnv, n_props=0, props_fn=None,
                 **extra):
    E = p.shape[0]
    mode_rates = getattr(systems, system + "_mode_rates")
    fA = getattr(systems, system + "_A")
    fD = getattr(systems, system + "_D")
    single = cls is LinearizedHOEnv

    def act(args):
        a = np.asarray(args[0], dtype=np.float64)
        return a.reshape(-1, a.shape[-1])[:, 0]

    class Env(cls):
        default_params = {}

        def reset_states(self):
            return y0[0].copy() if single else y0.copy()

        def get_mode_rates(self, t, modes, args):
            r = mode_rates(p, t, modes.reshape(E, m), act(args))
            return r[0] if single else r

        def get_A(self, t, modes, args):
            r = fA(p, t, modes.reshape(E, m), act(args))
            return r[0] if single else r

        def get_D(self, t, modes, args):
            r = fD(p, t, modes.reshape(E, m), act(args))
            return r[0] if single else r

        def get_reward(self):
            # toy reward: minus the trace of the covariance block of the *observations*
            V = self.Observations[..., 2 * m:].reshape(*self.Observations.shape[:-1], q, q)
            return -np.trace(V, axis1=-2, axis2=-1)

        def get_properties(self):
            return props_fn(self.Observations)

    fresh_backend()
    d = 2 * m + q * q
    # the env's grid is one action interval longer than what we step, so that the vectorized env never reaches
    # ``terminated`` (its auto-reset inside ``step_wait`` would wipe ``States``/``data``; base.py:1277-1291)
    kw = dict(name=system, desc="golden", params={}, num_modes=m, num_quads=q, t_norm_max=(S + K) * ssz + ssz / 2,
              t_norm_ssz=ssz, t_norm_mul=1.0, n_observations=d, n_properties=n_props, n_actions=1, action_maximums=[1.0],
              action_interval=K, data_idxs=[0, 1, 2, 2 + d] if n_props == 0 else list(range(1 + 1 + d + n_props + 1)),
              backend_library="numpy", ode_method=method,
              ode_atol=atol, ode_rtol=rtol, **COMMON, **extra)
    if not single:
        kw["n_envs"] = E
    return Env(**kw)


def run_vec(env, actions, single=False):
    obs0 = env.reset()
    obs0 = obs0[0] if isinstance(obs0, tuple) else obs0
    St = [np.array(env.States[-1])]
    Rw = []
    for a in range(actions.shape[0]):
        if single:
            env.step(actions[a, 0])
        else:
            env.step_async(actions[a])
            env.step_wait()
        St.extend(np.array(env.States[1:]))
        Rw.append(np.array(env.Reward[-1]))
    return np.array(St), np.array(Rw), np.array(obs0)


def rk4_on_reference_func(env, y0, T, K, actions, substeps=1):
    """G2: classic RK4 whose RHS *is the reference's own* ``env.func`` (copy its reused output buffer)."""
    f = lambda t, y, args: np.array(env.func(t, y, args), copy=True)  # noqa: E731
    Y = [y0.copy()]
    y = y0.copy()
    S = len(T) - 1
    for i in range(S):
        args = [actions[i // K], None, None]
        h = (T[i + 1] - T[i]) / substeps
        for s in range(substeps):
            t = T[i] + s * h
            k1 = f(t, y, args)
            k2 = f(t + 0.5 * h, y + (0.5 * h) * k1, args)
            k3 = f(t + 0.5 * h, y + (0.5 * h) * k2, args)
            k4 = f(t + h, y + h * k3, args)
            y = y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        Y.append(y.copy())
    return np.array(Y)


def golden_lyap(system, tag_sizes, fname):
    out = {}
    for tag, E, N in tag_sizes:
        K, S, ssz = 10, 100, 0.01
        if system == "optomech":
            m, q = 2, 4
            p = systems.optomech_default_params(E, seed=2)
            p[:, 4] = np.linspace(1e-2, 3e-2, E)  # vary g0 too
            y0 = systems.optomech_y0(p)
        else:
            m, q = N, 2 * N
            p = systems.kerr_default_params(E, N, seed=4)
            y0 = systems.kerr_y0(p, N)
            y0[:, :N] = 0.3  # non-zero initial amplitudes
        rng = np.random.default_rng(3)
        n_steps = S // K
        actions = rng.uniform(-1, 1, (n_steps, E, 1))
        res = {"p": p, "y0": y0, "actions": actions, "K": K, "S": S, "ssz": ssz, "m": m, "q": q}
        for name, method, atol, rtol in (("dop853_tight", "dop853", 1e-14, 1e-13), ("vode_tight", "vode", 1e-14, 1e-13),
                                         ("vode_default", "vode", 1e-9, 1e-6)):
            env = make_vec_env(system, p, m, q, y0, K, S, ssz, method, atol, rtol)
            St, Rw, _ = run_vec(env, actions)
            res["States_" + name] = St
            res["Reward_" + name] = Rw
            if name == "dop853_tight":
                res["data_" + name] = np.array(env.data)[:, :S + 1]
                res["T"] = np.array(env.T)[:S + 1]
                res["States_rk4_reffunc"] = rk4_on_reference_func(env, y0, res["T"], K, actions, 1)
                res["States_rk4x2_reffunc"] = rk4_on_reference_func(env, y0, res["T"], K, actions, 2)
        if q == 4:
            # reward-functor pin: the reference's own QCMSolver on the tight trajectory (measure.py:401-440)
            from quantrl.solvers.measure import QCMSolver
            St = res["States_dop853_tight"]
            Corrs = St[:, :, 2 * m:].reshape(-1, q, q)
            Modes = (St[:, :, :m] + 1j * St[:, :, m:2 * m]).reshape(-1, m)
            qcm = QCMSolver(Modes=Modes, Corrs=Corrs, params={"show_progress": False, "measure_codes": ["entan_ln_2"],
                                                              "indices": (0, 1)})
            res["entan_ln_ref"] = qcm.get_entanglement_logarithmic_negativity_2(pos_i=0, pos_j=2).reshape(St.shape[:2])
        # G4: single-env twin of env 0 (quantrl/envs/deterministic.py:18-437)
        env1 = make_vec_env(system, p[:1], m, q, y0[:1], K, S, ssz, "dop853", 1e-14, 1e-13, cls=LinearizedHOEnv)
        St1, _, _ = run_vec(env1, actions[:, :1], single=True)
        res["States_single_env0"] = St1
        out.update({f"{tag}_{k}": v for k, v in res.items()})
    np.savez_compressed(os.path.join(GOLDEN, fname), **out)
    print(fname, {k: np.shape(v) for k, v in out.items()})


# ---------------------------------------------------------------------------------------------------------
# G6: delay system through the reference's solver plugin (has_delay: step = integrate + interpolate)
# ---------------------------------------------------------------------------------------------------------
def golden_dde():
    """``SciPyIVPSolver.solve_ivp`` with ``has_delay=True`` (quantrl/solvers/base.py:96-112,198-256; numpy.py:180-188) on
    ``y' = -a * y(t - tau) + b * sin(t)``, ``tau`` = one step window, constant history ``y(t <= 0) = y_0``."""
    from quantrl.solvers.numpy import SciPyIVPSolver
    from quantrl.backends.context_manager import get_backend_instance
    fresh_backend()
    backend = get_backend_instance(library="numpy", precision="double", device="cpu")
    K, S, dt = 10, 120, 0.05
    tau = K * dt
    a, b = np.array([0.5, 1.0, 2.0]), np.array([0.0, 0.3, 1.0])
    y0 = np.array([1.0, -0.5, 0.25])
    T = np.arange(S + 1) * dt

    def func(t, y, args):
        return -a * np.asarray(args[2](t - tau), dtype=np.float64) + b * np.sin(t) * args[0]

    solver = SciPyIVPSolver(func=func, y_0=y0, T=T, solver_params={"method": "vode", "atol": 1e-13, "rtol": 1e-12,
                                                                    "step_interval": 3}, func_controls=None,
                            has_delay=True, func_delay=lambda t: y0, delay_interval=K, backend=backend)
    assert solver.step_interval == K       # delay_interval overrides step_interval (base.py:108-109)
    Y = solver.solve_ivp(y_0=y0, params=1.5, show_progress=False)
    np.savez_compressed(os.path.join(GOLDEN, "dde_scipy.npz"), Y=Y, T=T, a=a, b=b, y0=y0, K=K, tau=tau, params=1.5)
    print("dde_scipy.npz", Y.shape, Y[-1])


def optomech_props(Obs):
    """Two properties of the observed state (shared by the generator and the tests): the covariance trace and |alpha|^2."""
    return np.stack([Obs[..., 4] + Obs[..., 9] + Obs[..., 14] + Obs[..., 19], Obs[..., 0] ** 2 + Obs[..., 2] ** 2], axis=-1)


def golden_props():
    """G8: ``LinearizedHOVecEnv`` with ``n_properties = 2`` (``get_properties`` hook, envs/base.py:296-305, 464-469) and
    every column selected: the packed ``data`` cache ``[t, action, obs(20), props(2), cumulative reward]`` of
    ``BaseSB3Env.update_data`` (envs/base.py:1314-1372)."""
    E, K, S, ssz = 3, 10, 50, 0.01
    p = systems.optomech_default_params(E, seed=2)
    y0 = systems.optomech_y0(p)
    actions = np.random.default_rng(8).uniform(-1, 1, (S // K, E, 1))
    env = make_vec_env("optomech", p, 2, 4, y0, K, S, ssz, "dop853", 1e-14, 1e-13, n_props=2, props_fn=optomech_props)
    St, Rw, _ = run_vec(env, actions)
    out = {"p": p, "y0": y0, "actions": actions, "K": K, "S": S, "ssz": ssz, "States": St, "Reward": Rw,
           "data": np.array(env.data)[:, :S + 1], "T": np.array(env.T)[:S + 1]}
    np.savez_compressed(os.path.join(GOLDEN, "lyap_optomech_props.npz"), **out)
    print("lyap_optomech_props.npz", {k: np.shape(v) for k, v in out.items()})


def golden_protocol():
    """G9: the SB3 ``VecEnv`` protocol of the reference ``LinearizedHOVecEnv`` ACROSS episode ends (envs/base.py:1233-1293):
    S = 30, K = 10, seven ``step_async`` / ``step_wait`` calls — termination after every third step (the RESET observation
    is returned, no ``terminal_observation``), and the same with an observation range that truncates mid-episode (any env
    out of range ends the episode for all, base.py:485-499)."""
    E, K, S, ssz = 3, 10, 30, 0.01
    p = systems.optomech_default_params(E, seed=2)
    y0 = systems.optomech_y0(p)
    actions = np.random.default_rng(9).uniform(-1, 1, (7, E, 1))
    out = {"p": p, "y0": y0, "actions": actions, "K": K, "S": S, "ssz": ssz}
    for tag, rng_ in (("term", [-1e12, 1e12]), ("trunc", None)):
        if rng_ is None:
            # the optical amplitude grows monotonically from 0 under the drive: pick the range it crosses during step 2
            free = out["term_obs"]
            rng_ = [-1e12, float(0.5 * (np.abs(free[0]).max() + np.abs(free[1]).max()))]
            out["trunc_range"] = np.array(rng_)
        env = make_vec_env("optomech", p, 2, 4, y0, K, S - K, ssz, "dop853", 1e-14, 1e-13, observation_space_range=rng_)
        assert env.shape_T[0] == S + 1
        obs0 = env.reset()
        rec = {k: [] for k in ("obs", "reward", "done", "t_idx", "batch_idx", "n_data_rewards", "rewards_cum")}
        for a in range(actions.shape[0]):
            env.step_async(actions[a])
            o, r, d, info = env.step_wait()
            assert info == [{}] * E and len(set(d)) == 1
            for k, v in (("obs", o), ("reward", r), ("done", d[0]), ("t_idx", env.t_idx), ("batch_idx", env.batch_idx),
                         ("n_data_rewards", len(env.data_rewards)), ("rewards_cum", env.rewards)):
                rec[k].append(np.array(v))
        out[tag + "_obs0"] = np.array(obs0)
        out.update({f"{tag}_{k}": np.array(v) for k, v in rec.items()})
        out[tag + "_data_rewards"] = np.array(env.data_rewards)
    np.savez_compressed(os.path.join(GOLDEN, "lyap_protocol.npz"), **out)
    print("lyap_protocol.npz", {k: (np.shape(v) if np.ndim(v) else v) for k, v in out.items() if "obs" not in k})


def golden_config1():
    """BASELINE config 1 exactly: ONE reference ``LinearizedHOEnv`` (m = 2, q = 4), kappa = 0.1,
    gamma = 1e-3, omega_m = 1, Delta0 = -1, g0 = 1e-2, n_th = 10, A_l = 5, t_norm_max = 10 (S = 1000), K = 10, actions
    held at ``action_maximums`` (what ``evolve`` applies): SciPy vode at the reference's default tolerances and
    dop853 tight."""
    p = systems.optomech_default_params(1, seed=2, jitter=False)
    y0 = systems.optomech_y0(p)
    K, S, ssz = 10, 1000, 0.01
    actions = np.ones((S // K, 1, 1))
    out = {"p": p, "y0": y0, "K": K, "S": S, "ssz": ssz}
    for name, method, atol, rtol in (("vode_default", "vode", 1e-9, 1e-6), ("dop853_tight", "dop853", 1e-14, 1e-13)):
        env = make_vec_env("optomech", p, 2, 4, y0, K, S, ssz, method, atol, rtol, cls=LinearizedHOEnv)
        St, Rw, _ = run_vec(env, actions, single=True)
        assert St.shape == (S + 1, 20), St.shape
        out["States_" + name], out["Reward_" + name] = St, Rw
    np.savez_compressed(os.path.join(GOLDEN, "lyap_config1.npz"), **out)
    print("lyap_config1.npz", {k: np.shape(v) for k, v in out.items()},
          "default vs tight:", np.max(np.abs(out["States_vode_default"] - out["States_dop853_tight"])))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    if "--config1-only" in sys.argv:
        golden_config1()
        sys.exit(0)
    if "--props-only" in sys.argv:
        golden_props()
        sys.exit(0)
    if "--protocol-only" in sys.argv:
        golden_protocol()
        sys.exit(0)
    if "--dde-only" in sys.argv:
        golden_dde()
        sys.exit(0)
    if "--tdep-only" in sys.argv:          # add the G3b fixture without regenerating the others
        golden_em_time_dependent()
        sys.exit(0)
    golden_em()
    golden_em_time_dependent()
    golden_lyap("optomech", [("E5", 5, 0)], "lyap_optomech.npz")
    golden_lyap("kerr", [("N2", 3, 2), ("N4", 2, 4)], "lyap_kerr.npz")
    golden_dde()
    golden_config1()
    golden_props()
    golden_protocol()
