"""GPU: the public env surface (LinearVecEnv / LinearizedHOVecEnv, mirrors of the reference's env classes) against
reference-generated fixtures: step()/reset() returns, lock-step episode bookkeeping, the packed ``data`` cache of
BaseSB3Env.update_data, auto-reset, truncation."""
import os

import numpy as np
import pytest
import torch

import em_oracle as eo
import systems as sy

pytestmark = pytest.mark.gpu


def _em_env(g, tag, **kw):
    from quantrl_b200.systems import AffineLinearVecEnv
    K, S, ssz = int(g[f"{tag}_K"]), int(g[f"{tag}_S"]), float(g[f"{tag}_ssz"])
    E, n = g[f"{tag}_x0"].shape
    return AffineLinearVecEnv(
        n_envs=E, n=n, A0=g[f"{tag}_A0"], Au=g[f"{tag}_Au"], noise_prefixes=g[f"{tag}_nz"], x0=g[f"{tag}_x0"],
        t_norm_max=S * ssz + ssz / 2, t_norm_ssz=ssz, action_interval=K, observation_stds=list(g[f"{tag}_stds"]),
        reward_weights=g[f"{tag}_weights"], dir_prefix="/tmp/qrl_test", **kw)


@pytest.mark.parametrize("tag", ["n4", "n8"])
def test_linear_vec_env_fed_matches_reference_linear_env(native, cuda, golden_dir, tag):
    """Each row of the vectorized env == the reference's single LinearEnv fed the same draws (fixture G3)."""
    g = np.load(os.path.join(golden_dir, "em_linear_env.npz"))
    env = _em_env(g, tag, rng="fed")
    K, S = int(g[f"{tag}_K"]), int(g[f"{tag}_S"])
    env.set_fed_noise(np.transpose(g[f"{tag}_Ws"], (1, 0, 2)), np.transpose(g[f"{tag}_obs_noises"], (1, 0, 2)))
    obs0 = env.reset()
    np.testing.assert_allclose(obs0.cpu().numpy(), g[f"{tag}_Observations"][:, 0], rtol=1e-13, atol=1e-15)
    assert env.action_steps == S // K and env.shape_T == (S + 1,)
    cum = 0.0
    for a in range(S // K):
        acts = g[f"{tag}_actions"][:, a]
        obs, rew, done, info = env.step(acts)
        t1 = (a + 1) * K
        ref_R = g[f"{tag}_Reward"][:, t1]
        np.testing.assert_allclose(rew.cpu().numpy(), ref_R, rtol=1e-12, atol=1e-14)
        cum = cum + ref_R
        last = a == S // K - 1
        assert done == [last] * env.n_envs and info == [{}] * env.n_envs
        if not last:
            np.testing.assert_allclose(obs.cpu().numpy(), g[f"{tag}_Observations"][:, t1], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(env.States.cpu().numpy(), np.transpose(g[f"{tag}_States"][:, t1 - K:t1 + 1], (1, 0, 2)),
                                       rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(env.rewards.cpu().numpy(), cum, rtol=1e-12)
            assert env.t_idx == t1 and env.action_idx == a + 1
        else:  # auto-reset: the reset observation is returned, no terminal_observation (base.py:1277-1293)
            np.testing.assert_allclose(obs.cpu().numpy(), g[f"{tag}_Observations"][:, 0], rtol=1e-13, atol=1e-15)
            assert env.t_idx == 0 and env.batch_idx == 1 and len(env.data_rewards) == 1
            np.testing.assert_allclose(env.data_rewards[0].cpu().numpy(), cum, rtol=1e-12)
    env.close()


def test_linear_vec_env_time_index_dependent_hooks_match_reference(native, cuda, golden_dir):
    """Fixture G3b: user hooks that depend on the integer time index (``is_A_time_dependent=True``: evaluated for every
    sub-step ``t_idx + i`` of the interval and stacked, stochastic.py:218-233) — each row of the vectorized env == the
    reference's single LinearEnv with the same hooks, fed the same draws."""
    import math
    from quantrl_b200.envs import LinearVecEnv
    g = np.load(os.path.join(golden_dir, "em_linear_env_tdep.npz"))
    K, S, ssz = int(g["K"]), int(g["S"]), float(g["ssz"])
    E, n = g["x0"].shape
    A0, Au, B, nz, x0 = (torch.as_tensor(g[k], device="cuda") for k in ("A0", "Au", "B", "nz", "x0"))

    class Env(LinearVecEnv):
        def reset_states(self):
            return x0

        def get_A(self, t_idx, args):
            return A0 + torch.einsum("ea,aij->eij", args[0], Au) + 0.2 * math.sin(0.3 * t_idx) * B

        def get_noise_prefixes(self, t_idx, args):
            return nz * (1.0 + 0.5 * math.cos(0.1 * t_idx))

    env = Env(name="tdep", desc="golden", params={}, t_norm_max=S * ssz + ssz / 2, t_norm_ssz=ssz, t_norm_mul=1.0,
              n_envs=E, n_observations=n, n_properties=0, n_actions=1, action_maximums=[1.0], action_interval=K,
              data_idxs=list(range(1 + 1 + n + 1)), rng="fed", is_A_time_dependent=True,
              observation_stds=list(g["stds"]), reward_functor="neg_wsq_obs", reward_weights=g["weights"],
              cache_all_data=False, dir_prefix="/tmp/qrl_test")
    env.set_fed_noise(np.transpose(g["Ws"], (1, 0, 2)), np.transpose(g["obs_noises"], (1, 0, 2)))
    obs0 = env.reset()
    np.testing.assert_allclose(obs0.cpu().numpy(), g["Observations"][:, 0], rtol=1e-13, atol=1e-15)
    for a in range(S // K - 1):      # stop before the terminating step (auto-reset)
        obs, rew, done, _ = env.step(g["actions"][:, a])
        t1 = (a + 1) * K
        assert done == [False] * E and env.t_idx == t1
        np.testing.assert_allclose(env.States.cpu().numpy(), np.transpose(g["States"][:, t1 - K:t1 + 1], (1, 0, 2)),
                                   rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(obs.cpu().numpy(), g["Observations"][:, t1], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(rew.cpu().numpy(), g["Reward"][:, t1], rtol=1e-12, atol=1e-14)
    env.close()


def test_linear_vec_env_data_cache_rows(native, cuda, golden_dir):
    """env.data == BaseSB3Env.update_data's rows [t, actions, obs, cumulative reward] incl. the overlap rule
    (row 0 of a block overwrites the last row of the previous block, base.py:1370-1372), via the oracle's pack_rows."""
    g = np.load(os.path.join(golden_dir, "em_linear_env.npz"))
    tag = "n4"
    env = _em_env(g, tag, rng="fed")
    K, S, ssz = int(g[f"{tag}_K"]), int(g[f"{tag}_S"]), float(g[f"{tag}_ssz"])
    E = env.n_envs
    env.set_fed_noise(np.transpose(g[f"{tag}_Ws"], (1, 0, 2)), np.transpose(g[f"{tag}_obs_noises"], (1, 0, 2)))
    env.reset()
    T = np.arange(S + 1) * ssz
    Obs = np.transpose(g[f"{tag}_Observations"], (1, 0, 2))
    ref = np.zeros((E, S + 1, 7))
    cum = np.zeros(E)
    n_steps = S // K - 1     # stop before the terminating step (its auto-reset zeroes ``data``)
    for a in range(n_steps):
        env.step(g[f"{tag}_actions"][:, a])
        t0 = a * K
        cum = cum + g[f"{tag}_Reward"][:, t0 + K]
        ref[:, t0:t0 + K + 1] = eo.pack_rows(T[t0:t0 + K + 1], g[f"{tag}_actions"][:, a], Obs[t0:t0 + K + 1], None, cum)
    np.testing.assert_allclose(env.data, ref, rtol=1e-12, atol=1e-14)
    # the reference's own Gym-env data agrees on the time / action / observation columns of the surviving rows
    gd = g[f"{tag}_data"]
    np.testing.assert_allclose(env.data[:, 1:n_steps * K, 0], gd[:, 1:n_steps * K, 0], rtol=0, atol=1e-15)
    np.testing.assert_allclose(env.data[:, :n_steps * K, 2:6], gd[:, :n_steps * K, 2:6], rtol=1e-12, atol=1e-14)
    env.close()


def test_linear_vec_env_philox_numpy_api_and_truncation(native, cuda):
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 1, seed=0)
    E, K = 100, 10
    kw = dict(n_envs=E, n=4, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.305, t_norm_ssz=0.01,
              action_interval=K, observation_stds=[0.01] * 4, seed=9, dir_prefix="/tmp/qrl_test", return_numpy=True)
    env = AffineLinearVecEnv(**kw)
    obs0 = env.reset()
    assert isinstance(obs0, np.ndarray) and obs0.shape == (E, 4)
    u = np.random.default_rng(0).uniform(-1, 1, (E, 1))
    obs, rew, done, _ = env.step(u)
    assert isinstance(obs, np.ndarray) and isinstance(rew, np.ndarray) and rew.shape == (E,) and done == [False] * E
    z_w, z_o = eo.philox_noise_block(env._philox_seed, 0, 0, K, 0, E, 4)
    ref = eo.em_interval_vec(np.ones((E, 4)), sy.affine_A(A0, Au, u), np.full((E, 4), 0.05), np.sqrt(0.01) * z_w, 0.01)
    np.testing.assert_allclose(obs, ref[-1] + 0.01 * z_o[-1], rtol=1e-10, atol=1e-12)
    # same seed -> same episode; the second episode of one env differs from its first (episode counter in the stream)
    env_b = AffineLinearVecEnv(**kw)
    env_b.reset()
    obs_b, _, _, _ = env_b.step(u)
    np.testing.assert_array_equal(obs, obs_b)
    env.step(u); o3, _, done3, _ = env.step(u)
    assert done3 == [True] * E and env.batch_idx == 1
    obs_e2, _, _, _ = env.step(u)
    assert not np.array_equal(obs_e2, obs)
    # truncation: any observation outside the range ends the episode for ALL envs (base.py:485-499, 1271-1293)
    env_t = AffineLinearVecEnv(**dict(kw, observation_space_range=[-1.05, 1.05]))
    env_t.reset()
    _, _, done_t, _ = env_t.step(u)
    assert done_t == [True] * E and env_t.batch_idx == 1 and env_t.t_idx == 0
    for e in (env, env_b, env_t):
        e.close()


def _optomech_env(g, tag, method, atol=1e-9, rtol=1e-6, fused=True, **kw):
    from quantrl_b200.envs import LinearizedHOVecEnv
    p, y0 = g[f"{tag}_p"], g[f"{tag}_y0"]
    K, S, ssz = int(g[f"{tag}_K"]), int(g[f"{tag}_S"]), float(g[f"{tag}_ssz"])
    E, d = y0.shape

    class Env(LinearizedHOVecEnv):
        def reset_states(self):
            return torch.as_tensor(y0, device="cuda")

        def get_reward(self):
            V = self.Observations[..., 4:].reshape(*self.Observations.shape[:-1], 4, 4)
            return -torch.diagonal(V, dim1=-2, dim2=-1).sum(-1)

        # torch hooks for the stage-callback mode (same formulas as oracle/systems.py)
        def get_mode_rates(self, t, modes, args):
            P, u = self._P, args[0][:, 0]
            a, b = modes[:, 0], modes[:, 1]
            da = -(P[:, 0] / 2 - 1j * P[:, 3]) * a + 1j * P[:, 4] * a * (b + b.conj()) + P[:, 6] * (1.0 + u)
            db = -(P[:, 1] / 2 + 1j * P[:, 2]) * b + 1j * P[:, 4] * (a.real ** 2 + a.imag ** 2)
            return torch.stack([da, db], dim=1)

        def get_A(self, t, modes, args):
            P = self._P
            a, b = modes[:, 0], modes[:, 1]
            Dl, GR, GI = P[:, 3] + 2 * P[:, 4] * b.real, 2 * P[:, 4] * a.real, 2 * P[:, 4] * a.imag
            A = torch.zeros((E, 4, 4), dtype=torch.float64, device="cuda")
            A[:, 0, 0] = A[:, 1, 1] = -P[:, 0] / 2
            A[:, 2, 2] = A[:, 3, 3] = -P[:, 1] / 2
            A[:, 0, 1], A[:, 1, 0], A[:, 0, 2], A[:, 1, 2] = -Dl, Dl, -GI, GR
            A[:, 2, 3], A[:, 3, 2], A[:, 3, 0], A[:, 3, 1] = P[:, 2], -P[:, 2], GR, GI
            return A

        def get_D(self, t, modes, args):
            P = self._P
            D = torch.zeros((E, 4, 4), dtype=torch.float64, device="cuda")
            D[:, 0, 0] = D[:, 1, 1] = P[:, 0] / 2
            D[:, 2, 2] = D[:, 3, 3] = P[:, 1] * (P[:, 5] + 0.5)
            return D

    env = Env(name="optomech", desc="test", params={}, num_modes=2, num_quads=4, t_norm_max=(S + K) * ssz + ssz / 2,
              t_norm_ssz=ssz, t_norm_mul=1.0, n_envs=E, n_observations=d, n_properties=0, n_actions=1,
              action_maximums=[1.0], action_interval=K, data_idxs=[0, 1, 2, 2 + d], cache_all_data=False,
              dir_prefix="/tmp/qrl_test", ode_method=method, ode_atol=atol, ode_rtol=rtol,
              **(dict(system="optomech", system_params=p) if fused else {}), **kw)
    env._P = torch.as_tensor(p, device="cuda")
    return env


def test_lho_vec_env_matches_reference_vec_env(native, cuda, golden_dir):
    """Fused LinearizedHOVecEnv (dopri5, tight) vs the reference's LinearizedHOVecEnv + SciPy dop853 (fixture G1):
    States, per-step reward and the packed ``data`` cache produced by BaseSB3Env.update_data."""
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    tag = "E5"
    env = _optomech_env(g, tag, "dopri5", atol=1e-14, rtol=1e-13)
    K, S = int(g[f"{tag}_K"]), int(g[f"{tag}_S"])
    env.reset()
    ref = g[f"{tag}_States_dop853_tight"]
    scale = np.abs(ref).max(axis=(0, 1)) + 1e-300
    for a in range(S // K):
        obs, rew, done, _ = env.step(g[f"{tag}_actions"][a])
        np.testing.assert_allclose(rew.cpu().numpy(), g[f"{tag}_Reward_dop853_tight"][a], rtol=1e-9)
        assert np.max(np.abs(env.States.cpu().numpy() - ref[a * K:(a + 1) * K + 1]) / scale) < 1e-10
        assert done == [False] * env.n_envs
    np.testing.assert_allclose(env.data[:, :S + 1], g[f"{tag}_data_dop853_tight"], rtol=1e-9, atol=1e-10)
    env.close()


@pytest.mark.parametrize("method", ["tsit5", "DOP853", "dopri8"])
def test_lho_vec_env_accepts_the_reference_plugins_method_names(native, cuda, golden_dir, method):
    """``ode_method`` takes the explicit-RK names of the reference's torchdiffeq / diffrax / SciPy plugins (generic
    engine); at tight tolerance the env reproduces the reference's LinearizedHOVecEnv + SciPy trajectory (fixture G1)."""
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    tag = "E5"
    env = _optomech_env(g, tag, method, atol=1e-14, rtol=1e-13)
    K = int(g[f"{tag}_K"])
    env.reset()
    ref = g[f"{tag}_States_dop853_tight"]
    scale = np.abs(ref).max(axis=(0, 1)) + 1e-300
    for a in range(4):
        env.step(g[f"{tag}_actions"][a])
        assert np.max(np.abs(env.States.cpu().numpy() - ref[a * K:(a + 1) * K + 1]) / scale) < 1e-9
    env.close()


def test_lho_vec_env_stage_callback_equals_fused(native, cuda, golden_dir):
    """The same env with torch hooks (no device functor; RK4 stages call ``func``) == the fused kernel == fixture G2."""
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    tag = "E5"
    fused, cb = _optomech_env(g, tag, "rk4"), _optomech_env(g, tag, "rk4", fused=False)
    fused.reset(); cb.reset()
    K = int(g[f"{tag}_K"])
    for a in range(3):
        of, rf, _, _ = fused.step(g[f"{tag}_actions"][a])
        oc, rc, _, _ = cb.step(g[f"{tag}_actions"][a])
        torch.testing.assert_close(of, oc, rtol=1e-11, atol=1e-13)
        torch.testing.assert_close(rf, rc, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(fused.States.cpu().numpy(), g[f"{tag}_States_rk4_reffunc"][a * K:(a + 1) * K + 1],
                                   rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(fused.data, cb.data, rtol=1e-11, atol=1e-13)
    fused.close(); cb.close()


def test_lho_vec_env_evolve_tail_interval_and_noise(native, cuda, golden_dir):
    """``evolve`` (actions = action_maximums, base.py:1382-1449) with an episode length that is not a multiple of the
    action interval (tail slice, base.py:455) and in-kernel measurement noise."""
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    from quantrl_b200.envs import LinearizedHOVecEnv
    p, y0 = g["E5_p"], g["E5_y0"]

    class Env(LinearizedHOVecEnv):
        def reset_states(self):
            return torch.as_tensor(y0, device="cuda")

        def get_reward(self):
            return -self.Observations[..., 4]

    env = Env(name="optomech", desc="t", params={}, num_modes=2, num_quads=4, t_norm_max=0.255, t_norm_ssz=0.01,
              t_norm_mul=2.0, n_envs=5, n_observations=20, n_properties=0, n_actions=1, action_maximums=[0.5],
              action_interval=10, data_idxs=[0, -1], cache_all_data=False, dir_prefix="/tmp/qrl_test",
              system="optomech", system_params=p, observation_stds=[1e-3] * 20, seed=3)
    assert env.shape_T == (26,) and env.action_steps == 3 and env.t_ssz == 0.02
    env.evolve(close=False)
    assert env.t_idx == 25 and env.States.shape[0] == 6          # the tail interval has 5 sub-steps
    d = env.data
    np.testing.assert_allclose(d[0, :, 0], np.arange(26) * 0.02, rtol=0, atol=1e-15)
    assert np.all(d[:, 21:, 1] == d[:, 25:26, 1]) and len(env.data_rewards) == 1
    noise = (env.Observations - env.States).flatten()
    assert 0.5e-3 < noise.std().item() < 2e-3
    env.close()


def test_cuda_graph_replay_equals_eager(native, cuda):
    """``use_cuda_graph=True`` (one captured graph per action interval, device-resident time index) reproduces the
    eager env bit for bit over two episodes, including the fused ``data`` cache rows and a tail interval."""
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 1, seed=0)
    E, K = 300, 10
    kw = dict(n_envs=E, n=4, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.455, t_norm_ssz=0.01,
              action_interval=K, observation_stds=[0.01] * 4, seed=9, dir_prefix="/tmp/qrl_test")
    eager, graph = AffineLinearVecEnv(**kw), AffineLinearVecEnv(use_cuda_graph=True, **kw)
    # host-I/O variant: NumPy actions in, NumPy results out; H2D + kernel + D2H replayed as one graph
    hostg = AffineLinearVecEnv(use_cuda_graph=True, return_numpy=True, **kw)
    # ... and the same handing out VIEWS of its (double buffered) pinned result buffers instead of copies
    hostv = AffineLinearVecEnv(use_cuda_graph=True, return_numpy=True, host_output_views=True, fused_coefficients=True, **kw)
    assert eager.shape_T == (46,) and eager.action_steps == 5         # 4 full intervals + a tail of 5 sub-steps
    o_e, o_g, o_h = eager.reset(), graph.reset(), hostg.reset()
    hostv.reset()
    assert torch.equal(o_e, o_g) and np.array_equal(o_e.cpu().numpy(), o_h)
    rng = np.random.default_rng(0)
    prev_view = prev_copy = None
    for step in range(10):
        u = torch.as_tensor(rng.uniform(-1, 1, (E, 1)), device="cuda")
        if step == 3:
            d_e, d_g = eager.data.copy(), graph.data.copy()
            np.testing.assert_array_equal(d_e, d_g)
        oe, re_, de, _ = eager.step(u)
        og, rg, dg, _ = graph.step(u)
        oh, rh, dh, ih = hostg.step(u.cpu().numpy())
        assert torch.equal(oe, og) and torch.equal(re_, rg) and de == dg
        assert isinstance(oh, np.ndarray) and np.array_equal(oe.cpu().numpy(), oh) and np.array_equal(re_.cpu().numpy(), rh)
        assert dh == de and len(ih) == E and hostg.t_idx == eager.t_idx
        ov, rv, dv, _ = hostv.step(u.cpu().numpy())
        # in-kernel affine drift vs the torch hook: one fused multiply-add of rounding (see the test below)
        np.testing.assert_allclose(ov, oh, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(rv, rh, rtol=1e-12, atol=1e-14)
        assert dv == dh and isinstance(ov, np.ndarray)
        if prev_view is not None:     # the results of the previous step are still intact after this one
            np.testing.assert_array_equal(prev_view, prev_copy)
        prev_view, prev_copy = (ov, ov.copy()) if not dv[0] else (None, None)
        assert torch.equal(eager.States, graph.States) and torch.equal(eager.rewards, graph.rewards)
        assert eager.t_idx == graph.t_idx
    assert graph.graph_replays == 8 and graph.graph_native_launches == 1 and eager.batch_idx == graph.batch_idx == 2
    assert hostg.graph_replays == 8 and hostg.batch_idx == 2
    np.testing.assert_array_equal(hostg.data, eager.data)
    eager.close(); graph.close(); hostg.close(); hostv.close()


def test_device_resident_steps_interleaved_with_host_steps(native, cuda):
    """``step_device`` launches of the single-kernel step get their time index from the host and no longer advance the
    device-side one (no last-block protocol); a host round trip (captured graph / re-enqueued block, which READ the
    device-side index) that follows them must see it brought up to date.  Alternating the two kinds of step, over two
    episodes with a tail interval, reproduces an env stepped through ``step(host actions)`` only — bit for bit."""
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 1, seed=0)
    E, K = 200, 10
    kw = dict(n_envs=E, n=4, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.455, t_norm_ssz=0.01,
              action_interval=K, observation_stds=[0.01] * 4, seed=11, dir_prefix="/tmp/qrl_test", use_cuda_graph=True,
              fused_coefficients=True, return_numpy=True)
    ref, mix = AffineLinearVecEnv(**kw), AffineLinearVecEnv(**kw)
    np.testing.assert_array_equal(ref.reset(), mix.reset())
    rng = np.random.default_rng(1)
    pattern = [0, 0, 1, 0, 0, 1, 0, 1, 0, 1]      # 0 = device-resident step, 1 = host round trip (5 steps per episode, the last a tail)
    for step, host in enumerate(pattern):
        u = rng.uniform(-1, 1, (E, 1))
        o_r, r_r, d_r, _ = ref.step(u)
        if host:
            o_m, r_m, d_m, _ = mix.step(u)
            np.testing.assert_array_equal(o_r, o_m)
            np.testing.assert_array_equal(r_r, r_m)
            assert d_r == d_m
        else:
            o_m, r_m, term = mix.step_device(torch.as_tensor(u, device="cuda"))
            if not d_r[0]:
                np.testing.assert_array_equal(o_r, o_m.cpu().numpy())
                np.testing.assert_array_equal(r_r, r_m.cpu().numpy())
            assert term == d_r[0]
            if term:
                mix.reset()
        assert ref.t_idx == mix.t_idx and ref.batch_idx == mix.batch_idx
    np.testing.assert_array_equal(ref.data, mix.data)
    ref.close(); mix.close()


@pytest.mark.parametrize("views", [False, True])
@pytest.mark.parametrize("rng", ["philox", "fed"])
def test_lean_host_step_equals_the_general_step_path(native, cuda, monkeypatch, views, rng):
    """``step(host actions)`` of the single-kernel step takes a lean path (``LinearVecEnv.step``): replay of the captured
    one-kernel graph + inline bookkeeping, or — ``QRL_HOST_STEP_NATIVE`` — ONE native call (``qrl_em_step_host``, ABI 5:
    launch + spin on the completion word the last block stores to pinned memory, or launch + stream synchronisation).
    All three give the results of the general ``step_async`` / ``step_wait`` path (``QRL_NO_FAST_HOST_STEP``) bit for
    bit, over two episodes with a tail interval, with copies and with views of the pinned result buffers."""
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 1, seed=0)
    E, K = 200, 10
    kw = dict(n_envs=E, n=4, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.455, t_norm_ssz=0.01,
              action_interval=K, observation_stds=[0.01] * 4, seed=11, dir_prefix="/tmp/qrl_test", use_cuda_graph=True,
              fused_coefficients=True, return_numpy=True, host_output_views=views, rng=rng)
    fed = None
    if rng == "fed":      # the same caller-supplied draws for every env of the test
        g = np.random.default_rng(5)
        fed = (0.1 * g.standard_normal((45, E, 4)), 0.01 * g.standard_normal((46, E, 4)))
    rng_ = np.random.default_rng(2)
    actions = [rng_.uniform(-1, 1, (E, 1)) for _ in range(10)]
    envs, results = {}, {}
    for name, var, val in (("general", "QRL_NO_FAST_HOST_STEP", "1"), ("lean", None, None),
                           ("native", "QRL_HOST_STEP_NATIVE", "1"), ("native_sync", "QRL_HOST_STEP_NATIVE", "sync")):
        if var:
            monkeypatch.setenv(var, val)
        env = envs[name] = AffineLinearVecEnv(**kw)
        if fed is not None:
            env.set_fed_noise(*fed)
        results[name] = [env.reset().copy()]
        results[name].append([np.array(x) if isinstance(x, np.ndarray) else x for x in env.step(actions[0])])  # fixes the flavour
        if var:
            monkeypatch.delenv(var)
    assert envs["general"]._fast_host is False and envs["lean"]._fast_host == {"native": False}
    assert envs["native"]._fast_host["blocks"][0]["done"] is not None and envs["native_sync"]._fast_host["blocks"][0]["done"] is None
    general = envs["general"]
    for step in range(1, 10):
        res = {}
        for name, env in envs.items():
            n0 = native.launch_count()
            out = env.step(actions[step])
            res[name] = [np.array(x) if isinstance(x, np.ndarray) else x for x in out]
            if name.startswith("native") and env.t_idx != 0 and env.t_idx % K == 0:
                assert native.launch_count() - n0 == 1           # a full interval is one launch of the library
        for name, env in envs.items():
            if name == "general":
                continue
            a, b = res[name], res["general"]
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[1], b[1])
            assert a[2] == b[2] and len(a[3]) == E
            assert env.t_idx == general.t_idx and env.batch_idx == general.batch_idx and env.action_idx == general.action_idx
            assert torch.equal(env.States, general.States) and torch.equal(env.rewards, general.rewards)
    assert general.batch_idx == 2
    assert envs["lean"].graph_replays == 8                        # every full interval, incl. the one that captured the graph
    for name, env in envs.items():
        np.testing.assert_array_equal(env.data, general.data)
        env.close()


def test_lean_device_step_equals_the_general_device_step(native, cuda, monkeypatch):
    """``step_device`` on a full interval of the single-kernel step patches and enqueues the cached argument block itself
    and does the bookkeeping of ``update()`` / ``update_data()`` inline; same States, rewards, cache rows and counters as the
    general path (``QRL_NO_FAST_DEVICE_STEP``), bit for bit, over two episodes with a tail interval."""
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 1, seed=0)
    E, K = 200, 10
    kw = dict(n_envs=E, n=4, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.455, t_norm_ssz=0.01,
              action_interval=K, observation_stds=[0.01] * 4, seed=11, dir_prefix="/tmp/qrl_test", use_cuda_graph=True,
              fused_coefficients=True, return_numpy=True)
    lean = AffineLinearVecEnv(**kw)
    monkeypatch.setenv("QRL_NO_FAST_DEVICE_STEP", "1")
    gen = AffineLinearVecEnv(**kw)
    lean.reset(); gen.reset()
    g = torch.Generator(device="cuda").manual_seed(4)
    for step in range(10):
        u = torch.rand((E, 1), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
        o_g, r_g, t_g = gen.step_device(u)          # (the first call builds the block and decides the flavour)
        if step == 0:
            monkeypatch.delenv("QRL_NO_FAST_DEVICE_STEP")
        o_l, r_l, t_l = lean.step_device(u)
        assert torch.equal(o_l, o_g) and torch.equal(r_l, r_g) and t_l == t_g
        assert torch.equal(lean.States, gen.States) and torch.equal(lean.rewards, gen.rewards)
        assert (lean.t_idx, lean.action_idx, lean.batch_idx) == (gen.t_idx, gen.action_idx, gen.batch_idx)
        assert lean.truncation_pending() == gen.truncation_pending()
        if t_l:
            np.testing.assert_array_equal(lean.data, gen.data)
            assert len(lean.data_rewards) == len(gen.data_rewards)
            lean.reset(); gen.reset()
    assert lean._static["device"]["lean"] is True and gen._static["device"]["lean"] is False and lean.batch_idx == 2
    lean.close(); gen.close()


def test_fused_affine_coefficients_match_hook_path(native, cuda):
    """``fused_coefficients=True`` evaluates A = A0 + sum_a u_a A_a (u = actions * action_maximums) inside the kernel:
    same trajectories as the ``get_A`` torch hook up to the rounding of one fused multiply-add, bit-identical between
    eager and CUDA-graph replay, scaled actions in the cache rows."""
    from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices
    A0, Au = affine_matrices(4, 2, seed=3)
    E, K = 130, 10
    kw = dict(n_envs=E, n=4, A0=A0, Au=Au, noise_prefixes=[0.05, 0.04, 0.03, 0.02], x0=1.0, t_norm_max=0.405,
              t_norm_ssz=0.01, action_interval=K, action_maximums=[0.5, 2.0], observation_stds=[0.01] * 4, seed=2,
              dir_prefix="/tmp/qrl_test")
    hook, fused, fgraph = AffineLinearVecEnv(**kw), AffineLinearVecEnv(fused_coefficients=True, **kw), \
        AffineLinearVecEnv(fused_coefficients=True, use_cuda_graph=True, **kw)
    # device-resident loop: CUDA-tensor actions, the cached single-kernel argument block re-enqueued every step with
    # the host-supplied time index (QRL_EM_HOST_TIDX) and programmatic dependent launch
    fdev = AffineLinearVecEnv(fused_coefficients=True, use_cuda_graph=True, **kw)
    for e in (hook, fused, fgraph, fdev):
        e.reset()
    rng = np.random.default_rng(1)
    for step in range(3):
        u = rng.uniform(-1, 1, (E, 2))
        oh, rh, _, _ = hook.step(u)
        of, rf, _, _ = fused.step(u)
        og, rg, _, _ = fgraph.step(u)
        od, rd, _, _ = fdev.step(torch.as_tensor(u, device="cuda"))
        torch.testing.assert_close(of, oh, rtol=1e-12, atol=1e-14)
        torch.testing.assert_close(rf, rh, rtol=1e-12, atol=1e-14)
        assert torch.equal(of, og) and torch.equal(rf, rg) and torch.equal(fused.States, fgraph.States)
        assert torch.equal(of, od) and torch.equal(rf, rd) and torch.equal(fused.States, fdev.States)
    np.testing.assert_allclose(fused.data, hook.data, rtol=1e-12, atol=1e-14)
    np.testing.assert_array_equal(fused.data, fgraph.data)
    np.testing.assert_array_equal(fused.data, fdev.data)
    np.testing.assert_allclose(fused.data[:, 21:30, 1:3], np.broadcast_to((u * [0.5, 2.0])[:, None, :], (E, 9, 2)), rtol=1e-15)
    # across the episode boundary (4 action steps per episode) and into the second episode
    for step in range(4):
        u = rng.uniform(-1, 1, (E, 2))
        of, rf, df, _ = fused.step(u)
        od, rd, dd, _ = fdev.step(torch.as_tensor(u, device="cuda"))
        assert torch.equal(of, od) and torch.equal(rf, rd) and df == dd and fused.t_idx == fdev.t_idx
    assert fused.batch_idx == fdev.batch_idx == 1
    np.testing.assert_array_equal(fused.data, fdev.data)
    for e in (hook, fused, fgraph, fdev):
        e.close()


def test_lho_vec_env_fused_entanglement_reward(native, cuda, golden_dir):
    """``reward_functor='entan_ln_state'``: reward rows, per-step reward and cumulative reward come out of the
    integration launch; values == the oracle's restatement of the reference's QCMSolver on the env's own States."""
    import measure_oracle as mo
    from quantrl_b200.envs import LinearizedHOVecEnv
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    p, y0 = g["E5_p"], g["E5_y0"]

    class Env(LinearizedHOVecEnv):
        def reset_states(self):
            return torch.as_tensor(y0, device="cuda")

    env = Env(name="optomech", desc="t", params={}, num_modes=2, num_quads=4, t_norm_max=1.205, t_norm_ssz=0.01,
              t_norm_mul=1.0, n_envs=5, n_observations=20, n_properties=0, n_actions=1, action_maximums=[1.0],
              action_interval=10, data_idxs=[0, -1], cache_all_data=False, dir_prefix="/tmp/qrl_test", system="optomech",
              system_params=p, ode_method="dopri5", ode_atol=1e-14, ode_rtol=1e-13, reward_functor="entan_ln_state")
    with pytest.raises(AssertionError, match="reward_functor"):
        Env(name="optomech", desc="t", params={}, num_modes=2, num_quads=4, t_norm_max=1.205, t_norm_ssz=0.01,
            t_norm_mul=1.0, n_envs=5, n_observations=20, n_properties=0, n_actions=1, action_maximums=[1.0],
            action_interval=10, data_idxs=[0, -1], cache_all_data=False, dir_prefix="/tmp/qrl_test", system="optomech",
            system_params=p, reward_functor="no_such_measure")
    env.reset()
    cum = np.zeros(5)
    for a in range(10):
        obs, rew, done, _ = env.step(g["E5_actions"][a])
        ref = mo.entan_ln_2(env.States.cpu().numpy()[:, :, 4:].reshape(11, 5, 4, 4))
        np.testing.assert_allclose(env.Reward.cpu().numpy(), ref, rtol=0, atol=1e-12)
        np.testing.assert_allclose(rew.cpu().numpy(), ref[-1], rtol=0, atol=1e-12)
        np.testing.assert_allclose(rew.cpu().numpy(), g["E5_entan_ln_ref"][(a + 1) * 10], rtol=0, atol=1e-10)
        cum += ref[-1]
    np.testing.assert_allclose(env.rewards.cpu().numpy(), cum, rtol=0, atol=1e-11)
    np.testing.assert_allclose(env.data[:, 100, 1], cum, rtol=0, atol=1e-11)
    env.close()
