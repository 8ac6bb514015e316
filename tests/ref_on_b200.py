"""Shared by the CPU and GPU tests: the reference's OWN ``LinearizedHOVecEnv`` class (unmodified, imported through
oracle/ref_shim.py from /root/reference or from the staged copy oracle/_ref) running on ``backend_library='b200'`` after
``quantrl_b200.register()`` — the "fourth backend" of BASELINE.json's north_star.  The hooks are the canonical
optomechanical system written in torch."""
import numpy as np
import torch


def build_reference_env(quantrl, p, y0, K, S, ssz, device):
    LHO = quantrl.envs.deterministic.LinearizedHOVecEnv
    P = torch.as_tensor(p, device=device)
    kap, gam, om, D0, g0, nth, Al = (P[:, i] for i in range(7))
    E = p.shape[0]

    class RefOptomechEnv(LHO):
        backend_libraries = LHO.backend_libraries + ["b200"]       # the class-level allow-list (deterministic.py:515)
        default_ode_solver_params = dict(LHO.default_ode_solver_params, ode_method="rk4")

        def reset_states(self):
            return torch.as_tensor(y0, device=device)

        def get_reward(self):
            O = self.Observations
            return -(O[..., 4] + O[..., 9] + O[..., 14] + O[..., 19])

        def get_mode_rates(self, t, modes, args):
            a, b = modes[:, 0], modes[:, 1]
            u = args[0][:, 0]
            da = -(0.5 * kap - 1j * D0) * a + 1j * g0 * a * (b + b.conj()) + Al * (1.0 + u)
            db = -(0.5 * gam + 1j * om) * b + 1j * g0 * (a * a.conj())
            return torch.stack((da, db), dim=1)

        def get_A(self, t, modes, args):
            a, b = modes[:, 0], modes[:, 1]
            Delta = D0 + 2.0 * g0 * b.real
            GR, GI = 2.0 * g0 * a.real, 2.0 * g0 * a.imag
            A = torch.zeros((E, 4, 4), dtype=torch.float64, device=device)
            A[:, 0, 0] = A[:, 1, 1] = -0.5 * kap
            A[:, 0, 1], A[:, 1, 0] = -Delta, Delta
            A[:, 0, 2], A[:, 1, 2] = -GI, GR
            A[:, 2, 2] = A[:, 3, 3] = -0.5 * gam
            A[:, 2, 3], A[:, 3, 2] = om, -om
            A[:, 3, 0], A[:, 3, 1] = GR, GI
            return A

        def get_D(self, t, modes, args):
            D = torch.zeros((E, 4, 4), dtype=torch.float64, device=device)
            D[:, 0, 0] = D[:, 1, 1] = 0.5 * kap
            D[:, 2, 2] = D[:, 3, 3] = gam * (nth + 0.5)
            return D

    return RefOptomechEnv(name="optomech", desc="reference env class on the b200 backend", params={}, num_modes=2, num_quads=4,
                          t_norm_max=S * ssz + ssz / 2, t_norm_ssz=ssz, t_norm_mul=1.0, n_envs=E, n_observations=20,
                          n_properties=0, n_actions=1, action_maximums=[1.0], action_interval=K, data_idxs=[0, 1, 2, 22],
                          backend_library="b200", dir_prefix="/tmp/qrl_ref_b200", plot=False, cache_all_data=False)


def run_and_check(quantrl, golden_dir, device, n_steps=2):
    """Drive the reference env class for ``n_steps`` action steps on the b200 backend and hold it to fixture G2 (classic
    RK4 on the reference's own ``func`` with the NumPy backend) — same engine semantics, other backend."""
    import os
    g = np.load(os.path.join(golden_dir, "lyap_optomech.npz"))
    p, y0, K, ssz = g["E5_p"], g["E5_y0"], int(g["E5_K"]), float(g["E5_ssz"])
    S = int(g["E5_S"])
    ref = g["E5_States_rk4_reffunc"]                   # (S+1, E, d)
    env = build_reference_env(quantrl, p, y0, K, S, ssz, device)
    assert type(env.backend).__name__ == "B200Backend" and type(env.solver).__name__ == "B200IVPSolver"
    obs0 = env.reset()
    np.testing.assert_allclose(env.backend.convert_to_numpy(obs0), y0, rtol=0, atol=0)
    scale = np.maximum(np.abs(ref).max(axis=(0, 1)), 1.0)
    for a in range(n_steps):
        env.step_async(g["E5_actions"][a])               # (``step`` itself lives in SB3's VecEnv, which is a stand-in here)
        obs, rew, dones, infos = env.step_wait()
        Y = env.backend.convert_to_numpy(env.States)
        assert Y.shape == (K + 1, p.shape[0], 20)
        err = np.max(np.abs(Y - ref[a * K:(a + 1) * K + 1]) / scale)
        assert err < 1e-10, (a, err)
        np.testing.assert_allclose(env.backend.convert_to_numpy(obs), ref[(a + 1) * K], rtol=1e-9, atol=1e-9)
        assert dones == [False] * p.shape[0] and env.t_idx == (a + 1) * K
    assert env.data.shape == (p.shape[0], S + 1, 4)
    np.testing.assert_allclose(env.data[:, :n_steps * K + 1, 0], np.broadcast_to(np.arange(n_steps * K + 1) * ssz, (p.shape[0], n_steps * K + 1)), rtol=1e-12)
    return env
