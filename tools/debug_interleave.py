"""Development aid: device-resident steps interleaved with host round trips, per-step comparison against an env stepped
through ``step(host actions)`` only (the scenario of tests/test_envs_gpu.py::test_device_resident_steps_interleaved...)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantrl_b200.systems import AffineLinearVecEnv, affine_matrices  # noqa: E402


def run(pattern):
    A0, Au = affine_matrices(4, 1, seed=0)
    E, K = 200, 10
    kw = dict(n_envs=E, n=4, A0=A0, Au=Au, noise_prefixes=0.05, x0=1.0, t_norm_max=0.455, t_norm_ssz=0.01,
              action_interval=K, observation_stds=[0.01] * 4, seed=11, dir_prefix="/tmp/qrl_test", use_cuda_graph=True,
              fused_coefficients=True, return_numpy=True)
    ref, mix = AffineLinearVecEnv(**kw), AffineLinearVecEnv(**kw)
    assert np.array_equal(ref.reset(), mix.reset())
    rng = np.random.default_rng(1)
    out = []
    for step, host in enumerate(pattern):
        u = rng.uniform(-1, 1, (E, 1))
        t0 = ref.t_idx
        o_r, r_r, d_r, _ = ref.step(u)
        if host:
            o_m, r_m, d_m, _ = mix.step(u)
            term = d_m[0]
        else:
            ud = torch.as_tensor(u, device="cuda")
            o_m, r_m, term = mix.step_device(ud)
            o_m, r_m = o_m.cpu().numpy(), r_m.cpu().numpy()
            dyn = mix._dyn.cpu().numpy().tolist()
            if term:
                o_m = mix.reset()
        ok = np.array_equal(o_r, o_m)
        okr = bool(d_r[0]) or np.array_equal(r_r, r_m)
        out.append(f"{step}:{'H' if host else 'D'} t0={t0} obs={'ok' if ok else 'DIFF'} rew={'ok' if okr else 'DIFF'} "
                   f"done={bool(d_r[0])}/{bool(term)} mix.t_idx={mix.t_idx} dyn={mix._dyn.cpu().numpy().tolist()}")
    data_ok = np.array_equal(ref.data, mix.data)
    ref.close(); mix.close()
    return out, data_ok


if __name__ == "__main__":
    pats = {"test": [0, 0, 1, 0, 0, 1, 0, 1, 0, 1], "DDDD": [0] * 10, "HD": [1, 0, 1, 0, 1, 1, 0, 1, 0, 1],
            "DH": [0, 1, 1, 1, 1, 0, 1, 1, 1, 1]}
    for name, p in pats.items():
        lines, data_ok = run(p)
        print(f"== pattern {name} env={ {k: v for k, v in os.environ.items() if k.startswith('QRL_')} } data_equal={data_ok}")
        for ln in lines:
            print("  ", ln)
