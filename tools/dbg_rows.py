import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import systems as sy
from helpers import ode_step
from quantrl_b200 import _native as N
rng = np.random.default_rng(12)
E, m, q = 203, 4, 8
p = sy.kerr_default_params(E, m, seed=5); y0 = sy.kerr_y0(p, m)
kw = dict(params=p, actions=rng.uniform(-1, 1, E))
for noise in (False, True):
    nk = dict(philox=dict(seed=3, episode=2, t_idx=40, env_offset=11), obs_std=np.full(72, 1e-3)) if noise else {}
    rows = ode_step("kerr_chain", m, q, y0, 12, 0.4, 0.01, substeps=2, **nk, **kw)
    thr = ode_step("kerr_chain", m, q, y0, 12, 0.4, 0.01, substeps=2, flags=N.ODE_FORCE_THREAD, **nk, **kw)
    for key in ("states", "observations", "y_last", "obs_last", "minmax"):
        d = (rows[key] - thr[key]).abs()
        print(noise, key, float(d.max()), int((d > 0).sum()), tuple(d.shape))
    d = (rows["states"] - thr["states"]).abs()
    idx = (d > 0).nonzero()
    print(idx[:10].tolist())
