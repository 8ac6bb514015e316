"""``B200IVPSolver`` — the ``'b200'`` entry of the reference's IVP-solver registry.

Same constructor, attributes and ``step / integrate / interpolate / solve_ivp`` surface as
``quantrl.solvers.base.BaseIVPSolver`` (quantrl/solvers/base.py:21-256): ``integrate(T_step, y_0, params)`` returns
``Y`` with ``Y.shape == (len(T_step), *y_0.shape)`` and ``Y[0] == y_0``; ``args = [params, func_controls, func_delay]``
reaches ``func`` exactly as in quantrl/solvers/numpy.py:108.

Two execution modes:

* **fused** (the product path): the env that owns ``func`` declares a device functor (``system=...``); the whole
  interval — all grid points, all RK stages, the RHS and its hooks — is ONE launch of ``qrl_ode_step``
  (``'rk4'`` fixed step, ``'dopri5'`` per-env adaptive with dense output, and the generic explicit Runge–Kutta engine
  for the other explicit method names of the reference's plugins: ``'bosh3'/'RK23'``, ``'tsit5'``,
  ``'adaptive_heun'/'heun'``, ``'fehlberg2'``, ``'dop853'/'DOP853'/'dopri8'``, ``'euler'``, ``'midpoint'``).
* **stage-callback**: no functor, arbitrary user hooks written in torch; classic RK4 whose four stage evaluations
  call ``func`` (batched torch on the device).  This keeps every reference env runnable on the backend — including
  delay systems (``has_delay``: ``step`` refits ``func_delay`` to the rows of the last window, as
  quantrl/solvers/base.py:198-202 does) — it is not the accelerated path.
"""
import numpy as np
import torch

from . import _native as N

SYSTEMS = {"optomech": N.SYS_OPTOMECH, "kerr_chain": N.SYS_KERR_CHAIN, "const_AD": N.SYS_CONST_AD}
SYSTEM_NPARAMS = {"optomech": lambda m: 7, "kerr_chain": lambda m: 5 + m, "const_AD": lambda m: 0}


class B200IVPSolver:
    default_solver_params = {"method": "rk4", "atol": 1e-12, "rtol": 1e-9, "is_stiff": False, "step_interval": 10,
                             "complex": False, "substeps": 1}
    solver_methods = list(N.ODE_METHODS)   # fused mode; the stage-callback mode implements "rk4" only

    def __init__(self, func, y_0, T, solver_params: dict, func_controls=None, has_delay: bool = False, func_delay=None,
                 delay_interval: int = 0, backend=None, system=None):
        self.func, self.y_0, self.T = func, y_0, T
        self.func_controls, self.has_delay, self.func_delay, self.delay_interval = func_controls, has_delay, func_delay, delay_interval
        self.backend = backend
        # same validation as quantrl/solvers/base.py:96-112
        assert self.func_delay is not None if self.has_delay else True, \
            "delay function cannot be ``None`` if parameter ``has_delay`` is ``True``"
        assert not (has_delay and system is not None), \
            "delay (DDE) systems run in the stage-callback mode: the device functors have no history argument"
        self.shape_y = tuple(y_0.shape)
        self.shape_T = tuple(T.shape)
        self.solver_params = {k: solver_params.get(k, v) for k, v in self.default_solver_params.items()}
        if self.has_delay and self.delay_interval != 0:   # the history is refitted once per step: step == delay window
            self.solver_params["step_interval"] = self.delay_interval
        assert self.solver_params["method"] in self.solver_methods, \
            f"parameter ``method`` should be one of ``{self.solver_methods}``"
        assert isinstance(self.solver_params["step_interval"], int) and self.solver_params["step_interval"] < self.shape_T[0], \
            "parameter ``step_interval`` should be an integer with a value less than the total number of steps"
        self.step_interval = self.solver_params["step_interval"]
        # fused mode: a dict(system=, num_modes=, num_quads=, params=(E,np) tensor, A=, D=) supplied by the env
        self.system = system
        self.epilogue = {}          # optional output pointers the env wants filled by the same launch
        self.h_state = None
        self.n_steps = None

    # ---- reference surface ---------------------------------------------------------------------------------
    def step(self, T_step, y_0, params):
        """Integrate, then refit the delay function to this step's rows (quantrl/solvers/base.py:168-204)."""
        Y = self.integrate(T_step=T_step, y_0=y_0, params=params)
        if self.has_delay:
            self.func_delay = self.interpolate(T_step=T_step, Y=Y)
        return Y

    def interpolate(self, T_step, Y):
        """History function of a delay system: ``t -> y(t)`` on (and, like ``splev``, extrapolated beyond) the window
        ``T_step``.  Same interpolant as the reference's ``splrep``/``splev`` per component (quantrl/solvers/numpy.py:180-188:
        the cubic spline through the ``len(T_step)`` rows with not-a-knot ends), but evaluated for the whole batch at
        once: a spline through the rows is linear in them, ``y(t) = sum_i w_i(t) Y[i]``, so the weights ``w(t)`` come
        from one B-spline basis on the host and the combination is ONE ``tensordot`` on the device.  Returns a callable
        whose result has the shape of ``y_0`` (the reference's returns a list over components)."""
        from scipy.interpolate import make_interp_spline
        Th = T_step.detach().cpu().numpy() if isinstance(T_step, torch.Tensor) else np.asarray(T_step, dtype=np.float64)
        dim = Th.shape[0]
        assert dim >= 2 and Y.shape[0] == dim, "``interpolate`` needs at least two rows, one per element of ``T_step``"
        basis = make_interp_spline(Th, np.eye(dim), k=min(3, dim - 1))
        rows = Y.detach().clone()       # the env reuses its (K+1, E, d) block for the next interval

        def func_delay(t):
            w = torch.as_tensor(np.ascontiguousarray(basis(float(t))), dtype=rows.dtype).to(rows.device)
            return torch.tensordot(w, rows, dims=1)
        return func_delay

    def solve_ivp(self, y_0, params, show_progress=False):
        """quantrl/solvers/base.py:206-256."""
        Y = torch.empty((self.shape_T[0], *y_0.shape), dtype=torch.float64, device=y_0.device)
        Y[0] = y_0
        K = self.step_interval
        for i in range(K, self.shape_T[0], K):
            Y[i - K + 1:i + 1] = self.step(T_step=self.T[i - K:i + 1], y_0=Y[i - K], params=params)[1:]
        return Y

    def integrate(self, T_step, y_0, params=None, out=None):
        if self.system is not None:
            return self._integrate_fused(T_step, y_0, params, out)
        return self._integrate_callback(T_step, y_0, params)

    # ---- fused: one launch of qrl_ode_step ----------------------------------------------------------------------
    def _integrate_fused(self, T_step, y_0, params, out=None):
        sysd = self.system
        m, q = sysd["num_modes"], sysd["num_quads"]
        y_0 = y_0.contiguous()
        E, d = y_0.shape
        assert d == 2 * m + q * q, "y_0 should have 2 * num_modes + num_quads**2 columns"
        dim = int(T_step.shape[0])
        K = dim - 1
        Y = out if out is not None else torch.empty((dim, E, d), dtype=torch.float64, device=y_0.device)
        a = N.OdeArgs()
        a.system, a.num_modes, a.num_quads, a.n_envs, a.K = SYSTEMS[sysd["system"]], m, q, E, K
        a.method, adaptive = N.ODE_METHODS[self.solver_params["method"]]
        a.substeps = int(self.solver_params["substeps"])
        t_host = sysd.get("T_host")
        if t_host is not None:
            a.t0, t1 = float(t_host[0]), float(t_host[1] if K > 0 else t_host[0])
        else:
            a.t0 = float(T_step[0])
            t1 = float(T_step[1]) if K > 0 else a.t0
        a.t_ssz = sysd.get("t_ssz", (t1 - a.t0) if K > 0 else 1.0)
        a.atol, a.rtol = float(self.solver_params["atol"]), float(self.solver_params["rtol"])
        keep = []
        if sysd["system"] == "const_AD":
            A, D = sysd["A"].contiguous(), sysd["D"].contiguous()
            keep += [A, D]
            a.A, a.A_stride_e = N.ptr(A), (q * q if A.dim() == 3 else 0)
            a.D, a.D_stride_e = N.ptr(D), (q * q if D.dim() == 3 else 0)
        else:
            P = sysd["params"]
            a.params, a.params_stride_e = N.ptr(P), (P.shape[-1] if P.dim() == 2 else 0)
            act = params.contiguous()
            keep.append(act)
            a.actions, a.n_actions = N.ptr(act), act.shape[-1]
        a.y0 = N.ptr(y_0)
        a.states = N.ptr(Y)
        ep = self.epilogue
        a.observations, a.y_last, a.obs_last = ep.get("observations"), ep.get("y_last"), ep.get("obs_last")
        a.obs_noise, a.obs_minmax, a.nan_flag = ep.get("obs_noise"), ep.get("obs_minmax"), ep.get("nan_flag")
        if ep.get("reward_kind"):
            a.reward_kind, a.reward = ep["reward_kind"], ep.get("reward")
            a.reward_last, a.rewards_cum = ep.get("reward_last"), ep.get("rewards_cum")
        if ep.get("obs_std") is not None:
            a.obs_std, a.obs_std_stride_e = ep["obs_std"], ep["obs_std_stride_e"]
            a.seed, a.episode, a.t_idx, a.env_offset = ep["seed"], ep["episode"], ep["t_idx"], ep["env_offset"]
        if adaptive:
            if self.h_state is None or self.h_state.shape[0] != E:
                self.h_state = torch.zeros((E,), dtype=torch.float64, device=y_0.device)
                self.n_steps = torch.zeros((E, 2), dtype=torch.int32, device=y_0.device)
            a.h_state, a.n_steps = N.ptr(self.h_state), N.ptr(self.n_steps)
        N.ode_step(a)
        return Y

    def reset_controller(self):
        """Forget the carried dopri5 step sizes (called at episode reset)."""
        if self.h_state is not None:
            self.h_state.zero_()
            self.n_steps.zero_()

    # ---- stage-callback: arbitrary torch hooks, classic RK4 ----------------------------------------------------
    def _integrate_callback(self, T_step, y_0, params):
        assert self.solver_params["method"] == "rk4", "stage-callback mode (no device functor) supports ``'rk4'`` only"
        args = [params, self.func_controls, self.func_delay]
        dim = int(T_step.shape[0])
        Y = torch.empty((dim, *y_0.shape), dtype=y_0.dtype, device=y_0.device)
        Y[0] = y_0
        y = y_0.clone()
        nsub = int(self.solver_params["substeps"])
        Th = T_step.detach().cpu().numpy() if isinstance(T_step, torch.Tensor) else np.asarray(T_step)
        f = lambda t, v: self.func(t, v, args).clone()  # noqa: E731  (func returns a reused buffer, deterministic.py:739-746)
        for i in range(dim - 1):
            h = (Th[i + 1] - Th[i]) / nsub
            for s in range(nsub):
                t = Th[i] + s * h
                k1 = f(t, y)
                k2 = f(t + 0.5 * h, y + (0.5 * h) * k1)
                k3 = f(t + 0.5 * h, y + (0.5 * h) * k2)
                k4 = f(t + h, y + h * k3)
                y = y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
            Y[i + 1] = y
        return Y


_SOLVERS = {"b200": B200IVPSolver}


def get_IVP_solver(library: str = "b200"):
    """Same contract as quantrl/solvers/context_manager.py:16-46: returns the solver CLASS."""
    assert "b200" in library.lower(), "quantrl_b200 only provides the ``'b200'`` IVP solver"
    return _SOLVERS["b200"]
