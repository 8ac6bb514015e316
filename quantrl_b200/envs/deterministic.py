"""``LinearizedHOVecEnv`` for the ``'b200'`` backend — deterministic linearized harmonic-oscillator environments.

Mirror of the reference class of the same name (quantrl/envs/deterministic.py:439-867): same constructor arguments
(``num_modes``, ``num_quads``, time grid, ``n_envs``, spaces, ``ode_method/ode_atol/ode_rtol`` kwargs), same state
layout ``y = [Re a | Im a | vec(V)]`` (deterministic.py:667), same ``_update_states = solver.step(T_step, States[-1],
actions)`` (deterministic.py:646-651) and the same user hooks.

* With ``system='optomech' | 'kerr_chain' | 'const_AD'`` (+ ``system_params``) the hooks are CUDA device functors and a
  whole action interval is one kernel launch.
* Without ``system`` the reference's Python hooks ``get_mode_rates / get_A / get_D`` (deterministic.py:748-818) are
  used, written in torch on the device, through the solver's stage-callback mode; ``func`` below restates
  deterministic.py:653-746 for that case.
"""
import torch

from .. import _native as N
from ..solvers import SYSTEM_NPARAMS, SYSTEMS, get_IVP_solver
from .base import BaseVecEnv, default_backend


# fused reward functors of qrl_ode_step (csrc/measures.cuh): measure evaluated on the Observations or on the States
_REWARD_KINDS = {"entan_ln_obs": N.REWARD_ENTAN_LN_OBS, "entan_ln_state": N.REWARD_ENTAN_LN_STATE,
                 "sync_c_obs": N.REWARD_SYNC_C_OBS, "sync_c_state": N.REWARD_SYNC_C_STATE,
                 "sync_p_obs": N.REWARD_SYNC_P_OBS, "sync_p_state": N.REWARD_SYNC_P_STATE}


class LinearizedHOVecEnv(BaseVecEnv):
    default_params = {}
    default_ode_solver_params = {"ode_method": "rk4", "ode_atol": 1e-9, "ode_rtol": 1e-6, "ode_substeps": 1}
    backend_libraries = ["b200"]

    def __init__(self, name, desc, params, num_modes, num_quads, t_norm_max, t_norm_ssz, t_norm_mul, n_envs,
                 n_observations, n_properties, n_actions, action_maximums, action_interval, data_idxs,
                 backend_library="b200", backend_precision="double", backend_device="gpu", dir_prefix="data", **kwargs):
        assert backend_library in self.backend_libraries, \
            f"parameter ``backend_library`` should be one of ``{self.backend_libraries}``"
        backend = default_backend(backend_library, backend_precision, backend_device)
        IVPSolver = get_IVP_solver(library=backend_library)
        self.name, self.desc = name, desc
        self.num_modes = int(num_modes)
        self.dim_corrs = (int(num_quads), int(num_quads))
        self.num_corrs = int(num_quads) ** 2
        self.params = {key: params.get(key, val) for key, val in self.default_params.items()}
        for key, val in self.default_ode_solver_params.items():
            kwargs.setdefault(key, val)
        ode = {k: kwargs.pop(k) for k in list(self.default_ode_solver_params)}
        self.reward_functor = kwargs.pop("reward_functor", None)
        assert self.reward_functor in (None,) + tuple(_REWARD_KINDS), \
            f"``reward_functor`` should be None or one of {list(_REWARD_KINDS)}"
        system = kwargs.pop("system", None)
        system_params = kwargs.pop("system_params", None)
        assert n_observations == 2 * self.num_modes + self.num_corrs, \
            "``n_observations`` should equal 2 * num_modes + num_quads**2"
        self.is_A_constant = False
        self.is_D_constant = False
        super().__init__(
            backend=backend, t_norm_max=t_norm_max, t_norm_ssz=t_norm_ssz, t_norm_mul=t_norm_mul, n_envs=n_envs,
            n_observations=n_observations, n_properties=n_properties, n_actions=n_actions,
            action_maximums=action_maximums, action_interval=action_interval, data_idxs=data_idxs,
            dir_prefix=(dir_prefix if dir_prefix != "data" else ("data/" + self.name.lower()) + "/env"),
            file_prefix="lho_vec_env", **kwargs)
        E, q = self.n_envs, self.dim_corrs[0]
        self.A = torch.zeros((E, q, q), dtype=torch.float64, device=self.device)
        self.D = torch.zeros((E, q, q), dtype=torch.float64, device=self.device)

        sysd = None
        if system is not None:
            assert system in SYSTEMS, f"``system`` should be one of {list(SYSTEMS)}"
            sysd = {"system": system, "num_modes": self.num_modes, "num_quads": q, "t_ssz": self.t_ssz}
            if system == "const_AD":
                sysd["A"], sysd["D"] = self.A, self.D   # filled by the user (is_A_constant / is_D_constant semantics)
                self.is_A_constant = self.is_D_constant = True
            else:
                P = self.backend.convert_to_typed(tensor=system_params, dtype="real").contiguous()
                npar = SYSTEM_NPARAMS[system](self.num_modes)
                assert tuple(P.shape) in [(npar,), (E, npar)], f"``system_params`` should have shape ({npar},) or (n_envs, {npar})"
                sysd["params"] = P
        self.system = sysd
        if self.reward_functor is not None:
            assert sysd is not None and q == 4, "fused reward functors need a device functor system with num_quads == 4"
            self._reward_fused = self._cum_fused = True
        # delay systems (quantrl/envs/base.py:214-216, solvers/base.py:96-112, 198-202): the history function is an argument
        # of the user's torch hooks (``args[2]``), refitted by the solver after every action interval — stage-callback mode
        assert not (self.has_delay and sysd is not None), \
            "delay (DDE) systems run in the stage-callback mode (no ``system=``): the device functors have no history argument"
        self._func_delay_0 = getattr(self, "func_delay", None)
        self.solver = IVPSolver(
            func=self.func, y_0=self._States[-1], T=self.T,
            solver_params={"method": ode["ode_method"], "atol": ode["ode_atol"], "rtol": ode["ode_rtol"], "is_stiff": False,
                           "step_interval": self.action_interval, "substeps": ode["ode_substeps"]},
            func_controls=getattr(self, "func_controls", None), has_delay=self.has_delay,
            func_delay=self._func_delay_0, delay_interval=self.action_interval, backend=self.backend,
            system=sysd)

    # ---- user hooks (deterministic.py:748-818) ------------------------------------------------------------------
    def get_A(self, t, modes, args):
        raise NotImplementedError

    def get_D(self, t, modes, args):
        raise NotImplementedError

    def get_mode_rates(self, t, modes, args):
        raise NotImplementedError

    def get_mode_rates_real(self, t, modes_real, args):
        m = self.num_modes
        rates = self.get_mode_rates(t=t, modes=torch.complex(modes_real[:, :m], modes_real[:, m:]), args=args)
        return torch.cat((rates.real, rates.imag), dim=1)

    def func(self, t, y, args):
        """RHS for the stage-callback mode (restates deterministic.py:653-746 in torch)."""
        m, q, E = self.num_modes, self.dim_corrs[0], self.n_envs
        parts, modes = [], None
        if m != 0:
            modes = torch.complex(y[:, :m], y[:, m:2 * m])
            parts.append(self.get_mode_rates_real(t=t, modes_real=y[:, :2 * m], args=args))
        if self.num_corrs != 0:
            V = y[:, 2 * m:].reshape(E, q, q)
            A = self.A if self.is_A_constant else self.get_A(t=t, modes=modes, args=args)
            D = self.D if self.is_D_constant else self.get_D(t=t, modes=modes, args=args)
            parts.append(((torch.matmul(A, V) + torch.matmul(V, A.transpose(1, 2))) + D).reshape(E, q * q))
        return torch.cat(parts, dim=1)

    # ---- engine hooks ---------------------------------------------------------------------------------------
    def _noise_epilogue(self, t_idx):
        """Measurement-noise source for the launch: fed tensor slice or the in-kernel Philox stream."""
        ep = {}
        E, d = self.n_envs, self.n_observations
        if self.observation_stds is None:
            return ep
        if self.rng == "fed":
            ep["obs_noise"] = N.ptr(self.Observation_noises) + 8 * t_idx * E * d
        else:
            ep.update(obs_std=N.ptr(self.observation_stds), obs_std_stride_e=(d if self.observation_stds.dim() == 2 else 0),
                      seed=self._philox_seed, episode=self.batch_idx & 0xFFFFFFFF, t_idx=t_idx, env_offset=self.env_offset)
        return ep

    def _reset(self):
        if self.rng == "fed" and self.observation_stds is not None and not getattr(self, "_fed_assigned", False):
            S1, E, d = self.shape_T[0], self.n_envs, self.n_observations
            self.Observation_noises = self.backend.normal(
                generator=self.backend.generator(seed=self.seed), shape=(S1, E, d), mean=0.0, std=1.0,
                dtype="real") * self.observation_stds.reshape(1, -1, d)
        self.solver.reset_controller()
        if self.has_delay:      # an episode starts from the history the env was built with, not from the last episode's window
            self.solver.func_delay = self._func_delay_0
        return super()._reset()

    def set_fed_noise(self, Observation_noises):
        assert self.rng == "fed"
        self.Observation_noises = self.backend.convert_to_typed(tensor=Observation_noises, dtype="real").contiguous()
        assert tuple(self.Observation_noises.shape) == (self.shape_T[0], self.n_envs, self.n_observations)
        self._fed_assigned = True

    def _initial_observations(self, states_0):
        if self.observation_stds is None:
            self._obs_last.copy_(states_0)
            return states_0.clone()
        if self.system is None:
            noise0 = self.Observation_noises[0] if self.rng == "fed" else \
                self.backend.normal(generator=self.backend.generator(seed=self.seed), shape=tuple(states_0.shape),
                                    mean=0.0, std=1.0, dtype="real") * self.observation_stds
            obs = states_0 + noise0
            self._obs_last.copy_(obs)
            return obs
        # K = 0 launch of the fused kernel: row 0 only
        self.system["T_host"] = self._T_host[:1]
        self.solver.epilogue = dict(self._noise_epilogue(0), obs_last=N.ptr(self._obs_last))
        self.solver.integrate(T_step=self.T[:1], y_0=states_0, params=self.action_maximums_batch, out=self._States[:1])
        return self._obs_last.clone()

    # ---- fast host path of the fused mode (BASELINE config 2 through step()) ------------------------------------------
    # The step used to be an eager chain around the integration kernel: H2D of the actions, a scaling kernel, a freshly
    # built argument block (~40 ctypes stores), the launch, the pack launch, D2H (profiles/r01p: 153 us around a 54 us
    # kernel).  Now the kernel reads the RAW actions in place (pinned host memory over PCIe for step(host array), the
    # caller's tensor for CUDA tensors) and scales them itself (qrl_ode_args.action_scale), and the argument block of a
    # full interval is built once and patched (t0, time index, episode) per step.
    def _host_actions_in_place(self):
        if self.system is None or self.system["system"] == "const_AD" or self.n_actions > 4:
            return False
        if self.__dict__.get("_h_act_dev") is None:
            try:
                self._h_act_dev = N.host_device_pointer(self._h_actions)
            except N.NativeError:
                self._h_act_dev = False
        if not self._h_act_dev:
            return False
        h = self._h_actions     # ``self.actions`` (scaled, on the device) materialises only if something reads it
        self._raw_actions_ptr = self._h_act_dev
        self._actions_lazy = lambda: h.to(self.device, non_blocking=True) * self.action_maximums_batch
        return True

    def _zero_copy_out(self):
        """Device-side views of the pinned result buffer of this step (``return_numpy`` envs with a fused reward functor on
        platforms that map pinned memory), else ``None``: the D2H copy of ``[obs | reward | min,max]`` is then one node."""
        zc = self.__dict__.get("_zc_out")
        if zc is None:
            zc = False
            if getattr(self, "host_zero_copy", True) and self.return_numpy and self._reward_fused and self._obs_last_ptr is None:
                try:
                    E, no = self.n_envs, self.n_observations
                    zc = []
                    for b in self._h_out_bufs:
                        base = N.host_device_pointer(b)
                        zc.append({"obs": base, "reward": base + 8 * E * no, "tail_host": b[E * (no + 1):]})
                    self._step_tail = self._step_out[E * (no + 1):]
                    self._fetch_event = torch.cuda.Event()
                    n_data = 1 + self.n_actions + no + self.n_properties + 1
                    self._pack_reads_actions = self._packed is not None or any(
                        1 <= (i + n_data if i < 0 else i) <= self.n_actions for i in self.data_idxs)
                    self._host_event_after_pack = False
                except N.NativeError:
                    zc = False    # pinned memory not mapped on this platform: keep the D2H copy
            self._zc_out = zc
        return zc[self._h_sel] if zc else None

    def update_data(self):
        super().update_data()
        if self.__dict__.get("_host_event_after_pack"):
            self._host_event_after_pack = False
            self._fetch_event.record()
            self._host_event = self._fetch_event

    def _set_actions(self, a):
        if self.system is None or self.system["system"] == "const_AD" or self.n_actions > 4:
            return super()._set_actions(a)
        self._host_step = False
        a = a if a.is_cuda else a.to(self.device, non_blocking=True)
        self._raw_actions = a.reshape(self.n_envs, self.n_actions).to(torch.float64).contiguous()   # no copy when already so
        self._raw_actions_ptr = self._raw_actions.data_ptr()
        self._actions_lazy = lambda: self._raw_actions * self.action_maximums_batch

    def _fused_args(self, dim):
        """Cached ``qrl_ode_args`` of a full interval (``dim == K + 1``): everything that does not change from step to step."""
        a = self.__dict__.get("_ode_args")
        if a is not None:
            return a
        sysd, sp = self.system, self.solver.solver_params
        a = N.OdeArgs()
        a.system, a.num_modes, a.num_quads = SYSTEMS[sysd["system"]], sysd["num_modes"], sysd["num_quads"]
        a.n_envs, a.K = self.n_envs, dim - 1
        a.method, adaptive = N.ODE_METHODS[sp["method"]]
        a.substeps, a.t_ssz, a.atol, a.rtol = int(sp["substeps"]), self.t_ssz, float(sp["atol"]), float(sp["rtol"])
        P = sysd["params"]
        a.params, a.params_stride_e = N.ptr(P), (P.shape[-1] if P.dim() == 2 else 0)
        a.n_actions, a.action_scale = self.n_actions, N.ptr(self.action_maximums)
        a.y0, a.y_last, a.states = N.ptr(self._x_last), N.ptr(self._x_last), N.ptr(self._States)
        a.observations, a.obs_last = N.ptr(self._Observations), N.ptr(self._obs_last)
        self._obs_last_dev_ptr, self._reward_last_dev_ptr = N.ptr(self._obs_last), None
        a.obs_minmax, a.nan_flag = N.ptr(self._minmax), N.ptr(self._nan)
        if self.reward_functor is not None:
            a.reward_kind, a.reward = _REWARD_KINDS[self.reward_functor], N.ptr(self._Reward)
            a.reward_last, a.rewards_cum = N.ptr(self._reward_last), N.ptr(self.rewards)
            self._reward_last_dev_ptr = N.ptr(self._reward_last)
        if self.observation_stds is not None and self.rng != "fed":
            a.obs_std, a.obs_std_stride_e = N.ptr(self.observation_stds), (self.n_observations if self.observation_stds.dim() == 2 else 0)
            a.seed, a.env_offset = self._philox_seed, self.env_offset
        if adaptive:
            if self.solver.h_state is None or self.solver.h_state.shape[0] != self.n_envs:
                self.solver.h_state = torch.zeros((self.n_envs,), dtype=torch.float64, device=self.device)
                self.solver.n_steps = torch.zeros((self.n_envs, 2), dtype=torch.int32, device=self.device)
            a.h_state, a.n_steps = N.ptr(self.solver.h_state), N.ptr(self.solver.n_steps)
        self._ode_args = a
        return a

    def _update_states(self):
        dim = self._dim
        t0 = self.t_idx
        fast = (self.system is not None and self.system["system"] != "const_AD" and dim == self.action_interval + 1
                and self.n_actions <= 4 and (self._host_step or self.__dict__.get("_actions_lazy") is not None))
        if fast:
            a = self._fused_args(dim)
            a.actions = self._raw_actions_ptr      # pinned host memory (step(host array)) or the caller's CUDA tensor
            # (no per-step reset of the min/max accumulators: they run over the episode — reset() re-arms them — and
            # "max > hi or min < lo so far" is the reference's per-block check (base.py:485-499) at every step, because the
            # first block that leaves the range ends the episode)
            a.t0, a.t_idx, a.episode = float(self._T_host[t0]), t0, self.batch_idx & 0xFFFFFFFF
            if self.observation_stds is not None and self.rng == "fed":
                a.obs_noise = N.ptr(self.Observation_noises) + 8 * t0 * self.n_envs * self.n_observations
            zc = self._zero_copy_out() if self._host_step else None
            if zc is not None:
                # step(host actions) -> host results: the kernel writes the last observations and rewards straight into the
                # pinned result buffer (zero-copy over PCIe); only [min, max | status] (32 bytes) is copied, an event marks
                # the point the host waits for, and the cache-row launch that follows runs behind the host's back
                a.obs_last, a.reward_last = zc["obs"], zc["reward"]
                N.ode_step(a)
                zc["tail_host"].copy_(self._step_tail, non_blocking=True)
                if self._pack_reads_actions:      # the cache-row launch reads the pinned actions: the host waits for it too
                    self._host_event_after_pack = True
                else:
                    self._fetch_event.record()
                    self._host_event = self._fetch_event
            else:
                a.obs_last, a.reward_last = self._obs_last_dev_ptr, self._reward_last_dev_ptr
                N.ode_step(a)
            return self._States
        if self.system is not None:
            _ = self.actions                # (materialise the scaled actions for the generic launch below)
            self.system["T_host"] = self._T_step_host
            self.solver.epilogue = dict(
                self._noise_epilogue(t0), observations=N.ptr(self._Observations), y_last=N.ptr(self._x_last),
                obs_last=N.ptr(self._obs_last), obs_minmax=N.ptr(self._minmax), nan_flag=N.ptr(self._nan))
            if self.reward_functor is not None:
                self.solver.epilogue.update(
                    reward_kind=_REWARD_KINDS[self.reward_functor],
                    reward=N.ptr(self._Reward), reward_last=N.ptr(self._reward_last), rewards_cum=N.ptr(self.rewards))
            return self.solver.integrate(T_step=self.T_step, y_0=self._x_last, params=self.actions,
                                         out=self._States[:dim])
        # stage-callback mode: integrate with torch hooks, then the (unfused) noise add / reductions
        Y = self.solver.step(T_step=self.T_step, y_0=self._x_last.clone(), params=self.actions)
        self._States[:dim] = Y
        obs = Y
        if self.observation_stds is not None:
            noise = self.Observation_noises[t0:t0 + dim] if self.rng == "fed" else \
                self.backend.normal(generator=self.backend.generator(seed=self.seed), shape=tuple(Y.shape), mean=0.0,
                                    std=1.0, dtype="real") * self.observation_stds
            obs = Y + noise
        self._Observations[:dim] = obs
        self._x_last.copy_(Y[dim - 1])
        self._obs_last.copy_(obs[dim - 1])
        finite = obs[torch.isfinite(obs)]
        if finite.numel() > 0:
            self._minmax[0], self._minmax[1] = torch.minimum(self._minmax[0], finite.min()), torch.maximum(self._minmax[1], finite.max())
        return self._States[:dim]
