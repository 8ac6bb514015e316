"""``LinearVecEnv`` — vectorized stochastic linear environments (Euler–Maruyama with Wiener increments) on a B200.

The reference's stochastic env is single-env only (``LinearEnv(BaseGymEnv)``, quantrl/envs/stochastic.py:21-284).
This class keeps its per-env semantics and user hooks, batched over ``n_envs`` independent environments behind the
``BaseSB3Env``-style vectorized surface:

* ``get_A(t_idx, args) -> (n_envs, n, n)`` or ``(n, n)``      (stochastic.py:244-263; ``args = (actions, None, None)``)
* ``get_noise_prefixes(t_idx, args) -> (n_envs, n)`` or ``(n,)`` (stochastic.py:265-284)
* ``reset_states() -> (n_envs, n)``, ``get_reward() -> (K+1, n_envs)`` (optional when a fused ``reward_functor`` is set)

Per sub-step (stochastic.py:218-242):  ``Y[i+1] = (I + A(t_idx+i) * t_ssz) @ Y[i] + n(t_idx+i) * Ws[t_idx+i]``,
``Ws = sqrt(t_ssz) * N(0,1)`` (stochastic.py:162-170); observations add measurement noise
``N(0,1) * observation_stds`` (envs/base.py:408-426, 462).

The hooks are batched torch on the device and — because they depend only on ``(t_idx, actions)`` — are evaluated
once per action interval (``is_A_time_dependent=False``) or once per sub-step of the interval and stacked
(``True``); all ``K`` sub-steps then run in ONE launch of ``qrl_em_step``.

``rng='philox'``: increments and measurement noise are generated inside the kernel (no ``O(shape_T*E*n)`` tensors).
``rng='fed'``: ``Ws`` ``(shape_T-1, E, n)`` and ``Observation_noises`` ``(shape_T, E, n)`` are drawn at ``reset`` exactly
like the reference (or assigned by the caller) and fed to the kernel — used for pathwise parity with the reference.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _native as N
from .base import BaseVecEnv, default_backend

_REWARD_KINDS = {None: N.REWARD_NONE, "neg_wsq_obs": N.REWARD_NEG_WSQ_OBS, "neg_wsq_state": N.REWARD_NEG_WSQ_STATE}


class LinearVecEnv(BaseVecEnv):
    default_params = {}
    backend_libraries = ["b200"]

    def __init__(self, name, desc, params, t_norm_max, t_norm_ssz, t_norm_mul, n_envs, n_observations, n_properties,
                 n_actions, action_maximums, action_interval, data_idxs, backend_library="b200",
                 backend_precision="double", backend_device="gpu", dir_prefix="data", **kwargs):
        assert backend_library in self.backend_libraries, \
            f"parameter ``backend_library`` should be one of ``{self.backend_libraries}``"
        backend = default_backend(backend_library, backend_precision, backend_device)
        self.name, self.desc = name, desc
        self.params = {key: params.get(key, val) for key, val in self.default_params.items()}
        self.is_A_time_dependent = bool(kwargs.pop("is_A_time_dependent", False))
        self.fuse_cache_rows = bool(kwargs.pop("fuse_cache_rows", True))
        self.force_generic_kernel = bool(kwargs.pop("force_generic_kernel", False))
        # replay one captured CUDA graph per action interval (hooks + integration launch) instead of launching eagerly;
        # requires hooks that do not depend on ``t_idx`` (``is_A_time_dependent=False``)
        self.use_cuda_graph = bool(kwargs.pop("use_cuda_graph", False))
        # evaluate an action-affine drift inside the kernel (``_affine_drift``) instead of through the ``get_A`` hook
        self.fused_coefficients = bool(kwargs.pop("fused_coefficients", False))
        # host round trip of ``step(host actions)`` under ``use_cuda_graph`` + ``return_numpy``: the kernel reads the
        # actions from, and writes obs_last / reward_last / [min,max] to, pinned host memory directly (no copy nodes)
        self.host_zero_copy = bool(kwargs.pop("host_zero_copy", True))
        self._zc = None
        self._peer_barrier = None           # (flag pointer table, epoch, rank, world): ABI-3 fused barrier (kept for A/B)
        self._peer = None                   # distributed.PeerGather: publish-only exchange, outputs in a local ring
        self.last_step_peer_fenced = False
        self.reward_functor = kwargs.pop("reward_functor", None)
        reward_weights = kwargs.pop("reward_weights", None)
        assert self.reward_functor in _REWARD_KINDS, f"``reward_functor`` should be one of {list(_REWARD_KINDS)}"
        self.Ws = None
        super().__init__(
            backend=backend, t_norm_max=t_norm_max, t_norm_ssz=t_norm_ssz, t_norm_mul=t_norm_mul, n_envs=n_envs,
            n_observations=n_observations, n_properties=n_properties, n_actions=n_actions,
            action_maximums=action_maximums, action_interval=action_interval, data_idxs=data_idxs,
            dir_prefix=(dir_prefix if dir_prefix != "data" else ("data/" + self.name.lower()) + "/env") + "_" + "_".join(
                [str(val) for _, val in self.params.items()]),
            file_prefix="lin_vec_env", **kwargs)
        n = self.n_observations
        assert 1 <= n <= 16, "qrl_em_step supports 1 <= n_observations <= 16"
        assert not self.has_delay, "the Euler-Maruyama path has no history argument (the reference's LinearEnv has none either)"
        # (shape_T - 1) % action_interval != 0 is supported: the last interval is shorter (dim < K+1, base.py:454-458);
        # the reference's own LinearEnv raises there (its States block keeps K+1 rows, base.py:462)
        self.I = torch.eye(n, dtype=torch.float64, device=self.device)
        self._reward_fused = self.reward_functor is not None
        self._cum_fused = self._reward_fused
        assert not (self.use_cuda_graph and self.is_A_time_dependent), \
            "``use_cuda_graph`` needs coefficient hooks that do not depend on the time index"
        self._dyn = torch.zeros((2,), dtype=torch.int64, device=self.device)      # {t_idx, episode} read by the kernel
        self._actions_in = torch.zeros((self.n_envs, self.n_actions), dtype=torch.float64, device=self.device)
        self._graph = self._graph_host = self._graph_host1 = None
        self._static = {}                   # cached argument blocks of the single-kernel step (_relaunch_interval)
        self._actions_src = None
        self.graph_replays, self.graph_native_launches = 0, 0
        # last-block bookkeeping of the captured step (qrl_em_args.finish_counter / minmax_out / dyn_advance)
        self._finish = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._minmax_acc = torch.tensor([float("inf"), float("-inf")], dtype=torch.float64, device=self.device)
        self._reward_w = None
        if self._reward_fused:
            w = np.ones(n) if reward_weights is None else np.asarray(reward_weights, dtype=np.float64)
            assert w.shape == (n,), "``reward_weights`` should have shape (n_observations,)"
            self._reward_w = torch.as_tensor(w, device=self.device)

    # ---- user hooks ---------------------------------------------------------------------------------------
    def get_A(self, t_idx, args):
        raise NotImplementedError

    def get_noise_prefixes(self, t_idx, args):
        raise NotImplementedError

    def get_reward(self):
        raise NotImplementedError("override ``get_reward`` or pass ``reward_functor=``")

    # ---- reset: noise bookkeeping (stochastic.py:157-176, base.py:408-426) --------------------------------------
    def _reset(self):
        if self.rng == "fed" and not getattr(self, "_fed_assigned", False):
            S1, E, n = self.shape_T[0], self.n_envs, self.n_observations
            Ws = np.sqrt(self.t_ssz) * self.backend.normal(
                generator=self.backend.generator(seed=self.seed), shape=(S1 - 1, E, n), mean=0.0, std=1.0, dtype="real")
            # (drawn into the SAME tensors every episode: captured graphs and cached argument blocks hold their addresses)
            self.Ws = Ws if self.Ws is None or self.Ws.shape != Ws.shape else self.Ws.copy_(Ws)
            if self.observation_stds is not None:
                noises = self.backend.normal(
                    generator=self.backend.generator(seed=self.seed), shape=(S1, E, n), mean=0.0, std=1.0,
                    dtype="real") * self.observation_stds.reshape(1, -1, n)
                old = self.Observation_noises
                self.Observation_noises = noises if old is None or old.shape != noises.shape else old.copy_(noises)
        if self.use_cuda_graph:
            # scalar fills: the values travel as kernel arguments, so a host that runs whole episodes ahead of the GPU
            # (device-resident loops) cannot overwrite a staging buffer before an earlier copy has executed
            self._dyn[0].fill_(0)
            self._dyn[1].fill_((self.batch_idx + 1) & 0xFFFFFFFF)
            self._dyn_stale = False
        return super()._reset()

    def set_fed_noise(self, Ws, Observation_noises=None):
        """Use caller-supplied draws (e.g. the reference's own) for the next episode(s): ``Ws (shape_T-1, E, n)``
        already scaled by ``sqrt(t_ssz)``; ``Observation_noises (shape_T, E, n)`` already scaled by the stds."""
        assert self.rng == "fed"
        S1, E, n = self.shape_T[0], self.n_envs, self.n_observations
        self.Ws = self.backend.convert_to_typed(tensor=Ws, dtype="real").contiguous()
        assert tuple(self.Ws.shape) == (S1 - 1, E, n)
        self.Observation_noises = None
        if Observation_noises is not None:
            self.Observation_noises = self.backend.convert_to_typed(tensor=Observation_noises, dtype="real").contiguous()
            assert tuple(self.Observation_noises.shape) == (S1, E, n)
        self._fed_assigned = True

    def _affine_drift(self):
        """Optional hook for the in-kernel action-affine drift functor: return ``(A0, A_act)`` with ``A0`` ``(n,n)`` or
        ``(n_envs,n,n)`` and ``A_act`` ``(n_actions,n,n)`` to have ``A = A0 + sum_a u_a A_act[a]`` (u = actions *
        action_maximums) evaluated inside the integration launch instead of calling ``get_A``; ``None`` = not affine."""
        return None

    def _coefficients(self, K):
        args = (self.actions, None, None)
        E, n = self.n_envs, self.n_observations
        if self.is_A_time_dependent and K > 0:
            A = torch.stack([self._as_A(self.get_A(self.t_idx + i, args)) for i in range(K)]).contiguous()
            nz = torch.stack([self._as_nz(self.get_noise_prefixes(self.t_idx + i, args)) for i in range(K)]).contiguous()
            return A, E * n * n, n * n, nz, E * n, n
        A = self.backend.convert_to_typed(tensor=self.get_A(self.t_idx, args), dtype="real").contiguous()
        nz = self.backend.convert_to_typed(tensor=self.get_noise_prefixes(self.t_idx, args), dtype="real").contiguous()
        assert tuple(A.shape) in [(n, n), (E, n, n)], "``get_A`` should return (n, n) or (n_envs, n, n)"
        assert tuple(nz.shape) in [(n,), (E, n)], "``get_noise_prefixes`` should return (n,) or (n_envs, n)"
        return A, 0, (n * n if A.dim() == 3 else 0), nz, 0, (n if nz.dim() == 2 else 0)

    def _as_A(self, A):
        A = self.backend.convert_to_typed(tensor=A, dtype="real")
        return A.expand(self.n_envs, -1, -1) if A.dim() == 2 else A

    def _as_nz(self, nz):
        nz = self.backend.convert_to_typed(tensor=nz, dtype="real")
        return nz.expand(self.n_envs, -1) if nz.dim() == 1 else nz

    def _launch(self, K, x0, write_traj=True, graph_mode=None):
        a, keep = self.interval_args(K, x0, write_traj, graph_mode)
        peer = self._peer if K > 0 else None
        if peer is not None:
            self._peer_patch(a, host_io=getattr(self, "_host_step", False))
        ev = getattr(self, "_em_events", None)
        if ev is not None and K > 0:  # live per-launch timing on the launch stream (bench.py roofline)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.em_step(a)
            e1.record()
            ev.append((e0, e1))
        else:
            N.em_step(a)
        if peer is not None:
            peer.after_launch(N.current_stream())
        return keep

    def interval_args(self, K, x0=None, write_traj=True, graph_mode=None):
        """Build the ``qrl_em_args`` of one action interval (evaluates the coefficient hooks).  Returns
        ``(args, tensors_to_keep_alive)``; ``_launch`` enqueues it, ``bench.py`` re-launches it to time the kernel.
        ``graph_mode`` (``"device"`` / ``"host"``) is set while a step is captured into a CUDA graph: the kernel's last
        block then advances the device-side time index itself and, for the host round trip, delivers this step's
        [min, max] without a reset copy."""
        x0 = self._x_last if x0 is None else x0
        E, n = self.n_envs, self.n_observations
        a = N.EmArgs()
        a.n, a.n_envs, a.K, a.t_ssz = n, E, K, self.t_ssz
        keep = []
        affine = self._affine_drift() if (K > 0 and self.fused_coefficients) else None
        if affine is not None:
            # drift evaluated in-kernel from the RAW actions (scaling by action_maximums included): no hook launches
            A0, A_act = affine
            nz = self.backend.convert_to_typed(tensor=self.get_noise_prefixes(self.t_idx, (self.actions, None, None)), dtype="real").contiguous()
            keep += [A0, A_act, nz]
            a.A, a.A_stride_k, a.A_stride_e = N.ptr(A0), 0, (n * n if A0.dim() == 3 else 0)
            a.nz, a.nz_stride_k, a.nz_stride_e = N.ptr(nz), 0, (n if nz.dim() == 2 else 0)
            a.A_act, a.action_scale = N.ptr(A_act), N.ptr(self.action_maximums)
            a.actions, a.n_actions = N.ptr(self._actions_in), self.n_actions
        elif K > 0:
            A, a.A_stride_k, a.A_stride_e, nz, a.nz_stride_k, a.nz_stride_e = self._coefficients(K)
            keep += [A, nz]
            a.A, a.nz = N.ptr(A), N.ptr(nz)
        if self.rng == "philox":
            a.rng_mode = N.RNG_PHILOX
            a.seed, a.episode, a.t_idx, a.env_offset = self._philox_seed, self.batch_idx & 0xFFFFFFFF, self.t_idx, self.env_offset
            if self.use_cuda_graph:
                a.dyn = N.ptr(self._dyn)
            if self.observation_stds is not None:
                a.obs_std = N.ptr(self.observation_stds)
                a.obs_std_stride_e = n if self.observation_stds.dim() == 2 else 0
        else:
            a.rng_mode = N.RNG_FED
            t_off = 0 if self.use_cuda_graph else self.t_idx      # graph mode: the kernel adds dyn[0] rows itself
            if self.use_cuda_graph:
                a.dyn = N.ptr(self._dyn)
            if K > 0:
                a.W = N.ptr(self.Ws) + 8 * t_off * E * n
            if self.Observation_noises is not None:
                a.obs_noise = N.ptr(self.Observation_noises) + 8 * t_off * E * n
        a.x0 = N.ptr(x0)
        if write_traj and self.store_trajectory:
            a.states, a.observations = N.ptr(self._States), N.ptr(self._Observations)
        a.x_last = N.ptr(self._x_last)
        a.obs_last = self._obs_last_ptr if self._obs_last_ptr is not None else N.ptr(self._obs_last)
        if self._reward_fused and K > 0:
            a.reward_kind = _REWARD_KINDS[self.reward_functor]
            a.reward_w = N.ptr(self._reward_w)
            a.reward = N.ptr(self._Reward) if self.store_trajectory else None
            a.reward_last = self._reward_last_ptr if self._reward_last_ptr is not None else N.ptr(self._reward_last)
            a.rewards_cum = N.ptr(self.rewards)
        a.obs_minmax = N.ptr(self._minmax)
        a.nan_flag = N.ptr(self._nan)
        if self._peer is not None and K > 0 and self._peer.mode == "pull":
            # publish-only exchange: constant part (the per-launch part — ring slice, epoch — is patched in by _peer_patch)
            pg = self._peer
            a.finish_counter = N.ptr(self._finish)
            a.peer_flags, a.peer_rank, a.peer_world, a.peer_dst = N.ptr(pg.flag_ptrs), pg.rank, pg.world, pg.dst
            a.peer_pulled = pg.pulled_ptr
            a.flags |= N.EM_PEER_PUBLISH
        if graph_mode is not None:
            a.finish_counter, a.dyn_advance = N.ptr(self._finish), K
            if self._peer_barrier is not None:
                flag_ptrs, epoch, prank, pworld = self._peer_barrier
                a.peer_flags, a.peer_epoch, a.peer_rank, a.peer_world = N.ptr(flag_ptrs), N.ptr(epoch), prank, pworld
            if graph_mode == "host":
                a.obs_minmax, a.minmax_out = N.ptr(self._minmax_acc), N.ptr(self._minmax)
                zc = self._zero_copy()
                if zc is not None:
                    a.obs_last, a.minmax_out = zc["out"], zc["out"] + 8 * E * (n + 1)
                    a.nan_flag = zc["out"] + 8 * (E * (n + 1) + 2)
                    if a.reward_kind != N.REWARD_NONE:
                        a.reward_last = zc["out"] + 8 * E * n
                    if affine is not None:
                        a.actions = zc["actions"]
        self._packed_by_kernel = False
        if (self._reward_fused and K > 0 and self.n_properties == 0 and self.store_trajectory
                and not self.cache_all_data and self.fuse_cache_rows):
            # the (E, shape_T, n_sel) ``data`` cache rows of this interval are written by the same launch
            n_sel = len(self.data_idxs)
            t_off = 0 if self.use_cuda_graph else self.t_idx
            a.packed = N.ptr(self._data_dev) + 8 * t_off * self._data_pitch
            a.packed_stride_e = self.shape_T[0] * self._data_pitch
            a.packed_row_pitch = self._data_pitch
            a.T_step = N.ptr(self.T) + 8 * t_off
            if affine is None:
                a.actions, a.n_actions = N.ptr(self.actions), self.n_actions
            a.sel_idxs, a.n_sel = N.ptr(self._sel_idxs), n_sel
            self._packed_by_kernel = True
        if self.force_generic_kernel:
            a.flags |= N.EM_FORCE_GENERIC
        a.flags |= int(os.environ.get("QRL_EM_DEBUG_FLAGS", "0"))    # development knob (kernel diagnostics bits)
        return a, keep

    def attach_peer(self, peer):
        """Route the step outputs through ``distributed.PeerGather``'s ring: every integration launch (K > 0) writes its
        last rows into the next ring slice and publishes its epoch to the learner rank."""
        assert not (self.use_cuda_graph and not self._single_kernel_step()), \
            "the peer exchange rotates the output slice launch by launch: it needs eager launches or the single-kernel step"
        self._peer = peer
        self._graph = self._graph_host = self._graph_host1 = self._fast_host = None
        self._static = {}

    def _peer_patch(self, a, host_io):
        """Per-launch part of the publish-only exchange: the ring slice of this epoch, its number, the credit it needs."""
        e, obs_ptr, rew_ptr, pulled_min = self._peer.next_launch()
        # the ring gets the second copy of the last rows; the primary outputs stay where step() / step_device() read them
        # (the env's own buffer, or pinned host memory), so reset launches and host fetches never touch a published slice
        a.obs_last_b, a.reward_last_b = obs_ptr, rew_ptr
        a.peer_publish_value, a.peer_pulled_min = e, pulled_min

    def attach_peer_barrier(self, flag_ptrs, epoch, rank, world):
        """Fuse the per-step cross-GPU barrier of the sharded path into the captured step kernel (ABI 3)."""
        self._peer_barrier = (flag_ptrs, epoch, int(rank), int(world))
        self._graph = self._graph_host = self._graph_host1 = self._fast_host = None
        self._static = {}

    def _zero_copy(self):
        """Device-side addresses of the pinned step buffers (``_h_out``, ``_h_actions``) when the zero-copy host round
        trip applies, else ``None``."""
        if not (self.host_zero_copy and self.return_numpy and self._obs_last_ptr is None and self._reward_fused):
            return None
        if self._zc is None:
            try:
                act = N.host_device_pointer(self._h_actions)
                self._zc = [{"out": N.host_device_pointer(b), "actions": act} for b in self._h_out_bufs]
            except N.NativeError:
                self._zc = False    # pinned memory not mapped on this platform: keep the copy nodes
        return self._zc[self._h_sel] if self._zc else None

    def _initial_observations(self, states_0):
        self._launch(0, states_0, write_traj=False)  # row 0 only: obs_0 = states_0 + noise[0]
        return self._obs_last.clone()

    def _single_kernel_step(self):
        """True when a captured action interval consists of ONE kernel of the native library and nothing else (in-kernel
        drift functor, fused reward): its argument block is then cached and re-launched directly — no CUDA graph, and
        the actions pointer can change from step to step."""
        return self.use_cuda_graph and self.fused_coefficients and self._reward_fused and self._affine_drift() is not None

    def _direct_actions(self, a):
        if not (self._single_kernel_step() and isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64
                and a.is_contiguous() and a.numel() == self.n_envs * self.n_actions):
            self._actions_src = None
            return False
        self._actions_src = a           # consumed by the launch of this step (or copied for a tail interval)
        return True

    def _update_states(self):
        dim = self._dim
        self.last_step_peer_fenced = False
        src = getattr(self, "_actions_src", None)
        if src is not None and not (self.use_cuda_graph and dim == self.action_interval + 1):
            self._actions_in.copy_(src.reshape(self.n_envs, self.n_actions), non_blocking=True)   # not the fast path after all
            self._actions_src = src = None
        if not self.use_cuda_graph:
            self._launch(dim - 1, self._x_last)
        elif dim == self.action_interval + 1 and self._single_kernel_step() and self._peer is not None:
            # sharded path: the output ring slice rotates launch by launch, so the host round trip re-enqueues its block too
            self._relaunch_interval(host_io=getattr(self, "_host_step", False), actions_src=src)
            self._actions_src = None
        elif dim == self.action_interval + 1 and self._single_kernel_step() and not getattr(self, "_host_step", False):
            # device-resident step: direct launch, actions read in place.  (The host round trip keeps its captured
            # graph: its pointers never change and a replay costs ~4 us less host time than a ctypes launch.)
            self._relaunch_interval(host_io=False, actions_src=src)
            self._actions_src = None
            self.last_step_peer_fenced = self._peer_barrier is not None
        elif dim == self.action_interval + 1:
            self._replay_interval(host_io=getattr(self, "_host_step", False))
            self.last_step_peer_fenced = self._peer_barrier is not None
        else:  # tail interval (fewer sub-steps): eager launch, then advance the device-side time index
            if getattr(self, "_host_step", False):
                self._actions_in.copy_(self._h_actions, non_blocking=True)
                self._minmax.copy_(self._minmax_init, non_blocking=True)
            self.actions = (self._actions_in * self.action_maximums_batch)
            if self.__dict__.get("_dyn_stale"):
                self._sync_dyn()                # the eager launch reads the device-side index too
            self._launch(dim - 1, self._x_last)
            self._dyn[0] += dim - 1
        return self._States[:dim]

    def _relaunch_interval(self, host_io=False, actions_src=None):
        """Single-kernel step: the ``qrl_em_args`` block of one full action interval is built once (device-resident time
        index, last-block bookkeeping) and enqueued again every step with only the actions pointer updated — the
        caller's CUDA tensor for ``step_device``, the pinned host buffer (zero-copy) or the staging buffer otherwise."""
        key = ("host", self._h_sel) if host_io else "device"    # one block per pinned result buffer (host_output_views)
        st = self._static.get(key)
        if st is None:
            zc = self._zero_copy() if host_io else None
            a, keep = self.interval_args(self.action_interval, self._x_last, True, graph_mode="host" if host_io else "device")
            if not host_io:
                # back-to-back step kernels of a device-resident loop: programmatic dependent launch hides the launch
                # latency between them (measured 32.2 -> 30.3 us per step at config 3); the host re-enqueues the block
                # every step anyway, so it also supplies the time index (the kernel still advances the device-side copy
                # for the graph paths) and the kernel does not wait for a global load before its first Philox counter
                a.flags |= N.EM_PDL | N.EM_HOST_TIDX
                if os.environ.get("QRL_DEV_NO_PDL"):          # development knob (A/B of the device-resident loop)
                    a.flags &= ~N.EM_PDL
                # ... and because it does, the device-side index need not be advanced by this launch: without the
                # last-block protocol (one __syncthreads, a fence and an atomic round trip on every CTA's way out) a step
                # of the chained loop takes 23.0 instead of 24.9 us at config 3.  The index is brought up to date by a
                # scalar fill before the next launch that reads it (``_sync_dyn``).  The kernel-driven exchanges (pull /
                # fused barrier) keep the protocol: their last block publishes the epoch.
                if (self._peer is None or self._peer.mode == "push") and self._peer_barrier is None \
                        and not os.environ.get("QRL_DEV_KEEP_FINISH"):
                    a.finish_counter, a.dyn_advance = None, 0
            stream = torch.cuda.current_stream()
            self._replay_stream = stream       # looked up once: the call costs ~10 us per step otherwise
            st = self._static[key] = {"a": a, "keep": keep, "zc": zc, "packs": self._packed_by_kernel,
                                      "stream": C.c_void_p(stream.cuda_stream), "default_actions": a.actions,
                                      # bases of the t_idx-dependent pointers (interval_args leaves them at row 0 in graph mode)
                                      "W": a.W, "obs_noise": a.obs_noise, "T_step": a.T_step, "packed": a.packed}
            self.graph_native_launches = 1
        a = st["a"]
        if host_io and st["zc"] is None:
            self._actions_in.copy_(self._h_actions, non_blocking=True)
        if not host_io:
            a.actions = actions_src.data_ptr() if actions_src is not None else st["default_actions"]
            t = self.t_idx                      # EM_HOST_TIDX: what resolve_dyn() would derive from the device-side index
            a.t_idx, a.episode = t, self.batch_idx & 0xFFFFFFFF
            rows = 8 * t * self.n_envs * self.n_observations
            if st["W"]:
                a.W = st["W"] + rows
            if st["obs_noise"]:
                a.obs_noise = st["obs_noise"] + rows
            if st["T_step"]:
                a.T_step = st["T_step"] + 8 * t
            if st["packed"]:
                a.packed = st["packed"] + 8 * t * (a.packed_row_pitch if a.packed_row_pitch > 0 else a.n_sel)
        self._packed_by_kernel = st["packs"]
        if self._peer is not None:
            self._peer_patch(a, host_io)
        if not host_io and not a.finish_counter:
            self._dyn_stale = True              # this launch does not advance the device-side time index
        elif self.__dict__.get("_dyn_stale"):
            self._sync_dyn()
        N.em_step(a, st["stream"])
        if self._peer is not None:
            self._peer.after_launch(st["stream"])
        if host_io and st["zc"] is None:
            self._h_out.copy_(self._step_out, non_blocking=True)
        self._host_graph_done = host_io

    # ---- step(host actions): the lean host round trip of the single-kernel step -------------------------------------
    def _fast_host_setup(self):
        """State of the lean ``step(host actions)`` path, or ``False`` when it does not apply: it needs the single-kernel
        step with the zero-copy round trip (the launch reads the actions from and writes ``[obs | reward | min,max |
        status]`` to pinned memory) and cache rows written by the same launch — then a full action interval is one
        kernel and nothing else, and everything ``update()`` / ``update_data()`` do besides is host bookkeeping.

        Two flavours of the launch + wait (measured at config 3 with nothing else on the host,
        ``tools/e2e_breakdown.py``, ``profiles/r02y_e2e_breakdown.log``): replay of the captured one-kernel graph + stream
        synchronisation 35.7 us (default), ``qrl_em_step_host`` (ABI 5: direct launch + spin on a completion word in pinned
        memory, ``QRL_HOST_STEP_NATIVE=1``) 36.2 us — the kernel itself takes 22.9 us, the rest is launch latency and the
        164 KB of results crossing PCIe at the end of the launch."""
        if (os.environ.get("QRL_NO_FAST_HOST_STEP") or not self._single_kernel_step() or self._peer is not None
                or self._peer_barrier is not None or self.n_properties > 0 or self.cache_all_data
                or getattr(self, "_pre_sync_hook", None) is not None or getattr(self, "_em_events", None) is not None):
            return False
        if self._zero_copy() is None or (self.store_trajectory and not self.fuse_cache_rows):
            return False
        native = os.environ.get("QRL_HOST_STEP_NATIVE")      # "1": spin on the completion word, "sync": stream synchronisation
        if not native:
            return {"native": False}
        E, no, blocks, sel0 = self.n_envs, self.n_observations, [], self._h_sel
        for sel in range(len(self._h_out_bufs)):
            self._h_sel = sel
            zc = self._zero_copy()
            a, keep = self.interval_args(self.action_interval, self._x_last, True, graph_mode="host")
            a.flags |= N.EM_HOST_TIDX           # the host supplies the time index (dyn is still advanced by the last block)
            spin = native != "sync"
            if spin:
                a.done_flag = zc["out"] + 8 * (E * (no + 1) + 3)
            words = self._h_out_nps[sel].view(np.uint64)
            blocks.append({"a": a, "ref": C.byref(a), "keep": keep, "packs": self._packed_by_kernel,
                           "T_step": a.T_step, "packed": a.packed, "W": a.W, "obs_noise": a.obs_noise,
                           "pitch": 8 * (a.packed_row_pitch if a.packed_row_pitch > 0 else a.n_sel),
                           "done": C.c_void_p(words[E * (no + 1) + 3:].ctypes.data) if spin else None})
        self._h_sel = sel0
        stream = torch.cuda.current_stream()
        self._replay_stream = stream
        self._fast_seq = 0
        return {"native": True, "blocks": blocks, "stream": C.c_void_p(stream.cuda_stream), "fn": N.lib().qrl_em_step_host}

    def step(self, actions):
        """``step_async`` + ``step_wait`` (base.py:1233-1293).  Host actions on a full action interval of the single-kernel
        step take the lean path: copy the actions into pinned memory, replay the captured one-kernel graph (or call
        ``qrl_em_step_host``), wait, and do the bookkeeping of ``update()`` / ``update_data()`` inline."""
        fast = self.__dict__.get("_fast_host")
        K = self.action_interval
        t = self.t_idx
        host_full = not isinstance(actions, torch.Tensor) and t + K < self.shape_T[0]
        if fast is None and host_full:
            fast = self._fast_host = self._fast_host_setup()
        if not fast or not host_full:
            return super().step(actions)
        sel = self._h_sel ^ 1 if self.host_output_views else self._h_sel    # this step's results go to the other pinned buffer
        if not fast["native"]:
            graph = self._graph_host1 if sel else self._graph_host
            if graph is None:                   # not captured yet: the general path captures it
                return super().step(actions)
        self._h_actions_np[...] = np.asarray(actions, dtype=np.float64).reshape(self.n_envs, self.n_actions)
        if sel != self._h_sel:
            self._h_sel = sel
            self._h_out, self._h_out_np = self._h_out_bufs[sel], self._h_out_nps[sel]
        if fast["native"]:
            b = fast["blocks"][sel]
            a = b["a"]
            a.t_idx, a.episode = t, self.batch_idx & 0xFFFFFFFF
            if b["packed"]:
                a.T_step, a.packed = b["T_step"] + 8 * t, b["packed"] + b["pitch"] * t
            if b["W"]:                          # rng='fed': rows of the fed draws
                a.W = b["W"] + 8 * t * self.n_envs * self.n_observations
                if b["obs_noise"]:
                    a.obs_noise = b["obs_noise"] + 8 * t * self.n_envs * self.n_observations
            self._fast_seq = a.done_value = self._fast_seq + 1
            rc = fast["fn"](b["ref"], fast["stream"], b["done"], 0)
            if rc != 0:
                N.check(rc, "qrl_em_step_host")
            self._packed_by_kernel = b["packs"]
        else:
            if self.__dict__.get("_dyn_stale"):
                self._sync_dyn()
            self.graph_replays += 1
            graph.replay()
            self._replay_stream.synchronize()
            self._packed_by_kernel = self._graph_packs
        # bookkeeping of update() / update_data() for a full interval (base.py:440-483, 1295-1380)
        self._host_step, self._host_graph_done = True, False
        self._dim, self._t_range = K + 1, (t, t + K + 1)
        self.States, self.Observations, self.Reward = self._States, self._Observations, self._Reward
        self._data_host = None
        self.last_step_peer_fenced = False
        self.t_idx = t = t + K
        self.t = self._T_host[t]
        self.action_idx += 1
        return self._finish_host_step(self._h_out, self._obs_last, self._reward_last, not t + 1 < self.shape_T[0])

    def step_device(self, actions):
        """Device-resident step (``BaseVecEnv.step_device``).  A full action interval of the single-kernel step whose
        actions are a contiguous float64 CUDA tensor takes a lean path: the cached argument block of ``_relaunch_interval``
        is patched and enqueued here and the bookkeeping of ``update()`` / ``update_data()`` is done inline — the host side
        of a step is what limits a multi-GPU loop once the kernel takes 23 us (8 ranks on one box share its cores)."""
        st = self._static.get("device")
        K, t = self.action_interval, self.t_idx
        if (st is None or not st.get("lean") or t + K >= self.shape_T[0] or not isinstance(actions, torch.Tensor)
                or not actions.is_cuda or actions.dtype != torch.float64 or not actions.is_contiguous()
                or actions.numel() != self.n_envs * self.n_actions):
            out = super().step_device(actions)
            st = self._static.get("device")
            if st is not None and "lean" not in st:     # the general path has just built the block: does the lean path apply?
                st["lean"] = (self.n_properties == 0 and not self.cache_all_data and self._cum_fused
                              and (st["packs"] or not self.store_trajectory) and getattr(self, "_em_events", None) is None
                              and not os.environ.get("QRL_NO_FAST_DEVICE_STEP"))
                st["ref"], st["fn"] = C.byref(st["a"]), N.lib().qrl_em_step
                st["no_finish"] = not st["a"].finish_counter
                st["pitch"] = 8 * (st["a"].packed_row_pitch if st["a"].packed_row_pitch > 0 else st["a"].n_sel)
                st["episode"] = None
            return out
        a = st["a"]
        a.actions, a.t_idx = actions.data_ptr(), t
        if st["episode"] != self.batch_idx:
            st["episode"] = self.batch_idx
            a.episode = self.batch_idx & 0xFFFFFFFF
        if st["packed"]:
            a.T_step, a.packed = st["T_step"] + 8 * t, st["packed"] + st["pitch"] * t
        if st["W"]:                             # rng='fed': rows of the fed draws
            rows = 8 * t * self.n_envs * self.n_observations
            a.W = st["W"] + rows
            if st["obs_noise"]:
                a.obs_noise = st["obs_noise"] + rows
        peer = self._peer
        if peer is not None:
            e, a.obs_last_b, a.reward_last_b, pulled_min = peer.next_launch()
            if pulled_min or peer.mode != "push":
                a.peer_publish_value, a.peer_pulled_min = e, pulled_min
        if st["no_finish"]:
            self._dyn_stale = True              # this launch does not advance the device-side time index
        elif self.__dict__.get("_dyn_stale"):
            self._sync_dyn()
        rc = st["fn"](st["ref"], st["stream"])
        if rc != 0:
            N.check(rc, "qrl_em_step")
        if peer is not None:
            peer.after_launch(st["stream"])
        # bookkeeping of update() / update_data() for a full interval (base.py:440-483, 1295-1380)
        self._host_step = self._host_graph_done = False
        self._actions_src = None
        self._dim, self._t_range = K + 1, (t, t + K + 1)
        self.States, self.Observations, self.Reward = self._States, self._Observations, self._Reward
        self._packed_by_kernel, self._data_host = st["packs"], None
        self.last_step_peer_fenced = self._peer_barrier is not None
        self.t_idx = t = t + K
        self.t = self._T_host[t]
        self.action_idx += 1
        terminated = not t + 1 < self.shape_T[0]
        if terminated:
            self.data_rewards.append(self.rewards.clone())
        return self._obs_last, self._reward_last, terminated

    def _sync_dyn(self):
        """The device-side time index = the host's, before a launch that reads it (graph replays, the host round trip)
        follows device-resident steps that did not advance it."""
        self._dyn[0].fill_(self.t_idx)
        self._dyn_stale = False

    def _replay_interval(self, host_io=False):
        """One full action interval as ONE graph launch: actions scaling, the coefficient hooks (batched torch) and the
        fused integration kernel (which also advances the device-side time index) are captured once and replayed.
        ``host_io=True`` (the ``step(host actions)`` path) replays a second graph that also covers the host round trip:
        with ``host_zero_copy`` (default, ``return_numpy`` envs) the kernel itself reads the actions from and writes
        ``[obs_last | reward_last | min,max]`` to pinned host memory over PCIe, so the graph is still one kernel;
        otherwise it starts with the H2D copy of the actions and ends with the D2H copy of the results."""
        # the host graph is captured per pinned result buffer (two of them when step() hands out views)
        assert self._peer is None, "the peer exchange needs the single-kernel step (fused_coefficients + reward_functor)"
        key = ("_graph_host" if self._h_sel == 0 else "_graph_host1") if host_io else "_graph"
        if getattr(self, key, None) is None:
            K = self.action_interval
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            n0 = N.launch_count()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    affine = self.fused_coefficients and self._affine_drift() is not None
                    zc = self._zero_copy() if host_io else None
                    if host_io and not (zc is not None and affine):
                        self._actions_in.copy_(self._h_actions, non_blocking=True)
                    if not affine:
                        self.actions = self._actions_in * self.action_maximums_batch
                    # the kernel's last block advances dyn[0] by K and (host round trip) hands over this step's
                    # [min, max]: no reset copy, no second kernel
                    keep = self._launch(K, self._x_last, graph_mode="host" if host_io else "device")
                    if host_io and zc is None:
                        self._h_out.copy_(self._step_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(side)
            self._replay_stream = torch.cuda.current_stream()   # looked up once: the call costs ~10 us per step otherwise
            setattr(self, key, graph)
            setattr(self, key + "_keep", keep)
            self._graph_packs = self._packed_by_kernel
            self.graph_native_launches = N.launch_count() - n0   # kernels of libquantrl_b200.so inside one replay
        if self.__dict__.get("_dyn_stale"):
            self._sync_dyn()
        self.graph_replays += 1
        self._packed_by_kernel = self._graph_packs
        getattr(self, key).replay()
        self._host_graph_done = host_io
