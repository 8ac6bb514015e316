"""Host-side mirror of the reference's vectorized env surface for the ``'b200'`` backend.

``BaseVecEnv`` restates, from scratch, the lock-step semantics of ``BaseEnv`` + ``BaseSB3Env``
(quantrl/envs/base.py:30-667, 999-1492): the same constructor arguments and ``kwargs`` (base.py:92-111), the same
public attributes (``T``, ``t_idx``, ``States``, ``Observations``, ``Reward``, ``Properties``, ``actions``,
``rewards``, ``data``, ``data_rewards``, ``batch_idx`` …), the same user hooks (``reset_states``, ``get_reward``,
``get_properties``) and the Stable-Baselines3 ``VecEnv`` protocol (``reset/step_async/step_wait/step/close`` and the
``env_is_wrapped/env_method/get_attr/set_attr`` replication helpers, base.py:1097-1183).

What differs is where the work happens: ``_update_states`` of the subclasses is ONE launch of the fused CUDA kernel
(through the C ABI) that also produces the noisy observations, the reward rows, the cumulative-reward update and
the truncation reductions; ``update_data`` is one launch of the pack kernel that writes the ``(E, dim, n_data)``
cache rows and the selected ``data`` columns straight into a device-resident ``(E, shape_T, n_sel)`` cache that is
copied to the host lazily.  Index conventions follow SURVEY.md Appendix A line by line.
"""
import math
import os
import queue
import threading
import warnings

import numpy as np
import torch

from .. import _native as N
from ..backends import get_backend_instance

try:  # the real SB3 base class when it is installed, else a protocol-compatible stand-in
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv
except Exception:  # noqa: BLE001
    _SB3VecEnv = None


class _Box:
    """Stand-in for ``gymnasium.spaces.Box`` (only the attributes the reference touches, base.py:173-206)."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        lo, hi = max(self.low, -1e3), min(self.high, 1e3)
        return np.random.uniform(lo, hi, self.shape).astype(self.dtype)


class _MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
        self.shape = self.nvec.shape

    def sample(self):
        return np.array([np.random.randint(k) for k in self.nvec])


def _spaces():
    try:
        from gymnasium.spaces import Box, MultiDiscrete
        return Box, MultiDiscrete
    except Exception:  # noqa: BLE001
        return _Box, _MultiDiscrete


class PartCacheWriter:
    """Trajectory part files in the reference's on-disk format (quantrl/io.py:56-80): one ``.npz`` per action step,
    named ``{b*E}_{(b+1)*E-1}_{part}.npz`` with the array under ``arr_0``.  Unlike the reference (whose "async" dump
    joins its thread immediately, io.py:74-80) the write really happens on a background thread."""

    def __init__(self, disk_cache_dir, cache_dump_interval, compressed=True, env_offset=0, n_envs_total=None):
        self.disk_cache_dir = disk_cache_dir
        self.cache_dump_interval = cache_dump_interval
        # sharded batches (distributed.make_sharded_env): the names carry GLOBAL env indices, so that ranks sharing one
        # directory never overwrite each other's parts; an unsharded env gives exactly the reference's names
        self.env_offset = int(env_offset)
        self.n_envs_total = int(cache_dump_interval if n_envs_total is None else n_envs_total)
        self.compressed = compressed
        self._q = None
        self._thread = None
        self._error = None

    def _start(self):
        os.makedirs(self.disk_cache_dir, exist_ok=True)
        self._q = queue.Queue(maxsize=8)
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            path, data = item
            if self._error is None:     # after a failure the queue is still drained, so that the env never blocks on it
                try:
                    (np.savez_compressed if self.compressed else np.savez)(path, data)
                except Exception as e:  # noqa: BLE001
                    self._error = e

    def _raise_pending(self):
        """A write that failed on the background thread surfaces in the caller's thread, once."""
        if self._error is not None:
            e, self._error = self._error, None
            raise RuntimeError(f"writing a trajectory part file under {self.disk_cache_dir} failed") from e

    def dump_part_async(self, data, batch_idx, part_idx):
        self._raise_pending()
        if self._thread is None:
            self._start()
        self._q.put((self.part_path(batch_idx, part_idx), data))

    def flush(self):
        """Block until every queued part file is on disk (the writer thread restarts with the next dump)."""
        if self._thread is not None:
            self._q.put(None)
            self._thread.join()
            self._thread = None
        self._raise_pending()

    def part_path(self, batch_idx, part_idx):
        first = batch_idx * self.n_envs_total + self.env_offset
        return os.path.join(self.disk_cache_dir, "_".join([
            str(first), str(first + self.cache_dump_interval - 1), str(part_idx)]) + ".npz")

    def load_parts(self, batch_idx, idxs=None):
        """Read the part files of one batch back as ONE ``(E, rows, n_data)`` block (the reader the reference leaves
        as a TODO, io.py:20; its ``get_disk_cache`` only understands the un-parted ``{start}_{end}.npz`` names).  Part
        ``p`` holds rows ``t_idx .. t_idx+dim-1`` of action step ``p``; its row 0 repeats the last row of part ``p-1``
        (with that step's cumulative reward — the host cache keeps the LATER one, base.py:1370-1372), so the stitched
        block drops the last row of every part but the final one."""
        self.flush()
        parts, p = [], 0
        while os.path.exists(self.part_path(batch_idx, p)):
            with np.load(self.part_path(batch_idx, p)) as f:
                a = f["arr_0"]
            parts.append(a if idxs is None else a[:, :, idxs])
            p += 1
        if not parts:
            raise FileNotFoundError(self.part_path(batch_idx, 0))
        return np.concatenate([a[:, :-1] for a in parts[:-1]] + [parts[-1]], axis=1)

    def close(self, dump_cache=True):
        self.flush()


class BaseVecEnv(*(() if _SB3VecEnv is None else (_SB3VecEnv,))):
    """Vectorized, lock-stepped environment batch on one B200 (mirror of ``BaseSB3Env``, base.py:999-1492)."""

    base_env_kwargs = {
        "has_delay": False,
        "observation_space_range": [-1e12, 1e12],
        "observation_stds": None,
        "action_space_range": [-1.0, 1.0],
        "action_space_type": "box",
        "seed": None,
        "cache_all_data": True,
        "cache_dump_interval": 100,
        "average_over": 100,
        "plot": False,              # matplotlib live plots are out of scope
        # ---- additions of the b200 backend ----
        "rng": "philox",            # 'philox': in-kernel Philox4x32-10 noise; 'fed': pre-drawn tensors (reference layout)
        "return_numpy": False,      # step()/reset() return NumPy (pinned D2H) instead of device tensors
        "host_output_views": False,  # with return_numpy: step() returns VIEWS of the pinned step buffer (double buffered:
                                     # valid until the end of the NEXT step) instead of fresh copies — the reference
                                     # returns views of its own buffers too (base.py:1293); saves two host copies per step
        "env_offset": 0,            # global index of local env 0 (multi-GPU sharding)
        "n_envs_total": None,       # size of the global batch this env is a shard of (None: unsharded)
        "store_trajectory": True,   # keep the (K+1,E,n) States/Observations/Reward blocks (reference semantics)
        "cache_compressed": True,
    }

    def __init__(self, backend, t_norm_max, t_norm_ssz, t_norm_mul, n_envs, n_observations, n_properties, n_actions,
                 action_maximums, action_interval, data_idxs, dir_prefix, file_prefix, **kwargs):
        for key, val in self.base_env_kwargs.items():
            kwargs.setdefault(key, val)
        # same validation as base.py:135-148
        assert t_norm_max > t_norm_ssz, "maximum normalized time should be greater than the normalized step size"
        assert n_properties >= 0, "parameter ``n_properties`` should be non-negative"
        assert action_interval > 0, "parameter ``action_interval`` should be a positive integer"
        assert len(data_idxs) > 0, "parameter ``data_idxs`` should be a list containing at least one element"
        assert kwargs["observation_stds"] is None or isinstance(kwargs["observation_stds"], (list, np.ndarray, torch.Tensor)), \
            "parameter ``observation_stds`` should be a list"
        assert kwargs["seed"] is None or isinstance(kwargs["seed"], int), "parameter ``seed`` should be an integer or ``None``"
        assert isinstance(kwargs["cache_all_data"], bool), "parameter ``cache_all_data`` should be a boolean"
        assert len(kwargs["observation_space_range"]) == 2 and len(kwargs["action_space_range"]) == 2
        assert kwargs["action_space_type"] in ["binary", "box"], "parameter ``action_space_type`` can be either ``'binary'`` or ``'box'``"
        assert kwargs["rng"] in ["philox", "fed"], "parameter ``rng`` can be either ``'philox'`` or ``'fed'``"

        N.lib()  # fail loudly, now, if the CUDA extension is missing
        self.backend = backend
        self.device = backend.device
        self.numpy_int = backend.dtypes["numpy"][backend.precision]["integer"]
        self.numpy_real = backend.dtypes["numpy"][backend.precision]["real"]
        assert backend.precision == "double", "the b200 hot path computes in FP64 (backend_precision='double')"
        f64 = torch.float64

        # time grid (base.py:157-168)
        self.t_norm_max, self.t_norm_ssz, self.t_norm_mul = t_norm_max, t_norm_ssz, t_norm_mul
        self.shape_T = (int(self.numpy_int(t_norm_max / t_norm_ssz)) + 1,)
        self.T_norm = np.arange(self.shape_T[0], dtype=self.numpy_real) * t_norm_ssz
        self._T_host = self.T_norm * t_norm_mul
        self.T = torch.as_tensor(self._T_host, device=self.device)
        self.t_ssz = t_norm_ssz * t_norm_mul

        # spaces (base.py:170-212)
        Box, MultiDiscrete = _spaces()
        self.n_envs = int(n_envs)
        self.n_observations = int(n_observations)
        self.observation_space_range = kwargs["observation_space_range"]
        self.observation_space = Box(low=self.observation_space_range[0], high=self.observation_space_range[1],
                                     shape=(self.n_observations,), dtype=self.numpy_real)
        self.n_properties = int(n_properties)
        self.n_actions = int(n_actions)
        self.action_space_type = kwargs["action_space_type"]
        if self.action_space_type == "binary":
            self.action_space_range = [0, 1]
            self.action_space = MultiDiscrete(nvec=[2] * self.n_actions)
        else:
            self.action_space_range = kwargs["action_space_range"]
            self.action_space = Box(low=self.action_space_range[0], high=self.action_space_range[1],
                                    shape=(self.n_actions,), dtype=self.numpy_real)
        self.action_maximums = torch.as_tensor(np.asarray(action_maximums, dtype=np.float64), device=self.device)
        self.action_maximums_batch = self.action_maximums.reshape(1, self.n_actions).repeat(self.n_envs, 1).contiguous()
        self.action_interval = int(action_interval)
        self.action_steps = int(math.ceil((self.shape_T[0] - 1) / self.action_interval))
        self.has_delay = bool(kwargs["has_delay"])     # honoured by LinearizedHOVecEnv in its stage-callback mode (base.py:214-216)

        # measurement noise: per-env stds (E,n_obs) or shared (n_obs,) (base.py:179-184, 417-424)
        stds = kwargs["observation_stds"]
        self.observation_stds = None
        if stds is not None:
            stds = torch.as_tensor(np.asarray(stds, dtype=np.float64), device=self.device)
            assert stds.shape in [(self.n_observations,), (self.n_envs, self.n_observations)], \
                "``observation_stds`` should have shape (n_observations,) or (n_envs, n_observations)"
            self.observation_stds = stds.contiguous()

        self.seed_value = kwargs["seed"]
        self.seed = kwargs["seed"]
        self._philox_seed = int(np.random.SeedSequence(kwargs["seed"]).generate_state(1, dtype=np.uint64)[0])
        self.rng = kwargs["rng"]
        self.env_offset = int(kwargs["env_offset"])
        self.return_numpy = bool(kwargs["return_numpy"])
        self.host_output_views = bool(kwargs["host_output_views"]) and self.return_numpy
        self.store_trajectory = bool(kwargs["store_trajectory"])

        # data / cache constants (base.py:221-240)
        self.dir_path = dir_prefix + "/" + "_".join(
            [str(t_norm_max), str(t_norm_ssz), str(t_norm_mul), str(list(np.asarray(action_maximums).tolist())),
             str(action_interval)])
        self.file_path_prefix = self.dir_path + "/" + file_prefix
        self.n_data = 1 + self.n_actions + self.n_observations + self.n_properties + 1
        self.average_over = int(kwargs["average_over"])
        self.data_idxs = list(data_idxs)
        self.cache_all_data = kwargs["cache_all_data"]
        self.cache_dump_interval = self.n_envs  # n_envs overrides cache_dump_interval (base.py:1040)
        self.n_envs_total = self.n_envs if kwargs["n_envs_total"] is None else int(kwargs["n_envs_total"])
        self.io = PartCacheWriter(self.file_path_prefix + "_cache", self.cache_dump_interval,
                                  compressed=kwargs["cache_compressed"], env_offset=int(kwargs["env_offset"]),
                                  n_envs_total=self.n_envs_total)
        self.plot = False

        if _SB3VecEnv is not None:
            _SB3VecEnv.__init__(self, num_envs=self.n_envs, observation_space=self.observation_space,
                                action_space=self.action_space)
        self.num_envs = self.n_envs
        self.render_mode = None

        # device buffers (base.py:1063-1084)
        K1, E, no = self.action_interval + 1, self.n_envs, self.n_observations
        self.t_idx, self.t, self.action_idx, self._idx_s = 0, self.numpy_real(0.0), 0, 0
        self.batch_idx = -1
        self.actions = None
        self._raw_actions_ptr = None
        self._t_range = None
        self._States = torch.empty((K1, E, no), dtype=f64, device=self.device)
        self._Observations = torch.empty((K1, E, no), dtype=f64, device=self.device)
        self._Reward = torch.zeros((K1, E), dtype=f64, device=self.device)
        self._Properties = torch.empty((K1, E, self.n_properties), dtype=f64, device=self.device)
        self.States, self.Observations, self.Reward, self.Properties = \
            self._States, self._Observations, self._Reward, self._Properties
        self.Observation_noises = None
        self._x_last = torch.empty((E, no), dtype=f64, device=self.device)
        # obs_last | reward_last | min,max | status word (QRL_STATUS_* bits, int32 in the first half of the double) | pad
        self._step_out = torch.zeros((E * (no + 1) + 4,), dtype=f64, device=self.device)
        self._obs_last = self._step_out[:E * no].view(E, no)
        self._reward_last = self._step_out[E * no:E * (no + 1)]
        self._minmax = self._step_out[E * (no + 1):E * (no + 1) + 2]
        self._minmax_init = torch.tensor([float("inf"), float("-inf")], dtype=f64, device=self.device)
        self._nan = self._step_out[E * (no + 1) + 2:E * (no + 1) + 3].view(torch.int32)   # travels with every step's D2H
        self.nonfinite_detected = False
        # pinned result buffers: two when step() hands out views (the kernel of step i+1 must not write where the caller
        # may still read the results of step i)
        self._h_out_bufs = [torch.zeros(self._step_out.shape, dtype=f64).pin_memory()
                            for _ in range(2 if self.host_output_views else 1)]
        self._h_out_nps = [b.numpy() for b in self._h_out_bufs]
        self._h_status_nps = [b.view(np.int32) for b in self._h_out_nps]
        self._h_views = [(b[:E * no].reshape(E, no), b[E * no:E * (no + 1)]) for b in self._h_out_nps]   # (obs, reward) views
        self._h_sel = 0
        self._h_out, self._h_out_np = self._h_out_bufs[0], self._h_out_nps[0]
        self._h_actions = torch.empty((E, self.n_actions), dtype=f64).pin_memory()
        self._h_actions_np = self._h_actions.numpy()       # NumPy views of the pinned staging buffers (cheap indexing)
        self.rewards = None
        self.data_rewards = []
        self._data_dev = None
        self._data_host = None
        self._sel_idxs = torch.as_tensor(np.asarray(self.data_idxs, dtype=np.int32), device=self.device)
        self._packed = torch.empty((E, K1, self.n_data), dtype=f64, device=self.device) if self.cache_all_data else None
        self._reward_fused = False
        self._cum_fused = False
        # per-env replicas returned by step_wait (base.py:1293 builds ``[done] * n_envs, [{}] * n_envs`` every step)
        self._dones_true, self._dones_false, self._infos = [True] * E, [False] * E, [{}] * E
        self._host_step = False
        self._obs_last_ptr = self._reward_last_ptr = None

    # ``actions`` = float(action) * action_maximums (base.py:1244-1248).  Fast paths hand the RAW actions to the kernels
    # (which scale them in place) and materialise this tensor only when something reads it (user hooks, the pack launch)
    @property
    def actions(self):
        lazy = self.__dict__.get("_actions_lazy")
        if lazy is not None:
            self._actions_val, self._actions_lazy = lazy(), None
        return self.__dict__.get("_actions_val")

    @actions.setter
    def actions(self, value):
        self._actions_val, self._actions_lazy = value, None

    # ---- hooks the interfaced environment implements (base.py:272-317) ------------------------------------------
    def _update_states(self):
        raise NotImplementedError

    def _initial_observations(self, states_0):
        """obs_0 = states_0 + noise[0] (base.py:426)."""
        raise NotImplementedError

    def reset_states(self):
        raise NotImplementedError

    def get_properties(self):
        raise NotImplementedError

    def get_reward(self):
        raise NotImplementedError

    # ---- SB3 replication helpers (base.py:1097-1183) ------------------------------------------------------------
    def _n_indices(self, indices):
        return indices if isinstance(indices, int) else (self.n_envs if indices is None else len(indices))

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in range(self._n_indices(indices))]

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return [getattr(self, method_name)(*method_args, **method_kwargs) for _ in range(self._n_indices(indices))]

    def get_attr(self, attr_name, indices=None):
        return [getattr(self, attr_name) for _ in range(self._n_indices(indices))]

    def set_attr(self, attr_name, value, indices=None):
        return [setattr(self, attr_name, value) for _ in range(self._n_indices(indices))]

    def validate_base(self, shape_reset_states, shape_get_properties, shape_get_reward):
        """Shape checks of the user hooks on the initial state block (base.py:319-376): ``reset_states``, then — with
        ``States`` / ``Observations`` filled with the initial states — ``get_properties`` and ``get_reward`` (skipped when
        the reward is a fused functor of the kernel).  A missing hook raises (the reference prints and exits)."""
        states_0 = self.backend.convert_to_typed(tensor=self.reset_states(), dtype="real")
        assert tuple(states_0.shape) == tuple(shape_reset_states), \
            f"``reset_states`` should return an array with shape ``{tuple(shape_reset_states)}``"
        self._States[:] = states_0.unsqueeze(0)
        self._Observations[:] = self._States
        self.States, self.Observations = self._States, self._Observations
        if self.n_properties > 0:
            self.Properties = self.backend.convert_to_typed(tensor=self.get_properties(), dtype="real")
            assert tuple(self.Properties.shape) == tuple(shape_get_properties), \
                f"``get_properties`` should return an array with shape ``{tuple(shape_get_properties)}``"
        if not self._reward_fused:
            self.Reward = self.backend.convert_to_typed(tensor=self.get_reward(), dtype="real")
            assert tuple(self.Reward.shape) == tuple(shape_get_reward), \
                f"``get_reward`` should return an array with shape ``{tuple(shape_get_reward)}``"

    def validate(self):
        """``validate_base`` with the shapes of a vectorized env (base.py:1088-1095)."""
        K1, E = self.action_interval + 1, self.n_envs
        return self.validate_base(shape_reset_states=(E, self.n_observations),
                                  shape_get_properties=(K1, E, self.n_properties), shape_get_reward=(K1, E))

    # ---- reset (base.py:378-438, 1185-1231) ------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        obs = self._reset()
        return self._out(obs)

    def _reset(self):
        self.batch_idx += 1
        # zeroed in place: the buffer address is baked into captured CUDA graphs (use_cuda_graph)
        if self.rewards is None:
            self.rewards = torch.zeros((self.n_envs,), dtype=torch.float64, device=self.device)
        else:
            self.rewards.zero_()
        n_sel = len(self.data_idxs)
        if self._data_dev is None:
            # row pitch rounded up to an even number of doubles: every env's run of rows then starts 16-byte aligned
            # and the integration kernel can write it with TMA bulk stores; ``data`` drops the padding column
            self._data_pitch = n_sel + (n_sel & 1)
            self._data_dev = torch.zeros((self.n_envs, self.shape_T[0], self._data_pitch), dtype=torch.float64,
                                         device=self.device)
            self._data_stale_rows = 0
        else:
            # The reference allocates a zeroed cache every episode (base.py:1226).  Here the device-resident cache is NOT
            # re-zeroed (655 MB at config 3 = ~110 us of fills per reset, 1 us per step on average): an episode overwrites
            # rows 0 .. t_idx, and rows beyond that may hold an earlier episode's values up to the high-water mark kept here;
            # ``data`` — the only reader — zeroes exactly those rows in the host copy it hands out.
            self._data_stale_rows = max(self._data_stale_rows, self._data_rows_written())
        self._data_host = None
        self.t_idx, self.t, self.action_idx = 0, self.numpy_real(0.0), 0
        states_0 = self.backend.convert_to_typed(tensor=self.reset_states(), dtype="real").contiguous()
        self._x_last.copy_(states_0)
        self._States[:] = states_0.unsqueeze(0)
        self.States = self._States
        self._minmax.copy_(self._minmax_init, non_blocking=True)
        observations_0 = self._initial_observations(states_0)
        self._Observations[:] = observations_0.unsqueeze(0)
        self.Observations = self._Observations
        return observations_0

    # ---- one action interval (base.py:440-483, 1233-1293) --------------------------------------------------------
    def step_async(self, actions):
        if isinstance(actions, torch.Tensor):
            a = actions.to(device=self.device, dtype=torch.float64)
        else:
            # host actions: one NumPy copy into pinned memory, then an asynchronous H2D on the launch stream
            self._h_actions_np[...] = np.asarray(actions, dtype=np.float64).reshape(self.n_envs, self.n_actions)
            if self.host_output_views:      # this step's results go to the other pinned buffer
                self._h_sel ^= 1
                self._h_out, self._h_out_np = self._h_out_bufs[self._h_sel], self._h_out_nps[self._h_sel]
            if getattr(self, "use_cuda_graph", False):
                self._host_step = True      # H2D of the actions and D2H of the results are nodes of the captured graph
                return
            if self._host_actions_in_place():
                self._host_step = True      # the integration launch reads the pinned actions itself (zero-copy)
                return
            a = self._h_actions
        self._host_step = False
        self._set_actions(a)

    def _set_actions(self, a):
        """``actions = float(action) * action_maximums`` (base.py:1244-1248).  With ``use_cuda_graph`` the raw actions
        land in the graph's static input buffer and the scaling is part of the captured graph."""
        if getattr(self, "use_cuda_graph", False) or getattr(self, "fused_coefficients", False):
            if self._direct_actions(a):
                return      # single-kernel step: the launch reads the caller's tensor in place (no staging copy)
            self._actions_in.copy_(a.reshape(self.n_envs, self.n_actions), non_blocking=True)
            if not getattr(self, "use_cuda_graph", False):
                self.actions = self._actions_in * self.action_maximums_batch
        else:
            a = a.to(self.device, non_blocking=True) if not a.is_cuda else a
            self.actions = (a.reshape(self.n_envs, self.n_actions) * self.action_maximums_batch).contiguous()

    def _direct_actions(self, a):
        """Hook: environments whose captured step is a single native kernel may consume the actions tensor in place."""
        return False

    def _host_actions_in_place(self):
        """Hook: True when the env's integration launch reads ``_h_actions`` (pinned host memory) directly."""
        return False

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def update(self, reset_minmax=True):
        S1 = self.shape_T[0]
        K = self.action_interval
        dim = K + 1 if self.t_idx + K < S1 else S1 - self.t_idx
        self._dim = dim
        self._t_range = (self.t_idx, self.t_idx + dim)     # T_step / _T_step_host are views built on demand

        if reset_minmax:
            self._minmax.copy_(self._minmax_init, non_blocking=True)
        self.States = self._update_states()        # ONE fused launch: States, Observations (+Reward, rewards, min/max)
        self.Observations = self._Observations if dim == K + 1 else self._Observations[:dim]
        if self.n_properties > 0:
            self.Properties = self.backend.convert_to_typed(tensor=self.get_properties(), dtype="real")
        if not self._reward_fused:
            self.Reward = self.backend.convert_to_typed(tensor=self.get_reward(), dtype="real")
            self._reward_last.copy_(self.Reward[dim - 1])
        else:
            self.Reward = self._Reward if dim == K + 1 else self._Reward[:dim]

        self.t_idx += dim - 1
        self.t = self._T_host[self.t_idx]
        self.action_idx += 1
        terminated = not self.t_idx + 1 < S1
        return self._obs_last, self._reward_last, terminated

    @property
    def T_step(self):
        """``T[t_idx : t_idx+K+1]`` of the current action interval (base.py:454-458), device tensor."""
        r = self._t_range
        return None if r is None else self.T[r[0]:r[1]]

    @property
    def _T_step_host(self):
        r = self._t_range
        return None if r is None else self._T_host[r[0]:r[1]]

    def update_data(self):
        """Packed rows ``[t, actions, obs, props, cumulative reward]`` (base.py:1295-1380) — one pack-kernel launch."""
        dim = self._dim
        if not self._cum_fused:
            self.rewards += self._reward_last
        if not self.store_trajectory:
            return  # learner-only mode: no (K+1,E,n) blocks exist, nothing to pack
        if getattr(self, "_packed_by_kernel", False):
            self._data_host = None
            return  # the integration kernel already wrote this interval's cache rows
        a = self.__dict__.get("_pack_args")
        if a is None:       # the constant part of the argument block is built once
            a = self._pack_args = N.PackArgs()
            a.n_envs, a.n_actions, a.n_obs, a.n_props = self.n_envs, self.n_actions, self.n_observations, self.n_properties
            a.n_sel = len(self.data_idxs)
            a.observations = N.ptr(self._Observations)
            a.rewards_cum = N.ptr(self.rewards)
            a.idxs = N.ptr(self._sel_idxs)
            if self._packed is not None:
                a.packed, a.packed_stride_e = N.ptr(self._packed), (self.action_interval + 1) * self.n_data
            a.selected_stride_e = self.shape_T[0] * self._data_pitch
            a.selected_row_pitch = self._data_pitch
        a.dim = dim
        a.T_step = N.ptr(self.T) + 8 * (self.t_idx - dim + 1)
        raw = self.__dict__.get("_raw_actions_ptr")
        if raw is not None and self.__dict__.get("_actions_lazy") is not None:
            a.actions, a.action_scale = raw, N.ptr(self.action_maximums)     # raw actions, scaled by the pack kernel
        else:
            a.actions, a.action_scale = N.ptr(self.actions), None
        a.properties = N.ptr(self.Properties.contiguous()) if self.n_properties > 0 else None
        # rows t_idx-dim+1 .. t_idx of the (E, shape_T, n_sel) cache; row 0 of a block overwrites the last row of
        # the previous one exactly as in the reference (base.py:1370-1372)
        a.selected = N.ptr(self._data_dev) + 8 * (self.t_idx - dim + 1) * self._data_pitch
        N.pack_rows(a)
        self._data_host = None
        if self.cache_all_data:
            part = self._packed[:, :dim].cpu().numpy()
            self.io.dump_part_async(data=part, batch_idx=self.batch_idx, part_idx=self.action_idx - 1)

    def _fetch_step_outputs(self):
        """One D2H of [obs_last | reward_last | min,max] into pinned memory, then a stream sync.  An env whose launch has
        already delivered the results (zero-copy) leaves an event in ``_host_event``: the host then waits for that point of
        the stream only — launches enqueued after it (cache rows) do not hold the step up."""
        ev = self.__dict__.pop("_host_event", None)
        if ev is not None:
            ev.synchronize()
            return self._h_out
        self._h_out.copy_(self._step_out, non_blocking=True)
        raw = N.current_raw_stream()
        if raw != getattr(self, "_sync_raw", None):     # the Stream object is looked up again only when the stream changes
            self._sync_raw, self._sync_stream = raw, torch.cuda.current_stream()
        self._sync_stream.synchronize()
        return self._h_out

    def check_truncation(self, host=None):
        """``max(Observations) > range[1] or min(Observations) < range[0]`` over the whole block (base.py:485-499)."""
        if host is None:
            self._fetch_step_outputs()
        E, no = self.n_envs, self.n_observations
        self._check_status(int(self._h_status_nps[self._h_sel][2 * (E * (no + 1) + 2)]))
        lo, hi = self._h_out_np[E * (no + 1)], self._h_out_np[E * (no + 1) + 1]
        return bool(hi > self.observation_space_range[1] or lo < self.observation_space_range[0])

    def _check_status(self, status):
        """The kernels' status word, read wherever the host has just synchronised anyway.  A fused cross-GPU barrier that
        timed out is fatal (the gathered step results are stale or partial); a non-finite state is reported once — the
        reference carries on silently there (its truncation check is NaN-blind, base.py:495-499)."""
        if status & N.STATUS_PEER_TIMEOUT:
            raise RuntimeError("quantrl_b200: the fused cross-GPU barrier timed out (a peer rank never arrived); the "
                               "gathered observations / rewards of this step are not valid")
        if (status & N.STATUS_NONFINITE) and not self.nonfinite_detected:
            self.nonfinite_detected = True
            warnings.warn(f"quantrl_b200: non-finite states detected in batch #{self.batch_idx} (NaN / Inf); "
                          "``env.nonfinite_detected`` is set", RuntimeWarning, stacklevel=3)

    def step_wait(self):
        host_step = getattr(self, "_host_step", False)
        observations, reward, terminated = self.update(reset_minmax=not host_step)
        self.update_data()
        hook = getattr(self, "_pre_sync_hook", None)
        if hook is not None:        # e.g. the learner's gather wait of the sharded path: enqueued behind the step, covered by
            hook()                  # the one stream synchronisation below
        if host_step and getattr(self, "_host_graph_done", False):
            self._replay_stream.synchronize()     # the replayed graph already delivered the results to pinned memory
            host = self._h_out
            self._host_graph_done = False
        else:
            host = self._fetch_step_outputs()
        return self._finish_host_step(host, observations, reward, terminated)

    def _finish_host_step(self, host, observations, reward, terminated):
        """Tail of ``step_wait`` once this step's results are in pinned memory: truncation check, the returned arrays, the
        automatic reset (base.py:1262-1293)."""
        truncated = self.check_truncation(host)
        if truncated:
            print(f"Batch #{self.batch_idx} truncated")
        E, no = self.n_envs, self.n_observations
        if self.host_output_views:
            obs_out, reward_out = self._h_views[self._h_sel]
        elif self.return_numpy:
            reward_out = self._h_out_np[E * no:E * (no + 1)].copy()
            obs_out = self._h_out_np[:E * no].reshape(E, no).copy()
        else:
            reward_out, obs_out = reward.clone(), observations.clone()
        done = bool(terminated or truncated)
        if done:
            self.data_rewards.append(self.rewards.clone())
            obs_reset = self._reset()
            obs_out = self._out(obs_reset)
        return obs_out, reward_out, (self._dones_true if done else self._dones_false), self._infos

    def step_device(self, actions):
        """Device-resident step for on-GPU learners: same work as ``step`` (integration, observations, reward,
        cumulative reward, packed cache rows) but no host round trip — actions are a CUDA tensor, the returned
        ``(observations, reward)`` are views of device buffers that the next step overwrites, and the truncation
        reduction stays on the device (``truncation_pending()`` reads it).  Returns ``(obs, reward, terminated)``."""
        self._host_step = False     # (a preceding step(host actions) leaves it set: this step reads ``actions``, not the pinned buffer)
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float64), device=self.device)
        self._set_actions(actions)
        observations, reward, terminated = self.update(reset_minmax=False)  # min/max accumulate over the episode
        self.update_data()
        if terminated:
            self.data_rewards.append(self.rewards.clone())
        return observations, reward, terminated

    def truncation_pending(self):
        return self.check_truncation()

    def attach_step_output(self, obs_ptr, reward_ptr, obs_view=None, reward_view=None):
        """Redirect the per-step outputs (last observations, last reward) of the integration kernel to caller-owned
        device memory — e.g. the learner rank's buffer mapped over NVLink (``distributed.PeerGather``).  ``obs_ptr`` /
        ``reward_ptr`` are raw device addresses; the optional views replace the local tensors returned by
        ``step_device`` (ranks whose outputs live on another GPU keep returning their stale local buffers)."""
        self._obs_last_ptr, self._reward_last_ptr = int(obs_ptr), int(reward_ptr)
        if obs_view is not None:
            self._obs_last = obs_view
        if reward_view is not None:
            self._reward_last = reward_view
        self._graph = self._graph_host = self._graph_host1 = None
        if hasattr(self, "_static"):
            self._static = {}

    def _out(self, t):
        return t.detach().cpu().numpy() if self.return_numpy else t

    @property
    def data(self):
        """Host view of the selected-column cache ``(n_envs, shape_T, len(data_idxs))`` (base.py:1226, 1372)."""
        if self._data_host is None and self._data_dev is not None:
            host = self._data_dev[:, :, :len(self.data_idxs)].cpu().numpy()
            lo, hi = self._data_rows_written(), self._data_stale_rows
            if hi > lo:
                host[:, lo:hi] = 0.0        # rows this episode has not reached (they hold an earlier episode's values)
            self._data_host = host
        return self._data_host

    def _data_rows_written(self):
        """Leading rows of the episode cache written by the current episode (0 before its first action step)."""
        return self.t_idx + 1 if (self.action_idx > 0 and self.store_trajectory) else 0

    def evolve(self, show_progress=False, close=True, save=False):
        """Free evolution with ``actions = action_maximums`` (base.py:1382-1449)."""
        self.reset()
        for _ in range(self.action_steps):
            self._set_actions(torch.ones((self.n_envs, self.n_actions), dtype=torch.float64, device=self.device))
            self.update()
            self.update_data()
            if self.check_truncation():
                print("Batch truncated")
                break
        self.data_rewards.append(self.rewards.clone())
        if close:
            self.close(save=save)

    def close(self, save=False, axis_args=None):
        if save:
            os.makedirs(self.dir_path, exist_ok=True)
            rewards = self.backend.convert_to_numpy(tensor=self.data_rewards, dtype="real").ravel()
            np.savez_compressed(self.file_path_prefix + "_" + "_".join(
                ["learning_curve", str(self._idx_s), str(self._idx_s + rewards.shape[0] - 1)]) + ".npz", rewards)
        self.io.close(dump_cache=True)


def default_backend(backend_library="b200", backend_precision="double", backend_device="gpu"):
    assert "b200" in backend_library, "parameter ``backend_library`` should be ``'b200'``"
    return get_backend_instance(library=backend_library, precision=backend_precision, device=backend_device)
