"""TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy) of the reference's stochastic hot path.

Never imported by the product package ``quantrl_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it, and only as the checker / the thing timed as
"the CPU path".

Restates, per reference file:line (paths relative to the reference root):

* time grid ............ quantrl/envs/base.py:157-168, 212   (``time_grid``)
* Wiener increments .... quantrl/envs/stochastic.py:157-170  (``Ws = sqrt(t_ssz) * N(0,1)``, shape (S, n))
* measurement noise .... quantrl/envs/base.py:408-426        (``N(0,1) * stds``, shape (S+1, n); obs_0)
* EM recurrence ........ quantrl/envs/stochastic.py:178-242  (``Y[i+1] = (I + A dt) @ Y[i] + n * Ws[t_idx+i]``)
* interval bookkeeping . quantrl/envs/base.py:440-483        (T_step slice, Observations = States + noises,
                                                              t_idx advance, terminated)
* packed cache rows .... quantrl/envs/base.py:1295-1372      (``[t, actions, obs, props, cumulative reward]``)

Parity status: PINNED against the reference itself — ``oracle/make_golden.py`` runs the unmodified
``quantrl.envs.stochastic.LinearEnv`` (through ``oracle/ref_shim.py``) and stores its ``Ws``,
``Observation_noises``, trajectories and rewards in ``tests/golden/em_linear_env.npz``;
``tests/test_oracle_golden.py`` replays them through this file.  (The reference has no tests or golden vectors of
its own — SURVEY.md §4.)

The in-kernel RNG of the CUDA path (Philox4x32-10 + Box–Muller, ``quantrl_b200/csrc/philox.cuh``) has no
reference counterpart (the reference draws from NumPy PCG64); ``philox_normals`` restates *our* published stream
definition so that Philox-mode trajectories can also be compared pathwise.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------------------
# time grid (quantrl/envs/base.py:157-168, 212)
# ------------------------------------------------------------------------------------------------------------
def time_grid(t_norm_max, t_norm_ssz, t_norm_mul, action_interval):
    shape_T = int(np.int64(t_norm_max / t_norm_ssz)) + 1
    T_norm = np.arange(shape_T, dtype=np.float64) * t_norm_ssz
    T = T_norm * t_norm_mul
    t_ssz = t_norm_ssz * t_norm_mul
    action_steps = int(np.ceil((shape_T - 1) / action_interval))
    return shape_T, T, t_ssz, action_steps


# ------------------------------------------------------------------------------------------------------------
# faithful single-env port (what the reference executes: a Python ``for`` over sub-steps, backends/numpy.py:178-186)
# ------------------------------------------------------------------------------------------------------------
def linear_env_interval(Y0, t_idx, K, get_A, get_noise_prefixes, args, Ws, t_ssz):
    """One action interval of ``LinearEnv._update_states`` for ONE env.  Y0 (n,), Ws (S,n) -> States (K+1,n)."""
    n = Y0.shape[0]
    I = np.eye(n)
    Y = np.empty((K + 1, n))
    Y[0] = Y0
    for i in range(K):
        M_i = I + get_A(t_idx + i, args) * t_ssz
        n_i = get_noise_prefixes(t_idx + i, args)
        Y[i + 1] = np.matmul(M_i, Y[i]) + n_i * Ws[t_idx + i]
    return Y


def linear_env_episode(x0, actions_per_step, K, S, get_A, get_noise_prefixes, Ws, obs_noises, t_ssz,
                       reward_fn=None):
    """Whole episode for ONE env, lock-step bookkeeping of ``BaseEnv.update`` (base.py:440-483).

    actions_per_step (n_steps, n_act) already multiplied by ``action_maximums``.
    Returns States (S+1,n), Observations (S+1,n), Reward (S+1,) (row 0 of Reward is 0).
    """
    n = x0.shape[0]
    States = np.empty((S + 1, n))
    Obs = np.empty((S + 1, n))
    Rew = np.zeros(S + 1)
    States[0] = x0
    Obs[0] = x0 + (obs_noises[0] if obs_noises is not None else 0.0)
    t_idx = 0
    for a in range(actions_per_step.shape[0]):
        dim = min(K, S - t_idx) + 1
        args = (actions_per_step[a], None, None)
        Y = linear_env_interval(States[t_idx], t_idx, dim - 1, get_A, get_noise_prefixes, args, Ws, t_ssz)
        O = Y + (obs_noises[t_idx:t_idx + dim] if obs_noises is not None else 0.0)
        States[t_idx:t_idx + dim] = Y
        Obs[t_idx:t_idx + dim] = O
        if reward_fn is not None:
            Rew[t_idx:t_idx + dim] = reward_fn(Y, O, actions_per_step[a])
        t_idx += dim - 1
        if not t_idx + 1 < S + 1:
            break
    return States, Obs, Rew


# ------------------------------------------------------------------------------------------------------------
# env-vectorized restatement (same recurrence for E independent envs; used as the parity checker)
# ------------------------------------------------------------------------------------------------------------
def em_interval_vec(x0, A, nz, W, t_ssz):
    """x0 (E,n); A (E,n,n) or (K,E,n,n); nz (E,n) or (K,E,n); W (K,E,n) *already scaled by sqrt(t_ssz)*.

    Returns States (K+1,E,n) with row 0 = x0 (stochastic.py:180-184, 218-242).
    """
    K = W.shape[0]
    E, n = x0.shape
    I = np.eye(n)
    Y = np.empty((K + 1, E, n))
    Y[0] = x0
    for i in range(K):
        A_i = A[i] if A.ndim == 4 else A
        n_i = nz[i] if nz.ndim == 3 else nz
        M_i = I + A_i * t_ssz
        Y[i + 1] = np.einsum("eij,ej->ei", M_i, Y[i]) + n_i * W[i]
    return Y


def reward_neg_weighted_sq(States_or_Obs, weights):
    """Canonical EM reward functor: r[t,e] = -sum_c w_c * x[t,e,c]^2."""
    return -np.einsum("tec,c->te", States_or_Obs * States_or_Obs, np.asarray(weights, dtype=np.float64))


def pack_rows(T_step, actions, Obs, Props, rewards_cum):
    """``BaseSB3Env.update_data`` row layout (base.py:1314-1356): (E, dim, n_data).

    T_step (dim,), actions (E,n_act), Obs (dim,E,n_obs), Props (dim,E,n_props) or None, rewards_cum (E,).
    """
    dim = T_step.shape[0]
    E = actions.shape[0]
    cols = [np.broadcast_to(T_step.reshape(1, dim, 1), (E, dim, 1)),
            np.broadcast_to(actions.reshape(E, 1, -1), (E, dim, actions.shape[1])),
            np.transpose(Obs[:dim], (1, 0, 2))]
    if Props is not None and Props.shape[-1] > 0:
        cols.append(np.transpose(Props[:dim], (1, 0, 2)))
    cols.append(np.broadcast_to(rewards_cum.reshape(E, 1, 1), (E, dim, 1)))
    return np.concatenate(cols, axis=2)


# ------------------------------------------------------------------------------------------------------------
# Philox4x32-10 + Box–Muller: restatement of OUR stream definition (include/quantrl_b200.h, "RNG stream")
# ------------------------------------------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint64(0x9E3779B9)
_W1 = np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  All inputs broadcastable uint32-valued arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(k0) & _MASK
    k1 = np.uint64(k1) & _MASK
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def _bits_to_double(hi20, lo32):
    """double with exponent 0 (value in [1,2)) from a 20-bit high and a 32-bit low mantissa word."""
    bits = (np.uint64(0x3FF) << np.uint64(52)) | (hi20 << np.uint64(32)) | lo32
    return bits.view(np.float64)


def philox_normals(seed, episode, tau, env, comp):
    """Two N(0,1) per counter (tau, episode, env, comp), key = (seed lo, seed hi): the stream of
    include/quantrl_b200.h ("RNG stream"), evaluated with NumPy's libm in extended precision (np.longdouble: 64-bit
    mantissa on x86) instead of the kernel's table + polynomial kernels, then rounded to double.

      u1 = 2 - bits(0x3FF<<52 | x1<<20 | x0>>12 | 1)   in (0,1);   radius = sqrt(-2 ln u1)
      j = x3 >> 24 (8 bits);  frac = ((x3 & 0xFFFFFF) << 28 | x2 >> 4) / 2^52;  theta = (j + frac) * 2 pi / 256
      z0 = radius * sin theta,  z1 = radius * cos theta
    """
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    x0, x1, x2, x3 = philox4x32_10(tau, episode, env, comp, seed & 0xFFFFFFFF, seed >> 32)
    u64, ld = np.uint64, np.longdouble
    d1 = _bits_to_double(x1 >> u64(12), ((x1 << u64(20)) & _MASK) | (x0 >> u64(12)) | u64(1))
    radius = np.sqrt(-2 * np.log((2.0 - d1).astype(ld)))
    frac = _bits_to_double((x3 >> u64(4)) & u64(0xFFFFF), ((x3 << u64(28)) & _MASK) | (x2 >> u64(4))) - 1.0
    theta = ((x3 >> u64(24)).astype(ld) + frac.astype(ld)) * (ld("3.14159265358979323846264338327950288") / 128)
    return (radius * np.sin(theta)).astype(np.float64), (radius * np.cos(theta)).astype(np.float64)


def philox_noise_block(seed, episode, t_idx, K, env0, E, n):
    """(z_w (K,E,n), z_obs (K+1,E,n)) for the interval starting at global time index ``t_idx``.

    Stochastic-path convention: the call with tau = k gives z0 -> Wiener increment of sub-step k -> k+1 and
    z1 -> measurement noise of time index k+1; the measurement noise of time index 0 is z1 of tau = 0xFFFFFFFF."""
    env = (np.uint64(env0) + np.arange(E, dtype=np.uint64)).reshape(1, -1, 1)
    comp = np.arange(n, dtype=np.uint64).reshape(1, 1, -1)
    first = np.uint64(t_idx - 1) if t_idx > 0 else np.uint64(0xFFFFFFFF)
    tau = np.concatenate([[first], np.arange(t_idx, t_idx + K, dtype=np.uint64)]).reshape(-1, 1, 1)
    z0, z1 = philox_normals(seed, episode, tau, env, comp)
    return z0[1:], z1


def philox_obs_noise_ode(seed, episode, t_idx, K, env0, E, d):
    """Deterministic-path measurement noise (K+1,E,d): component c of time index t = z[c & 1] of the call with
    counter (t, episode, env, 0x10000 + c // 2)."""
    tau = np.arange(t_idx, t_idx + K + 1, dtype=np.uint64).reshape(-1, 1, 1)
    env = (np.uint64(env0) + np.arange(E, dtype=np.uint64)).reshape(1, -1, 1)
    comp = np.arange(d, dtype=np.uint64).reshape(1, 1, -1)
    z0, z1 = philox_normals(seed, episode, tau, env, np.uint64(0x10000) + comp // np.uint64(2))
    return np.where(comp % np.uint64(2) == 1, z1, z0)
