"""TEST INFRASTRUCTURE ONLY — builds and binds oracle/em_port.c (plain-C restatement of the stochastic hot path).

Used by tests/test_oracle_c_port.py (second implementation of the oracle, checked against the NumPy one and fixture
G3) and by tools/cpu_compiled_em.py (compiled, threaded CPU comparator).  Never imported by the product."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "em_port.c")
LIB = os.path.join(HERE, "_build", "libem_port.so")
LYAP_SRC = os.path.join(HERE, "lyap_port.c")
LYAP_LIB = os.path.join(HERE, "_build", "liblyap_port.so")
PHILOX_SRC = os.path.join(HERE, "philox_port.c")
PHILOX_LIB = os.path.join(HERE, "_build", "libphilox_port.so")
_lib = None
_lyap = None


def available():
    return shutil.which("gcc") is not None


def usable():
    """gcc is present AND the three C files compile here (OpenMP runtime, -march level): what the tests skip on."""
    if not available():
        return False
    try:
        build()
        return True
    except (subprocess.CalledProcessError, OSError):
        return False


def build(force=False):
    """Compile the C restatements (oracle/em_port.c, lyap_port.c, philox_port.c) into oracle/_build/; returns the EM library."""
    for src, out in ((SRC, LIB), (LYAP_SRC, LYAP_LIB), (PHILOX_SRC, PHILOX_LIB)):
        newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "rng_tables.h")))
        if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", out,
                            "-lm"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp, i32, f64, u64 = C.POINTER(C.c_double), C.c_int, C.c_double, C.c_uint64
        _lib.em_port_fed.argtypes = [i32, i32, i32, f64] + [dp] * 9
        _lib.em_port_fed.restype = None
        _lib.em_port_rng.argtypes = [i32, i32, i32, f64, u64, u64] + [dp] * 9
        _lib.em_port_rng.restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else np.ascontiguousarray(np.broadcast_to(a, shape))


def em_interval_fed(x0, A, nz, W, dt, obs_noise=None, weights=None):
    """x0 (E,n); A (E,n,n) or (n,n); nz (E,n) or (n,); W (K,E,n) scaled by sqrt(dt); obs_noise (K+1,E,n) or None.
    Returns States, Observations (K+1,E,n) and Reward (K+1,E) = -sum_c w_c obs_c^2."""
    E, n = x0.shape
    K = W.shape[0]
    x0, A, nz, W = _c(x0), _c(A, (E, n, n)), _c(nz, (E, n)), _c(W)
    ON, w = _c(obs_noise), _c(weights)
    St, Ob, Rw = np.empty((K + 1, E, n)), np.empty((K + 1, E, n)), np.empty((K + 1, E))
    lib().em_port_fed(E, n, K, float(dt), _p(x0), _p(A), _p(nz), _p(W), _p(ON), _p(w), _p(St), _p(Ob), _p(Rw))
    return St, Ob, Rw


class RngStepper:
    """Preallocated buffers for repeated ``em_port_rng`` intervals (the timing loop of tools/cpu_compiled_em.py)."""

    def __init__(self, E, n, K, dt, A, nz, stds, weights, seed=1):
        self.E, self.n, self.K, self.dt, self.seed = E, n, K, float(dt), int(seed)
        self.A, self.nz, self.stds, self.w = _c(A, (E, n, n)), _c(nz, (E, n)), _c(stds, (n,)), _c(weights, (n,))
        self.St, self.Ob, self.Rw = np.empty((K + 1, E, n)), np.empty((K + 1, E, n)), np.empty((K + 1, E))
        self.x = np.ones((E, n))
        self.interval = 0

    def step(self):
        lib().em_port_rng(self.E, self.n, self.K, self.dt, self.seed, self.interval, _p(self.x), _p(self.A), _p(self.nz),
                          _p(self.stds), _p(self.w), _p(self.St), _p(self.Ob), _p(self.Rw), _p(self.x))
        self.interval += 1
        return self.Ob[-1], self.Rw[-1]


def lyap_rk4_interval(params, actions, y0, K, t_ssz, substeps=1):
    """oracle/lyap_port.c: optomech RK4 interval.  params (E,7), actions (E,), y0 (E,20) -> Y (K+1,E,20)."""
    global _lyap
    if _lyap is None:
        build()
        _lyap = C.CDLL(LYAP_LIB)
        _lyap.lyap_port_rk4.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double] + [C.POINTER(C.c_double)] * 4
        _lyap.lyap_port_rk4.restype = None
    E = y0.shape[0]
    params, actions, y0 = _c(params, (E, 7)), _c(actions, (E,)), _c(y0)
    Y = np.empty((K + 1, E, 20))
    _lyap.lyap_port_rk4(E, K, int(substeps), float(t_ssz), _p(params), _p(actions), _p(y0), _p(Y))
    return Y


_philox = None


def _philox_lib():
    global _philox
    if _philox is None:
        build()
        _philox = C.CDLL(PHILOX_LIB)
        u32p = C.POINTER(C.c_uint32)
        _philox.philox_port_4x32_10.argtypes = [u32p, C.c_uint32, C.c_uint32, u32p]
        _philox.philox_port_4x32_10.restype = None
        _philox.philox_port_normals.argtypes = [C.c_uint64, C.c_uint32, u32p, u32p, u32p, C.c_int] + [C.POINTER(C.c_double)] * 2
        _philox.philox_port_normals.restype = None
    return _philox


def philox4x32_10(counter, key):
    """oracle/philox_port.c: one Philox4x32-10 block; counter (4 words), key (2 words) -> 4 words."""
    c = (C.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in counter])
    out = (C.c_uint32 * 4)()
    _philox_lib().philox_port_4x32_10(c, int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF, out)
    return [int(v) for v in out]


def philox_normals_poly(seed, episode, tau, env, comp):
    """The kernel's generator (polynomial ln / sin / cos, Goldschmidt sqrt) in C for flat uint32 arrays of counters."""
    tau, env, comp = (np.ascontiguousarray(np.broadcast_arrays(tau, env, comp)[i], dtype=np.uint32).ravel() for i in range(3))
    z0, z1 = np.empty(tau.size), np.empty(tau.size)
    u32p = C.POINTER(C.c_uint32)
    _philox_lib().philox_port_normals(int(seed) & 0xFFFFFFFFFFFFFFFF, int(episode) & 0xFFFFFFFF, tau.ctypes.data_as(u32p),
                                      env.ctypes.data_as(u32p), comp.ctypes.data_as(u32p), tau.size, _p(z0), _p(z1))
    return z0, z1
