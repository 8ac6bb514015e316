"""ctypes binding of ``include/quantrl_b200.h`` (the C ABI of the CUDA hot path).

There is deliberately NO fallback: if the shared library is missing or a call fails, a ``NativeError`` is raised.
"""
import ctypes as C
import os

from .build import LIB_PATH

ABI_VERSION = 5

RNG_FED, RNG_PHILOX = 0, 1
STATUS_NONFINITE, STATUS_PEER_TIMEOUT = 1, 2
REWARD_NONE, REWARD_NEG_WSQ_OBS, REWARD_NEG_WSQ_STATE, REWARD_ENTAN_LN_OBS, REWARD_ENTAN_LN_STATE = 0, 1, 2, 3, 4
REWARD_SYNC_C_OBS, REWARD_SYNC_C_STATE, REWARD_SYNC_P_OBS, REWARD_SYNC_P_STATE = 5, 6, 7, 8
SYS_OPTOMECH, SYS_KERR_CHAIN, SYS_CONST_AD = 1, 2, 3
ODE_RK4, ODE_DOPRI5, ODE_BOSH3, ODE_TSIT5, ODE_HEUN21, ODE_FEHLBERG2, ODE_DOP853, ODE_EULER, ODE_MIDPOINT = range(9)
# method names of the solver plugin -> (id, adaptive); aliases follow the reference's torchdiffeq / diffrax / SciPy names
ODE_METHODS = {"rk4": (ODE_RK4, False), "dopri5": (ODE_DOPRI5, True), "bosh3": (ODE_BOSH3, True), "RK23": (ODE_BOSH3, True),
               "tsit5": (ODE_TSIT5, True), "adaptive_heun": (ODE_HEUN21, True), "heun": (ODE_HEUN21, True),
               "fehlberg2": (ODE_FEHLBERG2, True), "dop853": (ODE_DOP853, True), "DOP853": (ODE_DOP853, True),
               "dopri8": (ODE_DOP853, True), "euler": (ODE_EULER, False), "midpoint": (ODE_MIDPOINT, False)}
EM_FORCE_GENERIC, EM_FORCE_TILED, EM_ITEM_OUTPUT, EM_PACK_TMA, EM_PDL, EM_HOST_TIDX = 1, 16, 32, 128, 256, 64
EM_PEER_PUBLISH = 8192
ODE_FORCE_THREAD, ODE_FORCE_LANES, ODE_LANES_SINGLE = 1, 2, 4

_dp = C.c_void_p  # device pointers travel as integers


class NativeError(RuntimeError):
    pass


class EmArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_envs", C.c_int32), ("K", C.c_int32), ("rng_mode", C.c_int32), ("t_ssz", C.c_double),
        ("A", _dp), ("A_stride_k", C.c_int64), ("A_stride_e", C.c_int64),
        ("nz", _dp), ("nz_stride_k", C.c_int64), ("nz_stride_e", C.c_int64),
        ("W", _dp), ("obs_noise", _dp),
        ("seed", C.c_uint64), ("episode", C.c_uint32), ("_pad0", C.c_uint32), ("t_idx", C.c_int64),
        ("env_offset", C.c_int64), ("obs_std", _dp), ("obs_std_stride_e", C.c_int64),
        ("x0", _dp), ("states", _dp), ("observations", _dp), ("x_last", _dp), ("obs_last", _dp),
        ("reward_kind", C.c_int32), ("flags", C.c_int32), ("reward_w", _dp), ("reward", _dp), ("reward_last", _dp),
        ("rewards_cum", _dp), ("obs_minmax", _dp), ("nan_flag", _dp),
        ("packed", _dp), ("packed_stride_e", C.c_int64), ("packed_row_pitch", C.c_int64), ("T_step", _dp), ("actions", _dp), ("sel_idxs", _dp),
        ("n_actions", C.c_int32), ("n_sel", C.c_int32), ("dyn", _dp), ("A_act", _dp), ("action_scale", _dp),
        ("finish_counter", _dp), ("minmax_out", _dp), ("dyn_advance", C.c_int64),
        ("peer_flags", _dp), ("peer_epoch", _dp), ("peer_rank", C.c_int32), ("peer_world", C.c_int32),
        ("peer_publish_value", C.c_uint64), ("peer_pulled", _dp), ("peer_pulled_min", C.c_uint64),
        ("peer_dst", C.c_int32), ("_pad1", C.c_int32), ("obs_last_b", _dp), ("reward_last_b", _dp),
        ("done_flag", _dp), ("done_value", C.c_uint64),
    ]


class OdeArgs(C.Structure):
    _fields_ = [
        ("system", C.c_int32), ("num_modes", C.c_int32), ("num_quads", C.c_int32), ("n_envs", C.c_int32),
        ("K", C.c_int32), ("method", C.c_int32), ("substeps", C.c_int32), ("n_actions", C.c_int32),
        ("t0", C.c_double), ("t_ssz", C.c_double), ("atol", C.c_double), ("rtol", C.c_double),
        ("params", _dp), ("params_stride_e", C.c_int64), ("actions", _dp),
        ("A", _dp), ("A_stride_e", C.c_int64), ("D", _dp), ("D_stride_e", C.c_int64),
        ("y0", _dp), ("states", _dp), ("observations", _dp), ("y_last", _dp), ("obs_last", _dp), ("obs_noise", _dp),
        ("seed", C.c_uint64), ("episode", C.c_uint32), ("_pad0", C.c_uint32), ("t_idx", C.c_int64),
        ("env_offset", C.c_int64), ("obs_std", _dp), ("obs_std_stride_e", C.c_int64),
        ("h_state", _dp), ("n_steps", _dp), ("obs_minmax", _dp), ("nan_flag", _dp),
        ("flags", C.c_int32), ("reward_kind", C.c_int32), ("reward", _dp), ("reward_last", _dp), ("rewards_cum", _dp),
        ("action_scale", _dp),
    ]


class PackArgs(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("dim", C.c_int32), ("n_actions", C.c_int32), ("n_obs", C.c_int32),
        ("n_props", C.c_int32), ("n_sel", C.c_int32),
        ("T_step", _dp), ("actions", _dp), ("observations", _dp), ("properties", _dp), ("rewards_cum", _dp),
        ("idxs", _dp), ("packed", _dp), ("packed_stride_e", C.c_int64), ("selected", _dp),
        ("selected_stride_e", C.c_int64), ("selected_row_pitch", C.c_int64), ("dyn", _dp), ("action_scale", _dp),
    ]


class PeerPullArgs(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("n_obs", C.c_int32), ("out_slots", C.c_int32), ("blocks_per_rank", C.c_int32),
        ("flags", _dp), ("peer_ring", _dp), ("slot_stride", _dp), ("env_offset", _dp), ("env_count", _dp),
        ("obs_all", _dp), ("reward_all", _dp), ("pull_epoch", _dp), ("peer_pulled", _dp), ("finish_counter", _dp),
        ("status", _dp),
    ]


class PeerPushArgs(C.Structure):
    _fields_ = [("src", _dp), ("dst", _dp), ("bytes", C.c_int64), ("src2", _dp), ("dst2", _dp), ("bytes2", C.c_int64),
                ("flag_src", _dp), ("flag_dst", _dp),
                ("step_done", _dp), ("copy_done", _dp)]


# every symbol include/quantrl_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "qrl_abi_version": (C.c_int, []),
    "qrl_last_error": (C.c_char_p, []),
    "qrl_device_count": (C.c_int, []),
    "qrl_release": (C.c_int, []),
    "qrl_launch_count": (C.c_uint64, []),
    "qrl_host_device_pointer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "qrl_em_step": (C.c_int, [C.POINTER(EmArgs), C.c_void_p]),
    "qrl_em_step_host": (C.c_int, [C.POINTER(EmArgs), C.c_void_p, C.c_void_p, C.c_int32]),
    "qrl_ode_step": (C.c_int, [C.POINTER(OdeArgs), C.c_void_p]),
    "qrl_pack_rows": (C.c_int, [C.POINTER(PackArgs), C.c_void_p]),
    "qrl_peer_pull": (C.c_int, [C.POINTER(PeerPullArgs), C.c_void_p]),
    "qrl_peer_push": (C.c_int, [C.POINTER(PeerPushArgs), C.c_void_p, C.c_void_p]),
    "qrl_peer_event_wait": (C.c_int, [C.c_void_p]),
    "qrl_peer_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p]),
    "qrl_peer_ctx_push": (C.c_uint64, [C.c_void_p, C.POINTER(PeerPushArgs), C.c_void_p]),
    "qrl_peer_ctx_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "qrl_peer_ctx_destroy": (C.c_int, [C.c_void_p]),
    "qrl_peer_wait": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "qrl_measure_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "qrl_measure_fp64_latency": (C.c_int, [C.POINTER(C.c_double)]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises NativeError if it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("QRL_B200_LIB", LIB_PATH)   # development override (kernel variants)
        if not os.path.exists(path):
            raise NativeError(
                f"{path} is missing: build it with `python -m quantrl_b200.build` (nvcc, sm_100a). "
                "quantrl_b200 has no CPU or PyTorch fallback for its hot path.")
        handle = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if handle.qrl_abi_version() != ABI_VERSION:
            raise NativeError(f"ABI mismatch: library {handle.qrl_abi_version()} vs binding {ABI_VERSION}")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().qrl_last_error().decode(errors="replace")
        kind = "invalid argument" if rc < 0 else f"CUDA error {rc}"
        raise NativeError(f"{what} failed ({kind}): {msg}")


def ptr(t):
    """Device pointer of a torch CUDA tensor (None -> NULL).  The tensor must be contiguous float64/int32."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("expected a CUDA tensor (quantrl_b200 has no CPU path)")
    if not t.is_contiguous():
        raise NativeError("expected a contiguous tensor")
    return t.data_ptr()


def current_raw_stream() -> int:
    """``cudaStream_t`` of torch's current stream on the current device.  ``torch.cuda.current_stream()`` builds a Stream
    object and re-validates the device every call (~10 us, measured per step); the raw handle costs well under 1 us."""
    import torch
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def current_stream():
    return C.c_void_p(current_raw_stream())


def em_step(args: EmArgs, stream=None):
    check(lib().qrl_em_step(C.byref(args), stream if stream is not None else current_stream()), "qrl_em_step")


def ode_step(args: OdeArgs, stream=None):
    check(lib().qrl_ode_step(C.byref(args), stream if stream is not None else current_stream()), "qrl_ode_step")


def pack_rows(args: PackArgs, stream=None):
    check(lib().qrl_pack_rows(C.byref(args), stream if stream is not None else current_stream()), "qrl_pack_rows")


def peer_pull(args: PeerPullArgs, stream=None):
    check(lib().qrl_peer_pull(C.byref(args), stream if stream is not None else current_stream()), "qrl_peer_pull")


def peer_push(args: PeerPushArgs, main_stream, side_stream):
    check(lib().qrl_peer_push(C.byref(args), main_stream, side_stream), "qrl_peer_push")


def peer_ctx_create(side_stream):
    ctx = C.c_void_p(0)
    check(lib().qrl_peer_ctx_create(C.byref(ctx), side_stream), "qrl_peer_ctx_create")
    return ctx


def peer_ctx_push(ctx, args: PeerPushArgs, main_stream) -> int:
    seq = int(lib().qrl_peer_ctx_push(ctx, C.byref(args), main_stream))
    if seq == 0:
        check(-1, "qrl_peer_ctx_push")
    return seq


def peer_ctx_wait(ctx, seq):
    check(lib().qrl_peer_ctx_wait(ctx, seq), "qrl_peer_ctx_wait")


def peer_ctx_destroy(ctx):
    lib().qrl_peer_ctx_destroy(ctx)


def peer_event_wait(event):
    check(lib().qrl_peer_event_wait(event), "qrl_peer_event_wait")


def peer_wait(flags_ptr, world, epoch, status_ptr, stream=None):
    check(lib().qrl_peer_wait(flags_ptr, world, epoch, status_ptr, stream if stream is not None else current_stream()), "qrl_peer_wait")


def host_device_pointer(t) -> int:
    """Device-side address of a pinned host tensor (raises NativeError if it is pageable / not mapped)."""
    out = C.c_void_p(0)
    check(lib().qrl_host_device_pointer(C.c_void_p(t.data_ptr()), C.byref(out)), "qrl_host_device_pointer")
    return int(out.value)


def launch_count() -> int:
    return int(lib().qrl_launch_count())


def measure_fp64_peak(iters=4096) -> float:
    out = C.c_double(0.0)
    check(lib().qrl_measure_fp64_peak(int(iters), C.byref(out)), "qrl_measure_fp64_peak")
    return float(out.value)


def measure_fp64_latency() -> float:
    out = C.c_double(0.0)
    check(lib().qrl_measure_fp64_latency(C.byref(out)), "qrl_measure_fp64_latency")
    return float(out.value)
