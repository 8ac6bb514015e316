"""``B200Backend`` — the fourth ``quantrl.backends`` backend (CUDA tensors on one B200, FP64 by default).

Mirrors the operator surface of the reference's ``BaseBackend`` (quantrl/backends/base.py:16-1033: 17 abstract
methods :122-552, the concrete NumPy-like ops :609-1033, the eight ``jit_*`` aliases :577-607) with the same method
names, keyword names and return conventions, so the reference's env/solver code can run on it unchanged when it is
registered as ``BACKEND_INSTANCES['b200']`` (quantrl/backends/context_manager.py:14,38).

Differences from the reference's ``TorchBackend`` (quantrl/backends/torch.py), all of them fixes of SURVEY.md F6:
the device is always a real ``torch.device('cuda:<current>')`` and is passed explicitly (the process-wide default
device is not touched), generators are seeded from an integer drawn from the SeedSequence child, and
``convert_to_numpy`` always goes through ``.detach().cpu()``.

These ops are plumbing (allocation, reshapes, host<->device, user hooks written in torch).  The per-step arithmetic
of the hot path never runs through them: it is launched through the C ABI (``quantrl_b200._native``).
"""
import numpy as np
import torch


class B200Backend:
    """Backend interfacing CUDA tensors; method-for-method compatible with ``quantrl.backends.base.BaseBackend``."""

    def __init__(self, precision: str = "double", device=None):
        assert precision in ["single", "double"], "parameter ``precision`` can be either ``'single'`` or ``'double'``."
        if not torch.cuda.is_available():
            raise RuntimeError("quantrl_b200.B200Backend needs a CUDA device: there is no CPU fallback")
        self.name = "b200"
        self.library = torch
        self.tensor_type = torch.Tensor
        self.precision = precision
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtypes = {
            "typed": {
                "single": {"integer": torch.int32, "real": torch.float32, "complex": torch.complex64},
                "double": {"integer": torch.int64, "real": torch.float64, "complex": torch.complex128},
            },
            "numpy": {
                "single": {"integer": np.int32, "real": np.float32, "complex": np.complex64},
                "double": {"integer": np.int64, "real": np.float64, "complex": np.complex128},
            },
        }
        self.seed_sequence = None

    # ---- helpers (quantrl/backends/base.py:74-120, 554-575) ------------------------------------------------
    def dtype_from_str(self, dtype: str = None, numpy: bool = False):
        if dtype is None:
            return None
        return self.dtypes["numpy" if numpy else "typed"][self.precision][dtype]

    def is_typed(self, tensor, dtype: str = None) -> bool:
        if isinstance(tensor, torch.Tensor) and tensor.device == self.device:
            return dtype is None or tensor.dtype == self.dtype_from_str(dtype)
        return False

    def get_seedsequence(self, seed: int = None) -> np.random.SeedSequence:
        if seed is None:
            entropy = np.random.randint(1234567890)
        else:
            entropy = np.random.default_rng(seed).integers(0, 1234567890, (1,))[0]
        return np.random.SeedSequence(entropy)

    # ---- the 17 abstract methods ---------------------------------------------------------------------------
    def convert_to_typed(self, tensor, dtype: str = None) -> torch.Tensor:
        if self.is_typed(tensor=tensor, dtype=dtype):
            return tensor
        if isinstance(tensor, torch.Tensor):
            return tensor.to(device=self.device, dtype=self.dtype_from_str(dtype))
        if isinstance(tensor, (list, tuple)) and len(tensor) > 0 and isinstance(tensor[0], torch.Tensor):
            return torch.stack([t.to(self.device) for t in tensor]).to(dtype=self.dtype_from_str(dtype))
        return torch.as_tensor(np.asarray(tensor), dtype=self.dtype_from_str(dtype), device=self.device)

    @staticmethod
    def _host(t):
        # torch.conj() / negation views are lazy bits on the tensor; NumPy needs them materialised
        return t.detach().resolve_conj().resolve_neg().cpu().numpy()

    def convert_to_numpy(self, tensor, dtype: str = None) -> np.ndarray:
        if isinstance(tensor, (list, tuple)):
            tensor = [self._host(t) if isinstance(t, torch.Tensor) else np.asarray(t) for t in tensor]
        elif isinstance(tensor, torch.Tensor):
            tensor = self._host(tensor)
        return np.asarray(tensor, dtype=self.dtype_from_str(dtype, numpy=True))

    def generator(self, seed: int = None) -> torch.Generator:
        # one SeedSequence per backend instance; every call spawns the next child (backends/base.py:101-120,
        # backends/numpy.py:54-59)
        if self.seed_sequence is None:
            self.seed_sequence = self.get_seedsequence(seed)
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(self.seed_sequence.spawn(1)[0].generate_state(1, dtype=np.uint64)[0] >> np.uint64(1)))
        return gen

    def integers(self, generator, shape: tuple, low: int = 0, high: int = 1000, dtype: str = None):
        return torch.randint(low, high, tuple(shape), generator=generator, dtype=self.dtype_from_str(dtype),
                             device=self.device)

    def normal(self, generator, shape: tuple, mean: float = 0.0, std: float = 1.0, dtype: str = None):
        return self.empty(shape=shape, dtype=dtype).normal_(mean, std, generator=generator)

    def uniform(self, generator, shape: tuple, low: float = 0.0, high: float = 1.0, dtype: str = None):
        return low + (high - low) * torch.rand(tuple(shape), generator=generator, dtype=self.dtype_from_str(dtype),
                                               device=self.device)

    def transpose(self, tensor, axis_0: int = None, axis_1: int = None):
        if axis_0 is None or axis_1 is None:
            return self.convert_to_typed(tensor=tensor).T
        return torch.transpose(tensor, dim0=axis_0, dim1=axis_1)

    def repeat(self, tensor, repeats, axis):
        return torch.repeat_interleave(tensor, repeats=repeats, dim=axis)

    def add(self, tensor_0, tensor_1, out):
        return torch.add(tensor_0, tensor_1, out=out)

    def matmul(self, tensor_0, tensor_1, out):
        return torch.matmul(tensor_0, tensor_1, out=out)

    def dot(self, tensor_0, tensor_1, out):
        return torch.dot(tensor_0, tensor_1, out=out)

    def norm(self, tensor, axis):
        return torch.norm(tensor, dim=axis)

    def concatenate(self, tensors: tuple, axis, out):
        return torch.concatenate(tensors, dim=axis, out=out)

    def stack(self, tensors: tuple, axis, out):
        return torch.stack(tensors, dim=axis, out=out)

    def update(self, tensor, indices, values):
        tensor[indices] = values
        return tensor

    def if_else(self, condition, func_true, func_false, args):
        if condition:
            return func_true(args)
        return func_false(args)

    def iterate_i(self, func, iterations_i: int, Y, args: tuple = None):
        for i in range(iterations_i):
            Y = func(i, Y, args)
        return Y

    # ---- jit aliases (backends/base.py:577-607) ----------------------------------------------------------------
    def jit_transpose(self, tensor, axis_0, axis_1):
        return self.transpose(tensor, axis_0, axis_1)

    def jit_repeat(self, tensor, repeats, axis):
        return self.repeat(tensor, repeats, axis)

    def jit_add(self, tensor_0, tensor_1, out):
        return self.add(tensor_0, tensor_1, out)

    def jit_matmul(self, tensor_0, tensor_1, out):
        return self.matmul(tensor_0, tensor_1, out)

    def jit_dot(self, tensor_0, tensor_1, out):
        return self.dot(tensor_0, tensor_1, out)

    def jit_concatenate(self, tensors, axis, out):
        return self.concatenate(tensors, axis, out)

    def jit_stack(self, tensors, axis, out):
        return self.stack(tensors, axis, out)

    def jit_update(self, tensor, indices, values):
        return self.update(tensor, indices, values)

    # ---- concrete NumPy-like ops (backends/base.py:609-1033) -----------------------------------------------
    def empty(self, shape: tuple, dtype: str = None):
        return torch.empty(tuple(shape), dtype=self.dtype_from_str(dtype), device=self.device)

    def zeros(self, shape: tuple, dtype: str = None):
        return torch.zeros(tuple(shape), dtype=self.dtype_from_str(dtype), device=self.device)

    def ones(self, shape: tuple, dtype: str = None):
        return torch.ones(tuple(shape), dtype=self.dtype_from_str(dtype), device=self.device)

    def eye(self, N: int, M: int = None, dtype: str = None):
        return torch.eye(N, M if M is not None else N, dtype=self.dtype_from_str(dtype), device=self.device)

    def diag(self, tensor, dtype: str = None):
        return torch.diag(self.convert_to_typed(tensor=tensor, dtype=dtype))

    def arange(self, start: float, stop: float, ssz: int, dtype: str = None):
        return torch.arange(start, stop, ssz, dtype=self.dtype_from_str(dtype), device=self.device)

    def linspace(self, start: float, stop: float, dim: int, dtype: str = None):
        return torch.linspace(start, stop, dim, dtype=self.dtype_from_str(dtype), device=self.device)

    def shape(self, tensor) -> tuple:
        return tuple(tensor.shape)

    def reshape(self, tensor, shape: tuple):
        return torch.reshape(tensor, tuple(shape))

    def flatten(self, tensor):
        return torch.flatten(tensor)

    def real(self, tensor):
        return torch.real(tensor)

    def imag(self, tensor):
        return torch.imag(tensor)

    def sqrt(self, tensor):
        return torch.sqrt(tensor)

    def sum(self, tensor, axis=None):
        return torch.sum(tensor) if axis is None else torch.sum(tensor, axis)

    def cumsum(self, tensor, axis=None):
        return torch.cumsum(tensor.flatten() if axis is None else tensor, 0 if axis is None else axis)

    def conj(self, tensor):
        return torch.conj(tensor)

    def min(self, tensor):
        return torch.min(tensor)

    def max(self, tensor):
        return torch.max(tensor)

    def argmax(self, tensor):
        return torch.argmax(tensor)


_INSTANCES = {}


def get_backend_instance(library: str = "b200", precision: str = "double", device: str = "gpu"):
    """Same contract as quantrl/backends/context_manager.py:16-53: one cached instance per library name."""
    assert "b200" in library.lower(), "quantrl_b200 only provides the ``'b200'`` backend"
    if "b200" not in _INSTANCES:
        _INSTANCES["b200"] = B200Backend(precision=precision)
    return _INSTANCES["b200"]
