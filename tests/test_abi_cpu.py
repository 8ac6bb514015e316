"""CPU: the C-ABI library builds, loads and exports every symbol include/quantrl_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(native):
    header = open(os.path.join(ROOT, "include", "quantrl_b200.h")).read()
    declared = set(re.findall(r"QRL_API\s+[\w\s\*]+?\b(qrl_\w+)\s*\(", header))
    assert declared, "no QRL_API declarations found"
    assert declared == set(native.SYMBOLS), (declared ^ set(native.SYMBOLS))
    lib = native.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.qrl_abi_version() == native.ABI_VERSION == int(re.search(r"#define QRL_ABI_VERSION (\d+)", header).group(1))


def test_struct_sizes_match_header(native):
    # natural alignment, all members 4 or 8 bytes: ctypes and nvcc must agree
    assert ctypes.sizeof(native.EmArgs) % 8 == 0 and ctypes.sizeof(native.OdeArgs) % 8 == 0
    assert native.EmArgs.t_ssz.offset == 16 and native.EmArgs.A.offset == 24
    assert native.EmArgs.seed.offset % 8 == 0 and native.EmArgs.t_idx.offset == native.EmArgs.seed.offset + 16


def test_invalid_arguments_are_reported_without_a_gpu(native):
    a = native.EmArgs()
    a.n, a.n_envs, a.K = 99, 1, 1
    with pytest.raises(native.NativeError, match="n must be in 1..16"):
        native.em_step(a, ctypes.c_void_p(0))
    assert native.lib().qrl_device_count() >= 0


def test_missing_library_fails_loudly(native, monkeypatch):
    import quantrl_b200._native as N
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", "/nonexistent/libquantrl_b200.so")
    with pytest.raises(N.NativeError, match="no CPU or PyTorch fallback"):
        N.lib()


def test_flag_and_enum_constants_match_header(native):
    """Every ``QRL_*`` integer constant of the header that the Python binding mirrors (flags, methods, systems, reward
    and RNG kinds) has the same value on both sides."""
    header = open(os.path.join(ROOT, "include", "quantrl_b200.h")).read()
    defines = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+QRL_(\w+)\s+(-?(?:0x[0-9a-fA-F]+|\d+))\b", header)}
    # enumerators written as `QRL_NAME = value`
    defines.update({m.group(1): int(m.group(2), 0) for m in re.finditer(r"\bQRL_(\w+)\s*=\s*(-?(?:0x[0-9a-fA-F]+|\d+))", header)})
    mirrored = {name: getattr(native, name) for name in dir(native)
                if name.isupper() and isinstance(getattr(native, name), int) and name in defines}
    assert len(mirrored) >= 12, sorted(mirrored)
    for name, value in mirrored.items():
        assert value == defines[name], f"{name}: binding {value} != header {defines[name]}"
    for flag in ("EM_PDL", "EM_HOST_TIDX", "EM_FORCE_GENERIC", "ODE_FORCE_THREAD", "ODE_FORCE_LANES", "ODE_LANES_SINGLE"):
        assert flag in mirrored, f"{flag} is not mirrored in the binding"


def test_integration_md_stub_matches_the_binding(native):
    """The ctypes stub INTEGRATION.md shows a reference maintainer declares the same ``qrl_em_args`` layout as the
    binding in use (same field names, types, offsets) and names the current ABI version."""
    import ctypes as C
    import os
    import re
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes as C\n.*?)\nlib = C\.CDLL", text, re.S).group(1)
    ns = {}
    exec(block, ns)   # noqa: S102  (our own documentation snippet: the struct declaration only)
    doc, ours = ns["EmArgs"], native.EmArgs
    assert C.sizeof(doc) == C.sizeof(ours)
    assert [(n, getattr(doc, n).offset, getattr(doc, n).size) for n, _ in doc._fields_] == \
           [(n, getattr(ours, n).offset, getattr(ours, n).size) for n, _ in ours._fields_]
    assert f"qrl_abi_version() == {native.ABI_VERSION}" in text
