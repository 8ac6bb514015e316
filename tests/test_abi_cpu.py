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


_P = 0x1000   # a non-NULL "device pointer": every case below is rejected before anything is dereferenced or launched


def _em(**kw):
    import quantrl_b200._native as N
    a = N.EmArgs()
    a.n, a.n_envs, a.K, a.rng_mode, a.t_ssz = 4, 8, 10, N.RNG_PHILOX, 0.01
    a.x0, a.A, a.nz = _P, _P, _P
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _ode(**kw):
    import quantrl_b200._native as N
    a = N.OdeArgs()
    a.system, a.num_modes, a.num_quads, a.n_envs, a.K, a.method, a.substeps = N.SYS_OPTOMECH, 2, 4, 8, 10, N.ODE_RK4, 1
    a.t_ssz, a.atol, a.rtol, a.y0, a.params, a.actions, a.n_actions = 0.01, 1e-9, 1e-6, _P, _P, _P, 1
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def _pack(**kw):
    import quantrl_b200._native as N
    a = N.PackArgs()
    a.n_envs, a.dim, a.n_actions, a.n_obs, a.n_props, a.n_sel = 8, 11, 1, 4, 0, 1
    a.T_step, a.actions, a.observations, a.rewards_cum, a.idxs = _P, _P, _P, _P, _P
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.mark.parametrize("entry,build,kw,msg", [
    ("em_step", _em, dict(n=0), "n must be in 1..16"),
    ("em_step", _em, dict(n=17), "n must be in 1..16"),
    ("em_step", _em, dict(K=-1), "non-negative"),
    ("em_step", _em, dict(n_envs=-3), "non-negative"),
    ("em_step", _em, dict(rng_mode=7), "bad rng_mode"),
    ("em_step", _em, dict(x0=None), "x0 is NULL"),
    ("em_step", _em, dict(A=None), "A / nz is NULL"),
    ("em_step", _em, dict(rng_mode=0), "fed mode needs W"),
    ("em_step", _em, dict(reward_kind=1), "reward_w is NULL"),
    ("em_step", _em, dict(t_idx=-1), "negative t_idx"),
    ("em_step", _em, dict(env_offset=-1), "negative t_idx / env_offset"),
    ("em_step", _em, dict(A_act=_P), "qrl_em_step"),                       # in-kernel drift functor without actions
    ("em_step", _em, dict(dyn_advance=10), "qrl_em_step"),                 # last-block bookkeeping without its counter
    ("em_step", _em, dict(minmax_out=_P, finish_counter=_P), "minmax_out needs obs_minmax"),
    ("em_step", _em, dict(peer_flags=_P), "qrl_em_step"),                  # fused barrier without epoch / counter
    ("em_step", _em, dict(packed=_P), "packed rows need a fused reward functor"),
    ("em_step", _em, dict(packed=_P, reward_kind=1, reward_w=_P, rewards_cum=_P), "packed rows need T_step"),
    ("em_step", _em, dict(packed=_P, reward_kind=1, reward_w=_P, rewards_cum=_P, T_step=_P, sel_idxs=_P, n_sel=3,
                          packed_row_pitch=2), "packed_row_pitch < n_sel"),
    ("em_step", _em, dict(packed=_P, reward_kind=1, reward_w=_P, rewards_cum=_P, T_step=_P, sel_idxs=_P, n_sel=3,
                          packed_stride_e=5), "packed_stride_e too small"),
    ("ode_step", _ode, dict(K=-1), "non-negative"),
    ("ode_step", _ode, dict(method=99), "unknown method"),
    ("ode_step", _ode, dict(substeps=0), "substeps must be >= 1"),
    ("ode_step", _ode, dict(method=1, atol=0.0), "atol and rtol must be positive"),
    ("ode_step", _ode, dict(t_ssz=0.0), "t_ssz must be positive"),
    ("ode_step", _ode, dict(y0=None), "y0 is NULL"),
    ("ode_step", _ode, dict(t_idx=-5), "negative t_idx"),
    ("ode_step", _ode, dict(num_quads=6), "optomech needs num_modes=2, num_quads=4"),
    ("ode_step", _ode, dict(params=None), "optomech needs params"),
    ("ode_step", _ode, dict(system=2, num_modes=3, num_quads=4), "kerr_chain needs num_quads = 2 \\* num_modes"),
    ("ode_step", _ode, dict(system=3, num_modes=2), "const_AD has no modes"),
    ("ode_step", _ode, dict(system=3, num_modes=0, num_quads=4), "const_AD needs A and D"),
    ("pack_rows", _pack, dict(n_obs=0), "qrl_pack_rows"),
    ("pack_rows", _pack, dict(observations=None), "NULL input"),
    ("pack_rows", _pack, dict(actions=None), "actions is NULL"),
    ("pack_rows", _pack, dict(n_props=2), "properties is NULL"),
    ("pack_rows", _pack, dict(selected=_P, idxs=None), "idxs is NULL"),
    ("pack_rows", _pack, dict(selected=_P, n_sel=3, selected_row_pitch=2), "selected_row_pitch < n_sel"),
    ("pack_rows", _pack, dict(packed=_P, packed_stride_e=3), "qrl_pack_rows"),
])
def test_every_argument_check_rejects_before_any_launch(native, entry, build, kw, msg):
    """The C ABI validates its argument block on the host (return code < 0 + ``qrl_last_error``) before touching the
    device: each ``QRL_REQUIRE`` of qrl_em_step / qrl_ode_step / qrl_pack_rows, exercised without a GPU."""
    n0 = native.launch_count()
    with pytest.raises(native.NativeError, match=msg) as e:
        getattr(native, entry)(build(**kw), ctypes.c_void_p(0))
    assert "invalid argument" in str(e.value) and native.launch_count() == n0


def test_null_argument_block_and_micro_benchmarks_are_rejected(native):
    lib = native.lib()
    for fn in (lib.qrl_em_step, lib.qrl_ode_step, lib.qrl_pack_rows):
        assert fn(None, None) < 0 and b"args is NULL" in lib.qrl_last_error()
    assert lib.qrl_measure_fp64_peak(0, None) < 0 and lib.qrl_measure_fp64_latency(None) < 0
    assert lib.qrl_host_device_pointer(None, None) < 0


def test_product_package_never_touches_the_oracle():
    """``oracle/`` is test infrastructure: no module of ``quantrl_b200`` (nor the example scripts) imports it, puts it on
    ``sys.path`` or reads the reference tree; ``bench.py`` does so only inside its CPU-baseline / reference-arm leg."""
    import ast
    banned = {"em_oracle", "lyap_oracle", "measure_oracle", "ref_shim", "make_golden", "em_port", "systems"}
    pkg = os.path.join(ROOT, "quantrl_b200")
    files = [os.path.join(d, f) for d, _, fs in os.walk(pkg) for f in fs if f.endswith(".py")]
    files += [os.path.join(ROOT, "examples", f) for f in os.listdir(os.path.join(ROOT, "examples")) if f.endswith(".py")]
    assert len(files) >= 12
    for path in files:
        src = open(path).read()
        tree = ast.parse(src)
        for node in ast.walk(tree):
            if isinstance(node, ast.Import):
                names = [a.name.split(".")[0] for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.level == 0:
                names = [(node.module or "").split(".")[0]]
            else:
                continue
            assert not banned.intersection(names) and "oracle" not in names, (path, names)
        code_only = "\n".join(ln.split("#")[0] for ln in src.splitlines())
        assert "/root/reference" not in code_only.replace('"""', ""), path
        assert 'os.path.join(ROOT, "oracle")' not in src and "'oracle'" not in src, path
    # bench.py: the oracle import sits inside the function that serves cpu_baseline / --impl reference
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in ast.dump(n) for n in top)


def test_header_is_plain_c_and_a_c_program_links_against_the_library(native, tmp_path):
    """The boundary is a C ABI: ``include/quantrl_b200.h`` compiles as strict C99 and as C++ (``-Wall -Wextra -Wpedantic
    -Werror``), and a plain C program links against ``libquantrl_b200.so``, reads the ABI version and gets an argument
    error back through the return code + ``qrl_last_error`` — no Python, no torch types in the way."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from quantrl_b200.build import LIB_PATH
    src = tmp_path / "abi_client.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "quantrl_b200.h"\n'
        "int main(void) {\n"
        "    qrl_em_args a;\n    memset(&a, 0, sizeof a);\n    a.n = 99; a.n_envs = 1; a.K = 1;\n"
        "    if (qrl_abi_version() != QRL_ABI_VERSION) return 2;\n"
        "    if (qrl_em_step(&a, NULL) >= 0) return 3;\n"
        "    if (qrl_ode_step(NULL, NULL) >= 0 || qrl_pack_rows(NULL, NULL) >= 0) return 4;\n"
        '    printf("%d %s\\n", qrl_abi_version(), qrl_last_error());\n    return 0;\n}\n')
    inc = os.path.join(ROOT, "include")
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        subprocess.run([cc, std, "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-I", inc, "-fsyntax-only"]
                       + (["-x", "c++"] if cc == "g++" else []) + [str(src)], check=True)
    exe = tmp_path / "abi_client"
    libdir = os.path.dirname(LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lquantrl_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, (p.returncode, p.stderr)
    assert p.stdout.split()[0] == str(native.ABI_VERSION) and "args is NULL" in p.stdout


def test_every_struct_field_offset_matches_the_header(native, tmp_path):
    """``offsetof`` of every field of qrl_em_args / qrl_ode_args / qrl_pack_args as a C compiler lays the header out ==
    the offsets of the ctypes mirror (same field names on both sides), plus the struct sizes."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"qrl_em_args": native.EmArgs, "qrl_ode_args": native.OdeArgs, "qrl_pack_args": native.PackArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "quantrl_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'    printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'    printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["    return 0;", "}"]
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "offsets"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = {}
    for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines():
        cname, fname, off = ln.split()
        got[(cname, fname)] = int(off)
    for cname, cls in structs.items():
        assert got[(cname, "sizeof")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
