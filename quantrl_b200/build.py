"""Build the C-ABI shared library ``quantrl_b200/lib/libquantrl_b200.so`` with nvcc for sm_100a (in-tree).

``python -m quantrl_b200.build`` or ``quantrl_b200.build.build_extension()``.  nvcc cross-compiles without a GPU.
Every ``csrc/*.cu`` is compiled to an object in parallel and linked into one shared library.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libquantrl_b200.so")
HEADER = os.path.join(PKG, "..", "include", "quantrl_b200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-DQRL_BUILD"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + [HEADER]:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if path.endswith((".cuh", ".h")):
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:12]


def _compile_one(job):
    src, obj, verbose = job
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build_extension(force=False, verbose=False):
    """Compile every CUDA source of the package (only the stale ones unless ``force``) and link one shared library."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    digest = _headers_digest()
    jobs, objs = [], []
    for f in sources():
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ_DIR, f"{f[:-3]}.{digest}.o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            jobs.append((src, obj, verbose))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            for src, rc, out in pool.map(_compile_one, jobs):
                if verbose or rc != 0:
                    sys.stderr.write(out)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    if jobs or not os.path.exists(LIB_PATH):
        res = subprocess.run([_nvcc(), "--shared", "-o", LIB_PATH] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("linking libquantrl_b200.so failed")
        for stale in os.listdir(OBJ_DIR):  # objects of older header digests
            if os.path.join(OBJ_DIR, stale) not in objs:
                os.remove(os.path.join(OBJ_DIR, stale))
    return LIB_PATH


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv))
