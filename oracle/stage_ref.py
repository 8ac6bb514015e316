"""TEST INFRASTRUCTURE ONLY — stage the UNMODIFIED reference package into ``oracle/_ref/`` (git-ignored).

The reference (Sampreet/quantrl v0.0.8) is pure Python: there is nothing to compile, so "building the reference where it
lies" means copying its package directory, byte for byte, next to the oracle — the same role a compiled ``.so`` plays
for a C reference.  ``oracle/_ref/`` is listed in ``.gitignore`` (it never enters the history) but not in
``.gpurunignore`` (it travels to the GPU box, where /root/reference does not exist).  Run by ``__graft_entry__.build()``
whenever /root/reference is present, or by hand:

    python oracle/stage_ref.py

Consumers (checkers only — the product never imports it): ``tests/test_reference_on_b200_gpu.py`` (the reference's own
``LinearizedHOVecEnv`` on ``backend_library='b200'``) and ``bench.py --impl reference`` / ``cpu_baseline`` (the reference's
own ``LinearEnv`` on the host cores).  A manifest with the SHA-256 of every staged file is written beside the copy.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("QRL_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(verbose=True):
    src_pkg = os.path.join(SRC, "quantrl")
    if not os.path.isdir(src_pkg):
        raise FileNotFoundError(f"reference package not found at {src_pkg}")
    dst_pkg = os.path.join(DST, "quantrl")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src_pkg, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(dst_pkg):
        for f in sorted(files):
            path = os.path.join(root, f)
            manifest[os.path.relpath(path, DST)] = hashlib.sha256(open(path, "rb").read()).hexdigest()
    # byte-for-byte: every staged file equals its source
    for rel, digest in manifest.items():
        assert hashlib.sha256(open(os.path.join(SRC, rel), "rb").read()).hexdigest() == digest, rel
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} files of the unmodified reference into {DST}")
    return DST


if __name__ == "__main__":
    try:
        stage()
    except FileNotFoundError as exc:
        sys.exit(str(exc))
