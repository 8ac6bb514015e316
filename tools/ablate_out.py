"""Output-stage ablation of the tiled EM kernel (development): variant libraries built with -DQRL_DBG_*."""
import sys, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
def run(lib, flags, traj=1):
    e = dict(os.environ)
    if lib: e["QRL_B200_LIB"] = os.path.join(HERE, "variants", f"lib_{lib}.so")
    code = (f"import sys; sys.path.insert(0, {HERE!r}); from quick_em_bench import bench; "
            f"print(round(bench(4096, 4, 100, True, bool({traj}), reps=30, flags={flags})['ms']*1e3, 1))")
    out = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    return out.stdout.strip().splitlines()[-1] if out.returncode == 0 else "ERR " + out.stderr[-200:]
for lib in ("", "NO_STG", "NO_MINMAX", "NO_REWARD"):
    print(f"{lib or 'default':10s} LSU: full {run(lib, 32)}  output-only {run(lib, 32 + 6)} | TMA: full {run(lib, 0)} output-only {run(lib, 6)} | traj off: full {run(lib, 32, 0)} output-only {run(lib, 38, 0)}", flush=True)
