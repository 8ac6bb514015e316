"""Stage ablation and tuning-knob sweep of the tiled EM kernel at config 3 (pure device time, CUDA-graph replay)."""
import sys, os, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

def one(flags, pack, E=4096, n=4):
    from quick_em_bench import bench
    r = bench(E, n, 100, True, True, reps=30, flags=flags, pack=pack)
    return r["ms"] * 1e3

if len(sys.argv) > 1 and sys.argv[1] == "child":
    flags, pack = int(sys.argv[2]), int(sys.argv[3])
    E = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    print(f"{one(flags, pack, E, n):.1f}")
    sys.exit(0)

def run(env, flags, pack, E=4096, n=4):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, __file__, "child", str(flags), str(pack), str(E), str(n)], env=e, capture_output=True, text=True)
    return out.stdout.strip().splitlines()[-1] if out.returncode == 0 else "ERR " + out.stderr[-300:]

rows = ((0, "full"), (2, "no produce"), (4, "no recurrence"), (8, "no output"), (12, "produce only"), (10, "rec only"),
        (6, "output only"), (14, "nothing"), (1, "generic kernel"))
for env in ({}, {"QRL_EM_TILE_ILP": "1"}, {"QRL_EM_TILE_C": "16"}, {"QRL_EM_TILE_C": "24"}, {"QRL_EM_TILE_ILP": "1", "QRL_EM_TILE_C": "20"}):
    print(env, "| pack=0:", run(env, 0, 0), "us | item output:", run(env, 32, 0), "us | pack=cum column:", run(env, 0, 2), "us | pack=all 7:", run(env, 0, 1), "us", flush=True)
print("ablation (default knobs, no pack):", " | ".join(f"{name}: {run({}, f, 0)}" for f, name in rows), flush=True)
for E, n in ((1024, 4), (8192, 4), (4096, 8), (2048, 2), (16384, 4)):
    print(f"E={E} n={n}: tiled {run({}, 16, 0, E, n)} us, generic {run({}, 1, 0, E, n)} us", flush=True)
