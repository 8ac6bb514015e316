"""Knob sweep of the third-generation tiled EM kernel (pure device time, CUDA-graph replay): chunk length C, ring depth D,
threads per CTA, against the second-generation kernel (QRL_EM_NO_T3=1) and the generic kernel, at the config 3 / 4 / 5
shard sizes.    python tools/t3_sweep.py [quick]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def child():
    sys.path.insert(0, HERE)
    from quick_em_bench import bench
    E, n, flags, pack = (int(v) for v in sys.argv[2:6])
    r = bench(E, n, 100, True, True, reps=30, flags=flags, pack=pack)
    print(f"{r['ms'] * 1e3:.1f}")


def run(env, E, n, flags=16, pack=2):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, __file__, "child", str(E), str(n), str(flags), str(pack)], env=e,
                         capture_output=True, text=True)
    return out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else "ERR " + out.stderr[-200:]


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = [(4096, 4), (8192, 8), (8192, 4)] if quick else [(4096, 4), (1024, 4), (8192, 4), (8192, 8), (16384, 8), (32768, 8), (65536, 8), (65536, 4)]
    for E, n in cases:
        print(f"E={E} n={n}: gen3 {run({}, E, n)} us | gen2 {run({'QRL_EM_NO_T3': '1'}, E, n)} us | generic {run({}, E, n, flags=1)} us", flush=True)
    if quick or not (len(sys.argv) > 1 and sys.argv[1] == "knobs"):
        sys.exit(0)
    for E, n in [(4096, 4), (8192, 8)]:
        for th in (512, 640, 768):
            row = []
            for D in (3, 4, 6):
                for C in (6, 12, 18, 24, 36):
                    row.append(f"D{D}C{C}:{run({'QRL_EM_T3_THREADS': str(th), 'QRL_EM_T3_D': str(D), 'QRL_EM_T3_C': str(C)}, E, n)}")
            print(f"E={E} n={n} threads={th}: " + " ".join(row), flush=True)
