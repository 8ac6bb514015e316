import sys, os
sys.path.insert(0, "tools")
from quick_em_bench import bench
for pack in (0, 2):
    for flags in (0, 32):
        r = [bench(4096, 4, 100, True, True, reps=40, flags=flags, pack=pack)["ms"] * 1e3 for _ in range(3)]
        print(f"pack={pack} flags={flags}: {min(r):.1f} us", flush=True)
