"""A/B timing of tiled-kernel flag variants at config 3 (pure device time); pack: 0 none, 2 cum column, 1 all columns."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
flag_sets = [int(x) for x in sys.argv[1:]] or [0, 32]
for pack in (0, 2, 1):
    for flags in flag_sets:
        r = [bench(4096, 4, 100, True, True, reps=40, flags=flags, pack=pack)["ms"] * 1e3 for _ in range(3)]
        print(f"pack={pack} flags={flags}: {min(r):.1f} us", flush=True)
