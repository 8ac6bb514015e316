import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_em_bench import bench
sel = sys.argv[1:] or ["all"]
rows = ((0, "full"), (2, "no produce"), (4, "no recurrence"), (8, "no output"), (6, "no produce, no recurrence"),
        (12, "no rec, no out (produce only)"), (10, "no prod, no out (rec only)"), (14, "nothing"), (1, "generic kernel"))
if sel != ["all"]:
    rows = [r for r in rows if str(r[0]) in sel]
out = []
for flags, name in rows:
    r = bench(4096, 4, 100, True, True, reps=30, flags=flags)
    out.append(f"{name}: {r['ms']*1e3:.1f} us")
print(os.environ.get("QRL_B200_LIB", "default")[-16:], " | ".join(out))
