"""Turn the scratch ncu outputs of one round-profile run (gpurun_out/) into the tracked summaries under profiles/.

usage: python tools/summarize_profiles.py <tag> [kernel-regex]     e.g. r02k em_tile3
       (reads gpurun_out/launches_<tag>.csv, prof_tile_<tag>.ncu-rep and, if present, prof_ring_<tag>.ncu-rep)
"""
import csv, io, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
kregex = sys.argv[2] if len(sys.argv) > 2 else "em_tile_kernel"
short = tag.replace("r01", "r1").replace("r02", "r2")
out = os.path.join(ROOT, "profiles")

# ---- launch list -----------------------------------------------------------------------------------------------
rows = [r for r in csv.reader(l for l in open(os.path.join(ROOT, "gpurun_out", f"launches_{short}.csv")) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    agg[r[ik]][0] += 1
    agg[r[ik]][1] += float(r[iv].replace(",", "")) / 1e3
step_total = sum(v[1] for k, v in agg.items() if "dfma_" not in k)
with open(os.path.join(out, f"{tag}_launches_bench_summary.txt"), "w") as f:
    f.write(f"# {tag} launch list of `python bench.py --steps 40 --warmup 5 --no-cpu-baseline` "
            "(ncu --metrics gpu__time_duration.sum --clock-control none -c 600)\n"
            "# shares exclude the FP64 peak micro-benchmark (dfma_*), which runs once after the timed region\n\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = "   n/a" if "dfma_" in k else f"{100 * us / step_total:5.1f}%"
        f.write(f"{n:5d} launches {us:10.1f} us total {us / n:9.2f} us avg {share}  {k[:110]}\n")
subprocess.run(["cp", os.path.join(ROOT, "gpurun_out", f"launches_{short}.csv"), os.path.join(out, f"{tag}_launches_bench.csv")])

# ---- full capture of the dominant kernel ---------------------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", f"prof_tile_{short}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_write.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
lines = [f"## {tag} — `{data[0][hdr.index('Kernel Name')]}` inside `bench.py --steps 40 --warmup 5` (config 3, reward-column cache fused)\n",
         f"command: `ncu --set full --clock-control none --import-source on -k regex:{kregex} -s 60 -c 2 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --only-headline`\n",
         "| metric | unit | value(s) per captured launch |", "|---|---|---|"]
vals = {}
for m in want:
    if m in hdr:
        i = hdr.index(m)
        vals[m] = [d[i] for d in data]
        lines.append(f"| {m} | {units[i]} | {', '.join(vals[m])} |")
def tobytes(m):
    i = hdr.index(m)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return [float(d[i].replace(",", "")) * scale for d in data]
rd, wr = tobytes("dram__bytes_read.sum"), tobytes("dram__bytes_write.sum")
traffic = sum(rd) / len(rd) + sum(wr) / len(wr)
lines.append(f"\nDRAM traffic per launch = dram__bytes_read.sum + dram__bytes_write.sum = **{traffic / 1e6:.2f} MB** against 32.8 MB "
             "algorithmic (80 B x 4096 envs x 100 sub-steps): the 33 MB of trajectory rows are rewritten in place every action step "
             "and stay resident in the 126 MB L2; only the episode-cache column streams out.\n")
tj = {"kernel": data[0][hdr.index("Kernel Name")], "capture": f"profiles/{tag}_ncu_summary.md",
      "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
      "traffic_bytes_per_launch": traffic, "workload": "bench.py config 3, E=4096 n=4 K=100"}
# the same launch with its trajectory blocks rotating over 8 sets (> L2): tools/prof_ring.py
ring_rep = os.path.join(ROOT, "gpurun_out", f"prof_ring_{short}.ncu-rep")
if os.path.exists(ring_rep):
    raw2 = subprocess.run(["ncu", "-i", ring_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r2 = list(csv.reader(io.StringIO(raw2)))
    h2, u2, d2 = r2[0], r2[1], r2[2:]
    def tob2(m):
        i = h2.index(m)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u2[i]]
        return [float(d[i].replace(",", "")) * scale for d in d2]
    rd2, wr2 = tob2("dram__bytes_read.sum"), tob2("dram__bytes_write.sum")
    t2 = sum(rd2) / len(rd2) + sum(wr2) / len(wr2)
    dur = [d[h2.index("gpu__time_duration.sum")] for d in d2]
    lines.append(f"### ring of 8 sets of trajectory blocks (> L2): `ncu --set full -k regex:{kregex} -s 20 -c 2 python tools/prof_ring.py`\n")
    lines.append(f"gpu__time_duration.sum ({u2[h2.index('gpu__time_duration.sum')]}): {', '.join(dur)}; DRAM read {sum(rd2) / len(rd2) / 1e6:.2f} MB + "
                 f"write {sum(wr2) / len(wr2) / 1e6:.2f} MB = **{t2 / 1e6:.2f} MB per launch** (every launch writes lines that are not L2 "
                 "resident and evicts dirty ones): the algorithmic 32.8 MB reach HBM, the kernel time does not change — it is issue bound.\n")
    tj["traffic_bytes_per_launch_ring"] = t2
open(os.path.join(out, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump(tj, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print("\n".join(lines))
print(open(os.path.join(out, f"{tag}_launches_bench_summary.txt")).read())
