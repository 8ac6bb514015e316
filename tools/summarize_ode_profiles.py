"""Summaries of the deterministic-path ncu captures (gpurun_out/prof_<name>_<tag>.ncu-rep) -> profiles/<tag>_ode_ncu_summary.md

usage: python tools/summarize_ode_profiles.py <tag> <name:description> ...
"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
short = tag.replace("r01", "r1")
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
lines = [f"# {tag} — ncu `--set full --clock-control none` captures of the deterministic path (`qrl_ode_step`), one launch each, "
         "driver `tools/prof_ode.py` (optomech, K = 100 grid rows per launch, trajectory on: States + Observations rows)\n"]
for spec in sys.argv[2:]:
    name, desc = spec.split(":", 1)
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{name}_{short}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines += [f"\n## {desc}\n", f"kernel: `{data[0][hdr.index('Kernel Name')]}`\n", "| metric | unit | value |", "|---|---|---|"]
    for m in want:
        if m in hdr:
            i = hdr.index(m)
            lines.append(f"| {m} | {units[i]} | {', '.join(d[i] for d in data)} |")
open(os.path.join(ROOT, "profiles", f"{tag}_ode_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
