"""Per-source-line summary of an ``ncu --set full --import-source on`` capture (compile with -lineinfo).

    ncu -i capture.ncu-rep --page source --csv --print-source sass,cuda > page.csv
    python tools/summarize_source_page.py page.csv [--top 25] [--kernel-regex em_tile_kernel] > profiles/<name>.md

Aggregates the SASS rows of the interleaved source page by (file, line): warp instructions executed, stall samples and
the dominant stall reasons, then prints per-file totals, the hottest lines and the instruction mix by pipe class.
Runs anywhere ncu's report reader runs (no GPU needed)."""
import argparse
import collections
import csv
import re
import sys

PIPE = [("fp64", r"^(DFMA|DMUL|DADD|DSETP|DMNMX|F2F\.F64|I2F\.F64|F2I\.\w*F64|MUFU\.(RCP64H|RSQ64H))"),
        ("imad.wide", r"^IMAD\.WIDE"), ("imad/int-mul", r"^(IMAD|IMUL)"),
        ("alu (lop/shf/iadd/sel/mov...)", r"^(LOP3|SHF|IADD|IADD3|LEA|SEL|MOV|ISETP|PRMT|FSEL|IABS|IMNMX|VIADD|PLOP3|P2R|R2P|CS2R|S2R|FSETP|FMNMX)"),
        ("shared ld/st", r"^(LDS|STS|LDSM)"), ("global ld/st/atom", r"^(LDG|STG|ATOMG|RED|LD\.|ST\.|ATOM|LDC|LDCU|UBLKCP|UTMA)"),
        ("shuffle/vote", r"^(SHFL|VOTE|MATCH|REDUX)"), ("barrier/sync", r"^(BAR|WARPSYNC|BSYNC|BSSY|MEMBAR|FENCE|NANOSLEEP|ERRBAR|DEPBAR|ACQBULK|SYNCS)"),
        ("branch", r"^(BRA|BRX|EXIT|RET|CALL|JMP|BREAK)"), ("uniform datapath", r"^(U[A-Z0-9]+|R2UR|S2UR)")]


def pipe_of(sass):
    op = sass.strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", op)
    for name, rx in PIPE:
        if re.match(rx, op):
            return name
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--kernel-regex", default=".")
    ap.add_argument("--regions", default="", help="roll-up of source lines: 'file.cu:lo-hi=name,file.cuh:lo-hi=name,...'")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv, newline="")))
    # The interleaved page lists a SASS instruction under EVERY level of its inline stack (the callee's line in the
    # header and the call site in the kernel file) and repeats everything per captured launch.  Keep launch 0 and give
    # each address to its innermost location: a file outside the repo, else a header (*.cuh), else the kernel file.
    def depth(path):
        return 2 if not path.startswith("/root/repo") else (1 if path.endswith((".cuh", ".h", ".hpp")) else 0)

    addr = {}
    src_of = {}
    file_path = func = header = cur = None
    use = False
    kernel_name = None
    seen_files, launch = set(), 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            file_path = r[1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            use = re.search(a.kernel_regex, func) is not None
            if use:
                kernel_name = kernel_name or func
                if (func, file_path) in seen_files:
                    launch += 1
                    seen_files = set()
                seen_files.add((func, file_path))
            continue
        if r[0] == "Line No":
            header = r
            i_inst, i_samp = header.index("Instructions Executed"), header.index("# Samples")
            stall_cols = [(i, h) for i, h in enumerate(header) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if header is None or not use or launch > 0:
            continue
        if r[0] != "":
            cur = (file_path, int(r[0]))
            src_of[cur] = r[1].strip()
            continue
        if cur is None or r[2] in ("...", "-"):
            continue
        try:
            inst, samp = int(r[i_inst]), int(r[i_samp])
        except ValueError:
            continue
        e = addr.get(r[2])
        if e is None or depth(cur[0]) > depth(e["loc"][0]):
            stalls = collections.Counter()
            for i, h in stall_cols:
                try:
                    stalls[h] += int(r[i])
                except ValueError:
                    pass
            addr[r[2]] = {"loc": cur, "inst": inst, "samples": samp, "sass": r[3], "stalls": stalls,
                          "callers": (e["callers"] | {e["loc"]}) if e else set()}
        else:
            e["callers"].add(cur)
    per_line = collections.defaultdict(lambda: {"inst": 0, "samples": 0, "src": "", "stalls": collections.Counter(), "n_sass": 0})
    per_site = collections.Counter()      # kernel-file call sites of inlined header code
    per_pipe, per_pipe_static = collections.Counter(), collections.Counter()
    for e in addr.values():
        d = per_line[e["loc"]]
        d["inst"] += e["inst"]; d["samples"] += e["samples"]; d["n_sass"] += 1; d["stalls"] += e["stalls"]
        d["src"] = src_of.get(e["loc"], "")
        p = pipe_of(e["sass"])
        per_pipe[p] += e["inst"]; per_pipe_static[p] += 1
        if depth(e["loc"][0]) > 0:
            sites = [c for c in e["callers"] if depth(c[0]) == 0]
            if sites:
                per_site[(e["loc"][0].split("/")[-1], max(sites, key=lambda c: c[1]))] += e["inst"]
    tot_i = sum(d["inst"] for d in per_line.values()) or 1
    tot_s = sum(d["samples"] for d in per_line.values()) or 1
    out = sys.stdout
    out.write(f"kernel: `{kernel_name}`\n\nwarp instructions executed (first captured launch): {tot_i}; stall samples: {tot_s}\n\n")
    out.write("### by file\n\n| file | warp instr | % | samples | % | SASS instr |\n|---|---|---|---|---|---|\n")
    by_file = collections.defaultdict(lambda: [0, 0, 0])
    for (f, _), d in per_line.items():
        by_file[f][0] += d["inst"]; by_file[f][1] += d["samples"]; by_file[f][2] += d["n_sass"]
    for f, (i, s, n) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        out.write(f"| `{f.split('/')[-1]}` | {i} | {100 * i / tot_i:.1f} | {s} | {100 * s / tot_s:.1f} | {n} |\n")
    if a.regions:
        regs = []
        for item in a.regions.split(","):
            loc, name = item.split("=")
            f, rng = loc.split(":")
            lo, hi = rng.split("-")
            regs.append((f, int(lo), int(hi), name))
        roll = collections.OrderedDict((name, [0, 0]) for *_, name in regs)
        roll["(elsewhere)"] = [0, 0]
        for (f, ln), d in per_line.items():
            for rf, lo, hi, name in regs:
                if f.endswith("/" + rf) and lo <= ln <= hi:
                    break
            else:
                name = "(elsewhere)"
            roll[name][0] += d["inst"]; roll[name][1] += d["samples"]
        out.write("\n### by code region (innermost source location of every SASS instruction)\n\n| region | warp instr | % | stall samples % |\n|---|---|---|---|\n")
        for name, (i, sm) in roll.items():
            out.write(f"| {name} | {i} | {100 * i / tot_i:.1f} | {100 * sm / tot_s:.1f} |\n")
    out.write("\n### instruction mix (executed warp instructions by issue class)\n\n| class | warp instr | % | SASS instr |\n|---|---|---|---|\n")
    for p, i in per_pipe.most_common():
        out.write(f"| {p} | {i} | {100 * i / tot_i:.1f} | {per_pipe_static[p]} |\n")
    out.write("\n### inlined header code by call site in the kernel file\n\n| header | call site | warp instr | % | source |\n|---|---|---|---|---|\n")
    for (hdr, site), i in per_site.most_common(12):
        src = src_of.get(site, "").replace("|", "\\|")[:110]
        out.write(f"| `{hdr}` | `{site[0].split('/')[-1]}:{site[1]}` | {i} | {100 * i / tot_i:.1f} | `{src}` |\n")
    out.write(f"\n### hottest source lines (top {a.top} by executed warp instructions)\n\n| file:line | warp instr | % | samples % | top stalls | source |\n|---|---|---|---|---|---|\n")
    for (f, ln), d in sorted(per_line.items(), key=lambda kv: -kv[1]["inst"])[:a.top]:
        st = ", ".join(f"{h[6:]} {100 * c / max(d['samples'], 1):.0f}%" for h, c in d["stalls"].most_common(3) if c)
        src = d["src"].replace("|", "\\|")[:110]
        out.write(f"| `{f.split('/')[-1]}:{ln}` | {d['inst']} | {100 * d['inst'] / tot_i:.1f} | {100 * d['samples'] / tot_s:.1f} | {st} | `{src}` |\n")
    out.write(f"\n### most-stalled source lines (top {a.top} by samples)\n\n| file:line | samples % | warp instr % | top stalls | source |\n|---|---|---|---|---|\n")
    for (f, ln), d in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:a.top]:
        st = ", ".join(f"{h[6:]} {100 * c / max(d['samples'], 1):.0f}%" for h, c in d["stalls"].most_common(3) if c)
        src = d["src"].replace("|", "\\|")[:110]
        out.write(f"| `{f.split('/')[-1]}:{ln}` | {100 * d['samples'] / tot_s:.1f} | {100 * d['inst'] / tot_i:.1f} | {st} | `{src}` |\n")


if __name__ == "__main__":
    main()
