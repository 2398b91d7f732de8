"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py full  gpurun_out/prof_X.ncu-rep   profiles/r1_strip_full.txt
    python profiles/summarize.py list  gpurun_out/launches_X.csv   profiles/r1_launches.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sass__inst_executed_register_spilling",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, dst):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {rep}", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"## {name}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                lines.append(f"{h:75s} {r[i]:>18s} {units[i]}")
        lines.append("# warp stall reasons per issued instruction (ratio > 0.05)")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    if float(r[i]) > 0.05:
                        lines.append(f"{h:75s} {float(r[i]):18.3f}")
                except ValueError:
                    pass
        lines.append("")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        hdr = src[1]
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        ops, samp = collections.Counter(), collections.Counter()
        for r in src[2:]:
            if len(r) <= max(iS, iI) or not r[iI].isdigit():
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += int(r[iI])
            samp[op] += int(r[iS])
        tot = max(1, sum(samp.values()))
        lines.append("# SASS opcode mix of the first captured kernel (warp-level instructions executed, share of stall samples)")
        for op, c in ops.most_common(16):
            lines.append(f"{op:10s} {c:14d} {100 * samp[op] / tot:6.1f}%")
    open(dst, "w").write("\n".join(lines) + "\n")


def launch_list(path, dst):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        agg[r[4]][0] += 1
        agg[r[4]][1] += float(r[-1]) / 1e6
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list (gpu__time_duration.sum, --clock-control none) from {path}",
             "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
             f"# {len(rows)} launches, {tot:.3f} ms total", f"{'kernel':90s} {'launches':>8s} {'ms total':>10s} {'ms avg':>9s} {'share':>7s}"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:90]:90s} {n:8d} {ms:10.3f} {ms / n:9.4f} {100 * ms / tot:6.1f}%")
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
