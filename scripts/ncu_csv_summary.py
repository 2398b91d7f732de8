"""Summarise an `ncu --page raw --csv` export (one kernel per row) into the few numbers the bench and DESIGN.md quote.
    python scripts/ncu_csv_summary.py gpurun_out/x.csv [out.txt]"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sass__inst_executed_register_spilling",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "derived__memory_l2_theoretical_sectors_global_excessive",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = [f"# ncu --set full --clock-control none, raw page: {sys.argv[1]}", ""]
for r in rows[2:]:
    out.append("## " + r[hdr.index("Kernel Name")])
    for i, h in enumerate(hdr):
        if h in KEYS:
            out.append(f"{h:75s} {r[i]:>20s} {units[i]}")
    out.append("# warp stall reasons per issued instruction (ratio > 0.05)")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            try:
                if float(r[i]) > 0.05:
                    out.append(f"{h:75s} {float(r[i]):20.3f}")
            except ValueError:
                pass
    out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
