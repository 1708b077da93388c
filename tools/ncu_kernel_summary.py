"""Markdown table of the key ncu metrics for every kernel in a report: python tools/ncu_kernel_summary.py rep.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, units = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp inst")]
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
    vals = []
    for k, _ in cols:
        if k in h:
            v, u = r[h.index(k)], units[h.index(k)]
            try: v = "%.4g" % float(v.replace(",", ""))
            except ValueError: pass
            vals.append((v + " " + u).strip())
        else:
            vals.append("n/a")
    print("| `%s` | " % name + " | ".join(vals) + " |")
