"""Launch list (ncu --metrics gpu__time_duration.sum,dram__bytes_* --csv) -> markdown table of this repo's kernels
and a CSV restricted to them.  python tools/launch_summary.py in.csv tag"""
import csv, sys, os, statistics
src, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = open(src).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
ours = [r for r in rows if r["Kernel Name"].startswith(("k_", "void k_", "fe::k_", "void fe::k_"))]
launch = {}
for r in ours:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "").replace("fe::", ""), "grid": r["Grid Size"], "block": r["Block Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    else:
        d[r["Metric Name"]] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
with open(os.path.join(root, "profiles", tag + "_launches.csv"), "w") as f:
    f.write("id,kernel,grid,block,ms,dram_read_bytes,dram_write_bytes\n")
    for i, d in launch.items():
        f.write("%s,%s,\"%s\",\"%s\",%.6f,%d,%d\n" % (i, d["name"], d["grid"], d["block"], d.get("ms", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)))
by_all = {}
for d in launch.values():
    by_all.setdefault(d["name"], []).append(d)
# the device-resident bench steps are the whole-shard launches; the end-to-end leg re-launches the same kernels on
# 192 MB chunks of the shard (host pipeline).  Shares are taken over the whole-shard launches.
by, chunked = {}, {}
for k, v in by_all.items():
    top = max(x.get("ms", 0) for x in v)
    by[k] = [x for x in v if x.get("ms", 0) >= 0.5 * top]
    chunked[k] = len(v) - len(by[k])
step = sum(statistics.median(x.get("ms", 0) for x in v) for k, v in by.items() if not k.startswith("k_fp32_peak"))
out = ["# %s -- launch list of the bench (ncu --metrics gpu__time_duration.sum,dram__bytes_*; serialised, cold cache: compare shares)" % tag, "",
       "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline` "
       "(this repo's kernels only: %s_launches.csv; torch's generator kernels that build the synthetic shard are left out)." % tag, "",
       "| kernel | whole-shard launches | median ms | share of step | DRAM read GB | DRAM write GB | chunked launches (e2e leg) |", "|---|---|---|---|---|---|---|"]
for k, v in by.items():
    ms = statistics.median(x.get("ms", 0) for x in v)
    share = "%.1f %%" % (100 * ms / step) if not k.startswith("k_fp32_peak") else "(peak probe)"
    out.append("| `%s` | %d | %.3f | %s | %.2f | %.2f | %d |" % (k, len(v), ms, share, statistics.median(x.get("dram__bytes_read.sum", 0) for x in v) / 1e9,
                                                          statistics.median(x.get("dram__bytes_write.sum", 0) for x in v) / 1e9, chunked[k]))
open(os.path.join(root, "profiles", tag + "_launches.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
