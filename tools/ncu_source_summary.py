"""Summarise an ncu source-page CSV: python tools/ncu_source_summary.py report.ncu-rep [kernel-regex]
Joins SASS addresses with nvdisasm -g line info from the in-tree .so when possible."""
import csv, subprocess, sys, re, collections, io, os
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hdr_i]; data = rows[hdr_i + 1:]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in data if len(r) > 10)
tot_samp = sum(int(r[col["# Samples"]]) for r in data if len(r) > 10)
print("total warp instructions executed:", tot_inst, " samples:", tot_samp)
st = collections.Counter()
for r in data:
    if len(r) <= 10: continue
    for n in stall_cols: st[n] += int(r[col[n]])
print("stall mix:", {k: round(100.0 * v / max(tot_samp, 1), 1) for k, v in st.most_common(10)})
op = collections.Counter(); ops = collections.Counter()
for r in data:
    if len(r) <= 10: continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]])
    if m: op[m.group(1)] += int(r[col["Instructions Executed"]]); ops[m.group(1)] += int(r[col["# Samples"]])
print("executed by opcode (% inst, % samples):")
for k, v in op.most_common(25): print("  %-10s %5.1f %5.1f" % (k, 100.0 * v / tot_inst, 100.0 * ops[k] / max(tot_samp, 1)))
exc = [(int(r[col["L1 Wavefronts Shared Excessive"]]), r[col["Source"]].strip()) for r in data if len(r) > 10 and int(r[col["L1 Wavefronts Shared Excessive"]]) > 0]
print("shared excessive wavefronts: total", sum(e for e, _ in exc))
for e, s in sorted(exc, reverse=True)[:12]: print("  ", e, s)
print("top stall instructions:")
top = sorted((r for r in data if len(r) > 10), key=lambda r: -int(r[col["# Samples"]]))[:25]
for r in top:
    ss = {n[6:]: int(r[col[n]]) for n in stall_cols if int(r[col[n]]) > 0}
    print("  %5.2f%% %-60s %s" % (100.0 * int(r[col["# Samples"]]) / max(tot_samp, 1), r[col["Source"]].strip()[:60], dict(sorted(ss.items(), key=lambda x: -x[1])[:3])))
