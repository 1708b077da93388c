"""Split a kernel's executed instructions by phase, using WARPSYNC / BAR / SYNCS instructions in the SASS
stream as phase markers.  usage: python tools/ncu_by_phase.py report.ncu-rep frames"""
import csv, subprocess, sys, re, collections, io
rep = sys.argv[1]; frames = float(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
col = {n: i for i, n in enumerate(rows[hi])}
data = [r for r in rows[hi + 1:] if len(r) > 10]
phase = 0; agg = collections.OrderedDict()
for r in data:
    src = r[col["Source"]].strip()
    m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", src); op = m.group(1) if m else "?"
    e = int(r[col["Instructions Executed"]]); sm = int(r[col["# Samples"]])
    a = agg.setdefault(phase, {"inst": 0, "samp": 0, "ops": collections.Counter(), "first": src})
    a["inst"] += e; a["samp"] += sm; a["ops"][op] += e
    if op == "BAR" or (op == "BRA" and "BRA.DIV" in src) or (op == "SYNCS" and "TRYWAIT" in src): phase += 1
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
print("total %.1f warp-instr/frame" % (ti / frames))
for ph, a in agg.items():
    if a["inst"] == 0: continue
    print("phase %2d: %6.1f/frame (%4.1f%% inst, %4.1f%% samples)  %s" % (ph, a["inst"] / frames, 100.0 * a["inst"] / ti, 100.0 * a["samp"] / max(ts, 1),
          {k: round(v / frames, 1) for k, v in a["ops"].most_common(9)}))
