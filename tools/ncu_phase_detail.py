"""Per-phase detail of a K1 ncu report: instructions, samples, shared wavefronts (total / ideal), stall mix.
usage: python tools/ncu_phase_detail.py report.ncu-rep frames [--dump]"""
import csv, subprocess, sys, re, collections, io
rep = sys.argv[1]; frames = float(sys.argv[2]); dump = "--dump" in sys.argv
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
col = {n: i for i, n in enumerate(rows[hi])}
data = [r for r in rows[hi + 1:] if len(r) > 10]
stall_cols = [n for n in rows[hi] if n.startswith("stall_") and "Not Issued" not in n]
phase = 0; agg = collections.OrderedDict()
for r in data:
    src = r[col["Source"]].strip()
    m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", src); op = m.group(1) if m else "?"
    e = int(r[col["Instructions Executed"]]); sm = int(r[col["# Samples"]])
    a = agg.setdefault(phase, {"inst": 0, "samp": 0, "wf": 0, "wfi": 0, "st": collections.Counter(), "wfop": collections.Counter()})
    a["inst"] += e; a["samp"] += sm
    wf = int(r[col["L1 Wavefronts Shared"]] or 0); wfi = int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    a["wf"] += wf; a["wfi"] += wfi
    if wf: a["wfop"][re.sub(r"\s+.*", "", src.split(" R")[0])[:12]] += wf
    for n in stall_cols: a["st"][n[6:]] += int(r[col[n]])
    if dump: print("%2d %9d %6d %8d  %s" % (phase, e, sm, wf, src))
    if op == "BAR" or (op == "BRA" and "BRA.DIV" in src) or (op == "SYNCS" and "TRYWAIT" in src): phase += 1
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values()); tw = sum(a["wf"] for a in agg.values())
print("total %.1f warp-instr/frame, %.1f shared wavefronts/frame" % (ti / frames, tw / frames))
for ph, a in agg.items():
    if a["inst"] == 0: continue
    print("phase %2d: %6.1f inst/frame %5.1f%% samples  wf %5.1f/frame (ideal %5.1f)  cyc/inst %.2f  stalls %s  wf by op %s" % (
        ph, a["inst"] / frames, 100.0 * a["samp"] / max(ts, 1), a["wf"] / frames, a["wfi"] / frames,
        (a["samp"] / max(ts, 1)) / max(a["inst"] / ti, 1e-9),
        {k: round(100.0 * v / max(a["samp"], 1)) for k, v in a["st"].most_common(5)},
        {k: round(v / frames, 1) for k, v in a["wfop"].most_common(4)}))
