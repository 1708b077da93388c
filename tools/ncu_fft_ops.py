"""Executed instructions per 4-frame FFT pass by opcode (FFT-warp branch of K1) and per tile for the epilogue branch.
usage: python tools/ncu_fft_ops.py report.ncu-rep frames"""
import csv, sys, re, collections, subprocess, io
rep, frames = sys.argv[1], float(sys.argv[2])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
col = {n: i for i, n in enumerate(rows[hi])}; data = [r for r in rows[hi + 1:] if len(r) > 10]
sec = "pre"; ops = {"pre": collections.Counter(), "epi": collections.Counter(), "fft": collections.Counter()}
for r in data:
    src = r[col["Source"]].strip(); e = int(r[col["Instructions Executed"]])
    if "USETMAXREG.DEALLOC" in src: sec = "epi"
    if "USETMAXREG.TRY_ALLOC" in src: sec = "fft"
    m = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", src); ops[sec][m.group(1) if m else "?"] += e
for sec, unit, div in (("fft", "4-frame pass", frames / 4), ("epi", "32-frame tile", frames / 32)):
    tot = sum(ops[sec].values()); packed = sum(v for k, v in ops[sec].items() if k in ("FADD2", "FMUL2", "FFMA2"))
    print("%s branch: %.0f instr per %s (%.1f per frame); packed f32x2 %.0f -> issue cycles (packed x2) %.1f per frame" % (
        sec, tot / div, unit, tot / frames, packed / div, (tot + packed) / frames))
    print("   " + "  ".join("%s %.1f" % (k, v / div) for k, v in ops[sec].most_common(28)))
