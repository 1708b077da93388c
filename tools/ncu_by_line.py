"""Executed warp-instructions and stall samples per SOURCE LINE: joins an ncu report's SASS page with
nvdisasm -g line info of the in-tree library (must be the same build that was profiled).
usage: python tools/ncu_by_line.py report.ncu-rep kernel-substring [frames]"""
import csv, subprocess, sys, re, collections, io, os, tempfile, glob
rep, ksub = sys.argv[1], sys.argv[2]
frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "automatic-speech-recognition_b200", "libasr_frontend.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and ksub in l][0]
end = [i for i, l in enumerate(dis) if l.startswith("//--------------------- ") and i > start][0]
off2line = {}; cur = None
for ln in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+\S', ln)
    if m and cur: off2line[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
col = {n: i for i, n in enumerate(rows[hi])}
data = [r for r in rows[hi + 1:] if len(r) > 10]
a0 = int(data[0][col["Address"]], 16)
inst = collections.Counter(); samp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in data:
    off = int(r[col["Address"]], 16) - a0
    key = off2line.get(off, ("?", 0))
    e = int(r[col["Instructions Executed"]])
    inst[key] += e; samp[key] += int(r[col["# Samples"]])
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", r[col["Source"]]); ops[key][m.group(1) if m else "?"] += e
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-instr %d  samples %d" % (ti, ts))
for key, v in sorted(inst.items(), key=lambda x: -x[1])[:45]:
    per = (" %6.1f/frame" % (v / frames)) if frames else ""
    print("%-18s:%-4d inst %5.1f%%%s  samples %5.1f%%  %s" % (key[0], key[1], 100.0 * v / ti, per, 100.0 * samp[key] / max(ts, 1),
          dict((k, round(c / (frames or 1), 1) if frames else c) for k, c in ops[key].most_common(4))))
