"""Writes profiles/<tag>_k1_summary.md from an ncu --set full report of k_frames_to_statics.
usage: python tools/make_profile_summary.py report.ncu-rep frames tag"""
import csv, io, subprocess, sys, os, json
rep, frames, tag = sys.argv[1], float(sys.argv[2]), sys.argv[3]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; u = rows[1]; v = rows[2]
def g(n):
    return (v[h.index(n)], u[h.index(n)]) if n in h else ("n/a", "")
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
out = ["# %s -- ncu summary of `k_frames_to_statics` (%s)" % (tag, os.path.basename(rep)), "",
       "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200;",
       "%d frames in the profiled launch.  Numbers under a profiler are for shares and counters, not for timing claims." % frames, "",
       "| metric | value |", "|---|---|"]
for k in keys:
    val, unit = g(k)
    out.append("| `%s` | %s %s |" % (k, val, unit))
def num(n):
    try: return float(g(n)[0].replace(",", ""))
    except Exception: return float("nan")
rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
ru, wu = g("dram__bytes_read.sum")[1], g("dram__bytes_write.sum")[1]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tot = rd * scale.get(ru, 1) + wr * scale.get(wu, 1)
cyc = num("sm__cycles_elapsed.max")
out += ["", "DRAM traffic per frame: **%.1f B** (algorithmic 372 B: 320 B int16 in + 52 B statics out; statics blocks are padded to 32 frames)." % (tot / frames),
        "Executed warp-instructions per frame: **%.1f**;  shared-memory wavefronts per frame: **%.1f**;  cycles per frame per SMSP: **%.0f**." % (
            num("smsp__inst_executed.sum") / frames, num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / frames, cyc * 592 / frames), ""]
ops = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_fft_ops.py"), rep, str(frames)], capture_output=True, text=True).stdout
out += ["## executed instructions by warp role", "",
        "A packed f32x2 instruction occupies the SMSP's issue port for two cycles (tools/ubench_issue.cu), so the",
        "issue-cycle figure counts FADD2 / FMUL2 / FFMA2 twice; the kernel is bound by it.", "", "```", ops.strip(), "```", ""]
ss = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_source_summary.py"), rep], capture_output=True, text=True).stdout
out += ["## stall mix / opcode mix / top stall sites", "", "```", "\n".join(ss.strip().split("\n")[:45]), "```", ""]
open(os.path.join(root, "profiles", "%s_k1_summary.md" % tag), "w").write("\n".join(out))
json.dump({"dram_bytes_per_frame": tot / frames, "source": "profiles/%s_k1_summary.md (ncu dram__bytes_read.sum + dram__bytes_write.sum / frames)" % tag},
          open(os.path.join(root, "profiles", "k1_traffic.json"), "w"))
print("\n".join(out[:32]))
