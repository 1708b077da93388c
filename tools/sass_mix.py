"""Static SASS opcode histogram of one kernel of libasr_frontend.so (no GPU needed).

    python tools/sass_mix.py [substring of the mangled name] [--lib path] [--dump]

Used to check instruction-count work on K1 before spending GPU time: the FFT warp's loop body is
straight-line code, so the static count of FADD2 / FFMA2 / FMUL2 / STS / LDS in the kernel tracks the
executed mix per 4-frame pass (profiles/*_k1_summary.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            body[cur].append(line)
    return body


def opcode(line):
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    return m.group(1) if m else None


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    pat = args[0] if args else "k_frames_to_staticsILi400ELi160ELi0ELi0ELi1E"
    lib = os.path.join(ROOT, "automatic-speech-recognition_b200", "libasr_frontend.so")
    if "--lib" in sys.argv:
        lib = sys.argv[sys.argv.index("--lib") + 1]
    for name, lines in functions(lib).items():
        if pat not in name:
            continue
        h = collections.Counter()
        for l in lines:
            op = opcode(l)
            if op:
                h[op.split(".")[0] + (".128" if ".128" in op else ".64" if ".64" in op and op.startswith(("LDS", "STS")) else "")] += 1
        print(name, "instructions:", sum(h.values()))
        print("  " + "  ".join("%s %d" % kv for kv in h.most_common(40)))
        if "--dump" in sys.argv:
            print("\n".join(lines))


if __name__ == "__main__":
    main()
