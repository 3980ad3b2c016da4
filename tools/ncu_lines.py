"""Summarise an ncu report by CUDA source line: share of executed instructions and of stall samples.
usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue            # keep only the per-CUDA-line aggregate rows (Address == '-')
    try:
        ins, smp = int(r[iI]), int(r[iS])
    except ValueError:
        continue
    stalls = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall_cols), reverse=True)[:3]
    data.append((ins, smp, r[0], r[1], stalls))
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total warp-instructions {ti}  samples {ts}")
key = (lambda d: -d[0]) if (len(sys.argv) > 3 and sys.argv[3] == "inst") else (lambda d: -d[1])
for ins, smp, ln, src, stalls in sorted(data, key=key)[:top]:
    st = " ".join(f"{h[6:]}:{v}" for v, h in stalls if v)
    print(f"{ln:>4} inst {100 * ins / ti:5.1f}% samp {100 * smp / ts:5.1f}% | {src.strip()[:84]:84} | {st}")
