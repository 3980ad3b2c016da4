"""Executed warp-instructions and stall samples per source line of ONE file, in line order.
usage: python tools/ncu_range.py rep.ncu-rep file.cuh lo hi"""
import csv
import subprocess
import sys

rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur, tot_i, tot_s, sel = None, "?", 0, 0, []
for r in rows:
    if r and r[0] == "File Path" and len(r) > 1:
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        ln, ins, smp = int(r[0]), int(r[iI]), int(r[iS])
    except ValueError:
        continue
    tot_i += ins
    tot_s += smp
    if cur == fname and lo <= ln <= hi and ins > 0:
        sel.append((ln, ins, smp, r[1].strip()[:100]))
si = sum(x[1] for x in sel)
ss = sum(x[2] for x in sel)
for ln, ins, smp, src in sorted(sel):
    print(f"{ln:5d} {ins / 1e6:8.2f}M {100 * ins / tot_i:5.2f}% samp {100 * smp / tot_s:5.2f}% | {src}")
print(f"range total {si / 1e6:.1f}M {100 * si / tot_i:.1f}% samples {100 * ss / tot_s:.1f}%")
