"""Executed warp-instructions of an ncu report grouped by source file and line ranges ("phases").
usage: python tools/ncu_phases.py rep.ncu-rep file.cuh:lo-hi:name [file.cuh:lo-hi:name ...]   (unmatched lines -> 'other')"""
import csv
import subprocess
import sys

rep = sys.argv[1]
phases = []
for spec in sys.argv[2:]:
    f, rng, name = spec.split(":")
    lo, hi = rng.split("-")
    phases.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
tot = {}
samp = {}
cur_file = "?"
for r in rows:
    if r and r[0] == "File Path" and len(r) > 1:
        cur_file = r[1].split("/")[-1]
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
    name = "other:" + cur_file
    for f, lo, hi, nm in phases:
        if f == cur_file and lo <= ln <= hi:
            name = nm
            break
    tot[name] = tot.get(name, 0) + ins
    samp[name] = samp.get(name, 0) + smp
T, S = sum(tot.values()), sum(samp.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:28s} inst {tot[k]:12d} {100 * tot[k] / T:5.1f}%   samples {100 * samp[k] / max(S, 1):5.1f}%")
print("total", T)
