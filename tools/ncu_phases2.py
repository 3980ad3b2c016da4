"""Like ncu_phases.py, plus the STATIC size (SASS instructions) of every phase and its stall-sample mix.
usage: python tools/ncu_phases2.py rep.ncu-rep file.cuh:lo-hi:name ..."""
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
tot, samp, stat, noinst = {}, {}, {}, {}
cur_file = "?"
for r in rows:
    if r and r[0] == "File Path" and len(r) > 1:
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        iN = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
        iSrc = hdr.index("Source")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    name = "other:" + cur_file
    for f, lo, hi, nm in phases:
        if f == cur_file and lo <= ln <= hi:
            name = nm
            break
    if r[2] != "-":            # a SASS row under its source line
        stat[name] = stat.get(name, 0) + 1
        continue
    try:
        ins, smp = int(r[iI]), int(r[iS])
    except ValueError:
        continue
    tot[name] = tot.get(name, 0) + ins
    samp[name] = samp.get(name, 0) + smp
T, S = sum(tot.values()), sum(samp.values())
print(f"{'phase':28s} {'warp-inst':>12s} {'%':>6s} {'samples%':>9s} {'SASS':>6s}")
for k in sorted(tot, key=lambda k: -samp.get(k, 0)):
    print(f"{k:28s} {tot[k]:12d} {100 * tot[k] / T:5.1f}% {100 * samp[k] / max(S, 1):8.1f}% {stat.get(k, 0):6d}")
print("total", T, "samples", S, "SASS", sum(stat.values()))
