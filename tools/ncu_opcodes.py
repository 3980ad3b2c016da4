"""Executed-instruction histogram by SASS opcode from an ncu report (source page, SASS view)."""
import csv, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iI = hdr.index("Instructions Executed"); iS = hdr.index("Source")
h = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr) or not r[iI].isdigit():
        continue
    src = r[iS].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    h[op.split(".")[0]] += int(r[iI])
tot = sum(h.values())
print("total", tot)
for op, v in h.most_common(40):
    print(f"  {op:12s} {100*v/tot:5.1f}%")
