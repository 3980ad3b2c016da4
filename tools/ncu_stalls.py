"""Overall warp-stall distribution of an ncu report (sum over all SASS instructions)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for _, h in cols}
ninst = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ninst += 1
    for i, h in cols:
        if r[i].isdigit():
            tot[h] += int(r[i])
s = sum(tot.values())
print("SASS instructions in kernel:", ninst, " total samples:", s)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {h[6:]:20s} {100 * v / s:5.1f}%")
