"""Instruction / stall-sample share per kernel phase (source line ranges of bigclam_kernels.cuh)."""
import csv, subprocess, sys, re
rep, src = sys.argv[1], sys.argv[2]
lines = open(src).read().split("\n")
def find(pat):
    for i, l in enumerate(lines, 1):
        if pat in l:
            return i
    raise KeyError(pat)
marks = [("helpers(exp/log/clamp/reduce)", 1), ("kernel prologue", find("__global__ void __launch_bounds__")),
         ("next-node loads + fusf/fufu", find("issue the next nodes")), ("PRE dots/eval/axpy", find("PRE (:157")),
         ("grad finalize + active set", find("grad = (sum - sumF)")), ("own terms + table", find("LS, active-space path")), ("tile gather", find("gather: val[e][t]")),
         ("tile dot", find("dot: lane (j, h)")), ("exp/log + decision", find("exp/log: two independent chains")),
         ("swap/write", find("SWAP (:183")), ("rotate + epilogue", find("rotate the pipeline")), ("end", 10**9)]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try: data.append((int(r[0]), int(r[iI]), int(r[iS]), r[1]))
    except ValueError: pass
ti = sum(d[1] for d in data); ts = sum(d[2] for d in data)
helper_names = {"exp/log/rcp": (find("__device__ __forceinline__ double exp_neg"), find("struct EdgeConst")),
                "edge_term": (find("struct EdgeConst"), find("// step(), bigclam4-7.scala:110-113")),
                "clamp": (find("// step(), bigclam4-7.scala:110-113"), find("// L2 prefetch of the rows")),
                "prefetch": (find("// L2 prefetch of the rows"), find("// Sum R per-lane partials")),
                "reduce_to_lanes+chunk_dots": (find("// Sum R per-lane partials"), find("// Dense line search (cold path)")),
                "warp_sum/ldg2": (find("__device__ __forceinline__ double warp_sum"), find("__device__ __forceinline__ double exp_neg"))}
print(f"total warp-inst {ti}, per-node {ti/334863:.0f}")
for name, (a, b) in helper_names.items():
    i = sum(d[1] for d in data if a <= d[0] < b); s = sum(d[2] for d in data if a <= d[0] < b)
    print(f"  helper {name:28s} inst {100*i/ti:5.1f}% ({i/334863:6.0f}/node)  samples {100*s/ts:5.1f}%")
for (name, a), (_, b) in zip(marks[1:], marks[2:]):
    i = sum(d[1] for d in data if a <= d[0] < b); s = sum(d[2] for d in data if a <= d[0] < b)
    print(f"  {name:36s} inst {100*i/ti:5.1f}% ({i/334863:6.0f}/node)  samples {100*s/ts:5.1f}%")
oth = sum(d[1] for d in data if d[0] >= marks[-2][1] + 60 or False)
