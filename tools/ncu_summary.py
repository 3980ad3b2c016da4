"""Writes a text summary of an ncu report (key metrics, stall reasons, opcode mix, hottest lines)."""
import csv, subprocess, sys, io, contextlib, runpy, os
rep, out = sys.argv[1], sys.argv[2]
HERE = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__cycles_elapsed.max", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
with open(out, "w") as fh:
    fh.write(f"# ncu summary of {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on)\n\n## key metrics\n")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            fh.write(f"{h:70s} {v} {u}\n")
    for title, script, extra in [("stall reasons (all SASS instructions)", "ncu_stalls.py", []),
                                 ("executed instructions by opcode", "ncu_opcodes.py", []),
                                 ("hottest source lines by stall samples", "ncu_lines.py", ["25"]),
                                 ("hottest source lines by executed instructions", "ncu_lines.py", ["25", "inst"])]:
        r = subprocess.run([sys.executable, os.path.join(HERE, script), rep] + extra, capture_output=True, text=True).stdout
        fh.write(f"\n## {title}\n{r}")
print("wrote", out)
