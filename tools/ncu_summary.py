"""Summarises an .ncu-rep (raw + source pages) into the few numbers DESIGN.md/profiles cite.  Runs on the CPU box."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]
for k, row in enumerate(data):
    print(f"--- launch {k}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:72s} {row[i]:>22s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
body = []
for r in rows[2:]:
    if len(r) != len(h) or r[0] == "Address":
        break
    body.append(r)
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: 0.0 for c in cols}
for r in body:
    for c in cols:
        try:
            tot[c] += float(r[h.index(c)])
        except ValueError:
            pass
s = sum(tot.values()) or 1.0
print("--- warp stall samples (launch 0, all samples)")
for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print(f"{c:28s} {100 * v / s:6.1f}%")
