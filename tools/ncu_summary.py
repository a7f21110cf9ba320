"""Summarise an .ncu-rep (read here, no GPU): key metrics + hot-loop instruction mix per warp point-evaluation.
usage: python tools/ncu_summary.py report.ncu-rep [warp_point_evals]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
wpe = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.05:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
iS, iN = h2.index("Source"), h2.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iN and r[iN].isdigit()]
tot = sum(int(r[iN]) for r in data)
mx = max(int(r[iN]) for r in data)
hot = [r for r in data if int(r[iN]) > 0.25 * mx]
print(f"total warp-instructions {tot:.4g}; hot-loop instructions {len(hot)} covering {sum(int(r[iN]) for r in hot) / tot:.1%}")
ops = collections.Counter()
for r in hot:
    t = r[iS].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += int(r[iN])
den = wpe or mx
print("per warp point-evaluation:" if wpe else "relative to hottest instruction:")
print("  " + "  ".join(f"{k} {v / den:.2f}" for k, v in ops.most_common(24)))
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"  fp64-pipe {fp64 / den:.2f}  all {sum(ops.values()) / den:.2f}")
