"""Summarise one kernel of an .ncu-rep: key raw metrics, SASS opcode mix and stall reasons.
usage: ncu_summary.py rep [kernel-name-regex]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ksel = ["-k", "regex:" + sys.argv[2]] if len(sys.argv) > 2 else []
raw = subprocess.run(["ncu", "-i", rep] + ksel + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed.sum", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:70s} {vals[i]} {unit[i]}")
src = subprocess.run(["ncu", "-i", rep] + ksel + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


ti = sum(f(r, "Instructions Executed") for r in data)
ts = sum(f(r, "# Samples") for r in data)
by = collections.defaultdict(lambda: [0, 0])
for r in data:
    op = [t for t in r[ix["Source"]].strip().split() if not t.startswith("@")]
    name = ".".join(op[0].split(".")[:2]) if op else "?"
    by[name][0] += f(r, "Instructions Executed")
    by[name][1] += f(r, "# Samples")
print(f"warp instructions {ti:.3e}; samples {ts:.0f}")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"  {k:24s} inst {100 * v[0] / ti:5.1f}%  samples {100 * v[1] / ts:5.1f}%")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stalls % of samples:", {s[6:]: round(100 * sum(f(r, s) for r in data) / ts, 1) for s in st if sum(f(r, s) for r in data) / ts > 0.004})
