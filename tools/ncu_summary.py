"""Summarise an .ncu-rep (read here, no GPU): key per-kernel metrics as a small table.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra-regex]
"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_bytes.sum", "sm__pipe_tmem_cycles_active.avg.pct_of_peak_sustained_elapsed",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in KEYS:
        for h in hdr:
            if h == k or h.endswith("." + k):
                print(f"{h:95s} {d[h]:>18s} {units[hdr.index(h)]}")
                break
    if extra:
        pat = re.compile(extra)
        for h in hdr:
            if pat.search(h):
                print(f"{h:95s} {d[h]:>18s} {units[hdr.index(h)]}")
