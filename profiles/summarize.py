"""Summarise an ncu report (--set full) into a few lines of JSON: python profiles/summarize.py report.ncu-rep [units]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid", "launch__block_size": "block", "inst_executed": "warp_instructions",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct2",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts", "sm__cycles_elapsed.avg": "cycles",
}
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")][:80]}
    for i, h in enumerate(hdr):
        if h in want:
            d[want[h]] = f"{r[i]} {units[i]}"
    out.append(d)
print(json.dumps(out, indent=1))
