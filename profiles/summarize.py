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
    # cycles-active (not instruction-issue) view of the pipes, issue-slot and L1 wavefront utilisation
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_cycles_active_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "fmaheavy_cycles_active_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_cycles_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_cycles_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.avg.per_cycle_active": "ipc_per_scheduler",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_cycles_active_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_pipe_inst_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_cycles_active_pct",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active": "uniform_pipe_pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle_per_issue",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct": "stall_barrier_pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct": "stall_short_scoreboard_pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct": "stall_long_scoreboard_pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct": "stall_math_pipe_pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct": "stall_mio_throttle_pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct": "stall_lg_throttle_pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct": "stall_wait_pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct": "stall_not_selected_pct",
}
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")][:80]}
    for i, h in enumerate(hdr):
        if h in want:
            d[want[h]] = f"{r[i]} {units[i]}"
    out.append(d)
print(json.dumps(out, indent=1))
