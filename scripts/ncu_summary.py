"""Key numbers of an ncu --set full report as markdown (run here, no GPU):  python scripts/ncu_summary.py rep [title]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
v = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_tma_ld.sum", "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_branch.sum",
]
print("```")
for k in want:
    if k in v:
        print(f"{k:72s} {v[k][0]:>20s} {v[k][1]}")
stalls = sorted(((float(v[h][0]), h.split('issue_stalled_')[1].split('_per_')[0]) for h in v if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h), reverse=True)
tot = sum(s for s, _ in stalls)
print("stall shares (warps per issue-active cycle, share of all): " + ", ".join(f"{n} {s / tot * 100:.0f}%" for s, n in stalls[:8]))
print("```")
