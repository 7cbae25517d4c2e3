"""The counters of one `ncu --set full` capture that profiles/README.md quotes, one per line (name unit value), plus the stall mix.
    python tools/ncu_summary.py gpurun_out/k2_r2b_full.ncu-rep > profiles/k2_r2_summary.txt"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__sass_inst_executed_op_local_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
kn = h.index("Kernel Name")
for v in rows[2:]:                     # one row per captured launch
    if len(rows) > 3:
        print("==", v[kn].split("(")[0], "grid", v[h.index("Grid Size")], "block", v[h.index("Block Size")])
    for w in WANT:
        for i, x in enumerate(h):
            if x == w:
                print(w, u[i], v[i])
    stalls = [(float(v[i]), x) for i, x in enumerate(h) if x.startswith("smsp__average_warps_issue_stalled") and v[i]]
    for s_, x in sorted(stalls, reverse=True)[:8]:
        print(x, "%.3f" % s_)
