"""Print the key metrics of an .ncu-rep (first profiled kernel). Dev tool."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "sm__inst_executed_pipe_lsu.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__inst_executed_op_ldgsts.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
d = dict(zip(hdr, zip(units, vals)))
for k in want:
    if k in d: print(f"{k:80s} {d[k][1]} {d[k][0]}")
st = [(float(v[1]), k) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
tot = sum(x for x, _ in st) or 1
print("stall samples:", ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_','')} {100*x/tot:.1f}%" for x, k in sorted(st, reverse=True)[:9]))
