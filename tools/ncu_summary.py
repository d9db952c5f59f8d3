"""Prints the handful of ncu metrics we track from a .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'smsp__inst_executed.sum','sm__cycles_elapsed.avg','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__block_size','launch__occupancy_limit_registers',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio']
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
    for k in keys:
        if k in hdr:
            d[k] = (r[hdr.index(k)], units[hdr.index(k)])
    res.append(d)
for d in res:
    print("==", d["kernel"][:90])
    for k in keys:
        if k in d: print(f"  {k:95s} {d[k][0]:>16s} {d[k][1]}")
