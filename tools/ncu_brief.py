#!/usr/bin/env python
"""Short summary of an .ncu-rep: duration, DRAM bytes, pipe / L1 / issue utilisation and the top stall reasons per kernel.
usage: tools/ncu_brief.py file.ncu-rep [...]"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64_op_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    for row in rows[2:]:
        print("==", f, row[h.index("Kernel Name")][:70])
        for k in KEYS:
            if k in h and row[h.index(k)] not in ("", "n/a"):
                print("   %-85s %s" % (k, row[h.index(k)]))
        st = [(float(row[i].replace(",", "")), k) for i, k in enumerate(h) if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and row[i] not in ("", "n/a")]
        for v, k in sorted(st, reverse=True)[:7]:
            print("   stall %-40s %.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
