#!/usr/bin/env python
"""Print the handful of ncu raw metrics the DESIGN/roofline discussion uses.  usage: ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__pcsamp_sample_count"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name"), "id", d.get("ID"))
    for k in KEYS:
        if k in d:
            print("  %-70s %s %s" % (k, d[k], rows[1][hdr.index(k)]))
    tot = float(d.get("smsp__pcsamp_sample_count", "0").replace(",", "") or 0)
    st = [(float(d[k].replace(",", "")), k) for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and d[k]]
    for v, k in sorted(st, reverse=True)[:9]:
        print("  stall %-40s %5.1f %%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100.0 * v / max(tot, 1)))
