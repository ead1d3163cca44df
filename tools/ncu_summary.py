#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into the few numbers the
roofline argument needs.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sass__thread_inst_executed_true_per_opcode", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "sass__inst_executed_shared_loads",
    "sass__inst_executed_shared_stores", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat and not pat.search(r[name_i]):
            continue
        print("== kernel:", r[name_i][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-82s %s %s" % (k, r[i], units[i]))
        print("  -- warp stall reasons (average warps stalled per issue-active cycle) --")
        st = []
        for i, k in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", k)
            if m:
                try:
                    st.append((float(r[i]), m.group(1)))
                except ValueError:
                    pass
        for v, n in sorted(st, reverse=True)[:8]:
            print("  %-40s %.2f" % (n, v))


if __name__ == "__main__":
    main()
