#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): the metrics the roofline argument uses + top stall reasons.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
    for r in data:
        name = r[idx["Kernel Name"]]
        if pat not in name:
            continue
        print("=" * 110)
        print("Kernel Name".ljust(78), name)
        for w in WANT:
            if w in idx:
                print(w.ljust(78), units[idx[w]].ljust(16), r[idx[w]])
        st = sorted(((float(r[idx[h]]), h) for h in stall), reverse=True)[:8]
        for v, h in st:
            print("   stall/issue", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "").ljust(40), "%.3f" % v)


if __name__ == "__main__":
    main()
