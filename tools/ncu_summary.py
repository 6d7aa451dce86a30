#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md and
profiles/ quote.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "inst_executed",
    "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_integer_pred_on.sum", "smsp__sass_thread_inst_executed_op_conversion_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"== kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        u = dict(zip(hdr, units))
        for k in WANT:
            if k in d and d[k] not in ("", "n/a"):
                out.append(f"{k:75s} {d[k]} {u[k]}")
        stalls = [(k, float(d[k])) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                  and d.get(k) not in ("", None, "n/a")]
        if not stalls:
            stalls = [(k, float(d[k])) for k in hdr if k.startswith("smsp__average_warp_latency_issue_stalled_") and d.get(k) not in ("", None, "n/a")]
        stalls.sort(key=lambda kv: -kv[1])
        out.append("-- warp stall reasons (warps stalled per issue-active cycle), top 8")
        for k, v in stalls[:8]:
            out.append(f"{k:75s} {v:.3f}")
        pipes = [(k, float(d[k])) for k in hdr if k.startswith("sm__inst_executed_pipe_") and k.endswith("pct_of_peak_sustained_active")
                 and d.get(k) not in ("", None, "n/a")]
        pipes.sort(key=lambda kv: -kv[1])
        out.append("-- pipe utilisation (% of peak sustained active), top 8")
        for k, v in pipes[:8]:
            out.append(f"{k:75s} {v:.2f}")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
