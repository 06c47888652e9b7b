#!/usr/bin/env python
"""Condense an .ncu-rep into the handful of counters DESIGN.md reasons about.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks/SM"),
    ("launch__occupancy_limit_warps", "occupancy limit (warps) blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "warp inst / cycle / SM (max 4)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slot utilisation %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst (max 32)"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", "warp execution efficiency %"),
    ("smsp__thread_inst_executed_pred_on_per_inst_executed.ratio", "pred-on threads / warp inst"),
    ("sm__sass_thread_inst_executed_op_fp32_pred_on.sum", "fp32 thread-inst (pred on)"),
    ("sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread-inst"),
    ("sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread-inst"),
    ("sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread-inst"),
    ("sm__sass_thread_inst_executed_op_integer_pred_on.sum", "integer thread-inst"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "FMA-heavy pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "CBU (branch) pipe %"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "ADU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors.sum", "L2 sectors"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sass__inst_executed_local_loads", "local-memory load inst"),
    ("sass__inst_executed_local_stores", "local-memory store inst"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__average_warp_latency_issue_stalled", "stall"),
    ("smsp__average_warps_issue_stalled", "stall"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for row in rows[2:]:
        if sub and sub not in row[name_col]:
            continue
        print(f"== {row[name_col]}  (id {row[0]})")
        d = {h: (v, u) for h, v, u in zip(hdr, row, units)}
        for key, label in WANT:
            if label == "stall":
                for h in sorted(d):
                    if h.startswith(key) and h.endswith("_ratio") or (h.startswith(key) and h.endswith(".ratio")):
                        v, u = d[h]
                        try:
                            if float(v) >= 0.05:
                                print(f"  {h:95s} {v} {u}")
                        except ValueError:
                            pass
                continue
            if key in d:
                v, u = d[key]
                print(f"  {label:45s} {v} {u}   [{key}]")
        print()


if __name__ == "__main__":
    main()
