#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the per-kernel numbers DESIGN.md and bench.py quote.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.md
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe inst %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe inst %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread inst"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread inst"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread inst"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary of `{rep.split('/')[-1]}`", "",
             "Per launch (cold cache, serialised by ncu: shares matter, not absolutes).", ""]
    for r in data:
        lines.append(f"## {r[col['Kernel Name']]}  (launch id {r[col['ID']]})")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k, label in KEYS:
            if k in col and r[col[k]] != "":
                lines.append(f"| {label} (`{k}`) | {r[col[k]]} | {units[col[k]]} |")
        stalls = []
        for h, i in col.items():
            if "average_warp_latency_issue_stalled" in h or ("warp_issue_stalled" in h and h.endswith("per_warp_active.pct")):
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        for v, h in stalls[:6]:
            lines.append(f"| stall `{h}` | {v:.2f} | {units[col[h]]} |")
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(data), "launches")


if __name__ == "__main__":
    main()
