#!/usr/bin/env python
"""Summarise ncu reports brought back in gpurun_out/ into small text files under profiles/.
usage: python profiles/summarize_ncu.py <tag> kernel1 kernel2 ...   (reads gpurun_out/prof_<kernel>_<tag>.ncu-rep)"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    tag, kernels = sys.argv[1], sys.argv[2:]
    for k in kernels:
        rep = f"gpurun_out/prof_{k}_{tag}.ncu-rep"
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, vals = rows[0], rows[1], rows[-1]
        lines = [f"# ncu --set full --clock-control none, kernel {k}, tag {tag} (profiles/run_ncu.sh)",
                 f"# {vals[hdr.index('Kernel Name')]}"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"{w:80s} {vals[i]} {units[i]}")
        stalls = []
        for h, v in zip(hdr, vals):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(v), h))
                except ValueError:
                    pass
        lines.append("# warp stall reasons (cycles per issued instruction), top 6")
        for v, h in sorted(stalls, reverse=True)[:6]:
            lines.append(f"{h:80s} {v:.3f}")
        open(f"profiles/ncu_{k}_{tag}_summary.txt", "w").write("\n".join(lines) + "\n")
        print("wrote", f"profiles/ncu_{k}_{tag}_summary.txt")


if __name__ == "__main__":
    main()
