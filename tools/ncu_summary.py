"""Condense `ncu --set full` reports into the few lines DESIGN.md and bench.py quote.

usage: python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep > profiles/ncu_summary_rNN.md
Reads each report with `ncu -i … --page raw --csv` (no GPU needed).
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
     "tensor (DMMA) pipe active, % of elapsed"),
    ("TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
     "FP64 (DFMA) pipe active, % of elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(head, units, r)}
        print(f"### `{d['Kernel Name'][0][:110]}`\n")
        print(f"report `{path.split('/')[-1]}`, grid {d['Grid Size'][0]}, block {d['Block Size'][0]}\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in d:
                v, u = d[k]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                print(f"| {label} | {v} {u} |")
        print()


if __name__ == "__main__":
    for p in sys.argv[1:]:
        report(p)
