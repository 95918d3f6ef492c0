"""Key metrics of every kernel in an .ncu-rep (reads `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes.sum.per_second", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(r[ix["Kernel Name"]][:110], r[ix["Grid Size"]])
        for w in WANT:
            if w in ix:
                print("   %-82s %s %s" % (w, r[ix[w]], units[ix[w]]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
