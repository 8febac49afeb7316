"""Per-kernel table from an `ncu --set full` report (or its `--page raw --csv` export):
    python profiles/ncu_kernel_table.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel", 34), ("gpu__time_duration.sum", "ms", 8), ("dram__bytes_read.sum", "rd", 9), ("dram__bytes_write.sum", "wr", 9),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 6), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 6),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 6), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6),
    ("launch__registers_per_thread", "regs", 5), ("l1tex__t_sector_hit_rate.pct", "l1hit", 6), ("lts__t_sector_hit_rate.pct", "l2hit", 6),
    ("smsp__inst_executed.sum", "inst", 12), ("launch__grid_size", "grid", 8),
]


def main(path):
    if path.endswith(".csv"):  # already exported with `ncu -i report --page raw --csv`
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k) if k in hdr else -1, n, w) for k, n, w in WANT]
    print(" ".join(n.ljust(w) for _, n, w in idx))
    print(" ".join((units[i] if i >= 0 else "")[:w].ljust(w) for i, _, w in idx))
    for r in rows[2:]:
        cells = []
        for i, n, w in idx:
            v = r[i] if i >= 0 else "-"
            if n == "kernel":
                v = v.replace("ipcb::", "").replace("void ", "").split("(")[0]
            else:
                try:
                    v = ("%.4g" % float(v.replace(",", "")))
                except ValueError:
                    pass
            cells.append(v[:w].ljust(w))
        print(" ".join(cells))


if __name__ == "__main__":
    main(sys.argv[1])
