"""Markdown table of an `ncu --set full` report: one row per profiled launch (duration, DRAM bytes, DRAM / L2 / tensor /
issue / XU utilisation, achieved warps, registers, grid).   python tools/ncu_full_summary.py report.ncu-rep > table.md"""
import csv
import subprocess
import sys

COLS = [("us", "gpu__time_duration.sum", 1e-3), ("MB rd", "dram__bytes_read.sum", 1e-6), ("MB wr", "dram__bytes_write.sum", 1e-6),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("tensor %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("XU %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
        ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("regs", "launch__registers_per_thread", 1),
        ("grid", "launch__grid_size", 1)]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for n, r in enumerate(body):
        name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        vals = []
        for label, metric, scale in COLS:
            if metric not in idx or r[idx[metric]] in ("", "n/a"):
                vals.append("-")
                continue
            v = float(r[idx[metric]].replace(",", ""))
            u = units[idx[metric]]
            if label == "us":
                v = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1e-3)
            elif label.startswith("MB"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            vals.append(f"{v:.1f}")
        print(f"| {n} | `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
