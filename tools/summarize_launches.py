"""Summarise an ``ncu --metrics gpu__time_duration.sum --csv`` launch list into a per-kernel table of ONE step.

    python tools/summarize_launches.py gpurun_out/launches.csv [--out profiles/r01_launches.md] [--title "..."]

The step is delimited by the optimiser kernel (``adamw_kernel``): everything after the previous step's last
``adamw_kernel`` launch up to and including this step's last one.  Times are ncu's serialised cold-cache
durations: compare SHARES, not absolutes (B200_PROFILING.md).
"""

from __future__ import annotations

import argparse
import csv
import io
import re
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(<[^()]*>)?)", name)
    s = m.group(1) if m else name
    s = s.replace("at::native::", "at::")
    return s[:110]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--out", default=None)
    ap.add_argument("--title", default="kernel launches of one step")
    ap.add_argument("--marker", default="adamw_kernel")
    ap.add_argument("--traffic-json", default=None, help="write the GEMM kernel's mean DRAM bytes per launch here")
    ap.add_argument("--step", type=int, default=-1, help="which step (index into marker-delimited steps)")
    a = ap.parse_args()
    text = open(a.csv, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    unit_ns = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
    unit_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    by_id: "OrderedDict[str, list]" = OrderedDict()
    for r in rows:
        e = by_id.setdefault(r["ID"], [r["Kernel Name"], 0.0, 0.0])
        v = float(r["Metric Value"].replace(",", "") or 0)
        if r.get("Metric Name") == "gpu__time_duration.sum":
            e[1] = v * unit_ns.get(r.get("Metric Unit", "ns"), 1.0)
        elif r.get("Metric Name", "").startswith("dram__bytes_"):
            e[2] += v * unit_b.get(r.get("Metric Unit", "byte"), 1.0)
    launches = [(n, ns) for n, ns, _ in by_id.values()]
    dram = [b for _, _, b in by_id.values()]
    marks = [i for i, (n, _) in enumerate(launches) if a.marker in n]
    # group consecutive marker launches (segments of the arena) into step ends
    ends = [m for j, m in enumerate(marks) if j + 1 == len(marks) or marks[j + 1] - m > 8]
    if len(ends) >= 2:
        idx = a.step if a.step >= 0 else len(ends) - 1
        lo, hi = ends[idx - 1] + 1, ends[idx] + 1
    else:
        lo, hi = 0, len(launches)
    step = launches[lo:hi]
    step_dram = dram[lo:hi]
    agg: "OrderedDict[str, list[float]]" = OrderedDict()
    for (n, ns), b in zip(step, step_dram):
        e = agg.setdefault(short(n), [0.0, 0, 0.0])
        e[0] += ns
        e[1] += 1
        e[2] += b
    total = sum(v[0] for v in agg.values())
    have_dram = any(v[2] > 0 for v in agg.values())
    lines = [f"# {a.title}", "",
             f"{len(step)} launches, {total / 1e6:.3f} ms summed ncu durations (serialised, cold cache: compare shares). "
             f"{len(ends)} steps found in the capture; launches {lo}..{hi} shown.", "",
             "| kernel | launches | total ms | share | avg us |" + (" DRAM MB / launch | DRAM GB/s |" if have_dram else ""),
             "|---|---:|---:|---:|---:|" + ("---:|---:|" if have_dram else "")]
    for n, (ns, c, b) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        row = f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / c / 1e3:.1f} |"
        if have_dram:
            row += f" {b / c / 1e6:.2f} | {b / ns:.0f} |"
        lines.append(row)
    out = "\n".join(lines) + "\n"
    if a.out:
        open(a.out, "w").write(out)
    if a.traffic_json and have_dram:
        import json

        g = [(ns, b) for (n, ns), b in zip(step, step_dram) if "gemm_bf16_kernel" in n]
        json.dump({"kernel": "gemm_bf16_kernel", "launches": len(g), "dram_bytes_per_launch": round(sum(b for _, b in g) / len(g)),
                   "avg_us": round(sum(ns for ns, _ in g) / len(g) / 1e3, 2),
                   "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one eager step (tools/gpu_launchlist.sh)"},
                  open(a.traffic_json, "w"))
    try:
        print(out)
    except BrokenPipeError:
        pass


if __name__ == "__main__":
    main()
