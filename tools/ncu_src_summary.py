"""Summarise an `ncu --page source --csv` export: samples per stall reason over the whole kernel, the hottest SASS
instructions, and samples grouped by opcode.   python tools/ncu_src_summary.py file.csv [top_n]"""
import csv
import sys
from collections import Counter


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    # the export holds one section per kernel: ["Kernel Name", name], header row, SASS rows
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(starts[:-1], starts[1:]):
        print("=" * 20, rows[a][1][:90])
        section(rows[a + 1], rows[a + 2:b], path, top)


def section(hdr, body, path, top):
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in body if len(r) == len(hdr)]
    tot = Counter()
    by_op = Counter()
    inst = Counter()
    n_all = 0
    for r in data:
        n = int(r[idx["# Samples"]] or 0)
        n_all += n
        op = r[idx["Source"]].split()
        op = (op[1] if op and op[0].startswith("@") else op[0]) if op else "?"
        by_op[op] += n
        inst[op] += int(r[idx["Instructions Executed"]] or 0)
        for c in stall_cols:
            tot[c] += int(r[idx[c]] or 0)
    print(f"{path}: {len(data)} SASS lines, {n_all} samples, {sum(inst.values())} warp instructions")
    print("stall reasons:", ", ".join(f"{k[6:]} {v * 100 / max(n_all, 1):.1f}%" for k, v in tot.most_common(9)))
    print("samples by opcode:", ", ".join(f"{k} {v * 100 / max(n_all, 1):.1f}%" for k, v in by_op.most_common(14)))
    print("warp instructions by opcode:", ", ".join(f"{k} {v}" for k, v in inst.most_common(16)))
    hot = sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:top]
    for r in hot:
        n = int(r[idx["# Samples"]] or 0)
        reasons = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"  {n * 100 / max(n_all, 1):5.2f}%  {r[idx['Source']].strip()[:70]:70s} x{r[idx['Instructions Executed']]:>8s}  " +
              " ".join(f"{b}:{a}" for a, b in reasons if a))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
