"""Build libcinema_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m cinema_b200.build [--force] [--verbose]

Sources: cinema_b200/csrc/*.cu -> cinema_b200/lib/libcinema_b200.so.  The library links the
static CUDA runtime, so it loads next to torch's own runtime without any extra search path.
nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo.
"""

from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libcinema_b200.so"
OBJDIR = LIBDIR / "obj"
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; set NVCC or put it on PATH")
    return cand


if os.environ.get("CB_GEMM_TRACE") == "1":  # diagnostics build: in-kernel clock stamps in the GEMM (tools/gemm_trace.py)
    NVCC_FLAGS.append("-DCB_GEMM_TRACE")


def _digest(paths: list[Path]) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = find_nvcc()
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    stamp = LIBDIR / "build.sha256"
    digest = _digest(sources + headers)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    OBJDIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path) -> tuple[Path, str]:
        obj = OBJDIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    log = "\n".join(f"== {o.name}\n{err}" for o, err in results)
    (LIBDIR / "ptxas.log").write_text(log)
    if verbose:
        print(log)
    objs = [str(o) for o, _ in results]
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    sys.exit(0)
