"""bench.py contract checks that run without a GPU: the reference arm (the reference's CPU path = the oracle port, timed on
the host cores) prints ONE JSON line with the agreed keys, and our arm refuses to run without a CUDA device."""

import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    res = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "mae_pretrain_volumes_per_sec"
    assert line["unit"] == "frame-set volumes/s" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["steps"] == 1 and line["warmup"] == 0 and line["value"] > 0 and line["ms_per_step"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    cpu = line["cpu_baseline"]
    # "reference": the unmodified reference modules installed under baseline/_ref; "port": the oracle when that install is absent
    want = "reference" if (ROOT / "baseline" / "_ref" / "cinema" / "mae" / "mae.py").exists() else "port"
    assert cpu["kind"] == want and cpu["cores"] >= 1 and cpu["value"] == line["value"] and cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None  # BASELINE.md holds no published number for this metric


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU refusal")
def test_our_arm_fails_loudly_without_a_gpu():
    res = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", timeout=300)
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr
