#!/bin/bash
# Round-2 GPU call A: full GPU tests (incl. the experimental conv GEMM), smoke, bench (ViT-B + stock arm), bench ViT-L,
# kernel timings, compute-sanitizer.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( CB_EXPERIMENTAL_CONV=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=40 2>&1 | tail -120 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench_base.log
( timeout 900 python bench.py --steps 10 --warmup 3 --size large --lax 256 --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/bench_large256.log
( timeout 300 python tools/prof_kernels.py --time 2>&1 | tail -40 ) > gpurun_out/kernel_times.log
SAN_TIMEOUT=420 bash tools/gpu_sanitize.sh > /dev/null 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/kernel_times.log; cat gpurun_out/sanitize_summary.txt
python - <<'PY'
import json
for f in ("bench_base", "bench_large256"):
    try:
        line = [l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline", {}).get("frac"), d.get("attention_roofline", {}).get("frac"),
              d.get("stock_gpu_baseline"))
    except Exception as e:
        print(f, "no line", e)
PY
