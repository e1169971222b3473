#!/bin/bash
# Round-2 artefact run (1 GPU): tests, smoke, bench (ViT-B with stock + CPU arms; ViT-L LAX 256), kernel timings,
# ncu launch list of one eager step, compute-sanitizer.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -2 ) > gpurun_out/bench_base.log
( timeout 600 python bench.py --steps 10 --warmup 3 --size large --lax 256 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/bench_large256.log
( timeout 600 python bench.py --steps 10 --warmup 3 --lax 256 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/bench_base256.log
( timeout 200 python tools/prof_kernels.py --time 2>&1 | tail -40 ) > gpurun_out/kernel_times.log
( timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-profile --skip-e2e --no-stock-gpu 2>&1 | tail -3 ) > gpurun_out/ncu_launches.log
SAN_TIMEOUT=300 bash tools/gpu_sanitize.sh > /dev/null 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; cat gpurun_out/kernel_times.log | head -12; cat gpurun_out/sanitize_summary.txt
python - <<'PY'
import json
for f in ("bench_base", "bench_large256", "bench_base256"):
    try:
        line = [l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1]
        d = json.loads(line)
        sg = d.get("stock_gpu_baseline", {})
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline", {}).get("frac"), d.get("attention_roofline", {}).get("frac"),
              d.get("step_roofline", {}).get("frac"), {k: sg.get(k) for k in ("value", "ms_per_step", "grad_ckpt", "speedup_e2e", "speedup_vs_reference_default_grad_ckpt")},
              d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
    except Exception as e:
        print(f, "no line", e)
PY
