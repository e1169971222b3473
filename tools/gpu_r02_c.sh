#!/bin/bash
# Round-2 GPU call C: tests, synccheck of the attention kernels, quick bench (replayed kernel table), GEMM timings and an
# ncu --set full capture of the GEMM shapes (source-level, for the epilogue analysis).
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
TOOLS="synccheck" TARGETS="attn,gemm" SAN_TIMEOUT=200 bash tools/gpu_sanitize.sh
( timeout 600 python bench.py --steps 10 --warmup 3 --no-stock-gpu --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/bench_c2.log
( timeout 200 python tools/prof_kernels.py --time --only gemm,attn 2>&1 | tail -40 ) > gpurun_out/kernel_times_c.log
cat gpurun_out/kernel_times_c.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 11 -o gpurun_out/prof_r02_gemm -f python tools/prof_kernels.py --only gemm 2>&1 | tail -3 ) > gpurun_out/ncu_gemm.log
python - <<'PY'
import json
line = [l for l in open("gpurun_out/bench_c2.log") if l.startswith("{")][-1]
d = json.loads(line)
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline", {}).get("frac"), d.get("attention_roofline"))
for k, v in list(d["kernel_profile"]["kernels"].items())[:16]:
    print(k, v)
print({k: v for k, v in d["kernel_profile"].items() if k.endswith("_ms")})
PY
