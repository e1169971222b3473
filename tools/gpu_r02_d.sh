#!/bin/bash
# Round-2 GPU call D: GEMM epilogue changes (bias one chunk ahead, predicate-free full tiles) + zoom kernel.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or zoom or scale_intensity" 2>&1 | tail -15 ) > gpurun_out/pytest_d.log
tail -4 gpurun_out/pytest_d.log
( timeout 200 python tools/prof_kernels.py --time --only gemm 2>&1 | tail -40 ) > gpurun_out/kernel_times_d.log
cat gpurun_out/kernel_times_d.log
for i in 1 2 3; do ( timeout 300 python -m pytest tests/test_model_gpu.py -q -x -k "prefetch_equals" 2>&1 | grep -E "passed|failed|direct" | cut -c1-600 ); done
( CB_EXPERIMENTAL_CONV=1 timeout 600 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-stock-gpu --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/bench_d2.log
python - <<'PY'
import json
line = [l for l in open("gpurun_out/bench_d2.log") if l.startswith("{")][-1]
d = json.loads(line)
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline", {}).get("frac"), d.get("attention_roofline", {}).get("frac"))
for k, v in list(d["kernel_profile"]["kernels"].items())[:8]:
    print(k, v)
for s in d["kernel_profile"]["gemm_top_shapes"][:24]:
    print(s["gemm"], s["launches"], s["ms"], s["tflops"], s["bound"], s["frac_of_bound"])
PY
