#!/bin/bash
# Round-2 GPU call B: second-generation attention kernels -- tests, A/B timings, bench.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_attention_gpu.py -q -x 2>&1 | tail -30 ) > gpurun_out/pytest_attn.log
tail -3 gpurun_out/pytest_attn.log
if grep -q "passed" gpurun_out/pytest_attn.log && ! grep -q "failed" gpurun_out/pytest_attn.log; then
  ( timeout 300 python tools/prof_kernels.py --time --only attn 2>&1 | tail -8 ) > gpurun_out/attn_times_gen2.log
  ( CB_ATTN_FWD=1 CB_ATTN_BWD=1 timeout 300 python tools/prof_kernels.py --time --only attn 2>&1 | tail -8 ) > gpurun_out/attn_times_gen1.log
  echo gen2; cat gpurun_out/attn_times_gen2.log; echo gen1; cat gpurun_out/attn_times_gen1.log
  ( timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
  tail -4 gpurun_out/pytest_gpu.log
  ( timeout 900 python bench.py --steps 10 --warmup 3 --no-stock-gpu --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/bench_b.log
  python - <<'PY'
import json
line = [l for l in open("gpurun_out/bench_b.log") if l.startswith("{")][-1]
d = json.loads(line)
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline", {}).get("frac"), d.get("attention_roofline"))
for k, v in list(d["kernel_profile"]["kernels"].items())[:12]:
    print(k, v)
PY
fi
TOOLS="racecheck initcheck" TARGETS="gemm,attn,ln" SAN_TIMEOUT=420 bash tools/gpu_sanitize.sh
