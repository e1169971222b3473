#!/bin/bash
# Round-2 GPU call J: wide-store epilogue for plain bf16 outputs -- full GPU tests, step A/B against the narrow path (same box).
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -6 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for W in 1 0 1 0; do
  echo "== CB_GEMM_WIDE=$W"
  CB_GEMM_WIDE=$W python bench.py --steps 20 --warmup 5 --no-stock-gpu --no-cpu-baseline --no-kernel-profile 2>&1 | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
python tools/prof_kernels.py --time --only gemm 2>&1 | tail -11
