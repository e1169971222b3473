#!/bin/bash
# One gpurun round trip: GPU tests, smoke, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -x ${PYTEST_ARGS:-} 2>&1 | tail -80 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -30 ) > gpurun_out/smoke.log
( timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -30 ) > gpurun_out/bench.log
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench.log
