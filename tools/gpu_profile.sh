#!/bin/bash
# One gpurun round trip: bench (graph mode), stock-torch GPU baseline, ncu launch list of eager steps.
mkdir -p gpurun_out
( timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -5 ) > gpurun_out/bench.log
( timeout 600 python tests/perf_stock_torch_gpu.py 2>&1 | tail -5 ) > gpurun_out/stock_torch.log
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-profile --no-stock-gpu 2>&1 | tail -5 ) > gpurun_out/ncu_launches.log
tail -c 600 gpurun_out/stock_torch.log; tail -c 300 gpurun_out/ncu_launches.log; ls -la gpurun_out
