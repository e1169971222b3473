#!/bin/bash
# ncu launch list (duration + DRAM bytes per launch) of eager bench steps -> gpurun_out/launches.csv
mkdir -p gpurun_out
( timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-profile --skip-e2e --no-stock-gpu 2>&1 | tail -5 ) > gpurun_out/ncu_launches.log
tail -c 300 gpurun_out/ncu_launches.log; ls -la gpurun_out
