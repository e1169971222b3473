#!/bin/bash
# Round-2 GPU call K: phased GELU / GELU' epilogue arithmetic -- parity + timings.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -4 ) > gpurun_out/pytest_k.log
tail -2 gpurun_out/pytest_k.log
python tools/perf_epi.py 2>&1 | egrep "^M |plain bf16 \(B k|GELU"; python tools/perf_epi.py --dec 2>&1 | egrep "^M |plain bf16 \(B k|GELU"
python tools/prof_kernels.py --time --only gemm 2>&1 | tail -11
