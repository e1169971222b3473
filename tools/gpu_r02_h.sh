#!/bin/bash
# Round-2 GPU call H: wide-store GEMM epilogue -- parity, A/B timings against the narrow path on the same box.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -6 ) > gpurun_out/pytest_h.log
tail -3 gpurun_out/pytest_h.log
echo "== wide (default)"; python tools/perf_epi.py 2>&1 | egrep "^M |plain bf16 \(B k|\+bias  |GELU \(|fp32 out \+"; python tools/perf_epi.py --dec 2>&1 | egrep "^M |plain bf16 \(B k|\+bias  |GELU \(|fp32 out \+"
echo "== narrow (CB_GEMM_WIDE=0)"; CB_GEMM_WIDE=0 python tools/perf_epi.py 2>&1 | egrep "^M |plain bf16 \(B k|\+bias  |GELU \(|fp32 out \+"; CB_GEMM_WIDE=0 python tools/perf_epi.py --dec 2>&1 | egrep "^M |plain bf16 \(B k|\+bias  |GELU \(|fp32 out \+"
echo "== kernel times, wide"; python tools/prof_kernels.py --time --only gemm 2>&1 | tail -11
