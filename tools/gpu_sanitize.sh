#!/bin/bash
# compute-sanitizer gate: memcheck + racecheck + synccheck + initcheck over tools/sanitize_target.py.
# Logs -> gpurun_out/sanitize_<tool>.log, one-line summaries -> gpurun_out/sanitize_summary.txt (copied to profiles/).
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1   # every tensor is its own cudaMalloc: out-of-bounds accesses hit unmapped memory
export CB_PDL=${CB_PDL:-1}
: > gpurun_out/sanitize_summary.txt
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report ${RACE_REPORT:-analysis}"
  ( timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 30 --launch-timeout 120 \
      python tools/sanitize_target.py ${TARGETS:-gemm,attn,ln,model} 2>&1 | tail -400 ) > gpurun_out/sanitize_$tool.log
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1) | $(grep -c 'sanitize target done' gpurun_out/sanitize_$tool.log) completed" >> gpurun_out/sanitize_summary.txt
done
cat gpurun_out/sanitize_summary.txt
