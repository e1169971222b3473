#!/bin/bash
# Round-2 GPU call E: exponentials moved from the MUFU to the FMA pipe (CB_ATTN_POLY = 0 / 2 / 3 of every 8): parity + timings;
# HBM write-bandwidth probe.
mkdir -p gpurun_out
for P in 0 2 3; do
  echo "== CB_ATTN_POLY=$P"
  ( CB_ATTN_POLY=$P timeout 600 python -m pytest tests/test_attention_gpu.py -q -x 2>&1 | tail -3 )
  ( CB_ATTN_POLY=$P timeout 300 python tools/prof_kernels.py --time --only attn 2>&1 | tail -4 ) | tee gpurun_out/attn_times_poly$P.log
done
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from bench import measured_write_gbs
print("hbm write GB/s (1 GiB fill):", measured_write_gbs(torch.device("cuda")))
a = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
best = 0
for i in range(6):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); b.copy_(a); e.record(); torch.cuda.synchronize()
    best = max(best, 2 * a.numel() * 4 / (s.elapsed_time(e) * 1e-3) / 1e9)
print("hbm copy GB/s (read + write):", round(best, 1))
best = 0
for i in range(6):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); x = a.sum(); e.record(); torch.cuda.synchronize()
    best = max(best, a.numel() * 4 / (s.elapsed_time(e) * 1e-3) / 1e9)
print("hbm read GB/s (1 GiB sum):", round(best, 1))
PY
