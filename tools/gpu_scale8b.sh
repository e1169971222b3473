#!/bin/bash
# 8-GPU follow-up: does a CTA-limited NCCL (fewer NVLS channels beside the compute kernels) make the overlapped
# gradient all-reduce pay at N = 8?  ViT-B, 20 steps each.
mkdir -p gpurun_out
run() {  # $1 = output tag, rest = env assignments
  local tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-stock-gpu --no-cpu-baseline --no-kernel-profile 2>&1 | grep '^{' | tail -1 > gpurun_out/$tag.json
}
run scale8_ctas8 NCCL_MAX_CTAS=8
run scale8_ctas16 NCCL_MAX_CTAS=16
run scale8_ctas4 NCCL_MAX_CTAS=4
run scale8_nooverlap_b CB_OVERLAP_ALLREDUCE=0
python - <<'PY'
import json
for f in ("scale8_ctas4", "scale8_ctas8", "scale8_ctas16", "scale8_nooverlap_b"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
