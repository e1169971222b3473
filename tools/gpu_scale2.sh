#!/bin/bash
# 2-GPU scaling check: N=1 reference, N=2 with the CTA-limited background communicator, N=2 without it.
mkdir -p gpurun_out
run() {  # $1 = nproc, rest = env assignments
  local n=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 30 --warmup 5 --no-stock-gpu --no-cpu-baseline --no-kernel-profile 2>&1 | grep '^{' | tail -1
}
python bench.py --steps 30 --warmup 5 --no-stock-gpu --no-cpu-baseline --no-kernel-profile 2>&1 | grep '^{' | tail -1 > gpurun_out/scale_n1.json
run 2 CB_OVERLAP_ALLREDUCE=0 > gpurun_out/scale_n2_nooverlap.json
run 2 CB_STAGE_POINTS=6 > gpurun_out/scale_n2_pts6.json
run 2 CB_STAGE_POINTS=9,6,3 > gpurun_out/scale_n2_pts963.json
run 2 CB_STAGE_POINTS=9,6,3,1 > gpurun_out/scale_n2_pts9631.json
run 2 CB_STAGE_POINTS=8,4,1 > gpurun_out/scale_n2_pts841.json
python - <<'PY'
import json
for f in ("scale_n1", "scale_n2_nooverlap", "scale_n2_pts6", "scale_n2_pts963", "scale_n2_pts9631", "scale_n2_pts841"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
