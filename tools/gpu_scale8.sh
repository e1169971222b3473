#!/bin/bash
# 8-GPU scaling run (gpurun --gpus 8): ViT-B (bench default, LAX 192) and ViT-L (BASELINE.json config 5, LAX 256) at N = 8 and,
# on the same box, N = 1; NCCL's algorithm / protocol choice for the gradient all-reduce recorded once (NCCL_DEBUG=INFO).
mkdir -p gpurun_out
run() {  # $1 = nproc, $2 = output tag, rest = bench arguments
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then
    python bench.py --steps 20 --warmup 5 --no-stock-gpu --no-cpu-baseline --no-kernel-profile "$@" 2>&1 | grep '^{' | tail -1 > gpurun_out/$tag.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 20 --warmup 5 --no-stock-gpu --no-cpu-baseline --no-kernel-profile "$@" > gpurun_out/$tag.log 2>&1
    grep '^{' gpurun_out/$tag.log | tail -1 > gpurun_out/$tag.json
  fi
}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING run 8 scale8_base
grep -E "NVLS|Using network|Channel|comm 0x.*nranks|AllReduce.*algo|Algo|protocol|nChannels|Connected|P2P/CUMEM|Tree|Ring" gpurun_out/scale8_base.log | cut -c1-220 | sort | uniq -c | sort -rn | head -60 > gpurun_out/nccl_info_n8.txt
run 8 scale8_large --size large --lax 256
CB_OVERLAP_ALLREDUCE=0 run 8 scale8_base_nooverlap
python - <<'PY'
import json
for f in ("scale8_base", "scale8_base_nooverlap", "scale8_large"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["workload"][:60])
    except Exception as e:
        print(f, "failed", e)
PY
head -30 gpurun_out/nccl_info_n8.txt
rm -f gpurun_out/scale8_base.log.tmp
