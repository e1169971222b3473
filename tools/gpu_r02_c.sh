#!/bin/bash
# ncu full capture of the second-generation attention kernels (one launch each, bench shapes)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_(fwd2|bwd2)_kernel' -s 4 -c 4 \
    -f -o gpurun_out/prof_r02_attn python tools/prof_kernels.py --only attn > gpurun_out/ncu_attn.log 2>&1
tail -5 gpurun_out/ncu_attn.log; ls -la gpurun_out/prof_r02_attn.ncu-rep
