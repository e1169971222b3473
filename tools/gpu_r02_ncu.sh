#!/bin/bash
# Round-2 end: ncu --set full of the hot kernels at the bench shapes (source-level), input-pipeline kernel timings.
mkdir -p gpurun_out
( timeout 200 python tools/prof_kernels.py --time --only misc 2>&1 | tail -8 ) | tee gpurun_out/kernel_times_misc.log
( timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_bf16|attn_fwd2|attn_bwd2|attn_delta|attn_dq_convert|ln_bwd_tma|ln_fwd_kernel|dwconv_fast|dwconv_wgrad_fast|colsum_bf16|zoom_resample|zoom_scale" \
    -o gpurun_out/prof_r02_final -f python tools/prof_kernels.py 2>&1 | tail -3 ) > gpurun_out/ncu_final.log
tail -2 gpurun_out/ncu_final.log; ls -la gpurun_out/prof_r02_final.ncu-rep
