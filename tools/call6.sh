#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 600 -k "folded or gemm" 2>&1 | tail -15) > gpurun_out/r2_t_kernels4.log 2>&1
for f in 1 0; do
  (SS_UNET_LNFOLD=$f timeout 300 python tools/perf_unet.py 2>&1 | tail -3) > gpurun_out/r2_perf_unet4_fold$f.log 2>&1
done
(timeout 1500 python -m pytest tests/test_acceptance_gpu.py tests/test_fullsize_gpu.py tests/test_sdxl_gpu.py -q --timeout 900 2>&1 | tail -60) > gpurun_out/r2_t_rest4.log 2>&1
# profiles: launch lists (eager single pass inside an NVTX range) and full captures of the hot kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "unet_forward/" --csv --log-file gpurun_out/r2_launches_unet_forward.csv python tools/launch_list.py unet > gpurun_out/r2_ll_unet.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "decode_step/" --csv --log-file gpurun_out/r2_launches_decode_step.csv python tools/launch_list.py decode > gpurun_out/r2_ll_decode.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|fmha_tc|attn_decode|skinny_gemm|groupnorm|layernorm" -o gpurun_out/r2_full python tools/ncu_targets.py all > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/r2_perf_unet4_fold1.log gpurun_out/r2_perf_unet4_fold0.log; tail -n 3 gpurun_out/r2_t_kernels4.log; tail -n 5 gpurun_out/r2_t_rest4.log
