#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_acceptance_gpu.py::test_gen_george_runs_unchanged -q --timeout 900 2>&1 | grep -v "^E    *<img" | tail -80 | cut -c1-400) > gpurun_out/r2_t_accept2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "unet_forward/" --csv --log-file gpurun_out/r2_launches_unet_forward.csv python tools/launch_list.py unet > gpurun_out/r2_ll_unet.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "decode_step/" --csv --log-file gpurun_out/r2_launches_decode_step.csv python tools/launch_list.py decode > gpurun_out/r2_ll_decode.log 2>&1
for part in gemm2 skinny attn; do
  timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc|fmha_tc|attn_decode|skinny_gemm|groupnorm" -c 14 -o /tmp/r2_full_$part python tools/ncu_targets.py $part > gpurun_out/r2_ncu_$part.log 2>&1
  ncu -i /tmp/r2_full_$part.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_$part.csv 2>/dev/null
done
ls -la gpurun_out | tail -12; tail -n 30 gpurun_out/r2_t_accept2.log
