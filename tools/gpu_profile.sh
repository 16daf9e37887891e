#!/bin/bash
# ncu evidence for profiles/ in one gpurun call (one GPU):  tools/gpu.sh 3000 gpurun_out/profile.log -- 'bash tools/gpu_profile.sh'
#   * launch lists: one eager UNet forward / one eager decode step inside an NVTX range, every launch with its duration
#   * --set full captures of the hot kernels (tools/ncu_targets.py), exported to CSV ON THE BOX: a .ncu-rep with dozens
#     of full captures exceeds the 64 MiB that gpurun copies back
# Afterwards, here:  python tools/launch_agg.py gpurun_out/<list>.csv   and   python tools/ncu_summary.py <raw.csv> <out.md> <title>
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "unet_forward/" --csv --log-file gpurun_out/launches_unet_forward.csv python tools/launch_list.py unet > gpurun_out/ll_unet.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "decode_step/" --csv --log-file gpurun_out/launches_decode_step.csv python tools/launch_list.py decode > gpurun_out/ll_decode.log 2>&1
for part in gemm2 skinny attn; do
  timeout 600 ncu --set full --clock-control none -k regex:"gemm_tc|fmha_tc|attn_decode|skinny_gemm|groupnorm" -c 14 -o /tmp/full_$part python tools/ncu_targets.py $part > gpurun_out/ncu_$part.log 2>&1
  ncu -i /tmp/full_$part.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$part.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
