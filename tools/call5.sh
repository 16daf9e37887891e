#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 600 2>&1 | tail -15) > gpurun_out/r2_t_kernels3.log 2>&1
(timeout 300 python tools/decode_kernels.py 1 1041 2>&1 | tail -30) > gpurun_out/r2_decode_kernels3.log 2>&1
(timeout 300 python tools/perf_llm.py 2>&1 | tail -8) > gpurun_out/r2_perf_llm3.log 2>&1
for f in 1 0; do
  (SS_UNET_LNFOLD=$f timeout 300 python tools/perf_unet.py 2>&1 | tail -3) > gpurun_out/r2_perf_unet3_fold$f.log 2>&1
done
(timeout 900 python -m pytest tests/test_models_gpu.py tests/test_acceptance_gpu.py -q --timeout 900 2>&1 | tail -60) > gpurun_out/r2_t_models3.log 2>&1
cat gpurun_out/r2_decode_kernels3.log; tail -6 gpurun_out/r2_perf_llm3.log; tail -2 gpurun_out/r2_perf_unet3_fold*.log; tail -3 gpurun_out/r2_t_kernels3.log gpurun_out/r2_t_models3.log
