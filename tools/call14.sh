#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/decode_kernels.py 1 1041 2>&1 | head -16) > gpurun_out/r2_decode_kernels5.log 2>&1
(timeout 300 python tools/perf_llm.py 2>&1 | tail -6) > gpurun_out/r2_perf_llm5.log 2>&1
(timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q --timeout 600 2>&1 | tail -4) > gpurun_out/r2_t5.log 2>&1
cat gpurun_out/r2_decode_kernels5.log gpurun_out/r2_perf_llm5.log; tail -n 3 gpurun_out/r2_t5.log
