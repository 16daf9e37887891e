#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/decode_kernels.py 1 1041 2>&1 | tail -30) > gpurun_out/r2_decode_kernels.log 2>&1
(timeout 300 python tools/decode_kernels.py 4 1041 2>&1 | tail -30) > gpurun_out/r2_decode_kernels_b4.log 2>&1
(timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 600 -x -k "decode or fused or skinny or rope" 2>&1 | tail -15) > gpurun_out/r2_t_kernels2.log 2>&1
(timeout 600 python tools/perf_llm.py 2>&1 | tail -9) > gpurun_out/r2_perf_llm2.log 2>&1
cat gpurun_out/r2_decode_kernels.log; tail -3 gpurun_out/r2_t_kernels2.log; cat gpurun_out/r2_perf_llm2.log
