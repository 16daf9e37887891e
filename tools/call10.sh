#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
  (SS_DECODE_L2_PREFETCH=$v timeout 300 python tools/perf_llm.py 2>&1 | tail -6) > gpurun_out/r2_perf_llm_pf$v.log 2>&1
done
(timeout 600 python -m pytest tests/test_models_gpu.py tests/test_kernels_gpu.py -q --timeout 600 2>&1 | tail -5) > gpurun_out/r2_t_pf.log 2>&1
for v in 1 0; do echo pf=$v; cat gpurun_out/r2_perf_llm_pf$v.log; done; tail -n 3 gpurun_out/r2_t_pf.log
