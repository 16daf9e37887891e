#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_models_gpu.py tests/test_sdxl_gpu.py -q --timeout 600 2>&1 | tail -40 | cut -c1-300) > gpurun_out/r2_t_batch.log 2>&1
for spg in 4 1; do
  (time timeout 900 python bench.py --config sink --turns 6 --stories-per-gpu $spg --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-eager-baseline) > gpurun_out/r2_bench_sink_spg$spg.log 2>&1
done
tail -n 12 gpurun_out/r2_t_batch.log; for spg in 4 1; do grep -E "^\{|Error|error" gpurun_out/r2_bench_sink_spg$spg.log | cut -c1-400; done
