#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 | cut -c1-300) > gpurun_out/r2_pytest4.log 2>&1
(time timeout 1700 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_bench_full2.log 2>&1
(time timeout 600 python bench.py --config sdxl --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline) > gpurun_out/r2_bench_sdxl.log 2>&1
(time timeout 1200 python bench.py --config sink --stories-per-gpu 4 --turns 25 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-eager-baseline) > gpurun_out/r2_bench_sink25_spg4.log 2>&1
tail -n 5 gpurun_out/r2_pytest4.log; for f in r2_bench_full2 r2_bench_sdxl r2_bench_sink25_spg4; do grep -E "^\{|^real|Error" gpurun_out/$f.log | cut -c1-300; done
