#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_bench_ref.log 2>&1
(time timeout 1700 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_bench_full.log 2>&1
tail -c 600 gpurun_out/r2_bench_ref.log; tail -c 5000 gpurun_out/r2_bench_full.log
