#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 | cut -c1-300) > gpurun_out/r2_pytest3.log 2>&1
(timeout 300 python tools/perf_llm.py 2>&1 | tail -7) > gpurun_out/r2_perf_llm4.log 2>&1
tail -n 8 gpurun_out/r2_pytest3.log; cat gpurun_out/r2_perf_llm4.log
