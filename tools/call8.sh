#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 | cut -c1-300) > gpurun_out/r2_pytest2.log 2>&1
(timeout 300 python tools/perf_unet.py 2>&1 | tail -3) > gpurun_out/r2_perf_unet5.log 2>&1
(timeout 600 python tools/determinism.py 2>&1 | tail -30) > gpurun_out/r2_determinism2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
tail -n 6 gpurun_out/r2_pytest2.log; cat gpurun_out/r2_perf_unet5.log; tail -n 8 gpurun_out/r2_determinism2.log; tail -n 2 gpurun_out/r2_smoke.log
