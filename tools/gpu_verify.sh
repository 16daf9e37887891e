#!/bin/bash
# What the driver runs at round end, in one gpurun call:  tools/gpu.sh 5000 gpurun_out/verify.log -- 'bash tools/gpu_verify.sh'
# (GPU test suite, smoke, both bench arms at the driver's --steps 20 --warmup 5; outputs under gpurun_out/).
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 | cut -c1-300) > gpurun_out/verify_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1
(time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/verify_bench_ref.log 2>&1
(time timeout 1700 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/verify_bench.log 2>&1
tail -n 5 gpurun_out/verify_pytest.log; tail -n 1 gpurun_out/verify_smoke.log
for f in verify_bench_ref verify_bench; do grep -E "^\{|^real|Error" gpurun_out/$f.log | cut -c1-400; done
