#!/bin/bash
# development aid: verbose re-run of the tests that failed, the new fused-kernel tests, stage timings, bench tail
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 600 -x 2>&1 | tail -60) > gpurun_out/r2_t_kernels.log 2>&1
(timeout 600 python -m pytest "tests/test_fullsize_gpu.py::test_unet_full_width_blocks_match_oracle" "tests/test_fullsize_gpu.py::test_dropins_generate_and_get_image_embeds_match_oracle" -q --timeout 600 2>&1 | tail -150) > gpurun_out/r2_t_full.log 2>&1
(SS_UNET_LNFOLD=0 timeout 600 python -m pytest "tests/test_fullsize_gpu.py::test_unet_full_width_blocks_match_oracle" -q --timeout 600 2>&1 | tail -60) > gpurun_out/r2_t_full_nofold.log 2>&1
(timeout 900 python -m pytest tests/test_acceptance_gpu.py::test_gen_george_runs_unchanged -q --timeout 900 2>&1 | tail -150) > gpurun_out/r2_t_accept.log 2>&1
(timeout 900 python -m pytest tests/test_models_gpu.py tests/test_sdxl_gpu.py -q --timeout 600 2>&1 | tail -60) > gpurun_out/r2_t_models.log 2>&1
(timeout 600 python tools/perf_llm.py 2>&1 | tail -30) > gpurun_out/r2_perf_llm.log 2>&1
(timeout 600 python tools/perf_unet.py -v 2>&1 | tail -60) > gpurun_out/r2_perf_unet.log 2>&1
(SS_UNET_LNFOLD=0 timeout 600 python tools/perf_unet.py 2>&1 | tail -5) > gpurun_out/r2_perf_unet_nofold.log 2>&1
(time timeout 900 python bench.py --gpus 1 --steps 1 --warmup 1 --e2e-steps 1) > gpurun_out/r2_bench_short.log 2>&1
tail -3 gpurun_out/r2_t_*.log; tail -12 gpurun_out/r2_perf_llm.log; tail -4 gpurun_out/r2_perf_unet.log; tail -c 1500 gpurun_out/r2_bench_short.log
