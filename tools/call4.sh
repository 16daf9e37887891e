#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python tools/determinism.py 2>&1 | tail -60) > gpurun_out/r2_determinism.log 2>&1
for c in 1 0; do
  (SS_CARVEOUT=$c timeout 300 python tools/perf_llm.py 2>&1 | tail -8) > gpurun_out/r2_perf_llm_carve$c.log 2>&1
  for f in 1 0; do
    (SS_CARVEOUT=$c SS_UNET_LNFOLD=$f timeout 300 python tools/perf_unet.py 2>&1 | tail -3) > gpurun_out/r2_perf_unet_carve${c}_fold$f.log 2>&1
  done
done
(timeout 300 python tools/decode_kernels.py 1 1041 2>&1 | tail -30) > gpurun_out/r2_decode_kernels2.log 2>&1
(timeout 900 python -m pytest "tests/test_fullsize_gpu.py::test_unet_full_width_blocks_match_oracle" "tests/test_fullsize_gpu.py::test_dropins_generate_and_get_image_embeds_match_oracle" tests/test_sdxl_gpu.py::test_story_pipeline_tiny_end_to_end -q --timeout 600 2>&1 | tail -80) > gpurun_out/r2_t_retry.log 2>&1
(time timeout 900 python bench.py --gpus 1 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline) > gpurun_out/r2_bench_short2.log 2>&1
cat gpurun_out/r2_determinism.log; tail -4 gpurun_out/r2_perf_llm_carve*.log; tail -2 gpurun_out/r2_perf_unet_carve*.log; tail -3 gpurun_out/r2_t_retry.log
