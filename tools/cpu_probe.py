"""Host-CPU probe for sizing bench.py's reference arm: oracle (fp32 torch) timings at several thread counts."""
import os
import sys
import time

import torch

sys.path[:0] = [".", "seed-story_b200"]
from oracle import llama_oracle as LO  # noqa: E402
from oracle import sdxl_oracle as SO  # noqa: E402
from seedstory import synthetic  # noqa: E402

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
a = torch.randn(4096, 4096)
for th in (8, 16, 32, 64, 128):
    if th > (os.cpu_count() or 1):
        break
    torch.set_num_threads(th)
    a @ a
    t0 = time.time()
    for _ in range(3):
        a @ a
    dt = (time.time() - t0) / 3
    print(f"threads {th}: 4096^3 matmul {dt*1e3:.0f} ms = {2*4096**3/dt/1e12:.2f} TFLOP/s")
cfg = dict(SO.SDXL_UNET_CONFIG) if hasattr(SO, "SDXL_UNET_CONFIG") else dict(synthetic.SDXL_UNET_CONFIG)
sd = synthetic.random_unet_state_dict(cfg, seed=1)
p = LO.LlamaParams.random(4096, 11008, 32, 4, 32066, lora_r=16, seed=1)
for th in (16, 32, 64):
    if th > (os.cpu_count() or 1):
        break
    torch.set_num_threads(th)
    with torch.no_grad():
        for B in (1, 2):
            x = torch.randn(B, 4, 128, 128)
            t0 = time.time()
            SO.unet_forward(sd, cfg, x, torch.full((B,), 981.0), torch.randn(B, 64, 2048), torch.randn(B, 1280),
                            torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]] * B))
            print(f"threads {th}: UNet full-res batch {B}: {time.time()-t0:.1f} s", flush=True)
        ctx = 1041
        emb = torch.randn(1, ctx, 4096) * 0.02
        t0 = time.time()
        _, _, kv = LO.model_forward(p, emb, torch.arange(ctx).unsqueeze(0), None)
        print(f"threads {th}: prefill 1041 tokens x 4 layers: {time.time()-t0:.2f} s")
        t0 = time.time()
        for i in range(3):
            LO.model_forward(p, emb[:, :1], torch.tensor([[ctx]]), kv)
        print(f"threads {th}: decode token x 4 layers (+lm_head): {(time.time()-t0)/3*1e3:.0f} ms", flush=True)
