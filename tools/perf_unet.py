"""UNet-only timing (graph replay + event-bracketed eager GEMM/conv/attention shares) — development aid for A/B runs
under the SS_* environment knobs; bench.py is the contract benchmark."""
import sys

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import ops, sdxl_engine, synthetic  # noqa: E402

dev = torch.device("cuda:0")
cfg = synthetic.SDXL_UNET_CONFIG
sd = synthetic.random_unet_state_dict(cfg, seed=0, dtype=torch.float16, device="cuda")
eng = sdxl_engine.UNetEngine(sd, cfg, dev)
del sd
ts, _ = sdxl_engine.euler_schedule(50)
ctx = torch.randn(2, 64, 2048, device=dev).half()
eng.set_conditioning(ctx, torch.randn(2, 1280, device=dev).half(), [[1024, 1024, 0, 0, 1024, 1024]] * 2, ts)
lat = torch.randn(1, 4, 128, 128, device=dev).half()
eng.sample(lat, 3)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    a.record()
    for _ in range(10):
        eng._graph.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"UNet CFG step (graph replay): {ms:.3f} ms -> {2*6.747e12/ms/1e9:.0f} TFLOP/s algorithmic")
if "-v" in sys.argv:
    ops.PROFILE = []
    eng.forward()
    torch.cuda.synchronize()
    agg = {}
    for name, fl, s, e in ops.PROFILE:
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += fl
        d[2] += s.elapsed_time(e)
    ops.PROFILE = None
    for k, (n, fl, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"  {k}: {n} launches, {fl/1e12:.2f} TFLOP, {ms_:.2f} ms (eager, event-bracketed) -> {fl/ms_/1e9:.0f} TFLOP/s")
