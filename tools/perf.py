"""Stage-level timing of the full-size pipeline (CUDA events) — development aid; bench.py is the contract benchmark."""
import sys
import time

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import _capi, ops, story  # noqa: E402

dev = torch.device("cuda:0")
t0 = time.time()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
pipe = story.StoryPipeline(device=dev, cfg=story.FULL, num_inference_steps=steps, verbose=True)
print(f"build {time.time()-t0:.1f}s, mem {torch.cuda.memory_allocated()/2**30:.1f} GiB")


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, n=1):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(n):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r


img = torch.randn(1, 3, 448, 448, device=dev).half()
cap = list(range(100, 164))
for rep in range(2):
    ms, emb = timed(lambda: pipe.visual_encoder(img))
    print(f"ViT forward: {ms:.2f} ms  ({4.22e12/ms/1e9:.0f} TFLOP/s)")

# LLM: prefill 1041 tokens, decode steps, 66-chunk
llm = pipe.agent.llm
eng = llm.engine(max_new=512)
eng.set_image_token_ids(pipe.image_ids, 2)
L = 1041
x = (torch.randn(L, 4096, device=dev) * 0.02).half()
for rep in range(2):
    eng.reset_sequence(0)
    ms, _ = timed(lambda: eng.forward_chunk(0, x, list(range(L))))
    print(f"prefill {L} tokens: {ms:.2f} ms ({L*12.95e9/ms/1e9:.0f} TFLOP/s)")
eng.begin_decode([5], [L])
eng.decode_step(1)  # capture
for rep in range(2):
    ms, _ = timed(lambda: eng.decode_step(1), n=20)
    print(f"decode step (graph, ctx~{L}): {ms:.3f} ms  -> {13.215e9/ms/1e6:.0f} GB/s weight streaming")
ms, _ = timed(lambda: eng.decode_step(1, use_graph=False), n=5)
print(f"decode step (eager launches): {ms:.3f} ms")
eng.seq_len_h[0] += 27
xc = (torch.randn(66, 4096, device=dev) * 0.02).half()
p0 = eng.seq_len_h[0]
for rep in range(2):
    ms, _ = timed(lambda: eng.forward_chunk(0, xc, list(range(p0, p0 + 66))))
    p0 += 66
    print(f"66-token chunk: {ms:.2f} ms")

# resamplers
ms, _ = timed(lambda: pipe.agent.input_resampler(emb))
print(f"input resampler: {ms:.3f} ms")

# SDXL
feat = torch.randn(1, 256, 4096, device=dev).half()
for rep in range(2):
    torch.cuda.synchronize()
    t1 = time.time()
    ms, out = timed(lambda: pipe.adapter.generate(image_embeds=feat, num_inference_steps=steps, output_type="pt"))
    print(f"adapter.generate ({steps} steps): {ms:.1f} ms (wall {1e3*(time.time()-t1):.1f} ms)")
ue = pipe.unet.engine()
ms, _ = timed(lambda: ue._graph.replay(), n=10)
print(f"UNet CFG step (graph replay): {ms:.3f} ms -> {2*6.747e12/ms/1e9:.0f} TFLOP/s algorithmic")
ops.PROFILE = []
ue.forward()
torch.cuda.synchronize()
agg = {}
for name, fl, a, b in ops.PROFILE:
    d = agg.setdefault(name, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += fl
    d[2] += a.elapsed_time(b)
ops.PROFILE = None
for k, (n, fl, ms_) in agg.items():
    print(f"  {k}: {n} launches, {fl/1e12:.2f} TFLOP, {ms_:.2f} ms (event-bracketed, eager) -> {fl/ms_/1e9:.0f} TFLOP/s")
ve = pipe.vae.engine()
lat = torch.randn(128 * 128, 4, device=dev).half() * 0.2
for rep in range(2):
    ms, _ = timed(lambda: ve.decode(lat, 128))
    print(f"VAE decode: {ms:.2f} ms ({10.47e12/ms/1e9:.0f} TFLOP/s)")

# one whole story turn x3
for rep in range(2):
    torch.cuda.synchronize()
    t1 = time.time()
    c0 = _capi.launch_count()
    pipe.run_story(img, cap, 3)
    torch.cuda.synchronize()
    print(f"3-turn story: {time.time()-t1:.2f} s wall, {(_capi.launch_count()-c0)} kernel launches")
