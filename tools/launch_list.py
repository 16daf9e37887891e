"""One eager pass of each stage so that `ncu --metrics gpu__time_duration.sum` lists every launch (development aid)."""
import sys

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import ops, sdxl_engine, story  # noqa: E402

dev = torch.device("cuda:0")
pipe = story.StoryPipeline(device=dev, cfg=story.FULL, num_inference_steps=50)
what = sys.argv[1] if len(sys.argv) > 1 else "unet,decode"
torch.cuda.synchronize()
if "unet" in what:
    ue = pipe.unet.engine()
    ts, _ = sdxl_engine.euler_schedule(50)
    ctx = torch.randn(2, 64, 2048, device=dev).half()
    ue.set_conditioning(ctx, torch.randn(2, 1280, device=dev).half(), [[1024, 1024, 0, 0, 1024, 1024]] * 2, ts)
    ue.temb_cur.copy_(ue.temb_all[0:2])
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("unet_forward")
    ue.forward()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
if "decode" in what:
    eng = pipe.agent.llm.engine(max_new=512)
    eng.set_image_token_ids(pipe.image_ids, 2)
    x = (torch.randn(1041, 4096, device=dev) * 0.02).half()
    eng.reset_sequence(0)
    eng.forward_chunk(0, x, list(range(1041)))
    eng.begin_decode([5], [1041])
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("decode_step")
    eng.decode_step(1, use_graph=False)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done")
