"""Times the ORACLE (torch restatement of the reference path) in fp16 on the GPU: eager PyTorch with cuBLAS / cuDNN /
SDPA, unmerged LoRA, per-step python greedy loop — i.e. the reference's own behaviour restated.
Development aid / context number; not part of bench.py's contract."""
import sys
import time

import torch

sys.path[:0] = [".", "seed-story_b200"]
from oracle import llama_oracle as LO  # noqa: E402
from oracle import sdxl_oracle as SO  # noqa: E402
from seedstory import synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / n


with torch.no_grad():
    cfg = SO.SDXL_UNET_CONFIG
    sd = {k: v.to(dev, torch.float16) for k, v in synthetic.random_unet_state_dict(cfg, seed=1).items()}
    x = torch.randn(2, 4, 128, 128, device=dev).half()
    ctx = torch.randn(2, 64, 2048, device=dev).half()
    pooled = torch.randn(2, 1280, device=dev).half()
    tid = torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]] * 2, device=dev).half()
    SO_timestep = SO.timestep_embedding

    def te(t, dim, max_period=10000):
        return SO_timestep(t.cpu(), dim, max_period).to(dev)
    SO.timestep_embedding = te
    t_unet = timed(lambda: SO.unet_forward(sd, cfg, x, torch.tensor([981.0, 981.0], device=dev), ctx, pooled, tid))
    print(f"oracle UNet CFG step fp16 eager on GPU: {t_unet*1e3:.1f} ms")
    del sd
    torch.cuda.empty_cache()

    p = LO.LlamaParams.random(4096, 11008, 32, 32, 32066, lora_r=16, seed=1, dtype=torch.float16)
    p = p.to(dtype=torch.float16, device=dev)
    L = 1041
    emb = (torch.randn(1, L, 4096, device=dev) * 0.02).half()
    LO_rope = LO.rope_tables

    def rope_dev(d, n, base=10000.0):
        c, s = LO_rope(d, n, base)
        return c.to(dev), s.to(dev)
    LO.rope_tables = rope_dev
    orig_attend = LO.attend_bottom_right

    def attend(q, k, v):
        tq, tk = q.shape[-2], k.shape[-2]
        qi = torch.arange(tq, device=q.device).unsqueeze(1)
        kj = torch.arange(tk, device=q.device).unsqueeze(0)
        return torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=(kj <= qi + (tk - tq)))
    LO.attend_bottom_right = attend
    pos = torch.arange(L, device=dev).unsqueeze(0)
    t_pre = timed(lambda: LO.model_forward(p, emb, pos, None), n=2)
    _, _, kv = LO.model_forward(p, emb, pos, None)
    one = emb[:, :1]

    def step():
        lg, hn, _ = LO.model_forward(p, one, torch.tensor([[L]], device=dev), kv)
        return int(lg[0, -1].argmax().item())
    t_dec = timed(step, n=5)
    print(f"oracle Llama prefill {L} tokens: {t_pre*1e3:.1f} ms; decode step (32 layers, unmerged LoRA, python loop): {t_dec*1e3:.2f} ms")
    turn = 131 * t_dec + t_pre + 50 * t_unet
    print(f"restated reference GPU path, one turn (131 decode steps + re-prefill + 50 UNet CFG steps; ViT/VAE excluded): "
          f"{turn:.2f} s -> {1/turn:.3f} turns/s")
