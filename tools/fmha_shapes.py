"""Per-shape attention timing inside a CUDA graph (development aid)."""
import math
import sys

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import ops  # noqa: E402

dev = torch.device("cuda:0")
REP = 10


def graph_time(fn):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * REP) * 1e3  # us


for (B, H, Lq, Lk, D, cnt) in [(2, 20, 1024, 1024, 64, 60), (2, 10, 4096, 4096, 64, 10), (2, 20, 1024, 64, 64, 60),
                               (2, 10, 4096, 64, 64, 10), (1, 16, 1024, 1024, 128, 48)]:
    C = H * D
    if Lq == Lk:
        qkv = torch.randn(B * Lq, 3 * C, device=dev).half()
        q, k, v = qkv[:, :C].view(B, Lq, C), qkv[:, C:2 * C].view(B, Lk, C), qkv[:, 2 * C:].view(B, Lk, C)
    else:
        q = torch.randn(B, Lq, C, device=dev).half()
        kv = torch.randn(B, Lk, 2 * C, device=dev).half()
        k, v = kv[..., :C], kv[..., C:]
    us = graph_time(lambda: ops.mha_packed(q, k, v, H, 1.0 / math.sqrt(D), False))
    fl = 4.0 * B * H * Lq * Lk * D
    print(f"fmha B={B} H={H} Lq={Lq} Lk={Lk} D={D} x{cnt}: {us:8.2f} us  {fl/us/1e6:6.0f} TF", flush=True)
