"""Per-shape GEMM/conv timing inside a CUDA graph (20 back-to-back launches, PDL overlap as in the real step) for
each N-tile choice — development aid used to tune the tile cost model in csrc/gemm_tc.cu."""
import sys

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import ops  # noqa: E402

dev = torch.device("cuda:0")
REP = 20


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


gemms = [(2048, 1280, 64, 0, -1), (8192, 640, 64, 0, -1), (2048, 1280, 1280, 0, 192), (2048, 1280, 5120, 0, 60), (2048, 3840, 1280, 0, 60), (2048, 10240, 1280, 1, 60),
         (8192, 640, 640, 0, 40), (8192, 5120, 640, 1, 10), (8192, 640, 2560, 0, 10), (8192, 1920, 640, 0, 10),
         (1041, 12288, 4096, 0, 0), (1041, 4096, 4096, 0, 0), (1041, 22016, 4096, 2, 0), (1041, 4096, 11008, 0, 0),
         (66, 12288, 4096, 0, 0), (1024, 4992, 1664, 0, 0), (1024, 1664, 1664, 0, 0), (1024, 8192, 1664, 0, 0)]
for (M, N, K, glu, cnt) in gemms:
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    bias = torch.randn(N, device=dev).half()
    n_out = N // 2 if glu else N
    res = None if (glu or cnt < 0) else torch.randn(M, n_out, device=dev).half()
    if cnt < 0:
        bias = None
    out = torch.empty(M, n_out, device=dev).half()
    row = []
    for bn in (0, 160, 256, 1256):
        us = graph_time(lambda: ops.gemm(a, w, bias=bias, residual=res, glu=glu, out=out, force_bn=bn))
        row.append(f"bn{bn}: {us:7.2f} us {2.0*M*N*K/us/1e6:6.0f} TF")
    print(f"gemm {M}x{N}x{K} glu={glu} x{cnt}: " + " | ".join(row), flush=True)

convs = [(2, 32, 32, 1280, 1280, 10), (2, 128, 128, 320, 320, 7), (2, 64, 64, 640, 640, 6), (2, 128, 128, 640, 320, 2),
         (2, 32, 32, 2560, 1280, 2), (1, 256, 256, 256, 256, 0), (1, 512, 512, 256, 256, 0), (1, 1024, 1024, 128, 128, 0)]
for (Ni, H, W, Cin, Cout, cnt) in convs:
    dt = torch.float16 if cnt else torch.bfloat16
    x = torch.randn(Ni, H, W, Cin, device=dev).to(dt)
    w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).to(dt)
    bias = torch.randn(Cout, device=dev).to(dt)
    out = torch.empty(Ni, H, W, Cout, device=dev).to(dt)
    row = []
    for bn in (0, 160, 256, 1256):
        us = graph_time(lambda: ops.conv3x3(x, w, bias=bias, out=out, force_bn=bn))
        row.append(f"bn{bn}: {us:7.2f} us {2.0*Ni*H*W*Cout*9*Cin/us/1e6:6.0f} TF")
    print(f"conv {Ni}x{H}x{W} {Cin}->{Cout} x{cnt}: " + " | ".join(row), flush=True)
