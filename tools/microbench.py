"""Ad-hoc microbenchmarks (CUDA events) for the first GPU runs; bench.py is the contract benchmark."""
import math
import sys

import torch

sys.path.insert(0, "seed-story_b200")
from seedstory import ops  # noqa: E402

dev = torch.device("cuda:0")
ops.require_device()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


SECTIONS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["skinny", "gemm", "conv", "fmha"]
print("== skinny GEMM (decode) bandwidth; rotating over 8 weight copies (> L2) ==")
for (N, K, name) in [] if "skinny" not in SECTIONS else [(12288, 4096, "qkv"), (4096, 4096, "o"), (22016, 4096, "gate_up"), (4096, 11008, "down"),
                     (32066, 4096, "lm_head")]:
    Ws = [(torch.randn(N, K, device=dev) * 0.02).half() for _ in range(8)]
    for B in (1, 8):
        x = torch.randn(B, K, device=dev).half()
        i = [0]

        def f():
            ops.skinny_gemm(x, Ws[i[0] % 8])
            i[0] += 1
        t = timeit(f, iters=40)
        print(f"{name:8s} B={B} N={N} K={K}: {t*1e6:8.1f} us  {N*K*2/t/1e9:8.1f} GB/s")
    del Ws

print("== tcgen05 GEMM ==")
for (M, N, K) in [] if "gemm" not in SECTIONS else [(1024, 4096, 4096), (4096, 4096, 4096), (8192, 8192, 8192), (2048, 1280, 1280), (8192, 640, 640),
                  (1024, 8192, 1664), (2048, 10240, 1280), (1041, 12288, 4096), (2048, 1280, 5120), (2048, 3840, 1280), (8192, 5120, 640), (66, 12288, 4096)]:
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    out = torch.empty(M, N, device=dev).half()
    for bn in (64, 128, 256):
        t = timeit(lambda: ops.gemm(a, w, out=out, force_bn=bn))
        print(f"gemm {M}x{N}x{K} bn={bn}: {t*1e6:8.1f} us  {2*M*N*K/t/1e12:7.1f} TFLOP/s")
    t = timeit(lambda: torch.matmul(a, w.t()))
    print(f"  cuBLAS ref          : {t*1e6:8.1f} us  {2*M*N*K/t/1e12:7.1f} TFLOP/s")

print("== conv3x3 implicit GEMM ==")
for (Nimg, H, W, Cin, Cout) in [] if "conv" not in SECTIONS else [(2, 128, 128, 320, 320), (2, 64, 64, 640, 640), (2, 32, 32, 1280, 1280),
                                (2, 32, 32, 2560, 1280), (1, 1024, 1024, 128, 128), (1, 512, 512, 256, 256)]:
    x = torch.randn(Nimg, H, W, Cin, device=dev).half()
    w = (torch.randn(Cout, 9 * Cin, device=dev) / math.sqrt(9 * Cin)).half()
    out = torch.empty(Nimg, H, W, Cout, device=dev).half()
    for bn in (128, 256):
        t = timeit(lambda: ops.conv3x3(x, w, out=out, force_bn=bn))
        fl = 2.0 * Nimg * H * W * Cout * 9 * Cin
        print(f"conv {Nimg}x{H}x{W} {Cin}->{Cout} bn={bn}: {t*1e6:8.1f} us  {fl/t/1e12:7.1f} TFLOP/s")

print("== fmha ==")
for (B, H, L, D, causal) in [] if "fmha" not in SECTIONS else [(2, 10, 4096, 64, False), (2, 20, 1024, 64, False), (1, 16, 1024, 128, False),
                             (1, 32, 1041, 128, True)]:
    q = torch.randn(B, L, H * D, device=dev).half()
    k = torch.randn(B, L, H * D, device=dev).half()
    v = torch.randn(B, L, H * D, device=dev).half()
    out = torch.empty_like(q)
    t = timeit(lambda: ops.mha_packed(q, k, v, H, 1 / math.sqrt(D), causal, out=out))
    fl = 4.0 * B * H * L * L * D * (0.5 if causal else 1.0)
    print(f"fmha B{B} H{H} L{L} D{D} causal={causal}: {t*1e6:8.1f} us  {fl/t/1e12:7.1f} TFLOP/s")
