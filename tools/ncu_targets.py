"""Representative launches of every hot kernel for `ncu --set full` captures (development aid)."""
import math
import sys

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def rnd(*s, sc=1.0):
    return (torch.randn(*s, device=dev) * sc).half()


if which in ("all", "gemm"):
    for (M, N, K, glu) in [(2048, 1280, 1280, 0), (2048, 10240, 1280, 1), (8192, 640, 640, 0), (2048, 1280, 5120, 0),
                           (1041, 12288, 4096, 0), (8192, 8192, 8192, 0)]:
        a, w = rnd(M, K), rnd(N, K, sc=1 / math.sqrt(K))
        bias = rnd(N)
        for _ in range(3):
            ops.gemm(a, w, bias=bias, glu=glu)
    for (Nimg, H, W, Cin, Cout) in [(2, 64, 64, 640, 640), (2, 32, 32, 1280, 1280), (2, 128, 128, 320, 320)]:
        x, w = rnd(Nimg, H, W, Cin), rnd(Cout, 9 * Cin, sc=1 / math.sqrt(9 * Cin))
        for _ in range(3):
            ops.conv3x3(x, w, bias=rnd(Cout))
if which == "gemm2":
    for (M, N, K, glu, res) in [(2048, 10240, 1280, 1, 0), (2048, 1280, 1280, 0, 1), (2048, 3840, 1280, 0, 0)]:
        a, w = rnd(M, K), rnd(N, K, sc=1 / math.sqrt(K))
        bias = rnd(N)
        r = rnd(M, N) if res else None
        for _ in range(2):
            ops.gemm(a, w, bias=bias, glu=glu, residual=r)
    q, k, v = rnd(2, 1024, 1280), rnd(2, 1024, 1280), rnd(2, 1024, 1280)
    for _ in range(2):
        ops.mha_packed(q, k, v, 20, 0.125)
if which in ("all", "skinny"):
    for (N, K) in [(4096, 4096), (4096, 11008), (32066, 4096)]:
        Ws = [rnd(N, K, sc=0.02) for _ in range(3)]
        x = rnd(1, K)
        res = rnd(1, N)
        for W_ in Ws:
            ops.skinny_gemm(x, W_, ops.EPI_RESIDUAL, residual=res)
    # the fused decode-layer kernels: RMSNorm + q/k/v + RoPE + append, RMSNorm + gate/up + SwiGLU
    Hh, D, K = 32, 128, 4096
    x, gam = rnd(1, K), rnd(K)
    kc, vc = rnd(40, Hh, 64, D), rnd(40, Hh, 64, D)
    q = torch.empty(1, Hh * D, device=dev, dtype=torch.float16)
    kvb = torch.zeros(1, dtype=torch.int64, device=dev)
    rcs, rsn = rnd(1, D), rnd(1, D)
    for _ in range(3):
        ops.decode_qkv_rope_append(x, gam, 1e-5, rnd(3 * Hh * D, K, sc=0.02), q, kc, vc, kvb, rcs, rsn, Hh, D)
    for _ in range(3):
        ops.skinny_gemm_rmsnorm(x, gam, 1e-5, rnd(22016, K, sc=0.02), ops.EPI_SWIGLU)
if which == "fmha1":
    q, k, v = rnd(2, 4096, 640), rnd(2, 4096, 640), rnd(2, 4096, 640)
    for _ in range(2):
        ops.mha_packed(q, k, v, 10, 0.125)
if which in ("all", "attn"):
    for (B, H, L, D) in [(2, 10, 4096, 64), (2, 20, 1024, 64), (1, 16, 1024, 128)]:
        q, k, v = rnd(B, L, H * D), rnd(B, L, H * D), rnd(B, L, H * D)
        for _ in range(2):
            ops.mha_packed(q, k, v, H, 1 / math.sqrt(D))
    Hh, D, B = 32, 128, 1
    kc = rnd(40, Hh, 64, D)
    vc = rnd(40, Hh, 64, D)
    pt = torch.arange(32, device=dev, dtype=torch.int32).view(1, 32)
    q = rnd(B, Hh * D)
    out = torch.empty_like(q)
    ws = ops.attn_decode_workspace(B, Hh, D, 12, dev)
    sl = torch.tensor([1100], device=dev, dtype=torch.int32)
    for _ in range(2):
        ops.attn_decode_paged(q, kc, vc, sl, pt, out, ws, Hh, D, 12, 0.088)
    x = rnd(2, 64, 64, 640)
    g, b = rnd(640), rnd(640)
    wsn = ops.groupnorm_ws(2, 64 * 64, 640, 32, dev)
    for _ in range(2):
        ops.groupnorm_nhwc(x, g, b, 32, 1e-5, True, wsn)
torch.cuda.synchronize()
print("done")
