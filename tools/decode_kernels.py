"""Per-kernel timing of the Llama decode layer (CUDA graph of nL rotating layers so the weights stream from HBM, CUDA
events): fused kernels against the unfused chain they replace — development aid; bench.py is the contract benchmark."""
import math
import sys

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
H, D, K, I = 32, 128, 4096, 11008
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1041
nL = 6
eps = 1e-5


def rnd(*s, sc=0.02):
    return (torch.randn(*s, device=dev) * sc).half()


Wqkv = [rnd(3 * H * D, K) for _ in range(nL)]
Wil = [ops.interleave_rope_rows(w, H, D) for w in Wqkv]
Wgu = [rnd(2 * I, K) for _ in range(nL)]
Wo = [rnd(K, K) for _ in range(nL)]
Wd = [rnd(K, I) for _ in range(nL)]
gam = (1 + 0.1 * torch.randn(K, device=dev)).half()
x = rnd(B, K, sc=1.0)
xn = torch.empty_like(x)
qkv = torch.empty(B, 3 * H * D, device=dev, dtype=torch.float16)
q = torch.empty(B, H * D, device=dev, dtype=torch.float16)
attn = torch.empty_like(q)
act = torch.empty(B, I, device=dev, dtype=torch.float16)
hbuf = rnd(B, K, sc=1.0)
max_pages = 40
npg = B * max_pages
kc = [rnd(npg, H, 64, D, sc=1.0) for _ in range(nL)]
vc = [rnd(npg, H, 64, D, sc=1.0) for _ in range(nL)]
pt = torch.randperm(npg, device=dev).int().view(B, max_pages).contiguous()
inv = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))
fr = torch.einsum("i,j->ij", torch.arange(4096).float(), inv)
emb = torch.cat((fr, fr), -1)
cos_t, sin_t = emb.cos().half().to(dev), emb.sin().half().to(dev)
seq = torch.arange(B, device=dev, dtype=torch.int32)
pos = torch.full((B,), ctx - 1, device=dev, dtype=torch.int32)
slot = torch.full((B,), ctx - 1, device=dev, dtype=torch.int32)
lens = torch.full((B,), ctx, device=dev, dtype=torch.int32)
scale = 1 / math.sqrt(D)


def timed(name, body, bytes_per_layer):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for li in range(nL):
            body(li)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for li in range(nL):
            body(li)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 10 / nL * 1e3
    print(f"{name:58s} {us:8.2f} us/layer   {bytes_per_layer / us / 1e3:7.0f} GB/s")
    return us


qkv_bytes = 3 * H * D * K * 2
print(f"B={B} ctx={ctx}")
timed("rmsnorm + skinny qkv + rope_append (unfused)", lambda li: (ops.rmsnorm(x, gam, eps, out=xn), ops.skinny_gemm(xn, Wqkv[li], out=qkv),
      ops.rope_kv_append(qkv, q, kc[li], vc[li], seq, pos, slot, pt, cos_t, sin_t, H, D)), qkv_bytes)
timed("skinny qkv alone", lambda li: ops.skinny_gemm(xn, Wqkv[li], out=qkv), qkv_bytes)
timed("rmsnorm alone", lambda li: ops.rmsnorm(x, gam, eps, out=xn), 1)
kv_base = torch.zeros(B, dtype=torch.int64, device=dev)
rcs = torch.zeros(B, D, dtype=torch.float16, device=dev)
rsn = torch.zeros_like(rcs)
ops.decode_rope_meta(seq, pos, slot, pt, cos_t, sin_t, H, D, kv_base, rcs, rsn)
timed("decode_qkv_rope_append (fused)", lambda li: ops.decode_qkv_rope_append(x, gam, eps, Wil[li], q, kc[li], vc[li], kv_base, rcs, rsn,
      H, D), qkv_bytes)
timed("skinny_gemm_rmsnorm qkv (norm prologue only)", lambda li: ops.skinny_gemm_rmsnorm(x, gam, eps, Wqkv[li], out=qkv), qkv_bytes)
gu_bytes = 2 * I * K * 2
timed("rmsnorm + skinny gate_up swiglu (unfused)", lambda li: (ops.rmsnorm(x, gam, eps, out=xn),
      ops.skinny_gemm(xn, Wgu[li], ops.EPI_SWIGLU, out=act)), gu_bytes)
timed("skinny gate_up swiglu alone", lambda li: ops.skinny_gemm(xn, Wgu[li], ops.EPI_SWIGLU, out=act), gu_bytes)
timed("skinny_gemm_rmsnorm gate_up swiglu (fused)", lambda li: ops.skinny_gemm_rmsnorm(x, gam, eps, Wgu[li], ops.EPI_SWIGLU, out=act), gu_bytes)
timed("skinny o_proj + residual", lambda li: ops.skinny_gemm(attn, Wo[li], ops.EPI_RESIDUAL, residual=hbuf, out=hbuf), K * K * 2)
timed("skinny down_proj + residual", lambda li: ops.skinny_gemm(act, Wd[li], ops.EPI_RESIDUAL, residual=hbuf, out=hbuf), K * I * 2)
kv_bytes = B * ctx * H * D * 2 * 2
ws12 = ops.attn_decode_workspace(B, H, D, 12, dev)


def layer_fused(li, with_attn=True):
    ops.decode_qkv_rope_append(x, gam, eps, Wil[li], q, kc[li], vc[li], kv_base, rcs, rsn, H, D)
    if with_attn:
        ops.attn_decode_paged(q, kc[li], vc[li], lens, pt, attn, ws12, H, D, 12, scale)
    ops.skinny_gemm(attn, Wo[li], ops.EPI_RESIDUAL, residual=hbuf, out=hbuf)
    ops.skinny_gemm_rmsnorm(hbuf, gam, eps, Wgu[li], ops.EPI_SWIGLU, out=act)
    ops.skinny_gemm(act, Wd[li], ops.EPI_RESIDUAL, residual=hbuf, out=hbuf)


def layer_unfused(li, with_attn=True):
    ops.rmsnorm(hbuf, gam, eps, out=xn)
    ops.skinny_gemm(xn, Wqkv[li], out=qkv)
    ops.rope_kv_append(qkv, q, kc[li], vc[li], seq, pos, slot, pt, cos_t, sin_t, H, D)
    if with_attn:
        ops.attn_decode_paged(q, kc[li], vc[li], lens, pt, attn, ws12, H, D, 12, scale)
    ops.skinny_gemm(attn, Wo[li], ops.EPI_RESIDUAL, residual=hbuf, out=hbuf)
    ops.rmsnorm(hbuf, gam, eps, out=xn)
    ops.skinny_gemm(xn, Wgu[li], ops.EPI_SWIGLU, out=act)
    ops.skinny_gemm(act, Wd[li], ops.EPI_RESIDUAL, residual=hbuf, out=hbuf)


layer_bytes = qkv_bytes + gu_bytes + K * K * 2 + K * I * 2 + kv_bytes
timed("whole layer, fused (5 launches)", layer_fused, layer_bytes)
timed("whole layer, fused, no attention", lambda li: layer_fused(li, False), layer_bytes - kv_bytes)
timed("whole layer, unfused (8 launches)", layer_unfused, layer_bytes)
timed("whole layer, unfused, no attention", lambda li: layer_unfused(li, False), layer_bytes - kv_bytes)
for splits in (6, 9, 12, 17, 24, 32):
    ws = ops.attn_decode_workspace(B, H, D, splits, dev)
    timed(f"attn_decode_paged splits={splits}", lambda li: ops.attn_decode_paged(q, kc[li], vc[li], lens, pt, attn, ws, H, D, splits, scale), kv_bytes)
