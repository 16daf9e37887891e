"""Op-level parity of the CUDA kernels (through the C-ABI) against plain fp32 torch math."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol=2e-3, atol=None, what=""):
    got = got.float()
    ref = ref.float()
    if atol is None:
        atol = 2e-3 * ref.abs().max().item() + 1e-6
    err = (got - ref).abs()
    bad = err > (atol + rtol * ref.abs())
    assert not bad.any(), f"{what}: {bad.sum().item()} / {bad.numel()} out of tolerance, max err {err.max().item():.4g} (atol {atol:.3g})"


def test_rmsnorm(cuda_dev):
    from seedstory import ops
    torch.manual_seed(0)
    for rows, K in [(1, 4096), (5, 4096), (3, 11008), (66, 4096)]:
        x = torch.randn(rows, K, device=cuda_dev).half()
        w = (1 + 0.1 * torch.randn(K, device=cuda_dev)).half()
        y = ops.rmsnorm(x, w, 1e-5)
        var = x.float().pow(2).mean(-1, keepdim=True)
        ref = w * (x * torch.rsqrt(var + 1e-5)).half()
        _close(y, ref, rtol=2e-3, atol=1e-3, what=f"rmsnorm {rows}x{K}")  # 1 fp16 ulp


def test_layernorm(cuda_dev):
    from seedstory import ops
    torch.manual_seed(0)
    for dt in (torch.float16, torch.bfloat16):
        x = torch.randn(77, 1664, device=cuda_dev).to(dt)
        g = (1 + 0.1 * torch.randn(1664, device=cuda_dev)).to(dt)
        b = (0.1 * torch.randn(1664, device=cuda_dev)).to(dt)
        add = torch.randn(7, 1664, device=cuda_dev).to(dt)
        y, y2 = ops.layernorm(x, g, b, 1e-6, add=add)
        ref = torch.nn.functional.layer_norm(x.float(), (1664,), g.float(), b.float(), 1e-6)
        tol = 2e-3 if dt == torch.float16 else 1.6e-2
        _close(y, ref, rtol=tol, atol=tol, what="layernorm")
        ref2 = ref.to(dt).float() + add.float().repeat(11, 1)
        _close(y2, ref2, rtol=tol, atol=2 * tol, what="layernorm+add")


@pytest.mark.parametrize("B", [1, 3, 8])
def test_skinny_gemm(cuda_dev, B):
    from seedstory import ops
    torch.manual_seed(B)
    for N, K in [(4096, 4096), (12288, 4096), (4096, 11008), (32066, 4096)]:
        x = torch.randn(B, K, device=cuda_dev).half()
        W = (torch.randn(N, K, device=cuda_dev) * 0.02).half()
        ref = x.float() @ W.float().t()
        y = ops.skinny_gemm(x, W)
        _close(y, ref, what=f"skinny {B}x{N}x{K}")
        res = torch.randn(B, N, device=cuda_dev).half()
        y = ops.skinny_gemm(x, W, ops.EPI_RESIDUAL, residual=res)
        _close(y, ref.half().float() + res.float(), what="skinny+res")
    # SwiGLU with pairwise interleaved gate/up rows
    K, I = 4096, 11008
    x = torch.randn(B, K, device=cuda_dev).half()
    Wg = (torch.randn(I, K, device=cuda_dev) * 0.02).half()
    Wu = (torch.randn(I, K, device=cuda_dev) * 0.02).half()
    Wp = torch.stack([Wg, Wu], dim=1).reshape(2 * I, K).contiguous()
    y = ops.skinny_gemm(x, Wp, ops.EPI_SWIGLU)
    g = (x.float() @ Wg.float().t()).half()
    u = (x.float() @ Wu.float().t()).half()
    ref = torch.nn.functional.silu(g.float()).half() * u
    _close(y, ref, what="skinny swiglu")


def _rope_tables(D, maxpos, dev):
    inv = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))
    t = torch.arange(maxpos).float()
    fr = torch.einsum("i,j->ij", t, inv)
    emb = torch.cat((fr, fr), -1)
    return emb.cos().half().to(dev), emb.sin().half().to(dev)


def test_rope_append_and_decode_attention(cuda_dev):
    from seedstory import ops
    torch.manual_seed(1)
    H, D, B = 32, 128, 3
    lens = [131, 1041, 700]
    max_pages = 32
    npages = B * max_pages
    kc = torch.zeros(npages, H, 64, D, device=cuda_dev, dtype=torch.float16)
    vc = torch.zeros_like(kc)
    # scrambled page table: sequence b uses pages in a permuted order
    perm = torch.randperm(npages, device=cuda_dev).int().view(B, max_pages).contiguous()
    cos_t, sin_t = _rope_tables(D, 4096, cuda_dev)
    ks, vs = [], []
    for b, n in enumerate(lens):
        qkv = torch.randn(n, 3 * H * D, device=cuda_dev).half()
        pos = torch.arange(n, device=cuda_dev, dtype=torch.int32) + 5 * b  # position != slot on purpose
        slot = torch.arange(n, device=cuda_dev, dtype=torch.int32)
        seq = torch.full((n,), b, device=cuda_dev, dtype=torch.int32)
        qo = torch.empty(n, H * D, device=cuda_dev, dtype=torch.float16)
        ops.rope_kv_append(qkv, qo, kc, vc, seq, pos, slot, perm, cos_t, sin_t, H, D)
        q, k, v = qkv.view(n, 3, H, D).unbind(1)
        cos = cos_t[pos.long()][:, None, :]
        sin = sin_t[pos.long()][:, None, :]

        def rot(x):
            return torch.cat((-x[..., D // 2:], x[..., :D // 2]), -1)
        q_ref = (q * cos) + (rot(q) * sin)
        k_ref = (k * cos) + (rot(k) * sin)
        assert torch.equal(qo.view(n, H, D), q_ref), "rope(q) must be bit-exact"
        # read back the cache through the page table
        pages = perm[b, : (n + 63) // 64].long()
        k_back = kc[pages].permute(0, 2, 1, 3).reshape(-1, H, D)[:n]
        v_back = vc[pages].permute(0, 2, 1, 3).reshape(-1, H, D)[:n]
        assert torch.equal(k_back, k_ref) and torch.equal(v_back, v)
        ks.append(k_ref)
        vs.append(v)
    # decode attention: one query per sequence attends to everything cached
    q = torch.randn(B, H * D, device=cuda_dev).half()
    seq_lens = torch.tensor(lens, device=cuda_dev, dtype=torch.int32)
    for splits in (1, 2, 8, 12, 32):   # 1 / 2: several pages per CTA (stage refill); 32: empty splits
        out = torch.empty(B, H * D, device=cuda_dev, dtype=torch.float16)
        ws = ops.attn_decode_workspace(B, H, D, splits, cuda_dev)
        ops.attn_decode_paged(q, kc, vc, seq_lens, perm, out, ws, H, D, splits, 1.0 / math.sqrt(D))
        for b, n in enumerate(lens):
            qq = q[b].view(H, 1, D).float()
            kk = ks[b].permute(1, 0, 2).float()
            vv = vs[b].permute(1, 0, 2).float()
            p = torch.softmax(qq @ kk.transpose(1, 2) / math.sqrt(D), -1)
            ref = (p @ vv).reshape(H * D)
            _close(out[b], ref, rtol=2e-3, atol=2e-3, what=f"decode attn b={b} splits={splits}")


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 4, 8])
def test_fused_decode_layer_kernels_equal_the_unfused_chain(cuda_dev, B):
    """RMSNorm folded into the q/k/v and gate/up projections, RoPE + cache append folded into the q/k/v epilogue:
    bit-identical to rmsnorm -> skinny GEMM -> rope_kv_append (same rounding chain, same reduction order)."""
    from seedstory import ops
    torch.manual_seed(10 + B)
    H, D, K, I = 32, 128, 4096, 11008
    eps = 1e-5
    x = (torch.randn(B, K, device=cuda_dev) * 1.7).half()
    gamma = (1.0 + 0.1 * torch.randn(K, device=cuda_dev)).half()
    xn = ops.rmsnorm(x, gamma, eps)
    # gate/up + SwiGLU, and a plain projection
    Wp = (torch.randn(2 * I, K, device=cuda_dev) * 0.02).half()
    assert torch.equal(ops.skinny_gemm_rmsnorm(x, gamma, eps, Wp, ops.EPI_SWIGLU), ops.skinny_gemm(xn, Wp, ops.EPI_SWIGLU))
    assert torch.equal(ops.skinny_gemm_rmsnorm(x, gamma, eps, Wp[:4100]), ops.skinny_gemm(xn, Wp[:4100]))
    # q/k/v + RoPE + append
    Wqkv = (torch.randn(3 * H * D, K, device=cuda_dev) * 0.02).half()
    Wil = ops.interleave_rope_rows(Wqkv, H, D)
    max_pages = 8
    npages = B * max_pages
    perm = torch.randperm(npages, device=cuda_dev).int().view(B, max_pages).contiguous()
    cos_t, sin_t = _rope_tables(D, 4096, cuda_dev)
    seq = torch.arange(B, device=cuda_dev, dtype=torch.int32)
    pos = torch.tensor([17 + 101 * b for b in range(B)], device=cuda_dev, dtype=torch.int32)
    slot = torch.tensor([(63 + 37 * b) % (max_pages * 64) for b in range(B)], device=cuda_dev, dtype=torch.int32)
    kc1 = torch.zeros(npages, H, 64, D, device=cuda_dev, dtype=torch.float16)
    vc1 = torch.zeros_like(kc1)
    kc2, vc2 = torch.zeros_like(kc1), torch.zeros_like(kc1)
    q1 = torch.empty(B, H * D, device=cuda_dev, dtype=torch.float16)
    q2 = torch.zeros_like(q1)
    qkv = ops.skinny_gemm(xn, Wqkv)
    ops.rope_kv_append(qkv, q1, kc1, vc1, seq, pos, slot, perm, cos_t, sin_t, H, D)
    kv_base = torch.zeros(B, dtype=torch.int64, device=cuda_dev)
    rcs = torch.zeros(B, D, dtype=torch.float16, device=cuda_dev)
    rsn = torch.zeros_like(rcs)
    ops.decode_rope_meta(seq, pos, slot, perm, cos_t, sin_t, H, D, kv_base, rcs, rsn)
    assert torch.equal(rcs, cos_t[pos.long()]) and torch.equal(rsn, sin_t[pos.long()])
    ops.decode_qkv_rope_append(x, gamma, eps, Wil, q2, kc2, vc2, kv_base, rcs, rsn, H, D)
    assert torch.equal(q1, q2), "fused q (RoPE) differs"
    assert torch.equal(kc1, kc2) and torch.equal(vc1, vc2), "fused K/V append differs"
    assert kc2.abs().sum() > 0


def test_logits_processor_argmax(cuda_dev):
    from seedstory import ops
    torch.manual_seed(2)
    V = 32066
    img_ids = torch.tensor([32000] + list(range(32002, 32066)) + [32001], device=cuda_dev, dtype=torch.int32)
    B = 4
    logits = torch.randn(B, V, device=cuda_dev).half()
    logits[0, 32010] = 50.0          # would win, but gets zeroed (not in an image run)
    logits[0, :32001] -= 10.0        # everything else (incl. the untouched BOI) negative -> zeroed ids win at 0.0, lowest index first
    last = torch.tensor([17, 32000, 32002 + 63, 32001], device=cuda_dev, dtype=torch.int32)
    ref_logits = logits.clone()
    nxt = torch.empty(B, device=cuda_dev, dtype=torch.int32)
    ops.logits_process_argmax(logits, last, img_ids, nxt)
    ids = img_ids.tolist()
    exp = []
    for b in range(B):
        row = ref_logits[b].clone()
        cur = last[b].item()
        if cur in ids[:-1]:
            row[ids[ids.index(cur) + 1]] = row.max() + 10.0
        else:
            row[torch.tensor(ids[1:], device=cuda_dev).long()] = 0.0
        exp.append(int(torch.argmax(row.float()).item()))
        assert torch.equal(row, logits[b]), "in-place edit of the logits row must match the reference processor"
    assert nxt.tolist() == exp, (nxt.tolist(), exp)
    assert exp[0] == 32001 and exp[1] == 32002 and exp[2] == 32001
    # processor disabled
    ops.logits_process_argmax(ref_logits, last, None, nxt)
    assert nxt.tolist() == torch.argmax(ref_logits.float(), -1).tolist()
    # SuppressTokensLogitsProcessor after the image processor: scores[:, ids] = -inf, in place
    from src.models_clm.generation import AutoImageTokenGenerationProcessor, SuppressTokensProcessor
    logits = torch.randn(B, V, device=cuda_dev).half()
    win = torch.argmax(logits[0].float()).item()
    sup = torch.tensor([2, 32000, win], device=cuda_dev, dtype=torch.int32)
    logits[3, 2] = 40.0               # EOS would win row 3
    ref = logits.clone().cpu()
    ops.logits_process_argmax(logits, last, img_ids, nxt, sup)

    class _Tok:
        def encode(self, text, add_special_tokens=False):
            return ids
    host = AutoImageTokenGenerationProcessor(_Tok(), 64)
    sp = SuppressTokensProcessor(sup.tolist())
    for b in range(B):
        row = sp(None, host(last[b:b + 1].cpu().view(1, 1).long(), ref[b:b + 1].clone()))[0]
        assert torch.equal(row, logits[b].cpu()), f"row {b}: in-place edits differ from the host processors"
        assert nxt[b].item() == int(torch.argmax(row.float()).item())
    assert nxt[0].item() != win and nxt[3].item() != 2


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_gemm_tn_shapes(cuda_dev, dt):
    from seedstory import ops
    torch.manual_seed(3)
    tol = 2e-3 if dt == torch.float16 else 1.6e-2
    shapes = [(128, 128, 64), (128, 256, 512), (200, 320, 328), (1041, 4096, 4096), (1024, 4992, 1664),
              (64, 4096, 4096), (4096, 640, 640), (333, 72, 200), (2048, 1280, 1280), (300, 168, 136)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device=cuda_dev).to(dt)
        w = (torch.randn(N, K, device=cuda_dev) / math.sqrt(K)).to(dt)
        ref = a.float() @ w.float().t()
        torch.cuda.synchronize()
        for bn in (0, 64, 128, 160, 256, 1256):   # 1256: CTA-pair kernel (256-wide pair tile)
            # alternate the weight-prefetch-before-PDL-wait path (load-time weights) and the activation-operand path
            c = ops.gemm(a, w, force_bn=bn, w_const=(bn % 128 == 0))
            _close(c, ref, rtol=tol, atol=tol * ref.abs().max().item(), what=f"gemm {M}x{N}x{K} bn={bn} {dt}")


@pytest.mark.parametrize("bn,M", [(0, 300), (160, 300), (160, 128 * 150 + 17), (256, 300), (1256, 300),
                                  (1256, 128 * 150 + 17)])
def test_gemm_tn_epilogues(cuda_dev, bn, M):
    """Every epilogue, on the auto tile, the 160-wide tile with one tile per CTA (the two epilogue groups take
    alternate fills + the register-stored tail) and with several tiles per CTA (groups alternate tiles)."""
    from seedstory import ops
    import functools
    torch.manual_seed(4)
    N, K = 640, 320
    ops_gemm = functools.partial(ops.gemm, force_bn=bn)
    a = torch.randn(M, K, device=cuda_dev).half()
    w = ops.register_const((torch.randn(N, K, device=cuda_dev) / math.sqrt(K)).half())   # a load-time weight
    assert ops.is_const_weight(w) and not ops.is_const_weight(a)
    bias = torch.randn(N, device=cuda_dev).half()
    res = torch.randn(M, N, device=cuda_dev).half()
    rpg = (M + 2) // 3
    b2 = torch.randn(3, N, device=cuda_dev).half()
    lin = (a.float() @ w.float().t() + bias.float()).half()
    c = ops_gemm(a, w, bias=bias, act=ops.ACT_GELU, residual=res)
    ref = torch.nn.functional.gelu(lin.float()).half().float() + res.float()
    _close(c, ref, what="bias+gelu+res")
    c = ops_gemm(a, w, bias=bias, bias2=b2, rows_per_group=rpg, act=ops.ACT_SILU)
    ref = torch.nn.functional.silu((lin.float() + b2.float().repeat_interleave(rpg, 0)[:M]).half().float())
    _close(c, ref, what="bias+bias2+silu")
    c = ops_gemm(a, w, alpha=0.125)
    _close(c, 0.125 * (a.float() @ w.float().t()), what="alpha")
    # GLU epilogues over interleaved column pairs
    wp = w.view(2, N // 2, K).permute(1, 0, 2).reshape(N, K).contiguous()          # rows (first_j, second_j)
    torch.cuda.synchronize()
    ops.register_const(wp)
    bp = bias.view(2, N // 2).t().reshape(N).contiguous()
    first, second = lin[:, : N // 2], lin[:, N // 2:]
    c = ops_gemm(a, wp, bias=bp, glu=ops.GLU_GEGLU)
    _close(c, first.float() * torch.nn.functional.gelu(second.float()).half().float(), what="geglu")
    c = ops_gemm(a, wp, bias=bp, glu=ops.GLU_SWIGLU)
    _close(c, torch.nn.functional.silu(first.float()).half().float() * second.float(), what="swiglu")


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conv3x3(cuda_dev, dt):
    from seedstory import ops
    torch.manual_seed(5)
    tol = 3e-3 if dt == torch.float16 else 2e-2
    for (Nimg, H, W, Cin, Cout) in [(2, 32, 32, 64, 64), (1, 64, 64, 320, 640), (2, 128, 128, 128, 320),
                                    (1, 256, 256, 64, 128)]:
        x = torch.randn(Nimg, Cin, H, W, device=cuda_dev).to(dt)
        w = (torch.randn(Cout, Cin, 3, 3, device=cuda_dev) / math.sqrt(9 * Cin)).to(dt)
        bias = torch.randn(Cout, device=cuda_dev).to(dt)
        temb = torch.randn(Nimg, Cout, device=cuda_dev).to(dt)
        res = torch.randn(Nimg, H, W, Cout, device=cuda_dev).to(dt)
        ref = torch.nn.functional.conv2d(x.float(), w.float(), bias.float(), padding=1)
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        w_p = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
        y = ops.conv3x3(x_nhwc, w_p, bias=bias)
        _close(y.permute(0, 3, 1, 2), ref, rtol=tol, atol=tol * ref.abs().max().item(), what=f"conv {H}x{W} {Cin}->{Cout}")
        if Cout % 160 == 0:
            y = ops.conv3x3(x_nhwc, w_p, bias=bias, force_bn=256)
            _close(y.permute(0, 3, 1, 2), ref, rtol=tol, atol=tol * ref.abs().max().item(), what="conv bn=256")
            y = ops.conv3x3(x_nhwc, w_p, bias=bias, force_bn=160)
            _close(y.permute(0, 3, 1, 2), ref, rtol=tol, atol=tol * ref.abs().max().item(), what="conv bn=160")
            y = ops.conv3x3(x_nhwc, w_p, bias=bias, force_bn=1256)
            _close(y.permute(0, 3, 1, 2), ref, rtol=tol, atol=tol * ref.abs().max().item(), what="conv pair")
        y = ops.conv3x3(x_nhwc, w_p, bias=bias, bias2=temb, residual=res)
        ref2 = (ref.to(dt).float() + temb.float()[:, :, None, None]).to(dt).float() + res.permute(0, 3, 1, 2).float()
        _close(y.permute(0, 3, 1, 2), ref2, rtol=tol, atol=tol * ref2.abs().max().item(), what="conv+temb+res")


def _ref_attn(q, k, v, scale, causal):
    # q [B,Lq,H,D] etc, fp32 math
    q, k, v = q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)
    s = q @ k.transpose(-1, -2) * scale
    if causal:
        tq, tk = q.shape[-2], k.shape[-2]
        qi = torch.arange(tq, device=q.device)[:, None]
        kj = torch.arange(tk, device=q.device)[None, :]
        s = s.masked_fill(~(kj <= qi + (tk - tq)), float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2)


@pytest.mark.parametrize("D", [64, 128])
def test_fmha_dense(cuda_dev, D):
    from seedstory import ops
    torch.manual_seed(6)
    for (B, H, Lq, Lk, causal) in [(2, 4, 64, 64, False), (1, 3, 100, 333, False), (2, 2, 257, 257, True),
                                   (1, 2, 66, 1107, True), (1, 16, 1024, 1024, False), (2, 2, 64, 320, False),
                                   (1, 2, 128, 100, False), (2, 3, 200, 128, False), (1, 2, 130, 130, True)]:  # one / two key tiles
        # fused qkv buffer for self-attention shapes, separate buffers otherwise
        q = torch.randn(B, Lq, H, D, device=cuda_dev).half()
        k = torch.randn(B, Lk, H, D, device=cuda_dev).half()
        v = torch.randn(B, Lk, H, D, device=cuda_dev).half()
        scale = 1.0 / math.sqrt(D)
        out = ops.mha_packed(q.view(B, Lq, H * D), k.view(B, Lk, H * D), v.view(B, Lk, H * D), H, scale, causal)
        ref = _ref_attn(q, k, v, scale, causal)
        _close(out.view(B, Lq, H, D), ref, rtol=3e-3, atol=3e-3, what=f"fmha D={D} {B},{H},{Lq},{Lk},{causal}")
    # strided views into one fused [B, L, 3*H*D] buffer
    B, H, L = 2, 4, 200
    qkv = torch.randn(B, L, 3 * H * D, device=cuda_dev).half()
    q, k, v = qkv[..., :H * D], qkv[..., H * D:2 * H * D], qkv[..., 2 * H * D:]
    out = ops.mha_packed(q, k, v, H, 0.1)
    ref = _ref_attn(q.reshape(B, L, H, D), k.reshape(B, L, H, D), v.reshape(B, L, H, D), 0.1, False)
    _close(out.view(B, L, H, D), ref, rtol=3e-3, atol=3e-3, what="fmha fused-qkv views")


@pytest.mark.parametrize("Lq,Lk", [(66, 1107), (1041, 1041), (128, 128), (200, 519), (40, 300)])
def test_fmha_paged_causal(cuda_dev, Lq, Lk):
    """The Llama prefill (Lq == Lk), the 66-token image-run chunk on a long cache, and odd page counts: paged K/V with
    the bottom-right causal mask.  Lq >= 64 runs on the tcgen05 kernel (pages through 3-D TMA maps; a 128-key tile = two
    pages), shorter chunks on the mma.sync kernel; unused pool pages hold NaN so that any read past the table shows."""
    from seedstory import ops
    torch.manual_seed(7)
    H, D = 32, 128
    max_pages = 24
    npg = (Lk + 63) // 64
    pt = torch.randperm(40, device=cuda_dev)[:max_pages].int().view(1, max_pages).contiguous()
    pool = torch.full((40, H, 64, D), float("nan"), device=cuda_dev, dtype=torch.float16)
    vpool = torch.full_like(pool, float("nan"))
    used = pt[0, :npg].long()
    pool[used] = torch.randn(npg, H, 64, D, device=cuda_dev).half()
    vpool[used] = torch.randn(npg, H, 64, D, device=cuda_dev).half()
    q = torch.randn(1, Lq, H * D, device=cuda_dev).half()
    out = torch.empty_like(q)
    ops.fmha_path_counts(reset=True)
    ops.fmha(q, pool, vpool, out, 1, H, Lq, Lk, D, (0, H * D, D), (0, 0, 0), (0, 0, 0), (0, H * D, D),
             1.0 / math.sqrt(D), causal=True, page_table=pt)
    torch.cuda.synchronize()
    tc, mma = ops.fmha_path_counts()
    assert (tc, mma) == ((1, 0) if Lq >= 64 else (0, 1)), (tc, mma)
    k = pool[used].permute(0, 2, 1, 3).reshape(-1, H, D)[:Lk][None]
    v = vpool[used].permute(0, 2, 1, 3).reshape(-1, H, D)[:Lk][None]
    ref = _ref_attn(q.view(1, Lq, H, D), k, v, 1.0 / math.sqrt(D), True)
    assert torch.isfinite(out).all()
    _close(out.view(1, Lq, H, D), ref, rtol=3e-3, atol=3e-3, what=f"fmha paged causal Lq={Lq} Lk={Lk}")


def test_groupnorm_and_glue(cuda_dev):
    from seedstory import ops
    torch.manual_seed(8)
    for dt, tol in ((torch.float16, 4e-3), (torch.bfloat16, 3e-2)):
        for (N, H, W, C) in [(2, 32, 32, 320), (1, 64, 64, 640), (2, 16, 16, 2560), (1, 128, 128, 128)]:
            x = (torch.randn(N, H, W, C, device=cuda_dev) * 2 + 0.5).to(dt)
            g = (1 + 0.1 * torch.randn(C, device=cuda_dev)).to(dt)
            b = (0.1 * torch.randn(C, device=cuda_dev)).to(dt)
            ws = ops.groupnorm_ws(N, H * W, C, 32, cuda_dev)
            for silu in (False, True):
                y = ops.groupnorm_nhwc(x, g, b, 32, 1e-5, silu, ws)
                ref = torch.nn.functional.group_norm(x.float().permute(0, 3, 1, 2), 32, g.float(), b.float(), 1e-5)
                if silu:
                    ref = torch.nn.functional.silu(ref.to(dt).float())
                _close(y.permute(0, 3, 1, 2), ref, rtol=tol, atol=tol, what=f"groupnorm {N},{H},{W},{C} silu={silu} {dt}")
    x = torch.randn(2, 8, 8, 64, device=cuda_dev).half()
    up = ops.upsample2x(x)
    ref = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2, mode="nearest")
    assert torch.equal(up.permute(0, 3, 1, 2).float(), ref)
    a = torch.randn(2, 8, 8, 64, device=cuda_dev).half()
    b = torch.randn(2, 8, 8, 128, device=cuda_dev).half()
    assert torch.equal(ops.concat_channels(a, b), torch.cat([a, b], -1))
    # stride-2 conv through im2col + GEMM
    Cin, Cout = 64, 128
    xi = torch.randn(2, Cin, 32, 32, device=cuda_dev).half()
    w = (torch.randn(Cout, Cin, 3, 3, device=cuda_dev) / math.sqrt(9 * Cin)).half()
    bias = torch.randn(Cout, device=cuda_dev).half()
    cols = ops.im2col3x3_s2(xi.permute(0, 2, 3, 1).contiguous())
    y = ops.gemm(cols, w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous(), bias=bias)
    ref = torch.nn.functional.conv2d(xi.float(), w.float(), bias.float(), stride=2, padding=1)
    _close(y.view(2, 16, 16, Cout).permute(0, 3, 1, 2), ref, rtol=3e-3, atol=3e-3 * ref.abs().max().item(), what="conv s2")
    # patchify
    img = torch.randn(2, 3, 56, 56, device=cuda_dev).half()
    pat = ops.im2col_patch(img, 14, 640)
    ref = torch.nn.functional.unfold(img.float(), 14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(pat[:, :588].float(), ref) and pat[:, 588:].abs().max().item() == 0
    # softmax rows / transpose / mean / l2norm / scatter / add_bcast / cast / uint8
    s = torch.randn(37, 1000, device=cuda_dev).half()
    ref = torch.softmax(s.float() * 0.3, -1)
    _close(ops.softmax_rows_(s.clone(), 0.3), ref, rtol=2e-3, atol=1e-5, what="softmax rows")
    t = torch.randn(70, 130, device=cuda_dev).half()
    assert torch.equal(ops.transpose2d(t), t.t().contiguous())
    m = torch.randn(2, 64, 256, device=cuda_dev).half()
    _close(ops.mean_tokens(m), m.float().mean(1), rtol=2e-3, atol=1e-3, what="mean tokens")
    _close(ops.l2norm_tokens(m), torch.nn.functional.normalize(m.float(), dim=1), rtol=3e-3, atol=1e-4, what="l2norm")
    dst = torch.zeros(50, 128, device=cuda_dev).half()
    src = torch.randn(5, 128, device=cuda_dev).half()
    rows = torch.tensor([3, 9, 10, 49, 0], device=cuda_dev, dtype=torch.int32)
    ops.scatter_rows(src, rows, dst)
    assert torch.equal(dst[rows.long()], src)
    xx = torch.randn(6, 4, 64, device=cuda_dev).half()
    pe = torch.randn(4, 64, device=cuda_dev).half()
    assert torch.equal(ops.add_bcast(xx, pe), xx + pe)
    assert torch.equal(ops.cast_scale(xx, torch.bfloat16, 2.0), (xx.float() * 2).bfloat16())
    im = torch.randn(100, 8, device=cuda_dev).half()
    u8 = ops.image_to_uint8(im, 3)
    ref = ((im[:, :3].float() / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
    assert (u8.int() - ref.int()).abs().max().item() <= 1


def test_cfg_euler_step(cuda_dev):
    from seedstory import ops
    torch.manual_seed(9)
    HW, C = 128 * 128, 4
    eps = torch.randn(2, HW, 8, device=cuda_dev).half()
    lat = torch.randn(HW, C, device=cuda_dev).half()
    nxt = torch.zeros(2, HW, 64, device=cuda_dev).half()
    sigma, sigma_next, g = 14.6, 11.2, 7.5
    eu, ec = eps[0, :, :C], eps[1, :, :C]
    e = eu + (g * (ec - eu))
    x = lat.float()
    pred = x - sigma * e.float()
    ref = (x + (x - pred) / sigma * (sigma_next - sigma)).half()
    ops.cfg_euler_step(eps, lat, nxt, C, g, sigma, sigma_next)
    _close(lat, ref, rtol=2e-3, atol=2e-2, what="euler")
    _close(nxt[1, :, :C], ref.float() / math.sqrt(sigma_next ** 2 + 1), rtol=2e-3, atol=2e-3, what="scaled input")
    assert nxt[:, :, C:].abs().max().item() == 0


@pytest.mark.parametrize("M,C", [(2048, 1280), (8192, 640), (300, 320)])
def test_gemm_with_folded_layernorm(cuda_dev, M, C):
    """LayerNorm -> Linear as one GEMM: the producer GEMM leaves per-row (sum, sum of squares) partials, the consumer
    applies mean / rstd / gamma / beta in its epilogue (ss_gemm_tn_ln).  Checked against layernorm -> gemm in fp32."""
    from seedstory import ops
    torch.manual_seed(M + C)
    a = torch.randn(M, C, device=cuda_dev).half()
    w0 = (torch.randn(C, C, device=cuda_dev) * 0.03).half()
    b0 = (torch.randn(C, device=cuda_dev) * 0.1).half()
    res = (torch.randn(M, C, device=cuda_dev) * 2 + 0.7).half()      # rows with a non-zero mean
    st = ops.row_stats_buffer(M, C, cuda_dev)
    x = ops.gemm(a, w0, bias=b0, residual=res, stats_out=st)            # the LN input, with its row statistics
    x_plain = ops.gemm(a, w0, bias=b0, residual=res)
    assert torch.equal(x, x_plain), "stats_out must not change the GEMM's output"
    s = st.sum(0)
    xf = x.float()
    assert torch.allclose(s[:, 0], xf.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[:, 1], (xf * xf).sum(1), rtol=1e-4, atol=1e-2)
    gamma = (1.0 + 0.2 * torch.randn(C, device=cuda_dev)).half()
    beta = (0.1 * torch.randn(C, device=cuda_dev)).half()
    for N, glu, with_bias in [(C, 0, False), (3 * C, 0, False), (8 * C, ops.GLU_GEGLU, True)]:
        w = (torch.randn(N, C, device=cuda_dev) * 0.03).half()
        bias = (torch.randn(N, device=cuda_dev) * 0.1).half() if with_bias else None
        f = ops.FoldedLN(w, gamma, beta, 1e-5, bias=bias)
        y = ops.gemm(x, f.w, glu=glu, ln=f, ln_stats=st)
        ln = torch.nn.functional.layer_norm(xf, (C,), gamma.float(), beta.float(), 1e-5)
        ref = ln @ w.float().t() + (bias.float() if bias is not None else 0.0)
        if glu:
            val, gate = ref[:, 0::2], ref[:, 1::2]                      # interleaved (value_j, gate_j) columns
            ref = val * torch.nn.functional.gelu(gate)
        # same bound as LN (rounded to fp16) -> GEMM: 2e-3 of the output scale
        _close(y, ref, rtol=4e-3, atol=4e-3 * ref.abs().max().item(), what=f"folded LN gemm N={N} glu={glu}")
        y2 = ops.gemm(ops.layernorm(x, gamma, beta, 1e-5), w, bias=bias, glu=glu)
        _close(y, y2, rtol=4e-3, atol=4e-3 * ref.abs().max().item(), what=f"folded vs unfolded N={N} glu={glu}")
