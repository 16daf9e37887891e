"""Run-to-run bit reproducibility of the hot-path components (development aid): each component is run repeatedly on the
same inputs and every output is compared bitwise with the first."""
import math
import sys

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims", "."]
from seedstory import ops, sdxl_engine, story, synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def check(name, fn, reps=20):
    ref = None
    bad = 0
    for i in range(reps):
        out = fn()
        torch.cuda.synchronize()
        out = [o.clone() for o in (out if isinstance(out, (list, tuple)) else [out])]
        if ref is None:
            ref = out
        elif not all(torch.equal(a, b) for a, b in zip(ref, out)):
            bad += 1
    print(f"{name:60s} {'OK' if bad == 0 else f'NONDETERMINISTIC in {bad}/{reps - 1} repeats'}", flush=True)


# GroupNorm
for (N, HW, C) in [(2, 1024, 1280), (2, 4096, 640), (2, 16384, 320), (2, 1024, 128), (2, 256, 256)]:
    x = torch.randn(N, HW, C, device=dev).half()
    g, b = torch.randn(C, device=dev).half(), torch.randn(C, device=dev).half()
    ws = ops.groupnorm_ws(N, HW, C, 32, dev)
    check(f"groupnorm {N}x{HW}x{C}", lambda: ops.groupnorm_nhwc(x.view(N, HW, 1, C), g, b, 32, 1e-5, True, ws))
# GEMM with row statistics / folded LN
for (M, C) in [(2048, 1280), (8192, 640), (2048, 128), (512, 256)]:
    a = torch.randn(M, C, device=dev).half()
    w = (torch.randn(C, C, device=dev) * 0.03).half()
    res = torch.randn(M, C, device=dev).half()
    st = ops.row_stats_buffer(M, C, dev)
    check(f"gemm +res +stats {M}x{C}", lambda: (ops.gemm(a, w, residual=res, stats_out=st), st))
    gamma, beta = torch.randn(C, device=dev).half(), torch.randn(C, device=dev).half()
    f = ops.FoldedLN((torch.randn(3 * C, C, device=dev) * 0.03).half(), gamma, beta, 1e-5)
    xx = ops.gemm(a, w, residual=res, stats_out=st)
    check(f"gemm ln {M}x{3 * C}x{C}", lambda: ops.gemm(xx, f.w, ln=f, ln_stats=st))
# UNet forward, tiny and full width, LayerNorm folded / separate
for name, cfg in [("tiny", dict(synthetic.SDXL_UNET_CONFIG, **story.TINY["unet"])), ("full", synthetic.SDXL_UNET_CONFIG)]:
    sd = synthetic.random_unet_state_dict(cfg, seed=0, dtype=torch.float16, device="cuda")
    for fold in (False, True):
        eng = sdxl_engine.UNetEngine(sd, cfg, dev, fold_ln=fold)
        ts, _ = sdxl_engine.euler_schedule(4)
        ctx = torch.randn(2, 64, cfg["cross_attention_dim"], device=dev).half()
        S = cfg["sample_size"]
        pooled = torch.randn(2, cfg["projection_class_embeddings_input_dim"] - 6 * cfg["addition_time_embed_dim"], device=dev).half()
        eng.set_conditioning(ctx, pooled, [[8 * S, 8 * S, 0, 0, 8 * S, 8 * S]] * 2, ts)
        eng.x_in.normal_()
        eng.temb_cur.copy_(eng.temb_all[0:2])
        check(f"UNet forward {name} fold_ln={fold} (eager)", lambda: eng.forward(), reps=8)
        lat = torch.randn(1, 4, S, S, device=dev).half()
        check(f"UNet sample 3 steps {name} fold_ln={fold} (graph)", lambda: eng.sample(lat, 3).clone(), reps=6)
        del eng
    del sd
# decode attention
H, D, B = 32, 128, 2
kc = torch.randn(80, H, 64, D, device=dev).half()
vc = torch.randn(80, H, 64, D, device=dev).half()
pt = torch.randperm(80, device=dev).int().view(B, 40).contiguous()
q = torch.randn(B, H * D, device=dev).half()
lens = torch.tensor([1041, 77], device=dev, dtype=torch.int32)
for splits in (1, 12, 32):
    ws = ops.attn_decode_workspace(B, H, D, splits, dev)
    out = torch.empty_like(q)
    check(f"attn_decode_paged splits={splits}", lambda: (ops.attn_decode_paged(q, kc, vc, lens, pt, out, ws, H, D, splits, 1 / math.sqrt(D)), out)[1], reps=50)
print("done")
