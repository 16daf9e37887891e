#!/usr/bin/env python
"""bench.py — story-turns/sec of the SEED-Story interleaved inference hot path on B200 (contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A *step* is one 10-turn synthetic StoryStream-shaped story per GPU (BASELINE.json configs[1]): start image 448x448,
64-token caption, each turn = 64 greedy text tokens + forced <img>, 64 image queries, </img>, EOS, then one SDXL
1024x1024 image (Euler, 50 steps, CFG 7.5, seed 42) decoded by the VAE; window of 8 images.  Stories are
independent, so N GPUs run N stories data-parallel with no data-path collective ("scaling": "weak").

  value : turns/s with the start image + caption already on the device, results left on the device
  e2e   : same turns through the reference-facing API (src.* drop-ins) from HOST buffers: pinned start image and
          caption copied H2D inside the timed region, every turn's token ids and 1024x1024 uint8 image copied D2H
  roofline     : the dominant kernel (tcgen05 GEMM / implicit-GEMM conv inside the UNet): every launch type of one UNet
                 step timed live as 10 back-to-back launches in a CUDA graph (CUDA events), weighted by its count
  cpu_baseline : the oracle (CPU restatement of the reference path) on a bounded sample, host cores of this box
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "seed-story_b200"), os.path.join(ROOT, "seed-story_b200", "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

TURNS = 10
DENOISE_STEPS = 50
METRIC = "story-turns/sec (text+image) at 10-turn seq"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.stop_flag, self.index = [], False, index
        self.t = None

    def start(self):
        def run():
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            while not self.stop_flag:
                try:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
                except Exception:
                    pass
                time.sleep(0.5)
        self.t = threading.Thread(target=run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=reasons, samples=len(sm))


def synthetic_story(s, vocab_text=32000):
    """seed 1000+s: uniform-random uint8 448x448x3 image through the CLIP transform, 64 caption ids in [3, 32000)."""
    g = torch.Generator().manual_seed(1000 + s)
    img = torch.randint(0, 256, (3, 448, 448), generator=g, dtype=torch.uint8).float() / 255.0
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(3, 1, 1)
    img = ((img - mean) / std).half().unsqueeze(0)
    cap = torch.randint(3, vocab_text, (64,), generator=g).tolist()
    return img, cap


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from seedstory import _capi, ops, story
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    ops.require_device()
    pipe = story.StoryPipeline(device=dev, cfg=story.FULL, num_inference_steps=args.denoise_steps, verbose=(rank == 0))
    turns = args.turns

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(n):
        if world == 1:
            return n
        t = torch.tensor([n], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    # ---- value: device-resident inputs -----------------------------------------------------------
    stories = [synthetic_story(rank * 1000 + i) for i in range(args.warmup + args.steps)]
    dev_inputs = [(im.to(dev), cap) for im, cap in stories]
    for i in range(args.warmup):
        pipe.run_story(dev_inputs[i][0], dev_inputs[i][1], turns)
    barrier()
    n_turns = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c0 = _capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        outs = pipe.run_story(dev_inputs[i][0], dev_inputs[i][1], turns)
        n_turns += sum(1 for o in outs if o["has_img_output"])   # a story ends early if a turn emits no image
    e1.record()
    barrier()
    launches = _capi.launch_count() - c0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    n_turns_all = sum_over_ranks(n_turns)
    value = n_turns_all / (ms * 1e-3)

    # ---- e2e: host buffers, H2D + D2H inside the timed region --------------------------------------
    pinned = [(im.pin_memory(), cap) for im, cap in stories[args.warmup:]]
    host_img = torch.empty((1024, 1024, 3), dtype=torch.uint8).pin_memory()
    h2d = d2h = 0

    n_turns_e2e = 0

    def e2e_story(im_host, cap):
        nonlocal h2d, d2h, n_turns_e2e
        im = im_host.to(dev, non_blocking=True)
        cap_dev = torch.tensor(cap, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)  # ids travel too
        h2d += im_host.numel() * 2 + cap_dev.numel() * 4
        outs = pipe.run_story(im, cap, turns, return_images=True)
        for o in outs:
            if not o["has_img_output"]:
                continue
            n_turns_e2e += 1
            host_img.copy_(o["image"], non_blocking=True)
            d2h += host_img.numel() + len(o["generate_ids"]) * 8
        torch.cuda.current_stream().synchronize()
    e2e_story(*pinned[0])  # warm the path (pinned allocations)
    h2d = d2h = n_turns_e2e = 0
    barrier()
    e0.record()
    for im_host, cap in pinned:
        e2e_story(im_host, cap)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = sum_over_ranks(n_turns_e2e) / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM/conv launches of one UNet CFG step), live CUDA events ----
    peaks = load_peaks()
    roof = None
    cpu_base = None
    if rank == 0:
        ue = pipe.unet.engine()
        # one eager UNet forward records every tcgen05 GEMM / conv launch together with a closure that re-issues it on
        # the same buffers; each distinct launch type is then timed as 10 back-to-back launches inside a CUDA graph
        # (CUDA events on the launching stream, no host gaps, programmatic-dependent-launch overlap as in the real
        # step), and the per-launch times are weighted by how often the type occurs in the step
        ops.RECORD = []
        torch.cuda.synchronize()
        ue.forward()
        torch.cuda.synchronize()
        rec = ops.RECORD
        ops.RECORD = None
        types = {}
        for name, fl, fn in rec:
            t = types.setdefault(name, [0, fl, fn])
            t[0] += 1
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        side = torch.cuda.Stream()
        tc_ms = 0.0
        per_type = []
        for name, (cnt, fl, fn) in types.items():
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(10):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            g0.record()
            for _ in range(3):
                gr.replay()
            g1.record()
            torch.cuda.synchronize()
            us = g0.elapsed_time(g1) / 30 * 1e3
            tc_ms += cnt * us * 1e-3
            per_type.append((cnt * us, name, cnt, round(us, 2), round(fl / us / 1e6, 1)))
            del gr
        per_type.sort(reverse=True)
        tc_fl = sum(f for (_, f, _) in rec)
        n_l = len(rec)
        g0.record()
        for _ in range(5):
            ue._graph.replay()
        g1.record()
        torch.cuda.synchronize()
        fwd_ms = g0.elapsed_time(g1) / 5
        ach = tc_fl / (tc_ms * 1e-3) / 1e12
        # the single most expensive launch type of the step (GEGLU projection 2048 x 10240 x 1280), timed alone
        a_ = torch.randn(2048, 1280, device=dev).half()
        w_ = (torch.randn(10240, 1280, device=dev) * 0.03).half()
        b_ = torch.randn(10240, device=dev).half()
        o_ = torch.empty(2048, 5120, device=dev).half()
        for _ in range(3):
            ops.gemm(a_, w_, bias=b_, glu=ops.GLU_GEGLU, out=o_)
        g0.record()
        for _ in range(20):
            ops.gemm(a_, w_, bias=b_, glu=ops.GLU_GEGLU, out=o_)
        g1.record()
        torch.cuda.synchronize()
        top_ms = g0.elapsed_time(g1) / 20
        top_fl = 2.0 * 2048 * 10240 * 1280
        roof = dict(bound="tensor", kernel="gemm_tc_pair_kernel / gemm_tc_persist_kernel (tcgen05 GEMM + implicit-GEMM conv3x3)",
                    achieved=round(ach, 1),
                    peak=peaks["tf_sustained"], unit="TFLOP/s", frac=round(ach / peaks["tf_sustained"], 4),
                    traffic=31.8e6, traffic_note="dram__bytes_read+write of the GEGLU launch below from ncu --set full "
                    "(profiles/r1_ncu_end_of_round_full.md); its algorithmic operand bytes are 31.5e6",
                    top_launch=dict(shape="GEGLU GEMM 2048x10240x1280 (60 launches per UNet step)", us=round(top_ms * 1e3, 1),
                                    achieved=round(top_fl / (top_ms * 1e-3) / 1e12, 1), peak=peaks["tf_burst"],
                                    frac=round(top_fl / (top_ms * 1e-3) / 1e12 / peaks["tf_burst"], 4)),
                    peak_source=peaks["src"] + ", sustained bf16 (kernel timed inside a long step)",
                    method="per launch type: 10 back-to-back launches in a CUDA graph, CUDA events; weighted by count",
                    top_types=[dict(launch=n, count=c, us=u, tflops=t) for (_, n, c, u, t) in per_type[:6]],
                    launches_per_unet_step=n_l, algorithmic_tflop_per_unet_step=round(tc_fl / 1e12, 3),
                    unet_step_ms=round(fwd_ms, 3), tc_share_of_unet_step=round(tc_ms / fwd_ms, 3))
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_reference_sample(args)

    if rank == 0:
        line = dict(metric=METRIC, value=round(value, 4), unit="story-turns/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=round(ms / args.steps, 2), higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f16", data="synthetic",
                    config=dict(workload="configs[1]: 10-turn interleaved story, batch=1 per GPU, fp16, 448^2 start image, "
                                         "64-token caption, 64 text tokens + 66-token image run per turn, SDXL 1024^2 "
                                         f"{args.denoise_steps} Euler steps CFG 7.5, window 8",
                                turns_per_step=turns, stories_per_gpu=1, parallelism=f"dp{world} (one story per GPU)",
                                timing="inputs (7B+ weights streamed per decode step, 2.6B UNet) exceed L2; no flush needed",
                                weights="seeded random, real shapes (no checkpoints offline)"),
                    e2e=dict(value=round(e2e_value, 4), unit="story-turns/s", h2d_bytes_per_step=h2d // args.steps,
                             d2h_bytes_per_step=d2h // args.steps),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu_base)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU restatement of the reference path) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(args):
    """Bounded sample of one story turn on the host cores (oracle = CPU restatement of the reference path, fp32):
      * Llama-2-7B decode: 2 of 32 decoder layers + lm_head, 2 tokens at context 256, scaled x16 layers
      * SDXL UNet: ONE batch-1 forward at HALF resolution (64x64 latents), scaled x4 (pixels) x2 (CFG batch); the
        attention terms grow faster than x4, so the extrapolation favours the CPU
    (prefill, ViT, resamplers and the fp32 VAE are left out, which also favours the CPU figure).
    Returns the cpu_baseline object; value is story-turns/s extrapolated from the sample."""
    from oracle import llama_oracle as LO
    from oracle import sdxl_oracle as SO
    from seedstory import synthetic
    cores = os.cpu_count() or 1
    threads = min(cores, 32)          # torch's CPU GEMMs stop scaling (and oversubscribe) beyond a few dozen threads
    torch.set_num_threads(threads)
    t_all = time.time()
    p = LO.LlamaParams.random(4096, 11008, 32, 2, 32066, lora_r=16, seed=1)
    ctx = 256
    emb = torch.randn(1, ctx, 4096) * 0.02
    with torch.no_grad():
        _, _, kv = LO.model_forward(p, emb, torch.arange(ctx).unsqueeze(0), None)
        t0 = time.time()
        for i in range(2):
            _, _, kv = LO.model_forward(p, emb[:, :1], torch.tensor([[ctx + i]]), kv)
        t_tok2 = (time.time() - t0) / 2
    t_token = t_tok2 * 16  # 2 -> 32 layers (lm_head counted 16x: small overestimate, noted)
    del p, kv
    cfg = dict(SO.SDXL_UNET_CONFIG)
    sd = synthetic.random_unet_state_dict(cfg, seed=1234)
    x = torch.randn(1, 4, 64, 64)
    with torch.no_grad():
        t0 = time.time()
        SO.unet_forward(sd, cfg, x, torch.tensor([981.0]), torch.randn(1, 64, 2048), torch.randn(1, 1280),
                        torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]]))
        t_half = time.time() - t0
    t_unet = 8.0 * t_half
    tokens_per_turn = 131
    t_turn = tokens_per_turn * t_token + args.denoise_steps * t_unet
    return dict(value=round(1.0 / t_turn, 6), unit="story-turns/s", cores=threads, kind="port",
                sample=f"oracle fp32 on {threads} threads ({cores} cores present): 2 decode tokens x 2/32 Llama layers "
                       f"(+lm_head) at ctx 256 ({t_tok2 * 1e3:.0f} ms/token/2 layers) and one batch-1 UNet forward at "
                       f"half resolution ({t_half:.1f} s, x8 for resolution and CFG) out of {args.denoise_steps} steps, "
                       f"extrapolated to a turn of {tokens_per_turn} decoded tokens + {args.denoise_steps} CFG steps; "
                       f"prefill/ViT/VAE omitted (favours CPU); sample wall {time.time() - t_all:.0f} s")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps_vals = []
    base = None
    for _ in range(args.warmup if args.warmup < 1 else 1):
        pass
    t0 = time.time()
    base = cpu_reference_sample(args)
    v = base["value"]
    line = dict(metric=METRIC, value=v, unit="story-turns/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=round(1e3 * TURNS / v, 1), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="configs[1] (same as our arm), CPU fp32 oracle restatement of the reference path; each "
                                     "step is a bounded sample extrapolated to a 10-turn story", turns_per_step=TURNS),
                cpu_baseline=base,
                e2e=dict(value=v, unit="story-turns/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note=f"reference stack (transformers 4.34 / diffusers / peft / xformers) is not installable offline; "
                     f"sample took {time.time() - t0:.0f} s")
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--turns", type=int, default=TURNS)
    ap.add_argument("--denoise-steps", type=int, default=DENOISE_STEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
