"""SDXL de-tokenizer parity on the GPU: UNet / sampler / VAE engines (CUDA, C-ABI) vs the CPU oracle restatement
(oracle/sdxl_oracle.py — parity UNPINNED against diffusers, see its header) on shared seeded weights, reduced dims."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

TINY_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 256), layers_per_block=2,
                 transformer_layers_per_block=(0, 1, 2), num_attention_heads=(1, 2, 4), cross_attention_dim=256,
                 addition_time_embed_dim=32, projection_class_embeddings_input_dim=160 + 6 * 32, norm_num_groups=32,
                 sample_size=64)
TINY_VAE = dict(latent_channels=4, out_channels=3, block_out_channels=(64, 64, 128, 128), layers_per_block=2,
                norm_num_groups=32, scaling_factor=0.13025)


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def test_unet_forward_and_sampler_vs_oracle(cuda_dev):
    from oracle import sdxl_oracle as SO
    from seedstory import sdxl_engine, synthetic
    cfg = TINY_UNET
    sd = synthetic.random_unet_state_dict(cfg, seed=5)
    sd16 = {k: v.half().float() for k, v in sd.items()}       # oracle sees the same fp16-rounded weights, fp32 math
    eng = sdxl_engine.UNetEngine(sd, cfg, cuda_dev)
    g = torch.Generator().manual_seed(1)
    S = cfg["sample_size"]
    ctx = (torch.randn(2, 64, 256, generator=g) * 0.5).half()
    pooled = (torch.randn(2, 160, generator=g) * 0.5).half()
    steps = 4
    ts, sig = sdxl_engine.euler_schedule(steps)
    tid = [[8 * S, 8 * S, 0, 0, 8 * S, 8 * S]] * 2
    eng.set_conditioning(ctx, pooled, tid, ts)
    # single forward at step 0
    x = (torch.randn(1, 4, S, S, generator=g)).half()
    eng.x_in.zero_()
    eng.x_in[:, :, :, :4] = x.permute(0, 2, 3, 1).to(cuda_dev)
    eng.temb_cur.copy_(eng.temb_all[0:2])
    eps = eng.forward()[..., :4].permute(0, 3, 1, 2)
    ref = SO.unet_forward(sd16, cfg, x.float().repeat(2, 1, 1, 1), torch.tensor([ts[0], ts[0]]), ctx.float(),
                          pooled.float(), torch.tensor(tid, dtype=torch.float32))
    r = _rel(eps, ref)
    assert r < 2e-2, f"UNet forward rel err {r}"
    # full sampler (CUDA graph path) vs oracle loop
    lat0 = torch.randn(1, 4, S, S, generator=g).half()
    lat = eng.sample(lat0.to(cuda_dev), steps, guidance=7.5, use_graph=True)
    ref_lat = SO.sdxl_sample(sd16, cfg, ctx[1:].float(), ctx[:1].float(), pooled[1:].float(), pooled[:1].float(),
                             lat0.float(), steps, 7.5, size=(8 * S, 8 * S))
    got = lat.view(S, S, 4).permute(2, 0, 1)[None]
    r = _rel(got, ref_lat)
    assert r < 3e-2, f"sampler rel err {r}"
    assert eng.total_launches > 0


def test_vae_decode_vs_oracle(cuda_dev):
    from oracle import sdxl_oracle as SO
    from seedstory import sdxl_engine, synthetic
    cfg = TINY_VAE
    sd = synthetic.random_vae_decoder_state_dict(cfg, seed=6)
    sdb = {k: v.bfloat16().float() for k, v in sd.items()}
    eng = sdxl_engine.VAEDecoderEngine(sd, cfg, cuda_dev)
    S = 32
    lat = (torch.randn(1, 4, S, S, generator=torch.Generator().manual_seed(2)) * 0.2).half()
    u8, img = eng.decode(lat.permute(0, 2, 3, 1).reshape(S * S, 4).contiguous().to(cuda_dev), S)
    ref = SO.vae_decode(sdb, cfg, lat.float())
    got = img[0, :, :, :3].permute(2, 0, 1)[None]
    r = _rel(got, ref)
    assert r < 5e-2, f"VAE decode (bf16 vs fp32 oracle) rel err {r}"
    ref_u8 = SO.postprocess(ref)[0]
    mse = ((u8.cpu().float() - ref_u8.float()) ** 2).mean().item()
    psnr = 10 * math.log10(255.0 ** 2 / max(mse, 1e-9))
    assert psnr > 35.0, f"decoded image PSNR vs fp32 oracle {psnr:.1f} dB"


def test_story_pipeline_tiny_end_to_end(cuda_dev):
    """Whole hot path through the src.* drop-ins at reduced dims: 3 turns, window 2 (eviction exercised)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "seed-story_b200", "shims"))
    from seedstory import story
    pipe = story.StoryPipeline(device=cuda_dev, cfg=story.TINY, num_inference_steps=3, n_text_tokens=5, window_size=2)
    img = torch.randn(1, 3, 56, 56, device=cuda_dev).half()
    cap = [5, 9, 17, 33]
    outs = pipe.run_story(img, cap, n_turns=3, return_images=True)
    tk = pipe.tokenizer
    for o in outs:
        g = o["generate_ids"]
        assert g[5] == tk.boi and g[6:70] == [tk.img0 + i for i in range(64)] and g[70] == tk.eoi and g[71] == tk.eos_token_id
        assert o["image"].shape == (512, 512, 3) and o["image"].dtype == torch.uint8
    # determinism: same inputs -> identical tokens and pixels
    outs2 = pipe.run_story(img, cap, n_turns=3, return_images=True)
    assert [o["generate_ids"] for o in outs] == [o["generate_ids"] for o in outs2]
    assert all(torch.equal(a["image"], b["image"]) for a, b in zip(outs, outs2))
    # the side-stream overlap of SDXL(t) with the MLLM of turn t+1 must not change any result
    outs3 = pipe.run_story(img, cap, n_turns=3, return_images=True, overlap=True)
    assert [o["generate_ids"] for o in outs] == [o["generate_ids"] for o in outs3]
    assert all(torch.equal(a["image"], b["image"]) for a, b in zip(outs, outs3))


def test_many_synthetic_stories_never_lose_the_image_run(cuda_dev):
    """Regression for the round-1 bench crash: with random weights a free greedy token can be EOS (or <img>); the
    synthetic schedule suppresses both in the free slots, so every turn of every story has the forced structure.
    30 stories x 3 turns at reduced dims (the bench seeds, 1000+s)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "seed-story_b200", "shims"))
    from seedstory import story
    pipe = story.StoryPipeline(device=cuda_dev, cfg=story.TINY, num_inference_steps=1, n_text_tokens=24, window_size=2)
    tk = pipe.tokenizer
    for s in range(30):
        g = torch.Generator().manual_seed(1000 + s)
        img = torch.randn(1, 3, 56, 56, generator=g).half().to(cuda_dev)
        cap = torch.randint(3, 254, (16,), generator=g).tolist()
        outs = pipe.run_story(img, cap, n_turns=3, decode_images=False)
        assert len(outs) == 3 and all(o["has_img_output"] for o in outs), f"story {s}"
        for o in outs:
            gi = o["generate_ids"]
            assert tk.eos_token_id not in gi[:24] and tk.boi not in gi[:24]
            assert gi[24] == tk.boi and gi[89] == tk.eoi and gi[90] == tk.eos_token_id and len(gi) == 91


def test_story_ends_like_the_reference_when_a_turn_has_no_image(cuda_dev):
    """gen_george.py:208 — `while output['has_img_output'] and …`: a turn whose generation stops before an image run
    ends the story (no assert, no crash).  Forced here by scheduling EOS as the first generated token."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "seed-story_b200", "shims"))
    from seedstory import story
    from src.models_clm.generation import ForcedScheduleProcessor
    pipe = story.StoryPipeline(device=cuda_dev, cfg=story.TINY, num_inference_steps=1, n_text_tokens=4, window_size=2)
    pipe._schedule = lambda: [ForcedScheduleProcessor([-1, -1, pipe.tokenizer.eos_token_id])]
    img = torch.randn(1, 3, 56, 56, device=cuda_dev).half()
    outs = pipe.run_story(img, [5, 9, 17], n_turns=3, decode_images=False)
    assert len(outs) == 1 and outs[0]["has_img_output"] is False and outs[0]["generate_ids"][-1] == pipe.tokenizer.eos_token_id


def test_batched_stories_equal_single_stories(cuda_dev):
    """StoryPipeline.run_stories (BASELINE configs[3]: several stories per GPU, MLLM decode steps batched over the paged
    KV cache) gives every story the tokens and pixels run_story() gives it alone — windowed re-prefill and live sink."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "seed-story_b200", "shims"))
    from seedstory import story
    pipe = story.StoryPipeline(device=cuda_dev, cfg=story.TINY, num_inference_steps=2, n_text_tokens=5, window_size=2)
    g = torch.Generator().manual_seed(21)
    imgs = [torch.randn(1, 3, 56, 56, generator=g).half().to(cuda_dev) for _ in range(3)]
    caps = [[5, 9, 17, 33], [7, 7, 21, 40, 41, 42], [100, 3]]
    for sink in (False, True):
        batched = pipe.run_stories(imgs, caps, n_turns=4, return_images=True, sink=sink)
        for i in range(3):
            alone = pipe.run_story(imgs[i], caps[i], n_turns=4, return_images=True, sink=sink)
            assert [o["generate_ids"] for o in batched[i]] == [o["generate_ids"] for o in alone], (sink, i)
            assert all(torch.equal(a["image"], b["image"]) for a, b in zip(batched[i], alone)), (sink, i)
