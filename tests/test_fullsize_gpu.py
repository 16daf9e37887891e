"""Full-dimension model-level parity (the configs BASELINE.json names), CUDA path vs the fp32 CPU oracle on shared
seeded weights.  Depth is reduced where the CPU oracle would take minutes (1 Llama layer, a 3-level UNet with one
ResnetBlock / one transformer layer per level), never width, sequence length or resolution: every kernel sees the
exact shapes of the real model (ViT 1664-wide / head_dim 104 padded to 128 / 1024 tokens; Llama 4096 / 32x128 /
11008 / vocab 32066 at context 1041; UNet 320@128^2, 640@64^2, 1280@32^2 with 2048-wide 64-token context)."""
import math

import pytest
import torch

from _parity import rel

pytestmark = pytest.mark.gpu


def test_vit_g_full_size_matches_oracle_and_runs_on_tcgen05_attention(cuda_dev):
    """One 448x448 image through the full 48-layer ViT-G + attention pool (reference qwen_visual.py:376-399) via the
    src.* drop-in; the attention of every layer must have been served by the tcgen05 kernel (fmha_tc_kernel<128>)."""
    from oracle import vision_oracle as VO
    from seedstory import ops, story
    from src.models.qwen_visual import VisionTransformerWithAttnPool
    cfg = story.FULL["vit"]
    torch.manual_seed(11)
    with torch.device(cuda_dev):
        m = VisionTransformerWithAttnPool(**cfg).eval()
    m = m.to(dtype=torch.float16)
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (1, 3, 448, 448), generator=g).float() / 255.0
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    img = ((img - mean) / std).half()
    ops.fmha_path_counts(reset=True)
    out = m(img.to(cuda_dev))
    torch.cuda.synchronize()
    tc, mma = ops.fmha_path_counts()
    assert tc >= cfg["layers"] + 1 and mma == 0, f"ViT attention must run on the tcgen05 kernels (tc={tc}, mma.sync={mma})"
    sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}   # same fp16-rounded weights, fp32 math
    torch.set_num_threads(min(32, torch.get_num_threads() or 8))
    with torch.no_grad():
        ref = VO.vit_forward(sd, img.float(), cfg["heads"], cfg["layers"], cfg["patch_size"])
    assert out.shape == ref.shape == (1, 256, 4096)
    r = rel(out, ref)
    assert r < 1e-2, f"ViT-G 448^2 vs fp32 oracle: rel err {r}"      # north star: 1e-2 rel


def test_llama_layer_full_dims_prefill_1041_and_decode_at_ctx_1041(cuda_dev):
    """One full-width Llama-2-7B decoder layer (LoRA r=16 merged in the engine, unmerged in the oracle) + final norm +
    lm_head 32066x4096: prefill of 1041 tokens (tcgen05 GEMMs, paged causal attention), then a decode step at context
    1041 (skinny GEMMs + split-KV decode attention), then the 66-token image-run chunk on top of the cache.
    Reference ops: modeling_llama_xformer.py:318-368 (layer), :217-301 (attention), :759 (lm_head)."""
    from oracle import llama_oracle as LO
    from seedstory import llama_engine
    hidden, inter, heads, vocab, T = 4096, 11008, 32, 32066, 1041
    p = LO.LlamaParams.random(hidden, inter, heads, 1, vocab, lora_r=16, seed=21, std=0.02)
    p16 = p.to(dtype=torch.float16)
    cfg = llama_engine.LlamaConfig(hidden=hidden, inter=inter, heads=heads, layers=1, vocab=vocab, eps=p.eps, max_pos=4096)
    eng = llama_engine.LlamaEngine(cfg, cuda_dev, max_batch=1, max_ctx=2048, max_new=128)
    eng.load_weights(p16.embed, p16.layers, p16.norm, p16.lm_head, lora_scaling=p.scaling)
    pf = p16.to(dtype=torch.float32)   # the oracle computes in fp32 on the same fp16-rounded tensors
    g = torch.Generator().manual_seed(3)
    emb = (torch.randn(1, T, hidden, generator=g) * 0.05).half()
    torch.set_num_threads(min(32, torch.get_num_threads() or 8))
    # ---- prefill
    hn, logits = eng.forward_chunk(0, emb[0].to(cuda_dev), list(range(T)))
    with torch.no_grad():
        lo_ref, hn_ref, kv = LO.model_forward(pf, emb.float(), torch.arange(T).unsqueeze(0), None)
    assert rel(hn, hn_ref[0]) < 1e-2, rel(hn, hn_ref[0])
    assert rel(logits[0], lo_ref[0, -1]) < 1e-2, rel(logits[0], lo_ref[0, -1])
    kview = llama_engine.PagedKVView(eng, 0)[0]
    assert rel(kview[0], kv[0][0]) < 1e-2 and rel(kview[1], kv[0][1]) < 1e-2   # post-RoPE K, raw V as cached
    # ---- one decode step at context 1041 (graph path)
    tok = 12345
    eng.begin_decode([tok], [T])
    eng.decode_step(1)
    torch.cuda.synchronize()
    with torch.no_grad():
        lo1, hn1, kv = LO.model_forward(pf, pf.embed[torch.tensor([[tok]])], torch.tensor([[T]]), kv)
    assert rel(eng.hist[0, 1], hn1[0, 0]) < 1e-2, rel(eng.hist[0, 1], hn1[0, 0])
    # d_logits was edited in place by the (disabled) processor only if image ids are set: none here -> raw logits
    assert rel(eng.d_logits[0], lo1[0, 0]) < 1e-2, rel(eng.d_logits[0], lo1[0, 0])
    assert int(eng.cur_ids[0].item()) == int(torch.argmax(lo1[0, 0]).item()) or \
        LO.top2_margin(lo1[0, 0]) < 2e-2
    eng.seq_len_h[0] += 1
    # ---- 66-token chunk on top of the 1042-token cache (the forced <img> run as the engine feeds it)
    C = 66
    emb_c = (torch.randn(1, C, hidden, generator=g) * 0.05).half()
    pos = list(range(T + 1, T + 1 + C))
    hn_c, lo_c = eng.forward_chunk(0, emb_c[0].to(cuda_dev), pos)
    with torch.no_grad():
        lo2, hn2, _ = LO.model_forward(pf, emb_c.float(), torch.tensor([pos]), kv)
    assert rel(hn_c, hn2[0]) < 1e-2, rel(hn_c, hn2[0])
    assert rel(lo_c[0], lo2[0, -1]) < 1e-2


FULLWIDTH_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=1,
                      transformer_layers_per_block=(0, 1, 1), num_attention_heads=(5, 10, 20), cross_attention_dim=2048,
                      addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816, norm_num_groups=32,
                      sample_size=128)


def test_unet_full_width_blocks_match_oracle(cuda_dev):
    """SDXL UNet at full width and resolution (ResnetBlocks 320@128^2 .. 1280@32^2 incl. the 2560->1280 / 960->320
    concat blocks, CrossAttn transformer blocks C=640 / 4096 tokens and C=1280 / 1024 tokens, 2048-wide 64-token
    context, CFG batch 2) with the depth cut to one block per level so the fp32 CPU oracle finishes in seconds."""
    from oracle import sdxl_oracle as SO
    from seedstory import ops, sdxl_engine, synthetic
    cfg = FULLWIDTH_UNET
    sd = synthetic.random_unet_state_dict(cfg, seed=7)
    sd16 = {k: v.half().float() for k, v in sd.items()}
    eng = sdxl_engine.UNetEngine(sd, cfg, cuda_dev)
    g = torch.Generator().manual_seed(2)
    S = cfg["sample_size"]
    ctx = (torch.randn(2, 64, 2048, generator=g) * 0.5).half()
    pooled = (torch.randn(2, 1280, generator=g) * 0.5).half()
    ts, sig = sdxl_engine.euler_schedule(20)
    tid = [[8 * S, 8 * S, 0, 0, 8 * S, 8 * S]] * 2
    eng.set_conditioning(ctx, pooled, tid, ts)
    x = torch.randn(1, 4, S, S, generator=g).half()
    eng.x_in.zero_()
    eng.x_in[:, :, :, :4] = x.permute(0, 2, 3, 1).to(cuda_dev)
    eng.temb_cur.copy_(eng.temb_all[0:2])
    ops.fmha_path_counts(reset=True)
    eps = eng.forward()[..., :4].permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    tc, mma = ops.fmha_path_counts()
    # every transformer layer: self-attention on tcgen05 (4096 / 1024 tokens), cross-attention (64 keys) on the
    # mma.sync tile (down 2 + mid 1 + up 2 x 2 Transformer2D blocks of depth 1 = 7 layers)
    n_layers = sum(t.depth for t in eng._all_t2d())
    assert n_layers == 7 and tc == n_layers and mma == n_layers, (tc, mma, n_layers)
    torch.set_num_threads(min(32, torch.get_num_threads() or 8))
    with torch.no_grad():
        ref = SO.unet_forward(sd16, cfg, x.float().repeat(2, 1, 1, 1), torch.tensor([ts[0], ts[0]]), ctx.float(),
                              pooled.float(), torch.tensor(tid, dtype=torch.float32))
    r = rel(eps, ref)
    assert r < 1e-2, f"full-width UNet forward vs fp32 oracle: rel err {r}"


def test_vae_full_width_top_levels_match_oracle(cuda_dev):
    """VAE decoder at the real channel widths (512/512/256/128, single-head 512-wide mid attention) on a 32x32
    latent (256x256 image): bf16 engine vs fp32 oracle."""
    from oracle import sdxl_oracle as SO
    from seedstory import sdxl_engine, synthetic
    cfg = dict(synthetic.SDXL_VAE_CONFIG)
    sd = synthetic.random_vae_decoder_state_dict(cfg, seed=9)
    sdb = {k: v.bfloat16().float() for k, v in sd.items()}
    eng = sdxl_engine.VAEDecoderEngine(sd, cfg, cuda_dev)
    S = 32
    lat = (torch.randn(1, 4, S, S, generator=torch.Generator().manual_seed(4)) * 0.2).half()
    u8, img = eng.decode(lat.permute(0, 2, 3, 1).reshape(S * S, 4).contiguous().to(cuda_dev), S)
    torch.set_num_threads(min(32, torch.get_num_threads() or 8))
    with torch.no_grad():
        ref = SO.vae_decode(sdb, cfg, lat.float())
    got = img[0, :, :, :3].permute(2, 0, 1)[None]
    r = rel(got, ref)
    assert r < 5e-2, f"VAE decode (bf16 activations vs fp32 oracle) rel err {r}"
    ref_u8 = SO.postprocess(ref)[0]
    mse = ((u8.cpu().float() - ref_u8.float()) ** 2).mean().item()
    psnr = 10 * math.log10(255.0 ** 2 / max(mse, 1e-9))
    assert psnr > 35.0, f"decoded image PSNR vs fp32 oracle {psnr:.1f} dB"


def test_dropins_generate_and_get_image_embeds_match_oracle(cuda_dev):
    """ContinuousLVLM.generate (models.py:98-221) and SDXLAdapter.get_image_embeds (adapter_modules.py:387-428) called
    through the src.* drop-ins, compared with the oracle's restatement of the same call chain: resampled image
    tokens scattered into the prompt, greedy loop with the image-token processor, hidden rows of the 64 queries before
    the last </img>, output resampler; then ViT(zero image) + ResamplerXLV2 for the unconditional branch."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "seed-story_b200", "shims"))
    from oracle import llama_oracle as LO
    from oracle import vision_oracle as VO
    from seedstory import story
    from src.models_clm.generation import AutoImageTokenGenerationProcessor, ForcedScheduleProcessor
    from _parity import check_greedy_ids
    cfg = story.TINY
    pipe = story.StoryPipeline(device=cuda_dev, cfg=cfg, num_inference_steps=1, n_text_tokens=6, window_size=4)
    tk = pipe.tokenizer
    g = torch.Generator().manual_seed(8)
    img = torch.randn(1, 3, 56, 56, generator=g).half()
    cap = [7, 19, 44, 101, 5]
    ids = [tk.bos_token_id] + cap + pipe.image_ids
    L = len(ids)
    ids_t = torch.tensor([ids], device=cuda_dev)
    image_embeds = pipe.visual_encoder(img.to(cuda_dev))                       # [1, 256, 256]
    boi, eoi = ids.index(tk.boi), ids.index(tk.eoi)
    cmp_mask = torch.zeros_like(ids_t, dtype=torch.bool)
    cmp_mask[0, boi + 1:eoi] = True
    n_free = 6
    sched = [-1] * n_free + [tk.boi]
    out = pipe.agent.generate(tokenizer=tk, input_ids=ids_t, image_embeds=image_embeds,
                              embeds_cmp_mask=torch.ones(1, dtype=torch.bool, device=cuda_dev), ids_cmp_mask=cmp_mask,
                              max_new_tokens=90, num_img_gen_tokens=64,
                              logits_processor=[AutoImageTokenGenerationProcessor(tk, 64), ForcedScheduleProcessor(sched)],
                              device=cuda_dev)
    assert out["has_img_output"] and out["img_gen_feat"].shape == (1, 256, cfg["agent_dim"])
    # ---- oracle restatement of the same call chain on the same (fp16-rounded) weights, fp32 math
    f32 = lambda sd_: {k: v.detach().float().cpu() for k, v in sd_.items()}
    vit_sd = f32(pipe.visual_encoder.state_dict())
    v = cfg["vit"]
    with torch.no_grad():
        emb_ref = VO.vit_forward(vit_sd, img.float(), v["heads"], v["layers"], v["patch_size"])
    assert rel(image_embeds, emb_ref) < 1e-2
    agent_sd = f32(pipe.agent.state_dict())
    lm_in = VO.resampler(agent_sd, emb_ref, cfg["agent_heads"], prefix="input_resampler.")      # [1, 64, E]
    llm = pipe.agent.llm.base_model.model
    lc = cfg["llama"]
    p = LO.LlamaParams(lc["hidden_size"], lc["intermediate_size"], lc["num_attention_heads"], lc["num_hidden_layers"],
                       cfg["vocab"], eps=lc["rms_norm_eps"], lora_r=16, scaling=2.0)
    p.embed = llm.model.embed_tokens.weight.detach().float().cpu()
    p.norm = llm.model.norm.weight.detach().float().cpu()
    p.lm_head = llm.lm_head.weight.detach().float().cpu()
    for layer in llm.model.layers:
        Ld = {}
        for parent in (layer.self_attn, layer.mlp):
            for name, mod in parent._modules.items():
                Ld[name] = mod.weight.detach().float().cpu()
                Ld[name + ".lora_A"] = mod.lora_A["default"].weight.detach().float().cpu()
                Ld[name + ".lora_B"] = mod.lora_B["default"].weight.detach().float().cpu()
        Ld["input_layernorm"] = layer.input_layernorm.weight.detach().float().cpu()
        Ld["post_attention_layernorm"] = layer.post_attention_layernorm.weight.detach().float().cpu()
        p.layers.append(Ld)
    emb = p.embed[torch.tensor([ids])].clone()
    emb[0, boi + 1:eoi] = lm_in[0]                                              # models.py:135
    margins = []
    seq, hid, _ = LO.greedy_generate(p, torch.tensor([ids]), emb, pipe.image_ids, tk.eos_token_id, 90,
                                     forced_schedule=[None] * n_free + [tk.boi], margins_out=margins)
    gen_ref = seq[L:]
    gen = out["generate_ids"].tolist()
    k = check_greedy_ids(gen, gen_ref, margins, what="ContinuousLVLM.generate")
    if k == min(len(gen), len(gen_ref)):      # identical inputs all the way: the regressed image features must agree
        feats = LO.lvlm_postprocess(gen_ref, hid[L:], tk.eoi, 64)
        feat_ref = VO.resampler(agent_sd, feats.unsqueeze(0), cfg["agent_heads"], prefix="output_resampler.")
        assert rel(out["img_gen_feat"], feat_ref) < 2e-2, rel(out["img_gen_feat"], feat_ref)
    else:
        assert margins[k] < 2e-2
    # ---- SDXLAdapter.get_image_embeds: conditional = the regressed features, unconditional = ViT(zeros) (cached)
    pe, ne, pp, npool = pipe.adapter.get_image_embeds(image_embeds=out["img_gen_feat"], image_size=v["image_size"])
    xl = cfg["xl"]
    xl_sd = f32(pipe.adapter.resampler.state_dict())
    with torch.no_grad():
        zero_emb = VO.vit_forward(vit_sd, torch.zeros(1, 3, 56, 56), v["heads"], v["layers"], v["patch_size"])
        both = torch.cat([out["img_gen_feat"].float().cpu(), zero_emb], 0)
        e_ref, pool_ref = VO.resampler_xl_v2(xl_sd, both, xl["depth"], xl["heads"])
    assert rel(pe, e_ref[:1]) < 1e-2 and rel(ne, e_ref[1:]) < 1e-2, (rel(pe, e_ref[:1]), rel(ne, e_ref[1:]))
    assert rel(pp, pool_ref[:1]) < 1e-2 and rel(npool, pool_ref[1:]) < 1e-2
    # second call hits the cached zero-image branch: identical result
    pe2, ne2, _, _ = pipe.adapter.get_image_embeds(image_embeds=out["img_gen_feat"], image_size=v["image_size"])
    assert torch.equal(ne, ne2) and torch.equal(pe, pe2)
