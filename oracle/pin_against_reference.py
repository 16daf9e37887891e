"""ORACLE PINNING — runs only in the build container (needs /root/reference; never on the GPU box).

Imports the reference's own modules, runs them on seeded random weights at small dims (CPU fp32) next to
the oracle restatements, asserts agreement, and freezes input/output vectors under tests/golden/ so that
the CPU test suite and the GPU parity tests have reference-generated fixtures.

    python oracle/pin_against_reference.py            # writes tests/golden/*.pt

Reference pieces exercised (file:line):
  src/models_clm/modeling_llama_xformer.py:703-794 (LlamaForCausalLM.forward, xformers -> SDPA stand-in)
  src/models_clm/modeling_llama_xformer.py:796-852 (prepare_inputs_for_generation, use_kv_cache_head=True KV reuse)
  src/models_clm/generation.py:9-31              (AutoImageTokenGenerationProcessor)
  src/models/qwen_visual.py:376-399, 138-150       (ViT + attn-pool, agent Resampler)
  src/models_ipa/resampler.py:228-284             (ResamplerXLV2)
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SEEDSTORY_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import llama_oracle as LO  # noqa: E402
from oracle import vision_oracle as VO  # noqa: E402


def _install_xformers_stub():
    """xformers is not installed: stand in for the two symbols the reference touches
    (modeling_llama_xformer.py:282-295) with SDPA + an explicit bottom-right causal mask."""
    xf = types.ModuleType("xformers")
    xops = types.ModuleType("xformers.ops")
    fmha = types.ModuleType("xformers.ops.fmha")
    ab = types.ModuleType("xformers.ops.fmha.attn_bias")

    class LowerTriangularFromBottomRightMask:
        pass

    class LowerTriangularMask:
        pass

    def memory_efficient_attention(q, k, v, attn_bias=None):
        # q,k,v: [B, T, H, D]
        tq, tk = q.shape[1], k.shape[1]
        qi = torch.arange(tq).unsqueeze(1)
        kj = torch.arange(tk).unsqueeze(0)
        if isinstance(attn_bias, LowerTriangularFromBottomRightMask):
            mask = kj <= qi + (tk - tq)
        else:
            mask = kj <= qi
        o = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                             attn_mask=mask)
        return o.transpose(1, 2)
    ab.LowerTriangularFromBottomRightMask = LowerTriangularFromBottomRightMask
    fmha.attn_bias = ab
    xops.fmha = fmha
    xops.LowerTriangularMask = LowerTriangularMask
    xops.memory_efficient_attention = memory_efficient_attention
    xf.ops = xops
    for n, m in {"xformers": xf, "xformers.ops": xops, "xformers.ops.fmha": fmha,
                 "xformers.ops.fmha.attn_bias": ab}.items():
        sys.modules[n] = m


def _maxdiff(a, b):
    return (a.float() - b.float()).abs().max().item()


def pin_llama():
    _install_xformers_stub()
    sys.path.insert(0, REF)
    from src.models_clm import modeling_llama_xformer as M  # the reference file itself
    from transformers import LlamaConfig
    hidden, inter, heads, layers, vocab = 256, 352, 2, 3, 320  # head_dim 128 like Llama-2-7B
    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_attention_heads=heads,
                      num_hidden_layers=layers, vocab_size=vocab, rms_norm_eps=1e-5, max_position_embeddings=512,
                      pad_token_id=0)
    cfg._attn_implementation = "eager"
    ref = M.LlamaForCausalLM(cfg).eval()
    p = LO.LlamaParams.random(hidden, inter, heads, layers, vocab, lora_r=0, seed=7, std=0.05)
    sd = {"model.embed_tokens.weight": p.embed, "model.norm.weight": p.norm, "lm_head.weight": p.lm_head}
    for i, L in enumerate(p.layers):
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[f"model.layers.{i}.self_attn.{n}.weight"] = L[n]
        for n in ("gate_proj", "up_proj", "down_proj"):
            sd[f"model.layers.{i}.mlp.{n}.weight"] = L[n]
        sd[f"model.layers.{i}.input_layernorm.weight"] = L["input_layernorm"]
        sd[f"model.layers.{i}.post_attention_layernorm.weight"] = L["post_attention_layernorm"]
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not [m for m in missing if "rotary" not in m and "inv_freq" not in m], missing
    ref.use_kv_cache_head = False
    g = torch.Generator().manual_seed(11)
    T0, T1 = 37, 5
    emb0 = torch.randn(1, T0, hidden, generator=g) * 0.5
    pos0 = torch.arange(T0).unsqueeze(0)
    with torch.no_grad():
        out0 = ref(input_ids=torch.zeros(1, T0, dtype=torch.long), inputs_embeds=emb0, position_ids=pos0,
                   use_cache=True, output_hidden_states=True, return_dict=True)
        lo0, hn0, kv0 = LO.model_forward(p, emb0, pos0, None, max_pos=512)
    d = _maxdiff(out0.logits, lo0)
    assert d < 2e-4, d
    assert _maxdiff(out0.hidden_states[-1], hn0) < 2e-4
    # second call: 5-token chunk on top of the cache, window-relative positions (bottom-right mask exercised)
    emb1 = torch.randn(1, T1, hidden, generator=g) * 0.5
    pos1 = (torch.arange(T1) + T0).unsqueeze(0)
    with torch.no_grad():
        out1 = ref(input_ids=torch.zeros(1, T1, dtype=torch.long), inputs_embeds=emb1, position_ids=pos1,
                   past_key_values=out0.past_key_values, use_cache=True, output_hidden_states=True, return_dict=True)
        lo1, hn1, kv1 = LO.model_forward(p, emb1, pos1, kv0, max_pos=512)
    assert _maxdiff(out1.logits, lo1) < 2e-4
    assert _maxdiff(out1.past_key_values[1][0], kv1[1][0]) < 1e-5  # post-RoPE keys
    print("llama forward pinned: max logits diff", d)
    torch.save({"cfg": dict(hidden=hidden, inter=inter, heads=heads, layers=layers, vocab=vocab, eps=1e-5, seed=7,
                            std=0.05),
                "emb0": emb0, "emb1": emb1, "logits0": out0.logits, "logits1": out1.logits,
                "hidden0": out0.hidden_states[-1], "k_layer1": out1.past_key_values[1][0]},
               os.path.join(GOLD, "llama_forward.pt"))

    # logits processor: the reference class with a fake tokenizer
    from src.models_clm.generation import AutoImageTokenGenerationProcessor

    class Tok:
        def encode(self, s, add_special_tokens=False):
            return [300] + list(range(302, 302 + 8)) + [301]
    proc = AutoImageTokenGenerationProcessor(Tok(), num_img_gen_tokens=8)
    img_ids = proc.img_ids_list
    cases = []
    for last in (5, 300, 303, 309, 301):
        sc = torch.randn(1, vocab, generator=g)
        exp = proc(torch.tensor([[1, 2, last]]), sc.clone())
        got = LO.image_token_processor(last, sc[0].clone(), img_ids)
        assert torch.equal(exp[0], got)
        cases.append({"last": last, "scores": sc[0], "out": exp[0]})
    torch.save({"img_ids": img_ids, "cases": cases}, os.path.join(GOLD, "logits_processor.pt"))
    print("logits processor pinned")


def pin_greedy_loop():
    """The greedy loop + hidden-state bookkeeping lives in `transformers`, which is not vendored in the
    reference (pinned there to 4.34.0; this container has 5.5).  Pin the oracle's restatement of it against the
    installed `GenerationMixin.generate` driven the way models.py:146-153 drives it (input_ids + inputs_embeds, greedy,
    the REFERENCE's AutoImageTokenGenerationProcessor, output_hidden_states) on a tiny HF LlamaForCausalLM carrying the
    same weights.  Two prompts: free-running text, and a prompt ending in <img> so the forced image run, </img> and the
    return to free decoding are exercised."""
    sys.path.insert(0, REF)
    from transformers import LlamaConfig, LlamaForCausalLM, LogitsProcessorList
    from src.models_clm.generation import AutoImageTokenGenerationProcessor
    hidden, inter, heads, layers, vocab = 256, 352, 2, 3, 320
    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_attention_heads=heads, num_key_value_heads=heads,
                      num_hidden_layers=layers, vocab_size=vocab, rms_norm_eps=1e-5, max_position_embeddings=512,
                      pad_token_id=0, bos_token_id=1, eos_token_id=2, attention_bias=False, mlp_bias=False,
                      tie_word_embeddings=False)
    cfg._attn_implementation = "eager"
    hf = LlamaForCausalLM(cfg).eval()
    p = LO.LlamaParams.random(hidden, inter, heads, layers, vocab, lora_r=0, seed=9, std=0.05)
    sd = {"model.embed_tokens.weight": p.embed, "model.norm.weight": p.norm, "lm_head.weight": p.lm_head}
    for i, L in enumerate(p.layers):
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[f"model.layers.{i}.self_attn.{n}.weight"] = L[n]
        for n in ("gate_proj", "up_proj", "down_proj"):
            sd[f"model.layers.{i}.mlp.{n}.weight"] = L[n]
        sd[f"model.layers.{i}.input_layernorm.weight"] = L["input_layernorm"]
        sd[f"model.layers.{i}.post_attention_layernorm.weight"] = L["post_attention_layernorm"]
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)

    class Tok:
        def encode(self, s, add_special_tokens=False):
            return [300] + list(range(302, 310)) + [301]
    proc = AutoImageTokenGenerationProcessor(Tok(), num_img_gen_tokens=8)
    img_ids = proc.img_ids_list
    g = torch.Generator().manual_seed(3)
    cases = []
    for name, tail in (("text", []), ("image_run", [300])):
        ids = torch.cat([torch.randint(3, 290, (1, 19), generator=g), torch.tensor([tail], dtype=torch.long)], 1)
        emb = p.embed[ids]
        with torch.no_grad():
            out = hf.generate(input_ids=ids, inputs_embeds=emb, max_new_tokens=24, do_sample=False, num_beams=1,
                              logits_processor=LogitsProcessorList([proc]), output_hidden_states=True,
                              return_dict_in_generate=True, eos_token_id=2, pad_token_id=0)
        seq_hf = out.sequences[0].tolist()
        hs_hf = torch.cat([st[-1][0] for st in out.hidden_states], 0)
        seq, hid, _ = LO.greedy_generate(p, ids, emb, img_ids, 2, 24)
        assert seq == seq_hf, (name, seq, seq_hf)
        d = _maxdiff(hs_hf, hid[:hs_hf.shape[0]])
        assert d < 1e-4, d
        if tail:
            assert seq[ids.shape[1]:ids.shape[1] + 9] == list(range(302, 310)) + [301], seq
        cases.append({"name": name, "input_ids": ids, "sequence": seq_hf, "hidden": hs_hf})
        print(f"greedy loop pinned against transformers.generate ({name}): ids equal, max hidden diff {d:.2e}")
    import transformers
    torch.save({"cfg": dict(hidden=hidden, inter=inter, heads=heads, layers=layers, vocab=vocab, eps=1e-5, seed=9,
                            std=0.05), "img_ids": img_ids, "eos": 2, "max_new_tokens": 24,
                "transformers_version": transformers.__version__, "cases": cases},
               os.path.join(GOLD, "hf_greedy_loop.pt"))


def pin_vision():
    sys.path.insert(0, REF)
    from src.models import qwen_visual as QV
    from src.models_ipa import resampler as RS
    torch.manual_seed(3)
    # small ViT: width 64, 4 heads (head_dim 16), 2 layers, 56x56 image, patch 14 -> 16 tokens; pos table is 256 (16x16)
    vit = QV.VisionTransformerWithAttnPool(image_size=56, patch_size=14, width=64, layers=2, heads=4,
                                           mlp_ratio=4.0, n_queries=16, output_dim=256).eval()
    with torch.no_grad():
        for prm in vit.parameters():
            if prm.requires_grad:
                prm.add_(torch.randn_like(prm) * 0.05)
    img = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        ref = vit(img)
        got = VO.vit_forward(vit.state_dict(), img, heads=4, layers=2, patch=14)
    d = _maxdiff(ref, got)
    assert d < 1e-4, d
    print("vit pinned: max diff", d)
    torch.save({"sd": vit.state_dict(), "img": img, "out": ref,
                "cfg": dict(image_size=56, patch_size=14, width=64, layers=2, heads=4, mlp_ratio=4.0, n_queries=16,
                            output_dim=256)}, os.path.join(GOLD, "vit_small.pt"))

    # agent resamplers (grid 8 over 256 keys -> bicubic up; grid 16 over 64 keys -> bicubic down)
    for name, grid, L in (("resampler_in", 8, 256), ("resampler_out", 16, 64)):
        rs = QV.Resampler(grid_size=grid, embed_dim=256, num_heads=2, kv_dim=256).eval()
        with torch.no_grad():
            for prm in rs.parameters():
                if prm.requires_grad:
                    prm.add_(torch.randn_like(prm) * 0.05)
        x = torch.randn(2, L, 256)
        with torch.no_grad():
            ref = rs(x)
            got = VO.resampler(rs.state_dict(), x, heads=2)
        d = _maxdiff(ref, got)
        assert d < 1e-4, d
        print(name, "pinned: max diff", d)
        torch.save({"sd": rs.state_dict(), "x": x, "out": ref, "grid": grid, "heads": 2},
                   os.path.join(GOLD, f"{name}.pt"))

    xl = RS.ResamplerXLV2(dim=256, depth=2, dim_head=64, heads=4, num_queries=16, embedding_dim=256, output1_dim=96,
                          output2_dim=160, ff_mult=4).eval()
    x = torch.randn(2, 64, 256)
    with torch.no_grad():
        r1, r2 = xl(x)
        g1, g2 = VO.resampler_xl_v2(xl.state_dict(), x, depth=2, heads=4)
    d = max(_maxdiff(r1, g1), _maxdiff(r2, g2))
    assert d < 1e-4, d
    print("ResamplerXLV2 pinned: max diff", d)
    torch.save({"sd": xl.state_dict(), "x": x, "out1": r1, "out2": r2,
                "cfg": dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=16, embedding_dim=256,
                            output1_dim=96, output2_dim=160, ff_mult=4)}, os.path.join(GOLD, "resampler_xlv2.pt"))


def pin_sink_kv_reuse():
    """KV-reuse generation (`use_kv_cache_head=True`, the mode src/inference/vis_george_sink.py:266-295 prepares):
    the reference's OWN prepare_inputs_for_generation (:796-852) and forward (:703-794, kv_cache_head update :780-784)
    driven by a greedy loop written the way transformers 4.34 greedy_search drives a model
    (prepare_inputs_for_generation -> forward -> argmax -> extend attention_mask, carry past_key_values; inputs_embeds
    stays in model_kwargs), on a past that was sliced attention-sink style.  Freezes tests/golden/sink_kv_reuse.pt."""
    _install_xformers_stub()
    sys.path.insert(0, REF)
    from src.models_clm import modeling_llama_xformer as M
    from src.models_clm.generation import AutoImageTokenGenerationProcessor
    from transformers import LlamaConfig
    hidden, inter, heads, layers, vocab = 256, 352, 2, 3, 320
    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_attention_heads=heads,
                      num_hidden_layers=layers, vocab_size=vocab, rms_norm_eps=1e-5, max_position_embeddings=512,
                      pad_token_id=0)
    cfg._attn_implementation = "eager"
    ref = M.LlamaForCausalLM(cfg).eval()
    p = LO.LlamaParams.random(hidden, inter, heads, layers, vocab, lora_r=0, seed=13, std=0.05)
    sd = {"model.embed_tokens.weight": p.embed, "model.norm.weight": p.norm, "lm_head.weight": p.lm_head}
    for i, L in enumerate(p.layers):
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[f"model.layers.{i}.self_attn.{n}.weight"] = L[n]
        for n in ("gate_proj", "up_proj", "down_proj"):
            sd[f"model.layers.{i}.mlp.{n}.weight"] = L[n]
        sd[f"model.layers.{i}.input_layernorm.weight"] = L["input_layernorm"]
        sd[f"model.layers.{i}.post_attention_layernorm.weight"] = L["post_attention_layernorm"]
    ref.load_state_dict(sd, strict=False)
    img_ids = [300] + list(range(302, 310)) + [301]
    eos = 2

    class _Tok:
        def encode(self, text, add_special_tokens=False):
            return img_ids
    proc = AutoImageTokenGenerationProcessor(_Tok(), 8)
    g = torch.Generator().manual_seed(17)
    # turn 1: a 60-token context [text(10) <img> 8 queries </img> text(40)] is run in full; its cache is then sliced
    # like the sink script does after evicting that image: first 4 slots, [boi-4, boi+8), [eoi-8, eoi+4), live tail
    T = 60
    ids_full = torch.randint(3, 290, (1, T), generator=g)
    boi, eoi = 10, 19
    ids_full[0, boi:eoi + 1] = torch.tensor(img_ids)
    ref.use_kv_cache_head = False
    with torch.no_grad():
        out = ref(input_ids=ids_full, position_ids=torch.arange(T).unsqueeze(0), use_cache=True, return_dict=True)
    keep = sorted(set(range(4)) | set(range(boi - 4, boi + 8)) | set(range(eoi - 8, eoi + 4)) | set(range(eoi + 1, T)))
    past = tuple((k[:, :, keep], v[:, :, keep]) for (k, v) in out.past_key_values)
    # turn 2: the windowed input_ids = live tail (40 tokens, already cached) + 7 new tokens ending in <img>
    live = ids_full[:, eoi + 1:]
    new = torch.cat([torch.randint(3, 290, (1, 6), generator=g), torch.tensor([[300]])], dim=1)
    ids2 = torch.cat([live, new], dim=1)
    head = live.shape[1]
    L = ids2.shape[1]
    emb2 = p.embed[ids2]
    ref.use_kv_cache_head = True
    ref.kv_cache_head = head
    max_new = 14
    seq = ids2.clone()
    kwargs = dict(past_key_values=past, inputs_embeds=emb2, attention_mask=torch.ones(1, L, dtype=torch.long),
                  use_cache=True)
    hiddens, logits0 = [], None
    with torch.no_grad():
        for step in range(max_new):
            mi = ref.prepare_inputs_for_generation(seq, **kwargs)
            o = ref(**mi, return_dict=True, output_hidden_states=True)
            if logits0 is None:
                logits0 = o.logits.clone()
            hiddens.append(o.hidden_states[-1][0])
            scores = proc(seq, o.logits[:, -1, :].clone())
            nxt = int(torch.argmax(scores[0]).item())
            seq = torch.cat([seq, torch.tensor([[nxt]])], dim=1)
            kwargs["past_key_values"] = o.past_key_values
            kwargs["attention_mask"] = torch.cat([kwargs["attention_mask"], torch.ones(1, 1, dtype=torch.long)], dim=1)
            if nxt == eos:
                break
    rows = torch.cat(hiddens, 0)
    assert ref.kv_cache_head == head + (L - head) + (len(hiddens) - 1), ref.kv_cache_head
    # oracle on the same sliced past
    kv_o = None
    with torch.no_grad():
        _, _, kv_o = LO.model_forward(p, p.embed[ids_full], torch.arange(T).unsqueeze(0), None, max_pos=512)
        past_o = [(k[:, :, keep], v[:, :, keep]) for (k, v) in kv_o]
        seq_o, hid_o, _ = LO.greedy_generate(p, ids2, emb2, img_ids, eos, max_new, past_kvs=past_o, head=head)
    assert seq_o == seq[0].tolist(), (seq_o, seq[0].tolist())
    d = _maxdiff(hid_o[:rows.shape[0]], rows)
    assert d < 2e-4, d
    assert seq_o[L:L + 9] == img_ids[1:], "prompt ends in <img>: forced run + </img> first"
    print("sink KV-reuse generation pinned against the reference's prepare_inputs_for_generation/forward: max hidden diff", d)
    torch.save({"cfg": dict(hidden=hidden, inter=inter, heads=heads, layers=layers, vocab=vocab, eps=1e-5, seed=13,
                            std=0.05),
                "img_ids": img_ids, "eos": eos, "ids_full": ids_full, "keep": keep, "ids2": ids2, "head": head,
                "max_new": max_new, "sequence": seq[0].tolist(), "rows": rows, "logits0": logits0,
                "kv_cache_head_after": int(ref.kv_cache_head)},
               os.path.join(GOLD, "sink_kv_reuse.pt"))


def _install_diffusers_stub():
    """src/models_ipa/adapter_modules.py imports pipeline / LoRA classes from diffusers at module level (:6-22); diffusers is
    not installed here.  Empty stand-ins are enough to IMPORT the file: the method pinned below (get_image_embeds) touches
    none of them."""
    import types
    if "diffusers" in sys.modules:
        return
    d = types.ModuleType("diffusers")
    for n in ("StableDiffusionPipeline", "StableDiffusionXLPipeline", "StableDiffusionXLInstructPix2PixPipeline",
              "StableDiffusionInstructPix2PixPipeline"):
        setattr(d, n, type(n, (), {}))
    loaders = types.ModuleType("diffusers.loaders")
    loaders.LoraLoaderMixin = type("LoraLoaderMixin", (), {})
    models = types.ModuleType("diffusers.models")
    lora = types.ModuleType("diffusers.models.lora")
    lora.LoRALinearLayer = type("LoRALinearLayer", (), {})
    blocks = types.ModuleType("diffusers.models.unet_2d_blocks")
    blocks.DownBlock2D = type("DownBlock2D", (), {})
    d.loaders, d.models, models.lora, models.unet_2d_blocks = loaders, models, lora, blocks
    sys.modules.update({"diffusers": d, "diffusers.loaders": loaders, "diffusers.models": models,
                        "diffusers.models.lora": lora, "diffusers.models.unet_2d_blocks": blocks})


def pin_adapter_image_embeds():
    """SDXLAdapter.get_image_embeds / encode_image_embeds (src/models_ipa/adapter_modules.py:387-428, :345-348; row a13) —
    the reference's OWN methods, called unbound on an object that carries the attributes they use (visual_encoder,
    discrete_model, resampler): the unconditional branch is the visual encoder run on a ZERO image, concatenated after
    the conditional embeds, pushed through the identity discrete model and ResamplerXLV2 together, then chunked.  The
    oracle-side restatement (torch.cat + vision_oracle.resampler_xl_v2 + chunk, as tests/test_fullsize_gpu.py uses it) is
    checked against it and tests/golden/adapter_image_embeds.pt is frozen."""
    _install_diffusers_stub()
    sys.path.insert(0, REF)
    from src.models_ipa import resampler as RS
    from src.models_ipa.adapter_modules import SDXLAdapter
    torch.manual_seed(29)
    E, n_tok, S = 256, 64, 32      # (sizes the CUDA kernels accept as well: the same golden drives a GPU test of the drop-in)

    class _Encoder(torch.nn.Module):            # stand-in visual encoder: [B,3,S,S] -> [B, n_tok, E], input dependent
        def __init__(self):
            super().__init__()
            self.proj = torch.nn.Linear(3 * S * S // n_tok, E)

        def forward(self, x):
            return self.proj(x.reshape(x.shape[0], n_tok, -1)) + 0.1

    class _Identity(torch.nn.Module):           # DiscreteModleIdentity.encode_image_embeds (discrete_models.py:120-130)
        def encode_image_embeds(self, x):
            return x

    xl_cfg = dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=16, embedding_dim=E, output1_dim=96,
                  output2_dim=160, ff_mult=4)
    obj = SDXLAdapter.__new__(SDXLAdapter)
    torch.nn.Module.__init__(obj)
    obj.resampler = RS.ResamplerXLV2(**xl_cfg).eval()
    # the weights already frozen in resampler_xlv2.pt (same configuration): the golden then only carries activations
    obj.resampler.load_state_dict(torch.load(os.path.join(GOLD, "resampler_xlv2.pt"), weights_only=False)["sd"])
    obj.visual_encoder = _Encoder().eval()
    obj.discrete_model = _Identity()
    obj.image_transform = None
    feat = torch.randn(1, n_tok, E)
    with torch.no_grad():
        pe, ne, pp, npool = obj.get_image_embeds(image_embeds=feat, return_negative=True, image_size=S)
        zero_emb = obj.visual_encoder(torch.zeros(1, 3, S, S))
        e, pool = VO.resampler_xl_v2(obj.resampler.state_dict(), torch.cat([feat, zero_emb], 0), depth=2, heads=4)
    d = max(_maxdiff(pe, e[:1]), _maxdiff(ne, e[1:]), _maxdiff(pp, pool[:1]), _maxdiff(npool, pool[1:]))
    assert d < 1e-5, d
    print("SDXLAdapter.get_image_embeds pinned: max diff of the oracle restatement", d)
    torch.save({"cfg": xl_cfg, "image_size": S, "feat": feat, "zero_embeds": zero_emb, "resampler_sd_from": "resampler_xlv2.pt",
                "prompt": pe, "negative": ne, "pooled": pp, "negative_pooled": npool},
               os.path.join(GOLD, "adapter_image_embeds.pt"))


def pin_transform():
    """src/processer/transforms.py:4-47 (row a1): the reference's own get_transform on a seeded 300x400 RGB image, every
    type x keep_ratio at 448 (and the 'sd' type at 1024).  The full outputs are megabytes; the golden keeps a 16x16
    crop, the sum and the absolute sum of each (tests/golden/transform.pt)."""
    sys.path.insert(0, REF)
    import importlib.util
    import numpy as np
    from PIL import Image
    spec = importlib.util.spec_from_file_location("ref_transforms", os.path.join(REF, "src", "processer", "transforms.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    px = np.random.RandomState(5).randint(0, 256, (300, 400, 3), dtype=np.uint8)
    img = Image.fromarray(px)
    out = {}
    for typ, size in (("clip", 448), ("clipa", 448), ("sd", 448), ("sd", 1024)):
        for keep in (False, True):
            t = ref.get_transform(type=typ, image_size=size, keep_ratio=keep)(img)
            assert tuple(t.shape) == (3, size, size)
            out[f"{typ}/{size}/{int(keep)}"] = dict(crop=t[:, 200:216, 200:216].clone(), sum=float(t.double().sum()),
                                                   abs_sum=float(t.double().abs().sum()))
    print("get_transform pinned:", ", ".join(out))
    torch.save({"pixels_seed": 5, "pixels_shape": (300, 400, 3), "cases": out}, os.path.join(GOLD, "transform.pt"))


def pin_lvlm_generate():
    """ContinuousLVLM.generate (src/models_clm/models.py:98-221) — the reference's OWN class, run on CPU around a fake
    `llm` that replays a fixed generation (ids + per-step hidden states), so that everything the method itself computes
    is pinned: the scatter of the resampled image tokens into the <img_i> slots (:127-135), the slicing of the
    generated ids, the search for the LAST </img> in both branches (past_key_values None / given, :182-197), the 64
    hidden rows handed to the output resampler, has_img_output / num_gen_imgs.  Freezes tests/golden/lvlm_generate.pt
    and checks the oracle's lvlm_postprocess / lvlm_postprocess_past against it."""
    sys.path.insert(0, REF)
    import types
    from src.models_clm.models import ContinuousLVLM
    torch.manual_seed(23)               # the fake llm's nn.Embedding initialises from the global generator
    E, n_q = 32, 8                      # hidden width, image tokens per image (num_img_gen_tokens)
    boi, eoi, img0 = 300, 301, 302
    img_ids = [boi] + [img0 + i for i in range(n_q)] + [eoi]
    g = torch.Generator().manual_seed(23)

    class _Tok:
        eos_token_id = 2

        def encode(self, text, add_special_tokens=False):
            return {"<img>": [boi], "</img>": [eoi]}.get(text, img_ids)

        def decode(self, ids, skip_special_tokens=False):
            return " ".join(str(int(i)) for i in ids)

    class _Resampler(torch.nn.Module):          # deterministic stand-ins with the resamplers' shapes
        def __init__(self, n_out, scale):
            super().__init__()
            self.n_out, self.scale = n_out, scale

        def forward(self, x):
            return x[:, :self.n_out] * self.scale + 0.25

    class _FakeLLM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(320, E)
            self.past_key_values = "kv-of-this-call"
            self.replay = None

        def get_input_embeddings(self):
            return self.emb

        def generate(self, input_ids=None, inputs_embeds=None, **kw):
            gen, hs = self.replay
            self.seen_embeds = inputs_embeds.detach().clone()
            return types.SimpleNamespace(sequences=torch.cat([input_ids, torch.tensor([gen])], dim=1), hidden_states=hs,
                                         attentions=None)

    llm = _FakeLLM()
    model = ContinuousLVLM(llm, _Resampler(n_q, 2.0), _Resampler(2 * n_q, -1.5)).eval()
    tok = _Tok()
    prompt = [1, 11, 12] + img_ids + [13, 14] + img_ids          # two input images
    ids = torch.tensor([prompt])
    mask = torch.zeros_like(ids, dtype=torch.bool)
    b = [i for i, t in enumerate(prompt) if t == boi]
    e = [i for i, t in enumerate(prompt) if t == eoi]
    for k in range(2):
        mask[0, b[k] + 1:e[k]] = True
    image_embeds = torch.randn(3, 2 * n_q, E, generator=g)        # three candidate images, the middle one unused
    emb_mask = torch.tensor([True, False, True])
    cases = {}
    gens = {"one_run": [21, 22] + img_ids + [2],
            "two_runs_last_wins": [21] + img_ids + [22, 23] + img_ids + [2],
            "no_image": [21, 22, 23, 2]}
    L = len(prompt)
    for name, gen in gens.items():
        for branch in ("none", "past"):
            T = len(gen)
            # HF hands back one tuple per step; the last element of each is the final hidden state of that forward:
            # step 0 covers the fed prompt (all of it without a past, its tail [text, <img>..</img>] of 12 tokens with one —
            # the prompt tail of the sink script ends in an image, which the KV branch's </img> search can find: with no
            # new image generated the reference reports the PROMPT's last image, a quirk this fixture pins), then one row
            # per step
            fed0 = L if branch == "none" else 12
            hs = ((torch.randn(1, fed0, E, generator=g),),) + tuple((torch.randn(1, 1, E, generator=g),) for _ in range(T - 1))
            llm.replay = (gen, hs)
            with torch.no_grad():
                out = model.generate(tokenizer=tok, input_ids=ids.clone(), image_embeds=image_embeds, embeds_cmp_mask=emb_mask,
                                     ids_cmp_mask=mask, num_img_gen_tokens=n_q, max_new_tokens=64,
                                     past_key_values=None if branch == "none" else "sliced-cache", device="cpu")
            rows = torch.cat([h[-1] for h in hs], dim=1)[0]
            if branch == "none":
                feats = LO.lvlm_postprocess(gen, rows[L:], eoi, n_q)
            else:
                feats = LO.lvlm_postprocess_past(prompt + gen, rows, eoi, n_q)
            if out["has_img_output"]:
                ref_in = (out["img_gen_feat"] - 0.25) / -1.5                     # undo the stand-in output resampler
                assert feats is not None and _maxdiff(feats[None], ref_in) < 1e-6, (name, branch)
            else:
                assert feats is None and out["img_gen_feat"] is None, (name, branch)
            assert out["generate_ids"].tolist() == gen and out["past_key_values"] == "kv-of-this-call"
            cases[f"{name}/{branch}"] = dict(gen=gen, hidden_states=[h[-1] for h in hs], has_img_output=out["has_img_output"],
                                             num_gen_imgs=out["num_gen_imgs"], img_gen_feat=out["img_gen_feat"],
                                             text=out["text"], input_embeds=llm.seen_embeds)
    print("ContinuousLVLM.generate pinned:", ", ".join(cases))
    torch.save({"E": E, "n_q": n_q, "boi": boi, "eoi": eoi, "img_ids": img_ids, "prompt": prompt, "ids_cmp_mask": mask,
                "image_embeds": image_embeds, "embeds_cmp_mask": emb_mask, "emb_weight": llm.emb.weight.detach().clone(),
                "cases": cases}, os.path.join(GOLD, "lvlm_generate.pt"))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    pin_llama()
    pin_greedy_loop()
    pin_vision()
    pin_sink_kv_reuse()
    pin_transform()
    pin_adapter_image_embeds()
    pin_lvlm_generate()
    print("golden vectors written to", GOLD)
