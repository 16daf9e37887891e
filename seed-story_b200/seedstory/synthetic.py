"""Seeded random weights with the upstream checkpoint key layouts (no checkpoints exist offline — SURVEY.md §8d).

Used by the `diffusers` façade's from_pretrained fallbacks, the tests and bench.py so that the CUDA engines and
the oracle are driven by the SAME tensors.  SDXL-base-1.0 shapes follow SURVEY.md Appendix C."""
import math
import os

import torch


def missing_checkpoint(path, what):
    """A checkpoint path that does not exist is an error, as in the reference (torch.load / HF from_pretrained
    raise).  Seeded random weights of the real shapes are only substituted when the caller opted in with
    SEEDSTORY_SYNTHETIC=1 (tests / benchmarks on machines without the checkpoints)."""
    if os.environ.get("SEEDSTORY_SYNTHETIC", "0") != "1":
        raise FileNotFoundError(f"{what}: checkpoint path {path!r} does not exist (set SEEDSTORY_SYNTHETIC=1 to run "
                                f"with seeded random weights of the real shapes instead)")
    print(f"[seedstory_b200] SEEDSTORY_SYNTHETIC=1: {path} not found, {what} uses seeded random weights")

SDXL_UNET_CONFIG = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280), layers_per_block=2,
                        transformer_layers_per_block=(0, 2, 10), num_attention_heads=(5, 10, 20),
                        cross_attention_dim=2048, addition_time_embed_dim=256,
                        projection_class_embeddings_input_dim=2816, norm_num_groups=32, sample_size=128)

SDXL_VAE_CONFIG = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                       norm_num_groups=32, scaling_factor=0.13025)


def random_unet_state_dict(cfg, seed=0, std=0.02, dtype=torch.float32, device="cpu"):
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def rn(*shape, s=std):
        return (torch.randn(*shape, generator=g, device=device) * s).to(dtype)

    def lin(pre, o, i, bias=True, s=None):
        sd[pre + ".weight"] = rn(o, i, s=s if s is not None else 1.0 / math.sqrt(i))
        if bias:
            sd[pre + ".bias"] = rn(o, s=0.02)

    def conv(pre, o, i, k):
        sd[pre + ".weight"] = rn(o, i, k, k, s=1.0 / math.sqrt(i * k * k))
        sd[pre + ".bias"] = rn(o, s=0.02)

    def norm(pre, c):
        sd[pre + ".weight"] = (1.0 + rn(c, s=0.05)).to(dtype)
        sd[pre + ".bias"] = rn(c, s=0.05)

    def resnet(pre, i, o, temb):
        norm(pre + ".norm1", i)
        conv(pre + ".conv1", o, i, 3)
        if temb:
            lin(pre + ".time_emb_proj", o, temb)
        norm(pre + ".norm2", o)
        conv(pre + ".conv2", o, o, 3)
        if i != o:
            conv(pre + ".conv_shortcut", o, i, 1)

    def t2d(pre, c, depth, cross):
        norm(pre + ".norm", c)
        lin(pre + ".proj_in", c, c)
        for k in range(depth):
            b = f"{pre}.transformer_blocks.{k}"
            for n, kv in (("attn1", c), ("attn2", cross)):
                norm(f"{b}.norm{1 if n == 'attn1' else 2}", c)
                lin(f"{b}.{n}.to_q", c, c, bias=False)
                lin(f"{b}.{n}.to_k", c, kv, bias=False)
                lin(f"{b}.{n}.to_v", c, kv, bias=False)
                lin(f"{b}.{n}.to_out.0", c, c)
            norm(b + ".norm3", c)
            lin(b + ".ff.net.0.proj", 8 * c, c)
            lin(b + ".ff.net.2", c, 4 * c)
        lin(pre + ".proj_out", c, c)
    ch = cfg["block_out_channels"]
    temb = 4 * ch[0]
    lin("time_embedding.linear_1", temb, ch[0])
    lin("time_embedding.linear_2", temb, temb)
    lin("add_embedding.linear_1", temb, cfg["projection_class_embeddings_input_dim"])
    lin("add_embedding.linear_2", temb, temb)
    conv("conv_in", ch[0], cfg["in_channels"], 3)
    nb = len(ch)
    skip_ch = [ch[0]]
    prev = ch[0]
    for i in range(nb):
        for j in range(cfg["layers_per_block"]):
            resnet(f"down_blocks.{i}.resnets.{j}", prev, ch[i], temb)
            prev = ch[i]
            if cfg["transformer_layers_per_block"][i]:
                t2d(f"down_blocks.{i}.attentions.{j}", ch[i], cfg["transformer_layers_per_block"][i],
                    cfg["cross_attention_dim"])
            skip_ch.append(prev)
        if i < nb - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", ch[i], ch[i], 3)
            skip_ch.append(prev)
    resnet("mid_block.resnets.0", prev, prev, temb)
    t2d("mid_block.attentions.0", prev, cfg["transformer_layers_per_block"][-1], cfg["cross_attention_dim"])
    resnet("mid_block.resnets.1", prev, prev, temb)
    for i in range(nb):
        ri = nb - 1 - i
        for j in range(cfg["layers_per_block"] + 1):
            s = skip_ch.pop()
            resnet(f"up_blocks.{i}.resnets.{j}", prev + s, ch[ri], temb)
            prev = ch[ri]
            if cfg["transformer_layers_per_block"][ri]:
                t2d(f"up_blocks.{i}.attentions.{j}", ch[ri], cfg["transformer_layers_per_block"][ri],
                    cfg["cross_attention_dim"])
        if i < nb - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", ch[ri], ch[ri], 3)
    norm("conv_norm_out", ch[0])
    conv("conv_out", cfg["out_channels"], ch[0], 3)
    return sd


def random_vae_decoder_state_dict(cfg, seed=0, dtype=torch.float32, device="cpu"):
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def rn(*shape, s):
        return (torch.randn(*shape, generator=g, device=device) * s).to(dtype)

    def conv(pre, o, i, k):
        sd[pre + ".weight"] = rn(o, i, k, k, s=1.0 / math.sqrt(i * k * k))
        sd[pre + ".bias"] = rn(o, s=0.02)

    def norm(pre, c):
        sd[pre + ".weight"] = (1.0 + rn(c, s=0.05)).to(dtype)
        sd[pre + ".bias"] = rn(c, s=0.05)

    def lin(pre, o, i):
        sd[pre + ".weight"] = rn(o, i, s=1.0 / math.sqrt(i))
        sd[pre + ".bias"] = rn(o, s=0.02)

    def resnet(pre, i, o):
        norm(pre + ".norm1", i)
        conv(pre + ".conv1", o, i, 3)
        norm(pre + ".norm2", o)
        conv(pre + ".conv2", o, o, 3)
        if i != o:
            conv(pre + ".conv_shortcut", o, i, 1)
    ch = list(reversed(cfg["block_out_channels"]))
    lc = cfg["latent_channels"]
    conv("post_quant_conv", lc, lc, 1)
    conv("decoder.conv_in", ch[0], lc, 3)
    resnet("decoder.mid_block.resnets.0", ch[0], ch[0])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", ch[0])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{n}", ch[0], ch[0])
    resnet("decoder.mid_block.resnets.1", ch[0], ch[0])
    prev = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg["layers_per_block"] + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev, c)
            prev = c
        if i < len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("decoder.conv_norm_out", prev)
    conv("decoder.conv_out", cfg["out_channels"], prev, 3)
    return sd
