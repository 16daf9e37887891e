"""Minimal `diffusers` façade for the names the reference imports (src/inference/gen_george.py:10,
src/models_ipa/adapter_modules.py:6-11,20-22) — used only when the real package is absent.

UNet2DConditionModel / AutoencoderKL are parameter containers with diffusers' state_dict key layout (so SDXL
checkpoints and the adapted de-tokenizer checkpoint load unchanged, and `unet.named_modules()` exposes the
`…to_k` / `…to_v` modules SDXLAdapter scans, adapter_modules.py:317-320); StableDiffusionXLPipeline runs the
sampler + VAE on seedstory.sdxl_engine (hand-written CUDA).  Nothing here computes with torch.
"""
import json
import os

import torch
from torch import nn

from seedstory import sdxl_engine, synthetic


def _build_tree(root, state_dict):
    """Register every tensor of `state_dict` as a parameter under nested nn.Module containers named after the
    dotted key path, so root.state_dict() reproduces the keys exactly."""
    for key, val in state_dict.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(val, requires_grad=False))


def _load_config(path, defaults, renames=()):
    """config.json of a diffusers model directory -> the subset of keys the engines use (tuples for lists)."""
    cfg = {}
    f = os.path.join(path, "config.json")
    if os.path.exists(f):
        with open(f) as fh:
            raw = json.load(fh)
        for src, dst in renames:
            if src in raw and dst not in raw:
                raw[dst] = raw[src]
        for k in defaults:
            if k in raw and raw[k] is not None:
                cfg[k] = tuple(raw[k]) if isinstance(raw[k], list) else raw[k]
    return cfg


def _load_dir(path):
    sd = {}
    for fn in sorted(os.listdir(path)):
        full = os.path.join(path, fn)
        if fn.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd.update(load_file(full))
        elif fn.endswith(".bin"):
            sd.update(torch.load(full, map_location="cpu"))
    return sd


class _Container(nn.Module):
    _engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype


class UNet2DConditionModel(_Container):
    def __init__(self, config=None, state_dict=None, seed=1234):
        super().__init__()
        self.config = dict(synthetic.SDXL_UNET_CONFIG)
        self.config.update(config or {})
        _build_tree(self, state_dict if state_dict is not None else synthetic.random_unet_state_dict(self.config, seed=seed))

    @classmethod
    def from_pretrained(cls, path=None, subfolder=None, config=None, seed=1234, **kwargs):
        d = os.path.join(path, subfolder) if (path and subfolder) else path
        if d and os.path.isdir(d):
            # SDXL's config.json calls the per-level head COUNT `attention_head_dim` (diffusers' historical naming)
            cfg = _load_config(d, synthetic.SDXL_UNET_CONFIG, renames=[("attention_head_dim", "num_attention_heads")])
            cfg.update(config or {})
            for k in ("transformer_layers_per_block", "num_attention_heads"):
                if k in cfg and not isinstance(cfg[k], tuple):
                    cfg[k] = (cfg[k],) * len(cfg.get("block_out_channels", synthetic.SDXL_UNET_CONFIG["block_out_channels"]))
            return cls(config=cfg, state_dict=_load_dir(d))
        synthetic.missing_checkpoint(d, "UNet2DConditionModel")
        return cls(config=config, seed=seed)

    def engine(self):
        if self._engine is None:
            self._engine = sdxl_engine.UNetEngine(self.state_dict(), self.config, self.device)
        return self._engine


class AutoencoderKL(_Container):
    def __init__(self, config=None, state_dict=None, seed=4321):
        super().__init__()
        self.config = dict(synthetic.SDXL_VAE_CONFIG)
        self.config.update(config or {})
        _build_tree(self, state_dict if state_dict is not None else synthetic.random_vae_decoder_state_dict(self.config, seed=seed))

    @classmethod
    def from_pretrained(cls, path=None, subfolder=None, config=None, seed=4321, **kwargs):
        d = os.path.join(path, subfolder) if (path and subfolder) else path
        if d and os.path.isdir(d):
            sd = {k: v for k, v in _load_dir(d).items() if k.startswith("decoder.") or k.startswith("post_quant_conv.")}
            cfg = _load_config(d, synthetic.SDXL_VAE_CONFIG)
            cfg.update(config or {})
            return cls(config=cfg, state_dict=sd)
        synthetic.missing_checkpoint(d, "AutoencoderKL")
        return cls(config=config, seed=seed)

    def engine(self):
        if self._engine is None:
            self._engine = sdxl_engine.VAEDecoderEngine(self.state_dict(), self.config, self.device)
        return self._engine


class EulerDiscreteScheduler:
    """Holds the SDXL scheduler constants; the schedule itself is computed in sdxl_engine.euler_schedule."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 timestep_spacing="leading", steps_offset=1, prediction_type="epsilon", **kwargs):
        assert beta_schedule == "scaled_linear" and timestep_spacing == "leading" and prediction_type == "epsilon", \
            "only the SDXL-base scheduler configuration is implemented"
        self.config = dict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           steps_offset=steps_offset)

    @classmethod
    def from_pretrained(cls, path=None, subfolder=None, **kwargs):
        d = os.path.join(path, subfolder, "scheduler_config.json") if (path and subfolder) else None
        if d and os.path.exists(d):
            with open(d) as f:
                c = json.load(f)
            return cls(**{k: v for k, v in c.items() if not k.startswith("_")})
        return cls()


class DDPMScheduler(EulerDiscreteScheduler):
    pass


class _PipeOutput:
    def __init__(self, images):
        self.images = images


class StableDiffusionXLPipeline:
    """prompt_embeds-only SDXL sampling (the way adapter_modules.py:455-466 calls it): CFG + Euler on the UNet
    engine, VAE decode, uint8 -> PIL."""

    def __init__(self, vae=None, unet=None, scheduler=None, tokenizer=None, tokenizer_2=None, text_encoder=None,
                 text_encoder_2=None, **kwargs):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.stats = {}

    @torch.no_grad()
    def __call__(self, prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                 negative_pooled_prompt_embeds=None, guidance_scale=5.0, num_inference_steps=50, generator=None,
                 height=1024, width=1024, output_type="pil", **kwargs):
        ue = self.unet.engine()
        S = ue.S
        assert height == width == 8 * S, f"engine was built for {8 * S}x{8 * S} images"
        dev = ue.dev
        ts, _ = sdxl_engine.euler_schedule(num_inference_steps)
        ctx = torch.cat([negative_prompt_embeds, prompt_embeds], 0)
        text = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], 0)
        tid = [[height, width, 0, 0, height, width]] * 2
        ue.set_conditioning(ctx, text, tid, ts)
        # initial noise: torch's CUDA Philox stream, same call the pipeline makes (randn_tensor, fp16)
        lat0 = torch.randn((1, 4, S, S), generator=generator, device=dev, dtype=torch.float16)
        lat = ue.sample(lat0, num_inference_steps, guidance=guidance_scale)
        img_u8, _ = self.vae.engine().decode(lat, S)
        self.stats = dict(unet_launches=ue.total_launches, vae_launches=self.vae.engine().launches)
        if output_type == "pil":
            from PIL import Image
            return _PipeOutput([Image.fromarray(img_u8.cpu().numpy())])
        return _PipeOutput([img_u8])


class StableDiffusionPipeline:  # imported by name only (adapter_modules.py:7); SD1.5 path is out of scope
    def __init__(self, *a, **k):
        raise NotImplementedError("SD1.5 pipelines are outside the SEED-Story hot path")


StableDiffusionXLInstructPix2PixPipeline = StableDiffusionPipeline
StableDiffusionInstructPix2PixPipeline = StableDiffusionPipeline
