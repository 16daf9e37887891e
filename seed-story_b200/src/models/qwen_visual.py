"""Drop-in for the reference's src/models/qwen_visual.py (same import path, class names, constructor kwargs and
state_dict keys — SURVEY.md §8b / Appendix A); forward() runs on the seedstory_b200 CUDA kernels.

  VisionTransformerWithAttnPool  <- reference qwen_visual.py:321-422   (hydra target in configs/visual_tokenizer/qwen_vitg_448.yaml)
  Resampler                      <- reference qwen_visual.py:95-153     (agent input/output resamplers, configs/clm_models/agent_7b_sft.yaml)

The nn.Module tree below only HOLDS parameters (so checkpoints written for the reference load unchanged);
no torch arithmetic runs in forward().  There is no CPU path: calling forward without a CUDA device raises.
"""
import math
from functools import partial

import numpy as np
import torch
from torch import nn

from seedstory import vision_engine


def get_2d_sincos_pos_embed(embed_dim, grid_size):
    """Fixed 2-D sin/cos table of the Resampler queries (MAE recipe: half the channels encode one grid axis,
    half the other; each half is [sin | cos] over 1/10000^(2i/d)).  Same values as reference qwen_visual.py:45-92."""
    coords = np.arange(grid_size, dtype=np.float32)
    gw, gh = np.meshgrid(coords, coords)  # w varies fastest

    def axis_table(dim, pos):
        freq = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float32) / (dim / 2.0))
        ang = np.einsum("m,d->md", pos.reshape(-1), freq)
        return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)
    return np.concatenate([axis_table(embed_dim // 2, gw), axis_table(embed_dim // 2, gh)], axis=1)


class _EngineMixin:
    """Lazily packs parameters into kernel layouts; any .to()/.half()/load_state_dict() invalidates the pack."""
    _engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)


class Resampler(_EngineMixin, nn.Module):
    def __init__(self, grid_size, embed_dim, num_heads, kv_dim=None, norm_layer=nn.LayerNorm):
        super().__init__()
        self.num_queries = grid_size ** 2
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.pos_embed = nn.Parameter(torch.from_numpy(get_2d_sincos_pos_embed(embed_dim, grid_size)).float(),
                                      requires_grad=False)
        self.query = nn.Parameter(torch.zeros(self.num_queries, embed_dim))
        nn.init.trunc_normal_(self.query, std=.02)
        if kv_dim is not None and kv_dim != embed_dim:
            self.kv_proj = nn.Linear(kv_dim, embed_dim, bias=False)
            nn.init.trunc_normal_(self.kv_proj.weight, std=.02)
            self.out_dim = kv_dim
        else:
            self.kv_proj = nn.Identity()
            self.out_dim = embed_dim
        self.attn = nn.MultiheadAttention(embed_dim, num_heads)  # parameter container only
        self.ln_q = norm_layer(embed_dim)
        self.ln_kv = norm_layer(embed_dim)

    def forward(self, x, attn_mask=None):
        assert attn_mask is None, "attn_mask is never passed on the inference path"
        if self._engine is None or self._engine.kv_len != x.shape[1]:
            self._engine = vision_engine.ResamplerEngine(self.state_dict(), self.num_heads, x.shape[1], x.device,
                                                         eps=self.ln_kv.eps)
        return self._engine(x.to(torch.float16)).to(x.dtype)


class _AttnParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.in_proj = nn.Linear(d, 3 * d)   # rows per head: [q | k | v] interleaved (Appendix A)
        self.out_proj = nn.Linear(d, d)


class _Block(nn.Module):
    def __init__(self, d, mlp_width, norm_layer):
        super().__init__()
        self.ln_1 = norm_layer(d)
        self.ln_2 = norm_layer(d)
        self.attn = _AttnParams(d)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(d, mlp_width))
        self.mlp.add_module("gelu", nn.GELU())
        self.mlp.add_module("c_proj", nn.Linear(mlp_width, d))


class _Blocks(nn.Module):
    def __init__(self, width, layers, mlp_width, norm_layer):
        super().__init__()
        self.resblocks = nn.ModuleList([_Block(width, mlp_width, norm_layer) for _ in range(layers)])


class VisionTransformerWithAttnPool(_EngineMixin, nn.Module):
    def __init__(self, image_size, patch_size, width, layers, heads, mlp_ratio, n_queries=256, output_dim=512,
                 **kwargs):
        super().__init__()
        self.image_size, self.patch_size = (image_size, image_size), (patch_size, patch_size)
        self.grid_size = (image_size // patch_size, image_size // patch_size)
        self.output_dim, self.width, self.layers, self.heads = output_dim, width, layers, heads
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        self.positional_embedding = nn.Parameter(width ** -0.5 * torch.randn(256, width))
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.ln_pre = norm_layer(width)
        self.transformer = _Blocks(width, layers, int(width * mlp_ratio), norm_layer)
        self.attn_pool = Resampler(grid_size=int(math.sqrt(n_queries)), embed_dim=output_dim,
                                   num_heads=output_dim // 128, kv_dim=width, norm_layer=norm_layer)
        self.ln_post = norm_layer(output_dim)
        self.proj = nn.Parameter(output_dim ** -0.5 * torch.randn(output_dim, output_dim))

    def forward(self, x):
        if self._engine is None:
            dev = self.proj.device
            self._engine = vision_engine.ViTEngine(self.state_dict(), self.image_size[0], self.patch_size[0],
                                                   self.width, self.layers, self.heads, dev)
        return self._engine(x).to(self.proj.dtype)

    @classmethod
    def from_pretrained(cls, pretrained_model_path=None, **kwargs):
        model = cls(**kwargs)
        if pretrained_model_path is not None:
            import os
            if os.path.exists(pretrained_model_path):
                ckpt = torch.load(pretrained_model_path, map_location="cpu")
                missing, unexpected = model.load_state_dict(ckpt, strict=False)
                print("Load ckpt of qwen visual encoder")
                print("missing keys: ", len(missing), "unexpected keys:", len(unexpected))
            else:
                from seedstory.synthetic import missing_checkpoint
                missing_checkpoint(pretrained_model_path, "VisionTransformerWithAttnPool")
        return model
