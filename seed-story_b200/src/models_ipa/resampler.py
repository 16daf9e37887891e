"""Drop-in for the one resampler class of src/models_ipa/resampler.py that the shipped de-tokenizer config
instantiates: ResamplerXLV2 (reference :228-284; hydra target in
configs/detokenizer/detokenizer_sdxl_qwen_vit_adapted.yaml:3-13).  Parameter names follow the reference's
state_dict; forward() runs on the seedstory_b200 kernels (PerceiverAttention :47-76 and
AttentionPool2d :90-118 included)."""
import torch
from torch import nn

from seedstory import vision_engine


class _PerceiverParams(nn.Module):
    def __init__(self, dim, dim_head, heads):
        super().__init__()
        inner = dim_head * heads
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)


def _ff_params(dim, mult):
    inner = int(dim * mult)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner, bias=False), nn.GELU(),
                         nn.Linear(inner, dim, bias=False))


class _AttnPoolParams(nn.Module):
    def __init__(self, seq_len, embed_dim, num_heads, output_dim):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(seq_len + 1, embed_dim) / embed_dim ** 0.5)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        self.num_heads = num_heads


class ResamplerXLV2(nn.Module):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output1_dim=768,
                 output2_dim=1280, ff_mult=4):
        super().__init__()
        self.depth, self.heads = depth, heads
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.norm_out = nn.LayerNorm(dim)
        self.in_dim = dim
        self.out_dim = output1_dim + output2_dim
        self.layers = nn.ModuleList([nn.ModuleList([_PerceiverParams(dim, dim_head, heads), _ff_params(dim, ff_mult)])
                                     for _ in range(depth)])
        self.unet_proj_1 = nn.Linear(dim, output1_dim)
        self.unet_proj_2 = nn.Linear(dim, output2_dim)
        self.unet_attnpool = _AttnPoolParams(num_queries, dim, heads, output2_dim)
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def forward(self, x, pooled_text_embeds=None):
        if self._engine is None:
            self._engine = vision_engine.ResamplerXLV2Engine(self.state_dict(), self.depth, self.heads,
                                                             self.latents.device)
        p, q = self._engine(x)
        return p.to(x.dtype), q.to(x.dtype)
