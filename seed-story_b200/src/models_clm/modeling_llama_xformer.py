"""Drop-in for the reference's src/models_clm/modeling_llama_xformer.py at the boundary the scripts use
(SURVEY.md §8b): `LlamaForCausalLM.from_pretrained(path, torch_dtype=, low_cpu_mem_usage=)`, the attribute
trio `use_kv_cache_head / kv_cache_head / past_key_values` (reference :676-678), `get_input_embeddings`,
`resize_token_embeddings`, and the HF state_dict key layout (`model.layers.N.self_attn.q_proj.weight`, …).

The module tree only holds parameters.  All arithmetic of LlamaModel.forward / LlamaAttention / LlamaMLP /
LlamaRMSNorm (reference :97-368, :532-666) runs in seedstory.llama_engine on the CUDA kernels.
"""
import json
import os

import torch
from torch import nn

from seedstory import llama_engine


class LlamaRMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps


class _Attn(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.q_proj = nn.Linear(h, h, bias=False)
        self.k_proj = nn.Linear(h, h, bias=False)
        self.v_proj = nn.Linear(h, h, bias=False)
        self.o_proj = nn.Linear(h, h, bias=False)


class _MLP(nn.Module):
    def __init__(self, h, i):
        super().__init__()
        self.gate_proj = nn.Linear(h, i, bias=False)
        self.down_proj = nn.Linear(i, h, bias=False)
        self.up_proj = nn.Linear(h, i, bias=False)


class LlamaDecoderLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.self_attn = _Attn(cfg["hidden_size"])
        self.mlp = _MLP(cfg["hidden_size"], cfg["intermediate_size"])
        self.input_layernorm = LlamaRMSNorm(cfg["hidden_size"], cfg["rms_norm_eps"])
        self.post_attention_layernorm = LlamaRMSNorm(cfg["hidden_size"], cfg["rms_norm_eps"])


class _TokenEmbedding(nn.Embedding):
    """Embedding lookup on the CUDA gather kernel (reference call site: src/models_clm/models.py:127)."""
    _engine_ref = None

    def forward(self, input_ids):
        from seedstory import ops
        ids = input_ids.reshape(-1).to(self.weight.device, torch.int32).contiguous()
        out = torch.empty((ids.numel(), self.weight.shape[1]), dtype=self.weight.dtype, device=self.weight.device)
        ops.gather_rows(self.weight, ids, out)
        return out.view(*input_ids.shape, -1)


class LlamaModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embed_tokens = _TokenEmbedding(cfg["vocab_size"], cfg["hidden_size"])
        self.layers = nn.ModuleList([LlamaDecoderLayer(cfg) for _ in range(cfg["num_hidden_layers"])])
        self.norm = LlamaRMSNorm(cfg["hidden_size"], cfg["rms_norm_eps"])


DEFAULT_7B = dict(hidden_size=4096, intermediate_size=11008, num_attention_heads=32, num_hidden_layers=32,
                  vocab_size=32000, rms_norm_eps=1e-5, max_position_embeddings=4096)


class LlamaForCausalLM(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        cfg = dict(DEFAULT_7B)
        cfg.update(config or {})
        self.config = cfg
        self.model = LlamaModel(cfg)
        self.lm_head = nn.Linear(cfg["hidden_size"], cfg["vocab_size"], bias=False)
        self.past_key_values = None
        self.kv_cache_head = None
        self.use_kv_cache_head = True
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.normal_(p, std=0.02)

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def get_output_embeddings(self):
        return self.lm_head

    def resize_token_embeddings(self, n):
        old_e, old_h = self.model.embed_tokens, self.lm_head
        if n == old_e.weight.shape[0]:
            return old_e
        dev, dt = old_e.weight.device, old_e.weight.dtype
        new_e = _TokenEmbedding(n, old_e.weight.shape[1]).to(dev, dt)
        new_h = nn.Linear(old_h.weight.shape[1], n, bias=False).to(dev, dt)
        k = min(n, old_e.weight.shape[0])
        with torch.no_grad():
            nn.init.normal_(new_e.weight, std=0.02)
            nn.init.normal_(new_h.weight, std=0.02)
            new_e.weight[:k] = old_e.weight[:k]
            new_h.weight[:k] = old_h.weight[:k]
        self.model.embed_tokens, self.lm_head = new_e, new_h
        self.config["vocab_size"] = n
        return new_e

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, torch_dtype=None, low_cpu_mem_usage=False,
                        config=None, device=None, **kwargs):
        """Loads an HF Llama directory (config.json + *.safetensors / pytorch_model*.bin) when it exists; otherwise
        builds the architecture with seeded random weights (no checkpoints exist offline — SURVEY.md §8d)."""
        path = pretrained_model_name_or_path
        cfg = dict(config or {})
        have = path is not None and os.path.isdir(path) and os.path.exists(os.path.join(path, "config.json"))
        if have:
            with open(os.path.join(path, "config.json")) as f:
                hf = json.load(f)
            for k in DEFAULT_7B:
                if k in hf:
                    cfg.setdefault(k, hf[k])
        ctx = torch.device(device) if device is not None else torch.device("cpu")
        with ctx:
            model = cls(cfg)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        if have:
            sd = {}
            for fn in sorted(os.listdir(path)):
                if fn.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd.update(load_file(os.path.join(path, fn)))
                elif fn.startswith("pytorch_model") and fn.endswith(".bin"):
                    sd.update(torch.load(os.path.join(path, fn), map_location="cpu"))
            missing, unexpected = model.load_state_dict(sd, strict=False)
            missing = [m for m in missing if "rotary_emb" not in m]
            if missing:
                raise RuntimeError(f"Llama checkpoint at {path} does not cover: {missing[:8]} …")
        else:
            from seedstory.synthetic import missing_checkpoint
            missing_checkpoint(path, "LlamaForCausalLM")
        return model

    def engine_config(self):
        c = self.config
        return llama_engine.LlamaConfig(hidden=c["hidden_size"], inter=c["intermediate_size"],
                                        heads=c["num_attention_heads"], layers=c["num_hidden_layers"],
                                        vocab=c["vocab_size"], eps=c["rms_norm_eps"],
                                        max_pos=max(4096, c.get("max_position_embeddings", 4096)))
