"""ORACLE — test infrastructure only (never imported by the product path).

Functional CPU/torch restatement of the visual side of the hot path, driven by state_dicts laid out exactly
as the reference modules' (SURVEY.md Appendix A):
  * VisionTransformerWithAttnPool.forward   src/models/qwen_visual.py:376-399
      get_abs_pos :23-39, VisualAttention :184-235, VisualAttentionBlock :275-287
  * Resampler.forward (agent input/output resamplers and the ViT attn-pool)   qwen_visual.py:138-150
  * ResamplerXLV2.forward + PerceiverAttention + AttentionPool2d   src/models_ipa/resampler.py:266-284, 47-76, 90-118
  * SDXLAdapter.get_image_embeds glue   src/models_ipa/adapter_modules.py:387-428

Pinned by oracle/pin_against_reference.py against the reference modules themselves (they import and run
on CPU in the build container); frozen vectors live in tests/golden/.
"""
import math

import torch
import torch.nn.functional as F


def abs_pos(pos, tgt_len):
    """qwen_visual.py:23-39 — bicubic resize of a square grid of positional embeddings, done in fp32."""
    src = int(math.sqrt(pos.shape[0]))
    tgt = int(math.sqrt(tgt_len))
    if src == tgt:
        return pos
    grid = pos.float().reshape(1, src, src, -1).permute(0, 3, 1, 2)
    out = F.interpolate(grid, size=(tgt, tgt), mode="bicubic", align_corners=False)
    return out.permute(0, 2, 3, 1).flatten(0, 2).to(pos.dtype)


def sincos_2d(embed_dim, grid_size):
    """qwen_visual.py:45-92 (numpy there; same arithmetic in float32/float64 mix reproduced with numpy)."""
    import numpy as np
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])

    def one_d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float32)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    emb = np.concatenate([one_d(embed_dim // 2, grid[0]), one_d(embed_dim // 2, grid[1])], axis=1)
    return torch.from_numpy(emb).float()


def mha_need_weights(q, k, v, in_w, in_b, out_w, out_b, heads):
    """nn.MultiheadAttention forward with default need_weights=True (seq-first inputs [L, N, E]):
    packed in-projection split Q;K;V, q scaled by 1/sqrt(d) before bmm, softmax in the activation dtype,
    out-projection (torch/nn/functional.py multi_head_attention_forward, explicit-math branch)."""
    Lq, N, E = q.shape
    Lk = k.shape[0]
    d = E // heads
    wq, wk, wv = in_w.chunk(3, dim=0)
    bq, bk, bv = in_b.chunk(3, dim=0)
    qh = F.linear(q, wq, bq).reshape(Lq, N * heads, d).transpose(0, 1)
    kh = F.linear(k, wk, bk).reshape(Lk, N * heads, d).transpose(0, 1)
    vh = F.linear(v, wv, bv).reshape(Lk, N * heads, d).transpose(0, 1)
    qh = qh * math.sqrt(1.0 / float(d))
    w = torch.bmm(qh, kh.transpose(-2, -1))
    w = torch.softmax(w, dim=-1)
    o = torch.bmm(w, vh).transpose(0, 1).contiguous().view(Lq * N, E)
    return F.linear(o, out_w, out_b).view(Lq, N, E)


def resampler(sd, x, heads, eps=1e-5, prefix=""):
    """Resampler.forward (qwen_visual.py:138-150). x [N, L, kv_dim] -> [N, nq, E]."""
    g = lambda k: sd[prefix + k]
    pos_k = abs_pos(g("pos_embed"), x.shape[1])
    if prefix + "kv_proj.weight" in sd:
        x = F.linear(x, g("kv_proj.weight"))
    E = x.shape[-1]
    x = F.layer_norm(x, (E,), g("ln_kv.weight"), g("ln_kv.bias"), eps).permute(1, 0, 2)
    N = x.shape[1]
    q = F.layer_norm(g("query"), (E,), g("ln_q.weight"), g("ln_q.bias"), eps)
    qq = q.unsqueeze(1).repeat(1, N, 1) + g("pos_embed").unsqueeze(1)
    out = mha_need_weights(qq, x + pos_k.unsqueeze(1), x, g("attn.in_proj_weight"), g("attn.in_proj_bias"),
                           g("attn.out_proj.weight"), g("attn.out_proj.bias"), heads)
    return out.permute(1, 0, 2)


def vit_attention(sd, pre, x, heads):
    """VisualAttention.forward (qwen_visual.py:184-235); x [L, N, W]; per-head interleaved q|k|v rows."""
    L, N, W = x.shape
    hd = W // heads
    mixed = F.linear(x, sd[pre + "in_proj.weight"], sd[pre + "in_proj.bias"]).view(L, N, heads, 3 * hd)
    q, k, v = mixed.split(hd, dim=-1)
    q = q.reshape(L, N * heads, hd).transpose(0, 1)
    k = k.reshape(L, N * heads, hd).transpose(0, 1)
    v = v.reshape(L, N * heads, hd).transpose(0, 1)
    p = torch.bmm(q / math.sqrt(hd), k.transpose(-2, -1)).softmax(dim=-1)
    ctx = torch.bmm(p, v).view(N, heads, L, hd).permute(2, 0, 1, 3).reshape(L, N, W)
    return F.linear(ctx, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def vit_forward(sd, img, heads, layers, patch, eps=1e-6):
    """VisionTransformerWithAttnPool.forward (qwen_visual.py:376-399). img [N,3,S,S] -> [N,256,out_dim]."""
    W = sd["conv1.weight"].shape[0]
    x = F.conv2d(img, sd["conv1.weight"], stride=patch)
    x = x.reshape(x.shape[0], W, -1).permute(0, 2, 1)
    x = x + abs_pos(sd["positional_embedding"], x.shape[1])
    x = F.layer_norm(x, (W,), sd["ln_pre.weight"], sd["ln_pre.bias"], eps)
    x = x.permute(1, 0, 2)
    for i in range(layers):
        pre = f"transformer.resblocks.{i}."
        y = F.layer_norm(x, (W,), sd[pre + "ln_1.weight"], sd[pre + "ln_1.bias"], eps)
        x = x + vit_attention(sd, pre + "attn.", y, heads)
        y = F.layer_norm(x, (W,), sd[pre + "ln_2.weight"], sd[pre + "ln_2.bias"], eps)
        y = F.linear(F.gelu(F.linear(y, sd[pre + "mlp.c_fc.weight"], sd[pre + "mlp.c_fc.bias"])),
                     sd[pre + "mlp.c_proj.weight"], sd[pre + "mlp.c_proj.bias"])
        x = x + y
    x = x.permute(1, 0, 2)
    E = sd["proj"].shape[0]
    x = resampler(sd, x, E // 128, eps=eps, prefix="attn_pool.")
    x = F.layer_norm(x, (E,), sd["ln_post.weight"], sd["ln_post.bias"], eps)
    return x @ sd["proj"]


# ---- ResamplerXLV2 ---------------------------------------------------------------------------
def perceiver_attention(sd, pre, x, latents, heads):
    """resampler.py:47-76."""
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    latents = F.layer_norm(latents, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    b, l, _ = latents.shape
    q = F.linear(latents, sd[pre + "to_q.weight"])
    kv = F.linear(torch.cat((x, latents), dim=-2), sd[pre + "to_kv.weight"])
    k, v = kv.chunk(2, dim=-1)

    def sh(t):
        return t.view(t.shape[0], t.shape[1], heads, -1).transpose(1, 2)
    q, k, v = sh(q), sh(k), sh(v)
    dh = q.shape[-1]
    scale = 1 / math.sqrt(math.sqrt(dh))
    w = (q * scale) @ (k * scale).transpose(-2, -1)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    out = (w @ v).permute(0, 2, 1, 3).reshape(b, l, -1)
    return F.linear(out, sd[pre + "to_out.weight"])


def attention_pool_2d(sd, pre, x, heads):
    """resampler.py:90-118 — prepend mean token, add positional embedding, MHA (need_weights=False), token 0."""
    x = x.permute(1, 0, 2)
    x = torch.cat([x.mean(dim=0, keepdim=True), x], dim=0)
    x = x + sd[pre + "positional_embedding"][:, None, :].to(x.dtype)
    L, N, E = x.shape
    d = E // heads
    q = F.linear(x, sd[pre + "q_proj.weight"], sd[pre + "q_proj.bias"])
    k = F.linear(x, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"])
    v = F.linear(x, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"])

    def sh(t):
        return t.reshape(L, N * heads, d).transpose(0, 1)
    o = F.scaled_dot_product_attention(sh(q), sh(k), sh(v))
    o = o.transpose(0, 1).reshape(L, N, E)
    o = F.linear(o, sd[pre + "c_proj.weight"], sd[pre + "c_proj.bias"])
    return o[0]


def resampler_xl_v2(sd, x, depth, heads):
    """ResamplerXLV2.forward (resampler.py:266-284). x [B,256,4096] -> ([B,64,2048], [B,1280])."""
    latents = sd["latents"].repeat(x.size(0), 1, 1)
    x = F.normalize(x)  # p=2 over dim=1 — the token axis (resampler.py:269)
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    D = latents.shape[-1]
    for i in range(depth):
        latents = perceiver_attention(sd, f"layers.{i}.0.", x, latents, heads) + latents
        y = F.layer_norm(latents, (D,), sd[f"layers.{i}.1.0.weight"], sd[f"layers.{i}.1.0.bias"])
        y = F.linear(F.gelu(F.linear(y, sd[f"layers.{i}.1.1.weight"])), sd[f"layers.{i}.1.3.weight"])
        latents = y + latents
    hid = F.layer_norm(latents, (D,), sd["norm_out.weight"], sd["norm_out.bias"])
    e1 = F.linear(hid, sd["unet_proj_1.weight"], sd["unet_proj_1.bias"])
    e2 = F.linear(hid, sd["unet_proj_2.weight"], sd["unet_proj_2.bias"])
    pooled = attention_pool_2d(sd, "unet_attnpool.", hid, heads)
    return torch.cat([e1, e2], dim=-1), pooled
