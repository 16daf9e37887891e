"""Qwen ViT-G + attention-pool, the agent Resamplers and ResamplerXLV2 over the seedstory_b200 kernels.

Replaces the arithmetic of
  src/models/qwen_visual.py      VisionTransformerWithAttnPool.forward :376-399, VisualAttention :184-235,
                                 VisualAttentionBlock :275-287, Resampler.forward :138-150, get_abs_pos :23-39
  src/models_ipa/resampler.py    ResamplerXLV2.forward :266-284, PerceiverAttention :47-76, AttentionPool2d :90-118
State dicts keep the reference's key names; `pack_*` repacks them once into kernel
layouts (head_dim 104 zero-padded to 128, fused/transposed projections, interpolated positional tables —
input-independent constants computed at load).
"""
import math

import torch
import torch.nn.functional as F

from . import ops


def _interp_pos(pos, tgt_len):
    """get_abs_pos (qwen_visual.py:23-39): bicubic, align_corners=False, fp32, cast back. Load-time constant."""
    src = int(math.sqrt(pos.shape[0]))
    tgt = int(math.sqrt(tgt_len))
    if src == tgt:
        return pos
    g = pos.float().reshape(1, src, src, -1).permute(0, 3, 1, 2)
    out = F.interpolate(g, size=(tgt, tgt), mode="bicubic", align_corners=False)
    return out.permute(0, 2, 3, 1).flatten(0, 2).to(pos.dtype)


def _dev16(t, dev):
    return t.detach().to(dev, torch.float16).contiguous()


class ResamplerEngine:
    """One cross-attention block: learnable queries over LN(kv_proj(x)) (+ positional tables)."""

    def __init__(self, sd, heads, kv_len, device, eps=1e-5, prefix=""):
        g = lambda k: sd[prefix + k]
        self.dev, self.heads, self.eps, self.kv_len = device, heads, eps, kv_len
        E = g("query").shape[1]
        self.E, self.nq = E, g("query").shape[0]
        self.kv_proj = _dev16(g("kv_proj.weight"), device) if (prefix + "kv_proj.weight") in sd else None
        self.ln_kv = (_dev16(g("ln_kv.weight"), device), _dev16(g("ln_kv.bias"), device))
        in_w, in_b = g("attn.in_proj_weight"), g("attn.in_proj_bias")
        wq, wk, wv = in_w.chunk(3, 0)
        bq, bk, bv = in_b.chunk(3, 0)
        self.wk, self.bk = _dev16(wk, device), _dev16(bk, device)
        self.wv, self.bv = _dev16(wv, device), _dev16(bv, device)
        self.wo, self.bo = _dev16(g("attn.out_proj.weight"), device), _dev16(g("attn.out_proj.bias"), device)
        # keys' positional table (interpolated to kv_len) and the input-independent query projection
        self.pos_k = _dev16(_interp_pos(g("pos_embed").detach().cpu(), kv_len), device)
        query = _dev16(g("query"), device)
        pos_q = _dev16(g("pos_embed"), device)
        _, q_in = ops.layernorm(query, _dev16(g("ln_q.weight"), device), _dev16(g("ln_q.bias"), device), eps, add=pos_q)
        # w_const=False: the fp16 copy of wq is produced on this stream right before the launch
        self.Q = ops.gemm(q_in, _dev16(wq, device), bias=_dev16(bq, device), w_const=False)  # [nq, E], constant
        self.scale = 1.0 / math.sqrt(E // heads)
        ops.register_const_tree(self)          # load-time weights: GEMMs may prefetch them before their PDL wait
        torch.cuda.current_stream().synchronize()

    def __call__(self, x):
        """x [N, L, kv_dim] fp16 -> [N, nq, E]."""
        N, L, _ = x.shape
        assert L == self.kv_len, (L, self.kv_len)
        x2 = x.reshape(N * L, -1)
        if self.kv_proj is not None:
            x2 = ops.gemm(x2, self.kv_proj)
        kn, kn_pos = ops.layernorm(x2, self.ln_kv[0], self.ln_kv[1], self.eps, add=self.pos_k)
        K = ops.gemm(kn_pos, self.wk, bias=self.bk)
        V = ops.gemm(kn, self.wv, bias=self.bv)
        E, H = self.E, self.heads
        D = E // H
        attn = torch.empty((N, self.nq, E), dtype=torch.float16, device=self.dev)
        ops.fmha(self.Q, K, V, attn, N, H, self.nq, L, D, (0, E, D), (L * E, E, D), (L * E, E, D),
                 (self.nq * E, E, D), self.scale)
        out = ops.gemm(attn.view(N * self.nq, E), self.wo, bias=self.bo)
        return out.view(N, self.nq, E)


class ViTEngine:
    def __init__(self, sd, image_size, patch_size, width, layers, heads, device, eps=1e-6):
        self.dev, self.width, self.layers, self.heads, self.eps = device, width, layers, heads, eps
        self.S, self.P = image_size, patch_size
        self.G = image_size // patch_size
        T = self.G * self.G
        self.T = T
        hd = width // heads
        self.hd = hd
        self.hdp = 64 if hd <= 64 else 128
        assert hd <= 128
        kraw = 3 * patch_size * patch_size
        self.kpad = (kraw + 63) // 64 * 64
        cw = sd["conv1.weight"].detach().reshape(width, kraw)
        cwp = torch.zeros(width, self.kpad, dtype=cw.dtype)
        cwp[:, :kraw] = cw
        self.conv_w = _dev16(cwp, device)
        self.pos = _dev16(_interp_pos(sd["positional_embedding"].detach().cpu(), T), device)
        self.ln_pre = (_dev16(sd["ln_pre.weight"], device), _dev16(sd["ln_pre.bias"], device))
        self.blocks = []
        hp = self.hdp
        for i in range(layers):
            pre = f"transformer.resblocks.{i}."
            w_in = sd[pre + "attn.in_proj.weight"].detach().view(heads, 3, hd, width)
            b_in = sd[pre + "attn.in_proj.bias"].detach().view(heads, 3, hd)
            w_pack = torch.zeros(heads, 3, hp, width, dtype=w_in.dtype)
            b_pack = torch.zeros(heads, 3, hp, dtype=b_in.dtype)
            w_pack[:, :, :hd] = w_in
            b_pack[:, :, :hd] = b_in
            w_out = sd[pre + "attn.out_proj.weight"].detach().view(width, heads, hd)
            w_out_p = torch.zeros(width, heads, hp, dtype=w_out.dtype)
            w_out_p[:, :, :hd] = w_out
            self.blocks.append(dict(
                ln1=(_dev16(sd[pre + "ln_1.weight"], device), _dev16(sd[pre + "ln_1.bias"], device)),
                ln2=(_dev16(sd[pre + "ln_2.weight"], device), _dev16(sd[pre + "ln_2.bias"], device)),
                w_in=_dev16(w_pack.view(heads * 3 * hp, width), device), b_in=_dev16(b_pack.view(-1), device),
                w_out=_dev16(w_out_p.view(width, heads * hp), device), b_out=_dev16(sd[pre + "attn.out_proj.bias"], device),
                w_fc=_dev16(sd[pre + "mlp.c_fc.weight"], device), b_fc=_dev16(sd[pre + "mlp.c_fc.bias"], device),
                w_proj=_dev16(sd[pre + "mlp.c_proj.weight"], device), b_proj=_dev16(sd[pre + "mlp.c_proj.bias"], device)))
        E = sd["proj"].shape[0]
        self.pool = ResamplerEngine(sd, E // 128, T, device, eps=eps, prefix="attn_pool.")
        self.ln_post = (_dev16(sd["ln_post.weight"], device), _dev16(sd["ln_post.bias"], device))
        self.proj_t = _dev16(sd["proj"].detach().t(), device)  # x @ proj  ==  x @ (proj^T)^T
        self.scale = 1.0 / math.sqrt(hd)
        ops.register_const_tree(self)          # load-time weights: GEMMs may prefetch them before their PDL wait
        torch.cuda.current_stream().synchronize()

    def __call__(self, img):
        """img [N,3,S,S] fp16 (CLIP-normalised) -> [N, nq, out_dim]."""
        N = img.shape[0]
        T, W, H, hp = self.T, self.width, self.heads, self.hdp
        patches = ops.im2col_patch(img.to(self.dev, torch.float16), self.P, self.kpad)
        x = ops.gemm(patches, self.conv_w)
        ops.add_bcast(x, self.pos, out=x)
        x = ops.layernorm(x, self.ln_pre[0], self.ln_pre[1], self.eps)
        qkv = torch.empty((N * T, H * 3 * hp), dtype=torch.float16, device=self.dev)
        attn = torch.empty((N * T, H * hp), dtype=torch.float16, device=self.dev)
        y = torch.empty_like(x)
        for blk in self.blocks:
            ops.layernorm(x, blk["ln1"][0], blk["ln1"][1], self.eps, out=y)
            ops.gemm(y, blk["w_in"], bias=blk["b_in"], out=qkv)
            ld = H * 3 * hp
            ops.fmha(qkv, qkv[:, hp:], qkv[:, 2 * hp:], attn, N, H, T, T, hp, (T * ld, ld, 3 * hp), (T * ld, ld, 3 * hp),
                     (T * ld, ld, 3 * hp), (T * H * hp, H * hp, hp), self.scale)
            ops.gemm(attn, blk["w_out"], bias=blk["b_out"], residual=x, out=x)
            ops.layernorm(x, blk["ln2"][0], blk["ln2"][1], self.eps, out=y)
            hdn = ops.gemm(y, blk["w_fc"], bias=blk["b_fc"], act=ops.ACT_GELU)
            ops.gemm(hdn, blk["w_proj"], bias=blk["b_proj"], residual=x, out=x)
        pooled = self.pool(x.view(N, T, W))
        E = pooled.shape[-1]
        out = ops.layernorm(pooled.view(-1, E), self.ln_post[0], self.ln_post[1], self.eps)
        return ops.gemm(out, self.proj_t).view(N, -1, E)


class ResamplerXLV2Engine:
    def __init__(self, sd, depth, heads, device):
        d = lambda k: _dev16(sd[k], device)
        self.dev, self.depth, self.heads = device, depth, heads
        self.latents = d("latents")[0]  # [nq, dim]
        self.nq, self.dim = self.latents.shape
        self.proj_in = (d("proj_in.weight"), d("proj_in.bias"))
        self.layers = []
        for i in range(depth):
            a, f = f"layers.{i}.0.", f"layers.{i}.1."
            self.layers.append(dict(
                n1=(d(a + "norm1.weight"), d(a + "norm1.bias")), n2=(d(a + "norm2.weight"), d(a + "norm2.bias")),
                to_q=d(a + "to_q.weight"), to_kv=d(a + "to_kv.weight"), to_out=d(a + "to_out.weight"),
                ff_ln=(d(f + "0.weight"), d(f + "0.bias")), ff1=d(f + "1.weight"), ff2=d(f + "3.weight")))
        self.norm_out = (d("norm_out.weight"), d("norm_out.bias"))
        self.p1 = (d("unet_proj_1.weight"), d("unet_proj_1.bias"))
        self.p2 = (d("unet_proj_2.weight"), d("unet_proj_2.bias"))
        ap = "unet_attnpool."
        self.ap_pos = d(ap + "positional_embedding")
        self.ap_q = (d(ap + "q_proj.weight"), d(ap + "q_proj.bias"))
        self.ap_k = (d(ap + "k_proj.weight"), d(ap + "k_proj.bias"))
        self.ap_v = (d(ap + "v_proj.weight"), d(ap + "v_proj.bias"))
        self.ap_c = (d(ap + "c_proj.weight"), d(ap + "c_proj.bias"))
        self.dh = self.to_q_dim() // heads
        ops.register_const_tree(self)          # load-time weights: GEMMs may prefetch them before their PDL wait
        torch.cuda.current_stream().synchronize()

    def to_q_dim(self):
        return self.layers[0]["to_q"].shape[0]

    def __call__(self, x):
        """x [B, T, emb] fp16 -> (prompt_embeds [B, nq, o1+o2], pooled [B, o2])."""
        B, T, _ = x.shape
        dim, nq, H = self.dim, self.nq, self.heads
        inner = self.to_q_dim()
        x = ops.l2norm_tokens(x.to(self.dev, torch.float16))
        x = ops.gemm(x.view(B * T, -1), self.proj_in[0], bias=self.proj_in[1])  # [B*T, dim]
        lat = self.latents.unsqueeze(0).repeat(B, 1, 1).reshape(B * nq, dim).contiguous()
        kv_in = torch.empty((B, T + nq, dim), dtype=torch.float16, device=self.dev)
        lat_n = torch.empty((B * nq, dim), dtype=torch.float16, device=self.dev)
        scale = 1.0 / math.sqrt(self.dh)
        for L in self.layers:
            for b in range(B):
                ops.layernorm(x[b * T:(b + 1) * T], L["n1"][0], L["n1"][1], 1e-5, out=kv_in[b, :T])
                ops.layernorm(lat[b * nq:(b + 1) * nq], L["n2"][0], L["n2"][1], 1e-5, out=kv_in[b, T:])
                lat_n[b * nq:(b + 1) * nq].copy_(kv_in[b, T:])
            q = ops.gemm(lat_n, L["to_q"])
            kv = ops.gemm(kv_in.view(B * (T + nq), dim), L["to_kv"])  # [B*(T+nq), 2*inner]: k | v
            o = torch.empty((B * nq, inner), dtype=torch.float16, device=self.dev)
            Lk = T + nq
            ops.fmha(q, kv, kv[:, inner:], o, B, H, nq, Lk, self.dh, (nq * inner, inner, self.dh),
                     (Lk * 2 * inner, 2 * inner, self.dh), (Lk * 2 * inner, 2 * inner, self.dh),
                     (nq * inner, inner, self.dh), scale)
            ops.gemm(o, L["to_out"], residual=lat, out=lat)
            y = ops.layernorm(lat, L["ff_ln"][0], L["ff_ln"][1], 1e-5)
            y = ops.gemm(y, L["ff1"], act=ops.ACT_GELU)
            ops.gemm(y, L["ff2"], residual=lat, out=lat)
        hid = ops.layernorm(lat, self.norm_out[0], self.norm_out[1], 1e-5)  # [B*nq, dim]
        o1, o2 = self.p1[0].shape[0], self.p2[0].shape[0]
        prompt = torch.empty((B * nq, o1 + o2), dtype=torch.float16, device=self.dev)
        ops.gemm(hid, self.p1[0], bias=self.p1[1], out=prompt[:, :o1])
        ops.gemm(hid, self.p2[0], bias=self.p2[1], out=prompt[:, o1:])
        # AttentionPool2d: [mean; tokens] + pos, MHA, keep token 0
        seq = torch.empty((B, nq + 1, dim), dtype=torch.float16, device=self.dev)
        seq[:, 0].copy_(ops.mean_tokens(hid.view(B, nq, dim)))
        seq[:, 1:].copy_(hid.view(B, nq, dim))
        ops.add_bcast(seq, self.ap_pos, out=seq)
        s2 = seq.view(B * (nq + 1), dim)
        q0 = ops.gemm(seq[:, 0].contiguous(), self.ap_q[0], bias=self.ap_q[1])  # [B, dim]
        K = ops.gemm(s2, self.ap_k[0], bias=self.ap_k[1])
        V = ops.gemm(s2, self.ap_v[0], bias=self.ap_v[1])
        dh = dim // H
        o = torch.empty((B, dim), dtype=torch.float16, device=self.dev)
        ops.fmha(q0, K, V, o, B, H, 1, nq + 1, dh, (dim, dim, dh), ((nq + 1) * dim, dim, dh), ((nq + 1) * dim, dim, dh),
                 (dim, dim, dh), 1.0 / math.sqrt(dh))
        pooled = ops.gemm(o, self.ap_c[0], bias=self.ap_c[1])
        return prompt.view(B, nq, o1 + o2), pooled
