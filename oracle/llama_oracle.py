"""ORACLE — test infrastructure only (never imported by the product path).

CPU/torch restatement of the MLLM half of the SEED-Story hot path:
  * Llama-2 decoder forward as vendored by the reference, src/models_clm/modeling_llama_xformer.py
      RMSNorm :107-115, RoPE :118-173, MLP :190-191, attention :217-301, layer :318-368,
      model :532-666, lm_head :759
  * xformers `memory_efficient_attention(..., LowerTriangularFromBottomRightMask)` (:289-295) restated as a
    masked softmax(QK^T/sqrt(d))V with the diagonal anchored bottom-right
  * peft 0.4.0 LoRA Linear (requirements.txt:24, config configs/clm_models/llama2chat7b_lora.yaml:7-27):
      y = x W^T + ((x A^T) B^T) * (alpha / r), dropout off in eval            [third-party, not in tree]
  * transformers 4.34 greedy_search as driven by src/models_clm/models.py:137-153 and
    prepare_inputs_for_generation (:796-852)                                    [third-party, not in tree]
  * AutoImageTokenGenerationProcessor, src/models_clm/generation.py:9-31
  * ContinuousLVLM.generate post-processing, src/models_clm/models.py:156-221

Pinning: `oracle/pin_against_reference.py` runs the reference's own modules (imported from /root/reference,
xformers replaced by an SDPA stand-in) against these functions and freezes tests/golden/*.pt.
The greedy loop is additionally pinned against `transformers.GenerationMixin.generate` as installed here (5.5; the
reference pins 4.34.0) — ids equal, hidden states within 3e-6 (tests/golden/hf_greedy_loop.pt).  The LoRA arithmetic is a
third-party restatement (peft is absent): parity for it is UNPINNED.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# layer math
# --------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps):
    """modeling_llama_xformer.py:107-115 — fp32 variance; the product is cast to the weight dtype
    before the weight multiply when the weight is half precision."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        y = y.to(weight.dtype)
    return weight * y


def rope_tables(head_dim, n_pos, base=10000.0):
    """:120-137 — fp32 tables [n_pos, head_dim] = cat(freqs, freqs)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(n_pos, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin, position_ids):
    """:165-173 — q,k [B,H,T,D]; tables are cast to the activation dtype first (:150-151)."""
    cos = cos.to(q.dtype)[position_ids].unsqueeze(1)
    sin = sin.to(q.dtype)[position_ids].unsqueeze(1)
    return q * cos + _rot_half(q) * sin, k * cos + _rot_half(k) * sin


def attend_bottom_right(q, k, v):
    """q [B,H,Tq,D], k/v [B,H,Tk,D]; query i sees keys j <= i + (Tk - Tq); softmax in fp32."""
    tq, tk, d = q.shape[-2], k.shape[-2], q.shape[-1]
    scores = torch.matmul(q.float(), k.float().transpose(-1, -2)) / math.sqrt(d)
    qi = torch.arange(tq).unsqueeze(1)
    kj = torch.arange(tk).unsqueeze(0)
    scores = scores.masked_fill(~(kj <= qi + (tk - tq)), float("-inf"))
    p = torch.softmax(scores, dim=-1)
    return torch.matmul(p, v.float()).to(q.dtype)


def lora_linear(x, w, a=None, b=None, scaling=2.0):
    y = F.linear(x, w)
    if a is not None:
        y = y + F.linear(F.linear(x, a), b) * scaling
    return y


class LlamaParams:
    """Plain container: per-layer dicts of tensors named like the reference's state_dict."""

    def __init__(self, hidden, inter, n_heads, n_layers, vocab, eps=1e-5, lora_r=0, scaling=2.0):
        self.hidden, self.inter, self.n_heads, self.n_layers, self.vocab = hidden, inter, n_heads, n_layers, vocab
        self.eps, self.lora_r, self.scaling = eps, lora_r, scaling
        self.layers = []
        self.embed = None
        self.norm = None
        self.lm_head = None

    @staticmethod
    def random(hidden, inter, n_heads, n_layers, vocab, lora_r=16, seed=0, std=0.02, dtype=torch.float32, eps=1e-5):
        g = torch.Generator().manual_seed(seed)
        p = LlamaParams(hidden, inter, n_heads, n_layers, vocab, eps=eps, lora_r=lora_r)

        def rn(*shape, s=std):
            return (torch.randn(*shape, generator=g) * s).to(dtype)
        p.embed = rn(vocab, hidden)
        for _ in range(n_layers):
            L = {}
            for name, (o, i) in {"q_proj": (hidden, hidden), "k_proj": (hidden, hidden), "v_proj": (hidden, hidden),
                                 "o_proj": (hidden, hidden), "gate_proj": (inter, hidden), "up_proj": (inter, hidden),
                                 "down_proj": (hidden, inter)}.items():
                L[name] = rn(o, i)
                if lora_r:
                    L[name + ".lora_A"] = rn(lora_r, i)
                    L[name + ".lora_B"] = rn(o, lora_r)
            L["input_layernorm"] = (1.0 + rn(hidden, s=0.1)).to(dtype)
            L["post_attention_layernorm"] = (1.0 + rn(hidden, s=0.1)).to(dtype)
            p.layers.append(L)
        p.norm = (1.0 + rn(hidden, s=0.1)).to(dtype)
        p.lm_head = rn(vocab, hidden)
        return p

    def to(self, dtype=None, device=None):
        def cv(t):
            return t.to(dtype=dtype, device=device)
        q = LlamaParams(self.hidden, self.inter, self.n_heads, self.n_layers, self.vocab, self.eps, self.lora_r,
                        self.scaling)
        q.embed, q.norm, q.lm_head = cv(self.embed), cv(self.norm), cv(self.lm_head)
        q.layers = [{k: cv(v) for k, v in L.items()} for L in self.layers]
        return q


def _proj(L, name, x, scaling):
    return lora_linear(x, L[name], L.get(name + ".lora_A"), L.get(name + ".lora_B"), scaling)


def decoder_layer(p, L, h, cos, sin, position_ids, past_kv):
    """:318-368 with attention :217-301.  Returns (h, (k, v)) — k cached post-RoPE, v raw (:236-244)."""
    B, T, _ = h.shape
    H, D = p.n_heads, p.hidden // p.n_heads
    x = rms_norm(h, L["input_layernorm"], p.eps)
    q = _proj(L, "q_proj", x, p.scaling).view(B, T, H, D).transpose(1, 2)
    k = _proj(L, "k_proj", x, p.scaling).view(B, T, H, D).transpose(1, 2)
    v = _proj(L, "v_proj", x, p.scaling).view(B, T, H, D).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin, position_ids)
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=2)
        v = torch.cat([past_kv[1], v], dim=2)
    attn = attend_bottom_right(q, k, v).transpose(1, 2).reshape(B, T, p.hidden)
    h = h + _proj(L, "o_proj", attn, p.scaling)
    x = rms_norm(h, L["post_attention_layernorm"], p.eps)
    mlp = _proj(L, "down_proj", F.silu(_proj(L, "gate_proj", x, p.scaling)) * _proj(L, "up_proj", x, p.scaling),
                p.scaling)
    return h + mlp, (k, v)


def model_forward(p, inputs_embeds, position_ids, past_kvs=None, max_pos=4096):
    """LlamaModel.forward + lm_head (:532-666, :759).  Returns (logits, final_hidden (post-norm), new_kvs)."""
    D = p.hidden // p.n_heads
    cos, sin = rope_tables(D, max_pos)
    cos, sin = cos.to(inputs_embeds.device), sin.to(inputs_embeds.device)
    h = inputs_embeds
    new = []
    for i, L in enumerate(p.layers):
        h, kv = decoder_layer(p, L, h, cos, sin, position_ids, None if past_kvs is None else past_kvs[i])
        new.append(kv)
    hn = rms_norm(h, p.norm, p.eps)
    logits = F.linear(hn, p.lm_head)
    return logits, hn, new


# --------------------------------------------------------------------------------------------
# generation
# --------------------------------------------------------------------------------------------
def image_token_processor(last_id, scores, img_ids):
    """generation.py:19-31 for one sequence; `scores` [V] is edited in place like the reference."""
    if last_id in img_ids[:-1]:
        out_id = img_ids[img_ids.index(last_id) + 1]
        scores[out_id] = scores.max() + 10.0
    else:
        scores[torch.tensor(img_ids[1:], dtype=torch.long)] = 0.0
    return scores


def top2_margin(scores):
    """(top-1 minus top-2 logit) / max|logit| of a processed score row: how far the greedy choice is from a tie.
    Test helper for the "token ids exact wherever the oracle's margin exceeds fp16 noise" rule."""
    s = scores.float()
    s = s[torch.isfinite(s)]
    top = torch.topk(s, 2).values
    return float((top[0] - top[1]) / s.abs().max().clamp_min(1e-12))


def first_divergence(got, ref):
    n = min(len(got), len(ref))
    for i in range(n):
        if got[i] != ref[i]:
            return i
    return n


def greedy_generate(p, input_ids, inputs_embeds, img_ids, eos_id, max_new_tokens, use_processor=True,
                    forced_schedule=None, margins_out=None, past_kvs=None, head=0):
    """HF 4.34 greedy_search as configured by models.py:137-153 (batch 1).

    step 0 feeds inputs_embeds with position_ids = arange(L) (cumsum(ones)-1, :832-837); later steps feed the
    embedding of the last token at position len-1.  Returns (sequence ids list, final_hidden rows [L+T-1, hidden]
    where row i is the post-norm hidden state of the position whose input is sequence[i], kv caches).
    `forced_schedule`: optional list of token ids that override argmax for the first steps (synthetic-weights
    benchmark schedule; the reference exposes the same hook through `logits_processor=`).
    `margins_out`: optional list that receives top2_margin() of every generated step.
    `past_kvs`/`head`: KV reuse with `use_kv_cache_head=True` (prepare_inputs_for_generation :804-826): step 0 feeds
    only inputs_embeds[:, head:] at positions head..L-1 on top of the given cache; the returned hidden rows then start
    at input position `head` (models.py:186-189 keeps all of them)."""
    assert input_ids.shape[0] == 1
    seq = input_ids[0].tolist()
    L = len(seq)
    if past_kvs is None:
        head = 0
    pos = torch.arange(head, L).unsqueeze(0)
    logits, hn, kvs = model_forward(p, inputs_embeds[:, head:], pos, past_kvs)
    hiddens = [hn[0]]
    n_new = 0
    while True:
        scores = logits[0, -1].clone()
        if use_processor:
            scores = image_token_processor(seq[-1], scores, img_ids)
        nxt = int(torch.argmax(scores.float()).item())
        forced = forced_schedule is not None and n_new < len(forced_schedule) and forced_schedule[n_new] is not None
        if forced:
            nxt = forced_schedule[n_new]
        if margins_out is not None:   # a forced step cannot flip: infinite margin
            margins_out.append(float("inf") if forced else top2_margin(scores))
        seq.append(nxt)
        n_new += 1
        if nxt == eos_id or n_new >= max_new_tokens:
            break
        emb = p.embed[torch.tensor([[nxt]])]
        pos = torch.tensor([[len(seq) - 1]])
        logits, hn, kvs = model_forward(p, emb, pos, kvs)
        hiddens.append(hn[0])
    return seq, torch.cat(hiddens, dim=0), kvs


def lvlm_postprocess_past(sequence, rows, eoi_id, num_img_gen_tokens=64):
    """models.py:186-197, past_key_values branch: ALL hidden rows of the call are kept and </img> is searched in the
    last len(rows) ids of the full sequence (prompt + generated) — one position later than the rows' own inputs."""
    tail = list(sequence)[-rows.shape[0]:]
    eoi = [i for i, t in enumerate(tail) if t == eoi_id]
    if not eoi:
        return None
    e = eoi[-1]
    return rows[e - num_img_gen_tokens:e]


def lvlm_postprocess(generate_ids, last_hidden_states, eoi_id, num_img_gen_tokens=64):
    """models.py:182-205 (past_key_values=None branch): rows of the LAST </img>; hidden rows are indexed
    relative to the first generated token's input position (last_hidden_states[L:])."""
    eoi = [i for i, t in enumerate(generate_ids) if t == eoi_id]
    if not eoi:
        return None
    e = eoi[-1]
    return last_hidden_states[e - num_img_gen_tokens:e]
