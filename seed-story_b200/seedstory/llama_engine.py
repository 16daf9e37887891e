"""Llama-2 decode/prefill engine over the seedstory_b200 kernels (host side: buffers, page tables, launches).

Replaces, for inference, the arithmetic of the reference's
  src/models_clm/modeling_llama_xformer.py  LlamaModel.forward :532-666, LlamaDecoderLayer :318-368,
                                            LlamaAttention :217-301, lm_head :759,
                                            prepare_inputs_for_generation :796-852 (position rule)
  transformers 4.34 greedy_search (call site src/models_clm/models.py:146-153)
  src/models_clm/generation.py:19-31        (logits processor, fused with the argmax on the device)
  peft 0.4 LoRA Linear                      (folded into the base weights once at load)

Data layout in HBM
  weights   per layer: qkv [3*hid, hid] (q|k|v rows), o [hid, hid], gate_up [2*inter, hid] with rows
            interleaved (gate_j, up_j), down [hid, inter]; fp16, LoRA merged (fp32 accumulate, one rounding)
  KV cache  k_pages/v_pages [layers, pages, heads, 64, head_dim] fp16; a sequence owns an ordered list of
            pages (page table row); sink compaction copies retained tokens into fresh pages
  decode    every per-step scalar (ids, positions, slots, lengths, done flags) lives on the device so one
            decode step is a single CUDA-graph replay with no host round trip except the 4-byte id read.
"""
import math

import torch

from . import _capi, ops

PAGE = ops.KV_PAGE


class LlamaConfig:
    def __init__(self, hidden=4096, inter=11008, heads=32, layers=32, vocab=32066, eps=1e-5, max_pos=4096):
        self.hidden, self.inter, self.heads, self.layers, self.vocab, self.eps, self.max_pos = \
            hidden, inter, heads, layers, vocab, eps, max_pos
        self.head_dim = hidden // heads


def rope_tables_f16(head_dim, n_pos, device, base=10000.0):
    """fp32 tables as the reference builds them (modeling_llama_xformer.py:120-137), cast to fp16 (:150-151).
    Constant precompute on the host at load time."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(n_pos, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().half().to(device), emb.sin().half().to(device)


class LlamaEngine:
    def __init__(self, cfg, device, max_batch=1, max_ctx=4096, max_new=512, decode_splits=12):
        ops.require_device()
        self.cfg, self.dev = cfg, device
        self.max_batch = max_batch
        self.max_pages = (max_ctx + PAGE - 1) // PAGE
        self.max_new = max_new
        self.splits = max(1, decode_splits)
        c = cfg
        assert c.head_dim == 128, "attention kernels are specialised for head_dim 128"
        self.w = None
        n_pages = max_batch * self.max_pages + 8
        self.k_pages = torch.zeros((c.layers, n_pages, c.heads, PAGE, c.head_dim), dtype=torch.float16, device=device)
        self.v_pages = torch.zeros_like(self.k_pages)
        self.free_pages = list(range(n_pages))
        self.page_table_h = torch.zeros((max_batch, self.max_pages), dtype=torch.int32)
        self.page_table = torch.zeros((max_batch, self.max_pages), dtype=torch.int32, device=device)
        self.n_pages_owned = [0] * max_batch
        self.seq_len_h = [0] * max_batch
        self.cos, self.sin = rope_tables_f16(c.head_dim, c.max_pos, device)
        i32 = dict(dtype=torch.int32, device=device)
        B = max_batch
        # decode state (device resident)
        self.cur_ids = torch.zeros(B, **i32)
        self.next_ids = torch.zeros(B, **i32)
        self.tok_pos = torch.zeros(B, **i32)
        self.tok_slot = torch.zeros(B, **i32)
        self.seq_lens = torch.zeros(B, **i32)
        self.tok_seq = torch.arange(B, **i32)
        self.n_out = torch.zeros(B, **i32)
        self.done = torch.zeros(B, **i32)
        self.out_ids = torch.zeros((B, max_new), **i32)
        self.schedule = torch.full((B, max_new), -1, **i32)
        self.hist = torch.zeros((B, max_new, c.hidden), dtype=torch.float16, device=device)
        f16 = dict(dtype=torch.float16, device=device)
        self.d_h = torch.zeros((B, c.hidden), **f16)
        self.d_xn = torch.zeros((B, c.hidden), **f16)
        self.d_qkv = torch.zeros((B, 3 * c.hidden), **f16)
        self.d_q = torch.zeros((B, c.hidden), **f16)
        self.d_attn = torch.zeros((B, c.hidden), **f16)
        self.d_act = torch.zeros((B, c.inter), **f16)
        self.d_logits = torch.zeros((B, c.vocab), **f16)
        self.d_ws = ops.attn_decode_workspace(B, c.heads, c.head_dim, self.splits, device)
        # per-step constants of the fused q/k/v kernel (cache destination + rotary factors of each sequence's new token)
        self.d_kv_base = torch.zeros(B, dtype=torch.int64, device=device)
        self.d_rope_cs = torch.zeros((B, c.head_dim), **f16)
        self.d_rope_sn = torch.zeros((B, c.head_dim), **f16)
        self.img_ids = None
        self.img_ids_h = [-1, -2]   # no image-token processor until set_image_token_ids()
        self.suppress_ids = None
        self.suppress_ids_h = []
        self.eos_id = 2
        self._graphs = {}
        self._pinned_ids = torch.zeros(B, dtype=torch.int32).pin_memory()
        self._pinned_done = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.decode_block = 16          # decode steps queued per host sync in generate()
        self._pinned_blk = torch.zeros(64, dtype=torch.int32).pin_memory()

    # ------------------------------------------------------------------ weights
    def load_weights(self, embed, layers, norm, lm_head, lora_scaling=2.0):
        """`layers`: list of dicts with q_proj..down_proj [out,in] fp16 (+ optional '<name>.lora_A' [r,in],
        '<name>.lora_B' [out,r]), input_layernorm, post_attention_layernorm.  Packs into kernel layouts."""
        def merged(L, n):
            W = L[n].to(self.dev, torch.float16).contiguous()
            if n + ".lora_A" in L:
                W = ops.lora_merge(W, L[n + ".lora_A"].to(self.dev, torch.float16),
                                   L[n + ".lora_B"].to(self.dev, torch.float16), lora_scaling)
            return W
        packed = []
        for L in layers:
            q, k, v = merged(L, "q_proj"), merged(L, "k_proj"), merged(L, "v_proj")
            g, u = merged(L, "gate_proj"), merged(L, "up_proj")
            packed.append(dict(
                qkv=torch.cat([q, k, v], 0).contiguous(),
                qkv_dec=ops.interleave_rope_rows(torch.cat([q, k, v], 0), self.cfg.heads, self.cfg.head_dim),
                o=merged(L, "o_proj"),
                gate_up=torch.stack([g, u], dim=1).reshape(2 * g.shape[0], g.shape[1]).contiguous(),
                down=merged(L, "down_proj"),
                ln1=L["input_layernorm"].to(self.dev, torch.float16).contiguous(),
                ln2=L["post_attention_layernorm"].to(self.dev, torch.float16).contiguous()))
            del q, k, v, g, u
        self.w = dict(layers=packed, embed=embed.to(self.dev, torch.float16).contiguous(),
                      norm=norm.to(self.dev, torch.float16).contiguous(),
                      lm_head=lm_head.to(self.dev, torch.float16).contiguous())
        ops.register_const_tree(self.w)        # packed once here, never written again
        torch.cuda.current_stream().synchronize()
        self._graphs.clear()

    def set_image_token_ids(self, img_ids, eos_id=2):
        """img_ids = [BOI, IMG_0..IMG_{n-1}, EOI] as produced by the reference processor's tokenizer.encode
        (src/models_clm/generation.py:14-17)."""
        self.img_ids = torch.tensor(list(img_ids), dtype=torch.int32, device=self.dev)
        self.img_ids_h = list(img_ids)
        self.eos_id = eos_id
        self._graphs.clear()

    def set_suppress_ids(self, ids):
        """transformers' SuppressTokensLogitsProcessor (scores[:, ids] = -inf every step), fused with the argmax."""
        ids = sorted(set(int(i) for i in (ids or [])))
        if ids == self.suppress_ids_h:
            return
        self.suppress_ids_h = ids
        self.suppress_ids = torch.tensor(ids, dtype=torch.int32, device=self.dev) if ids else None
        self._graphs.clear()

    def embed_tokens(self, ids):
        ids = ids.to(self.dev, torch.int32).reshape(-1).contiguous()
        out = torch.empty((ids.numel(), self.cfg.hidden), dtype=torch.float16, device=self.dev)
        ops.gather_rows(self.w["embed"], ids, out)
        return out

    # ------------------------------------------------------------------ KV pages
    def reset_sequence(self, b):
        n = self.n_pages_owned[b]
        self.free_pages.extend(self.page_table_h[b, :n].tolist())
        self.n_pages_owned[b] = 0
        self.seq_len_h[b] = 0

    def _ensure_pages(self, b, n_tokens):
        need = (n_tokens + PAGE - 1) // PAGE
        assert need <= self.max_pages, f"sequence needs {need} pages > max {self.max_pages}"
        changed = False
        while self.n_pages_owned[b] < need:
            self.page_table_h[b, self.n_pages_owned[b]] = self.free_pages.pop()
            self.n_pages_owned[b] += 1
            changed = True
        if changed:
            self.page_table[b].copy_(self.page_table_h[b], non_blocking=True)

    def truncate(self, b, n):
        """Drop every cache slot >= n of sequence b (e.g. the generated part of a turn, vis_george_sink.py:243: the
        cache is cut back to the prompt).  No data moves; surplus pages return to the pool (later kernels on the
        stream are the only possible writers, so stream order keeps this safe)."""
        assert 0 <= n <= self.seq_len_h[b]
        need = (n + PAGE - 1) // PAGE
        owned = self.n_pages_owned[b]
        if need < owned:
            self.free_pages.extend(self.page_table_h[b, need:owned].tolist())
            self.n_pages_owned[b] = need
        self.seq_len_h[b] = n

    def retain_tokens(self, b, keep):
        """Keep only cache slots `keep` (sorted list of slot indices) of sequence b, compacted into fresh pages:
        the window cut (gen_george.py:235-239) and the attention-sink retention (vis_george_sink.py:266-291).
        Keys keep their original RoPE phase (they are cached post-RoPE, modeling_llama_xformer.py:236-244)."""
        keep = list(keep)
        n = len(keep)
        assert all(0 <= k < self.seq_len_h[b] for k in keep) and keep == sorted(set(keep))
        old_n = self.n_pages_owned[b]
        old_pages = self.page_table_h[b, :old_n].clone()
        need = (n + PAGE - 1) // PAGE
        new_pages = torch.tensor([self.free_pages.pop() for _ in range(need)], dtype=torch.int32)
        i32 = dict(dtype=torch.int32, device=self.dev)
        if n:
            ops.kv_gather_tokens(self.k_pages, self.v_pages, old_pages.to(self.dev), new_pages.to(self.dev),
                                 torch.tensor(keep, **i32), self.cfg.heads, self.cfg.head_dim)
        # the old pages go back to the pool without a host sync: every later writer of those pages is queued on this
        # stream behind the gather that reads them (the same argument as in truncate())
        self.free_pages.extend(old_pages.tolist())
        self.page_table_h[b].zero_()
        self.page_table_h[b, :need] = new_pages
        self.page_table[b].copy_(self.page_table_h[b])
        self.n_pages_owned[b] = need
        self.seq_len_h[b] = n

    # ------------------------------------------------------------------ prefill / chunk
    def forward_chunk(self, b, embeds, positions, want_logits=True):
        """Run T tokens of sequence b on top of its cache (prefill when the cache is empty).
        embeds [T, hidden] fp16 device; positions: int tensor/list [T] (window-relative RoPE positions).
        Returns (final-norm hidden [T, hidden], logits [1, vocab] of the last row or None)."""
        c, w = self.cfg, self.w
        T = embeds.shape[0]
        base = self.seq_len_h[b]
        self._ensure_pages(b, base + T)
        i32 = dict(dtype=torch.int32, device=self.dev)
        pos = torch.as_tensor(positions, dtype=torch.int32).to(self.dev)
        slot = torch.arange(base, base + T, **i32)
        seq = torch.full((T,), b, **i32)
        h = embeds.contiguous().clone()
        q = torch.empty((T, c.hidden), dtype=torch.float16, device=self.dev)
        attn = torch.empty_like(q)
        xn = torch.empty_like(q)
        qkv = torch.empty((T, 3 * c.hidden), dtype=torch.float16, device=self.dev)
        act = torch.empty((T, c.inter), dtype=torch.float16, device=self.dev)
        scale = 1.0 / math.sqrt(c.head_dim)
        pt_row = self.page_table[b:b + 1]
        Lk = base + T
        for li, L in enumerate(w["layers"]):
            ops.rmsnorm(h, L["ln1"], c.eps, out=xn)
            ops.gemm(xn, L["qkv"], out=qkv)
            ops.rope_kv_append(qkv, q, self.k_pages[li], self.v_pages[li], seq, pos, slot, self.page_table, self.cos,
                               self.sin, c.heads, c.head_dim)
            ops.fmha(q, self.k_pages[li], self.v_pages[li], attn, 1, c.heads, T, Lk, c.head_dim,
                     (0, c.hidden, c.head_dim), (0, 0, 0), (0, 0, 0), (0, c.hidden, c.head_dim), scale, causal=True,
                     page_table=pt_row)
            ops.gemm(attn, L["o"], residual=h, out=h)
            ops.rmsnorm(h, L["ln2"], c.eps, out=xn)
            ops.gemm(xn, L["gate_up"], glu=ops.GLU_SWIGLU, out=act)
            ops.gemm(act, L["down"], residual=h, out=h)
        hn = ops.rmsnorm(h, w["norm"], c.eps)
        self.seq_len_h[b] = Lk
        logits = None
        if want_logits:
            logits = ops.skinny_gemm(hn[T - 1:T], w["lm_head"])
        return hn, logits

    # ------------------------------------------------------------------ decode step (graph body)
    def _decode_body(self, B):
        c, w = self.cfg, self.w
        scale = 1.0 / math.sqrt(c.head_dim)
        h, xn, q, attn, act = (t[:B] for t in (self.d_h, self.d_xn, self.d_q, self.d_attn, self.d_act))
        ops.gather_rows(w["embed"], self.cur_ids[:B], h)
        for li, L in enumerate(w["layers"]):
            # 5 launches per layer: [RMSNorm + q/k/v + RoPE + cache append] [attention] [o_proj + residual]
            # [RMSNorm + gate/up + SwiGLU] [down_proj + residual]
            ops.decode_qkv_rope_append(h, L["ln1"], c.eps, L["qkv_dec"], q, self.k_pages[li], self.v_pages[li],
                                       self.d_kv_base[:B], self.d_rope_cs[:B], self.d_rope_sn[:B], c.heads, c.head_dim)
            ops.attn_decode_paged(q, self.k_pages[li], self.v_pages[li], self.seq_lens[:B], self.page_table, attn,
                                  self.d_ws, c.heads, c.head_dim, self.splits, scale)
            ops.skinny_gemm(attn, L["o"], ops.EPI_RESIDUAL, residual=h, out=h)
            ops.skinny_gemm_rmsnorm(h, L["ln2"], c.eps, L["gate_up"], ops.EPI_SWIGLU, out=act)
            ops.skinny_gemm(act, L["down"], ops.EPI_RESIDUAL, residual=h, out=h)
        ops.rmsnorm(h, w["norm"], c.eps, out=xn)
        ops.store_rows_indexed(xn, self.hist[:B], self.n_out[:B])
        ops.skinny_gemm(xn, w["lm_head"], out=self.d_logits[:B])
        ops.logits_process_argmax(self.d_logits[:B], self.cur_ids[:B], self.img_ids, self.next_ids[:B],
                                  self.suppress_ids)
        ops.decode_advance(self.next_ids[:B], self.cur_ids[:B], self.tok_pos[:B], self.tok_slot[:B], self.seq_lens[:B],
                           self.out_ids[:B], self.n_out[:B], self.done[:B], self.eos_id, self.schedule[:B])
        self._rope_meta(B)      # for the NEXT step

    def _rope_meta(self, B):
        c = self.cfg
        ops.decode_rope_meta(self.tok_seq[:B], self.tok_pos[:B], self.tok_slot[:B], self.page_table, self.cos, self.sin,
                             c.heads, c.head_dim, self.d_kv_base[:B], self.d_rope_cs[:B], self.d_rope_sn[:B])

    def decode_step(self, B, use_graph=True):
        if not use_graph:
            self._decode_body(B)
            return
        g = self._graphs.get(B)
        if g is None:
            # warm up outside capture (module-level lazy init: func attributes, tensor-map cache)
            state = [t.clone() for t in (self.cur_ids, self.tok_pos, self.tok_slot, self.seq_lens, self.n_out,
                                         self.done, self.out_ids)]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._decode_body(B)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            for t, sv in zip((self.cur_ids, self.tok_pos, self.tok_slot, self.seq_lens, self.n_out, self.done,
                              self.out_ids), state):
                t.copy_(sv)
            g = torch.cuda.CUDAGraph()
            n0 = _capi.launch_count()
            with torch.cuda.graph(g):
                self._decode_body(B)
            self._graph_launches = _capi.launch_count() - n0
            for t, sv in zip((self.cur_ids, self.tok_pos, self.tok_slot, self.seq_lens, self.n_out, self.done,
                              self.out_ids), state):
                t.copy_(sv)
            self._rope_meta(B)     # the warm-up step advanced it
            self._graphs[B] = g
        g.replay()
        _capi.add_launches(self._graph_launches)

    # ------------------------------------------------------------------ greedy generation (batch of 1..B)
    def begin_decode(self, first_ids, positions, schedule=None):
        """Arm the device-side decode state after prefill.  first_ids[b] is the first generated token (already
        chosen from the prefill logits); positions[b] its RoPE position."""
        B = len(first_ids)
        for b in range(B):
            self._ensure_pages(b, min(self.seq_len_h[b] + self.max_new + 1, self.max_pages * PAGE))
        i32 = dict(dtype=torch.int32)
        self.cur_ids[:B].copy_(torch.tensor(first_ids, **i32))
        self.tok_pos[:B].copy_(torch.tensor(positions, **i32))
        self.tok_slot[:B].copy_(torch.tensor(self.seq_len_h[:B], **i32))
        self.seq_lens[:B].copy_(torch.tensor([n + 1 for n in self.seq_len_h[:B]], **i32))
        self.n_out[:B].fill_(1)
        self.done[:B].copy_(torch.tensor([1 if t == self.eos_id else 0 for t in first_ids], **i32))
        self.out_ids[:B, 0].copy_(torch.tensor(first_ids, **i32))
        if schedule is not None:
            self.schedule[:B].copy_(schedule)
        else:
            self.schedule[:B].fill_(-1)
        self._rope_meta(B)

    def read_step(self, B):
        """4-byte-per-sequence read-back of the ids just emitted (the only host sync of a decode step)."""
        self._pinned_ids[:B].copy_(self.cur_ids[:B], non_blocking=True)
        self._pinned_done[:B].copy_(self.done[:B], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._pinned_ids[:B].tolist(), self._pinned_done[:B].tolist()

    def read_block(self, start, k):
        """Ids emitted by the last k queued decode steps of sequence 0 (fewer than k if EOS stopped it): ONE host sync."""
        self._pinned_blk[:k].copy_(self.out_ids[0, start:start + k], non_blocking=True)
        self._pinned_ids[:1].copy_(self.n_out[:1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = int(self._pinned_ids[0]) - start
        return self._pinned_blk[:n].tolist()

    def first_token(self, logits, last_id, sched0=-1):
        """Processor + argmax on prefill logits (same kernel as the decode step)."""
        last = torch.tensor([last_id], dtype=torch.int32, device=self.dev)
        nxt = torch.zeros(1, dtype=torch.int32, device=self.dev)
        ops.logits_process_argmax(logits, last, self.img_ids, nxt, self.suppress_ids)
        t = int(nxt.item())
        return sched0 if sched0 >= 0 else t

    def load_past(self, b, past):
        """Make `past` the cache of sequence b: a sequence over layers of (K, V) tensors [1, heads, n, head_dim] — the
        reference's `past_key_values` tuple (modeling_llama_xformer.py:241-244), e.g. the sliced / concatenated
        attention-sink cache of src/inference/vis_george_sink.py:266-291.  K rows keep the RoPE phase they were
        cached with.  Returns n."""
        c = self.cfg
        past = list(past)
        assert len(past) == c.layers, f"past_key_values has {len(past)} layers, the model {c.layers}"
        n = int(past[0][0].shape[2])
        self.reset_sequence(b)
        self._ensure_pages(b, n)
        pages = self.page_table[b]
        for li, (k, v) in enumerate(past):
            assert k.shape[0] == 1 and tuple(k.shape[1:]) == (c.heads, n, c.head_dim) and v.shape == k.shape, \
                f"layer {li}: past K/V must be [1, {c.heads}, n, {c.head_dim}]"
            k3 = k[0].to(self.dev, torch.float16)
            v3 = v[0].to(self.dev, torch.float16)
            if k3.stride(2) != 1 or k3.stride(0) % 8 or k3.stride(1) % 8:
                k3 = k3.contiguous()
            if v3.stride(2) != 1 or v3.stride(0) % 8 or v3.stride(1) % 8:
                v3 = v3.contiguous()
            ops.kv_scatter_tokens(self.k_pages[li], self.v_pages[li], k3, v3, pages)
        self.seq_len_h[b] = n
        return n

    def generate(self, b, input_ids, inputs_embeds, max_new_tokens, schedule=None, chunk_image_run=True,
                 use_graph=True, past_len=None, head=0, return_chunk_hidden=False):
        """Greedy generation for sequence slot b (reference semantics: HF greedy_search with the image-token
        processor, stop on EOS or max_new_tokens).  `schedule`: optional list (len <= max_new) of forced ids
        (-1 = free).  Returns (generated ids list, hidden rows [T-1, hidden] where row i is the post-norm hidden
        state of the position whose input is generated id i) [+ the post-norm hidden rows of the fed prompt chunk].

        past_len=None: the sequence is reset and the whole prompt is prefilled at positions 0..L-1.
        past_len=n (live sink-KV mode, `use_kv_cache_head=True` in the reference, prepare_inputs_for_generation
        :804-826): the cache of slot b already holds n tokens (load_past / retain_tokens); only input_ids[head:] are
        fed, at positions head..L-1 — positions count the CURRENT (windowed) input_ids, while the retained sink keys
        keep the phase they were cached with — and generated token t sits at position L+t."""
        assert self.max_batch >= 1 and b == 0, "single-story generate uses slot 0; batched decode goes through begin_decode/decode_step"
        c = self.cfg
        L = len(input_ids)
        n_cached = 0 if past_len is None else int(past_len)
        if past_len is None:
            head = 0
        assert 0 <= head < L, "at least one prompt token must be fed on top of the cache"
        total = n_cached + (L - head) + max_new_tokens + 1
        if max_new_tokens + 1 > self.max_new or total > self.max_pages * PAGE or L + max_new_tokens + 1 > c.max_pos:
            raise _capi.SeedStoryError(f"generate({L} prompt + {max_new_tokens} new tokens on {n_cached} cached) exceeds "
                                       f"the engine's capacity (max_new {self.max_new}, max_ctx {self.max_pages * PAGE}, "
                                       f"max_pos {c.max_pos})")
        if past_len is None:
            self.reset_sequence(b)
        else:
            assert self.seq_len_h[b] == n_cached, (self.seq_len_h[b], n_cached)
        hn, logits = self.forward_chunk(b, inputs_embeds[head:], list(range(head, L)))
        sched = [-1] * self.max_new
        if schedule is not None:
            sched[:len(schedule)] = schedule
        first = self.first_token(logits, int(input_ids[-1]), sched[0])
        gen = [first]
        hid_rows = []
        sched_t = torch.tensor([sched], dtype=torch.int32)
        eoi = self.img_ids_h[-1]
        armed = False
        while gen[-1] != self.eos_id and len(gen) < max_new_tokens:
            # inside an image run the remaining ids are input-determined (generation.py:23-26): feed what is left of
            # [<img>, IMG_0.., </img>] as ONE chunk (normally the whole run: the last id is <img>)
            i_run = self.img_ids_h.index(gen[-1]) if gen[-1] in self.img_ids_h[:-1] else -1
            n_left = len(self.img_ids_h) - 1 - i_run
            if chunk_image_run and i_run >= 0 and len(gen) + n_left < max_new_tokens \
                    and all(s < 0 for s in sched[len(gen):len(gen) + n_left]):
                run = self.img_ids_h[i_run:]
                emb = self.embed_tokens(torch.tensor(run))
                p0 = L + len(gen) - 1
                hn_c, logits = self.forward_chunk(b, emb, list(range(p0, p0 + len(run))))
                hid_rows.append(hn_c)
                gen.extend(run[1:])
                nxt = self.first_token(logits, eoi, sched[len(gen)] if len(gen) < len(sched) else -1)
                gen.append(nxt)
                armed = False
                continue
            if not armed:
                self.begin_decode([gen[-1]], [L + len(gen) - 1], sched_t)
                self.n_out[:1].fill_(len(gen))
                armed = True
            # a BLOCK of decode steps is queued before the host looks (one sync per block, not per token): the device
            # applies the processor, the schedule and the EOS stop itself (after EOS the steps are no-ops on the state).
            # A block ends at the next scheduled id (it may be <img>: the run then takes the chunk path) and at
            # max_new_tokens; a free-running <img> inside a block is followed by forced query ids decoded one by one,
            # exactly as the reference decodes them, and the rest of the run is chunked above.
            start = len(gen)
            k = min(self.decode_block, max_new_tokens - start)
            for j in range(start, start + k):
                if sched[j] >= 0:
                    k = j - start + 1
                    break
            for _ in range(k):
                self.decode_step(1, use_graph)
            ids = self.read_block(start, k)
            n = len(ids)
            self.seq_len_h[b] += n
            hid_rows.append(self.hist[0, start:start + n].clone())
            gen.extend(ids)
            if n < k:
                assert gen[-1] == self.eos_id, (gen[-4:], n, k)
        hidden = torch.cat(hid_rows, 0) if hid_rows else torch.empty((0, c.hidden), dtype=torch.float16, device=self.dev)
        if return_chunk_hidden:
            return gen, hidden, hn
        return gen, hidden


    # ------------------------------------------------------------------ greedy generation, several sequences at once
    def generate_batch(self, reqs, chunk_image_run=True, use_graph=True):
        """Greedy generation for len(reqs) <= max_batch independent sequences that share every decode step (one pass over
        the weights per step for all of them): continuous batching over the paged KV cache — BASELINE configs[3].  The
        reference ignores padding masks (modeling_llama_xformer.py:289-295, 844), so a sequence's result must not depend on
        its neighbours: every sequence gets exactly what generate() would give it alone (same kernels row by row).

        reqs[b] = dict(input_ids=list[int], inputs_embeds=[L, hidden] fp16, max_new_tokens=int, schedule=list|None,
                       past_len=None|int, head=int), or None for an empty slot (sequence b keeps its engine slot b, so
        a finished story leaves a hole).  Returns [(generated ids, hidden rows, prompt-chunk hidden rows) | None]."""
        c = self.cfg
        B = len(reqs)
        assert 1 <= B <= self.max_batch, f"{B} sequences, engine built for {self.max_batch}"
        boi, eoi = self.img_ids_h[0], self.img_ids_h[-1]
        n_img = len(self.img_ids_h) - 2
        run = [boi] + self.img_ids_h[1:-1] + [eoi]
        gens, hids, chunk0, Ls, scheds, maxnew = [], [], [], [], [], []
        for b, r in enumerate(reqs):
            if r is None:     # an empty slot (a story that has ended): rides along masked, result None
                self.reset_sequence(b)
                gens.append([self.eos_id]); hids.append([]); chunk0.append(None); Ls.append(1)
                scheds.append([-1] * self.max_new); maxnew.append(0)
                continue
            ids, emb = list(r["input_ids"]), r["inputs_embeds"]
            L = len(ids)
            mx = int(r["max_new_tokens"])
            past_len, head = r.get("past_len"), int(r.get("head", 0) or 0)
            n_cached = 0 if past_len is None else int(past_len)
            if past_len is None:
                head = 0
            assert 0 <= head < L, "at least one prompt token must be fed on top of the cache"
            total = n_cached + (L - head) + mx + 1
            if mx + 1 > self.max_new or total > self.max_pages * PAGE or L + mx + 1 > c.max_pos:
                raise _capi.SeedStoryError(f"generate_batch[{b}]({L} prompt + {mx} new tokens on {n_cached} cached) exceeds the "
                                           f"engine's capacity (max_new {self.max_new}, max_ctx {self.max_pages * PAGE}, "
                                           f"max_pos {c.max_pos})")
            if past_len is None:
                self.reset_sequence(b)
            else:
                assert self.seq_len_h[b] == n_cached, (self.seq_len_h[b], n_cached)
            hn, logits = self.forward_chunk(b, emb[head:], list(range(head, L)))
            sched = [-1] * self.max_new
            if r.get("schedule") is not None:
                sched[:len(r["schedule"])] = r["schedule"]
            gens.append([self.first_token(logits, int(ids[-1]), sched[0])])
            hids.append([])
            chunk0.append(hn)
            Ls.append(L)
            scheds.append(sched)
            maxnew.append(mx)
        sched_t = torch.tensor(scheds, dtype=torch.int32)

        def active(b):
            return gens[b][-1] != self.eos_id and len(gens[b]) < maxnew[b]

        def chunkable(b):
            g, sc = gens[b], scheds[b]
            return (chunk_image_run and g[-1] == boi and len(g) + n_img + 1 < maxnew[b]
                    and all(x < 0 for x in sc[len(g):len(g) + n_img + 1]))

        while any(active(b) for b in range(B)):
            did_chunk = False
            for b in range(B):
                if active(b) and chunkable(b):
                    # the next n_img+1 ids are input-determined (generation.py:23-26): one tensor-core chunk
                    g = gens[b]
                    emb = self.embed_tokens(torch.tensor(run))
                    p0 = Ls[b] + len(g) - 1
                    hn_c, logits = self.forward_chunk(b, emb, list(range(p0, p0 + len(run))))
                    hids[b].append(hn_c)
                    g.extend(run[1:])
                    g.append(self.first_token(logits, eoi, scheds[b][len(g)] if len(g) < len(scheds[b]) else -1))
                    did_chunk = True
            if did_chunk:
                continue
            # token-by-token phase for every sequence that is still running; the others ride along masked (done = 1):
            # their state is not advanced and whatever the step writes for them lands in slots nobody has read yet
            live = [b for b in range(B) if active(b)]
            self.begin_decode([gens[b][-1] for b in range(B)], [Ls[b] + len(gens[b]) - 1 for b in range(B)], sched_t[:B])
            self.n_out[:B].copy_(torch.tensor([len(gens[b]) for b in range(B)], dtype=torch.int32))
            self.done[:B].copy_(torch.tensor([0 if b in live else 1 for b in range(B)], dtype=torch.int32))
            while True:
                self.decode_step(B, use_graph)
                ids, _ = self.read_step(B)
                stop = False
                for b in live:
                    self.seq_len_h[b] += 1
                    hids[b].append(self.hist[b, len(gens[b]):len(gens[b]) + 1].clone())
                    gens[b].append(ids[b])
                    if not active(b) or chunkable(b):
                        stop = True      # a sequence ended or reached an image run: re-plan on the host
                if stop:
                    break
        out = []
        for b in range(B):
            if reqs[b] is None:
                out.append(None)
                continue
            hidden = torch.cat(hids[b], 0) if hids[b] else torch.empty((0, c.hidden), dtype=torch.float16, device=self.dev)
            out.append((gens[b], hidden, chunk0[b]))
        return out


class RetainedKV:
    """Engine-native `past_key_values`: "the cache of sequence b exactly as it is now" (typically right after
    LlamaEngine.retain_tokens applied the window / attention-sink policy).  Handing this to generate() costs nothing;
    a tuple of (K, V) tensors (the reference's data structure) is accepted as well and copied in by load_past()."""

    def __init__(self, engine, b=0):
        self.e, self.b, self.n = engine, b, engine.seq_len_h[b]


class PagedKVView:
    """Lazy stand-in for the reference's tuple-of-(K, V) `past_key_values` (models.py:156, 220): indexing layer l
    gathers that layer's K/V of sequence b out of the page pool as [1, heads, n, head_dim] tensors."""

    def __init__(self, engine, b):
        self.e, self.b = engine, b
        self.n = engine.seq_len_h[b]
        self.pages = engine.page_table_h[b, :(self.n + PAGE - 1) // PAGE].clone().long()

    def __len__(self):
        return self.e.cfg.layers

    def __getitem__(self, l):
        e = self.e
        pg = self.pages.to(e.dev)

        def g(pool):
            t = pool[l][pg]  # [np, H, 64, D]
            return t.permute(1, 0, 2, 3).reshape(e.cfg.heads, -1, e.cfg.head_dim)[:, :self.n].unsqueeze(0)
        return (g(e.k_pages), g(e.v_pages))

    def __iter__(self):
        for l in range(len(self)):
            yield self[l]


def sink_retained_slots(seq_len, evicted_images, live_from, n_sink=4, head=(4, 8), tail=(8, 4)):
    """Retention set of the multimodal attention sink (vis_george_sink.py:266-291, SURVEY.md §7 "Sink semantics"):
    slots {0..n_sink-1} U for every evicted image (boi, eoi): [boi-4, boi+8) U [eoi-8, eoi+4) U the live tail
    [live_from, seq_len).  Returns a sorted list of slot indices."""
    keep = set(range(min(n_sink, seq_len)))
    for boi, eoi in evicted_images:
        keep.update(range(max(0, boi - head[0]), min(seq_len, boi + head[1])))
        keep.update(range(max(0, eoi - tail[0]), min(seq_len, eoi + tail[1])))
    keep.update(range(live_from, seq_len))
    return sorted(keep)
