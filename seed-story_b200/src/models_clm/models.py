"""Drop-in for the reference's src/models_clm/models.py::ContinuousLVLM (reference :20-230; hydra target
configs/clm_models/agent_7b_sft.yaml:1) — inference surface only (`generate`, `from_pretrained`).

generate() keeps the reference's signature, return dict (reference :213-221) and semantics:
  embed lookup (:127) -> input_resampler on the comprehension images (:133) -> scatter into the <img_i> slots
  (:135) -> greedy generation with the image-token processor (:146-153) -> hidden rows of the 64 image queries
  before the LAST </img> (:182-197) -> output_resampler (:205) -> tokenizer.decode (:211)
with every arithmetic step on the seedstory_b200 kernels.  `past_key_values=` (live sink-KV mode) is honoured: a tuple
of per-layer (K, V) tensors as vis_george_sink.py:266-291 builds it, or llama_engine.RetainedKV.  The training `forward` (reference :33-96) is out of scope.
"""
import os

import torch
from torch import nn

from .generation import AutoImageTokenGenerationProcessor

BOI_TOKEN = '<img>'
EOI_TOKEN = '</img>'
IMG_TOKEN = '<img_{:05d}>'


class ContinuousLVLM(nn.Module):
    def __init__(self, llm, input_resampler, output_resampler, lm_loss_scale=1.0, rec_loss_scale=1.0) -> None:
        super().__init__()
        self.llm = llm
        self.input_resampler = input_resampler
        self.output_resampler = output_resampler
        self.lm_loss_scale = lm_loss_scale
        self.rec_loss_scale = rec_loss_scale

    def forward(self, *args, **kwargs):
        raise NotImplementedError("training forward (reference models.py:33-96) is outside the inference hot path")

    def _input_embeds(self, input_ids, image_embeds, embeds_cmp_mask, ids_cmp_mask):
        """Token embeddings with the resampled image tokens scattered into the <img_i> slots (models.py:127-135)."""
        from seedstory import ops
        input_embeds = self.llm.get_input_embeddings()(input_ids)
        dim = input_embeds.shape[-1]
        if image_embeds is not None:
            assert embeds_cmp_mask is not None and ids_cmp_mask is not None
            image_embeds_lm = self.input_resampler(image_embeds)[embeds_cmp_mask]
            rows = torch.nonzero(ids_cmp_mask.reshape(-1)).reshape(-1).to(torch.int32)
            src = image_embeds_lm.reshape(-1, dim).to(input_embeds.dtype).contiguous()
            assert rows.numel() == src.shape[0], "number of <img_i> slots must match the resampled image tokens"
            ops.scatter_rows(src, rows.to(input_embeds.device), input_embeds.view(-1, dim))
        return input_embeds

    def _postprocess(self, tokenizer, input_ids, output, past_key_values, num_img_gen_tokens, output_past_key_values):
        """models.py:156-221: slice the generated ids, find the last </img>, gather the 64 hidden rows before it, run the
        output resampler, decode the text."""
        generate_ids = output.sequences[0][input_ids.shape[1]:]
        eoi_token_id = tokenizer.encode(EOI_TOKEN, add_special_tokens=False)[0]
        attn_weights = ()
        rows = torch.cat([step[-1] for step in output.hidden_states], dim=1)[0]
        if past_key_values is None:
            # rows of the prompt forward are dropped; row j belongs to the position fed with generated id j (:184-185)
            last_hidden_states = rows[input_ids.shape[1]:]
            eoi_indices = [j for j, t in enumerate(generate_ids.tolist()) if t == eoi_token_id]
        else:
            # KV-reuse branch (:186-189): every row of this call is kept (the fed prompt tail + the generated
            # positions) and </img> is searched in the LAST len(rows) ids of the sequence — i.e. shifted by one
            # against the rows, exactly as the reference does it
            last_hidden_states = rows
            tail = output.sequences[0][-rows.shape[0]:].tolist()
            eoi_indices = [j for j, t in enumerate(tail) if t == eoi_token_id]
        num_gen_imgs = 1 if len(eoi_indices) > 0 else 0
        has_img_output = num_gen_imgs > 0
        if has_img_output:
            e = eoi_indices[-1]   # the LAST </img> wins (reference :197)
            img_gen_feats = last_hidden_states[e - num_img_gen_tokens:e].unsqueeze(0).contiguous()
            img_gen_feat = self.output_resampler(img_gen_feats)
        else:
            img_gen_feat = None
        generate_text = tokenizer.decode(generate_ids, skip_special_tokens=False)
        return {
            'text': generate_text,
            'generate_ids': generate_ids,
            'has_img_output': has_img_output,
            'img_gen_feat': img_gen_feat,
            'num_gen_imgs': num_gen_imgs,
            'attn_weights': attn_weights,
            'past_key_values': output_past_key_values
        }

    @torch.no_grad()
    def generate(self, tokenizer, prompt=None, input_ids=None, image_embeds=None, embeds_cmp_mask=None,
                 ids_cmp_mask=None, logits_processor=None, num_img_gen_tokens=64, temperature=0.7, num_beams=1,
                 max_new_tokens=120, top_p=0.5, past_key_values=None, dtype=torch.float16, device='cuda'):
        if logits_processor is None:
            logits_processor = [AutoImageTokenGenerationProcessor(tokenizer=tokenizer,
                                                                  num_img_gen_tokens=num_img_gen_tokens)]
        if prompt is not None:
            input_ids = tokenizer(prompt, return_tensors="pt").input_ids
        if isinstance(input_ids, list):
            input_ids = torch.tensor(input_ids)
        input_ids = input_ids.to(device=device)
        input_embeds = self._input_embeds(input_ids, image_embeds, embeds_cmp_mask, ids_cmp_mask)
        output = self.llm.generate(input_ids=input_ids, inputs_embeds=input_embeds, output_hidden_states=True,
                                   return_dict_in_generate=True, logits_processor=logits_processor,
                                   past_key_values=past_key_values, temperature=temperature, num_beams=num_beams,
                                   max_new_tokens=max_new_tokens, top_p=top_p, do_sample=False,
                                   eos_token_id=getattr(tokenizer, "eos_token_id", 2) or 2)
        return self._postprocess(tokenizer, input_ids, output, past_key_values, num_img_gen_tokens,
                                 self.llm.past_key_values)

    @torch.no_grad()
    def generate_batch(self, tokenizer, requests, logits_processor=None, num_img_gen_tokens=64, max_new_tokens=120,
                       device='cuda'):
        """generate() for several independent stories whose MLLM decode steps are batched (one pass over the weights per
        step for all of them; BASELINE configs[3]).  requests[b] = dict(input_ids, image_embeds, embeds_cmp_mask,
        ids_cmp_mask[, past_len, head]) or None for an empty slot; returns the list of generate()'s result dicts (None for
        empty slots; 'past_key_values' is not materialised)."""
        if logits_processor is None:
            logits_processor = [AutoImageTokenGenerationProcessor(tokenizer=tokenizer,
                                                                  num_img_gen_tokens=num_img_gen_tokens)]
        ids_l, emb_l, past_lens, heads = [], [], [], []
        for r in requests:
            if r is None:       # an empty slot (a story that has ended)
                ids_l.append(None); emb_l.append(None); past_lens.append(None); heads.append(0)
                continue
            ids = r["input_ids"].to(device=device)
            ids_l.append(ids)
            emb_l.append(self._input_embeds(ids, r.get("image_embeds"), r.get("embeds_cmp_mask"), r.get("ids_cmp_mask")))
            past_lens.append(r.get("past_len"))     # live KV reuse: tokens already cached in the story's engine slot
            heads.append(r.get("head", 0))          # ... and the first prompt token that still has to be fed
        outs = self.llm.generate_batch(ids_l, emb_l, logits_processor=logits_processor, max_new_tokens=max_new_tokens,
                                       eos_token_id=getattr(tokenizer, "eos_token_id", 2) or 2, past_lens=past_lens,
                                       heads=heads)
        return [None if outs[b] is None else
                self._postprocess(tokenizer, ids_l[b], outs[b], None if past_lens[b] is None else True, num_img_gen_tokens, None)
                for b in range(len(outs))]

    @classmethod
    def from_pretrained(cls, llm, input_resampler, output_resampler, pretrained_model_path=None, **kwargs):
        model = cls(llm=llm, input_resampler=input_resampler, output_resampler=output_resampler, **kwargs)
        if pretrained_model_path is not None:
            if os.path.exists(pretrained_model_path):
                ckpt = torch.load(pretrained_model_path, map_location='cpu')
                missing, unexpected = model.load_state_dict(ckpt, strict=False)
                print('agent model, missing keys: ', len(missing), 'unexpected keys:', len(unexpected))
            else:
                from seedstory.synthetic import missing_checkpoint
                missing_checkpoint(pretrained_model_path, "ContinuousLVLM (agent)")
        return model
