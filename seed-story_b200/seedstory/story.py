"""End-to-end interleaved story generation through the reference-facing API (the `src.*` drop-in modules).

Mirrors the call sequence of the reference's scripts —
  src/inference/gen_george.py:152-270 (turn loop, window of 8 images :235-239) with the append-ids prompt
  bookkeeping of src/inference/vis_george_sink.py:247-263 —
but as a function, with synthetic StoryStream-shaped inputs: no checkpoints, tokenizer files or
datasets exist offline, so weights are seeded random tensors of the real shapes and the turn structure
(64 text tokens -> <img> -> 64 image queries -> </img> -> EOS) is forced through the `logits_processor=` hook
that ContinuousLVLM.generate exposes.  Text tokens are still the model's true greedy choices.
"""
import time

import torch

BOI_TOKEN = '<img>'
EOI_TOKEN = '</img>'
IMG_TOKEN = '<img_{:05d}>'


class SyntheticTokenizer:
    """Stand-in for pretrained/cvlm_llama2_tokenizer (32000 Llama ids + <img>, </img>, <img_00000..63>).
    Id assignment follows SURVEY.md §7: BOI=32000, EOI=32001, IMG_i=32002+i."""
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def __init__(self, vocab=32066, n_img=64):
        self.vocab, self.n_img = vocab, n_img
        self.boi, self.eoi, self.img0 = vocab - n_img - 2, vocab - n_img - 1, vocab - n_img
        self._special = {BOI_TOKEN: self.boi, EOI_TOKEN: self.eoi}
        for i in range(n_img):
            self._special[IMG_TOKEN.format(i)] = self.img0 + i
        self._inv = {v: k for k, v in self._special.items()}

    def encode(self, text, add_special_tokens=False):
        import re
        ids = []
        for piece in re.split(r"(<img_\d{5}>|</img>|<img>)", text):
            if not piece:
                continue
            if piece in self._special:
                ids.append(self._special[piece])
            else:  # "words" are space-separated decimal ids, e.g. "t17 t905"
                ids.extend(int(w[1:]) for w in piece.split() if w.startswith("t"))
        return ids

    def decode(self, ids, skip_special_tokens=False):
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        return " ".join(self._inv.get(i, f"t{i}") for i in ids)


FULL = dict(
    llama=dict(hidden_size=4096, intermediate_size=11008, num_attention_heads=32, num_hidden_layers=32,
               vocab_size=32000, rms_norm_eps=1e-5),
    vocab=32066,
    vit=dict(heads=16, image_size=448, layers=48, mlp_ratio=4.9231, output_dim=4096, patch_size=14, width=1664),
    agent_dim=4096, agent_heads=32,
    xl=dict(dim=1024, depth=4, dim_head=64, heads=16, num_queries=64, embedding_dim=4096, output1_dim=768,
            output2_dim=1280, ff_mult=4),
    unet=None, vae=None, image=1024,
)

# reduced-size configuration for tests (same structure, every kernel path exercised)
TINY = dict(
    llama=dict(hidden_size=256, intermediate_size=352, num_attention_heads=2, num_hidden_layers=2,
               vocab_size=254, rms_norm_eps=1e-5),
    vocab=320,
    vit=dict(heads=4, image_size=56, layers=2, mlp_ratio=4.0, output_dim=256, patch_size=14, width=64,
             n_queries=256),
    agent_dim=256, agent_heads=2,
    xl=dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=64, embedding_dim=256, output1_dim=96,
            output2_dim=160, ff_mult=4),
    unet=dict(block_out_channels=(64, 128, 256), num_attention_heads=(1, 2, 4), transformer_layers_per_block=(0, 1, 2),
              cross_attention_dim=256, projection_class_embeddings_input_dim=160 + 6 * 32, addition_time_embed_dim=32,
              sample_size=64),
    vae=dict(block_out_channels=(64, 64, 128, 128)), image=512,
)


class StoryPipeline:
    def __init__(self, device="cuda:0", cfg=None, seed=1234, num_inference_steps=50, n_text_tokens=64,
                 window_size=8, verbose=False):
        import diffusers
        from src.models.discrete_models import DiscreteModleIdentity
        from src.models.qwen_visual import Resampler, VisionTransformerWithAttnPool
        from src.models_clm.modeling_llama_xformer import LlamaForCausalLM
        from src.models_clm.models import ContinuousLVLM
        from src.models_clm.peft_models import LoraConfig, get_peft_model_with_resize_embedding
        from src.models_ipa.adapter_modules import SDXLAdapter
        from src.models_ipa.resampler import ResamplerXLV2
        cfg = cfg or FULL
        self.cfg, self.dev = cfg, torch.device(device)
        self.steps, self.n_text, self.window = num_inference_steps, n_text_tokens, window_size
        torch.manual_seed(seed)
        dt = torch.float16
        t0 = time.time()
        with torch.device(self.dev):
            self.tokenizer = SyntheticTokenizer(cfg["vocab"], 64)
            self.visual_encoder = VisionTransformerWithAttnPool(**cfg["vit"]).eval().to(dtype=dt)
            llama = LlamaForCausalLM(cfg["llama"]).to(dtype=dt)
            peft_cfg = LoraConfig(r=16, lora_alpha=32, target_modules=["q_proj", "v_proj", "k_proj", "o_proj",
                                                                       "gate_proj", "down_proj", "up_proj"],
                                  modules_to_save=["input_layernorm", "post_attention_layernorm", "norm"],
                                  lora_dropout=0.05, task_type="CAUSAL_LM")
            llm = get_peft_model_with_resize_embedding(llama, peft_config=peft_cfg, vocab_size=cfg["vocab"],
                                                       torch_dtype="fp16")
            # peft initialises lora_B to zero; give it N(0, 0.02^2) so the LoRA path carries signal
            for n, p in llm.named_parameters():
                if "lora_B" in n:
                    torch.nn.init.normal_(p, std=0.02)
            E, Hh = cfg["agent_dim"], cfg["agent_heads"]
            self.agent = ContinuousLVLM(llm, Resampler(8, E, Hh, kv_dim=E), Resampler(16, E, Hh, kv_dim=E)).eval().to(dtype=dt)
            self.scheduler = diffusers.EulerDiscreteScheduler()
            self.vae = diffusers.AutoencoderKL(config=cfg["vae"]).to(self.dev, dtype=dt)
            self.unet = diffusers.UNet2DConditionModel(config=cfg["unet"]).to(self.dev, dtype=dt)
            self.adapter = SDXLAdapter(self.unet, ResamplerXLV2(**cfg["xl"])).to(self.dev, dtype=dt).eval()
            self.discrete = DiscreteModleIdentity().eval()
        self.adapter.init_pipe(vae=self.vae, scheduler=self.scheduler, visual_encoder=self.visual_encoder,
                               image_transform=None, discrete_model=self.discrete, dtype=dt, device=self.dev)
        tk = self.tokenizer
        self.image_ids = [tk.boi] + [tk.img0 + i for i in range(64)] + [tk.eoi]
        if verbose:
            print(f"[story] models built in {time.time() - t0:.1f}s")

    def _schedule(self):
        """text tokens free (greedy), then <img> forced; the 64 queries + </img> come from the image-token processor;
        EOS forced afterwards.  EOS and <img> are suppressed in the free slots (HF SuppressTokensLogitsProcessor
        semantics): with random weights a greedy text token can be any id, and an early EOS would end the turn before
        its image (the reference then ends the story, gen_george.py:208) while an early <img> would add a second
        image run — either would make turns of different stories incomparable."""
        from src.models_clm.generation import ForcedScheduleProcessor, SuppressTokensProcessor
        tk = self.tokenizer
        sched = [-1] * self.n_text + [tk.boi] + [-1] * 65 + [tk.eos_token_id]
        return [ForcedScheduleProcessor(sched), SuppressTokensProcessor([tk.eos_token_id, tk.boi])]

    @torch.no_grad()
    def run_story(self, image_tensor, caption_ids, n_turns, decode_images=True, return_images=False, overlap=False,
                  sink=False):
        """image_tensor [1,3,S,S] fp16 on the device (CLIP-normalised); caption_ids: list[int].
        Returns list of per-turn dicts(generate_ids, image_uint8 | None).

        sink=True: LIVE multimodal attention-sink mode — what src/inference/vis_george_sink.py
        :243-295 prepares but never switches on (it passes past_key_values=None, :316): the KV cache of the prompt is
        kept across turns (`use_kv_cache_head=True`, kv_cache_head = previous prompt length, modeling_llama_xformer.py
        :804-826), only the new tail [text + <img>…</img>] is fed, and when the image window overflows the cache is cut
        to {first 4 slots} U per evicted image [<img>-4, <img>+8) U [</img>-8, </img>+4) U live tail
        (llama_engine.sink_retained_slots); retained keys keep their RoPE phase, new tokens get window-relative
        positions.  The image features follow the reference's KV-reuse branch (models.py:186-197).

        overlap=True (off by default: measured neutral on one B200, both phases already fill the GPU) issues the SDXL de-tokenizer of turn t on a side stream while the MLLM already decodes turn
        t+1: the next turn only needs `img_gen_feat` (gen_george.py:224), the pixels are merely saved
        (gen_george.py:210-222) — SURVEY.md §8f rank 1.  Results are identical to the sequential order."""
        from src.models_clm.generation import AutoImageTokenGenerationProcessor
        tk, dev = self.tokenizer, self.dev
        input_ids = [tk.bos_token_id] + list(caption_ids) + self.image_ids
        image_embeds = self.visual_encoder(image_tensor)
        procs = [AutoImageTokenGenerationProcessor(tk, 64)] + self._schedule()
        outs = []
        res = self.cfg["image"]
        main = torch.cuda.current_stream()
        if overlap and decode_images:
            if not hasattr(self, "_side"):
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
            side.wait_stream(main)
        else:
            side = None
        model = self.agent.llm.base_model.model
        model.use_kv_cache_head, model.kv_cache_head, past = bool(sink), None, None
        n_sink = 0          # retained sink slots in front of the windowed prompt inside the cache
        for turn in range(n_turns):
            ids_t = torch.tensor([input_ids], dtype=torch.long, device=dev)
            boi = [i for i, t in enumerate(input_ids) if t == tk.boi]
            eoi = [i for i, t in enumerate(input_ids) if t == tk.eoi]
            ids_cmp_mask = torch.zeros_like(ids_t, dtype=torch.bool)
            for i in range(image_embeds.shape[0]):
                ids_cmp_mask[0, boi[i] + 1:eoi[i]] = True
            embeds_cmp_mask = torch.ones(image_embeds.shape[0], dtype=torch.bool, device=dev)
            out = self.agent.generate(tokenizer=tk, input_ids=ids_t, image_embeds=image_embeds,
                                      embeds_cmp_mask=embeds_cmp_mask, ids_cmp_mask=ids_cmp_mask,
                                      max_new_tokens=500, num_img_gen_tokens=64, logits_processor=procs, device=dev,
                                      past_key_values=past)
            if not out["has_img_output"]:
                # reference behaviour (gen_george.py:208 `while output['has_img_output'] and …`): a turn without an
                # image run ends the story; the caller counts the turns actually produced
                outs.append(dict(generate_ids=out["generate_ids"].tolist(), image=None, has_img_output=False))
                break
            img = None
            if decode_images:
                feat = out["img_gen_feat"]
                if side is not None:
                    side.wait_stream(main)       # img_gen_feat was produced on the main stream
                    feat.record_stream(side)
                    with torch.cuda.stream(side):
                        imgs = self.adapter.generate(image_embeds=feat, num_inference_steps=self.steps, height=res,
                                                     width=res, output_type="pt",
                                                     input_image_size=self.cfg["vit"]["image_size"])
                else:
                    imgs = self.adapter.generate(image_embeds=feat, num_inference_steps=self.steps, height=res,
                                                 width=res, output_type="pt",
                                                 input_image_size=self.cfg["vit"]["image_size"])
                img = imgs[0]
            gen = out["generate_ids"].tolist()
            outs.append(dict(generate_ids=gen, image=img if return_images else None, has_img_output=True))
            image_embeds = torch.cat((image_embeds, out["img_gen_feat"]), dim=0)
            text_ids = [t for t in gen if t < tk.boi and t != tk.eos_token_id]
            if not sink:
                input_ids = input_ids + text_ids + self.image_ids
                while image_embeds.shape[0] > self.window:  # evict the oldest image and all text before it
                    first_eoi = input_ids.index(tk.eoi)
                    input_ids = [tk.bos_token_id] + input_ids[first_eoi + 1:]
                    image_embeds = image_embeds[1:]
                continue
            # ---- live sink mode: the cache holds [n_sink sink slots | this turn's prompt | generated ids]; keep the
            # prompt part (vis_george_sink.py:243-244), append the next tail to the ids (:247-249)
            from . import llama_engine
            eng = self.agent.llm.engine()
            input_ids, image_embeds, n_sink, L_prev = self._sink_advance(eng, 0, input_ids, text_ids, image_embeds, n_sink)
            model.kv_cache_head = L_prev
            past = llama_engine.RetainedKV(eng, 0)
        if side is not None:
            main.wait_stream(side)               # every image is complete before the caller touches the results
        return outs

    def _sink_advance(self, eng, b, input_ids, text_ids, image_embeds, n_sink):
        """End-of-turn bookkeeping of the live attention-sink mode for the story in engine slot b: cut the cache back to
        the prompt, append the next tail to the ids, and when the image window overflows keep only the sink slots
        (llama_engine.sink_retained_slots) of every evicted image.  Returns (input_ids, image_embeds, n_sink, L_prev)."""
        from . import llama_engine
        tk = self.tokenizer
        L_prev = len(input_ids)
        eng.truncate(b, n_sink + L_prev)
        input_ids = input_ids + text_ids + self.image_ids
        while image_embeds.shape[0] > self.window:
            b0, e0 = input_ids.index(tk.boi), input_ids.index(tk.eoi)        # oldest image, windowed indices
            cache_len = n_sink + L_prev
            fresh = llama_engine.sink_retained_slots(cache_len, [(n_sink + b0, n_sink + e0)], n_sink + e0 + 1,
                                                     n_sink=4 if n_sink == 0 else 0)
            keep = sorted(set(range(n_sink)) | set(fresh))
            n_live = cache_len - (n_sink + e0 + 1)
            input_ids = input_ids[e0 + 1:]
            image_embeds = image_embeds[1:]
            L_prev -= e0 + 1
            eng.retain_tokens(b, keep)            # compaction: slots are contiguous again, sink slots in front
            n_sink = len(keep) - n_live
        return input_ids, image_embeds, n_sink, L_prev

    @torch.no_grad()
    def run_stories(self, image_tensors, captions, n_turns, decode_images=True, return_images=False, sink=False):
        """Several independent stories per GPU with their MLLM decode steps BATCHED (continuous batching over the paged KV
        cache, BASELINE configs[3]: every decode step streams the weights once for all stories); per story the results are
        those of run_story().  image_tensors: list of [1,3,S,S]; captions: list of id lists.  Story i lives in engine slot
        i for the whole call (in sink mode its cache stays there across turns).  A story whose turn emits no image ends
        (gen_george.py:208) and simply stops riding along.  Returns a list (per story) of run_story()-style lists."""
        from src.models_clm.generation import AutoImageTokenGenerationProcessor
        tk, dev = self.tokenizer, self.dev
        S = len(captions)
        procs = [AutoImageTokenGenerationProcessor(tk, 64)] + self._schedule()
        res = self.cfg["image"]
        eng = self.agent.llm.engine(max_batch=S)
        st = []
        for i in range(S):
            st.append(dict(ids=[tk.bos_token_id] + list(captions[i]) + self.image_ids,
                           emb=self.visual_encoder(image_tensors[i]), outs=[], alive=True, n_sink=0, past_len=None, head=0))
        for turn in range(n_turns):
            reqs = []
            for s_ in st:
                if not s_["alive"]:
                    reqs.append(None)        # the story keeps its engine slot; the hole rides along masked
                    continue
                ids_t = torch.tensor([s_["ids"]], dtype=torch.long, device=dev)
                boi = [i for i, t in enumerate(s_["ids"]) if t == tk.boi]
                eoi = [i for i, t in enumerate(s_["ids"]) if t == tk.eoi]
                mask = torch.zeros_like(ids_t, dtype=torch.bool)
                for i in range(s_["emb"].shape[0]):
                    mask[0, boi[i] + 1:eoi[i]] = True
                reqs.append(dict(input_ids=ids_t, image_embeds=s_["emb"], ids_cmp_mask=mask,
                                 embeds_cmp_mask=torch.ones(s_["emb"].shape[0], dtype=torch.bool, device=dev),
                                 past_len=s_["past_len"], head=s_["head"]))
            results = self.agent.generate_batch(tk, reqs, logits_processor=procs, max_new_tokens=500, num_img_gen_tokens=64,
                                                device=dev)
            for b, (s_, out) in enumerate(zip(st, results)):
                if out is None:
                    continue
                if not out["has_img_output"]:
                    s_["outs"].append(dict(generate_ids=out["generate_ids"].tolist(), image=None, has_img_output=False))
                    s_["alive"] = False
                    continue
                img = None
                if decode_images:
                    img = self.adapter.generate(image_embeds=out["img_gen_feat"], num_inference_steps=self.steps, height=res,
                                                width=res, output_type="pt", input_image_size=self.cfg["vit"]["image_size"])[0]
                gen = out["generate_ids"].tolist()
                s_["outs"].append(dict(generate_ids=gen, image=img if return_images else None, has_img_output=True))
                s_["emb"] = torch.cat((s_["emb"], out["img_gen_feat"]), dim=0)
                text_ids = [t for t in gen if t < tk.boi and t != tk.eos_token_id]
                if not sink:
                    s_["ids"] = s_["ids"] + text_ids + self.image_ids
                    while s_["emb"].shape[0] > self.window:
                        first_eoi = s_["ids"].index(tk.eoi)
                        s_["ids"] = [tk.bos_token_id] + s_["ids"][first_eoi + 1:]
                        s_["emb"] = s_["emb"][1:]
                else:
                    s_["ids"], s_["emb"], s_["n_sink"], L_prev = self._sink_advance(eng, b, s_["ids"], text_ids, s_["emb"],
                                                                                    s_["n_sink"])
                    s_["head"], s_["past_len"] = L_prev, eng.seq_len_h[b]
            if not any(s_["alive"] for s_ in st):
                break
        return [s_["outs"] for s_ in st]
