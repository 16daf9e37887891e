"""Drop-in for src/models_clm/peft_models.py::get_peft_model_with_resize_embedding (reference :21-65) and for
the slice of peft 0.4.0 it relies on (requirements.txt:24): the returned object exposes `.base_model.model`
(the LlamaForCausalLM), `.generate`, `.get_input_embeddings`, `.past_key_values`, and a state_dict whose keys
follow peft's layout (`base_model.model.model.layers.N.self_attn.q_proj.lora_A.default.weight`,
`…input_layernorm.modules_to_save.default.weight`, …) so the agent checkpoint loads.

LoRA is folded into the base weights when the engine packs them (fp32 accumulate, one rounding).
"""
import torch
from torch import nn

from seedstory import llama_engine

from .generation import AutoImageTokenGenerationProcessor, ForcedScheduleProcessor, SuppressTokensProcessor


class LoraConfig:
    """Minimal stand-in for peft.LoraConfig (hydra target in configs/clm_models/llama2chat7b_lora.yaml:7-27)."""

    def __init__(self, r=8, lora_alpha=8, target_modules=None, lora_dropout=0.0, modules_to_save=None,
                 task_type=None, **kwargs):
        self.r, self.lora_alpha, self.lora_dropout = r, lora_alpha, lora_dropout
        self.target_modules = list(target_modules or [])
        self.modules_to_save = list(modules_to_save or [])
        self.task_type = task_type


class _LoraLinear(nn.Module):
    """Parameter layout of peft 0.4 lora.Linear: .weight, .lora_A.default.weight [r,in], .lora_B.default.weight [out,r]."""

    def __init__(self, base, r, alpha):
        super().__init__()
        self.weight = base.weight
        self.in_features, self.out_features = base.in_features, base.out_features
        kw = dict(device=base.weight.device, dtype=base.weight.dtype)
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False, **kw)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False, **kw)})
        nn.init.zeros_(self.lora_B["default"].weight)  # peft init: B = 0
        self.scaling = alpha / r


class _ModulesToSave(nn.Module):
    """peft ModulesToSaveWrapper layout: .original_module / .modules_to_save.default (the active copy)."""

    def __init__(self, mod):
        super().__init__()
        import copy
        self.original_module = mod
        self.modules_to_save = nn.ModuleDict({"default": copy.deepcopy(mod)})

    @property
    def weight(self):
        return self.modules_to_save["default"].weight


class _LoraModel(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model


class GenerateOutput:
    def __init__(self, sequences, hidden_states, attentions=None):
        self.sequences, self.hidden_states, self.attentions = sequences, hidden_states, attentions


class PeftModelForCausalLM(nn.Module):
    def __init__(self, model, peft_config):
        super().__init__()
        self.peft_config = peft_config
        for layer in model.model.layers:
            for parent in (layer.self_attn, layer.mlp):
                for name in list(parent._modules):
                    if name in peft_config.target_modules:
                        setattr(parent, name, _LoraLinear(getattr(parent, name), peft_config.r, peft_config.lora_alpha))
            for name in ("input_layernorm", "post_attention_layernorm"):
                if name in peft_config.modules_to_save:
                    setattr(layer, name, _ModulesToSave(getattr(layer, name)))
        if "norm" in peft_config.modules_to_save:
            model.model.norm = _ModulesToSave(model.model.norm)
        self.base_model = _LoraModel(model)
        self._engine = None
        self._engine_key = None

    # peft forwards unknown attributes to the wrapped model (models.py:156 reads llm.past_key_values)
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model.model, name)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def get_input_embeddings(self):
        return self.base_model.model.get_input_embeddings()

    def get_output_embeddings(self):
        return self.base_model.model.get_output_embeddings()

    def print_trainable_parameters(self):
        n = sum(p.numel() for p in self.parameters())
        print(f"all params: {n}")

    # ---------------------------------------------------------------- engine
    def engine(self, max_new=512, max_ctx=4096, max_batch=1):
        m = self.base_model.model
        dev = m.lm_head.weight.device
        if self._engine is not None and (self._engine.max_new < max_new or self._engine.max_batch < max_batch
                                         or self._engine.max_pages * llama_engine.PAGE < max_ctx):
            # capacity of the cached engine is too small for this call: rebuild it (never shrink)
            max_new, max_batch = max(max_new, self._engine.max_new), max(max_batch, self._engine.max_batch)
            max_ctx = max(max_ctx, self._engine.max_pages * llama_engine.PAGE)
            self._engine = None
            torch.cuda.empty_cache()
        if self._engine is None:
            eng = llama_engine.LlamaEngine(m.engine_config(), dev, max_batch=max_batch, max_ctx=max_ctx, max_new=max_new)
            layers = []
            for layer in m.model.layers:
                L = {}
                for parent in (layer.self_attn, layer.mlp):
                    for name, mod in parent._modules.items():
                        L[name] = mod.weight.data
                        if isinstance(mod, _LoraLinear):
                            L[name + ".lora_A"] = mod.lora_A["default"].weight.data
                            L[name + ".lora_B"] = mod.lora_B["default"].weight.data
                L["input_layernorm"] = layer.input_layernorm.weight.data
                L["post_attention_layernorm"] = layer.post_attention_layernorm.weight.data
                layers.append(L)
            scaling = self.peft_config.lora_alpha / self.peft_config.r
            eng.load_weights(m.model.embed_tokens.weight.data, layers, m.model.norm.weight.data, m.lm_head.weight.data,
                             lora_scaling=scaling)
            self._engine = eng
        return self._engine

    def _configure_engine(self, eng, logits_processor, eos_token_id):
        """Map the `logits_processor=` list of models.py:146-153 onto the engine's device-side implementations."""
        img_ids, schedule, suppress = None, None, []
        for proc in (logits_processor or []):
            if isinstance(proc, ForcedScheduleProcessor):
                schedule = proc.schedule
            elif isinstance(proc, SuppressTokensProcessor) or hasattr(proc, "suppress_tokens"):
                suppress = list(proc.suppress_tokens)
            elif hasattr(proc, "img_ids_list"):
                img_ids = proc.img_ids_list
            else:
                raise NotImplementedError(f"logits processor {type(proc).__name__} has no device implementation")
        if img_ids is None:
            eng.img_ids, eng.img_ids_h = None, [-1, -2]
            eng._graphs.clear()
        elif getattr(eng, "img_ids_h", None) != list(img_ids) or eng.eos_id != eos_token_id:
            eng.set_image_token_ids(img_ids, eos_token_id)
        eng.set_suppress_ids(suppress)
        return schedule

    @torch.no_grad()
    def generate_batch(self, input_ids_list, inputs_embeds_list, logits_processor=None, max_new_tokens=120,
                       eos_token_id=2, past_lens=None, heads=None):
        """Several independent greedy generations sharing every decode step (LlamaEngine.generate_batch; BASELINE
        configs[3]: stories batched per rank over the paged KV cache).  Sequence b lives in engine slot b; per sequence
        the result is what generate() returns for it alone.  Returns a list of GenerateOutput."""
        B = len(input_ids_list)
        Lmax = max(int(i.shape[1]) for i in input_ids_list if i is not None)
        eng = self.engine(max_new=max(512, max_new_tokens + 2), max_ctx=max(4096, Lmax + max_new_tokens + 2), max_batch=B)
        schedule = self._configure_engine(eng, logits_processor, eos_token_id)
        reqs = []
        for b in range(B):
            if input_ids_list[b] is None:
                reqs.append(None)
                continue
            reqs.append(dict(input_ids=input_ids_list[b][0].tolist(), inputs_embeds=inputs_embeds_list[b][0].to(torch.float16),
                             max_new_tokens=max_new_tokens, schedule=schedule,
                             past_len=None if past_lens is None else past_lens[b],
                             head=0 if heads is None else heads[b]))
        outs = []
        for b, res in enumerate(eng.generate_batch(reqs)):
            if res is None:
                outs.append(None)
                continue
            gen, hidden, chunk_hidden = res
            seq = torch.tensor([reqs[b]["input_ids"] + gen], dtype=torch.long, device=input_ids_list[b].device)
            hs = ((chunk_hidden.unsqueeze(0),),) + tuple((hidden[i:i + 1].unsqueeze(0),) for i in range(hidden.shape[0]))
            outs.append(GenerateOutput(seq, hs, None))
        return outs

    # ---------------------------------------------------------------- HF-style generate (greedy only)
    @torch.no_grad()
    def generate(self, input_ids=None, inputs_embeds=None, output_hidden_states=False, return_dict_in_generate=False,
                 logits_processor=None, past_key_values=None, max_new_tokens=120, do_sample=False, num_beams=1,
                 temperature=None, top_p=None, eos_token_id=2, **kwargs):
        if do_sample or num_beams != 1:
            raise NotImplementedError("seedstory_b200 implements the reference's greedy path (do_sample=False, num_beams=1)")
        assert input_ids.shape[0] == 1, "reference generate is batch-1 (models.py:157)"
        eng = self.engine(max_new=max(512, max_new_tokens + 2),
                          max_ctx=max(4096, input_ids.shape[1] + max_new_tokens + 2))
        schedule = self._configure_engine(eng, logits_processor, eos_token_id)
        ids = input_ids[0].tolist()
        if inputs_embeds is None:
            inputs_embeds = self.get_input_embeddings()(input_ids)
        emb = inputs_embeds[0].to(torch.float16)
        m = self.base_model.model
        L = len(ids)
        past_len, head = None, 0
        if past_key_values is not None:
            # live KV reuse (modeling_llama_xformer.py:804-826): with use_kv_cache_head the tokens from kv_cache_head on
            # are fed on top of the given cache at positions kv_cache_head..L-1; without it only the last token is (:827-831)
            head = int(m.kv_cache_head) if (m.use_kv_cache_head and m.kv_cache_head is not None) else L - 1
            if isinstance(past_key_values, llama_engine.RetainedKV):
                assert past_key_values.e is eng and eng.seq_len_h[0] == past_key_values.n, \
                    "RetainedKV handle is stale (the engine's cache changed since it was taken)"
                past_len = past_key_values.n
            else:
                past_len = eng.load_past(0, past_key_values)
        gen, hidden, chunk_hidden = eng.generate(0, ids, emb, max_new_tokens, schedule=schedule, past_len=past_len,
                                                 head=head, return_chunk_hidden=True)
        seq = torch.tensor([ids + gen], dtype=torch.long, device=input_ids.device)
        m.past_key_values = llama_engine.PagedKVView(eng, 0)
        if m.use_kv_cache_head:
            # LlamaForCausalLM.forward advances kv_cache_head by the tokens of every forward (:780-784)
            fed = (L - head) + max(len(gen) - 1, 0)
            m.kv_cache_head = fed if m.kv_cache_head is None else m.kv_cache_head + fed
        if not return_dict_in_generate:
            return seq
        hs = None
        if output_hidden_states:
            # step 0 carries the post-final-norm hidden rows of the fed prompt chunk (read only in the past_key_values
            # branch of ContinuousLVLM.generate, models.py:186-189); later steps the row of each generated token's
            # input position
            step0 = (chunk_hidden.unsqueeze(0),)
            hs = (step0,) + tuple((hidden[i:i + 1].unsqueeze(0),) for i in range(hidden.shape[0]))
        return GenerateOutput(seq, hs, None)


def get_peft_model(model, peft_config):
    return PeftModelForCausalLM(model, peft_config)


def get_peft_model_with_resize_embedding(model, peft_config=None, model_id=None, vocab_size=None, torch_dtype='bf16'):
    if torch_dtype in ('bf16', 'bfloat16'):
        torch_dtype = torch.bfloat16
    elif torch_dtype in ('fp16', 'float16'):
        torch_dtype = torch.float16
    else:
        torch_dtype = torch.float32
    if isinstance(model, dict) or type(model).__name__ == "DictConfig":
        import hydra
        model = hydra.utils.instantiate(model, torch_dtype=torch_dtype)
    assert (peft_config is None) + (model_id is None) == 1
    if vocab_size is not None:
        print(f'Length of tokenizer and resize embedding: {vocab_size}')
        model.resize_token_embeddings(vocab_size)
    if peft_config is None:
        raise NotImplementedError("PeftModel.from_pretrained(model_id) is not part of the inference path")
    print('peft config: ', vars(peft_config) if hasattr(peft_config, "__dict__") else peft_config)
    peft_model = get_peft_model(model=model, peft_config=peft_config)
    peft_model.print_trainable_parameters()
    return peft_model
