"""Stage-level timing of the full-size pipeline (CUDA events) — development aid; bench.py is the contract benchmark."""
import sys
import time

import torch

sys.path[:0] = ["seed-story_b200", "seed-story_b200/shims"]
from seedstory import _capi, ops, story  # noqa: E402

dev = torch.device("cuda:0")
t0 = time.time()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
from src.models_clm.modeling_llama_xformer import LlamaForCausalLM
from src.models_clm.peft_models import LoraConfig, get_peft_model_with_resize_embedding
with torch.device(dev):
    llama = LlamaForCausalLM(story.FULL["llama"]).to(dtype=torch.float16)
    llm_m = get_peft_model_with_resize_embedding(llama, peft_config=LoraConfig(r=16, lora_alpha=32, target_modules=["q_proj", "v_proj", "k_proj", "o_proj", "gate_proj", "down_proj", "up_proj"], modules_to_save=["input_layernorm", "post_attention_layernorm", "norm"]), vocab_size=32066, torch_dtype="fp16")
print(f"build {time.time()-t0:.1f}s, mem {torch.cuda.memory_allocated()/2**30:.1f} GiB")


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, n=1):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(n):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r


# LLM: prefill 1041 tokens, decode steps, 66-chunk
llm = llm_m
eng = llm.engine(max_new=512)
eng.set_image_token_ids([32000] + list(range(32002, 32066)) + [32001], 2)
L = 1041
x = (torch.randn(L, 4096, device=dev) * 0.02).half()
for rep in range(2):
    eng.reset_sequence(0)
    ms, _ = timed(lambda: eng.forward_chunk(0, x, list(range(L))))
    print(f"prefill {L} tokens: {ms:.2f} ms ({L*12.95e9/ms/1e9:.0f} TFLOP/s)")
eng.begin_decode([5], [L])
eng.decode_step(1)  # capture
for rep in range(2):
    ms, _ = timed(lambda: eng.decode_step(1), n=20)
    print(f"decode step (graph, ctx~{L}): {ms:.3f} ms  -> {13.215e9/ms/1e6:.0f} GB/s weight streaming")
ms, _ = timed(lambda: eng.decode_step(1, use_graph=False), n=5)
print(f"decode step (eager launches): {ms:.3f} ms")
eng.seq_len_h[0] += 27
xc = (torch.randn(66, 4096, device=dev) * 0.02).half()
p0 = eng.seq_len_h[0]
for rep in range(2):
    ms, _ = timed(lambda: eng.forward_chunk(0, xc, list(range(p0, p0 + 66))))
    p0 += 66
    print(f"66-token chunk: {ms:.2f} ms")

