"""Builds a complete fake SEED-Story project root for the script acceptance test (test infrastructure).

The reference's inference scripts (src/inference/gen_george.py, vis_george_sink.py) read everything through relative
paths: `configs/*.yaml` (hydra targets), `pretrained/…` (tokenizer, Llama-2, ViT, agent, SDXL, de-tokenizer
checkpoints), `data/json/val.jsonl`, `data/image/george_full/…`, and write `output/val_N/…`.  No checkpoint exists
offline, so this module fabricates every one of them in the upstream on-disk layout:

  * configs: the reference's own `_target_` paths and key names with reduced widths / depths (resolutions and token
    counts are the real ones: 448^2 input, 1024^2 output, 64 image tokens);
  * pretrained/cvlm_llama2_tokenizer: a real `transformers.LlamaTokenizer` (byte-fallback BPE) + the 66 image tokens;
  * pretrained/Llama-2-7b-hf: HF directory (config.json + model.safetensors);
  * pretrained/visual_tokenizer/qwen_vit_G.pt, pretrained/seed_story/george_sft/pytorch_model.bin (agent: peft key
    layout), pretrained/stable-diffusion-xl-base-1.0/{scheduler,vae,unet}, the adapted de-tokenizer .bin.

Random weights never emit `<img>` by themselves, so the agent checkpoint is *crafted*: token embeddings carry a one-hot
"successor" code in a few hidden dimensions that the lm_head rows read, so that greedy decoding emits twelve letters and
then `<img>` (the reference's processor then forces the 64 queries and `</img>`), over and over until max_new_tokens.
This gives the turn loop of both scripts real text and image output to work with.
"""
import json
import os

import torch

BOI_TOKEN = '<img>'
EOI_TOKEN = '</img>'
IMG_TOKEN = '<img_{:05d}>'

DIM = 256          # agent / Llama hidden width
VIT = dict(heads=4, image_size=448, layers=2, mlp_ratio=4.0, output_dim=DIM, patch_size=14, width=64)
LLAMA = dict(hidden_size=DIM, intermediate_size=352, num_attention_heads=2, num_hidden_layers=2, rms_norm_eps=1e-5,
             max_position_embeddings=4096)
XL = dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=64, embedding_dim=DIM, output1_dim=96, output2_dim=160,
          ff_mult=4)
UNET = dict(in_channels=4, out_channels=4, block_out_channels=[64, 128, 256], layers_per_block=2,
            transformer_layers_per_block=[0, 1, 2], attention_head_dim=[1, 2, 4], cross_attention_dim=256,
            addition_time_embed_dim=32, projection_class_embeddings_input_dim=160 + 6 * 32, norm_num_groups=32,
            sample_size=128)
VAE = dict(latent_channels=4, out_channels=3, block_out_channels=[64, 64, 128, 128], layers_per_block=2,
           norm_num_groups=32, scaling_factor=0.13025)


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)


def build_tokenizer(path):
    """LlamaTokenizer as transformers 5.x builds it (tokenizers BPE, byte fallback, Metaspace) with a byte-level
    vocabulary, plus the image tokens added exactly as the reference's tokenizer carries them."""
    from transformers import LlamaTokenizer
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    for b in range(256):
        vocab[f"<0x{b:02X}>"] = len(vocab)
    vocab["▁"] = len(vocab)
    # whitespace runs (▁▁, ▁x4 … ▁x64) as the real Llama vocabulary has them: gen_george.py turns every image tag of a
    # generated run into a space (:196), and 66 single-space tokens per run would blow the prompt past max positions
    merges, piece = [], "▁"
    for _ in range(6):
        vocab[piece + piece] = len(vocab)
        merges.append((piece, piece))
        piece = piece + piece
    tk = LlamaTokenizer(vocab=vocab, merges=merges)
    tk.add_tokens([BOI_TOKEN, EOI_TOKEN] + [IMG_TOKEN.format(i) for i in range(64)], special_tokens=True)
    os.makedirs(path, exist_ok=True)
    tk.save_pretrained(path)
    return len(vocab), len(tk)


def build_project(root, n_stories=1, n_captions=27, seed=0):
    """Creates configs/, pretrained/, data/ under `root`.  Must run with the drop-in `src` package and the shims on
    sys.path (it instantiates the same hydra configs the scripts load, on CPU, to produce the checkpoints)."""
    import hydra
    from omegaconf import OmegaConf
    from PIL import Image
    from safetensors.torch import save_file
    torch.manual_seed(seed)
    cwd = os.getcwd()
    os.makedirs(root, exist_ok=True)
    os.chdir(root)
    try:
        n_text, n_vocab = build_tokenizer("pretrained/cvlm_llama2_tokenizer")
        # ---------------- configs (reference _target_ paths and key names; reduced widths) -------------------
        _write("configs/tokenizer/clm_llama_tokenizer.yaml",
               "_target_: transformers.LlamaTokenizer.from_pretrained\n"
               "pretrained_model_name_or_path: pretrained/cvlm_llama2_tokenizer\n")
        _write("configs/processer/qwen_448_transform.yaml",
               "_target_: src.processer.transforms.get_transform\ntype: clip\nimage_size: 448\nkeep_ratio: False\n")
        _write("configs/visual_tokenizer/qwen_vitg_448.yaml",
               "_target_: src.models.qwen_visual.VisionTransformerWithAttnPool.from_pretrained\n"
               + "".join(f"{k}: {v}\n" for k, v in VIT.items())
               + "pretrained_model_path: pretrained/visual_tokenizer/qwen_vit_G.pt\n")
        _write("configs/clm_models/llama2chat7b_lora.yaml",
               "_target_: src.models_clm.peft_models.get_peft_model_with_resize_embedding\n"
               "model:\n"
               "  _target_: src.models_clm.modeling_llama_xformer.LlamaForCausalLM.from_pretrained\n"
               "  pretrained_model_name_or_path: pretrained/Llama-2-7b-hf\n"
               "  low_cpu_mem_usage: True\n"
               "peft_config:\n"
               "  _target_: peft.LoraConfig\n  _convert_: object\n  r: 16\n  lora_alpha: 32\n"
               "  modules_to_save:\n    - input_layernorm\n    - post_attention_layernorm\n    - norm\n"
               "  target_modules:\n    - q_proj\n    - v_proj\n    - k_proj\n    - o_proj\n    - gate_proj\n"
               "    - down_proj\n    - up_proj\n  task_type: CAUSAL_LM\n  lora_dropout: 0.05\n\n"
               f"vocab_size: {n_vocab}\n")
        _write("configs/clm_models/agent_7b_sft.yaml",
               "_target_: src.models_clm.models.ContinuousLVLM.from_pretrained\n"
               f"input_resampler:\n  _target_: src.models.qwen_visual.Resampler\n  grid_size: 8\n  embed_dim: {DIM}\n"
               f"  num_heads: 2\n  kv_dim: {DIM}\n\n"
               f"output_resampler:\n  _target_: src.models.qwen_visual.Resampler\n  grid_size: 16\n  embed_dim: {DIM}\n"
               f"  num_heads: 2\n  kv_dim: {DIM}\n\n"
               "lm_loss_scale: 1.0\nrec_loss_scale: 1.0\n"
               "pretrained_model_path: pretrained/seed_story/george_sft/pytorch_model.bin\n")
        _write("configs/detokenizer/detokenizer_sdxl_qwen_vit_adapted.yaml",
               "_target_: src.models_ipa.adapter_modules.SDXLAdapter.from_pretrained\n\nresampler:\n"
               "  _target_: src.models_ipa.resampler.ResamplerXLV2\n"
               + "".join(f"  {k}: {v}\n" for k, v in XL.items())
               + "\npretrained_model_path: pretrained/detokenizer/detokenizer_george_adapted/checkpoint-4000/pytorch_model.bin\n")
        _write("configs/discrete_model/discrete_identity.yaml",
               "_target_: src.models.discrete_models.DiscreteModleIdentity\n")

        # ---------------- checkpoints -------------------------------------------------------------------------
        from src.models.qwen_visual import VisionTransformerWithAttnPool
        from src.models_clm.modeling_llama_xformer import LlamaForCausalLM
        vit = VisionTransformerWithAttnPool(**VIT)
        os.makedirs("pretrained/visual_tokenizer", exist_ok=True)
        torch.save(vit.state_dict(), "pretrained/visual_tokenizer/qwen_vit_G.pt")

        hf = dict(LLAMA, vocab_size=n_text, architectures=["LlamaForCausalLM"], model_type="llama")
        os.makedirs("pretrained/Llama-2-7b-hf", exist_ok=True)
        with open("pretrained/Llama-2-7b-hf/config.json", "w") as f:
            json.dump(hf, f)
        llama = LlamaForCausalLM(dict(LLAMA, vocab_size=n_text))
        save_file({k: v.contiguous() for k, v in llama.state_dict().items()},
                  "pretrained/Llama-2-7b-hf/model.safetensors")

        # the agent checkpoint: instantiate through the same configs the scripts use (missing .bin tolerated HERE only)
        os.environ["SEEDSTORY_SYNTHETIC"] = "1"
        try:
            llm = hydra.utils.instantiate(OmegaConf.load("configs/clm_models/llama2chat7b_lora.yaml"), torch_dtype="fp16")
            agent = hydra.utils.instantiate(OmegaConf.load("configs/clm_models/agent_7b_sft.yaml"), llm=llm)
        finally:
            os.environ.pop("SEEDSTORY_SYNTHETIC", None)
        sd = {k: v.float().clone() for k, v in agent.state_dict().items()}
        for k in sd:                                     # LoRA B is zero at init: give the adapters signal
            if "lora_B" in k:
                sd[k] = torch.randn_like(sd[k]) * 0.02
        # craft a successor chain: hidden dims 0..n carry a one-hot "what comes next" code in every token embedding and
        # the lm_head rows read it.  Every token is followed by the first of 12 letters, each letter by the next one and
        # the last letter by <img> (the reference's processor then forces the 64 queries and </img>, after which the
        # chain starts over until max_new_tokens).  The turn text is therefore non-empty, as in real use: gen_george.py
        # slices len('[INST]') characters past </img> when it evicts an image (:236), which needs text there.
        emb_k = [k for k in sd if k.endswith("embed_tokens.weight")][0]
        head_k = [k for k in sd if k.endswith("lm_head.weight")][0]
        boi = n_text          # first added token
        letters = [3 + ord(c) for c in "abcdefghijkl"]          # byte-fallback ids: <unk>, <s>, </s>, then <0x00>..
        nl = len(letters)
        sd[emb_k][:, :nl + 1] = 0.0
        sd[head_k][:, :nl + 1] = 0.0
        sd[emb_k][:, 0] = 8.0                                    # default successor: the first letter
        sd[head_k][letters[0], 0] = 4.0
        for j, t in enumerate(letters):
            sd[emb_k][t, 0] = 0.0
            sd[emb_k][t, j + 1] = 8.0                            # letter j -> letter j + 1 ... last letter -> <img>
            sd[head_k][letters[j + 1] if j + 1 < nl else boi, j + 1] = 4.0
        os.makedirs("pretrained/seed_story/george_sft", exist_ok=True)
        torch.save({k: v.half() for k, v in sd.items()}, "pretrained/seed_story/george_sft/pytorch_model.bin")

        import diffusers
        sdxl = "pretrained/stable-diffusion-xl-base-1.0"
        os.makedirs(f"{sdxl}/scheduler", exist_ok=True)
        with open(f"{sdxl}/scheduler/scheduler_config.json", "w") as f:
            json.dump(dict(_class_name="EulerDiscreteScheduler", num_train_timesteps=1000, beta_start=0.00085,
                           beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=1,
                           prediction_type="epsilon"), f)
        unet_cfg = {k: (tuple(v) if isinstance(v, list) else v) for k, v in UNET.items()}
        unet_cfg["num_attention_heads"] = unet_cfg.pop("attention_head_dim")
        unet = diffusers.UNet2DConditionModel(config=unet_cfg, seed=11)
        os.makedirs(f"{sdxl}/unet", exist_ok=True)
        with open(f"{sdxl}/unet/config.json", "w") as f:
            json.dump(dict(UNET, _class_name="UNet2DConditionModel"), f)
        save_file({k: v.half().contiguous() for k, v in unet.state_dict().items()},
                  f"{sdxl}/unet/diffusion_pytorch_model.safetensors")
        vae = diffusers.AutoencoderKL(config={k: (tuple(v) if isinstance(v, list) else v) for k, v in VAE.items()}, seed=12)
        os.makedirs(f"{sdxl}/vae", exist_ok=True)
        with open(f"{sdxl}/vae/config.json", "w") as f:
            json.dump(dict(VAE, _class_name="AutoencoderKL"), f)
        save_file({k: v.contiguous() for k, v in vae.state_dict().items()},
                  f"{sdxl}/vae/diffusion_pytorch_model.safetensors")

        from src.models_ipa.adapter_modules import SDXLAdapter
        from src.models_ipa.resampler import ResamplerXLV2
        adapter = SDXLAdapter(unet, ResamplerXLV2(**XL))
        det = "pretrained/detokenizer/detokenizer_george_adapted/checkpoint-4000"
        os.makedirs(det, exist_ok=True)
        torch.save({k: v for k, v in adapter.state_dict().items()
                    if k.startswith("resampler.") or k.endswith("to_k.weight") or k.endswith("to_v.weight")},
                   f"{det}/pytorch_model.bin")

        # ---------------- data -------------------------------------------------------------------------------
        os.makedirs("data/json", exist_ok=True)
        os.makedirs("data/image/george_full", exist_ok=True)
        words = ["george", "the", "monkey", "sees", "a", "yellow", "hat", "and", "runs", "home"]
        g = torch.Generator().manual_seed(seed + 1)
        with open("data/json/val.jsonl", "w") as f:
            for s in range(n_stories):
                caps = [" ".join(words[int(i)] for i in torch.randint(0, len(words), (6,), generator=g))
                        for _ in range(n_captions)]
                name = f"story{s}/000.jpg"
                os.makedirs(os.path.dirname(os.path.join("data/image/george_full", name)), exist_ok=True)
                px = torch.randint(0, 256, (300, 400, 3), generator=g, dtype=torch.uint8).numpy()
                Image.fromarray(px).save(os.path.join("data/image/george_full", name))
                f.write(json.dumps(dict(images=[name], captions=caps)) + "\n")
    finally:
        os.chdir(cwd)
    return dict(n_text=n_text, n_vocab=n_vocab)
