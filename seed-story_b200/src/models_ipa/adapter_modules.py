"""Drop-in for the one adapter class the shipped config instantiates: SDXLAdapter
(reference src/models_ipa/adapter_modules.py:281-468; hydra target configs/detokenizer/detokenizer_sdxl_qwen_vit_adapted.yaml:1).

Same constructor / from_pretrained / init_pipe / get_image_embeds / generate signatures and return values.
Two input-independent pieces of work the reference redoes on every call are computed once and cached here
(results are identical because their inputs are constants):
  * the unconditional branch — ViT on an all-zeros 448x448 image (reference :406-414) followed by the resampler.
The training forward (reference :327-343) is out of scope.
"""
import os

import torch
from torch import nn


class SDXLAdapter(nn.Module):
    def __init__(self, unet, resampler, full_ft=False) -> None:
        super().__init__()
        self.unet = unet
        self.resampler = resampler
        self.full_ft = full_ft
        self._neg_embeds = None

    def forward(self, *a, **k):
        raise NotImplementedError("training forward (reference adapter_modules.py:327-343) is outside the hot path")

    def encode_image_embeds(self, image_embeds):
        return self.resampler(image_embeds)

    @classmethod
    def from_pretrained(cls, unet, resampler, pretrained_model_path=None, **kwargs):
        model = cls(unet=unet, resampler=resampler, **kwargs)
        if pretrained_model_path is not None:
            if os.path.exists(pretrained_model_path):
                ckpt = torch.load(pretrained_model_path, map_location='cpu')
                missing, unexpected = model.load_state_dict(ckpt, strict=False)
                print('missing keys: ', len(missing), 'unexpected keys:', len(unexpected))
            else:
                from seedstory.synthetic import missing_checkpoint
                missing_checkpoint(pretrained_model_path, "SDXLAdapter (de-tokenizer)")
        return model

    def init_pipe(self, vae, scheduler, visual_encoder, image_transform, discrete_model=None, dtype=torch.float16,
                  device='cuda'):
        from diffusers import StableDiffusionXLPipeline
        self.device = device
        self.dtype = dtype
        self.sdxl_pipe = StableDiffusionXLPipeline(tokenizer=None, tokenizer_2=None, text_encoder=None,
                                                   text_encoder_2=None, vae=vae, unet=self.unet, scheduler=scheduler)
        self.visual_encoder = visual_encoder.to(self.device, dtype=self.dtype)
        self.discrete_model = discrete_model.to(self.device, dtype=self.dtype) if discrete_model is not None else None
        self.image_transform = image_transform
        self._neg_embeds = None

    @torch.inference_mode()
    def get_image_embeds(self, image_pil=None, image_tensor=None, image_embeds=None, return_negative=True,
                         image_size=448):
        assert int(image_pil is not None) + int(image_tensor is not None) + int(image_embeds is not None) == 1
        if image_pil is not None:
            image_tensor = self.image_transform(image_pil).unsqueeze(0).to(self.device, dtype=self.dtype)
        if image_tensor is not None:
            if return_negative:
                image_tensor = torch.cat([image_tensor, torch.zeros_like(image_tensor)], dim=0)
            image_embeds = self.visual_encoder(image_tensor)
        elif return_negative:
            key = (image_size, tuple(image_embeds.shape[1:]))
            if self._neg_embeds is None or self._neg_embeds[0] != key:
                zeros = torch.zeros(1, 3, image_size, image_size).to(image_embeds.device, dtype=image_embeds.dtype)
                self._neg_embeds = (key, self.visual_encoder(zeros))   # constant: computed once (reference :406-413)
            image_embeds = torch.cat([image_embeds, self._neg_embeds[1]], dim=0)
        if self.discrete_model is not None:
            image_embeds = self.discrete_model.encode_image_embeds(image_embeds)
        image_embeds, pooled_image_embeds = self.encode_image_embeds(image_embeds)
        if return_negative:
            image_embeds, image_embeds_neg = image_embeds.chunk(2)
            pooled_image_embeds, pooled_image_embeds_neg = pooled_image_embeds.chunk(2)
        else:
            image_embeds_neg = None
            pooled_image_embeds_neg = None
        return image_embeds, image_embeds_neg, pooled_image_embeds, pooled_image_embeds_neg

    def generate(self, image_pil=None, image_tensor=None, image_embeds=None, seed=42, height=1024, width=1024,
                 guidance_scale=7.5, num_inference_steps=30, input_image_size=448, **kwargs):
        prompt, neg, pooled, neg_pooled = self.get_image_embeds(image_pil=image_pil, image_tensor=image_tensor,
                                                                image_embeds=image_embeds, return_negative=True,
                                                                image_size=input_image_size)
        generator = torch.Generator(self.device).manual_seed(seed) if seed is not None else None
        return self.sdxl_pipe(prompt_embeds=prompt, negative_prompt_embeds=neg, pooled_prompt_embeds=pooled,
                              negative_pooled_prompt_embeds=neg_pooled, guidance_scale=guidance_scale,
                              num_inference_steps=num_inference_steps, generator=generator, height=height,
                              width=width, **kwargs).images
