"""Minimal `peft` stand-in exporting the names the reference configs / imports touch
(configs/clm_models/llama2chat7b_lora.yaml:8 `_target_: peft.LoraConfig`; src/models_clm/peft_models.py:1-13).
The working implementation lives in src/models_clm/peft_models.py of the drop-in."""
from src.models_clm.peft_models import (LoraConfig, PeftModelForCausalLM, get_peft_model)  # noqa: F401

PeftModel = PeftModelForCausalLM
LoraModel = PeftModelForCausalLM
