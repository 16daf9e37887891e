class LoRALinearLayer:  # name imported at adapter_modules.py:21; unused on the inference path
    pass
