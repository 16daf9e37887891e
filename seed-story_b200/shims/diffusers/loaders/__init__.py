class LoraLoaderMixin:  # name imported at adapter_modules.py:20; unused on the inference path
    pass
