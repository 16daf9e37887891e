class DownBlock2D:  # name imported at adapter_modules.py:22; unused on the inference path
    pass
