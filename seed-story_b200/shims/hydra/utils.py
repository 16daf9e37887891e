import importlib

from omegaconf import DictConfig, ListConfig, OmegaConf


def _locate(path):
    parts = path.split(".")
    for i in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:i]))
        except ImportError:
            continue
        for p in parts[i:]:
            obj = getattr(obj, p)
        return obj
    raise ImportError(f"cannot locate {path}")


def instantiate(config, *args, **overrides):
    """Nested dicts carrying `_target_` are instantiated first (hydra's default _recursive_=True); a nested
    config WITHOUT being instantiated is what `_recursive_: false` asks for.  `_convert_: object|all` turns
    DictConfig/ListConfig arguments into plain containers."""
    if config is None:
        return None
    cfg = dict(config)
    cfg.update(overrides)
    target = cfg.pop("_target_")
    convert = cfg.pop("_convert_", "none")
    recursive = cfg.pop("_recursive_", True)
    cfg.pop("_partial_", None)

    def build(v):
        if isinstance(v, dict) and "_target_" in v and recursive:
            # the reference passes the *config* of the Llama model into get_peft_model_with_resize_embedding,
            # which instantiates it itself (peft_models.py:35-36): hydra resolves that through the callee's
            # isinstance(model, DictConfig) check only when the callee asks for it — keep hydra's behaviour:
            return instantiate(v)
        if convert in ("all", "object", "partial"):
            return OmegaConf.to_container(v) if isinstance(v, (dict, list)) else v
        return v
    fn = _locate(target)
    kwargs = {}
    for k, v in cfg.items():
        kwargs[k] = build(v)
    return fn(*args, **kwargs)
