"""Minimal `hydra` stand-in: only `hydra.utils.instantiate(cfg, **overrides)` as the reference uses it
(one YAML = one object, nested `_target_`, `_convert_`, `_recursive_`; gen_george.py:40-71)."""
from . import utils  # noqa: F401
