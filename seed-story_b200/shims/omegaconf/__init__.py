"""Minimal `omegaconf` stand-in (the reference uses only OmegaConf.load + DictConfig, e.g. gen_george.py:39,
peft_models.py:35).  PyYAML-backed; used only when the real package is absent."""
import yaml


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ListConfig(list):
    pass


def _wrap(o):
    if isinstance(o, dict):
        return DictConfig({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return ListConfig(_wrap(v) for v in o)
    return o


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def create(obj):
        return _wrap(obj)

    @staticmethod
    def to_container(cfg, resolve=True):
        if isinstance(cfg, dict):
            return {k: OmegaConf.to_container(v) for k, v in cfg.items()}
        if isinstance(cfg, list):
            return [OmegaConf.to_container(v) for v in cfg]
        return cfg
