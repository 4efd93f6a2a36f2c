"""Config surface of the drop-in: the reference's Hydra/OmegaConf model YAMLs, read with plain PyYAML.

The reference assembles its model from string-keyed registries driven by ``conf/experiment/model/*.yaml``
(SURVEY.md section 0); Hydra and OmegaConf are not needed to *read* one model group, so this module gives a
``DictConfig``-like view (``Cfg``) over ``yaml.safe_load`` output and applies the one config rewrite the
reference performs before building the model (``*_epoch(s)`` -> ``*_iter(s)``, nlf/__init__.py:306-315).
"""
from __future__ import annotations

import copy
from typing import Any

_EPOCH_PREFIXES = ("max_freq", "wait", "stop", "falloff", "window", "no_bias", "window_bias",
                   "window_bias_start", "decay", "warmup")


class Cfg(dict):
    """Attribute-access dict: ``cfg.k``, ``cfg['k']``, ``'k' in cfg``, assignment -- the DictConfig subset
    the reference's constructors rely on."""

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = v

    def __deepcopy__(self, memo):
        return Cfg({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_cfg(o: Any) -> Any:
    if isinstance(o, dict):
        return Cfg({k: to_cfg(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return [to_cfg(v) for v in o]
    return o


def to_plain(o: Any) -> Any:
    if isinstance(o, dict):
        return {k: to_plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [to_plain(v) for v in o]
    return o


def _yaml_loader():
    """PyYAML's SafeLoader with the float resolver OmegaConf installs (omegaconf/_utils.py:get_yaml_loader): YAML 1.1 reads
    ``1e-4`` (no dot) as a string, Hydra/OmegaConf -- what the reference is launched through -- as a float
    (``rm_weight_mask_thre: 1e-4`` in shiny_z_tensorf_cascaded.yaml is compared with a tensor, tensorf_no_sample.py:201)."""
    import re

    import yaml

    class Loader(yaml.SafeLoader):
        pass

    Loader.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile(r"""^(?:
         [-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
        |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
        |\.[0-9_]+(?:[eE][-+][0-9]+)?
        |[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+\.[0-9_]*
        |[-+]?\.(?:inf|Inf|INF)
        |\.(?:nan|NaN|NAN))$""", re.X),
        list("-+0123456789."))
    return Loader


def load_model_yaml(path: str) -> Cfg:
    """Read one ``conf/experiment/model/<name>.yaml`` of a reference checkout (unforked, read-only)."""
    import yaml

    with open(path) as f:
        return to_cfg(yaml.load(f, Loader=_yaml_loader()))


def epochs_to_iters(cfg: Any, iters_per_epoch: int) -> Any:
    """In-place: for every ``<p>_epoch`` / ``<p>_epochs`` key add ``<p>_iter`` / ``<p>_iters`` =
    value * iters_per_epoch (lists of lists element-wise); does not descend below a rewritten key."""
    if isinstance(cfg, dict):
        for key in list(cfg.keys()):
            base = None
            for p in _EPOCH_PREFIXES:
                if key == p + "_epoch" or key == p + "_epochs":
                    base = p
            if base is None:
                epochs_to_iters(cfg[key], iters_per_epoch)
                continue
            v = cfg[key]
            new_key = key.replace("epoch", "iter")
            if isinstance(v, list):
                cfg[new_key] = [[x * iters_per_epoch for x in row] for row in v]
            else:
                cfg[new_key] = v * iters_per_epoch
    elif isinstance(cfg, list):
        for v in cfg:
            epochs_to_iters(v, iters_per_epoch)
    return cfg
