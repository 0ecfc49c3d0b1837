"""A small Hydra-compatible config composer for the reference's ``configurations/`` tree
(main.py:24-29 uses ``@hydra.main(config_path="configurations", config_name="config")``).

Hydra / OmegaConf are not installed in the build image, so this implements the subset the
reference's command lines use: a root ``defaults`` list of ``group: option`` entries, per-file
``defaults: [base]`` inheritance inside a group, ``group=option`` selection, ``a.b.c=value`` and
``+a.b.c=value`` overrides, with attribute access on the result.  When OmegaConf *is* importable
``compose(..., as_omegaconf=True)`` returns a ``DictConfig`` instead.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Iterable, List, Optional

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configurations")


class Config(dict):
    """dict with attribute access (the part of DictConfig the algorithm class uses)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(x):
        if isinstance(x, dict):
            return Config({k: Config.wrap(v) for k, v in x.items()})
        if isinstance(x, list):
            return [Config.wrap(v) for v in x]
        return x


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _load_group_file(config_dir: str, group: str, option: str) -> Dict[str, Any]:
    path = os.path.join(config_dir, group, option + ".yaml")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no config '{option}' in group '{group}' ({path})")
    with open(path) as f:
        data = _numify(yaml.safe_load(f) or {})
    out: Dict[str, Any] = {}
    for parent in data.pop("defaults", []) or []:
        if isinstance(parent, str):
            _merge(out, _load_group_file(config_dir, group, parent))
    return _merge(out, data)


def _parse_value(text: str):
    return _numify(yaml.safe_load(text))


def _numify(x):
    """PyYAML (YAML 1.1) reads ``1e-5`` as a string; Hydra/OmegaConf read it as a float."""
    if isinstance(x, dict):
        return {k: _numify(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_numify(v) for v in x]
    if isinstance(x, str):
        try:
            return float(x) if any(c in x for c in "eE.") and x.strip() not in ("", ".") else x
        except ValueError:
            return x
    return x


def compose(overrides: Optional[Iterable[str]] = None, config_dir: str = CONFIG_DIR, config_name: str = "config",
            as_omegaconf: bool = False):
    overrides = list(overrides or [])
    with open(os.path.join(config_dir, config_name + ".yaml")) as f:
        root = yaml.safe_load(f) or {}
    groups: Dict[str, str] = {}
    for entry in root.pop("defaults", []) or []:
        if isinstance(entry, dict):
            groups.update({str(k): str(v) for k, v in entry.items()})
    dotted: List[str] = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        if key.lstrip("+") in groups and "." not in key:
            groups[key.lstrip("+")] = val
        else:
            dotted.append(ov)
    cfg: Dict[str, Any] = {}
    for group, option in groups.items():
        cfg[group] = _load_group_file(config_dir, group, option)
    _merge(cfg, root)
    for ov in dotted:
        key, _, val = ov.partition("=")
        add = key.startswith("+")
        parts = key.lstrip("+").split(".")
        node = cfg
        for p in parts[:-1]:
            if p not in node:
                if not add:
                    raise KeyError(f"override '{ov}': '{p}' does not exist (use +{key.lstrip('+')}=... to add)")
                node[p] = {}
            node = node[p]
        if not add and parts[-1] not in node:
            raise KeyError(f"override '{ov}': key does not exist (use +{key.lstrip('+')}=... to add)")
        node[parts[-1]] = _parse_value(val)
    if as_omegaconf:
        from omegaconf import OmegaConf      # raises if absent
        return OmegaConf.create(cfg)
    return Config.wrap(cfg)
