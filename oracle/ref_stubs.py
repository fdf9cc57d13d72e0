"""Dependency shims that let the UNMODIFIED reference (`/root/reference/hulc`) import in this container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (run in the build container, where
`/root/reference` exists) to produce the committed fixtures under `tests/golden/`.  Nothing on the
GPU box imports this file: `/root/reference` does not exist there.

The reference needs four third-party packages that are not installed (and cannot be, there is no network):
`omegaconf`, `hydra`, `pytorch_lightning`, `calvin_agent`.  Only the handful of symbols the
training hot path touches are provided; they carry no arithmetic.
"""
from __future__ import annotations

import importlib
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


class DictConfig(dict):
    """Attribute-style dict standing in for `omegaconf.DictConfig`."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class ListConfig(list):
    pass


def _instantiate(cfg, *args, **kwargs):
    """`hydra.utils.instantiate` for `_recursive_: false` configs: import `_target_`, call it."""
    if cfg is None:
        return None
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_", None)
    mod, _, name = target.rpartition(".")
    fn = getattr(importlib.import_module(mod), name)
    cfg.update(kwargs)
    return fn(*args, **cfg)


class _LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.logged = {}

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, name, value, **kw):
        self.logged[name] = value.detach().clone() if torch.is_tensor(value) else value

    def save_hyperparameters(self, *a, **k):
        pass


def install():
    """Insert the shims into `sys.modules` and put the reference on `sys.path`."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    om = types.ModuleType("omegaconf")
    om.DictConfig = DictConfig
    om.ListConfig = ListConfig

    class OmegaConf:
        @staticmethod
        def load(path):
            raise FileNotFoundError(path)

    om.OmegaConf = OmegaConf
    sys.modules["omegaconf"] = om

    hy = types.ModuleType("hydra")
    hu = types.ModuleType("hydra.utils")
    hu.instantiate = _instantiate
    hy.utils = hu
    sys.modules["hydra"] = hy
    sys.modules["hydra.utils"] = hu

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = _LightningModule
    pl.Trainer = type("Trainer", (), {})
    plu = types.ModuleType("pytorch_lightning.utilities")
    plu.rank_zero_info = lambda *a, **k: None
    plu.rank_zero_only = lambda f: f
    pl.utilities = plu
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = plu

    ca = types.ModuleType("calvin_agent")
    cam = types.ModuleType("calvin_agent.models")
    cab = types.ModuleType("calvin_agent.models.calvin_base_model")
    cab.CalvinBaseModel = type("CalvinBaseModel", (), {})
    sys.modules["calvin_agent"] = ca
    sys.modules["calvin_agent.models"] = cam
    sys.modules["calvin_agent.models.calvin_base_model"] = cab

    # `transformers.get_constant_schedule` exists in the image; nothing to shim.
    return DictConfig
