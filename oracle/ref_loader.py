"""TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference from /root/reference (build container only;
the path does not exist on the GPU box) behind `sys.modules` shims for the third-party packages that are not
installed here: timm.models.layers/registry, skimage.measure/transform, kornia.morphology, lightning.

Used by tests/golden/make_golden.py to generate fixtures and by the CPU tests that pin `oracle/l4p_oracle.py`
against the real reference.
"""
from __future__ import annotations

import collections.abc
import os
import sys
import types
from itertools import repeat

REFERENCE_ROOT = os.environ.get("L4P_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "l4p"))


def _install_shims() -> None:
    import torch

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        registry = types.ModuleType("timm.models.registry")

        def to_2tuple(x):
            if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
                return tuple(x)
            return tuple(repeat(x, 2))

        def drop_path(x, drop_prob: float = 0.0, training: bool = False):
            assert not (training and drop_prob > 0.0), "shim: eval-only"
            return x

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        layers.to_2tuple, layers.drop_path, layers.trunc_normal_ = to_2tuple, drop_path, trunc_normal_
        registry.register_model = lambda f: f
        timm.models, models.layers, models.registry = models, layers, registry
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                            "timm.models.registry": registry})
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        measure = types.ModuleType("skimage.measure")
        transform = types.ModuleType("skimage.transform")

        def _absent(*a, **k):
            raise RuntimeError("scikit-image is not installed; the multi-window joint aligner cannot run "
                               "through the reference here (SURVEY.md §8c)")

        measure.ransac = _absent
        transform.SimilarityTransform = _absent
        sk.measure, sk.transform = measure, transform
        sys.modules.update({"skimage": sk, "skimage.measure": measure, "skimage.transform": transform})
    if "kornia" not in sys.modules:   # l4p/data/l4p_dataset_mini.py:13 (used for segmentation masks only, not the rgb key)
        kornia = types.ModuleType("kornia")
        morphology = types.ModuleType("kornia.morphology")

        def erosion(*a, **k):
            raise RuntimeError("kornia is not installed; mask erosion cannot run through the reference here")

        morphology.erosion = erosion
        kornia.morphology = morphology
        sys.modules.update({"kornia": kornia, "kornia.morphology": morphology})
    if "lightning" not in sys.modules:
        L = types.ModuleType("lightning")

        class LightningModule(torch.nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

        L.LightningModule = LightningModule
        sys.modules["lightning"] = L


def load():
    """Returns the reference's `l4p` package (unmodified source, shimmed dependencies)."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import l4p  # noqa: F401

    return sys.modules["l4p"]


def instantiate(cfg):
    """~20-line replacement for jsonargparse's class_path/init_args instantiation (l4p/models/utils.py:47-49)."""
    import importlib

    if isinstance(cfg, dict) and "class_path" in cfg:
        mod, _, name = cfg["class_path"].rpartition(".")
        cls = getattr(importlib.import_module(mod), name)
        kwargs = {k: instantiate(v) for k, v in (cfg.get("init_args") or {}).items()}
        return cls(**kwargs)
    if isinstance(cfg, dict):
        return {k: instantiate(v) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [instantiate(v) for v in cfg]
    return cfg
