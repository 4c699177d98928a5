"""Config-driven construction: the reference instantiates its module tree from `configs/model.yaml` with
jsonargparse (`class_path` / `init_args`, l4p/models/utils.py:37-49; LightningCLI, l4p/main.py:10-11).
jsonargparse is not a dependency here: this is the same recursive instantiation in ~30 lines. Class paths
that point at the reference package (`l4p.…`) are transparently mapped to the drop-in classes (`l4p_b200.…`).
"""
from __future__ import annotations

import importlib
from pathlib import Path
from typing import Any, Dict, Optional, Union

import torch
import yaml

_REMAP = {
    "l4p.l4p.": "l4p_b200.l4p.",
    "l4p.models.": "l4p_b200.models.",
}

DEFAULT_CONFIG = Path(__file__).resolve().parents[1] / "configs" / "model.yaml"


def _resolve(class_path: str):
    for old, new in _REMAP.items():
        if class_path.startswith(old):
            class_path = new + class_path[len(old):]
            break
    mod, _, name = class_path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(cfg: Any, device=None) -> Any:
    if isinstance(cfg, dict) and "class_path" in cfg:
        cls = _resolve(cfg["class_path"])
        kwargs = {k: instantiate(v, device) for k, v in (cfg.get("init_args") or {}).items()}
        if device is not None and cls.__module__.startswith("l4p_b200.models") and "device" in cls.__init__.__code__.co_varnames:
            kwargs.setdefault("device", device)
        return cls(**kwargs)
    if isinstance(cfg, dict):
        return {k: instantiate(v, device) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [instantiate(v, device) for v in cfg]
    return cfg


def load_model(config: Union[str, Path, Dict, None] = None, ckpt_path: Optional[str] = None, device=None,
               max_queries: Optional[int] = None, compute_dtype: Optional[torch.dtype] = None):
    """Equivalent of the reference's `prepare_model` (l4p/models/utils.py:15-60) without Fabric: build the module
    tree from the yaml, optionally override `max_queries` (:41-44), load `ckpt["state_dict"]` strictly (:52-54)."""
    if config is None:
        config = DEFAULT_CONFIG
    cfg = yaml.safe_load(open(config)) if not isinstance(config, dict) else config
    if max_queries is not None:
        cfg["init_args"]["l4p_model"]["init_args"]["task_heads"]["init_args"]["modules"]["track_2d"]["init_args"][
            "max_queries"] = max_queries
    model = instantiate(cfg, device)
    if compute_dtype is not None:
        model.l4p_model.set_compute_dtype(compute_dtype)
    if ckpt_path is not None:
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=True)
        model.load_state_dict(ckpt["state_dict"], strict=getattr(model, "strict_loading", True))
    return model
