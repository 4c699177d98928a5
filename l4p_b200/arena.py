"""Packed weight arena: after a checkpoint is loaded (`l4p/models/utils.py:52-54` semantics, strict
`state_dict` with the reference's keys incl. the aliased duplicates), every kernel-ready operand of the model - 16-bit GEMM /
conv weights in their packed K-major layouts, fp32 biases, LayerNorm parameters, folded positional tables - is moved into ONE
contiguous device allocation, and the fp32 master parameters can be released.

HBM per GPU for the shipped model (1.42 B parameters): 5.7 GB fp32 masters + 2.9 GB per-module 16-bit copies before, one
2.9 GB arena after `pack_weights(model, release_masters=True)`. One allocation also means one `cudaMemcpy` to replicate the
weights (multi-GPU start-up) and a stable address range for the whole step (CUDA-graph friendly).

The modules keep working unchanged: their `_packed` dictionaries now hold views into the arena. With the masters released
the model is inference-frozen: `state_dict()` returns empty tensors and weights can only change by building a new model.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Tuple

import torch

RELEASE_MIN_NUMEL = 8192
ALIGN = 256   # bytes: TMA global addresses need 16, 256 keeps every tensor on its own L2 sector group


@dataclass
class ArenaInfo:
    buffer: torch.Tensor          # uint8 [total_bytes]
    tensors: int
    bytes_16bit: int
    bytes_fp32: int
    masters_released_bytes: int

    @property
    def total_bytes(self) -> int:
        return self.buffer.numel()


def _owners(model: torch.nn.Module) -> List[torch.nn.Module]:
    """Modules that own a kernel-ready `_packed` dictionary (encoder, DPT adapters, track head)."""
    return [m for m in model.modules() if hasattr(m, "_packed") and hasattr(m, "prepare")]


def _prepare_all(model: torch.nn.Module, device: torch.device, dtype: torch.dtype) -> None:
    from .models.videomae import VideoMAEEncoder

    for m in _owners(model):
        if isinstance(m, VideoMAEEncoder):
            m.compute_dtype = dtype
            m.prepare(device)
        else:
            m.prepare(device, dtype)


def _walk(node: Any, fn):
    """Rebuild a nested dict / list / tuple structure with every tensor leaf replaced by fn(tensor)."""
    if torch.is_tensor(node):
        return fn(node)
    if isinstance(node, dict):
        return {k: _walk(v, fn) for k, v in node.items()}
    if isinstance(node, list):
        return [_walk(v, fn) for v in node]
    if isinstance(node, tuple):
        return tuple(_walk(v, fn) for v in node)
    return node


def pack_weights(model: torch.nn.Module, device=None, dtype: torch.dtype = None, release_masters: bool = False) -> ArenaInfo:
    """Pack every kernel operand of `model` (an `L4P_VideoMAE` or `L4PLitModule`) into one device buffer."""
    core = getattr(model, "l4p_model", model)
    if device is None:
        device = next(core.parameters()).device
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:   # the modules compare against tensor.device (always indexed)
        device = torch.device("cuda", torch.cuda.current_device())
    if dtype is None:
        dtype = core.compute_dtype
    _prepare_all(core, device, dtype)
    owners = _owners(core)
    # 1. collect distinct tensors (aliased entries - e.g. the DPT's duplicate `layerN_rn` keys - share storage already)
    seen: Dict[Tuple[int, int], torch.Tensor] = {}

    def collect(t: torch.Tensor):
        if t.device == device:
            seen.setdefault((t.data_ptr(), t.numel() * t.element_size()), t)
        return t

    for m in owners:
        _walk(m._packed, collect)
    offsets: Dict[Tuple[int, int], int] = {}
    total = 0
    b16 = b32 = 0
    for key, t in seen.items():
        offsets[key] = total
        n = t.numel() * t.element_size()
        total += (n + ALIGN - 1) // ALIGN * ALIGN
        if t.element_size() == 2:
            b16 += n
        else:
            b32 += n
    buf = torch.empty(max(total, ALIGN), device=device, dtype=torch.uint8)
    views: Dict[Tuple[int, int], torch.Tensor] = {}
    for key, t in seen.items():
        n = t.numel() * t.element_size()
        v = buf[offsets[key]:offsets[key] + n].view(t.dtype).view(t.shape)
        v.copy_(t.contiguous())
        views[key] = v

    def rebind(t: torch.Tensor):
        return views.get((t.data_ptr(), t.numel() * t.element_size()), t)

    for m in owners:
        m._packed = _walk(m._packed, rebind)
    released = 0
    if release_masters:
        # weight matrices / conv kernels only: the few small embedding tables the host code reads directly (prompt encoder,
        # mask token: < 50 KB in total) stay
        for p in core.parameters():
            if p.numel() >= RELEASE_MIN_NUMEL:
                released += p.numel() * p.element_size()
                p.data = torch.empty(0, device=p.device, dtype=p.dtype)
        core._masters_released = True
    info = ArenaInfo(buf, len(seen), b16, b32, released)
    core._weight_arena = info
    return info
