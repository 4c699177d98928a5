"""CUDA-graph replay of the fixed-shape inference step.

One all-heads window is ~520 kernel launches on 5 streams; enqueueing them from Python costs ~30 ms of host time per step on
the GPU box's cores - as much as the GPU needs to execute them (tools/e2e_diag.py, round 2), so the step is host-bound as soon
as anything else competes for the core (8 ranks on one host, a data loader, ...). Everything in the step is shape-static and
free of host synchronisation (no `.item()`, no host round trips: the pose / alignment solves run on the device), so the whole
forward - encoder, the concurrent head streams, the geometry solves - is captured ONCE into a CUDA graph per input signature
and replayed with one launch: the tensor maps of the TMA kernels are kernel parameters and are baked in with the addresses of
the graph's private memory pool, the programmatic-dependent-launch edges are kept by the capture.

`StepGraph` owns static input buffers (new inputs are copied in, host->device included) and the static outputs of the
captured call; outputs are overwritten by the next replay, like any graph-pool tensor.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Hashable, Tuple

import torch

from . import ops


def signature(batch: Dict[str, Any], extra: Hashable = ()) -> Tuple:
    """Hashable description of everything a captured graph depends on besides tensor values."""
    items = []
    for k in sorted(batch):
        v = batch[k]
        if torch.is_tensor(v):
            items.append((k, tuple(v.shape), str(v.dtype)))
        else:
            items.append((k, repr(v)))
    return (tuple(items), extra)


class StepGraph:
    def __init__(self, fn: Callable[[Dict[str, Any]], Dict[str, Any]], example: Dict[str, Any], device: torch.device,
                 warmup: int = 2) -> None:
        self.device = torch.device(device)
        self.static_in: Dict[str, Any] = {}
        for k, v in example.items():
            self.static_in[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device) if torch.is_tensor(v) else v
        self._copy_in(example)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):     # lazy initialisation (weight packing, workspaces, func attributes) happens here
                fn(dict(self.static_in))
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.LAUNCHES
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(dict(self.static_in))
        self.launches = ops.LAUNCHES - l0       # kernels of this library inside one replay

    def _copy_in(self, batch: Dict[str, Any]) -> None:
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static_in[k].copy_(v, non_blocking=True)

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        self._copy_in(batch)
        self.graph.replay()
        ops.LAUNCHES += self.launches
        return self.static_out
