"""Drop-in mirror of l4p/l4p.py: `L4PLitModule(tasks, l4p_model, ..., strict_loading)` with `forward(batch, tasks)`
(:37-39) and `predict_step(batch, batch_idx)` (:107-109, :54-66). Training hooks (loss / metrics / optimizers,
:68-126) are outside the inference hot path and raise.

If `lightning` is importable the class derives from `lightning.LightningModule` so it drops under the
reference's LightningCLI unchanged; otherwise it is a plain `torch.nn.Module` with the same surface."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

try:  # pragma: no cover - lightning is not installed in the build image
    import lightning as L

    _Base = L.LightningModule
except Exception:  # noqa: BLE001
    _Base = torch.nn.Module


class L4PLitModule(_Base):
    def __init__(self, tasks: List[str], l4p_model: torch.nn.Module, loss_module: Optional[torch.nn.Module] = None,
                 metrics_module: Optional[torch.nn.Module] = None, optimizer_opts: Optional[Dict[str, Any]] = None,
                 scheduler_opts: Optional[Dict[str, Any]] = None, strict_loading: bool = True):
        super().__init__()
        self.tasks = tasks
        self.l4p_model = l4p_model
        self.loss_module = loss_module
        self.metrics_module = metrics_module
        self.optimizer_opts = optimizer_opts
        self.scheduler_opts = scheduler_opts
        self.strict_loading = strict_loading
        # opt-in CUDA-graph replay of predict_step (l4p_b200/graph.py): one captured graph per input signature
        self.use_cuda_graph = False
        self._graphs: Dict[Any, Any] = {}

    def _device(self) -> torch.device:
        dev = getattr(super(), "device", None)
        if isinstance(dev, torch.device):
            return dev
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cuda")

    def forward(self, batch, tasks):
        return self.l4p_model.forward(batch, tasks)

    def enable_cuda_graph(self, enabled: bool = True) -> None:
        """Replay `predict_step` from a CUDA graph (captured on the first call with a given set of input shapes / tasks).
        The returned tensors are the graph's static outputs: consume or copy them before the next call."""
        self.use_cuda_graph = bool(enabled)
        if not enabled:
            self._graphs.clear()

    def step(self, phase, batch, batch_idx):
        dev = self._device()
        if phase == "predict" and self.use_cuda_graph and dev.type == "cuda":
            from .graph import StepGraph, signature

            key = signature(batch, tuple(self.tasks))
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = StepGraph(lambda b: self.forward(b, self.tasks), batch, dev)
            return g(batch)
        for key in batch.keys():
            if torch.is_tensor(batch[key]):
                # non_blocking: a pinned host batch is copied asynchronously on the compute stream (ordering with the kernels
                # that read it is the stream's), so the host keeps enqueueing instead of waiting for the previous step
                batch[key] = batch[key].to(device=dev, non_blocking=True)
        out = self.forward(batch, self.tasks)
        if phase == "predict":
            return out
        raise NotImplementedError("l4p_b200 implements the inference path only (predict); the reference ships no loss "
                                  "module either (configs/model.yaml)")

    def predict_step(self, batch, batch_idx):
        return self.step("predict", batch, batch_idx)

    def training_step(self, batch, batch_idx):
        return self.step("train", batch, batch_idx)

    def validation_step(self, batch, batch_idx):
        return self.step("val", batch, batch_idx)

    def test_step(self, batch, batch_idx):
        return self.step("val", batch, batch_idx)

    def configure_optimizers(self):
        return {"optimizer": None, "lr_scheduler": None}
