"""Window-overlap aligners on the device. Mirrors l4p/models/aligner.py (LstSqAffineAligner :29-66,
LinearAligner :69-118, KabaschUmeyama3DAligner :158-265): same class names and solve/apply protocol."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch

from .. import lib as _l
from .. import ops as _ops
from ..ops import _dev_init, _stream


class WindowOverlapAligner(ABC):
    @abstractmethod
    def solve(self, pred, target, intrinsics, img_info):
        pass

    @abstractmethod
    def apply(self, pred):
        pass


class LstSqAffineAligner(WindowOverlapAligner):
    """Scale + shift between two depth windows on their overlap; 5-moment reduction + closed-form solve on
    the device (csrc/geometry.cu) instead of torch.linalg.lstsq on a [401408, 2] matrix."""

    def __init__(self, pre_post_fn: Optional[str] = "identity") -> None:
        if pre_post_fn in ("identity", None):
            self.inverse = 0
        elif pre_post_fn == "inverse":
            self.inverse = 1
        else:
            raise ValueError(f"Unknown pre_post_fn: {pre_post_fn}")
        self.sol = None

    def solve(self, pred, target, intrinsics=None, img_info=None, pred_conf=None, target_conf=None):
        _dev_init(pred)
        bs = pred.shape[0]
        p = pred.reshape(bs, -1).float().contiguous()
        t = target.reshape(bs, -1).float().contiguous()
        n = p.shape[1]
        ws = torch.empty(5 * bs, device=p.device, dtype=torch.float64)
        self.sol = torch.empty(bs, 2, device=p.device, dtype=torch.float32)
        _l.check(_l.load().l4p_affine_align_solve(p.data_ptr(), t.data_ptr(), bs, n, n, n, self.inverse, ws.data_ptr(),
                                                  self.sol.data_ptr(), _stream()), "l4p_affine_align_solve")
        _ops._count(2)

    def apply(self, pred):
        bs = pred.shape[0]
        x = pred.reshape(bs, -1).float().contiguous()
        y = torch.empty_like(x)
        _l.check(_l.load().l4p_affine_align_apply(x.data_ptr(), y.data_ptr(), self.sol.data_ptr(), bs, x.shape[1],
                                                  self.inverse, _stream()), "l4p_affine_align_apply")
        _ops._count(1)
        return y.reshape(pred.shape).to(pred.dtype)


class LinearAligner(WindowOverlapAligner):
    """aligner.py:69-118. Not reachable from configs/model.yaml (depth head defaults to align_type='affine',
    dense_heads.py:158,166): SURVEY.md §2 row 7 marks it out of scope, so it is not built."""

    def __init__(self, *a, **k) -> None:
        raise NotImplementedError("LinearAligner is out of scope (SURVEY.md §2 row 7); use align_type='affine'")

    def solve(self, pred, target, intrinsics, img_info):
        raise NotImplementedError

    def apply(self, pred):
        raise NotImplementedError


class KabaschUmeyama3DAligner(WindowOverlapAligner):
    """Joint depth + pose similarity alignment of overlapping windows (aligner.py:158-265) on the device."""

    def __init__(self, calc_scale: bool = True) -> None:
        self.rel_T_b44 = None
        self.calc_scale = calc_scale
        self.min_samples = 10
        self.reprojection_threshold = 0.01
        self.confidence = 0.99
        self.frame_sample_step = 3
        self.point_sample_ratio = 0.1

    def solve(self, pred, target, img_info):
        from .sim3 import solve_sim3

        self.rel_T_b44 = solve_sim3(pred, target, self.frame_sample_step, self.reprojection_threshold)

    def apply(self, pred):
        assert self.rel_T_b44 is not None, "rel_T_b44 is not set"
        bs, _, T, H, W = pred["depth"].shape
        out = {}
        for name, v in pred.items():
            if name == "camray":
                pose = torch.einsum("bij,bjkt->bikt", self.rel_T_b44["T"], v.reshape(bs, 4, 4, T)).clone()
                pose[:, :3, :3] = pose[:, :3, :3] / self.rel_T_b44["s"].reshape(bs, 1, 1, 1)
                out[name] = pose.reshape(bs, -1, T)
            elif name == "depth":
                out[name] = v * self.rel_T_b44["s"].reshape(bs, 1, 1, 1, 1)
            elif name == "camray_intrinsics_est":
                out[name] = v
            else:
                raise ValueError(f"Unknown task name: {name}")
        return out
