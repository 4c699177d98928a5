"""Device-side solve for KabaschUmeyama3DAligner (l4p/models/aligner.py:177-237), B == 1 per the reference's own
reshape(1, ...) in joint_windowed_estimation (dense_heads.py:457-462)."""
from __future__ import annotations

from typing import Dict

import torch

from .. import lib as _l
from .. import ops as _ops
from ..ops import _dev_init, _stream


def solve_sim3(pred: Dict[str, torch.Tensor], target: Dict[str, torch.Tensor], frame_step: int = 3,
               rel_threshold: float = 0.01, iters: int = 8, min_points: int = 10) -> Dict[str, torch.Tensor]:
    d_s, d_t = pred["depth"], target["depth"]
    bs, _, ov, H, W = d_s.shape
    assert bs == 1, "joint alignment is per clip (batch 1), like the reference"
    _dev_init(d_s)
    dev = d_s.device
    f32 = lambda t: t.contiguous().float()
    ds, dt_ = f32(d_s[0, 0]), f32(d_t[0, 0])
    Ks, Kt = f32(pred["camray_intrinsics"].reshape(4, 4, ov)), f32(target["camray_intrinsics"].reshape(4, 4, ov))
    Ps, Pt = f32(pred["camray"].reshape(16, ov)), f32(target["camray"].reshape(16, ov))
    # threshold = 0.01 * 98th percentile of the predicted overlap depth (aligner.py:187-188)
    thr = (torch.quantile(ds.reshape(-1), 0.98) * rel_threshold).reshape(1).float().contiguous()
    ws = torch.empty(34, device=dev, dtype=torch.float64)
    _l.check(_l.load().l4p_sim3_align(ds.data_ptr(), Ks.data_ptr(), Ps.data_ptr(), ov, dt_.data_ptr(), Kt.data_ptr(),
                                      Pt.data_ptr(), ov, ov, frame_step, H, W, thr.data_ptr(), iters, min_points,
                                      ws.data_ptr(), _stream()), "l4p_sim3_align")
    _ops._count(2 * iters)
    T = ws[17:33].reshape(1, 4, 4).to(d_s.dtype)
    s = ws[33:34].to(d_s.dtype)
    return {"T": T, "s": s, "R": T[:, :3, :3] / s, "t": T[:, :3, 3]}
