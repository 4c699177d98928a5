"""Device-side mirror of the reference's l4p/utils/geometry_utils.py functions that sit on the hot path
(SURVEY.md §2 row 9). Same names / argument meaning; the solves run in one CUDA kernel per call
(csrc/geometry.cu) instead of the reference's lstsq + per-frame Python SVD loop + cv2 host round trip."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import lib as _l
from .. import ops as _ops
from ..ops import _dev_init, _stream


def normalize_intrinsics(intrinsics_b44t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """geometry_utils.py:110-116."""
    k = torch.clone(intrinsics_b44t).detach()
    k[:, :2, 2] += 0.5
    k[:, 0] = k[:, 0] / w
    k[:, 1] = k[:, 1] / h
    return k


def denormalize_intrinsics(intrinsics_b44t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """geometry_utils.py:119-125."""
    k = torch.clone(intrinsics_b44t).detach()
    k[:, 0] *= w
    k[:, 1] *= h
    k[:, :2, 2] -= 0.5
    return k


def _pose_call(camray_b6thw: torch.Tensor, k_norm: Optional[torch.Tensor], mode: int, output_size, thr: float,
               refits: int = 3):
    rays = camray_b6thw.contiguous().float()
    _dev_init(rays)
    B, six, T, h, w = rays.shape
    assert six == 6
    dev = rays.device
    ext = torch.empty(B, 4, 4, T, device=dev, dtype=torch.float32)
    pose = torch.empty_like(ext)
    centers = torch.empty(B, T, 3, device=dev, dtype=torch.float32)
    kest = torch.empty_like(ext) if mode == 1 else None
    kgrid = torch.empty(B * 9, device=dev, dtype=torch.float64) if mode == 1 else None
    kn = k_norm.contiguous().float() if k_norm is not None else None
    H, W = output_size
    _l.check(_l.load().l4p_pose_from_rays(
        rays.data_ptr(), None if kn is None else kn.data_ptr(), mode, B, T, h, w, H, W, float(thr), refits,
        None if kgrid is None else kgrid.data_ptr(), ext.data_ptr(), pose.data_ptr(), centers.data_ptr(),
        None if kest is None else kest.data_ptr(), _stream()), "l4p_pose_from_rays")
    _ops._count(2 if mode == 1 else 1)
    return ext, pose, centers, kest


def solve_cameras(camray_b6thw: torch.Tensor, intrinsics_norm_b44t: Optional[torch.Tensor] = None,
                  output_size: Tuple[int, int] = (16, 16), reproj_threshold: float = 0.2):
    """One call -> (extrinsics_b44t, pose_b44t = extrinsics^-1, centers_bt3, intrinsics_est_b44t | None).
    With intrinsics: rays_to_cameras; without: fixed intrinsics estimated from frame 0."""
    if intrinsics_norm_b44t is not None:
        return _pose_call(camray_b6thw, intrinsics_norm_b44t, 0, (1, 1), 0.0)
    return _pose_call(camray_b6thw, None, 1, output_size, reproj_threshold)


def rays_to_cameras(camray_b6thw: torch.Tensor, intrinsics_b44t: torch.Tensor, ctr_only: bool = False):
    """geometry_utils.py:331-406: (extrinsics_b44t | None, camera_centers_bt3) given NORMALISED intrinsics."""
    ext, _, centers, _ = _pose_call(camray_b6thw, intrinsics_b44t, 0, (1, 1), 0.0)
    return (None, centers) if ctr_only else (ext, centers)


def rays_to_cameras_and_intrinsics(camray_b6thw: torch.Tensor, ctr_only: bool = False, reproj_threshold: float = 0.2,
                                   output_size: Tuple[int, int] = (16, 16), fixed_intrinsics: bool = False):
    """geometry_utils.py:459-490 (fixed_intrinsics=True branch, :493-579)."""
    if not fixed_intrinsics:
        raise NotImplementedError("variable per-frame intrinsics are not reachable from configs/model.yaml:45")
    ext, _, centers, kest = _pose_call(camray_b6thw, None, 1, output_size, reproj_threshold)
    return (None, centers) if ctr_only else (ext, centers, kest)
