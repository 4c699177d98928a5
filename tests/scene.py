"""Test infrastructure: a geometrically consistent synthetic scene for the joint depth + pose window-alignment chain
(`joint_windowed_estimation`, reference dense_heads.py:360-492 + aligner.py:158-265).

One scene (per-frame depth maps, camera poses, fixed intrinsics) is cut into overlapping windows; window w > 0 is expressed in
its own similarity frame x = s_w R_w x' + t_w (depth / s_w, pose -> [R_w^T R | R_w^T (c - t_w) / s_w]), which is exactly the
ambiguity the chain has to remove: aligned to window 0 the stitched result must reproduce the scene. `StubHead` plays the
depth / camera-ray heads and returns the prepared per-window outputs (the same trick tests/test_live_reference.py uses to run
the reference's own chain), including the batched / sharded call patterns of the drop-in.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


def _rot(gen: torch.Generator, amount: float) -> torch.Tensor:
    q, _ = torch.linalg.qr(torch.eye(3) + amount * torch.randn(3, 3, generator=gen))
    if torch.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def make_scene(T: int, starts: Sequence[int], Tw: int, H: int, W: int, seed: int = 3, focal: Optional[float] = None,
               outlier_frac: float = 0.0, noise: float = 0.0) -> Dict[str, object]:
    gen = torch.Generator().manual_seed(seed)
    focal = float(min(H, W)) if focal is None else focal
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = focal
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    intr = K[None, :, :, None].repeat(1, 1, 1, T)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = torch.stack([2.0 + 0.3 * torch.sin(xx / (W / 6.0) + 0.4 * t) + 0.2 * torch.cos(yy / (H / 8.0) - 0.3 * t)
                         for t in range(T)])[None, None]
    pose = torch.zeros(1, 4, 4, T)
    for t in range(T):
        pose[0, :3, :3, t] = _rot(gen, 0.1)
        pose[0, :3, 3, t] = torch.randn(3, generator=gen) * 0.2
        pose[0, 3, 3, t] = 1.0
    sims = []
    d_w: List[torch.Tensor] = []
    p_w: List[torch.Tensor] = []
    for w, s0 in enumerate(starts):
        if w == 0:
            s, R, t = 1.0, torch.eye(3), torch.zeros(3)
        else:
            s = float(0.6 + 1.2 * torch.rand(1, generator=gen))
            R, t = _rot(gen, 0.3), torch.randn(3, generator=gen) * 0.3
        sims.append((s, R, t))
        sl = slice(s0, s0 + Tw)
        pw = pose[..., sl].clone()
        pw[0, :3, :3] = torch.einsum("ji,jkt->ikt", R, pose[0, :3, :3, sl])
        pw[0, :3, 3] = torch.einsum("ji,jt->it", R, pose[0, :3, 3, sl] - t[:, None]) / s
        dw = depth[:, :, sl] / s
        if noise > 0:
            dw = dw * (1.0 + noise * torch.randn(dw.shape, generator=gen))
        if outlier_frac > 0 and w > 0:     # a "moving object": a block of pixels whose depth disagrees between the windows
            hh, ww = int(H * outlier_frac ** 0.5), int(W * outlier_frac ** 0.5)
            dw = dw.clone()
            dw[..., :hh, :ww] *= 1.5
        d_w.append(dw.contiguous())
        p_w.append(pw.reshape(1, 16, Tw).contiguous())
    return dict(T=T, Tw=Tw, starts=list(starts), intr=intr, depth=depth, pose=pose, depth_w=d_w, pose_w=p_w, sims=sims)


class StubHead(torch.nn.Module):
    """Returns prepared per-window outputs. Window identity travels in the "features": feats[0] is a tensor whose entries are
    the window ids of the rows of the call (one id for the per-window call, several for a batched / sharded call)."""

    def __init__(self, task_name: str, task_suffix: str, outputs: Sequence[torch.Tensor]):
        super().__init__()
        self.task_name, self.task_suffix, self.outputs = task_name, task_suffix, list(outputs)
        self.first_window_intrinsics_b44t = None

    def _ids(self, feats) -> List[int]:
        return [int(v) for v in feats[0].reshape(-1).tolist()]

    def _cat(self, feats) -> torch.Tensor:
        return torch.cat([self.outputs[w] for w in self._ids(feats)], dim=0)

    def forward(self, feats, img_info=None, intrinsics_b44t=None, win_id=None, **kw):
        out = self.outputs[win_id] if win_id is not None else self._cat(feats)
        return {f"{self.task_name}_est_{self.task_suffix}": out.clone()}

    def rays(self, feats, img_info=None):      # the "rays" of the stub camera head are the prepared poses themselves
        return self._cat(feats)

    def pose_from_rays(self, rays, img_info, intrinsics_b44t=None, win_id=None, **kw):
        return {f"{self.task_name}_est_{self.task_suffix}": rays.clone()}


def stub_heads(scene, device) -> torch.nn.ModuleDict:
    return torch.nn.ModuleDict(dict(
        depth=StubHead("depth", "b1thw", [d.to(device) for d in scene["depth_w"]]),
        camray=StubHead("traj3d", "b16t", [p.to(device) for p in scene["pose_w"]])))


def window_feats(ids: Sequence[int], device) -> List[List[torch.Tensor]]:
    """Per-window "feature lists" carrying the window id."""
    return [[torch.tensor([float(i)], device=device)] for i in ids]


def oracle_chain(scene, O, seed: int = 0):
    """The reference chain on the CPU oracle: per window > 0, point maps of every 3rd overlap frame -> similarity
    (RANSAC restatement, aligner.py:177-237) -> scale depth / transform poses (aligner.py:239-265) -> overwrite the buffer."""
    import numpy as np

    np.random.seed(seed)
    T, Tw, starts, intr = scene["T"], scene["Tw"], scene["starts"], scene["intr"]
    depth = torch.zeros(1, 1, T, *scene["depth"].shape[-2:])
    pose = torch.zeros(1, 16, T)
    for w, s in enumerate(starts):
        d, p = scene["depth_w"][w].clone(), scene["pose_w"][w].clone()
        if w > 0:
            ov = starts[w - 1] + Tw - s
            Kov = intr[..., s:s + ov]
            src = O.generate_point_map(d[:, :, 0:ov:3], Kov[..., ::3], p.reshape(1, 4, 4, Tw)[..., 0:ov:3])
            dst = O.generate_point_map(depth[:, :, s:s + ov:3], Kov[..., ::3], pose.reshape(1, 4, 4, T)[..., s:s + ov:3])
            src = src[0].reshape(3, -1).T.double().numpy()
            dst = dst[0].reshape(3, -1).T.double().numpy()
            n = src.shape[0]
            idx = np.random.choice(n, int(n * 0.1), replace=False)    # point_sample_ratio = 0.1 (aligner.py:171-175)
            thr = float(torch.quantile(d[:, :, :ov].reshape(-1), 0.98)) * 0.01
            Tm, _ = O.similarity_ransac(src[idx], dst[idx], thr, min_samples=10)
            d, p = O.sim3_apply(Tm, d, p)
        depth[:, :, s:s + Tw] = d
        pose[:, :, s:s + Tw] = p
    return depth, pose
