"""Dense per-task heads (flow / depth / dynamic mask / camera rays->pose) on the B200 kernels.

Drop-in mirrors of l4p/models/task_heads/dense_heads.py: same class names, constructor arguments, state-dict
keys, `forward` / `forward_windowed` signatures, output keys and windowed stitching rules
(VideoMAEFlowDPTHead :20-143, VideoMAEDepthDPTHead :146-182, VideoMAEDynMaskDPTHead :185-217,
VideoMAETraj3DDPTHead :257-352, joint_windowed_estimation :360-492).

Differences by design: all windows of a clip go through the DPT kernels as one batch (the reference loops),
the depth `exp` is fused into the last conv epilogue, and the pose / alignment solves run on the device.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Type

import torch
from torch import nn

from ... import ops
from ...utils.misc import apply_fn
from ...utils import geometry_utils as G
from ..aligner import KabaschUmeyama3DAligner, LinearAligner, LstSqAffineAligner, WindowOverlapAligner
from .dpt import PixelwiseTaskWithDPT


def _taps16(feats: Sequence[torch.Tensor], hooks: Sequence[int], dtype: torch.dtype) -> Tuple[List[torch.Tensor], int]:
    """16-bit [B*tokens, C] operands for the hooked features (reuses the encoder's copies when present)."""
    cached = getattr(feats, "taps16", None)
    out = []
    B = None
    for h in hooks:
        f = feats[h]
        if f is None:
            raise ValueError(f"encoder feature {h} was not materialised (VideoMAEEncoder.keep_features)")
        B = f.shape[0]
        if cached is not None and h in cached and cached[h].dtype == dtype:
            out.append(cached[h])
        else:
            t16 = torch.empty(f.shape[0] * f.shape[1], f.shape[2], device=f.device, dtype=dtype)
            ops.cast16(f.contiguous().float(), t16)
            out.append(t16)
    return out, B


def _decode_chunked(fn, batched, chunk: Optional[int]) -> torch.Tensor:
    """`fn(feature_list) -> [rows, ...]` over the rows (windows x clips) of a batched encoder result in chunks of `chunk` rows,
    concatenated. Keeps the DPT activation memory (~1.5-2 GB per window and head) independent of the video length, like the
    reference's one-window-at-a-time decode (dense_heads.py:99-106), while each pass still batches `chunk` windows."""
    n = next(f for f in batched if f is not None).shape[0]
    if not chunk or n <= chunk:
        return fn(batched)
    from ..videomae import slice_features
    return torch.cat([fn(slice_features(batched, i, min(n, i + chunk))) for i in range(0, n, chunk)], dim=0)


class VideoMAEFlowDPTHead(nn.Module):
    """2D optical flow DPT head (dense_heads.py:20-143)."""

    compute_dtype = torch.float16

    def __init__(self, task_name: str, out_nchan: int = 2, depth: int = 40, embed_dim: int = 1408,
                 hooks_idx: Optional[List[int]] = None,
                 actpost_scale_factors: Tuple[Tuple, ...] = ((1, 2, 2), (1, 1, 1), (0, 0, 0), (-1, -1, -1)),
                 fusion_scale_factors: Tuple[Tuple, ...] = ((1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)),
                 output_size: Optional[Tuple[int, int, int]] = None,
                 overlap_aligner_type: Optional[Type[WindowOverlapAligner]] = None, aligner_kwargs: Dict = {},
                 device=None) -> None:
        super().__init__()
        feature_dim = 256
        self.out_nchan = out_nchan
        if hooks_idx is None:
            hooks_idx = [depth * 2 // 5, depth * 3 // 5, depth * 4 // 5, depth]
        self.hooks_idx = list(hooks_idx)
        self.task_name = task_name
        self.overlap_aligner_type = overlap_aligner_type
        self.aligner_kwargs = aligner_kwargs
        self.output_size = output_size
        self.task_suffix = f"b{out_nchan}thw"
        self.task_head = PixelwiseTaskWithDPT(
            num_channels=out_nchan, feature_dim=feature_dim, last_dim=feature_dim // 2, hooks_idx=self.hooks_idx,
            layer_dims=[256, 512, 1024, 1024], dim_tokens=[embed_dim] * 4, is_use_conv3d=True, head_type="regression",
            actpost_scale_factors=actpost_scale_factors, fusion_scale_factors=fusion_scale_factors,
            output_size=self.output_size, device=device)

    # hook for subclasses: function fused into the last conv epilogue
    _exp_out = False

    # set by L4P_VideoMAE.forward when this head's DPT ran as part of a group of identical decoders (dpt.forward_grouped):
    # (id of the feature list it was computed from, output, CUDA event recorded behind it on the group's stream)
    _dpt_pre = None

    def _run_dpt(self, enc_features_bpc_list, img_info) -> torch.Tensor:
        pre = self._dpt_pre
        if pre is not None and pre[0] == id(enc_features_bpc_list) and pre[3] == tuple(img_info):
            if pre[2] is not None:
                cur = torch.cuda.current_stream(pre[1].device)
                cur.wait_event(pre[2])
                if not torch.cuda.is_current_stream_capturing():
                    pre[1].record_stream(cur)
            return pre[1]
        taps, B = _taps16(enc_features_bpc_list, self.hooks_idx, self.compute_dtype)
        return self.task_head.dpt(taps, B, tuple(img_info), exp_out=self._exp_out)

    def forward(self, enc_features_bpc_list, img_info: Tuple[int, int, int] = (16, 224, 224), **kwargs):
        task_out = self._run_dpt(enc_features_bpc_list, img_info)
        return {f"{self.task_name}_est_{self.task_suffix}": task_out[:, : self.out_nchan]}

    def _window_outputs(self, enc_features_bpc_2dlist, img_info, intrinsics_b44t, time_strides, **kwargs):
        """Per-window `forward` outputs. Windows are independent here, so this is the hook where a batched
        encoder result (L4P_VideoMAE passes `_batched_windows`) is consumed in one DPT launch sequence."""
        batched = kwargs.pop("_batched_windows", None)
        shard = kwargs.pop("_window_shard", None)
        chunk = kwargs.pop("_max_windows_per_pass", None)
        window_size = img_info[0] if self.output_size is None else self.output_size[0]
        key = f"{self.task_name}_est_{self.task_suffix}"
        run = lambda f: self.forward(f, img_info=img_info, **kwargs)[key]
        if shard is not None:
            # multi-GPU: this rank holds the features of windows [shard.start, shard.start + shard.count) only
            from ...parallel import gather_window_outputs
            local = []
            if shard.count > 0:
                out = _decode_chunked(run, batched, chunk)
                local = list(out.reshape(shard.count, -1, *out.shape[1:]).unbind(0))
            return gather_window_outputs([local], shard)[0]
        if batched is not None:
            nW = len(enc_features_bpc_2dlist)
            out = _decode_chunked(run, batched, chunk)  # [(B*nW), C, T, H, W], window-major
            return list(out.reshape(nW, -1, *out.shape[1:]).unbind(0))
        outs = []
        for win_id in range(len(enc_features_bpc_2dlist)):
            s = int(time_strides[win_id])
            kw = dict(kwargs, win_id=win_id)
            intr = intrinsics_b44t[..., s:s + window_size] if intrinsics_b44t is not None else None
            outs.append(self.forward(enc_features_bpc_2dlist[win_id], img_info=img_info, intrinsics_b44t=intr, **kw)[key])
        return outs

    def forward_windowed(self, enc_features_bpc_2dlist, img_info: Tuple[int, int, int] = (16, 224, 224),
                         time_strides: Optional[torch.Tensor] = None, intrinsics_b44t: Optional[torch.Tensor] = None,
                         **kwargs) -> Dict[str, torch.Tensor]:
        if time_strides is None:
            kwargs.pop("_batched_windows", None)
            kwargs.pop("_max_windows_per_pass", None)
            return self.forward(enc_features_bpc_2dlist[0], img_info=img_info, intrinsics_b44t=intrinsics_b44t, **kwargs)
        if self.output_size is None:
            window_size, H, W = img_info
        else:
            window_size, H, W = self.output_size
        T = int(time_strides[-1] + window_size)
        outs = self._window_outputs(enc_features_bpc_2dlist, img_info, intrinsics_b44t, time_strides, **kwargs)
        est_bktS = None
        for win_id in range(time_strides.shape[0]):
            s = int(time_strides[win_id])
            out = outs[win_id]
            if est_bktS is None:
                sz = list(out.shape)
                sz[2] = T
                est_bktS = torch.zeros(*sz, dtype=out.dtype, device=out.device)
            if win_id > 0 and self.overlap_aligner_type is not None:
                aligner = self.overlap_aligner_type(**self.aligner_kwargs)
                overlap_sz = int(time_strides[win_id - 1]) + window_size - s
                aligner.solve(out[:, :, :overlap_sz], est_bktS[:, :, s:s + overlap_sz],
                              intrinsics_b44t[..., s:s + overlap_sz] if intrinsics_b44t is not None else None, img_info)
                out = aligner.apply(out)
            if self.task_name == "flow_2d_backward" and win_id > 0:
                est_bktS[:, :, s + 1:s + window_size] = out[:, :, 1:]  # first frame of later windows is not valid
            else:
                est_bktS[:, :, s:s + window_size] = out
        return {f"{self.task_name}_est_{self.task_suffix}": est_bktS}


class VideoMAEDepthDPTHead(VideoMAEFlowDPTHead):
    """Depth DPT head (dense_heads.py:146-182)."""

    def __init__(self, task_name: str, out_nchan: int = 1, depth: int = 40, embed_dim: int = 1408,
                 depth_fn: str = "linear", hooks_idx: Optional[List[int]] = None,
                 align_window_overlap_fn: Optional[str] = None, align_type: str = "affine", device=None) -> None:
        super().__init__(task_name, out_nchan, depth, embed_dim, hooks_idx,
                         overlap_aligner_type=LstSqAffineAligner if align_type == "affine" else LinearAligner,
                         aligner_kwargs=dict(pre_post_fn=align_window_overlap_fn), device=device)
        self.depth_fn = depth_fn
        self._exp_out = depth_fn == "exp"   # fused into the last conv's epilogue (also read by the grouped DPT job)

    def forward(self, enc_features_bpc_list, img_info: Tuple[int, int, int] = (16, 224, 224), **kwargs):
        self._exp_out = self.depth_fn == "exp"  # fused into the conv epilogue
        task_out = self._run_dpt(enc_features_bpc_list, img_info)[:, : self.out_nchan]
        depth = task_out if self._exp_out else apply_fn(task_out, fn_type=self.depth_fn)
        return {f"{self.task_name}_est_{self.task_suffix}": depth}


class VideoMAEDynMaskDPTHead(VideoMAEFlowDPTHead):
    """Dynamic mask DPT head (dense_heads.py:185-217)."""

    def __init__(self, task_name: str, out_nchan: int = 1, depth: int = 40, embed_dim: int = 1408,
                 apply_fn: str = "linear", hooks_idx: Optional[List[int]] = None, device=None) -> None:
        super().__init__(task_name, out_nchan, depth, embed_dim, hooks_idx, overlap_aligner_type=None, device=device)
        self.apply_fn = apply_fn

    def forward(self, enc_features_bpc_list, img_info: Tuple[int, int, int] = (16, 224, 224), **kwargs):
        dyn_mask = self._run_dpt(enc_features_bpc_list, img_info)
        if self.apply_fn != "linear":
            dyn_mask[:, 0] = apply_fn(dyn_mask[:, 0], fn_type=self.apply_fn)
        return {f"{self.task_name}_est_{self.task_suffix}": dyn_mask}


class VideoMAETraj3DDPTHead(VideoMAEFlowDPTHead):
    """Plücker ray-map head + device-side rays->camera solve (dense_heads.py:257-352)."""

    def __init__(self, task_name: str, depth: int = 40, embed_dim: int = 1408, hooks_idx: Optional[List[int]] = None,
                 actpost_scale_factors: Tuple[Tuple, ...] = ((1, 0, 0), (1, 0, 0), (0, 0, 0), (-1, -1, -1)),
                 fusion_scale_factors: Tuple[Tuple, ...] = ((1, 1, 1), (1, 1, 1), (2, 1, 1), (2, 2, 2)),
                 output_size: Optional[Tuple[int, int, int]] = (16, 16, 16), use_intrinsics: bool = True,
                 fixed_intrinsics: bool = False, device=None) -> None:
        super().__init__(task_name, 6, depth, embed_dim, hooks_idx, actpost_scale_factors, fusion_scale_factors,
                         output_size, device=device)
        self.task_suffix = "b16t"
        self.use_intrinsics = use_intrinsics
        self.fixed_intrinsics = fixed_intrinsics
        self.first_window_intrinsics_b44t = None
        self.last_rays_b6thw: Optional[torch.Tensor] = None  # exposed for parity tests (the DPT output)

    def rays(self, enc_features_bpc_list, img_info=(16, 224, 224)) -> torch.Tensor:
        rays = self._run_dpt(enc_features_bpc_list, img_info).to(dtype=torch.float32)
        self.last_rays_b6thw = rays
        return rays

    def pose_from_rays(self, rays_est_b6thw, img_info, intrinsics_b44t=None, **kwargs) -> Dict[str, torch.Tensor]:
        T, H, W = img_info
        intrinsics_est_b44t = None
        if not self.use_intrinsics and self.fixed_intrinsics:
            assert "win_id" in kwargs, "win_id is required when setting fixed intrinsics as True"
            if kwargs["win_id"] == 0:
                self.first_window_intrinsics_b44t = None
        k_norm = G.normalize_intrinsics(intrinsics_b44t, H, W).float() if intrinsics_b44t is not None else None
        if self.use_intrinsics:
            _, pose_est_b44t, _, _ = G.solve_cameras(rays_est_b6thw, k_norm)
        elif self.fixed_intrinsics:
            if self.first_window_intrinsics_b44t is None:
                _, pose_est_b44t, _, intrinsics_est_b44t = G.solve_cameras(rays_est_b6thw, None, (H, W), 0.2)
                self.first_window_intrinsics_b44t = intrinsics_est_b44t.clone()
            else:
                # reference quirk kept: later windows use the INPUT intrinsics for the solve but report the
                # first-window estimate (dense_heads.py:327-334)
                _, pose_est_b44t, _, _ = G.solve_cameras(rays_est_b6thw, k_norm)
                intrinsics_est_b44t = self.first_window_intrinsics_b44t.clone()
        else:
            raise NotImplementedError("variable per-frame intrinsics (fixed_intrinsics=False) is not reachable from "
                                      "configs/model.yaml:45 and is not built (SURVEY.md §2 row 9)")
        B = pose_est_b44t.shape[0]
        out = {f"{self.task_name}_est_{self.task_suffix}": pose_est_b44t.reshape(B, 16, T)}
        if intrinsics_est_b44t is not None:
            out[f"{self.task_name}_intrinsics_est_{self.task_suffix}"] = intrinsics_est_b44t.reshape(B, 16, T)
        return out

    def forward(self, enc_features_bpc_list, img_info: Tuple[int, int, int] = (16, 224, 224),
                intrinsics_b44t: Optional[torch.Tensor] = None, **kwargs) -> Dict[str, torch.Tensor]:
        rays = self.rays(enc_features_bpc_list, img_info)
        self.last_rays_b6thw = rays
        return self.pose_from_rays(rays, img_info, intrinsics_b44t, **kwargs)

    def _window_outputs(self, enc_features_bpc_2dlist, img_info, intrinsics_b44t, time_strides, **kwargs):
        raise NotImplementedError  # pose windows are stateful: handled in forward_windowed below

    def forward_windowed(self, enc_features_bpc_2dlist, img_info=(16, 224, 224), time_strides=None,
                         intrinsics_b44t=None, **kwargs):
        batched = kwargs.pop("_batched_windows", None)
        shard = kwargs.pop("_window_shard", None)
        chunk = kwargs.pop("_max_windows_per_pass", None)
        if time_strides is None:
            return self.forward(enc_features_bpc_2dlist[0], img_info=img_info, intrinsics_b44t=intrinsics_b44t, **kwargs)
        window_size = self.output_size[0] if self.output_size is not None else img_info[0]
        T = int(time_strides[-1] + window_size)
        nW = time_strides.shape[0]
        rays_all = None
        run = lambda f: self.rays(f, img_info)
        if shard is not None:
            from ...parallel import gather_window_outputs
            local = []
            if shard.count > 0:
                r = _decode_chunked(run, batched, chunk)
                local = list(r.reshape(shard.count, -1, *r.shape[1:]).unbind(0))
            rays_all = gather_window_outputs([local], shard)[0]
        elif batched is not None:
            r = _decode_chunked(run, batched, chunk)
            rays_all = r.reshape(nW, -1, *r.shape[1:])
        key = f"{self.task_name}_est_{self.task_suffix}"
        est = None
        for win_id in range(nW):
            s = int(time_strides[win_id])
            rays = rays_all[win_id] if rays_all is not None else self.rays(enc_features_bpc_2dlist[win_id], img_info)
            self.last_rays_b6thw = rays
            out = self.pose_from_rays(rays, img_info, intrinsics_b44t[..., s:s + window_size], win_id=win_id)[key]
            if est is None:
                est = torch.zeros(out.shape[0], 16, T, dtype=out.dtype, device=out.device)
            est[:, :, s:s + window_size] = out
        return {key: est}


########################################################################################
# JOINT DEPTH AND CAMERA ALIGNMENT
########################################################################################


def joint_windowed_estimation(task_names: List[str], task_heads: nn.ModuleDict, enc_features_bpc_2dlist,
                              time_strides: Optional[torch.Tensor] = None, intrinsics_b44t: Optional[torch.Tensor] = None,
                              img_info: Tuple[int, int, int] = (16, 224, 224), **kwargs) -> Dict[str, torch.Tensor]:
    """Joint depth + camera windowed estimation with similarity alignment on the overlap
    (dense_heads.py:360-492). Per-window head outputs are computed as one batch; the alignment chain is
    sequential across windows and runs on the device (KabaschUmeyama3DAligner)."""
    batched = kwargs.pop("_batched_windows", None)
    shard = kwargs.pop("_window_shard", None)
    chunk = kwargs.pop("_max_windows_per_pass", None)
    out_all_tasks: Dict[str, torch.Tensor] = {}
    if time_strides is None:
        for task_name in task_names:
            out_all_tasks.update(task_heads[task_name].forward(enc_features_bpc_2dlist[0], img_info=img_info,
                                                               intrinsics_b44t=intrinsics_b44t, **kwargs))
        return out_all_tasks

    window_size = img_info[0]
    T = int(time_strides[-1] + window_size)
    nW = time_strides.shape[0]
    depth_head, cam_head = task_heads["depth"], task_heads["camray"]
    dkey = f"{depth_head.task_name}_est_{depth_head.task_suffix}"
    ckey = f"{cam_head.task_name}_est_{cam_head.task_suffix}"
    ikey = f"{cam_head.task_name}_intrinsics_est_{cam_head.task_suffix}"

    # independent per-window network outputs, batched over windows when the encoder result is batched
    if shard is not None:
        # multi-GPU (cfg 4): depth maps and ray maps of this rank's windows -> ONE all-gather -> every rank runs the chain
        from ...parallel import gather_window_outputs
        dl, rl = [], []
        if shard.count > 0:
            d = _decode_chunked(lambda f: depth_head.forward(f, img_info=img_info)[dkey], batched, chunk)
            dl = list(d.reshape(shard.count, -1, *d.shape[1:]).unbind(0))
            r = _decode_chunked(lambda f: cam_head.rays(f, img_info), batched, chunk)
            rl = list(r.reshape(shard.count, -1, *r.shape[1:]).unbind(0))
        depth_w, rays_w = gather_window_outputs([dl, rl], shard)
    elif batched is not None:
        d = _decode_chunked(lambda f: depth_head.forward(f, img_info=img_info)[dkey], batched, chunk)
        depth_w = list(d.reshape(nW, -1, *d.shape[1:]).unbind(0))
        r = _decode_chunked(lambda f: cam_head.rays(f, img_info), batched, chunk)
        rays_w = list(r.reshape(nW, -1, *r.shape[1:]).unbind(0))
    else:
        depth_w = [depth_head.forward(enc_features_bpc_2dlist[w], img_info=img_info)[dkey] for w in range(nW)]
        rays_w = [cam_head.rays(enc_features_bpc_2dlist[w], img_info) for w in range(nW)]

    est: Dict[str, Optional[torch.Tensor]] = {"depth": None, "camray": None, "camray_intrinsics_est": None}
    for win_id in range(nW):
        s = int(time_strides[win_id])
        cam_out = cam_head.pose_from_rays(rays_w[win_id], img_info, intrinsics_b44t[..., s:s + window_size],
                                          win_id=win_id)
        cur = {"depth": depth_w[win_id], "camray": cam_out[ckey]}
        if ikey in cam_out:
            cur["camray_intrinsics_est"] = cam_out[ikey]
        else:
            cur["camray_intrinsics_est"] = intrinsics_b44t[..., s:s + window_size].clone().reshape(1, 16, window_size)
        if est["depth"] is None:
            for k, v in cur.items():
                sz = list(v.shape)
                sz[2] = T
                est[k] = torch.zeros(*sz, dtype=v.dtype, device=v.device)
        if win_id > 0:
            aligner = KabaschUmeyama3DAligner()
            ov = int(time_strides[win_id - 1]) + window_size - s
            pred = {"depth": cur["depth"][:, :, :ov], "camray": cur["camray"][:, :, :ov],
                    "camray_intrinsics": cur["camray_intrinsics_est"][:, :, :ov].reshape(1, 4, 4, ov).clone()}
            target = {"depth": est["depth"][:, :, s:s + ov], "camray": est["camray"][:, :, s:s + ov],
                      "camray_intrinsics": est["camray_intrinsics_est"][:, :, s:s + ov].reshape(1, 4, 4, ov)}
            aligner.solve(pred, target, img_info)
            cur = aligner.apply(cur)
        for k in ("depth", "camray", "camray_intrinsics_est"):
            est[k][:, :, s:s + window_size] = cur[k]

    out_all_tasks[dkey] = est["depth"]
    out_all_tasks[ckey] = est["camray"]
    if est["camray_intrinsics_est"] is not None:
        out_all_tasks[ikey] = est["camray_intrinsics_est"]
    return out_all_tasks
