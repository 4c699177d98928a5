"""TEST INFRASTRUCTURE ONLY — not product code; see oracle/__init__.py.

CPU (torch, fp32) restatement of the reference algorithm of NVlabs/L4P's feed-forward inference hot path,
written as pure functions over a *state dict with the reference's key names*. Every function cites the
reference file:line it follows (paths relative to the reference repo root).

Pinning: `tests/test_oracle_golden.py` checks these functions against (a) golden vectors produced by running
the UNMODIFIED reference (imported through oracle/ref_loader.py) on seeded synthetic weights
(tests/golden/make_golden.py, fixtures under tests/golden/), and (b) live against the imported reference when
/root/reference is present. The reference ships no tests or golden data of its own, and its two
RANSAC steps (cv2.findHomography, skimage.ransac) are randomised / third-party: those two functions are
"parity unpinned" beyond closed-form known-answer tests (see `similarity_ransac`, `homography_intrinsics`).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# =====================================================================================================
# encoder: l4p/models/l4p_videomae.py:80-122, VideoMAEv2/models/modeling_finetune.py, modeling_pretrain.py
# =====================================================================================================
def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """modeling_finetune.py:288-299 (float64 numpy, cast to fp32)."""
    tab = np.array([[pos / np.power(10000, 2 * (j // 2) / d_hid) for j in range(d_hid)] for pos in range(n_position)])
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.tensor(tab, dtype=torch.float32).unsqueeze(0)


_POS_CACHE: Dict[Tuple[int, int], torch.Tensor] = {}


def _pos(n: int, d: int) -> torch.Tensor:
    if (n, d) not in _POS_CACHE:
        j = np.arange(d)
        ang = np.arange(n, dtype=np.float64)[:, None] / np.power(10000.0, 2.0 * (j // 2) / d)[None, :]
        tab = np.empty_like(ang)
        tab[:, 0::2] = np.sin(ang[:, 0::2])
        tab[:, 1::2] = np.cos(ang[:, 1::2])
        _POS_CACHE[(n, d)] = torch.tensor(tab, dtype=torch.float32).unsqueeze(0)
    return _POS_CACHE[(n, d)]


def patch_embed(sd: SD, pre: str, x: torch.Tensor, tubelet=(2, 14, 14)) -> torch.Tensor:
    """PatchEmbed.forward, modeling_finetune.py:276-283: Conv3d k=s=tubelet, flatten(2).transpose(1,2)."""
    y = F.conv3d(x, sd[pre + "proj.weight"], sd[pre + "proj.bias"], stride=tubelet)
    return y.flatten(2).transpose(1, 2)


def vit_attention(sd: SD, pre: str, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    """Attention.forward, modeling_finetune.py:169-190."""
    B, N, C = x.shape
    bias = None
    if pre + "q_bias" in sd:
        bias = torch.cat([sd[pre + "q_bias"], torch.zeros_like(sd[pre + "v_bias"]), sd[pre + "v_bias"]])
    qkv = F.linear(x, sd[pre + "qkv.weight"], bias).reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (q.shape[-1] ** -0.5)
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def vit_mlp(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """Mlp.forward, modeling_finetune.py:62-69 (exact-erf GELU)."""
    return F.linear(F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"])), sd[pre + "fc2.weight"],
                    sd[pre + "fc2.bias"])


def vit_block(sd: SD, pre: str, x: torch.Tensor, num_heads: int, eps: float = 1e-6) -> torch.Tensor:
    """Block.forward with gamma=None, modeling_finetune.py:245-248; LN eps 1e-6 (l4p_videomae.py:177)."""
    C = x.shape[-1]
    x = x + vit_attention(sd, pre + "attn.", F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps),
                          num_heads)
    x = x + vit_mlp(sd, pre + "mlp.", F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    return x


def encoder_forward(sd: SD, pre: str, rgb: torch.Tensor, depth: int = 40, num_heads: int = 16,
                    tubelet=(2, 14, 14), eps: float = 1e-6) -> List[torch.Tensor]:
    """VideoMAEEncoder.forward, l4p_videomae.py:80-122: list of depth+1 tensors [B,N,C]; entry 0 = patch-embed +
    pos-embed, entry i = output of block i, last entry replaced by norm(last)."""
    x = patch_embed(sd, pre + "patch_embed.", rgb, tubelet)
    x = x + _pos(x.shape[1], x.shape[2])
    feats = [x]
    for i in range(depth):
        feats.append(vit_block(sd, f"{pre}blocks.{i}.", feats[-1], num_heads, eps))
    C = x.shape[-1]
    feats[-1] = F.layer_norm(feats[-1], (C,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps)
    return feats


# =====================================================================================================
# DPT dense head: task_heads/dpt/dust3r/dpt_head.py:41-86, task_heads/dpt/croco/dpt_block.py
# =====================================================================================================
def _rcu(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """ResidualConvUnit_custom.forward, dpt_block.py:136-157 (ReLU not in place, convs with bias)."""
    out = F.conv3d(F.relu(x), sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    out = F.conv3d(F.relu(out), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return out + x


def _fusion(sd: SD, pre: str, scale, x0: torch.Tensor, x1: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FeatureFusionBlock_custom.forward, dpt_block.py:210-238."""
    out = x0
    if x1 is not None:
        out = out + _rcu(sd, pre + "resConfUnit1.", x1)
    out = _rcu(sd, pre + "resConfUnit2.", out)
    out = F.interpolate(out, scale_factor=tuple(float(s) for s in scale), mode="trilinear", align_corners=True)
    return F.conv3d(out, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def _reassemble(sd: SD, pre: str, x: torch.Tensor, sf: Sequence[int]) -> torch.Tensor:
    """act_postprocess[i] = 1x1x1 conv -> make_conv3d_custom, dpt_block.py:255-278,447-505."""
    x = F.conv3d(x, sd[pre + "0.weight"], sd[pre + "0.bias"])
    if any(s > 0 for s in sf):
        stride = tuple(2 ** s for s in sf)
        x = F.conv_transpose3d(x, sd[pre + "1.weight"], sd[pre + "1.bias"], stride=stride)
    elif any(s < 0 for s in sf):
        stride = tuple(2 ** (-s) for s in sf)
        pad = tuple(s // 2 for s in stride)
        x = F.conv3d(x, sd[pre + "1.weight"], sd[pre + "1.bias"], stride=stride, padding=pad)
    return x


DENSE_ACTPOST = ((1, 2, 2), (1, 1, 1), (0, 0, 0), (-1, -1, -1))
DENSE_FUSION = ((1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
CAMRAY_ACTPOST = ((1, 0, 0), (1, 0, 0), (0, 0, 0), (-1, -1, -1))
CAMRAY_FUSION = ((1, 1, 1), (1, 1, 1), (2, 1, 1), (2, 2, 2))


def dpt_forward(sd: SD, pre: str, feats: Sequence[torch.Tensor], hooks=(14, 21, 28, 36), img_info=(16, 224, 224),
                actpost=DENSE_ACTPOST, fusion=DENSE_FUSION, output_size=None, patch=(2, 14, 14),
                debug: Optional[dict] = None) -> torch.Tensor:
    """DPTOutputAdapter_fix.forward, dpt_head.py:41-86. `pre` ends with 'task_head.dpt.'."""
    T, H, W = img_info
    nt, nh, nw = T // patch[0], H // patch[1], W // patch[2]
    layers = []
    for i, h in enumerate(hooks):
        t = feats[h]
        B, _, C = t.shape
        t = t.reshape(B, nt, nh, nw, C).permute(0, 4, 1, 2, 3).contiguous()
        t = _reassemble(sd, f"{pre}act_postprocess.{i}.", t, actpost[i])
        layers.append(F.conv3d(t, sd[f"{pre}scratch.layer_rn.{i}.weight"], None, padding=1))
    p4 = _fusion(sd, pre + "scratch.refinenet4.", fusion[3], layers[3])[:, :, : layers[2].shape[2], : layers[2].shape[3]]
    p3 = _fusion(sd, pre + "scratch.refinenet3.", fusion[2], p4, layers[2])
    p2 = _fusion(sd, pre + "scratch.refinenet2.", fusion[1], p3, layers[1])
    p1 = _fusion(sd, pre + "scratch.refinenet1.", fusion[0], p2, layers[0])
    out = F.conv3d(p1, sd[pre + "head1.0.weight"], sd[pre + "head1.0.bias"], padding=1)
    if debug is not None:
        debug.update(l0=layers[0], l1=layers[1], l2=layers[2], l3=layers[3], p4=p4, p3=p3, p2=p2, p1=p1, h1=out)
    osz = tuple(img_info) if output_size is None else tuple(output_size)
    if tuple(out.shape[-3:]) != osz:
        out = F.interpolate(out, size=osz, mode="trilinear", align_corners=True)
    out = F.relu(F.conv3d(out, sd[pre + "head2.0.weight"], sd[pre + "head2.0.bias"], padding=1))
    return F.conv3d(out, sd[pre + "head2.2.weight"], sd[pre + "head2.2.bias"])


def apply_fn(x: torch.Tensor, fn_type: str) -> torch.Tensor:
    """l4p/utils/misc.py:11-38 (subset reachable from configs/model.yaml)."""
    if fn_type == "exp":
        return torch.exp(x)
    if fn_type == "linear":
        return x
    if fn_type == "sigmoid":
        return torch.sigmoid(x)
    if fn_type == "log":
        return torch.log(x)
    if fn_type == "inverse":
        out = torch.zeros_like(x)
        m = x.abs() > 1e-8
        out[m] = 1.0 / x[m]
        return out
    raise NotImplementedError(fn_type)


def safe_inverse(x: torch.Tensor, keep_above: float = 0.0) -> torch.Tensor:
    """l4p/utils/misc.py:48-62."""
    out = torch.zeros_like(x)
    m = x > keep_above
    out[m] = 1.0 / x[m]
    return out


def lstsq_affine_solve(pred: torch.Tensor, target: torch.Tensor, inverse: bool = True) -> torch.Tensor:
    """LstSqAffineAligner.solve, aligner.py:45-57 -> [B,2] (scale, shift)."""
    if inverse:
        pred, target = safe_inverse(pred), safe_inverse(target)
    bs = pred.shape[0]
    a = torch.cat([pred.reshape(bs, -1, 1), torch.ones_like(pred.reshape(bs, -1, 1))], dim=-1)
    return torch.linalg.lstsq(a.float(), target.reshape(bs, -1, 1).float()).solution[..., 0]


def lstsq_affine_apply(sol: torch.Tensor, pred: torch.Tensor, inverse: bool = True) -> torch.Tensor:
    """LstSqAffineAligner.apply, aligner.py:59-66."""
    shape = (sol.shape[0],) + (1,) * (pred.ndim - 1)
    p = safe_inverse(pred) if inverse else pred
    p = sol[:, 0].reshape(shape) * p + sol[:, 1].reshape(shape)
    return safe_inverse(p) if inverse else p


def window_starts(T: int, window: int = 16, stride: int = 8) -> List[int]:
    """l4p_videomae.py:267-270."""
    assert T % stride == 0
    return list(range(0, T - window + 1, stride))


def dense_head_windowed(per_window: Sequence[torch.Tensor], starts: Sequence[int], task_name: str,
                        align_inverse_affine: bool, window: int = 16) -> torch.Tensor:
    """VideoMAEFlowDPTHead.forward_windowed, dense_heads.py:76-143, given each window's head output
    [B,C,16,H,W]: depth aligns the whole current window to the buffer on the overlap (LstSqAffineAligner with
    pre_post_fn='inverse'), later windows overwrite; flow skips frame 0 of windows > 0."""
    T = int(starts[-1] + window)
    est = None
    for wi, s in enumerate(starts):
        out = per_window[wi]
        if est is None:
            sz = list(out.shape)
            sz[2] = T
            est = torch.zeros(*sz, dtype=out.dtype)
        if wi > 0 and align_inverse_affine:
            ov = starts[wi - 1] + window - s
            sol = lstsq_affine_solve(out[:, :, :ov], est[:, :, s:s + ov], inverse=True)
            out = lstsq_affine_apply(sol, out, inverse=True)
        if task_name == "flow_2d_backward" and wi > 0:
            est[:, :, s + 1:s + window] = out[:, :, 1:]
        else:
            est[:, :, s:s + window] = out
    return est


# =====================================================================================================
# geometry: l4p/utils/geometry_utils.py
# =====================================================================================================
def normalize_intrinsics(k_b44t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """geometry_utils.py:110-116."""
    k = k_b44t.clone()
    k[:, :2, 2] += 0.5
    k[:, 0] = k[:, 0] / w
    k[:, 1] = k[:, 1] / h
    return k


def denormalize_intrinsics(k_b44t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """geometry_utils.py:119-125."""
    k = k_b44t.clone()
    k[:, 0] *= w
    k[:, 1] *= h
    k[:, :2, 2] -= 0.5
    return k


def plucker_to_point_direction(ray_b6thw: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """geometry_utils.py:308-328."""
    d = ray_b6thw[:, :3]
    m = ray_b6thw[:, 3:] / torch.linalg.norm(d, dim=1, keepdim=True)
    return torch.cross(d, m, dim=1), d


def camera_centers(ray_b6thw: torch.Tensor) -> torch.Tensor:
    """intersect_skew_lines_high_dim on every frame, geometry_utils.py:249-282,362-366 -> [B,T,3]."""
    B, _, T, h, w = ray_b6thw.shape
    o, d = plucker_to_point_direction(ray_b6thw)
    o = o.permute(0, 2, 3, 4, 1).reshape(-1, h * w, 3)
    d = F.normalize(d.permute(0, 2, 3, 4, 1).reshape(-1, h * w, 3), dim=-1)
    eye = torch.eye(3)[None, None]
    imc = eye - d[..., None] * d[..., None, :]
    rhs = imc.matmul(o[..., None]).sum(dim=-3)
    c = torch.linalg.lstsq(imc.sum(dim=-3), rhs).solution[..., 0]
    return c.reshape(B, T, 3)


def kabsch(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """compute_optimal_rotation_alignment, geometry_utils.py:285-305: R minimising ||A - B R||_F, returns R^T."""
    h = (b.T @ a).float()
    u, _, vh = torch.linalg.svd(h, full_matrices=True)
    s = torch.linalg.det(u @ vh)
    r = u @ torch.diag(torch.tensor([1.0, 1.0, float(torch.sign(s))])) @ vh
    return r.T


def _ideal_rays(k_b33t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """geometry_utils.py:372-387: normalised K^-1 (i,j,1) at integer ray-grid coordinates -> [B,T,h,w,3]."""
    B = k_b33t.shape[0]
    j, i = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    pix = torch.stack([i, j, torch.ones_like(i)], dim=-1).expand(B, -1, -1, -1)
    r = torch.einsum("btmn,bhwn->bthwm", torch.inverse(k_b33t.permute(0, 3, 1, 2)), pix)
    return r / r.norm(dim=-1, keepdim=True)


def _extrinsics_from(ray_b6thw: torch.Tensor, rays_d: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    B, _, T, h, w = ray_b6thw.shape
    d = ray_b6thw[:, :3]
    ext = torch.zeros(B, 4, 4, T)
    ext[:, 3, 3] = 1.0
    for b in range(B):
        for t in range(T):
            ext[b, :3, :3, t] = kabsch(rays_d[b, t].reshape(-1, 3), d[b, :, t].reshape(3, -1).T)
    tr = -torch.matmul(ext[:, :3, :3].permute(0, 3, 1, 2), centers[..., None]).squeeze(3)
    ext[:, :3, -1] = tr.permute(0, 2, 1)
    return ext


def rays_to_cameras(ray_b6thw: torch.Tensor, k_norm_b44t: torch.Tensor) -> torch.Tensor:
    """geometry_utils.py:331-406 (ctr_only=False) -> extrinsics [B,4,4,T]."""
    B, _, T, h, w = ray_b6thw.shape
    ray_b6thw = ray_b6thw.float()
    c = camera_centers(ray_b6thw)
    k = denormalize_intrinsics(k_norm_b44t.float(), h, w)[:, :3, :3]
    return _extrinsics_from(ray_b6thw, _ideal_rays(k, h, w), c)


def homography_intrinsics(rays_origin: torch.Tensor, rays_target: torch.Tensor, reproj_threshold: float = 0.2,
                          robust: bool = True) -> torch.Tensor:
    """compute_optimal_rotation_intrinsics, geometry_utils.py:409-456 -> K [3,3].

    robust=True follows the reference literally (cv2.findHomography(RANSAC) + cv2.RQDecomp3x3; third-party,
    version-unpinned). robust=False is the closed form on all z-valid correspondences
    (cv2.findHomography(method=0) = normalised DLT + LM refinement), which equals the RANSAC result whenever
    every correspondence is an inlier; it is the comparator for the device solver."""
    import cv2

    zmask = torch.logical_and(rays_target.abs() > 1e-4, rays_origin.abs() > 1e-4)[:, 2]
    rt, ro = rays_target[zmask], rays_origin[zmask]
    ro = (ro[:, :2] / ro[:, -1:]).numpy()
    rt = (rt[:, :2] / rt[:, -1:]).numpy()
    if robust:
        a, _ = cv2.findHomography(ro, rt, cv2.RANSAC, reproj_threshold)
    else:
        a, _ = cv2.findHomography(ro, rt, 0)
    a = torch.from_numpy(a).float()
    if torch.linalg.det(a) < 0:
        a = -a
    hmat = torch.linalg.inv(a)
    out = cv2.RQDecomp3x3(hmat.numpy())
    k = out[1]
    return torch.from_numpy(k / k[2, 2]).float()


def rays_to_cameras_fixed_intrinsics(ray_b6thw: torch.Tensor, output_size=(224, 224), reproj_threshold: float = 0.2,
                                     robust: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """rays_to_cameras_and_fixed_per_frame_intrinsics, geometry_utils.py:493-579 -> (extrinsics, intrinsics)."""
    ray_b6thw = ray_b6thw.float()
    B, _, T, h, w = ray_b6thw.shape
    c = camera_centers(ray_b6thw)
    d = ray_b6thw[:, :3]
    eye = torch.eye(3)[None, :, :, None].repeat(B, 1, 1, T)
    rays_id = _ideal_rays(eye, h, w)
    kest = torch.zeros(B, 4, 4, T)
    kest[:, 3, 3] = 1.0
    kest[:, 2, 2] = 1.0
    for b in range(B):
        k = homography_intrinsics(rays_id[b, 0].reshape(-1, 3), d[b, :, 0].reshape(3, -1).T, reproj_threshold, robust)
        kest[b, :3, :3, :] = k[:, :, None].repeat(1, 1, T)
    ext = _extrinsics_from(ray_b6thw, _ideal_rays(kest[:, :3, :3], h, w), c)
    H, W = output_size
    return ext, denormalize_intrinsics(normalize_intrinsics(kest, h, w), H, W)


def get_rays_plucker(k_norm_b44t: torch.Tensor, ext_b44t: torch.Tensor, emb_hw=(16, 16)) -> torch.Tensor:
    """geometry_utils.py:165-241 with make_first_cam_ref=True, normalize_dist=False (test generator: the inverse
    of rays_to_cameras)."""
    B, _, _, T = k_norm_b44t.shape
    h, w = emb_hw
    cam_T_world = ext_b44t.permute(0, 3, 1, 2)
    ref_T_cam = torch.matmul(cam_T_world[:, :1], torch.linalg.inv(cam_T_world))
    k = denormalize_intrinsics(k_norm_b44t, h, w)[:, :3, :3]
    rd = _ideal_rays(k, h, w)
    rd = torch.einsum("btmn,bthwn->bthwm", ref_T_cam[..., :3, :3], rd)
    ro = ref_T_cam[..., :3, 3]
    oxd = torch.cross(ro.reshape(B, T, 1, 1, 3).expand_as(rd), rd, dim=-1)
    return torch.cat([rd, oxd], dim=-1).permute(0, 4, 1, 2, 3)


def traj3d_head_window(rays_b6thw: torch.Tensor, k_in_b44t: torch.Tensor, win_id: int, first_k: Optional[torch.Tensor],
                       img_hw=(224, 224), robust: bool = True):
    """VideoMAETraj3DDPTHead.forward with use_intrinsics=False, fixed_intrinsics=True, dense_heads.py:292-352.
    Returns (pose [B,16,T], intrinsics [B,16,T], first_window_intrinsics)."""
    H, W = img_hw
    T = rays_b6thw.shape[2]
    if win_id == 0 or first_k is None:
        ext, kest = rays_to_cameras_fixed_intrinsics(rays_b6thw, (H, W), 0.2, robust)
        first_k = kest.clone()
    else:
        ext = rays_to_cameras(rays_b6thw, normalize_intrinsics(k_in_b44t, H, W).float())
        kest = first_k.clone()
    pose = torch.linalg.inv(ext.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    return pose.reshape(pose.shape[0], 16, T), kest.reshape(kest.shape[0], 16, T), first_k


def generate_point_map(depth_b1thw, k_b44t, pose_b44t):
    """geometry_utils.py:13-53: X = pose [depth K^-1 (u,v,1); 1], integer pixel coordinates."""
    B, _, T, H, W = depth_b1thw.shape
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    pm = torch.zeros(B, 3, T, H, W)
    pm[:, 0] = i
    pm[:, 1] = j
    pm[:, 2] = 1
    kinv = torch.inverse(k_b44t[:, :3, :3].permute(0, 3, 1, 2).float()).permute(0, 2, 3, 1)
    pm = torch.einsum("bmnt,bnthw->bmthw", kinv, pm) * depth_b1thw
    pm4 = torch.cat([pm, torch.ones_like(pm[:, :1])], dim=1)
    return torch.einsum("bmnt,bnthw->bmthw", pose_b44t, pm4)[:, :3]


def umeyama(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """skimage.transform.SimilarityTransform.estimate (Umeyama 1991 with scale): 4x4 T, dst ~ s R src + t.
    skimage is not installed: restated from the published algorithm."""
    n, dim = src.shape
    sm, dm = src.mean(0), dst.mean(0)
    sc, dc = src - sm, dst - dm
    a = dc.T @ sc / n
    d = np.ones(dim)
    if np.linalg.det(a) < 0:
        d[-1] = -1
    u, s, vt = np.linalg.svd(a)
    r = u @ np.diag(d) @ vt
    scale = (s * d).sum() / sc.var(axis=0).sum()
    T = np.eye(dim + 1)
    T[:dim, :dim] = scale * r
    T[:dim, dim] = dm - scale * r @ sm
    return T


def similarity_from_T(T: np.ndarray) -> Dict[str, np.ndarray]:
    """get_similarity_3d_transform's return dict, aligner.py:148-153 (scale = cbrt(det))."""
    s = np.cbrt(np.linalg.det(T[:3, :3]))
    return {"T": T, "R": T[:3, :3] / s, "t": T[:3, 3], "s": np.array(s)}


def similarity_ransac(src: np.ndarray, dst: np.ndarray, threshold: float, min_samples: int = 10, max_trials: int = 100,
                      stop_probability: float = 0.99, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """skimage.measure.ransac(SimilarityTransform) as called at aligner.py:139-146, restated (seeded here; the
    reference's generator is unseeded, so only statistical agreement is meaningful)."""
    rng = np.random.default_rng(seed)
    n = src.shape[0]
    best_inl, best_cnt, best_res = None, 0, np.inf
    trials = 0
    while trials < max_trials:
        trials += 1
        idx = rng.choice(n, min_samples, replace=False)
        T = umeyama(src[idx], dst[idx])
        res = np.linalg.norm(dst - (src @ T[:3, :3].T + T[:3, 3]), axis=1)
        inl = res < threshold
        cnt, rs = int(inl.sum()), float((res ** 2).sum())
        if cnt > best_cnt or (cnt == best_cnt and rs < best_res):
            best_inl, best_cnt, best_res = inl, cnt, rs
            ratio = cnt / n
            denom = 1 - ratio ** min_samples
            need = 0 if denom <= 0 else (np.inf if denom >= 1 else math.ceil(math.log(1 - stop_probability) / math.log(denom)))
            if trials >= need:
                break
    if best_inl is None or best_cnt < min_samples:
        return umeyama(src, dst), np.ones(n, bool)
    return umeyama(src[best_inl], dst[best_inl]), best_inl


def sim3_apply(T: np.ndarray, depth: torch.Tensor, pose_b16t: torch.Tensor):
    """KabaschUmeyama3DAligner.apply, aligner.py:239-265 (B=1)."""
    s = float(np.cbrt(np.linalg.det(T[:3, :3])))
    Tt = torch.from_numpy(T).float()[None]
    bs, _, Tn = pose_b16t.shape
    pose = torch.einsum("bij,bjkt->bikt", Tt, pose_b16t.reshape(bs, 4, 4, Tn)).clone()
    pose[:, :3, :3] = pose[:, :3, :3] / s
    return depth * s, pose.reshape(bs, 16, Tn)


# =====================================================================================================
# track head: task_heads/sparse_heads.py, task_heads/sam/*
# =====================================================================================================
def _lin(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[pre + "weight"], sd[pre + "bias"])


def _ln(sd: SD, pre: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], eps)


def sam_attention(sd: SD, pre: str, q, k, v, num_heads: int = 8) -> torch.Tensor:
    """sam/transformer.py:223-245 (scores divided by sqrt(c_head) after QK^T)."""
    q, k, v = _lin(sd, pre + "q_proj.", q), _lin(sd, pre + "k_proj.", k), _lin(sd, pre + "v_proj.", v)

    def sep(x):
        b, n, c = x.shape
        return x.reshape(b, n, num_heads, c // num_heads).transpose(1, 2)

    q, k, v = sep(q), sep(k), sep(v)
    attn = torch.softmax((q @ k.permute(0, 1, 3, 2)) / math.sqrt(q.shape[-1]), dim=-1)
    out = (attn @ v).transpose(1, 2)
    out = out.reshape(out.shape[0], out.shape[1], -1)
    return _lin(sd, pre + "out_proj.", out)


def two_way_block(sd: SD, pre: str, queries, keys, query_pe, key_pe, skip_first_layer_pe: bool):
    """TwoWayAttentionBlock.forward, sam/transformer.py:156-187 (MLP activation is ReLU, :28,146)."""
    if skip_first_layer_pe:
        queries = sam_attention(sd, pre + "self_attn.", queries, queries, queries)
    else:
        q = queries + query_pe
        queries = queries + sam_attention(sd, pre + "self_attn.", q, q, queries)
    queries = _ln(sd, pre + "norm1.", queries)
    q, k = queries + query_pe, keys + key_pe
    queries = _ln(sd, pre + "norm2.", queries + sam_attention(sd, pre + "cross_attn_token_to_image.", q, k, keys))
    mlp = _lin(sd, pre + "mlp.lin2.", F.relu(_lin(sd, pre + "mlp.lin1.", queries)))
    queries = _ln(sd, pre + "norm3.", queries + mlp)
    q, k = queries + query_pe, keys + key_pe
    keys = _ln(sd, pre + "norm4.", keys + sam_attention(sd, pre + "cross_attn_image_to_token.", k, q, queries))
    return queries, keys


def two_way_transformer(sd: SD, pre: str, image_embedding, image_pe, point_embedding, depth: int = 2):
    """TwoWayTransformer.forward, sam/transformer.py:67-111."""
    queries, keys = point_embedding, image_embedding
    for i in range(depth):
        queries, keys = two_way_block(sd, f"{pre}layers.{i}.", queries, keys, point_embedding, image_pe, i == 0)
    q, k = queries + point_embedding, keys + image_pe
    queries = _ln(sd, pre + "norm_final_attn.", queries + sam_attention(sd, pre + "final_attn_token_to_image.", q, k, keys))
    return queries, keys


def pe_encoding(gauss: torch.Tensor, coords01: torch.Tensor) -> torch.Tensor:
    """PositionEmbeddingRandom3D._pe_encoding, prompt_encoder.py:196-203."""
    c = (2 * coords01 - 1) @ gauss
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(gauss: torch.Tensor, size=(8, 16, 16)) -> torch.Tensor:
    """PositionEmbeddingRandom3D.forward, prompt_encoder.py:205-219 -> [C,t,h,w] (coords stacked t,x,y)."""
    t, h, w = size
    grid = torch.ones(t, h, w)
    te = (grid.cumsum(0) - 0.5) / t
    ye = (grid.cumsum(1) - 0.5) / h
    xe = (grid.cumsum(2) - 0.5) / w
    return pe_encoding(gauss, torch.stack([te, xe, ye], dim=-1)).permute(3, 0, 1, 2)


def prompt_encode(sd: SD, pre: str, coords_n13, labels_n1, feat_n1c, feat_labels_n, image_size=(16, 224, 224)):
    """PromptEncoder.forward with points + features, prompt_encoder.py:78-180 -> [Nq,3,C]
    (point, 'not a point' pad, track-feature)."""
    gauss = sd[pre + "pe_layer.positional_encoding_gaussian_matrix"]
    n = coords_n13.shape[0]
    pts = torch.cat([coords_n13, torch.zeros(n, 1, 3)], dim=1)
    lab = torch.cat([labels_n1, -torch.ones(n, 1)], dim=1)
    c = pts.clone()
    c[:, :, 0] = c[:, :, 0] / image_size[0]
    c[:, :, 1] = c[:, :, 1] / image_size[2]
    c[:, :, 2] = c[:, :, 2] / image_size[1]
    pe = pe_encoding(gauss, c.float())
    pe[lab == -1] = 0.0
    pe[lab == -1] += sd[pre + "not_a_point_embed.weight"]
    i = 0
    while f"{pre}point_embeddings.{i}.weight" in sd:
        pe[lab == i] += sd[f"{pre}point_embeddings.{i}.weight"]
        i += 1
    fe = torch.zeros_like(feat_n1c)
    fe[feat_labels_n == 0] = feat_n1c[feat_labels_n == 0] + sd[pre + "prompt_feature_embeddings.0.weight"]
    fe[feat_labels_n == 1] = feat_n1c[feat_labels_n == 1] + sd[pre + "prompt_feature_embeddings.1.weight"]
    return torch.cat([pe, fe], dim=1)


def layernorm3d(sd: SD, pre: str, x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm3d.forward, sam/mask_decoder.py:145-157."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[pre + "weight"][:, None, None, None] * x + sd[pre + "bias"][:, None, None, None]


def mask_decode(sd: SD, pre: str, image_embeddings_1npc, image_pe_1cthw, sparse, num_mask_tokens: int = 3):
    """MaskDecoder.predict_masks, sam/mask_decoder.py:99-141 -> (masks [Nq,3,16,64,64], io tokens, enc features)."""
    nq = sparse.shape[0]
    tokens = torch.cat([sd[pre + "mask_tokens.weight"].unsqueeze(0).expand(nq, -1, -1), sparse], dim=1)
    src = image_embeddings_1npc[0]
    if src.shape[0] == 1:
        src = torch.repeat_interleave(src, nq, dim=0)
    pos = torch.repeat_interleave(image_pe_1cthw, nq, dim=0)
    b, c, t, h, w = pos.shape
    pos = pos.flatten(2).transpose(1, 2)
    hs, src = two_way_transformer(sd, pre + "transformer.", src, pos, tokens)
    io_features, enc_features = hs.clone(), src.clone()
    hyper = []
    for i in range(num_mask_tokens):
        x = hs[:, i, :]
        p = f"{pre}output_hypernetworks_mlps.{i}.layers."
        x = F.relu(_lin(sd, p + "0.", x))
        x = F.relu(_lin(sd, p + "1.", x))
        hyper.append(_lin(sd, p + "2.", x))
    hyper = torch.stack(hyper, dim=1)
    up = src.transpose(1, 2).reshape(b, c, t, h, w)
    up = F.conv_transpose3d(up, sd[pre + "output_upscaling.0.weight"], sd[pre + "output_upscaling.0.bias"], stride=2)
    up = F.gelu(layernorm3d(sd, pre + "output_upscaling.1.", up))
    up = F.gelu(F.conv_transpose3d(up, sd[pre + "output_upscaling.3.weight"], sd[pre + "output_upscaling.3.bias"],
                                   stride=(1, 2, 2)))
    b, c, t, h, w = up.shape
    masks = (hyper @ up.reshape(b, c, t * h * w)).reshape(b, -1, t, h, w)
    return masks, io_features, enc_features


def softargmax_xy(logits_nthw: torch.Tensor) -> torch.Tensor:
    """VideoMAETrack2DSamHead.softargmax, sparse_heads.py:140-155: pixel-centre (+0.5) expectation -> [N,T,2]."""
    N, T, H, W = logits_nthw.shape
    hm = torch.softmax(logits_nthw.reshape(N, T, 1, H * W), dim=-1)
    gx, gy = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="xy")
    grid = torch.stack([gx, gy], 0).reshape(2, -1) + 0.5
    return (hm * grid[None, None]).sum(-1)


def track_head_window(sd: SD, pre: str, enc_features, queries_bn3, labels_bn, prompt_feats_bnc=None,
                      prompt_labels_bn=None, image_size=(16, 224, 224)) -> Dict[str, torch.Tensor]:
    """VideoMAETrack2DSamHead.forward / forward_single_batch, sparse_heads.py:497-667, B == 1.
    enc_features: [1,P,C] or [1,Nq,P,C] (with per-query history)."""
    assert queries_bn3.shape[0] == 1
    if enc_features.dim() == 3:
        enc_features = enc_features.unsqueeze(1)
    nq = queries_bn3.shape[1]
    C = enc_features.shape[-1]
    coords = queries_bn3[0].unsqueeze(-2)
    labels = labels_bn[0].unsqueeze(-1)
    pf = prompt_feats_bnc[0].unsqueeze(-2) if prompt_feats_bnc is not None else torch.zeros(nq, 1, C)
    pl = prompt_labels_bn[0] if prompt_labels_bn is not None else torch.zeros(nq)
    sparse = prompt_encode(sd, pre + "prompt_encoder.", coords, labels, pf, pl, image_size)
    gauss = sd[pre + "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    emb = (image_size[0] // 2, image_size[1] // 14, image_size[2] // 14)
    masks, io, enc = mask_decode(sd, pre + "mask_decoder.", enc_features[0:1], dense_pe(gauss, emb).unsqueeze(0), sparse)
    logits = F.interpolate(masks, size=image_size, mode="trilinear", align_corners=False)  # [Nq,3,16,224,224]
    out = {}
    out["track_2d_prompt_features_bnc"] = _lin(sd, pre + "prompt_feature_linear_layer.", io[:, 5:6, :])[None][:, :, 0]
    out["track_2d_enc_features_with_track_history_bnpc"] = _lin(sd, pre + "processed_video_features_proj.", enc)[None]
    out["track_2d_traj_est_bn2t"] = softargmax_xy(logits[:, 0]).permute(0, 2, 1)[None]
    out["track_2d_vis_est_bn1t"] = logits[:, 1].mean(dim=[-1, -2])[None].unsqueeze(2)
    out["track_2d_depth_est_bn1t"] = torch.exp(logits[:, 2].mean(dim=[-1, -2]))[None].unsqueeze(2)
    return out


def track_windowed(sd: SD, pre: str, feats_last_per_window: Sequence[torch.Tensor], queries_bn3: torch.Tensor,
                   labels_bn: torch.Tensor, starts: Sequence[int], image_size=(16, 224, 224),
                   patch=(2, 14, 14)) -> Dict[str, torch.Tensor]:
    """VideoMAETrack2DSamHead.forward_windowed_core, sparse_heads.py:213-495, forward direction (sign = +1), B == 1,
    with the shipped flags (prompt_using_features, attend_to_past, modify_pointlabels_for_windowing).
    feats_last_per_window[w]: final-norm encoder tokens [1,P,C] of window w."""
    Tw = image_size[0]
    et, eh, ew = image_size[0] // patch[0], image_size[1] // patch[1], image_size[2] // patch[2]
    B, N = queries_bn3.shape[:2]
    assert B == 1
    C = feats_last_per_window[0].shape[-1]
    Pn = et * eh * ew
    T = int(starts[-1] + Tw)
    traj = torch.zeros(B, N, 2, T)
    vis = -torch.ones(B, N, 1, T) * 10.0
    dep = torch.zeros(B, N, 1, T)
    pfeat = torch.zeros(B, N, C)
    plab = torch.zeros(B, N)
    mask_tok = sd[pre + "processed_video_mask_token.weight"][0]
    hist = mask_tok[None, None, None, :].repeat(B, N, Pn, 1)
    cur_q, cur_lab = queries_bn3.clone(), labels_bn.clone()
    nW = len(starts)
    for wi in range(nW):
        s = int(starts[wi])
        nxt = int(starts[wi + 1]) if wi < nW - 1 else int(starts[wi - 1])
        q_off = cur_q.clone()
        valid_t = (torch.arange(Tw).repeat(B, N, 1) + s + 0.5 - q_off[:, :, 0:1]) >= 0
        valid_t = valid_t[:, :, None, :]
        valid = valid_t.sum(dim=-1)[..., 0] > 0
        q_off[:, :, 0] -= s
        cur_lab[~valid] = 0
        cur_lab[valid] = 1
        same = (cur_q == queries_bn3).sum(dim=-1) > 0
        cur_lab[same] = 1
        cur_lab[torch.logical_and(valid, ~same)] = 2
        enc = feats_last_per_window[wi].unsqueeze(1) + hist
        out = track_head_window(sd, pre, enc, q_off, cur_lab, pfeat, plab, image_size)
        sl = slice(s, s + Tw)
        vis[..., sl][valid_t] = out["track_2d_vis_est_bn1t"][valid_t]
        traj[..., 0:1, sl][valid_t] = out["track_2d_traj_est_bn2t"][:, :, 0:1][valid_t]
        traj[..., 1:2, sl][valid_t] = out["track_2d_traj_est_bn2t"][:, :, 1:2][valid_t]
        dep[..., sl][valid_t] = out["track_2d_depth_est_bn1t"][valid_t]
        if wi == nW - 1:
            continue
        pfeat[valid] = out["track_2d_prompt_features_bnc"][valid]
        plab[valid] = 1
        h = out["track_2d_enc_features_with_track_history_bnpc"].reshape(B, N, et, eh * ew, C)
        pad = mask_tok[None, None, None, None, :].repeat(B, N, et // 2, eh * ew, 1)
        hist = torch.cat([h[:, :, et // 2:], pad], dim=2).reshape(B, N, Pn, C)
        ov0, ov1 = nxt, s + Tw
        best = torch.argmax(vis[..., ov0:ov1], dim=-1)
        new_q = []
        for i in range(N):
            xy = traj[0, i, :, ov0:ov1][:, best[0, i, 0]]
            new_q.append(torch.tensor([float(best[0, i, 0]) + nxt + 0.5, float(xy[0]), float(xy[1])])[None, None])
        new_q = torch.cat(new_q, dim=1)
        later = new_q[0, :, 0] > cur_q[0, :, 0]
        cur_q[0, later, :] = new_q[0, later, :]
    return {"track_2d_traj_est_bn2t": traj, "track_2d_vis_est_bn1t": vis, "track_2d_depth_est_bn1t": dep}
