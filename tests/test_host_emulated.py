"""Host mirror on CPU against the REFERENCE goldens (tests/golden/golden_small.pt, produced by the unmodified reference:
tests/golden/make_golden.py), with the kernel layer replaced by its plain-torch per-op definitions (tests/emu.py).

What this pins without a GPU: weight packing and layouts (conv K order, ConvT row order, folded positional tables),
the kernel sequence of every module (`VideoMAEEncoder.forward`, `DPTOutputAdapter_fix.forward` incl. the out_conv /
upsample commutation, the two-way transformer / mask decoder of the track head), window batching in
`L4P_VideoMAE._encode_windows`, the stitching rules and the sliding-window tracker state machine -- i.e. SURVEY.md §8
rows a1-a11 above the C ABI. The kernels themselves are held to the same per-op definitions by `pytest -m gpu`.

Tolerances: the emulation keeps the product's rounding points (fp16 operands, fp32 accumulation), so the distance to
the fp32 reference is the 16-bit operand noise of a tiny (dim 64, 3 blocks) model; integer state is exact.
"""
from functools import partial
from pathlib import Path

import pytest
import torch

from l4p_b200 import weights
from tests import emu
from tests.util import rel_l2

GOLD = Path(__file__).resolve().parent / "golden"
HOOKS = [1, 2, 3, 3]
IMG = (4, 56, 56)
Q = torch.tensor([[[0.5, 10.5, 12.5], [1.5, 40.5, 30.5], [0.5, 28.0, 28.0], [5.5, 20.5, 44.5]]])


@pytest.fixture(scope="module")
def g():
    return torch.load(GOLD / "golden_small.pt")


@pytest.fixture()
def cpu_kernels(monkeypatch):
    emu.install(monkeypatch)
    return emu


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def _encoder():
    from l4p_b200.models.videomae import VideoMAEEncoder

    enc = VideoMAEEncoder(img_size=56, patch_size=14, embed_dim=64, depth=3, num_heads=4, mlp_ratio=4, qkv_bias=True,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.0, tubelet_size=2, all_frames=4)
    weights.fill_module_(enc, seed=11)
    enc.keep_features = "all"
    return enc


def _heads():
    from l4p_b200.models.task_heads import dense_heads as D

    depth = D.VideoMAEDepthDPTHead("depth", depth=3, embed_dim=64, depth_fn="exp", hooks_idx=HOOKS,
                                   align_window_overlap_fn="inverse")
    weights.fill_module_(depth, seed=13)
    flow = D.VideoMAEFlowDPTHead("flow_2d_backward", out_nchan=2, depth=3, embed_dim=64, hooks_idx=HOOKS)
    weights.fill_module_(flow, seed=14)
    cam = D.VideoMAETraj3DDPTHead("traj3d", depth=3, embed_dim=64, hooks_idx=HOOKS, output_size=(4, 4, 4),
                                  use_intrinsics=False, fixed_intrinsics=True)
    weights.fill_module_(cam, seed=15)
    return depth, flow, cam


def _tracker(max_queries=192):
    from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead

    trk = VideoMAETrack2DSamHead(task_name="track_2d", prompt_embed_dim=64, image_size=IMG, estimate_vis=True,
                                 estimate_depth=True, sam_head_depth=2, num_point_embeddings=2,
                                 modify_pointlabels_for_windowing=True, prompt_using_features=True, attend_to_past=True,
                                 estimation_directions=[1], depth_fn="exp", vis_fn="linear", max_queries=max_queries)
    weights.fill_module_(trk, seed=17)
    return trk


def test_without_the_emulator_cpu_tensors_are_refused():
    from l4p_b200.lib import L4PError

    with pytest.raises(L4PError, match="no CPU fallback"):
        _encoder()(rnd((1, 3, 4, 56, 56), 12))


def test_encoder_mirror_vs_reference(g, cpu_kernels):
    feats = _encoder()(rnd((1, 3, 4, 56, 56), 12))
    assert len(feats) == 4 and all(f.dtype == torch.float32 for f in feats)
    for i in range(4):
        r = rel_l2(feats[i], g["enc_feats"][i])
        assert r < 1e-3, f"feature {i}: rel-L2 {r:.3e}"
    # the 16-bit operand copies the heads consume are the same tensors, rounded once
    for i, t16 in feats.taps16.items():
        assert torch.equal(t16.view(1, -1, 64), feats[i].to(t16.dtype))
    assert cpu_kernels.CALLS["attention"] == 3 and cpu_kernels.CALLS["linear_qkv"] == 3
    assert cpu_kernels.CALLS["layernorm"] == 2 * 3 + 1


def test_dpt_heads_mirror_vs_reference(g, cpu_kernels):
    feats = _encoder()(rnd((1, 3, 4, 56, 56), 12))
    depth, _, cam = _heads()
    d = depth.forward(feats, img_info=IMG)["depth_est_b1thw"]
    assert d.shape == g["depth_single"].shape
    assert rel_l2(torch.log(d), torch.log(g["depth_single"])) < 1e-3 and rel_l2(d, g["depth_single"]) < 1e-3
    rays = cam.rays(feats, IMG)
    assert rays.shape == g["cam_rays"].shape
    assert rel_l2(rays, g["cam_rays"]) < 1.5e-3


def _windows(enc, rgb, starts):
    return [enc(rgb[:, :, s:s + 4]) for s in starts]


def test_dense_windowed_stitching_mirror_vs_reference(g, cpu_kernels):
    """3 overlapping windows through `forward_windowed` (per-window loop AND the batched `_batched_windows` path)."""
    enc = _encoder()
    depth, flow, _ = _heads()
    rgb = rnd((1, 3, 8, 56, 56), 16)
    starts = torch.arange(0, 8 - 4 + 1, 2)
    intr = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, 8)
    f2d = _windows(enc, rgb, starts)
    batched = enc(torch.cat([rgb[:, :, s:s + 4] for s in starts], dim=0))
    for extra in ({}, {"_batched_windows": batched}):
        d = depth.forward_windowed(f2d, img_info=IMG, time_strides=starts, intrinsics_b44t=intr, **extra)["depth_est_b1thw"]
        f = flow.forward_windowed(f2d, img_info=IMG, time_strides=starts, intrinsics_b44t=intr, **extra)["flow_2d_backward_est_b2thw"]
        assert d.shape == g["depth_windowed"].shape and f.shape == g["flow_windowed"].shape
        assert rel_l2(d, g["depth_windowed"]) < 1e-3, rel_l2(d, g["depth_windowed"])
        assert rel_l2(f, g["flow_windowed"]) < 1.5e-3, rel_l2(f, g["flow_windowed"])


def test_track_single_window_mirror_vs_reference(g, cpu_kernels):
    feats = _encoder()(rnd((1, 3, 4, 56, 56), 12))
    o = _tracker().forward(feats, Q[:, :3], torch.ones(1, 3))
    for k in ("track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t", "track_2d_prompt_features_bnc",
              "track_2d_enc_features_with_track_history_bnpc"):
        ref = g["trk_single/" + k]
        assert o[k].shape == ref.shape, k
    assert (o["track_2d_traj_est_bn2t"] - g["trk_single/track_2d_traj_est_bn2t"]).abs().max() < 0.01      # pixels (of 56)
    assert (o["track_2d_vis_est_bn1t"] - g["trk_single/track_2d_vis_est_bn1t"]).abs().max() < 2e-3
    assert rel_l2(o["track_2d_depth_est_bn1t"], g["trk_single/track_2d_depth_est_bn1t"]) < 1e-3
    assert rel_l2(o["track_2d_prompt_features_bnc"], g["trk_single/track_2d_prompt_features_bnc"]) < 3e-3
    assert rel_l2(o["track_2d_enc_features_with_track_history_bnpc"],
                  g["trk_single/track_2d_enc_features_with_track_history_bnpc"]) < 3e-3


def _check_tracks(o, g):
    for k in ("track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t"):
        v, ref = o[k], g["trk_windowed/" + k]
        assert v.shape == ref.shape, k
        # frames a query never wrote keep the buffer initialisation exactly: the integer state machine matches
        assert torch.equal(v == 0, ref == 0) and torch.equal(v == -10, ref == -10), k
    assert (o["track_2d_traj_est_bn2t"] - g["trk_windowed/track_2d_traj_est_bn2t"]).abs().max() < 0.01
    assert (o["track_2d_vis_est_bn1t"] - g["trk_windowed/track_2d_vis_est_bn1t"]).abs().max() < 2e-3
    assert rel_l2(o["track_2d_depth_est_bn1t"], g["trk_windowed/track_2d_depth_est_bn1t"]) < 2e-3


def test_track_windowed_mirror_vs_reference(g, cpu_kernels):
    """Sliding-window memory tracker over 3 windows: history roll, labels {0,1,2}, argmax-visibility re-query; and the
    `max_queries` chunking (sparse_heads.py:162-211) gives the same tracks."""
    enc = _encoder()
    rgb = rnd((1, 3, 8, 56, 56), 16)
    starts = torch.arange(0, 8 - 4 + 1, 2)
    f2d = _windows(enc, rgb, starts)
    o = _tracker().forward_windowed(f2d, Q, torch.ones(1, 4), time_strides=starts)
    _check_tracks(o, g)
    oc = _tracker(max_queries=2).forward_windowed(f2d, Q, torch.ones(1, 4), time_strides=starts)
    _check_tracks(oc, g)
    for k in o:
        assert torch.equal(o[k] == 0, oc[k] == 0)
        # chunks of 2 vs all 4 queries: the CPU matmuls behind the stand-in block differently per shape, and their fp32
        # round-off passes the 16-bit rounding points of the head (more of them since the folded attention forms): 1.2e-3 px
        assert (o[k] - oc[k]).abs().max() < 3e-3


def _tiny_model(tasks, joint=False):
    """L4P_VideoMAE with the reference's hard-wired ViT-giant (built on the meta device, never run) swapped for the tiny
    golden encoder: exercises the orchestrator itself (window schedule, batching, head dispatch)."""
    from l4p_b200.models.l4p_videomae import L4P_VideoMAE

    depth, flow, cam = _heads()
    heads = {"depth": depth, "flow_2d_backward": flow, "track_2d": _tracker()}
    model = L4P_VideoMAE(torch.nn.ModuleDict({t: heads[t] for t in tasks}), window_size=IMG, window_stride_T=2,
                         always_use_windowed_version=True, joint_alignment=joint, device="meta")
    model.video_encoder = _encoder()
    return model


def test_orchestrator_windowed_forward_vs_reference(g, cpu_kernels):
    tasks = ["depth", "flow_2d_backward", "track_2d"]
    model = _tiny_model(tasks)
    T = 8
    rgb = rnd((1, 3, T, 56, 56), 16)
    data = dict(rgb_b3thw=rgb, intrinsics_b44t=torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T),
                track_2d_pointquerries_bn3=Q, track_2d_pointlabels_bn=torch.ones(1, 4), img_info=IMG)  # the reference's
    # heads default img_info to (16,224,224); `data` is splatted into the head calls, so a tiny window rides along
    out = model.forward(data, tasks)
    assert len(out["enc_features_bpc_2dlist"]) == 3
    assert rel_l2(out["depth_est_b1thw"], g["depth_windowed"]) < 1e-3
    assert rel_l2(out["flow_2d_backward_est_b2thw"], g["flow_windowed"]) < 1.5e-3
    _check_tracks(out, g)
    # windows were encoded as ONE batch (3 blocks x 1 pass), not one pass per window
    assert cpu_kernels.CALLS["attention"] == 3
    # reference assertions (l4p_videomae.py:260,267-269)
    with pytest.raises(AssertionError, match="fixed spatial size"):
        model.forward(dict(data, rgb_b3thw=rnd((1, 3, T, 28, 56), 1)), tasks)
    with pytest.raises(AssertionError, match="multiple of window stride"):
        model.forward(dict(data, rgb_b3thw=rnd((1, 3, 7, 56, 56), 1)), tasks)


def test_orchestrator_single_window_dispatch(g, cpu_kernels):
    """T == window length and always_use_windowed_version=False -> forward_single_window (l4p_videomae.py:262-263)."""
    model = _tiny_model(["depth"])
    model.always_use_windowed_version = False
    out = model.forward(dict(rgb_b3thw=rnd((1, 3, 4, 56, 56), 12), img_info=IMG), ["depth"])
    assert "enc_features_bpc_list" in out and "enc_features_bpc_2dlist" not in out
    assert rel_l2(out["depth_est_b1thw"], g["depth_single"]) < 1e-3


def test_window_chunking_is_transparent(cpu_kernels):
    """`max_windows_per_pass` only changes how many encoder / per-head decoder passes are made, not the result (the decode of
    a long video is chunked too since round 2; the stand-in's CPU convolutions pick batch-size-dependent algorithms, hence
    fp32 round-off that flips 16-bit roundings (~2e-4 rel-L2 on flow, like batched vs per-window decode) instead of bit equality)."""
    tasks = ["flow_2d_backward"]
    rgb = rnd((1, 3, 8, 56, 56), 16)
    outs = []
    for per_pass in (8, 2, 1):
        model = _tiny_model(tasks)
        model.max_windows_per_pass = per_pass
        outs.append(model.forward(dict(rgb_b3thw=rgb, intrinsics_b44t=None, img_info=IMG), tasks)["flow_2d_backward_est_b2thw"])
    assert rel_l2(outs[1], outs[0]) < 1e-3 and rel_l2(outs[2], outs[0]) < 1e-3, (rel_l2(outs[1], outs[0]), rel_l2(outs[2], outs[0]))


def test_two_window_stitching_rules_against_oracle(cpu_kernels):
    """CPU twin of tests/test_windowed_gpu.py (same calls, keys and oracle functions, tiny sizes): a 6-frame clip = two
    overlapping 4-frame windows through `L4P_VideoMAE.forward`, compared with the oracle's stitching of the per-window
    outputs of the same path."""
    from l4p_b200.models.task_heads import dense_heads as D
    from l4p_b200.models.l4p_videomae import L4P_VideoMAE
    from oracle import l4p_oracle as O

    depth, flow, _ = _heads()
    mask = D.VideoMAEDynMaskDPTHead("dyn_mask", out_nchan=1, depth=3, embed_dim=64, apply_fn="linear", hooks_idx=HOOKS)
    weights.fill_module_(mask, seed=19)
    tasks = ["depth", "flow_2d_backward", "dyn_mask"]
    keys = {"depth": "depth_est_b1thw", "flow_2d_backward": "flow_2d_backward_est_b2thw", "dyn_mask": "dyn_mask_est_b1thw"}
    model = L4P_VideoMAE(torch.nn.ModuleDict(dict(depth=depth, flow_2d_backward=flow, dyn_mask=mask)), window_size=IMG,
                         window_stride_T=2, always_use_windowed_version=True, joint_alignment=False, device="meta")
    model.video_encoder = _encoder()
    T, starts = 6, [0, 2]
    rgb = rnd((1, 3, T, 56, 56), 5)
    intr = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T)
    full = model.forward(dict(rgb_b3thw=rgb, intrinsics_b44t=intr, img_info=IMG), tasks)
    assert len(full["enc_features_bpc_2dlist"]) == 2
    per = []
    for s in starts:
        one = model.forward(dict(rgb_b3thw=rgb[:, :, s:s + 4].contiguous(), intrinsics_b44t=intr[..., s:s + 4].contiguous(),
                                 img_info=IMG), tasks)
        per.append({t: one[keys[t]] for t in tasks})
    for t in tasks:
        assert tuple(full[keys[t]].shape[2:]) == (T, 56, 56)
        assert rel_l2(full[keys[t]][:, :, :2], per[0][t][:, :, :2]) < 1e-3       # frames only window 0 writes
    # (batched vs one-at-a-time windows: fp32 summation order differs, 16-bit roundings flip -> up to ~3e-4, not bit-equal)
    for t in ("flow_2d_backward", "dyn_mask"):
        ref = O.dense_head_windowed([w[t] for w in per], starts, t, False, window=4)
        assert rel_l2(full[keys[t]], ref) < 1e-3, t
    flow_full = full[keys["flow_2d_backward"]]
    w0, w1 = per[0]["flow_2d_backward"], per[1]["flow_2d_backward"]
    assert rel_l2(flow_full[:, :, 2], w0[:, :, 2]) < 1e-3 < rel_l2(flow_full[:, :, 2], w1[:, :, 0])   # window 1 frame 0 skipped
    assert rel_l2(flow_full[:, :, 3], w1[:, :, 1]) < 1e-3
    assert rel_l2(full[keys["dyn_mask"]][:, :, 2:], per[1]["dyn_mask"]) < 1e-3                         # later window wins
    ref = O.dense_head_windowed([w["depth"] for w in per], starts, "depth", True, window=4)
    assert rel_l2(full[keys["depth"]], ref) < 1e-3
    raw = O.dense_head_windowed([w["depth"] for w in per], starts, "depth", False, window=4)
    ov_prev = O.safe_inverse(per[0]["depth"][:, :, 2:4])
    assert rel_l2(O.safe_inverse(full[keys["depth"]][:, :, 2:4]), ov_prev) <= rel_l2(O.safe_inverse(raw[:, :, 2:4]), ov_prev) + 1e-5


def test_joint_depth_pose_alignment_chain(cpu_kernels, monkeypatch):
    """`joint_windowed_estimation` (dense_heads.py:360-492) on two windows with a known similarity between them: the
    chain must hand the aligner the overlap slices (depth, pose, intrinsics of both sides), apply the returned
    scale / transform to the new window (aligner.py:239-265) and write later windows over the buffer."""
    from l4p_b200.models.task_heads import dense_heads as D

    depth, _, _ = _heads()
    cam = D.VideoMAETraj3DDPTHead("camray", depth=3, embed_dim=64, hooks_idx=HOOKS, output_size=(4, 4, 4),
                                  use_intrinsics=True, fixed_intrinsics=False)
    weights.fill_module_(cam, seed=15)
    seen = []

    def fake_solve(pred, target, frame_step, thr):
        seen.append(({k: v.clone() for k, v in pred.items()}, {k: v.clone() for k, v in target.items()}, frame_step, thr))
        return sim()

    def sim():
        Tm = torch.eye(4)[None].clone()
        Tm[:, :3, :3] *= 2.0
        Tm[:, :3, 3] = torch.tensor([1.0, -2.0, 0.5])
        return {"T": Tm, "s": torch.tensor([2.0]), "R": Tm[:, :3, :3] / 2.0, "t": Tm[:, :3, 3]}

    import l4p_b200.models.sim3 as S
    monkeypatch.setattr(S, "solve_sim3", fake_solve)
    enc = _encoder()
    T, starts = 6, torch.tensor([0, 2])
    rgb = rnd((1, 3, T, 56, 56), 5)
    k = torch.eye(4)
    k[0, 0] = k[1, 1] = 56.0
    k[0, 2] = k[1, 2] = 28.0
    intr = k[None, :, :, None].repeat(1, 1, 1, T)
    f2d = _windows(enc, rgb, starts)
    heads = torch.nn.ModuleDict(dict(depth=depth, camray=cam))
    out = D.joint_windowed_estimation(["depth", "camray"], heads, f2d, time_strides=starts, intrinsics_b44t=intr, img_info=IMG)
    d0 = depth.forward(f2d[0], img_info=IMG)["depth_est_b1thw"]
    d1 = depth.forward(f2d[1], img_info=IMG)["depth_est_b1thw"]
    p1 = cam.forward(f2d[1], img_info=IMG, intrinsics_b44t=intr[..., 2:6], win_id=1)["camray_est_b16t"]
    assert len(seen) == 1
    pred, target, frame_step, thr = seen[0]
    assert pred["depth"].shape == (1, 1, 2, 56, 56) and torch.equal(pred["depth"], d1[:, :, :2])
    assert torch.equal(target["depth"], d0[:, :, 2:4])
    assert pred["camray"].shape == (1, 16, 2) and pred["camray_intrinsics"].shape == (1, 4, 4, 2)
    assert torch.equal(pred["camray_intrinsics"], intr[..., 2:4]) and (frame_step, thr) == (3, 0.01)
    assert out["depth_est_b1thw"].shape == (1, 1, T, 56, 56) and out["camray_est_b16t"].shape == (1, 16, T)
    assert torch.equal(out["depth_est_b1thw"][:, :, :2], d0[:, :, :2])
    assert torch.allclose(out["depth_est_b1thw"][:, :, 2:], 2.0 * d1)
    pose = torch.einsum("bij,bjkt->bikt", sim()["T"], p1.reshape(1, 4, 4, 4)).clone()
    pose[:, :3, :3] /= 2.0
    assert torch.allclose(out["camray_est_b16t"][:, :, 2:], pose.reshape(1, 16, 4), atol=1e-6)
    assert torch.equal(out["camray_intrinsics_est_b16t"], intr.reshape(1, 16, T))


def test_orchestrator_joint_alignment_path_vs_oracle_chain(cpu_kernels):
    """joint_alignment=True with depth + camray routes both through `joint_windowed_estimation` (l4p_videomae.py:299-311);
    its result must equal the oracle's chain (point maps of every 3rd overlap frame -> similarity -> apply, aligner.py
    :177-265) run on the per-window outputs of the same heads."""
    import numpy as np

    from l4p_b200.models.l4p_videomae import L4P_VideoMAE
    from l4p_b200.models.task_heads import dense_heads as D
    from oracle import l4p_oracle as O

    depth, flow, _ = _heads()
    cam = D.VideoMAETraj3DDPTHead("camray", depth=3, embed_dim=64, hooks_idx=HOOKS, output_size=(4, 4, 4),
                                  use_intrinsics=True, fixed_intrinsics=False)
    weights.fill_module_(cam, seed=15)
    heads = torch.nn.ModuleDict(dict(depth=depth, camray=cam, flow_2d_backward=flow))
    model = L4P_VideoMAE(heads, window_size=IMG, window_stride_T=2, always_use_windowed_version=True, joint_alignment=True,
                         device="meta")
    model.video_encoder = _encoder()
    T, starts = 6, [0, 2]
    rgb = rnd((1, 3, T, 56, 56), 5)
    k = torch.eye(4)
    k[0, 0] = k[1, 1] = 56.0
    k[0, 2] = k[1, 2] = 28.0
    intr = k[None, :, :, None].repeat(1, 1, 1, T)
    tasks = ["depth", "camray", "flow_2d_backward"]
    out = model.forward(dict(rgb_b3thw=rgb, intrinsics_b44t=intr, img_info=IMG), tasks)
    assert cpu_kernels.CALLS["sim3_align"] == 1 and cpu_kernels.CALLS.get("affine_align_solve", 0) == 0
    assert set(["depth_est_b1thw", "camray_est_b16t", "camray_intrinsics_est_b16t", "flow_2d_backward_est_b2thw"]) <= set(out)
    # the chain by hand from single-window outputs
    per = []
    for w, s in enumerate(starts):
        f = model.video_encoder(rgb[:, :, s:s + 4])
        d = depth.forward(f, img_info=IMG)["depth_est_b1thw"]
        p = cam.forward(f, img_info=IMG, intrinsics_b44t=intr[..., s:s + 4], win_id=w)["camray_est_b16t"]
        per.append((d, p))
    (d0, p0), (d1, p1) = per
    Kov = intr[..., 2:4]
    src = O.generate_point_map(d1[:, :, 0:2:3], Kov[..., ::3], p1.reshape(1, 4, 4, 4)[..., 0:2:3])[0].reshape(3, -1).T.double().numpy()
    dst = O.generate_point_map(d0[:, :, 2:4:3], Kov[..., ::3], p0.reshape(1, 4, 4, 4)[..., 2:4:3])[0].reshape(3, -1).T.double().numpy()
    thr = float(torch.quantile(d1[:, :, :2].reshape(-1), 0.98)) * 0.01
    Tm, _ = O.similarity_ransac(src, dst, thr, min_samples=10)
    d1a, p1a = O.sim3_apply(Tm, d1, p1)
    assert rel_l2(out["depth_est_b1thw"][:, :, :2], d0[:, :, :2]) < 1e-3
    assert rel_l2(out["depth_est_b1thw"][:, :, 2:], d1a) < 2e-3
    assert rel_l2(out["camray_est_b16t"][:, :, :2], p0[:, :, :2]) < 1e-3
    assert rel_l2(out["camray_est_b16t"][:, :, 2:], p1a) < 5e-3
    assert np.isfinite(Tm).all()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_tracker_state_machine_random_queries_vs_oracle(cpu_kernels, seed):
    """Random query times / positions (some starting in later windows, some after the clip's last frame -> never valid)
    over 4 windows: written-frame masks, label state machine and argmax re-query must follow the oracle's restatement of
    sparse_heads.py:213-495 exactly; values to 16-bit operand noise."""
    from oracle import l4p_oracle as O

    gen = torch.Generator().manual_seed(100 + seed)
    T, starts = 10, [0, 2, 4, 6]
    enc = _encoder()
    rgb = rnd((1, 3, T, 56, 56), 200 + seed)
    f2d = _windows(enc, rgb, starts)
    n = 7
    q = torch.cat([torch.randint(0, 12, (1, n, 1), generator=gen).float() + 0.5,
                   torch.rand(1, n, 2, generator=gen) * 50 + 3], dim=-1)
    lab = torch.ones(1, n)
    trk = _tracker()
    sd = {k: v.clone() for k, v in trk.state_dict().items()}
    got = trk.forward_windowed(f2d, q, lab, time_strides=torch.tensor(starts))
    ref = O.track_windowed(sd, "", [f[-1] for f in f2d], q, lab, starts, image_size=IMG)
    for k in ("track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t"):
        a, b = got[k], ref[k]
        assert a.shape == b.shape == (1, n, b.shape[2], T)
        assert torch.equal(a == 0, b == 0) and torch.equal(a == -10, b == -10), f"{k}: written-frame mask differs"
    late = q[0, :, 0] > T                                             # queried after the last frame: never written
    assert (got["track_2d_vis_est_bn1t"][0, late] == -10).all()
    assert (got["track_2d_traj_est_bn2t"] - ref["track_2d_traj_est_bn2t"]).abs().max() < 0.02
    assert (got["track_2d_vis_est_bn1t"] - ref["track_2d_vis_est_bn1t"]).abs().max() < 3e-3
    assert rel_l2(got["track_2d_depth_est_bn1t"], ref["track_2d_depth_est_bn1t"]) < 2e-3


def test_bf16_operands_through_the_same_path(g, cpu_kernels):
    """`set_compute_dtype(torch.bfloat16)` (prepare_model precision 'bf16-mixed') switches every operand buffer of the
    encoder and heads; results stay within bf16 operand noise of the reference goldens."""
    model = _tiny_model(["depth", "flow_2d_backward"])
    model.set_compute_dtype(torch.bfloat16)
    T = 8
    rgb = rnd((1, 3, T, 56, 56), 16)
    out = model.forward(dict(rgb_b3thw=rgb, intrinsics_b44t=torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T), img_info=IMG),
                        ["depth", "flow_2d_backward"])
    feats = out["enc_features_bpc_2dlist"][0]
    assert all(t.dtype == torch.bfloat16 for t in model.video_encoder._ws[next(iter(model.video_encoder._ws))].values()
               if t.dtype != torch.float32)
    assert feats[-1].dtype == torch.float32
    assert rel_l2(out["depth_est_b1thw"], g["depth_windowed"]) < 5e-3
    assert rel_l2(out["flow_2d_backward_est_b2thw"], g["flow_windowed"]) < 1.5e-2
