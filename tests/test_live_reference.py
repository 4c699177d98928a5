"""Host mirror vs the UNMODIFIED reference run live (build container only: /root/reference through oracle/ref_loader.py;
skipped where the tree is absent, e.g. on the GPU box). Fresh seeds, i.e. not the committed fixtures: both sides get the
same key-addressed synthetic weights (`l4p_b200.weights`), the mirror runs on tests/emu.py's per-op torch definitions.

Covers what the fixtures do not: the dynamic-mask head, the camera-ray head's three intrinsics modes including the
reference's first-window quirk (dense_heads.py:327-334), prompt-feature / label inputs of the single-window track head."""
from functools import partial

import pytest
import torch

from l4p_b200 import weights
from oracle import ref_loader
from tests import emu
from tests.util import rel_l2

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

HOOKS = [1, 2, 3, 3]
IMG = (4, 56, 56)


@pytest.fixture()
def cpu_kernels(monkeypatch):
    emu.install(monkeypatch)
    return emu


@pytest.fixture(scope="module")
def ref():
    ref_loader.load()
    import l4p.models.l4p_videomae as RV
    import l4p.models.task_heads.dense_heads as RD
    import l4p.models.task_heads.sparse_heads as RS
    import l4p.utils.geometry_utils as RG

    return dict(V=RV, D=RD, S=RS, G=RG)


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


ENC_KW = dict(img_size=56, patch_size=14, embed_dim=64, depth=3, num_heads=4, mlp_ratio=4, qkv_bias=True,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.0, tubelet_size=2, all_frames=4)


def _pair(ref_cls, our_cls, seed, *args, **kw):
    r, o = ref_cls(*args, **kw).eval(), our_cls(*args, **kw)
    weights.fill_module_(r, seed=seed)
    weights.fill_module_(o, seed=seed)
    assert list(r.state_dict().keys()) == list(o.state_dict().keys())
    return r, o


@torch.no_grad()
def test_dense_heads_windowed_fresh_seed(ref, cpu_kernels):
    from l4p_b200.models.task_heads import dense_heads as D
    from l4p_b200.models.videomae import VideoMAEEncoder

    renc, oenc = _pair(ref["V"].VideoMAEEncoder, VideoMAEEncoder, 41, **ENC_KW)
    oenc.keep_features = "all"
    T, starts = 10, torch.arange(0, 10 - 4 + 1, 2)
    rgb = rnd((1, 3, T, 56, 56), 42)
    intr = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T)
    rf2d = [renc(rgb[:, :, s:s + 4]) for s in starts]
    of2d = [oenc(rgb[:, :, s:s + 4]) for s in starts]
    for i in range(4):
        assert rel_l2(of2d[1][i], rf2d[1][i]) < 1e-3
    cases = [
        ("depth", ref["D"].VideoMAEDepthDPTHead, D.VideoMAEDepthDPTHead,
         dict(depth=3, embed_dim=64, depth_fn="exp", hooks_idx=HOOKS, align_window_overlap_fn="inverse"), "depth_est_b1thw"),
        ("flow_2d_backward", ref["D"].VideoMAEFlowDPTHead, D.VideoMAEFlowDPTHead,
         dict(out_nchan=2, depth=3, embed_dim=64, hooks_idx=HOOKS), "flow_2d_backward_est_b2thw"),
        ("dyn_mask", ref["D"].VideoMAEDynMaskDPTHead, D.VideoMAEDynMaskDPTHead,
         dict(out_nchan=1, depth=3, embed_dim=64, apply_fn="sigmoid", hooks_idx=HOOKS), "dyn_mask_est_b1thw"),
    ]
    for seed, (name, rc, oc, kw, key) in enumerate(cases, start=43):
        rh, oh = _pair(rc, oc, seed, name, **kw)
        want = rh.forward_windowed(rf2d, img_info=IMG, time_strides=starts, intrinsics_b44t=intr)[key]
        got = oh.forward_windowed(of2d, img_info=IMG, time_strides=starts, intrinsics_b44t=intr)[key]
        assert got.shape == want.shape == (1, want.shape[1], T, 56, 56)
        assert rel_l2(got, want) < 1.5e-3, (name, rel_l2(got, want))


class _Rays(torch.nn.Module):
    """Stands in for the reference head's DPT (`task_head`): returns prepared ray maps in call order."""

    def __init__(self, rays):
        super().__init__()
        self.rays, self.i = list(rays), 0

    def forward(self, feats, img_info):
        r = self.rays[self.i % len(self.rays)]
        self.i += 1
        return r


def _camera_windows(G, starts, Tw, T):
    """Valid Plücker ray maps of a synthetic moving camera, one per window (each window in its own first-camera frame,
    as the head would predict them: get_rays_plucker, geometry_utils.py:165-241)."""
    gen = torch.Generator().manual_seed(7)
    K = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2] = 0.9, 1.1, 0.5, 0.5          # normalised
    ext = torch.zeros(1, 4, 4, T)
    for t in range(T):
        qm, _ = torch.linalg.qr(torch.eye(3) + 0.2 * torch.randn(3, 3, generator=gen))
        if torch.linalg.det(qm) < 0:
            qm[:, 0] = -qm[:, 0]
        ext[0, :3, :3, t] = qm
        ext[0, :3, 3, t] = torch.randn(3, generator=gen) * 0.3
        ext[0, 3, 3, t] = 1
    rays = [G.get_rays_plucker(K[..., s:s + Tw], ext[..., s:s + Tw], (4, 4))[0] for s in starts]
    return rays, G.denormalize_intrinsics(K, 56, 56)


@torch.no_grad()
@pytest.mark.parametrize("use_intrinsics,fixed", [(True, False), (False, True)])
def test_camray_head_modes_and_first_window_quirk(ref, cpu_kernels, use_intrinsics, fixed):
    """Both heads are fed the SAME valid ray maps (the DPT is pinned elsewhere), so this isolates the pose logic: mode
    dispatch, `win_id` / first-window intrinsics state, later windows solved with the INPUT intrinsics while reporting
    the first-window estimate, window writes into the [B,16,T] buffer."""
    from l4p_b200.models.task_heads import dense_heads as D

    kw = dict(depth=3, embed_dim=64, hooks_idx=HOOKS, output_size=(4, 4, 4), use_intrinsics=use_intrinsics,
              fixed_intrinsics=fixed)
    rh, oh = _pair(ref["D"].VideoMAETraj3DDPTHead, D.VideoMAETraj3DDPTHead, 51, "camray", **kw)
    T, starts = 8, torch.arange(0, 8 - 4 + 1, 2)
    rays, intr = _camera_windows(ref["G"], [int(s) for s in starts], 4, T)
    it_o = iter(rays)
    rh.task_head = _Rays(rays)
    oh.rays = lambda feats, img_info=IMG: next(it_o)
    dummy = [[torch.zeros(1)]] * len(starts)      # the reference reads dtype / device off the first feature
    want = rh.forward_windowed(dummy, img_info=IMG, time_strides=starts, intrinsics_b44t=intr)["camray_est_b16t"]
    got = oh.forward_windowed(dummy, img_info=IMG, time_strides=starts, intrinsics_b44t=intr)["camray_est_b16t"]
    assert got.shape == want.shape == (1, 16, T)
    assert (got - want).abs().max() < 2e-3, (got - want).abs().max()
    # single-window call: the intrinsics estimate is reported only in the fixed-intrinsics mode
    rh.task_head = _Rays(rays[:1])
    oh.rays = lambda feats, img_info=IMG: rays[0]
    w1 = rh.forward([None], img_info=IMG, intrinsics_b44t=intr[..., :4], win_id=0)
    g1 = oh.forward([None], img_info=IMG, intrinsics_b44t=intr[..., :4], win_id=0)
    assert set(w1) == set(g1)
    for k in w1:
        assert (g1[k] - w1[k]).abs().max() < 2e-3 * max(1.0, float(w1[k].abs().max())), k
    if fixed:
        # window 1 after window 0: pose from the INPUT intrinsics, reported intrinsics = window 0's estimate
        rh.task_head = _Rays(rays[1:2])
        oh.rays = lambda feats, img_info=IMG: rays[1]
        w2 = rh.forward([None], img_info=IMG, intrinsics_b44t=intr[..., 2:6], win_id=1)
        g2 = oh.forward([None], img_info=IMG, intrinsics_b44t=intr[..., 2:6], win_id=1)
        assert torch.equal(g2["camray_intrinsics_est_b16t"], g1["camray_intrinsics_est_b16t"])
        assert (g2["camray_intrinsics_est_b16t"] - w2["camray_intrinsics_est_b16t"]).abs().max() < 2e-3 * 56
        assert (g2["camray_est_b16t"] - w2["camray_est_b16t"]).abs().max() < 2e-3


@torch.no_grad()
def test_track_head_prompt_features_and_labels(ref, cpu_kernels):
    """Single-window track head with explicit prompt features / feature labels and mixed point labels {0,1,2}
    (the inputs the windowed driver feeds from the second window on, sparse_heads.py:497-591)."""
    from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead
    from l4p_b200.models.videomae import VideoMAEEncoder

    renc, oenc = _pair(ref["V"].VideoMAEEncoder, VideoMAEEncoder, 61, **ENC_KW)
    kw = dict(task_name="track_2d", prompt_embed_dim=64, image_size=IMG, estimate_vis=True, estimate_depth=True,
              sam_head_depth=2, num_point_embeddings=2, modify_pointlabels_for_windowing=True, prompt_using_features=True,
              attend_to_past=True, estimation_directions=[1], depth_fn="exp", vis_fn="linear")
    rt, ot = _pair(ref["S"].VideoMAETrack2DSamHead, VideoMAETrack2DSamHead, 62, **kw)
    rgb = rnd((1, 3, 4, 56, 56), 63)
    rf, of = renc(rgb), oenc(rgb)
    q = torch.tensor([[[0.5, 10.5, 12.5], [1.5, 40.5, 30.5], [2.5, 28.0, 28.0], [3.5, 5.5, 50.5], [0.5, 33.3, 8.1]]])
    lab = torch.tensor([[1.0, 2.0, 0.0, 1.0, 2.0]])
    pf = rnd((1, 5, 64), 64, 0.5)
    pl = torch.tensor([[1.0, 0.0, 1.0, 0.0, 1.0]])
    # per-query history tokens (as from the second window on): [B, N, P, C]
    hist = rnd((1, 5, 32, 64), 65, 0.3)
    for enc_r, enc_o in ((rf[-1], of[-1]), (rf[-1].unsqueeze(1) + hist, of[-1].unsqueeze(1) + hist)):
        want = rt.forward([enc_r], q, lab, pf, pl)
        got = ot.forward([enc_o], q, lab, pf, pl)
        assert set(want) == set(got)
        assert (got["track_2d_traj_est_bn2t"] - want["track_2d_traj_est_bn2t"]).abs().max() < 0.02
        assert (got["track_2d_vis_est_bn1t"] - want["track_2d_vis_est_bn1t"]).abs().max() < 3e-3
        assert rel_l2(got["track_2d_depth_est_bn1t"], want["track_2d_depth_est_bn1t"]) < 2e-3
        assert rel_l2(got["track_2d_prompt_features_bnc"], want["track_2d_prompt_features_bnc"]) < 5e-3
        assert rel_l2(got["track_2d_enc_features_with_track_history_bnpc"],
                      want["track_2d_enc_features_with_track_history_bnpc"]) < 5e-3


@torch.no_grad()
def test_full_size_dpt_heads_reference_oracle_mirror(ref, cpu_kernels):
    """At the model's real size (1408-channel taps of a 16x224x224 window, hooks 14/21/28/36 of a 41-entry list) the
    unmodified reference, the oracle restatement and the host mirror (per-op torch definitions) agree: pins the oracle the
    GPU tests compare against at the size they use it, not only at the tiny golden size."""
    from l4p_b200.models.task_heads import dense_heads as D
    from oracle import l4p_oracle as O

    hooks = [14, 21, 28, 36]
    feats = [None] * 41
    for n, i in enumerate(hooks):
        feats[i] = rnd((1, 2048, 1408), 70 + n)
    # depth head: full 224x224 output
    rh, oh = _pair(ref["D"].VideoMAEDepthDPTHead, D.VideoMAEDepthDPTHead, 81, "depth", depth_fn="exp", hooks_idx=hooks,
                   align_window_overlap_fn="inverse")
    want = rh.forward(feats, img_info=(16, 224, 224))["depth_est_b1thw"]
    sd = {k: v.clone() for k, v in rh.state_dict().items()}
    orc = torch.exp(O.dpt_forward(sd, "task_head.dpt.", feats, hooks, img_info=(16, 224, 224)))
    got = oh.forward(feats, img_info=(16, 224, 224))["depth_est_b1thw"]
    assert want.shape == orc.shape == got.shape == (1, 1, 16, 224, 224)
    assert rel_l2(orc, want) < 2e-5, rel_l2(orc, want)
    assert rel_l2(got, want) < 1e-3 and rel_l2(torch.log(got), torch.log(want)) < 2e-3, rel_l2(got, want)
    # camera-ray head: other reassemble / fusion scale factors, fixed 16x16x16 output
    rc, oc = _pair(ref["D"].VideoMAETraj3DDPTHead, D.VideoMAETraj3DDPTHead, 82, "traj3d", hooks_idx=hooks,
                   use_intrinsics=False, fixed_intrinsics=True)
    want = rc.task_head(feats, (16, 224, 224))
    sd = {k: v.clone() for k, v in rc.state_dict().items()}
    orc = O.dpt_forward(sd, "task_head.dpt.", feats, hooks, img_info=(16, 224, 224), actpost=O.CAMRAY_ACTPOST,
                        fusion=O.CAMRAY_FUSION, output_size=(16, 16, 16))
    got = oc.rays(feats, (16, 224, 224))
    assert want.shape == orc.shape == got.shape == (1, 6, 16, 16, 16)
    assert rel_l2(orc, want) < 2e-5
    assert rel_l2(got, want) < 1.5e-3, rel_l2(got, want)


@torch.no_grad()
def test_full_size_track_head_reference_oracle_mirror(ref, cpu_kernels):
    """Track head at its real size (2048 x 1408 tokens, 8 heads of 88 / 176, 16x64x64 mask logits -> 224x224 read-out) for
    two queries, first-window form (shared tokens) and with per-query history: reference == oracle == host mirror."""
    from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead
    from oracle import l4p_oracle as O

    kw = dict(task_name="track_2d", estimate_vis=True, estimate_depth=True, sam_head_depth=2, num_point_embeddings=2,
              modify_pointlabels_for_windowing=True, prompt_using_features=True, attend_to_past=True,
              estimation_directions=[1], depth_fn="exp", vis_fn="linear")
    rt, ot = _pair(ref["S"].VideoMAETrack2DSamHead, VideoMAETrack2DSamHead, 91, **kw)
    sd = {k: v.clone() for k, v in rt.state_dict().items()}
    feat = rnd((1, 2048, 1408), 92)
    q = torch.tensor([[[0.5, 50.5, 60.5], [6.5, 150.5, 100.5]]])
    lab = torch.tensor([[1.0, 2.0]])
    pf = rnd((1, 2, 1408), 93, 0.5)
    pl = torch.tensor([[0.0, 1.0]])
    hist = rnd((1, 2, 2048, 1408), 94, 0.3)
    for enc in (feat, feat.unsqueeze(1) + hist):
        want = rt.forward([enc], q, lab, pf, pl)
        orc = O.track_head_window(sd, "", enc, q, lab, pf, pl)
        got = ot.forward([enc], q, lab, pf, pl)
        for k in ("track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t",
                  "track_2d_prompt_features_bnc", "track_2d_enc_features_with_track_history_bnpc"):
            assert want[k].shape == orc[k].shape == got[k].shape, k
            assert rel_l2(orc[k], want[k]) < 5e-5, (k, rel_l2(orc[k], want[k]))
        assert (got["track_2d_traj_est_bn2t"] - want["track_2d_traj_est_bn2t"]).abs().max() < 0.05      # pixels of 224
        assert (got["track_2d_vis_est_bn1t"] - want["track_2d_vis_est_bn1t"]).abs().max() < 5e-3
        assert rel_l2(got["track_2d_depth_est_bn1t"], want["track_2d_depth_est_bn1t"]) < 2e-3
        assert rel_l2(got["track_2d_prompt_features_bnc"], want["track_2d_prompt_features_bnc"]) < 5e-3
        assert rel_l2(got["track_2d_enc_features_with_track_history_bnpc"],
                      want["track_2d_enc_features_with_track_history_bnpc"]) < 5e-3


@torch.no_grad()
def test_full_width_encoder_reference_oracle_mirror(ref, cpu_kernels):
    """Two ViT-giant blocks at full width on one 16x224x224 window (what `__graft_entry__.smoke()` runs on the GPU against
    the oracle): tubelet patch embed + position table, 16 heads of 88 (padded to 96 in the kernel layout), MLP 6144, final
    norm on the last entry - reference == oracle == host mirror."""
    from l4p_b200.models.videomae import VideoMAEEncoder
    from oracle import l4p_oracle as O
    from tests.util import synth_rgb

    kw = dict(img_size=224, patch_size=14, embed_dim=1408, depth=2, num_heads=16, mlp_ratio=48 / 11, qkv_bias=True,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.0, tubelet_size=2, all_frames=16)
    renc, oenc = _pair(ref["V"].VideoMAEEncoder, VideoMAEEncoder, 0, **kw)
    rgb = synth_rgb(1, 16)
    want = renc(rgb)
    orc = O.encoder_forward({k: v.clone() for k, v in renc.state_dict().items()}, "", rgb, depth=2)
    got = oenc(rgb)
    assert len(want) == len(orc) == len(got) == 3
    for i in range(3):
        assert rel_l2(orc[i], want[i]) < 2e-5, (i, rel_l2(orc[i], want[i]))
        assert rel_l2(got[i], want[i]) < 1.5e-3, (i, rel_l2(got[i], want[i]))      # smoke()'s bound on the GPU


@torch.no_grad()
def test_whole_model_orchestrator_vs_reference(ref, cpu_kernels, monkeypatch):
    """`L4P_VideoMAE.forward` of the unmodified reference (its hard-wired ViT-giant constructor call answered with the tiny
    encoder) against the drop-in on a 10-frame clip = 4 sliding windows, all task heads, per-task alignment
    (l4p_videomae.py:256-330): same output keys, shapes, dtypes; values to 16-bit operand noise; tracker state exact."""
    from l4p_b200.models import l4p_videomae as OV
    from l4p_b200.models.task_heads import dense_heads as D
    from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead
    from l4p_b200.models.videomae import VideoMAEEncoder

    RV, RD, RS = ref["V"], ref["D"], ref["S"]
    renc, oenc = _pair(RV.VideoMAEEncoder, VideoMAEEncoder, 101, **ENC_KW)
    specs = {
        "depth": (RD.VideoMAEDepthDPTHead, D.VideoMAEDepthDPTHead, ("depth",),
                  dict(depth=3, embed_dim=64, depth_fn="exp", hooks_idx=HOOKS, align_window_overlap_fn="inverse")),
        "flow_2d_backward": (RD.VideoMAEFlowDPTHead, D.VideoMAEFlowDPTHead, ("flow_2d_backward",),
                             dict(out_nchan=2, depth=3, embed_dim=64, hooks_idx=HOOKS)),
        "dyn_mask": (RD.VideoMAEDynMaskDPTHead, D.VideoMAEDynMaskDPTHead, ("dyn_mask",),
                     dict(out_nchan=1, depth=3, embed_dim=64, apply_fn="linear", hooks_idx=HOOKS)),
        "track_2d": (RS.VideoMAETrack2DSamHead, VideoMAETrack2DSamHead, (),
                     dict(task_name="track_2d", prompt_embed_dim=64, image_size=IMG, estimate_vis=True, estimate_depth=True,
                          sam_head_depth=2, num_point_embeddings=2, modify_pointlabels_for_windowing=True,
                          prompt_using_features=True, attend_to_past=True, estimation_directions=[1], depth_fn="exp",
                          vis_fn="linear")),
    }
    rheads, oheads = {}, {}
    for seed, (task, (rc, oc, args, kw)) in enumerate(specs.items(), start=102):
        rheads[task], oheads[task] = _pair(rc, oc, seed, *args, **kw)
    monkeypatch.setattr(RV, "VideoMAEEncoder", lambda **kw: renc)        # the reference constructs ViT-giant unconditionally
    common = dict(window_size=IMG, window_stride_T=2, always_use_windowed_version=True, joint_alignment=True)
    rmodel = RV.L4P_VideoMAE(torch.nn.ModuleDict(rheads), **common).eval()
    omodel = OV.L4P_VideoMAE(torch.nn.ModuleDict(oheads), device="meta", **common)
    omodel.video_encoder = oenc
    T = 10
    tasks = list(specs)
    q = torch.tensor([[[0.5, 10.5, 12.5], [3.5, 40.5, 30.5], [8.5, 28.0, 28.0], [5.5, 20.5, 44.5], [11.5, 9.0, 9.0]]])
    data = dict(rgb_b3thw=rnd((1, 3, T, 56, 56), 110), intrinsics_b44t=torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T),
                track_2d_pointquerries_bn3=q, track_2d_pointlabels_bn=torch.ones(1, 5), img_info=IMG, seq_name=["clip"])
    want = rmodel.forward(dict(data), tasks)     # joint_alignment=True but no camray task -> per-task path (:318-328)
    got = omodel.forward(dict(data), tasks)
    assert set(want) == set(got)
    assert len(got["enc_features_bpc_2dlist"]) == len(want["enc_features_bpc_2dlist"]) == 4
    for k, w in want.items():
        if k == "enc_features_bpc_2dlist":
            continue
        g = got[k]
        assert g.shape == w.shape and g.dtype == w.dtype, k
        if k.startswith("track_2d"):
            assert torch.equal(g == 0, w == 0) and torch.equal(g == -10, w == -10), f"{k}: written-frame mask differs"
    assert rel_l2(got["depth_est_b1thw"], want["depth_est_b1thw"]) < 1e-3
    assert rel_l2(got["flow_2d_backward_est_b2thw"], want["flow_2d_backward_est_b2thw"]) < 1.5e-3
    assert rel_l2(got["dyn_mask_est_b1thw"], want["dyn_mask_est_b1thw"]) < 1.5e-3
    assert (got["track_2d_traj_est_bn2t"] - want["track_2d_traj_est_bn2t"]).abs().max() < 0.02
    assert (got["track_2d_vis_est_bn1t"] - want["track_2d_vis_est_bn1t"]).abs().max() < 3e-3
    assert rel_l2(got["track_2d_depth_est_bn1t"], want["track_2d_depth_est_bn1t"]) < 2e-3
    # single-window dispatch (T == window length, always_use_windowed_version=False -> forward_single_window, :262-263):
    # heads are called as head(enc_features_bpc_list=..., **data), outputs carry `enc_features_bpc_list`
    rmodel.always_use_windowed_version = omodel.always_use_windowed_version = False
    one = dict(data, rgb_b3thw=data["rgb_b3thw"][:, :, :4].contiguous(), intrinsics_b44t=data["intrinsics_b44t"][..., :4],
               track_2d_pointquerries_bn3=q[:, :2], track_2d_pointlabels_bn=torch.ones(1, 2))
    want1, got1 = rmodel.forward(dict(one), tasks), omodel.forward(dict(one), tasks)
    assert set(want1) == set(got1) and "enc_features_bpc_list" in got1
    for k, w in want1.items():
        if k == "enc_features_bpc_list":
            continue
        assert got1[k].shape == w.shape, k
        tol = 0.02 if "traj" in k else 5e-3
        assert (got1[k] - w).abs().max() < tol * max(1.0, float(w.abs().max())), (k, (got1[k] - w).abs().max())


class _StubHead(torch.nn.Module):
    """A task head that returns prepared per-window outputs: lets the REFERENCE's joint_windowed_estimation and
    KabaschUmeyama3DAligner run live on a geometrically consistent synthetic scene."""

    def __init__(self, task_name, task_suffix, outputs):
        super().__init__()
        self.task_name, self.task_suffix, self.outputs, self.calls = task_name, task_suffix, outputs, []

    def _out(self, win_id):
        return {f"{self.task_name}_est_{self.task_suffix}": self.outputs[win_id].clone()}

    def forward(self, feats, img_info=IMG, intrinsics_b44t=None, win_id=None, **kw):
        w = int(feats[0].item()) if win_id is None else win_id       # the mirror's depth call carries no win_id
        self.calls.append(w)
        return self._out(w)

    # the drop-in asks the camera head for rays and poses separately (pose solves are stateful per window)
    def rays(self, feats, img_info=IMG):
        return feats[0]

    def pose_from_rays(self, rays, img_info, intrinsics_b44t=None, win_id=None, **kw):
        return self._out(win_id)


@torch.no_grad()
def test_joint_alignment_chain_live_reference_on_consistent_scene(ref, cpu_kernels, monkeypatch):
    """Two overlapping windows whose depth maps and camera poses describe ONE scene in two similarity frames
    (x = s R x' + t). The reference's own chain (dense_heads.py:360-492, aligner.py:158-265; scikit-image is absent, so its
    `ransac` / `SimilarityTransform` are answered by the oracle's restatement) and the drop-in's chain must both bring
    window 1 into window 0's frame: depth * s, pose -> [R R' | s R t' + t], later windows written over the buffer."""
    import numpy as np

    import l4p.models.aligner as RA
    from l4p_b200.models.task_heads import dense_heads as D
    from oracle import l4p_oracle as O

    class SimT:
        def __init__(self, params):
            self.params = params
            self.rotation, self.translation = params[:3, :3], params[:3, 3]
            self.scale = float(np.cbrt(np.linalg.det(params[:3, :3])))

    def ransac(data, model_class, min_samples, residual_threshold, stop_probability, max_trials):
        Tm, inl = O.similarity_ransac(data[0].astype(np.float64), data[1].astype(np.float64), float(residual_threshold),
                                      min_samples=min_samples, max_trials=max_trials, stop_probability=stop_probability)
        return SimT(Tm), inl

    monkeypatch.setattr(RA, "ransac", ransac)
    monkeypatch.setattr(RA, "SimilarityTransform", SimT)
    gen = torch.Generator().manual_seed(3)
    T, Tw, starts = 6, 4, torch.tensor([0, 2])
    H = W = 56
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 50.0
    K[0, 2] = K[1, 2] = 28.0
    intr = K[None, :, :, None].repeat(1, 1, 1, T)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = torch.stack([2.0 + 0.3 * torch.sin(xx / 9 + t) + 0.2 * torch.cos(yy / 7 - t) for t in range(T)])[None, None]
    pose = torch.zeros(1, 4, 4, T)
    for t in range(T):
        qm, _ = torch.linalg.qr(torch.eye(3) + 0.1 * torch.randn(3, 3, generator=gen))
        if torch.linalg.det(qm) < 0:
            qm[:, 0] = -qm[:, 0]
        pose[0, :3, :3, t], pose[0, :3, 3, t], pose[0, 3, 3, t] = qm, torch.randn(3, generator=gen) * 0.2, 1.0
    s_true = 1.6
    Rm, _ = torch.linalg.qr(torch.eye(3) + 0.3 * torch.randn(3, 3, generator=gen))
    if torch.linalg.det(Rm) < 0:
        Rm[:, 0] = -Rm[:, 0]
    t_true = torch.tensor([0.4, -0.3, 0.2])
    # window 1 lives in the primed frame: x' = R^T (x - t) / s
    pose1 = pose[..., 2:6].clone()
    pose1[0, :3, :3] = torch.einsum("ji,jkt->ikt", Rm, pose[0, :3, :3, 2:6])
    pose1[0, :3, 3] = torch.einsum("ji,jt->it", Rm, pose[0, :3, 3, 2:6] - t_true[:, None]) / s_true
    d_w = [depth[:, :, 0:4], depth[:, :, 2:6] / s_true]
    p_w = [pose[..., 0:4].reshape(1, 16, Tw), pose1.reshape(1, 16, Tw)]
    feats = [[torch.tensor([0.0])], [torch.tensor([1.0])]]

    def heads():
        return torch.nn.ModuleDict(dict(depth=_StubHead("depth", "b1thw", d_w), camray=_StubHead("traj3d", "b16t", p_w)))

    np.random.seed(0)
    want = ref["D"].joint_windowed_estimation(["depth", "camray"], heads(), enc_features_bpc_2dlist=feats, time_strides=starts,
                                              intrinsics_b44t=intr, img_info=IMG)
    got = D.joint_windowed_estimation(["depth", "camray"], heads(), feats, time_strides=starts, intrinsics_b44t=intr,
                                      img_info=IMG)
    assert set(want) == set(got) == {"depth_est_b1thw", "traj3d_est_b16t", "traj3d_intrinsics_est_b16t"}
    for out in (want, got):      # both recover the one consistent scene
        assert rel_l2(out["depth_est_b1thw"], depth) < 1e-4
        assert (out["traj3d_est_b16t"] - pose.reshape(1, 16, T)).abs().max() < 1e-3
        assert torch.equal(out["traj3d_intrinsics_est_b16t"], intr.reshape(1, 16, T))
    for k in want:
        assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype
        assert (got[k] - want[k]).abs().max() < 1e-3, k


@torch.no_grad()
@pytest.mark.skipif(__import__("os").environ.get("L4P_SLOW_TESTS", "0") != "1", reason="~2 min, 15 GB: set L4P_SLOW_TESTS=1")
def test_cfg1_full_model_depth_reference_vs_mirror(ref, cpu_kernels):
    """BASELINE.json configs[0] as a parity case: one 16x224x224 clip, depth head only, the reference's full ViT-giant
    (40 blocks) + depth DPT head in PyTorch eager on CPU against the drop-in on the per-op torch definitions."""
    from l4p_b200.models.task_heads import dense_heads as D
    from l4p_b200.models.videomae import VideoMAEEncoder
    from tests.util import synth_rgb

    kw = dict(img_size=224, patch_size=14, embed_dim=1408, depth=40, num_heads=16, mlp_ratio=48 / 11, qkv_bias=True,
              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.0, tubelet_size=2, all_frames=16)
    renc, oenc = _pair(ref["V"].VideoMAEEncoder, VideoMAEEncoder, 0, **kw)
    rh, oh = _pair(ref["D"].VideoMAEDepthDPTHead, D.VideoMAEDepthDPTHead, 1, "depth", depth_fn="exp", hooks_idx=[14, 21, 28, 36],
                   align_window_overlap_fn="inverse")
    rgb = synth_rgb(1, 16)
    rf = renc(rgb)
    want = rh.forward(rf, img_info=(16, 224, 224))["depth_est_b1thw"]
    of = oenc(rgb)
    got = oh.forward(of, img_info=(16, 224, 224))["depth_est_b1thw"]
    taps = {i: rel_l2(of[i], rf[i]) for i in (14, 21, 28, 36, 40)}
    print("cfg1: encoder taps rel-L2", {k: f"{v:.2e}" for k, v in taps.items()}, "depth rel-L2 %.2e" % rel_l2(got, want),
          "log-depth rel-L2 %.2e" % rel_l2(torch.log(got), torch.log(want)))
    assert max(taps.values()) < 1.5e-3
    assert rel_l2(got, want) < 1e-3 and rel_l2(torch.log(got), torch.log(want)) < 2e-3
