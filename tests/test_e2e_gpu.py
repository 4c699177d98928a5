"""End-to-end GPU parity at the BASELINE config sizes (16x224x224 window, ViT-giant, full DPT heads) against the
CPU oracle (oracle/l4p_oracle.py) on identical seeded synthetic weights and inputs.

Tolerances (stated per assertion): fp16 operands / fp32 accumulation and statistics. The reference's own
fp16-autocast run differs from its fp32 run by rel-L2 1.2-1.7e-3 on encoder taps and by max-rel 4.9e-4 on depth
(SURVEY.md §7 hard part 3); north_star asks 1e-3 relative on depth/flow tensors.
"""
import pytest
import torch

from tests.util import max_rel, rel_l2, synth_intrinsics, synth_rgb

pytestmark = pytest.mark.gpu

HOOKS = [14, 21, 28, 36]


@pytest.fixture(scope="module")
def setup():
    from functools import partial

    from l4p_b200 import weights
    from l4p_b200.models.l4p_videomae import L4P_VideoMAE
    from l4p_b200.models.task_heads.dense_heads import (VideoMAEDepthDPTHead, VideoMAEFlowDPTHead,
                                                        VideoMAETraj3DDPTHead)
    from oracle import l4p_oracle as O

    torch.manual_seed(0)
    heads = torch.nn.ModuleDict(dict(
        depth=VideoMAEDepthDPTHead("depth", out_nchan=1, depth_fn="exp", hooks_idx=HOOKS,
                                   align_window_overlap_fn="inverse"),
        flow_2d_backward=VideoMAEFlowDPTHead("flow_2d_backward", out_nchan=2, hooks_idx=HOOKS),
        camray=VideoMAETraj3DDPTHead("traj3d", hooks_idx=HOOKS, use_intrinsics=False, fixed_intrinsics=True),
    ))
    model = L4P_VideoMAE(heads, always_use_windowed_version=True, joint_alignment=False)
    weights.fill_module_(model, seed=0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    rgb = synth_rgb(1, 16)
    data = dict(rgb_b3thw=rgb.cuda(), intrinsics_b44t=synth_intrinsics(1, 16).cuda())
    with torch.no_grad():
        out = model.forward(data, ["depth", "flow_2d_backward", "camray"])
        torch.cuda.synchronize()
        rays = model.task_heads["camray"].last_rays_b6thw
        feats_ref = O.encoder_forward(sd, "video_encoder.", rgb)
    return dict(model=model, sd=sd, out=out, rays=rays, feats_ref=feats_ref, O=O, rgb=rgb)


def test_encoder_taps(setup):
    feats = setup["out"]["enc_features_bpc_2dlist"][0]
    assert len(feats) == 41
    for i in (0, 14, 21, 28, 36, 40):
        got, ref = feats[i], setup["feats_ref"][i]
        assert got.shape == ref.shape and got.dtype == torch.float32
        r = rel_l2(got, ref)
        # fp16 operands, fp32 residual stream: well inside the reference's own fp16-autocast noise (1.2-1.7e-3)
        assert r < 1.5e-3, f"tap {i}: rel-L2 {r:.3e}"
    for i in (1, 13, 39):
        assert feats[i] is None  # documented placeholders (VideoMAEEncoder.keep_features)


def test_depth_head(setup):
    O, sd = setup["O"], setup["sd"]
    with torch.no_grad():
        ref_logit = O.dpt_forward(sd, "task_heads.depth.task_head.dpt.", setup["feats_ref"], HOOKS)
        ref = torch.exp(ref_logit)
    got = setup["out"]["depth_est_b1thw"]
    assert got.shape == (1, 1, 16, 224, 224) and got.dtype == torch.float32
    # (1) weight-independent statement: the network output (log-depth) is reproduced to 16-bit-pipeline accuracy
    rl = rel_l2(torch.log(got), ref_logit)
    # (2) north_star tolerance: 1e-3 relative on the depth tensor itself (depth = exp(logit), so its relative
    #     error is the ABSOLUTE logit error; holds for reference-style initialisation where |logit| is O(0.1-1))
    r, m = rel_l2(got, ref), max_rel(got, ref)
    print(f"depth: logit rel-L2 {rl:.3e} (logit rms {ref_logit.pow(2).mean().sqrt():.3f}); depth rel-L2 {r:.3e} "
          f"max-rel {m:.3e}")
    assert rl < 2e-3, rl
    assert r < 1e-3, r
    assert m < 4e-3, m


def test_flow_head(setup):
    O, sd = setup["O"], setup["sd"]
    with torch.no_grad():
        ref = O.dpt_forward(sd, "task_heads.flow_2d_backward.task_head.dpt.", setup["feats_ref"], HOOKS)
    got = setup["out"]["flow_2d_backward_est_b2thw"]
    assert got.shape == (1, 2, 16, 224, 224)
    r = rel_l2(got, ref)
    print(f"flow rel-L2 {r:.3e}")
    assert r < 1e-3, r


def test_camray_rays_and_pose(setup):
    O, sd = setup["O"], setup["sd"]
    with torch.no_grad():
        ref = O.dpt_forward(sd, "task_heads.camray.task_head.dpt.", setup["feats_ref"], HOOKS, actpost=O.CAMRAY_ACTPOST,
                            fusion=O.CAMRAY_FUSION, output_size=(16, 16, 16))
    got = setup["rays"]
    assert got.shape == (1, 6, 16, 16, 16)
    r = rel_l2(got, ref)
    print(f"rays rel-L2 {r:.3e}")
    # 6-channel *linear* output at the end of a 16-bit conv pyramid: accumulated operand rounding, no exp to hide in
    assert r < 2.5e-3, r
    out = setup["out"]
    assert out["traj3d_est_b16t"].shape == (1, 16, 16)
    # the non-joint windowed path stitches only the pose key, like the reference (dense_heads.py:142)
    assert "traj3d_intrinsics_est_b16t" not in out
    assert torch.isfinite(out["traj3d_est_b16t"]).all()
