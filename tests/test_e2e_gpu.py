"""End-to-end GPU parity of THE PATH bench.py TIMES: `load_model()` on the shipped configs/model.yaml (windowed path, joint
alignment, all five tasks), one 16x224x224 clip, 128 track queries (BASELINE.json configs[1]), through
`L4PLitModule.predict_step` - every output key against the CPU oracle (oracle/l4p_oracle.py) on identical seeded synthetic
weights and inputs.

Tolerances (stated per assertion): fp16 operands / fp32 accumulation and statistics. The reference's own fp16-autocast run
differs from its fp32 run by rel-L2 1.2-1.7e-3 on encoder taps and by max-rel 4.9e-4 on depth (SURVEY.md §7 hard part 3);
north_star asks 1e-3 relative on depth/flow tensors. The achieved figures are written to gpurun_out/parity_r2.json (copied
to profiles/ and quoted by bench.py / DESIGN.md).
"""
import json
from pathlib import Path

import pytest
import torch

from tests.util import max_rel, rel_l2, synth_intrinsics, synth_rgb

pytestmark = pytest.mark.gpu

HOOKS = [14, 21, 28, 36]
NQ = 128
TASKS = ["flow_2d_backward", "track_2d", "depth", "dyn_mask", "camray"]
PARITY = {}


def _record(**kw):
    PARITY.update({k: float(f"{v:.4g}") for k, v in kw.items()})
    out = Path(__file__).resolve().parents[1] / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_r2.json").write_text(json.dumps(dict(
        PARITY, source="pytest -m gpu tests/test_e2e_gpu.py on B200: shipped config, one 16x224x224 clip, 128 queries, fp16 "
                       "operands, vs the fp32 CPU oracle on the same seeded weights"), indent=1))


def bench_queries():
    xs = torch.linspace(7, 217, 16)
    ys = torch.linspace(14, 210, 8)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([torch.full_like(gx, 0.5), gx + 0.5, gy + 0.5], dim=-1).reshape(1, NQ, 3)


@pytest.fixture(scope="module")
def setup():
    from l4p_b200 import weights
    from l4p_b200.config import load_model
    from oracle import l4p_oracle as O

    torch.manual_seed(0)
    lit = load_model(device=torch.device("cuda"), max_queries=NQ + 1)     # configs/model.yaml as shipped
    model = lit.l4p_model
    assert model.joint_alignment and model.always_use_windowed_version and lit.tasks == TASKS
    weights.fill_module_(model, seed=0)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    rgb = synth_rgb(1, 16)
    q = bench_queries()
    q[0, 5, 0] = 6.5          # one query that starts mid-window (frames before it keep the buffer initialisation)
    batch = dict(rgb_b3thw=rgb.clone(), intrinsics_b44t=synth_intrinsics(1, 16), track_2d_pointquerries_bn3=q.clone(),
                 track_2d_pointlabels_bn=torch.ones(1, NQ))
    with torch.no_grad():
        out = lit.predict_step(batch, 0)          # host batch in, like the reference's predict loop (l4p/l4p.py:54-66,107-109)
        torch.cuda.synchronize()
        rays = model.task_heads["camray"].last_rays_b6thw
        feats_ref = O.encoder_forward(sd, "video_encoder.", rgb)
    return dict(model=model, sd=sd, out=out, rays=rays, feats_ref=feats_ref, O=O, rgb=rgb, q=q)


def test_output_keys_of_the_shipped_config(setup):
    out = setup["out"]
    want = {"enc_features_bpc_2dlist", "depth_est_b1thw", "flow_2d_backward_est_b2thw", "dyn_mask_est_b1thw",
            "traj3d_est_b16t", "traj3d_intrinsics_est_b16t", "track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t",
            "track_2d_depth_est_bn1t"}
    assert set(out) == want, set(out) ^ want
    for k, v in out.items():
        if torch.is_tensor(v):
            assert v.is_cuda and torch.isfinite(v).all(), k


def test_encoder_taps(setup):
    feats = setup["out"]["enc_features_bpc_2dlist"][0]
    assert len(feats) == 41
    worst = 0.0
    for i in (0, 14, 21, 28, 36, 40):
        got, ref = feats[i], setup["feats_ref"][i]
        assert got.shape == ref.shape and got.dtype == torch.float32
        r = rel_l2(got, ref)
        worst = max(worst, r)
        # fp16 operands, fp32 residual stream: well inside the reference's own fp16-autocast noise (1.2-1.7e-3)
        assert r < 1.5e-3, f"tap {i}: rel-L2 {r:.3e}"
    _record(encoder_taps_rel_l2_max=worst)
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
    _record(depth_rel_l2=r, depth_max_rel=m, depth_logit_rel_l2=rl)
    assert rl < 2e-3, rl
    assert r < 1e-3, r
    # worst single pixel of 802 816 (measured on B200, round 2: 1.3e-4; the reference's own fp16-autocast run reaches 4.9e-4
    # on its worst pixel): the north_star bound holds per pixel, not only in the L2 sense
    assert m < 1e-3, m


def test_flow_head(setup):
    O, sd = setup["O"], setup["sd"]
    with torch.no_grad():
        ref = O.dpt_forward(sd, "task_heads.flow_2d_backward.task_head.dpt.", setup["feats_ref"], HOOKS)
    got = setup["out"]["flow_2d_backward_est_b2thw"]
    assert got.shape == (1, 2, 16, 224, 224)
    r = rel_l2(got, ref)
    # flow changes sign, so a per-pixel relative error is meaningless at its zero crossings: worst pixel relative to the rms
    m = ((got.cpu() - ref).abs().max() / ref.pow(2).mean().sqrt()).item()
    print(f"flow rel-L2 {r:.3e}, max abs err / rms {m:.3e}")
    _record(flow_rel_l2=r, flow_max_abs_over_rms=m)
    assert r < 1e-3, r
    assert m < 6e-3, m


def test_dyn_mask_head(setup):
    O, sd = setup["O"], setup["sd"]
    with torch.no_grad():
        ref = O.dpt_forward(sd, "task_heads.dyn_mask.task_head.dpt.", setup["feats_ref"], HOOKS)   # apply_fn: linear
    got = setup["out"]["dyn_mask_est_b1thw"]
    assert got.shape == (1, 1, 16, 224, 224)
    r = rel_l2(got, ref)
    m = ((got.cpu() - ref).abs().max() / ref.pow(2).mean().sqrt()).item()
    print(f"dyn-mask logits rel-L2 {r:.3e}, max abs err / rms {m:.3e}")
    _record(dyn_mask_rel_l2=r, dyn_mask_max_abs_over_rms=m)
    # a raw logit map (apply_fn: linear) of small dynamic range, the same quantity as log-depth above (1.1e-3, bound 2e-3);
    # not one of the tensors north_star puts the 1e-3 bound on (depth / flow)
    assert r < 2e-3, r
    assert m < 8e-3, m


def test_camray_rays_and_pose(setup):
    O, sd = setup["O"], setup["sd"]
    with torch.no_grad():
        ref = O.dpt_forward(sd, "task_heads.camray.task_head.dpt.", setup["feats_ref"], HOOKS, actpost=O.CAMRAY_ACTPOST,
                            fusion=O.CAMRAY_FUSION, output_size=(16, 16, 16))
    got = setup["rays"]
    assert got.shape == (1, 6, 16, 16, 16)
    r = rel_l2(got, ref)
    print(f"rays rel-L2 {r:.3e}")
    _record(rays_rel_l2=r)
    # 6-channel *linear* output at the end of a 16-bit conv pyramid: accumulated operand rounding, no exp to hide in
    assert r < 2.5e-3, r
    out = setup["out"]
    pose = out["traj3d_est_b16t"].cpu().reshape(1, 4, 4, 16)
    kest = out["traj3d_intrinsics_est_b16t"].cpu().reshape(1, 4, 4, 16)
    # joint path, first window: fixed intrinsics estimated from frame 0 and reported for every frame (dense_heads.py:327-334)
    assert (kest - kest[..., :1]).abs().max() == 0
    # pose = inverse extrinsics: a rigid transform per frame whose translation is the least-squares camera centre of the
    # ray bundle (geometry_utils.py:249-282) - a closed form that does not depend on the intrinsics estimate
    R = pose[0, :3, :3].permute(2, 0, 1)
    assert (R @ R.transpose(1, 2) - torch.eye(3)).abs().max() < 1e-4 and (torch.linalg.det(R) - 1).abs().max() < 1e-4
    centers = O.camera_centers(got.cpu().float())                      # [1,T,3] from the SAME rays the device solver saw
    ct = (pose[0, :3, 3].T - centers[0]).abs().max().item()
    # the remaining part (K from a homography fit, R from Kabsch against the ideal rays of that K) against the oracle's
    # closed-form comparator on the same rays; random-weight ray maps are not a camera, so the fit is loose by nature and the
    # numbers are reported, the hard check of this solver is tests/test_geometry_gpu.py on the reference's golden rays
    rp, rk, _ = O.traj3d_head_window(got.cpu().float(), synth_intrinsics(1, 16), 0, None, robust=False)
    dk = ((kest - rk.reshape(1, 4, 4, 16)).abs().max() / rk.abs().max()).item()
    dp = (pose - rp.reshape(1, 4, 4, 16)).abs().max().item()
    print(f"pose: centre err {ct:.2e}; vs oracle closed form on the same rays: K rel {dk:.2e}, pose abs {dp:.2e}")
    _record(pose_center_abs=ct, pose_vs_oracle_abs=dp, intrinsics_vs_oracle_rel=dk)
    assert ct < 1e-3


def test_track_head_128_queries(setup):
    """All 128 bench queries through the windowed driver (first-window semantics) vs the oracle."""
    O, sd, q = setup["O"], setup["sd"], setup["q"]
    pre = "task_heads.track_2d."
    lab = torch.ones(1, NQ)
    with torch.no_grad():
        enc = setup["feats_ref"][40] + sd[pre + "processed_video_mask_token.weight"][0]
        ref = O.track_head_window(sd, pre, enc, q, lab, torch.zeros(1, NQ, 1408), torch.zeros(1, NQ))
    out = setup["out"]
    traj, vis, dep = (out[k].cpu() for k in ("track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t"))
    assert traj.shape == (1, NQ, 2, 16) and vis.shape == (1, NQ, 1, 16) and dep.shape == (1, NQ, 1, 16)
    valid = (torch.arange(16).view(1, 1, 1, 16) + 0.5 - q[:, :, 0:1, None]) >= 0
    assert (~valid).sum() == 6
    # frames before the query time keep the buffer initialisation (exact): traj 0, vis -10, depth 0
    assert (traj[~valid.expand_as(traj)] == 0).all() and (vis[~valid] == -10).all() and (dep[~valid] == 0).all()
    rt, rv, rd = ref["track_2d_traj_est_bn2t"], ref["track_2d_vis_est_bn1t"], ref["track_2d_depth_est_bn1t"]
    err_px = (traj - rt)[valid.expand_as(traj)].abs().max().item()
    ev = (vis - rv)[valid].abs().max().item()
    ed = rel_l2(dep[valid], rd[valid])
    print(f"tracks (128 queries): traj max err {err_px:.4f} px; vis max abs err {ev:.3e}; depth rel {ed:.3e}")
    _record(track_traj_max_px=err_px, track_vis_max_abs=ev, track_depth_rel_l2=ed)
    assert err_px < 0.05                                        # pixels (soft-argmax over 224x224)
    assert ev < 5e-3                                            # logits
    assert ed < 2e-3
