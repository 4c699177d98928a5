"""Multi-window dense inference on the GPU: a 24-frame clip = two overlapping 16-frame
windows (t = 0 and t = 8, `l4p_videomae.py:267-270`) through `L4P_VideoMAE.forward`, i.e. all windows batched through
the encoder / DPT kernels, then the reference's stitching rules (`dense_heads.py:76-143`): depth aligns each new window
to the buffer on the 8 overlap frames with the inverse-depth least-squares affine (`aligner.py:45-66`, device solver
`l4p_affine_align_solve/apply`), flow skips frame 0 of later windows, the dynamic mask is overwritten.

Checked against the CPU oracle's stitching (`oracle.l4p_oracle.dense_head_windowed`, pinned on reference goldens by
tests/test_oracle_golden.py::test_dense_windowed_stitching) fed with the per-window head outputs of the same GPU path
run one window at a time, so the comparison isolates batching + stitching + alignment; the per-window heads themselves
are held to the oracle at full size by tests/test_e2e_gpu.py.

Tolerances: batched vs one-at-a-time windows differ by 16-bit rounding noise of order-dependent split-K sums
(<= 5e-4 rel-L2 measured); depth goes through `exp` and an fp32 (oracle) vs fp64 (device) solve.
"""
import pytest
import torch

from tests.util import rel_l2, synth_intrinsics, synth_rgb

pytestmark = pytest.mark.gpu

HOOKS = [14, 21, 28, 36]
TASKS = ["depth", "flow_2d_backward", "dyn_mask"]
KEYS = {"depth": "depth_est_b1thw", "flow_2d_backward": "flow_2d_backward_est_b2thw", "dyn_mask": "dyn_mask_est_b1thw"}


@pytest.fixture(scope="module")
def run():
    from l4p_b200 import weights
    from l4p_b200.models.l4p_videomae import L4P_VideoMAE
    from l4p_b200.models.task_heads.dense_heads import (VideoMAEDepthDPTHead, VideoMAEDynMaskDPTHead,
                                                        VideoMAEFlowDPTHead)

    heads = torch.nn.ModuleDict(dict(
        depth=VideoMAEDepthDPTHead("depth", out_nchan=1, depth_fn="exp", hooks_idx=HOOKS,
                                   align_window_overlap_fn="inverse"),
        flow_2d_backward=VideoMAEFlowDPTHead("flow_2d_backward", out_nchan=2, hooks_idx=HOOKS),
        dyn_mask=VideoMAEDynMaskDPTHead("dyn_mask", out_nchan=1, apply_fn="linear", hooks_idx=HOOKS),
    ))
    model = L4P_VideoMAE(heads, always_use_windowed_version=True, joint_alignment=False)
    weights.fill_module_(model, seed=1)
    T, starts = 24, [0, 8]
    rgb = synth_rgb(1, T, seed=5)
    intr = synth_intrinsics(1, T)
    with torch.no_grad():
        full = model.forward(dict(rgb_b3thw=rgb.cuda(), intrinsics_b44t=intr.cuda()), TASKS)
        per_window = []
        for s in starts:
            one = model.forward(dict(rgb_b3thw=rgb[:, :, s:s + 16].contiguous().cuda(),
                                     intrinsics_b44t=intr[..., s:s + 16].contiguous().cuda()), TASKS)
            per_window.append({t: one[KEYS[t]].float().cpu() for t in TASKS})
        torch.cuda.synchronize()
    return dict(full={t: full[KEYS[t]].float().cpu() for t in TASKS}, per_window=per_window, starts=starts, T=T,
                n_windows=len(full["enc_features_bpc_2dlist"]))


def test_window_schedule_and_shapes(run):
    assert run["n_windows"] == 2
    assert tuple(run["full"]["depth"].shape) == (1, 1, run["T"], 224, 224)
    assert tuple(run["full"]["flow_2d_backward"].shape) == (1, 2, run["T"], 224, 224)
    assert tuple(run["full"]["dyn_mask"].shape) == (1, 1, run["T"], 224, 224)
    for t in TASKS:
        assert torch.isfinite(run["full"][t]).all()
    assert (run["full"]["depth"][:, :, :8] > 0).all()


def test_first_window_frames_are_the_single_window_result(run):
    # frames [0, 8) are written by window 0 only and never touched again
    for t in TASKS:
        r = rel_l2(run["full"][t][:, :, :8], run["per_window"][0][t][:, :, :8])
        assert r < 3e-3, f"{t}: frames 0-7 rel-L2 {r:.3e}"


def test_flow_and_mask_stitching_rules(run):
    from oracle import l4p_oracle as O

    for t in ("flow_2d_backward", "dyn_mask"):
        ref = O.dense_head_windowed([w[t] for w in run["per_window"]], run["starts"], t, False)
        r = rel_l2(run["full"][t], ref)
        assert r < 3e-3, f"{t}: stitched rel-L2 {r:.3e}"
    # the rule itself: frame 8 of the flow buffer is window 0's (frame 0 of window 1 is invalid), frame 9 is window 1's;
    # the dynamic mask takes window 1 from frame 8 on
    flow, w0, w1 = run["full"]["flow_2d_backward"], run["per_window"][0]["flow_2d_backward"], run["per_window"][1]["flow_2d_backward"]
    assert rel_l2(flow[:, :, 8], w0[:, :, 8]) < 3e-3 and rel_l2(flow[:, :, 8], w0[:, :, 8]) < rel_l2(flow[:, :, 8], w1[:, :, 0])
    assert rel_l2(flow[:, :, 9], w1[:, :, 1]) < 3e-3
    mask, m1 = run["full"]["dyn_mask"], run["per_window"][1]["dyn_mask"]
    assert rel_l2(mask[:, :, 8:], m1) < 3e-3


def test_depth_overlap_alignment_vs_oracle(run):
    from oracle import l4p_oracle as O

    per = [w["depth"] for w in run["per_window"]]
    ref = O.dense_head_windowed(per, run["starts"], "depth", True)
    got = run["full"]["depth"]
    r = rel_l2(got, ref)
    assert r < 5e-3, f"aligned + stitched depth rel-L2 {r:.3e}"
    raw = O.dense_head_windowed(per, run["starts"], "depth", False)
    print(f"depth: stitched vs oracle {r:.3e}; effect of the alignment on frames 8-23: {rel_l2(raw[:, :, 8:], ref[:, :, 8:]):.3e}")
    # the least-squares property: on the overlap frames the aligned window agrees with window 0 (in inverse depth) at
    # least as well as the raw one (identity is a feasible affine map), up to the batching noise
    ov_prev = O.safe_inverse(per[0][:, :, 8:16])
    assert rel_l2(O.safe_inverse(got[:, :, 8:16]), ov_prev) <= rel_l2(O.safe_inverse(per[1][:, :, :8]), ov_prev) + 2e-3


def test_track_query_chunking_matches_single_pass():
    """Query chunking by `max_queries` (sparse_heads.py:162-211): queries are independent, so 6 queries tracked over two
    windows in chunks of 3 must give what one pass over all 6 gives (written-frame masks exactly, values to kernel
    noise: the chunked GEMMs see different M)."""
    from l4p_b200 import weights
    from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead

    def head(max_queries):
        h = VideoMAETrack2DSamHead(task_name="track_2d", estimate_vis=True, estimate_depth=True, sam_head_depth=2,
                                   num_point_embeddings=2, prompt_using_features=True, attend_to_past=True,
                                   modify_pointlabels_for_windowing=True, estimation_directions=[1], depth_fn="exp",
                                   vis_fn="linear", max_queries=max_queries)
        weights.fill_module_(h, seed=3)
        return h.cuda()

    g = torch.Generator().manual_seed(11)
    feats = [torch.randn(1, 2048, 1408, generator=g) for _ in range(2)]
    f2d = [[None] * 40 + [f.cuda()] for f in feats]
    q = torch.tensor([[[0.5, 50.5, 60.5], [3.5, 150.5, 100.5], [12.5, 100.5, 180.5], [18.5, 30.5, 200.5],
                       [0.5, 200.5, 20.5], [9.5, 112.5, 112.5]]]).cuda()
    lab = torch.ones(1, 6).cuda()
    ts = torch.tensor([0, 8])
    with torch.no_grad():
        one = head(192).forward_windowed(f2d, q, lab, time_strides=ts)
        chunked = head(3).forward_windowed(f2d, q, lab, time_strides=ts)
    torch.cuda.synchronize()
    for k in ("track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t"):
        a, b = chunked[k].cpu(), one[k].cpu()
        assert a.shape == b.shape and a.shape[1] == 6 and a.shape[-1] == 24
        assert torch.equal(a == 0, b == 0) and torch.equal(a == -10, b == -10), f"{k}: written-frame mask differs"
    assert (chunked["track_2d_traj_est_bn2t"] - one["track_2d_traj_est_bn2t"]).abs().max().item() < 0.05   # pixels
    assert (chunked["track_2d_vis_est_bn1t"] - one["track_2d_vis_est_bn1t"]).abs().max().item() < 5e-3
    assert rel_l2(chunked["track_2d_depth_est_bn1t"], one["track_2d_depth_est_bn1t"]) < 2e-3
