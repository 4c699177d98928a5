"""Host-side logic that needs no GPU: window schedule, config instantiation, deterministic weights, tiling helpers."""
import torch

from l4p_b200 import ops, weights
from oracle import l4p_oracle as O


def test_window_starts_reference_rule():
    # time_strides = arange(0, T-16+1, 8)  (l4p_videomae.py:270): T=16 -> [0]; T=512 -> 63 windows; T=264 -> 32
    assert O.window_starts(16) == [0]
    assert len(O.window_starts(512)) == 63 and O.window_starts(512)[-1] == 496
    assert len(O.window_starts(264)) == 32
    assert torch.arange(0, 512 - 16 + 1, 8).tolist() == O.window_starts(512)


def test_synthetic_weights_are_deterministic_and_key_addressed():
    a = weights.synth_tensor("video_encoder.blocks.3.attn.qkv.weight", (12, 8), seed=0)
    b = weights.synth_tensor("video_encoder.blocks.3.attn.qkv.weight", (12, 8), seed=0)
    c = weights.synth_tensor("video_encoder.blocks.4.attn.qkv.weight", (12, 8), seed=0)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert weights.synth_tensor("x.norm1.weight", (8,)).mean().item() > 0.8  # norm scales ~1
    assert weights.synth_tensor("x.bias", (8,)).abs().max().item() <= 0.02
    assert not torch.equal(weights.synth_tensor("attn.qkv.weight", (6, 4), peaky=True),
                           weights.synth_tensor("attn.qkv.weight", (6, 4)))


def test_config_instantiates_reference_yaml_schema():
    """Both our configs/model.yaml and a reference-style config (class paths under `l4p.`) build the same tree."""
    import yaml

    from l4p_b200.config import DEFAULT_CONFIG, load_model

    cfg = yaml.safe_load(open(DEFAULT_CONFIG))
    txt = yaml.safe_dump(cfg, sort_keys=False).replace("l4p_b200.", "l4p.")
    ref_style = yaml.safe_load(txt)
    assert ref_style["class_path"] == "l4p.l4p.L4PLitModule"
    m1 = load_model(device="meta")
    m2 = load_model(ref_style, device="meta", max_queries=64)
    assert type(m1).__name__ == "L4PLitModule" and list(m1.state_dict()) == list(m2.state_dict())
    assert m1.tasks == ["flow_2d_backward", "track_2d", "depth", "dyn_mask", "camray"]
    assert m2.l4p_model.task_heads["track_2d"].max_queries == 64
    assert m1.l4p_model.always_use_windowed_version and m1.l4p_model.joint_alignment
    assert m1.l4p_model.task_heads["camray"].task_suffix == "b16t"


def test_conv_box_tiling():
    for shape in [(16, 224, 224), (16, 128, 128), (16, 64, 64), (16, 32, 32), (8, 16, 16), (4, 8, 8), (16, 16, 16), (3, 20, 24)]:
        bt, bh, bw = ops.pick_box(*shape)
        assert bt * bh * bw == 128
    assert ops.pick_box(4, 8, 8) == (2, 8, 8)
    T, H, W = 16, 224, 224
    bt, bh, bw = ops.pick_box(T, H, W)
    assert T % bt == 0 and H % bh == 0 and W % bw == 0  # no padded voxels at the headline resolution


def test_feature_list_placeholders():
    from l4p_b200.models.videomae import FeatureList

    f = FeatureList([None, torch.zeros(1), None], {1: torch.zeros(1, dtype=torch.float16)})
    assert isinstance(f, list) and f[-1] is None and 1 in f.taps16


def test_track_windowed_host_state_machine_matches_oracle_labels():
    """Label state machine {0,1,2} and valid masks of the sliding-window tracker (sparse_heads.py:306-335),
    evaluated on the host with the same tensor expressions the device driver uses."""
    q0 = torch.tensor([[[0.5, 1.0, 1.0], [9.5, 2.0, 2.0], [20.5, 3.0, 3.0]]])
    cur = q0.clone()
    cur[0, 0] = torch.tensor([12.5, 5.0, 5.0])  # re-queried by a previous window
    s, Tw = 8, 16
    ar = torch.arange(Tw)
    valid_t = ((ar.view(1, 1, Tw) + s + 0.5 - cur[:, :, 0:1]) >= 0)
    valid = valid_t.sum(-1) > 0
    lab = torch.where(valid, torch.ones(1, 3), torch.zeros(1, 3))
    same = (cur == q0).sum(dim=-1) > 0
    lab = torch.where(same, torch.ones_like(lab), lab)
    lab = torch.where(torch.logical_and(valid, ~same), torch.full_like(lab, 2), lab)
    assert lab.tolist() == [[2.0, 1.0, 1.0]]
    assert valid_t[0, 0].tolist() == [False] * 4 + [True] * 12   # frames before t=12.5 are never written
    assert valid_t[0, 2].tolist() == [False] * 12 + [True] * 4


def test_preprocess_plan_matches_dataset_restatement():
    """Shapes / crop origin of the fused GPU preprocessing (l4p_b200.data.plan) vs the CPU restatement of the reference's
    dataset pipeline (oracle/preprocess_oracle.py), including temporal mirror padding and the single-frame case."""
    import torch
    from l4p_b200.data import plan
    from oracle.preprocess_oracle import mirror_and_pad, preprocess
    x = torch.arange(5.0).view(1, 5, 1, 1)
    assert mirror_and_pad(x).flatten().tolist() == [0, 1, 2, 3, 4, 3, 2, 1, 0]
    for (T0, H0, W0, resize, crop) in [(10, 60, 80, (256, 320), None), (1, 50, 70, (240, 300), (16, 224, 224)),
                                       (40, 224, 224, (224, 224), None), (5, 300, 300, None, (24, 224, 224))]:
        frames = torch.zeros(T0, H0, W0, 3, dtype=torch.uint8)
        ref = preprocess(frames, resize, crop)
        pl = plan(T0, H0, W0, resize, crop)
        assert tuple(ref.shape) == (1, 3, pl["To"], pl["Hc"], pl["Wc"])
        assert pl["T_pad"] >= pl["To"] and pl["t0"] == 0
    try:
        plan(16, 100, 100, None, (16, 224, 224))
        assert False, "crop larger than the frame must fail like the reference's assert"
    except AssertionError:
        pass
