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


def _plan(**kw):
    """l4p_gemm_plan: the host-side launch decisions of l4p_gemm (no device access, fake non-null pointers)."""
    import ctypes as C
    from l4p_b200 import lib
    L = lib.load()
    d = lib.GemmDesc()
    fake = 0x10000
    d.a, d.w = fake, fake
    d.bf16, d.a_mode, d.store_mode = 0, lib.A_MATRIX, lib.STORE_ROWMAJOR
    for k, v in kw.items():
        setattr(d, k, v)
    if d.lda == 0:
        d.lda = d.K
    if d.ldw == 0:
        d.ldw = d.K
    if d.store_mode == lib.STORE_ROWMAJOR:
        if not (d.out_f32 or d.out_16 or d.out_16_relu):
            d.out_16 = fake
        d.ld_out = d.ld_out or d.N
        d.ld_res = d.ld_res or d.N
    out = (C.c_int * 6)()
    rc = L.l4p_gemm_plan(C.byref(d), out)
    assert rc == 0, L.l4p_last_error().decode()
    return dict(block_n=out[0], split_k=out[1], pair=out[2], stages=out[3], grid=out[4], threads=out[5])


def test_gemm_plan_tile_selection():
    """Host logic of the GEMM launcher: N-tile width from the UMMA cost model, CTA pairing,
    split-K for few-tile / long-K problems, grid sizing for 148 SMs."""
    from l4p_b200 import lib
    # encoder at one clip (M = 2048)
    fc1 = _plan(M=2048, N=6144, K=1408)
    assert fc1["block_n"] == 240 and fc1["pair"] == 1 and fc1["grid"] == 148 and fc1["threads"] == 384   # 3 rounds like 24 x 256, narrower
    proj = _plan(M=2048, N=1408, K=1408, out_f32=0x10000, res_f32=0x10000)
    assert proj["block_n"] == 160 and proj["pair"] == 1 and proj["grid"] == 144     # 72 pair tiles on 74 pairs, one round
    assert proj["stages"] == 3                                                       # 128-wide K stages: 2 x (16 + 10) KiB each
    assert _plan(M=16384, N=6144, K=1408)["block_n"] == 256                          # many rounds: the widest tile
    qkv = _plan(M=2048, N=4224, K=1408, store_mode=lib.STORE_QKV, q=0x10000, k=0x10000, vt=0x10000, heads=16, head_dim=88,
                head_dim_pad=96, tokens=2048)
    assert qkv["block_n"] in (240, 256) and qkv["pair"] == 1                        # ragged wide tiles (2 rounds), not 22 x 192 (3 rounds)
    # track head projections over 128 queries x 2048 video tokens: 3 x 240 instead of 4 x 176
    kproj = _plan(M=262144, N=704, K=1408)
    assert kproj["block_n"] == 240 and kproj["pair"] == 1 and kproj["split_k"] == 1
    # low-resolution DPT level: 256 output voxels, K = 27 * 1024 -> split-K when (and only when) a workspace is given
    conv = dict(M=256, N=256, K=27 * 1024, a_mode=lib.A_CONV3D, cB=1, cT=4, cH=8, cW=8, cCin=1024, kT=3, kH=3, kW=3, bT=2, bH=8, bW=8,
                lda=1024, ldw=27 * 1024)
    no_ws = _plan(**conv)
    assert no_ws["split_k"] == 1 and no_ws["pair"] == 0
    ws = _plan(**conv, splitk_ws=0x10000, splitk_ws_bytes=256 * 256 * 4)
    assert ws["split_k"] > 8 and ws["block_n"] == 256 and ws["grid"] <= 148 and ws["grid"] == 2 * ws["split_k"]
    # fused-dot modes run four epilogue warpgroups + the helper warp
    hyper = _plan(M=16384 * 4, N=704, K=352, store_mode=lib.STORE_HYPER, cB=4, cT=16, cH=32, cW=32, sT=1, sH=2, sW=2, ctCout=176,
                  out_f32=0x10000, w2=0x10000, c2=3, rows_per_group=16384)
    assert hyper["block_n"] == 176 and hyper["threads"] == 640
    # invalid descriptors are rejected with an error code, not a crash
    import ctypes as C
    d = lib.GemmDesc(); d.a = d.w = 0x10000; d.M, d.N, d.K = 128, 24, 64; d.lda = d.ldw = 64
    assert lib.load().l4p_gemm_plan(C.byref(d), (C.c_int * 6)()) != 0


def test_gemm_plan_k128_stages():
    """128-wide K stages (two k-blocks per operand and ring stage) are planned for matrix-mode problems with K % 64 == 0 when three
    such stages fit; K % 64 != 0, short K, convolutions and L4P_GEMM_K128=0 keep 64-wide stages (more, smaller stages)."""
    import subprocess
    import sys
    fc2 = _plan(M=2048, N=1408, K=6144, out_f32=0x10000, res_f32=0x10000)
    assert fc2["block_n"] == 160 and fc2["pair"] == 1 and fc2["stages"] == 3           # 2 x (16 + 10) KiB per stage
    fc1 = _plan(M=16384, N=6144, K=1408)
    assert fc1["block_n"] == 256 and fc1["stages"] == 5                                 # 256-wide pair tiles: 64-wide stages of 32 KiB
    ragged_k = _plan(M=2048, N=1408, K=264, out_f32=0x10000, res_f32=0x10000)
    assert ragged_k["stages"] >= 5                                                       # K % 64 != 0: plain 2-D boxes
    narrow = _plan(M=768, N=1408, K=1408)                                                # narrow-tile rule: one wave of 64-wide tiles
    assert narrow["block_n"] <= 96 and narrow["split_k"] == 1 and 3 <= narrow["stages"] <= 8
    code = ("import ctypes as C; from l4p_b200 import lib; d = lib.GemmDesc(); d.a = d.w = d.out_f32 = d.res_f32 = 0x10000;"
            "d.M, d.N, d.K, d.lda, d.ldw, d.ld_out, d.ld_res = 2048, 1408, 6144, 6144, 6144, 1408, 1408;"
            "d.a_mode, d.store_mode = lib.A_MATRIX, lib.STORE_ROWMAJOR; o = (C.c_int * 6)();"
            "assert lib.load().l4p_gemm_plan(C.byref(d), o) == 0; print(o[3])")
    import os
    env = dict(os.environ, L4P_GEMM_K128="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert out.returncode == 0, out.stderr
    assert int(out.stdout.strip()) >= 6                                                  # 64-wide stages of 26 KiB


def test_prepare_model_mirrors_reference_loader(tmp_path):
    """l4p/models/utils.py:15-60: yaml -> max_queries override -> strict state_dict load -> eval; precision selects the
    16-bit operand type; CPU / fp32 requests raise instead of falling back."""
    import pytest

    from l4p_b200.config import DEFAULT_CONFIG
    from l4p_b200.lib import L4PError
    from l4p_b200.models.utils import compute_dtype_for, prepare_model

    m = prepare_model(str(DEFAULT_CONFIG), None, max_queries=32, precision="bf16-mixed", device="meta")
    assert not m.training and m.l4p_model.compute_dtype == torch.bfloat16
    assert m.l4p_model.task_heads["track_2d"].max_queries == 32
    assert compute_dtype_for("16-mixed") == torch.float16
    # Lightning checkpoint layout: {"state_dict": ..., <trainer state>}; meta tensors keep the file small
    sd = {k: torch.empty(v.shape, dtype=v.dtype, device="meta") for k, v in m.state_dict().items()}
    ckpt = tmp_path / "model.ckpt"
    torch.save({"state_dict": sd, "epoch": 3, "global_step": 7}, ckpt)
    m2 = prepare_model(str(DEFAULT_CONFIG), str(ckpt), device="meta")
    assert list(m2.state_dict()) == list(sd) and m2.l4p_model.compute_dtype == torch.float16
    sd.pop("l4p_model.task_heads.depth.task_head.dpt.act_postprocess.3.1.bias")
    torch.save({"state_dict": sd}, ckpt)
    with pytest.raises(RuntimeError, match="Missing key"):      # strict_loading: true (configs/model.yaml:11)
        prepare_model(str(DEFAULT_CONFIG), str(ckpt), device="meta")
    with pytest.raises(L4PError):
        prepare_model(str(DEFAULT_CONFIG), None, precision="32-true", device="meta")
    with pytest.raises(L4PError):
        prepare_model(str(DEFAULT_CONFIG), None, accelerator="cpu", device="meta")


def test_forward_argument_errors_match_reference_asserts():
    """Shape errors surface as the reference's assertions before any device work (l4p_videomae.py:260, :267-269)."""
    import pytest

    from l4p_b200.models.l4p_videomae import L4P_VideoMAE

    model = L4P_VideoMAE(torch.nn.ModuleDict(), always_use_windowed_version=True, device="meta")
    with pytest.raises(AssertionError, match="fixed spatial size"):
        model.forward(dict(rgb_b3thw=torch.zeros(1, 3, 16, 200, 224, device="meta")), [])
    with pytest.raises(AssertionError, match="multiple of window stride"):
        model.forward(dict(rgb_b3thw=torch.zeros(1, 3, 20, 224, 224, device="meta")), [])


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under l4p_b200/ may import it, and bench.py only in its CPU-baseline leg."""
    import re
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference", re.M)
    for f in (root / "l4p_b200").rglob("*.py"):
        assert not pat.search(f.read_text()), f"{f} references the oracle / the reference tree"
    bench = (root / "bench.py").read_text()
    uses = [m.start() for m in re.finditer(r"^\s*(from|import)\s+oracle\b", bench, re.M)]
    # only the comparator legs (`_port_sample`, `reference_cpu`: the cpu_baseline block and the --impl reference arm)
    lo, hi = bench.index("def _port_sample"), bench.index("def run_reference")
    assert uses and all(lo < u < hi for u in uses)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from l4p_b200 import lib as L

    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", tmp_path / "libl4p_b200.so")
    import pytest

    with pytest.raises(L.L4PError, match="no CPU fallback"):
        L.load()


def test_attention_synchronisation_protocol_model():
    """tools/att_protocol_sim.py: discrete-event model of the attention kernels' mbarrier / TMA / tcgen05.commit protocol
    (production kernel, S-first and P-alias build variants, the CTA-pair kernel and its P-alias variant). Random
    interleavings must terminate with every buffer consumed at the right version and no parity aliasing; deliberately
    broken protocols must be rejected."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("att_protocol_sim", Path(__file__).resolve().parents[1] / "tools" / "att_protocol_sim.py")
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    caught = sim.self_test(seeds=20)
    assert all(v > 0 for v in caught.values()), caught
    sim.check_all(seeds=25, verbose=False)


def test_malformed_conv_descriptor_is_an_argument_error():
    """A zeroed / malformed conv descriptor must come back as an error code from the C ABI (argument validation before any
    use of the box as a divisor), never as a SIGFPE inside the library (round-1 advisor finding on gemm.cu)."""
    import ctypes as C

    from l4p_b200 import lib

    L = lib.load()
    out = (C.c_int * 6)()
    d = lib.GemmDesc()
    d.a, d.w, d.out_16 = 0x10000, 0x10000, 0x10000
    d.M, d.N, d.K, d.lda, d.ldw, d.ld_out = 256, 256, 27 * 64, 64, 27 * 64, 256
    d.a_mode, d.store_mode = lib.A_CONV3D, lib.STORE_ROWMAJOR          # every conv field left at zero
    assert L.l4p_gemm_plan(C.byref(d), out) != 0
    assert b"conv box" in L.l4p_last_error()
    d.bT, d.bH, d.bW = 2, 8, 7                                           # 112 voxels: not a 128-row tile
    assert L.l4p_gemm_plan(C.byref(d), out) != 0
    d.bW = 8                                                             # box ok, geometry still zero
    assert L.l4p_gemm_plan(C.byref(d), out) != 0
    assert b"conv geometry" in L.l4p_last_error()
    d.a_mode = 7
    assert L.l4p_gemm_plan(C.byref(d), out) != 0
