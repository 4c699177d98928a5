"""CPU twin of tests/test_joint_gpu.py / tests/test_dist_gpu.py (host logic of SURVEY.md §8 rows a11 / e): the joint depth +
pose window-alignment chain on the consistent scene of tests/scene.py at a small size, with the device solver replaced by
the oracle's restatement (tests/emu.py) - against the ground truth, against the hand-written oracle chain, and window-sharded
over two gloo ranks (ragged: 3 + 1 windows) against the unsharded run."""
import socket

import torch
import torch.multiprocessing as mp

from tests import emu, scene as S
from tests.util import rel_l2

IMG = (4, 56, 56)
STARTS = [0, 2, 4, 6]
T = 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_chain_on_consistent_scene_vs_truth_and_oracle(monkeypatch):
    from l4p_b200.models.task_heads import dense_heads as D
    from oracle import l4p_oracle as O

    emu.install(monkeypatch)
    sc = S.make_scene(T, STARTS, *IMG, seed=3)
    dev = torch.device("cpu")
    out = D.joint_windowed_estimation(["depth", "camray"], S.stub_heads(sc, dev), S.window_feats(range(4), dev),
                                      time_strides=torch.tensor(STARTS), intrinsics_b44t=sc["intr"], img_info=IMG)
    d, p = out["depth_est_b1thw"], out["traj3d_est_b16t"]
    assert rel_l2(d, sc["depth"]) < 1e-3 and (p - sc["pose"].reshape(1, 16, T)).abs().max() < 2e-3
    ref_d, ref_p = S.oracle_chain(sc, O)
    assert rel_l2(d, ref_d) < 1e-3 and (p - ref_p).abs().max() < 2e-3
    # batched call pattern (all windows in one head call) gives the same result
    bat = D.joint_windowed_estimation(["depth", "camray"], S.stub_heads(sc, dev), S.window_feats(range(4), dev),
                                      time_strides=torch.tensor(STARTS), intrinsics_b44t=sc["intr"], img_info=IMG,
                                      _batched_windows=[torch.arange(4.0)])
    for k in out:
        assert torch.equal(out[k], bat[k]), k


def test_two_rank_sharded_chain_equals_unsharded(tmp_path):
    from tests.joint_worker import run_rank

    world = 2
    mp.spawn(run_rank, args=(world, _free_port(), "cpu", "gloo", IMG, STARTS, T, str(tmp_path), True), nprocs=world, join=True)
    shards = []
    for r in range(world):
        res = torch.load(tmp_path / f"joint_{r}.pt")
        shards.append(res["shard"])
        for k in ("depth_est_b1thw", "traj3d_est_b16t", "traj3d_intrinsics_est_b16t"):
            assert res[k] < 1e-6, (r, k, res[k])          # fp32 outputs packed as fp32: the gather is exact
        assert res["pose_vs_truth"] < 2e-3 and res["depth_vs_truth"] < 1e-3
    assert shards == [(0, 2), (2, 2)]
