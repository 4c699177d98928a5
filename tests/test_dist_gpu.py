"""Multi-rank equality ON THE DEVICE (SURVEY.md §4 "multi-GPU" row, §8e): two ranks run the window-sharded joint depth + pose
chain (CUDA kernels: point maps + `l4p_sim3_align` on every rank after ONE gather of the per-window outputs) on the consistent
scene of tests/scene.py; every rank must return what the unsharded run returns, and both must reproduce the scene.

With >= 2 GPUs: NCCL, one GPU per rank (the product configuration). On a single-GPU box: both ranks share cuda:0 and the
exchange runs over gloo (device tensors staged through the host by l4p_b200.parallel._all_gather_into) - same kernels, same
sharding logic, only the transport differs. bench.py repeats the check on real NCCL ranks at every N > 1
(`configs.cfg4_sharded_vs_unsharded_rel_l2`)."""
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
IMG = (16, 224, 224)
STARTS = [0, 8, 16]      # ragged over two ranks: 2 + 1 windows
T = 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, multi_gpu, out_dir):
    from tests.joint_worker import run_rank

    device = f"cuda:{rank}" if multi_gpu else "cuda:0"
    run_rank(rank, world, port, device, "nccl" if multi_gpu else "gloo", IMG, STARTS, T, out_dir, False)


def test_two_rank_sharded_joint_chain_equals_unsharded_on_device(tmp_path):
    world = 2
    multi_gpu = torch.cuda.device_count() >= 2
    mp.spawn(_entry, args=(world, _free_port(), multi_gpu, str(tmp_path)), nprocs=world, join=True)
    shards = []
    for r in range(world):
        res = torch.load(tmp_path / f"joint_{r}.pt")
        shards.append(res["shard"])
        print(f"rank {r} ({'nccl' if multi_gpu else 'gloo, shared GPU'}): sharded vs unsharded "
              + ", ".join(f"{k} {res[k]:.1e}" for k in ("depth_est_b1thw", "traj3d_est_b16t")) +
              f"; vs truth: pose {res['pose_vs_truth']:.1e} depth {res['depth_vs_truth']:.1e}")
        for k in ("depth_est_b1thw", "traj3d_est_b16t", "traj3d_intrinsics_est_b16t"):
            assert res[k] <= 1e-3, (r, k, res[k])
        assert res["pose_vs_truth"] < 2e-3 and res["depth_vs_truth"] < 1e-3
    assert shards == [(0, 2), (2, 1)]
