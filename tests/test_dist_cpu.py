"""N>1 path on CPU: world_size-2 gloo run of the clip sharding + single all-gather used by bench.py / cfg 3.
Sharding must not change per-clip results: the gathered buffer equals the single-process buffer."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _clip_result(clip_id: int) -> torch.Tensor:
    """Stand-in for one clip's packed head outputs (deterministic function of the clip id)."""
    g = torch.Generator().manual_seed(1000 + clip_id)
    return torch.rand(257, generator=g)


def _worker(rank, world, port, clips, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [c for c in range(clips) if c % world == rank]           # clip i -> rank i mod N
    packed = torch.cat([_clip_result(c) for c in mine])
    gathered = torch.empty(world * packed.numel())
    dist.all_gather_into_tensor(gathered, packed)                    # the path's only exchange step
    if rank == 0:
        torch.save(gathered, os.path.join(out_dir, "gathered.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    world, clips = 2, 6
    mp.spawn(_worker, args=(world, _free_port(), clips, str(tmp_path)), nprocs=world, join=True)
    gathered = torch.load(tmp_path / "gathered.pt").view(world, clips // world, -1)
    for c in range(clips):
        r, slot = c % world, c // world
        assert torch.equal(gathered[r, slot], _clip_result(c)), f"clip {c} changed under sharding"


# ---------------------------------------------------------------------------------------------------------------
# cfg 4: windows of one long video sharded in contiguous blocks, all per-window outputs in ONE all-gather
# (l4p_b200.parallel.gather_window_outputs, the function the model calls under enable_window_sharding()).
# ---------------------------------------------------------------------------------------------------------------
def _window_result(w: int):
    g = torch.Generator().manual_seed(2000 + w)
    return torch.rand(1, 1, 4, 6, 5, generator=g), torch.rand(1, 6, 4, 3, 3, generator=g).double()


def _window_worker(rank, world, port, n_windows, out_dir):
    from l4p_b200.parallel import WindowShard, contiguous_partition, gather_window_outputs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = WindowShard.for_rank(n_windows)
    assert (shard.start, shard.count) == contiguous_partition(n_windows, world)[rank]
    mine = [_window_result(w) for w in range(shard.start, shard.start + shard.count)]
    depth, rays = gather_window_outputs([[m[0] for m in mine], [m[1] for m in mine]], shard)
    torch.save((depth, rays), os.path.join(out_dir, f"gathered_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_window_shard(tmp_path):
    from l4p_b200.parallel import contiguous_partition, round_robin_partition
    assert contiguous_partition(5, 2) == [(0, 3), (3, 2)]
    assert contiguous_partition(3, 4) == [(0, 1), (1, 1), (2, 1), (3, 0)]
    assert round_robin_partition(5, 2) == [[0, 2, 4], [1, 3]]
    world, n_windows = 2, 5   # ragged: 3 + 2 windows
    mp.spawn(_window_worker, args=(world, _free_port(), n_windows, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):   # every rank ends up with the outputs of ALL windows, in window order, dtype preserved
        depth, rays = torch.load(tmp_path / f"gathered_{rank}.pt")
        assert len(depth) == n_windows and len(rays) == n_windows
        for w in range(n_windows):
            d, r = _window_result(w)
            assert torch.equal(depth[w], d) and depth[w].dtype == torch.float32
            assert torch.allclose(rays[w], r.float().double()) and rays[w].dtype == torch.float64


# ---------------------------------------------------------------------------------------------------------------
# cfg 4 end to end on CPU: the tiny golden model (tests/test_host_emulated.py) with `enable_window_sharding()` on two
# gloo ranks -- each rank encodes / decodes only its windows, one all-gather per head, stitching + alignment on every
# rank -- must return on EVERY rank what the unsharded model returns. Kernel layer = tests/emu.py (test-only).
# ---------------------------------------------------------------------------------------------------------------
def _sharded_model_worker(rank, world, port, T, out_dir):
    import pytest

    from tests import emu
    from tests import test_host_emulated as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    mp_ = pytest.MonkeyPatch()
    emu.install(mp_)
    try:
        tasks = ["depth", "flow_2d_backward"]
        model = H._tiny_model(tasks)
        model.enable_window_sharding(True)
        rgb = H.rnd((1, 3, T, 56, 56), 16)
        intr = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T)
        out = model.forward(dict(rgb_b3thw=rgb, intrinsics_b44t=intr, img_info=H.IMG), tasks)
        torch.save({"depth": out["depth_est_b1thw"], "flow": out["flow_2d_backward_est_b2thw"],
                    "local_windows": len(out["enc_features_bpc_2dlist"]), "encoder_passes": emu.CALLS.get("patchify", 0)},
                   os.path.join(out_dir, f"sharded_{T}_{rank}.pt"))
        dist.barrier()
    finally:
        mp_.undo()
        dist.destroy_process_group()


def test_two_rank_window_sharded_model_matches_unsharded(tmp_path, monkeypatch):
    from tests import emu
    from tests import test_host_emulated as H
    from tests.util import rel_l2

    world = 2
    for T, local in ((8, [2, 1]), (4, [1, 0])):        # 3 windows -> 2 + 1; 1 window -> 1 + 0 (a rank with nothing to do)
        mp.spawn(_sharded_model_worker, args=(world, _free_port(), T, str(tmp_path)), nprocs=world, join=True)
        emu.install(monkeypatch)
        tasks = ["depth", "flow_2d_backward"]
        rgb = H.rnd((1, 3, T, 56, 56), 16)
        intr = torch.eye(4)[None, :, :, None].repeat(1, 1, 1, T)
        ref = H._tiny_model(tasks).forward(dict(rgb_b3thw=rgb, intrinsics_b44t=intr, img_info=H.IMG), tasks)
        for rank in range(world):
            got = torch.load(tmp_path / f"sharded_{T}_{rank}.pt")
            assert got["local_windows"] == local[rank] and got["encoder_passes"] == (1 if local[rank] else 0)
            # per-rank batches differ from the single-process batch -> fp32 summation order -> 16-bit rounding flips
            assert got["depth"].shape == ref["depth_est_b1thw"].shape
            assert rel_l2(got["depth"], ref["depth_est_b1thw"]) < 1e-3
            assert rel_l2(got["flow"], ref["flow_2d_backward_est_b2thw"]) < 1e-3
        a, b = (torch.load(tmp_path / f"sharded_{T}_{r}.pt") for r in range(world))
        # every rank ends with the same result: flow is pure data movement after the gather (bit-equal); depth goes through
        # the overlap least-squares solve, whose multi-threaded CPU reductions are not bit-reproducible across processes
        assert torch.equal(a["flow"], b["flow"]) and rel_l2(a["depth"], b["depth"]) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# Track queries sharded across ranks (queries are independent): `enable_query_sharding()` on the tiny
# golden tracker, two gloo ranks, 3 windows with the sliding-window memory -> every rank returns all tracks.
# ---------------------------------------------------------------------------------------------------------------
def _sharded_tracker_worker(rank, world, port, n_queries, out_dir):
    import pytest

    from tests import emu
    from tests import test_host_emulated as H
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    mp_ = pytest.MonkeyPatch()
    emu.install(mp_)
    try:
        enc = H._encoder()
        starts = torch.arange(0, 8 - 4 + 1, 2)
        f2d = H._windows(enc, H.rnd((1, 3, 8, 56, 56), 16), starts)
        trk = H._tracker(max_queries=2)
        trk.enable_query_sharding(True)
        q = _queries(n_queries)
        out = trk.forward_windowed(f2d, q, torch.ones(1, n_queries), time_strides=starts)
        torch.save({k: v for k, v in out.items()}, os.path.join(out_dir, f"tracks_{n_queries}_{rank}.pt"))
        dist.barrier()
    finally:
        mp_.undo()
        dist.destroy_process_group()


def _queries(n):
    g = torch.Generator().manual_seed(77)
    return torch.cat([torch.randint(0, 7, (1, n, 1), generator=g).float() + 0.5, torch.rand(1, n, 2, generator=g) * 50 + 3], -1)


def test_two_rank_query_sharded_tracker_matches_unsharded(tmp_path, monkeypatch):
    from tests import emu
    from tests import test_host_emulated as H

    world = 2
    emu.install(monkeypatch)
    enc = H._encoder()
    starts = torch.arange(0, 8 - 4 + 1, 2)
    f2d = H._windows(enc, H.rnd((1, 3, 8, 56, 56), 16), starts)
    for n in (5, 1):                                  # 3 + 2 queries; 1 + 0 (a rank without work)
        mp.spawn(_sharded_tracker_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
        ref = H._tracker().forward_windowed(f2d, _queries(n), torch.ones(1, n), time_strides=starts)
        outs = [torch.load(tmp_path / f"tracks_{n}_{r}.pt") for r in range(world)]
        for k, v in ref.items():
            for o in outs:
                assert o[k].shape == v.shape and o[k].dtype == v.dtype, k
                assert torch.equal(o[k] == 0, v == 0) and torch.equal(o[k] == -10, v == -10), f"{k}: written-frame mask"
                assert (o[k] - v).abs().max() < 2e-3, (k, (o[k] - v).abs().max())
            assert torch.equal(outs[0][k], outs[1][k])          # gathered result is identical on every rank
