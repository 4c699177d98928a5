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
