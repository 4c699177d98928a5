"""Multi-GPU sharding of the hot path: one process per GPU, weights replicated, independent units
(clips, windows of a long video) partitioned across ranks, ONE all-gather of the packed per-unit outputs per step.
The sequential parts (alignment chains, the sliding-window track memory) run after the gather, redundantly on
every rank, so every rank ends up with the full result (like the reference's single-process output).

Only `torch.distributed` plumbing lives here (NCCL over NVLink on the GPU box, gloo in the CPU tests); no kernel
of the path has a collective inside."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def _all_gather_into(out: torch.Tensor, inp: torch.Tensor, group=None) -> None:
    """`dist.all_gather_into_tensor`; gloo has no CUDA all-gather, so under gloo (the CPU test backend, also used to put two
    test ranks on ONE GPU) device tensors are staged through the host. NCCL - the product path - gathers in place."""
    if inp.is_cuda and dist.get_backend(group) == "gloo":
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(host, inp.detach().cpu().contiguous(), group=group)
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, inp, group=group)


def contiguous_partition(n_units: int, world: int) -> List[Tuple[int, int]]:
    """(start, count) per rank: contiguous blocks of ceil(n/world) units, so that overlap neighbours of a long video
    are mostly local (cfg 4); trailing ranks may get fewer (or zero) units."""
    per = -(-n_units // world)
    return [(min(r * per, n_units), max(0, min(n_units, (r + 1) * per) - r * per)) for r in range(world)]


def round_robin_partition(n_units: int, world: int) -> List[List[int]]:
    """clip i -> rank i mod N (cfg 3)."""
    return [[i for i in range(n_units) if i % world == r] for r in range(world)]


@dataclass
class WindowShard:
    """This rank's slice of the windows of one long video."""
    start: int
    count: int
    n_windows: int
    group: Optional[Any] = None

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group)

    @property
    def per_rank(self) -> int:
        return -(-self.n_windows // self.world)

    @staticmethod
    def for_rank(n_windows: int, group=None) -> "WindowShard":
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        start, count = contiguous_partition(n_windows, world)[rank]
        return WindowShard(start, count, n_windows, group)


def gather_window_outputs(local: Sequence[Sequence[torch.Tensor]], shard: WindowShard) -> List[List[torch.Tensor]]:
    """local[k][w]: output k (e.g. depth, rays) of this rank's w-th window. Returns the same structure for ALL
    windows of the video, identical on every rank. All outputs of all local windows travel in ONE
    all_gather_into_tensor (packed fp32 buffer, padded to ceil(nW/world) windows per rank)."""
    per, world = shard.per_rank, shard.world
    # shapes come from whatever rank has at least one window; every rank has the same per-window shapes by construction
    ref = [o[0] for o in local] if shard.count > 0 else None
    meta = [None]
    if ref is not None:
        meta = [[(tuple(t.shape), t.dtype) for t in ref]]
    # every rank computes the same partition: when ALL ranks hold at least one window, each knows the per-window shapes from
    # its own outputs and the host-side object exchange (a blocking, pickled collective, not capturable in a CUDA graph) is
    # skipped; it is only needed when some rank has no window at all (more ranks than windows)
    all_have = all(cnt > 0 for _, cnt in contiguous_partition(shard.n_windows, world))
    if world > 1 and not all_have:
        metas = [None] * world
        dist.all_gather_object(metas, meta[0], group=shard.group)
        meta0 = next(m for m in metas if m is not None)
    else:
        meta0 = meta[0]
    sizes = [int(torch.tensor(s).prod()) if len(s) else 1 for s, _ in meta0]
    unit = sum(sizes)
    dev = ref[0].device if ref is not None else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    packed = torch.zeros(per, unit, device=dev, dtype=torch.float32)
    for w in range(shard.count):
        off = 0
        for k, n in enumerate(sizes):
            packed[w, off:off + n] = local[k][w].reshape(-1).float()
            off += n
    gathered = torch.empty(world * per, unit, device=dev, dtype=torch.float32)
    _all_gather_into(gathered, packed, shard.group)
    parts = contiguous_partition(shard.n_windows, world)
    out: List[List[torch.Tensor]] = [[] for _ in sizes]
    for r, (_, cnt) in enumerate(parts):
        for w in range(cnt):
            row = gathered[r * per + w]
            off = 0
            for k, (n, (shape, dtype)) in enumerate(zip(sizes, meta0)):
                out[k].append(row[off:off + n].reshape(shape).to(dtype))
                off += n
    return out


def gather_query_outputs(local: Sequence[Optional[torch.Tensor]], n_queries: int, group=None) -> List[torch.Tensor]:
    """Track queries are independent (sparse_heads.py:181-211): rank r tracks the contiguous slice
    `contiguous_partition(n_queries, world)[r]` of the queries. local[k]: this rank's output k, shape [1, n_r, ...] (None
    when the rank has no query). Returns every output for ALL queries ([1, n_queries, ...], query order preserved),
    identical on every rank; all outputs travel in ONE all_gather_into_tensor."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = contiguous_partition(n_queries, world)
    per = -(-n_queries // world)
    count = parts[rank][1]
    meta = [(tuple(t.shape[2:]), t.dtype) for t in local] if count > 0 else None
    metas: List[Any] = [None] * world
    dist.all_gather_object(metas, meta, group=group)
    meta0 = next(m for m in metas if m is not None)
    sizes = [int(torch.tensor(s).prod()) if len(s) else 1 for s, _ in meta0]
    unit = sum(sizes)
    if count > 0:
        dev = local[0].device
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    packed = torch.zeros(per, unit, device=dev, dtype=torch.float32)
    if count > 0:
        packed[:count] = torch.cat([t[0].reshape(count, -1).float() for t in local], dim=1)
    gathered = torch.empty(world * per, unit, device=dev, dtype=torch.float32)
    _all_gather_into(gathered, packed, group)
    rows = torch.cat([gathered[r * per:r * per + cnt] for r, (_, cnt) in enumerate(parts)], dim=0)   # [n_queries, unit]
    out, off = [], 0
    for n, (shape, dtype) in zip(sizes, meta0):
        out.append(rows[:, off:off + n].reshape(1, n_queries, *shape).to(dtype))
        off += n
    return out


def gather_clip_outputs(packed: torch.Tensor, group=None) -> torch.Tensor:
    """cfg 3: every rank contributes the packed head outputs of its clips ([clips_per_rank, unit]); returns
    [world, clips_per_rank, unit] on every rank (clip i sits at [i % world, i // world])."""
    world = dist.get_world_size(group)
    out = torch.empty(world, *packed.shape, device=packed.device, dtype=packed.dtype)
    _all_gather_into(out.view(world * packed.shape[0], *packed.shape[1:]), packed.contiguous(), group)
    return out
