"""Test infrastructure: one rank of a multi-process run of the joint depth + pose window-alignment chain on the consistent
scene of tests/scene.py, window-sharded (`_window_shard` + `gather_window_outputs`, the path `enable_window_sharding()`
takes) AND unsharded in the same process; every rank must end up with the unsharded result.

Used by tests/test_joint_cpu.py (gloo, CPU tensors, kernels replaced by tests/emu.py) and tests/test_dist_gpu.py (CUDA
kernels; NCCL with one GPU per rank, or gloo with both ranks on cuda:0 when the box has a single GPU)."""
import os

import torch
import torch.distributed as dist


def run_rank(rank: int, world: int, port: int, device: str, backend: str, img, starts, T, out_dir: str, emulate: bool):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    mp_ctx = None
    if emulate:
        import pytest

        from tests import emu

        mp_ctx = pytest.MonkeyPatch()
        emu.install(mp_ctx)
    dev = torch.device(device)
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    kw = dict(device_id=dev) if backend == "nccl" else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    from l4p_b200.models.task_heads import dense_heads as D
    from l4p_b200.parallel import WindowShard
    from tests import scene as S

    sc = S.make_scene(T, starts, img[0], img[1], img[2], seed=3)
    heads = S.stub_heads(sc, dev)
    common = dict(time_strides=torch.tensor(starts), intrinsics_b44t=sc["intr"].to(dev), img_info=tuple(img))
    feats = S.window_feats(range(len(starts)), dev)
    with torch.no_grad():
        base = D.joint_windowed_estimation(["depth", "camray"], heads, feats, **common)
        shard = WindowShard.for_rank(len(starts))
        ids = torch.tensor([float(i) for i in range(shard.start, shard.start + shard.count)], device=dev)
        sh = D.joint_windowed_estimation(["depth", "camray"], S.stub_heads(sc, dev), feats, _window_shard=shard,
                                         _batched_windows=[ids] if shard.count > 0 else None, **common)
    res = {"shard": (shard.start, shard.count)}
    for k in base:
        a, b = sh[k].float().cpu(), base[k].float().cpu()
        res[k] = float((a - b).norm() / (b.norm() + 1e-30))
    gt = sc["pose"].reshape(1, 16, T)
    res["pose_vs_truth"] = float((sh["traj3d_est_b16t"].cpu() - gt).abs().max())
    res["depth_vs_truth"] = float((sh["depth_est_b1thw"].cpu() - sc["depth"]).norm() / sc["depth"].norm())
    torch.save(res, os.path.join(out_dir, f"joint_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()
    if mp_ctx is not None:
        mp_ctx.undo()
