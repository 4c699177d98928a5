"""TEST/BENCH INFRASTRUCTURE ONLY. Bounded CPU timing of the oracle port of the reference path
(`bench.py`'s cpu_baseline leg and `--impl reference` arm; see oracle/__init__.py).

A full all-heads 16x224x224 window is ~23.3 TFLOP, i.e. minutes of CPU time, so one "step"
times a bounded SAMPLE of the same workload and extrapolates by the exact repetition structure of the model:

    encoder   : BLOCKS of the 40 identical ViT blocks, chained, on the full [1,2048,1408] tokens  x 40 / BLOCKS
    dense DPT : one full depth-head DPT on a SLAB_FRAMES-frame slab (img_info=(SLAB_FRAMES,224,224);
                every conv of the head is linear in the number of frames)                        x 16 / SLAB_FRAMES per head, 3 heads
    camray    : the full 16x16x16 ray-map DPT + rays->camera solve                                x 1
    track     : the full two-way transformer + mask decoder + read-outs for QUERIES queries       x n_queries / QUERIES

Defaults (BLOCKS=8, SLAB_FRAMES=16, QUERIES=8) are ~5-6 s of wall time on 8-16 host cores (about 1/5 of the window).
"""
from __future__ import annotations

import time
from typing import Dict

import torch

from . import l4p_oracle as O


BLOCKS, SLAB_FRAMES, QUERIES = 8, 16, 8


def sample(sd_block: Dict[str, torch.Tensor], sd_depth, sd_cam, sd_track, n_queries: int = 128, seed: int = 0,
           blocks: int = BLOCKS, slab_frames: int = SLAB_FRAMES, queries: int = QUERIES) -> Dict[str, float]:
    assert 40 % blocks == 0 and 16 % slab_frames == 0 and slab_frames % 2 == 0 and n_queries % queries == 0
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 2048, 1408, generator=g)
    t = {}
    with torch.no_grad():
        t0 = time.perf_counter()
        y = x
        for _ in range(blocks):
            y = O.vit_block(sd_block, "", y, 16)
        t["block_s"] = (time.perf_counter() - t0) / blocks
        slab_tokens = (slab_frames // 2) * 256
        feats4 = [torch.randn(1, slab_tokens, 1408, generator=g) for _ in range(4)]
        t0 = time.perf_counter()
        O.dpt_forward(sd_depth, "task_head.dpt.", feats4, hooks=(0, 1, 2, 3), img_info=(slab_frames, 224, 224))
        t["dense_head_s"] = (time.perf_counter() - t0) * (16 // slab_frames)
        feats = [torch.randn(1, 2048, 1408, generator=g) for _ in range(4)]
        t0 = time.perf_counter()
        rays = O.dpt_forward(sd_cam, "task_head.dpt.", feats, hooks=(0, 1, 2, 3), actpost=O.CAMRAY_ACTPOST,
                             fusion=O.CAMRAY_FUSION, output_size=(16, 16, 16))
        try:
            O.rays_to_cameras_fixed_intrinsics(rays)
        except Exception:  # random-weight rays can defeat cv2's RANSAC; the solve is negligible either way
            pass
        t["camray_s"] = time.perf_counter() - t0
        q = torch.tensor([[[0.5, 20.5 + 25.0 * i, 120.5] for i in range(queries)]])
        t0 = time.perf_counter()
        O.track_head_window(sd_track, "", x, q, torch.ones(1, queries), torch.zeros(1, queries, 1408), torch.zeros(1, queries))
        t["track_query_s"] = (time.perf_counter() - t0) / queries
    t["sample_s"] = blocks * t["block_s"] + t["dense_head_s"] * slab_frames / 16 + t["camray_s"] + queries * t["track_query_s"]
    t["window_s"] = 40 * t["block_s"] + 3 * t["dense_head_s"] + t["camray_s"] + n_queries * t["track_query_s"]
    t["frames_per_s"] = 16.0 / t["window_s"]
    return t


SAMPLE_DESC = (f"oracle port, fp32 eager, all host threads: {BLOCKS} of the 40 ViT blocks (x{40 // BLOCKS}) + one full dense DPT head "
               f"on {SLAB_FRAMES} frames (x3 heads) + the full camray head and pose solve + the track head for {QUERIES} of the "
               f"128 queries (x{128 // QUERIES}), extrapolated by repetition count to one all-heads 16x224x224 window")
