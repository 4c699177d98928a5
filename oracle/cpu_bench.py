"""TEST/BENCH INFRASTRUCTURE ONLY. Bounded CPU timing of the oracle port of the reference path
(`bench.py`'s cpu_baseline leg and `--impl reference` arm; see oracle/__init__.py).

A full all-heads 16x224x224 window is ~23.3 TFLOP, i.e. minutes of CPU time, so one "step"
times a bounded SAMPLE of the same workload and extrapolates by the exact repetition structure of the model:

    encoder   : 1 of the 40 identical ViT blocks on the full [1,2048,1408] token tensor         x 40
    dense DPT : one full depth-head DPT on a 4-frame slab (img_info=(4,224,224); every conv of the
                head is linear in the number of frames)                                          x 4 per head, 3 heads
    camray    : the full 16x16x16 ray-map DPT + rays->camera solve                                x 1
    track     : the full two-way transformer + mask decoder + read-outs for ONE query             x n_queries
"""
from __future__ import annotations

import time
from typing import Dict

import torch

from . import l4p_oracle as O


def sample(sd_block: Dict[str, torch.Tensor], sd_depth, sd_cam, sd_track, n_queries: int = 128, seed: int = 0) -> Dict[str, float]:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 2048, 1408, generator=g)
    t = {}
    with torch.no_grad():
        t0 = time.perf_counter()
        O.vit_block(sd_block, "", x, 16)
        t["block_s"] = time.perf_counter() - t0
        feats4 = [torch.randn(1, 512, 1408, generator=g) for _ in range(4)]
        t0 = time.perf_counter()
        O.dpt_forward(sd_depth, "task_head.dpt.", feats4, hooks=(0, 1, 2, 3), img_info=(4, 224, 224))
        t["dense_slab_s"] = time.perf_counter() - t0
        feats = [torch.randn(1, 2048, 1408, generator=g) for _ in range(4)]
        t0 = time.perf_counter()
        rays = O.dpt_forward(sd_cam, "task_head.dpt.", feats, hooks=(0, 1, 2, 3), actpost=O.CAMRAY_ACTPOST,
                             fusion=O.CAMRAY_FUSION, output_size=(16, 16, 16))
        try:
            O.rays_to_cameras_fixed_intrinsics(rays)
        except Exception:  # random-weight rays can defeat cv2's RANSAC; the solve is negligible either way
            pass
        t["camray_s"] = time.perf_counter() - t0
        q = torch.tensor([[[0.5, 100.5, 120.5]]])
        t0 = time.perf_counter()
        O.track_head_window(sd_track, "", x, q, torch.ones(1, 1), torch.zeros(1, 1, 1408), torch.zeros(1, 1))
        t["track_query_s"] = time.perf_counter() - t0
    t["window_s"] = 40 * t["block_s"] + 3 * 4 * t["dense_slab_s"] + t["camray_s"] + n_queries * t["track_query_s"]
    t["frames_per_s"] = 16.0 / t["window_s"]
    return t


SAMPLE_DESC = ("oracle port, fp32 eager: 1/40 ViT blocks x40 + one dense DPT head on a 4-frame slab x4 x3 heads + "
               "full camray head + track head for 1 query x128, extrapolated to one all-heads 16x224x224 window")
