"""L4P model = shared video encoder + task heads. Drop-in mirror of l4p/models/l4p_videomae.py
(VideoMAEEncoder :17-122 re-exported from .videomae, L4P_VideoMAE :125-330): same constructor arguments,
state-dict keys (`video_encoder.*`, `task_heads.<task>.*`), `forward(data, tasks)`, `forward_single_window`,
`encode_features`, assertion messages and output keys.

B200-first difference: the reference encodes the sliding windows one after the other (:278-293); here all
windows of the clip (and all clips of the batch) are stacked and go through the encoder kernels as one batch,
in chunks of `max_windows_per_pass`, and the per-window lists the heads expect are views of that batch.
"""
from __future__ import annotations

from functools import partial
from typing import Any, Dict, List, Optional, Tuple

import torch

from .task_heads.dense_heads import joint_windowed_estimation
from .videomae import FeatureList, VideoMAEEncoder

__all__ = ["VideoMAEEncoder", "L4P_VideoMAE"]


class L4P_VideoMAE(torch.nn.Module):
    def __init__(self, task_heads: torch.nn.ModuleDict, video_encoder_ckpt_path: Optional[str] = None,
                 window_size: Tuple[int, int, int] = (16, 224, 224), window_stride_T: int = 8,
                 freeze_video_encoder: bool = False, freeze_heads: Optional[List[str]] = None,
                 unfreeze_blocks: Optional[List[int]] = None, always_use_windowed_version: bool = False,
                 joint_alignment: bool = False, cam_emb_placed_at_enc: Optional[str] = None, cam_emb_type: str = "add",
                 compute_dtype: torch.dtype = torch.float16, max_windows_per_pass: int = 8, parallel_heads: bool = True,
                 device=None) -> None:
        super().__init__()
        # Same hyper-parameters as the reference (l4p_videomae.py:163-186): ViT-giant, patch 14, tubelet 2.
        self.video_encoder = VideoMAEEncoder(
            img_size=224, patch_size=14, in_chans=3, num_classes=0, embed_dim=1408, depth=40, num_heads=16,
            mlp_ratio=48 / 11, qkv_bias=True, qk_scale=None, drop_rate=0, attn_drop_rate=0, drop_path_rate=0,
            norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), init_values=0.0, tubelet_size=2,
            use_learnable_pos_emb=False, with_cp=False, all_frames=16, cos_attn=False,
            cam_emb_placed_at=cam_emb_placed_at_enc, cam_emb_type=cam_emb_type, compute_dtype=compute_dtype,
            device=device)
        if video_encoder_ckpt_path is not None:
            print(f"Loading video model: {video_encoder_ckpt_path}")
            ckpt = torch.load(video_encoder_ckpt_path, map_location="cpu", weights_only=True)
            self.video_encoder.load_state_dict(ckpt, strict=False)
            print("Successfully loaded video model")
        self.task_heads = task_heads
        self.window_size = window_size
        self.window_stride_T = window_stride_T
        self.always_use_windowed_version = always_use_windowed_version
        self.joint_alignment = joint_alignment
        self.max_windows_per_pass = max_windows_per_pass
        # run independent task heads on concurrent CUDA streams (L4P_SERIAL_HEADS=1 turns it off for per-op profiling)
        self.parallel_heads = parallel_heads and __import__("os").environ.get("L4P_SERIAL_HEADS", "0") != "1"
        self._streams: Dict[Any, List[torch.cuda.Stream]] = {}
        # multi-GPU long-video mode (cfg 4): set by enable_window_sharding(); every rank encodes and
        # decodes only its contiguous block of windows, one all-gather per head assembles the per-window outputs and
        # the (cheap, sequential) stitching / alignment chain runs redundantly on every rank
        self.window_shard_group: Any = None
        self.shard_windows: bool = False
        self.set_compute_dtype(compute_dtype)
        # freeze_* / unfreeze_blocks are training-only knobs: accepted for config compatibility, parameters of
        # this inference-only implementation never require grad.

    def set_compute_dtype(self, dtype: torch.dtype) -> None:
        assert dtype in (torch.float16, torch.bfloat16)
        self.compute_dtype = dtype
        self.video_encoder.compute_dtype = dtype
        self.video_encoder.invalidate()
        for head in self.task_heads.values():
            if hasattr(head, "compute_dtype"):
                head.compute_dtype = dtype

    def enable_window_sharding(self, enabled: bool = True, group: Any = None) -> None:
        """Shard the windows of a long video across the ranks of `group` (default: the world group). Dense tasks only
        (depth / flow / dyn-mask / camray incl. joint alignment): the track head's window memory is sequential."""
        self.shard_windows = bool(enabled)
        self.window_shard_group = group

    def enable_query_sharding(self, enabled: bool = True, group: Any = None) -> None:
        """Shard the track queries across the ranks of `group`: every rank encodes the clip, tracks its slice of the queries
        and one all-gather returns all tracks everywhere (the demo's 625-query grids are 5 chunks of 128)."""
        self.task_heads["track_2d"].enable_query_sharding(enabled, group)

    def encode_features(self, data: Dict[str, Any]):
        """Generates video encoder features for a single window (l4p_videomae.py:222-232)."""
        return self.video_encoder(data["rgb_b3thw"])

    def forward_single_window(self, data: Dict[str, Any], tasks: List[str]) -> Dict[str, Any]:
        """l4p_videomae.py:234-254."""
        enc_features_bpc_list = self.video_encoder(data["rgb_b3thw"])
        out = {"enc_features_bpc_list": enc_features_bpc_list}
        for task in tasks:
            out.update(self.task_heads[task](enc_features_bpc_list=enc_features_bpc_list, **data))
        return out

    def _encode_windows(self, rgb: torch.Tensor, starts: List[int]):
        """All windows of all clips as one encoder batch (window-major: index = w*B + b)."""
        B = rgb.shape[0]
        Tw = self.window_size[0]
        wins = torch.cat([rgb[:, :, s:s + Tw] for s in starts], dim=0)  # [nW*B, 3, Tw, H, W]
        n = wins.shape[0]
        chunks: List[FeatureList] = []
        for i in range(0, n, self.max_windows_per_pass):
            chunks.append(self.video_encoder(wins[i:i + self.max_windows_per_pass]))
        if len(chunks) == 1:
            batched = chunks[0]
        else:
            L = len(chunks[0])
            feats = [None if chunks[0][i] is None else torch.cat([c[i] for c in chunks], dim=0) for i in range(L)]
            taps = {k: torch.cat([c.taps16[k] for c in chunks], dim=0) for k in chunks[0].taps16}
            batched = FeatureList(feats, taps)
        per_window = []
        for w in range(len(starts)):
            sl = slice(w * B, (w + 1) * B)
            per_window.append([None if f is None else f[sl] for f in batched])
        return batched, per_window

    group_dense_heads = __import__("os").environ.get("L4P_DPT_GROUP", "1") == "1"

    def _dense_group(self, tasks, batched, shard):
        """Heads of `tasks` whose DPT decoders can run as one group on the batched encoder result (>= 2 of them), else []."""
        if not self.group_dense_heads or batched is None or shard is not None:
            return []
        from .task_heads.dense_heads import VideoMAEFlowDPTHead
        rows = next(f for f in batched if f is not None).shape[0]
        if self.max_windows_per_pass and rows > self.max_windows_per_pass:
            return []
        by_sig = {}
        for t in tasks:
            h = self.task_heads[t] if t in self.task_heads else None
            if isinstance(h, VideoMAEFlowDPTHead) and type(h)._run_dpt is VideoMAEFlowDPTHead._run_dpt and hasattr(h.task_head, "dpt"):
                sig = (h.task_head.dpt.group_signature(), tuple(h.hooks_idx), h.compute_dtype)
                by_sig.setdefault(sig, []).append(h)
        best = max(by_sig.values(), key=len, default=[])
        return best if len(best) >= 2 else []

    def _run_jobs(self, jobs, device):
        if not self.parallel_heads or len(jobs) < 2 or device.type != "cuda":
            return [j() for j in jobs]
        main = torch.cuda.current_stream(device)
        streams = self._streams.setdefault(device, [])
        # (a higher scheduling priority for the track head's stream - a long chain of small token-side kernels between its
        # big ones - was measured in round 2: no effect on the captured step, 24.4 vs 24.8 ms)
        while len(streams) < len(jobs):
            streams.append(torch.cuda.Stream(device=device))
        start = torch.cuda.Event()
        start.record(main)
        results, done = [], []
        for job, st in zip(jobs, streams):
            st.wait_event(start)
            with torch.cuda.stream(st):
                res = job()
                if not torch.cuda.is_current_stream_capturing():   # (a graph's private pool never recycles mid-replay)
                    for v in res.values():
                        if torch.is_tensor(v):
                            v.record_stream(main)  # the caller consumes the outputs on its own stream
                e = torch.cuda.Event()
                e.record(st)
            results.append(res)
            done.append(e)
        for e in done:
            main.wait_event(e)
        return results

    def forward(self, data: Dict[str, Any], tasks: List[str]) -> Dict[str, Any]:
        """Main forward pass for both single and multi-window inference (l4p_videomae.py:256-330)."""
        B, _, T, H, W = data["rgb_b3thw"].shape
        assert H == self.window_size[1] and W == self.window_size[2], "Supports only fixed spatial size"
        if (not self.always_use_windowed_version) and (T == self.window_size[0]):
            return self.forward_single_window(data, tasks)
        assert T % self.window_stride_T == 0, "Temporal window needs to be a multiple of window stride, for now!"
        time_strides = torch.arange(0, T - self.window_size[0] + 1, self.window_stride_T)

        starts = [int(s) for s in time_strides]
        shard = None
        if self.shard_windows:
            import torch.distributed as dist
            from ..parallel import WindowShard
            assert dist.is_initialized(), "window sharding needs an initialised torch.distributed process group"
            assert "track_2d" not in tasks, "the track head's window memory is sequential: not window-shardable (SURVEY §8e)"
            assert B == 1, "window sharding: one video at a time"
            shard = WindowShard.for_rank(len(starts), self.window_shard_group)
            local_starts = starts[shard.start:shard.start + shard.count]
            if shard.count > 0:
                batched, enc_features_bpc_2dlist = self._encode_windows(data["rgb_b3thw"], local_starts)
            else:
                batched, enc_features_bpc_2dlist = None, []
        else:
            batched, enc_features_bpc_2dlist = self._encode_windows(data["rgb_b3thw"], starts)
        out: Dict[str, Any] = {"enc_features_bpc_2dlist": enc_features_bpc_2dlist}

        # The heads only read the encoder features: they are independent jobs (the reference runs them one after the
        # other, l4p_videomae.py:299-328). Each job runs on its own CUDA stream so that the low-occupancy kernels of
        # one head (low-resolution pyramid levels, token-side GEMMs, small solves) overlap the big kernels of another.
        common = dict(enc_features_bpc_2dlist=enc_features_bpc_2dlist, time_strides=time_strides, _batched_windows=batched,
                      _max_windows_per_pass=self.max_windows_per_pass)
        if shard is not None:
            common["_window_shard"] = shard

        def head_job(task: str):
            head = self.task_heads[task]
            if task != "track_2d" or B == 1:
                return head.forward_windowed(**common, **data)
            # Batched clips (BASELINE.json configs[2]): the encoder and the dense heads run on all clips at once; the tracker
            # keeps the reference's per-clip semantics (sparse_heads.py:241 asserts B == 1), so it is called clip by clip
            # on that clip's rows of the batched encoder result and the per-clip tracks are stacked along the batch axis.
            per_clip = []
            for b in range(B):
                feats_b = [[None if f is None else f[b:b + 1] for f in win] for win in enc_features_bpc_2dlist]
                data_b = {k: (v[b:b + 1] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B else v)
                          for k, v in data.items()}
                per_clip.append(head.forward_windowed(enc_features_bpc_2dlist=feats_b, time_strides=time_strides, **data_b))
            return {k: torch.cat([o[k] for o in per_clip], dim=0) for k in per_clip[0]}

        jobs = []
        # Dense heads whose DPT decoders share one architecture and read the same taps (flow / depth / motion mask in the shipped
        # config) run their decoders as ONE grouped launch sequence (dpt.forward_grouped) in a job of its own, enqueued first;
        # the heads' own jobs pick the result up behind a CUDA event (VideoMAEFlowDPTHead._run_dpt). Only when the whole batch
        # of windows is decoded in a single pass (no window sharding, rows <= max_windows_per_pass).
        group = self._dense_group(tasks, batched, shard)
        if group:
            def dense_group_job(group=group):
                from .task_heads.dense_heads import _taps16
                from .task_heads.dpt import forward_grouped
                h0 = group[0]
                img_info = tuple(self.window_size)
                taps, rows = _taps16(batched, h0.hooks_idx, h0.compute_dtype)
                outs = forward_grouped([h.task_head.dpt for h in group], taps, rows, img_info, exp_outs=[h._exp_out for h in group])
                evt = None
                if outs[0].is_cuda:       # (the CPU stand-in of the host tests runs the jobs in order on one thread)
                    evt = torch.cuda.Event()
                    evt.record()
                for h, o in zip(group, outs):
                    h._dpt_pre = (id(batched), o, evt, img_info)
                return {}
            jobs.append(dense_group_job)
        joint_alignment_possible = "depth" in tasks and "camray" in tasks
        if self.joint_alignment and joint_alignment_possible:
            for task in ["track_2d", "dyn_mask", "flow_2d_backward"]:
                if task in tasks:
                    jobs.append(lambda task=task: head_job(task))
            jobs.append(lambda: joint_windowed_estimation(["depth", "camray"], self.task_heads, **common, **data))
        else:
            if self.joint_alignment:
                print("Joint alignment is not possible as depth or camray tasks are not present")
            for task in tasks:
                jobs.append(lambda task=task: head_job(task))
        try:
            for res in self._run_jobs(jobs, data["rgb_b3thw"].device):
                out.update(res)
        finally:
            for h in group:
                h._dpt_pre = None
        return out
