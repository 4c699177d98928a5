#!/usr/bin/env python
"""Headline benchmark of the L4P inference hot path on B200 (contract: see the build brief / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline (BASELINE.json configs[1]): one step = one pass of the all-heads hot path (encoder + flow / depth / dyn-mask /
camray DPT heads + 128-query track head, the shipped config: windowed path + joint alignment) over ONE synthetic 16x224x224
clip per GPU through `L4PLitModule.predict_step`. Metric: frames/s = 16 * clips / step time. N > 1: one clip per GPU
(weak scaling) + ONE all-gather of the packed head outputs over NVLink.

The same process then times (block `configs` of the same JSON line, skipped with --skip-configs):
  cfg3  BASELINE.json configs[2]: 4 clips per GPU as ONE batch (encoder + DPT heads batched over the clips, tracker per clip)
  cfg4  BASELINE.json configs[3]: one long video (T=264 -> 32 windows, T=512 -> 63 windows), depth + camray with joint
        Sim(3) window alignment on the GPU; N > 1 shards the windows across ranks (`enable_window_sharding`), and a
        4-window clip is run sharded AND unsharded to report their difference.

`--impl reference` times the UNMODIFIED reference (baseline/_ref, see oracle/ref_arm.py) on the same full workload: fp32
eager on the host cores (the comparator), plus its eager fp16-autocast path on the B200 (informational).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TASKS = ["flow_2d_backward", "track_2d", "depth", "dyn_mask", "camray"]
NQ = 128
ATT_FLOPS_PER_BLOCK_WINDOW = 4 * 2048 * 2048 * 88 * 16  # 23 622 320 128 (unpadded d=88)
STEP_GFLOP = 23316.6   # one all-heads window in the reference's order of operations (cfg 2, 128 track queries)
WORKLOAD = ("single 16x224x224 clip per GPU, all heads (flow, depth, dyn-mask, camray pose, 128-query 2D/3D tracks), "
            "shipped config (windowed + joint alignment), BASELINE.json configs[1]")
REF_ARM_FILES = [Path("/tmp/l4p_reference_arm.json"), ROOT / "gpurun_out" / "reference_arm_last.json"]


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(tflops=j.get("bf16_tflops_sustained", j["bf16_tflops"]), hbm=j["hbm_gbs"], src="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def synth_batch(clips: int, T: int = 16, queries: bool = True):
    from tests.util import synth_intrinsics, synth_rgb

    batch = dict(rgb_b3thw=synth_rgb(clips, T, seed=0), intrinsics_b44t=synth_intrinsics(clips, T))
    if queries:
        xs = torch.linspace(7, 217, 16)
        ys = torch.linspace(14, 210, 8)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        q = torch.stack([torch.full_like(gx, 0.5), gx + 0.5, gy + 0.5], dim=-1).reshape(1, NQ, 3).repeat(clips, 1, 1)
        batch.update(track_2d_pointquerries_bn3=q, track_2d_pointlabels_bn=torch.ones(clips, NQ))
    return batch


OUT_KEYS = ["depth_est_b1thw", "flow_2d_backward_est_b2thw", "dyn_mask_est_b1thw", "traj3d_est_b16t",
            "traj3d_intrinsics_est_b16t", "track_2d_traj_est_bn2t", "track_2d_vis_est_bn1t", "track_2d_depth_est_bn1t"]


def pack_outputs(out, keys, clips):
    """[clips, unit] fp32: every head output of every clip, the buffer that is exchanged / copied to the host."""
    return torch.cat([out[k].reshape(clips, -1).float() for k in keys], dim=1)


# ------------------------------------------------------------------------------------------------ reference arm
def _port_sample():
    """Fallback comparator when the reference sources did not travel (no baseline/_ref): the oracle port on a bounded sample."""
    from l4p_b200 import weights
    from l4p_b200.models.task_heads.dense_heads import VideoMAEDepthDPTHead, VideoMAETraj3DDPTHead
    from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead
    from l4p_b200.models.videomae import Block
    from oracle import cpu_bench

    torch.set_num_threads(os.cpu_count() or 1)
    hooks = [0, 1, 2, 3]
    mods = dict(
        block=Block(1408, 16, 48 / 11, True, None, 1e-6, 0.0, device="meta"),
        depth=VideoMAEDepthDPTHead("depth", hooks_idx=hooks, device="meta"),
        cam=VideoMAETraj3DDPTHead("traj3d", hooks_idx=hooks, use_intrinsics=False, fixed_intrinsics=True, device="meta"),
        track=VideoMAETrack2DSamHead(estimate_vis=True, estimate_depth=True, prompt_using_features=True, attend_to_past=True,
                                     modify_pointlabels_for_windowing=True, estimation_directions=[1], depth_fn="exp",
                                     device="meta"))
    sds = {n: weights.synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=0) for n, m in mods.items()}
    s = cpu_bench.sample(sds["block"], sds["depth"], sds["cam"], sds["track"], NQ)
    return s, cpu_bench.SAMPLE_DESC


def reference_cpu(steps: int, warmup: int, budget_s: float, with_cuda: bool):
    """(cpu_baseline dict, extra dict). Real reference when its sources are on the box, else the port sample."""
    from oracle import ref_arm

    cores = os.cpu_count() or 1
    if ref_arm.root() is None:
        s, desc = _port_sample()
        return ({"value": s["frames_per_s"], "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc,
                 "same_config": False}, {"window_ms": 1e3 * s["window_s"], "steps_timed": 1})
    t0 = time.perf_counter()
    model = ref_arm.build_model(NQ)
    build_s = time.perf_counter() - t0
    batch = synth_batch(1)
    r = ref_arm.time_cpu(model, batch, TASKS, steps, warmup, budget_s)
    cb = {"value": 16.0 / r["mean_s"], "unit": "frames/s", "cores": r["threads"], "kind": "reference", "same_config": True,
          "sample": f"UNMODIFIED reference (baseline/_ref: l4p.l4p.L4PLitModule from its own configs/model.yaml, fp32 eager, "
                    f"{r['threads']} host threads): the full all-heads window, {r['steps_timed']} timed pass(es) after "
                    f"{min(warmup, 1)} warm-up, no sampling, no extrapolation"}
    extra = {"window_ms": 1e3 * r["mean_s"], "steps_timed": r["steps_timed"], "model_build_s": round(build_s, 1),
             "step_s": [round(x, 2) for x in r["step_s"]], "out_keys": r["out_keys"]}
    if with_cuda and torch.cuda.is_available():
        try:
            g = ref_arm.time_cuda_eager(model, batch, TASKS)
            extra["reference_eager_cuda"] = {
                "value": g["frames_per_s"], "unit": "frames/s", "ms_per_step": g["ms_per_step"], "autocast": g["autocast"],
                "peak_mem_gb": round(g["peak_mem_gb"], 2), "steps": g["steps"],
                "what": "the same unmodified reference modules on this B200: eager ATen/cuBLAS/cuDNN kernels under "
                        "torch.autocast(fp16) (= Lightning 16-mixed, demo/demo.py:22), inputs resident on the device"}
        except Exception as e:  # noqa: BLE001 - informational leg only
            extra["reference_eager_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    del model
    return cb, extra


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t_all = time.perf_counter()
    cb, extra = reference_cpu(args.steps, args.warmup, budget_s=float(os.environ.get("L4P_REF_BUDGET_S", "150")), with_cuda=True)
    v = cb["value"]
    line = {"impl": "reference", "metric": "frames/sec (16x224x224, all heads)", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": extra["window_ms"],
            "steps_timed": extra["steps_timed"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "clips_per_gpu": 1, "track_queries": NQ},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": None}
    line.update({k: x for k, x in extra.items() if k not in ("window_ms", "steps_timed")})
    line["wall_s"] = round(time.perf_counter() - t_all, 1)
    line["note"] = ("each step is one full un-sampled window; the run stops early once the time budget is spent "
                    "(steps_timed < steps), value = 16 frames / mean step time")
    for p in REF_ARM_FILES:
        try:
            p.parent.mkdir(parents=True, exist_ok=True)
            p.write_text(json.dumps(line))
        except OSError:
            pass
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=1)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="headline only (no cfg3 / cfg4 legs)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying the "
                    "captured CUDA graph of the step (l4p_b200/graph.py)")
    ap.add_argument("--ncu-range", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop "
                    "(run under `ncu --profile-from-start off`)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from l4p_b200 import build as _build

    if local == 0:
        _build.build()  # no-op when the in-tree library is current; the harness never runs without the CUDA library
    if world > 1:
        dist.barrier()
    from l4p_b200 import ops, weights
    from l4p_b200.config import load_model
    from l4p_b200.parallel import gather_clip_outputs

    dt = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    lit = load_model(device=dev, max_queries=NQ + 1, compute_dtype=dt)   # configs/model.yaml: all five tasks, joint alignment
    model = lit.l4p_model
    weights.fill_module_fast_(model, seed=0)    # the same (replicated) weights on every rank, as in deployment
    use_graph = not args.no_graph
    lit.enable_cuda_graph(use_graph)     # predict_step replays one captured graph per input signature
    copy_stream = torch.cuda.Stream(device=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(make_step, host, steps, warmup, instrument=False):
        """Device-resident and end-to-end timing of `make_step(batch) -> packed [units, n] fp32` (this rank's outputs; the
        step itself contains the exchange collective when world > 1). Returns ms, ms_e2e (max over ranks), byte counts."""
        batch = {k: v.to(dev) for k, v in host.items()}
        res = None
        for _ in range(warmup):
            res = make_step(batch)
        pinned = [torch.empty(res.shape, dtype=torch.float32).pin_memory() for _ in range(2)]
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = res.numel() * 4
        sync_all()
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as cs:
            if args.ncu_range and instrument:
                torch.cuda.profiler.start()
            e0.record()
            for _ in range(steps):
                make_step(batch)
            e1.record()
            sync_all()
            if args.ncu_range and instrument:
                torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1) / steps
        launches = (ops.LAUNCHES - l0) // steps
        # end to end: the user-facing call on PINNED HOST inputs (predict_step moves them to the device: H2D inside the timed
        # region), D2H of this rank's packed head outputs into double-buffered pinned memory on a copy stream, so that the copy
        # of step i overlaps the compute of step i+1; everything is complete before the closing event.
        sync_all()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_stream = torch.cuda.current_stream()
        e2.record()
        for i in range(steps):
            r = make_step(dict(host))
            done = torch.cuda.Event()
            done.record(main_stream)
            copy_stream.wait_event(done)
            with torch.cuda.stream(copy_stream):
                pinned[i % 2].copy_(r, non_blocking=True)
                r.record_stream(copy_stream)
        main_stream.wait_stream(copy_stream)
        e3.record()
        sync_all()
        ms_e2e = e2.elapsed_time(e3) / steps
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        return dict(ms=ms, ms_e2e=ms_e2e, h2d=h2d, d2h=d2h, launches=int(launches), clocks=cs.summary(), batch=batch)

    # ---------------------------------------------------------------- headline: cfg 2, one clip per GPU
    clips = args.clips_per_gpu

    def step_clips(n_clips):
        def step(b):
            out = lit.predict_step(b, 0)                       # the reference-facing call (l4p/l4p.py:107-109)
            packed = pack_outputs(out, OUT_KEYS, n_clips)      # [clips, unit]
            if world > 1:
                gather_clip_outputs(packed)                    # the single exchange step: head outputs over NVLink (stays on device)
            return packed                                      # this rank's shard: what its host process reads back
        return step

    with torch.no_grad():
        host = {k: v.pin_memory() for k, v in synth_batch(clips).items()}
        m = measure(step_clips(clips), host, args.steps, args.warmup, instrument=True)
        # roofline kernel: CUDA-event pair around every attention launch of two extra instrumented steps (kept out of the
        # timed region: an event record between two kernels breaks their programmatic-dependent-launch overlap)
        lit.use_cuda_graph = False          # eager for the instrumented steps (timing events are not capturable)
        ops.ATTN_EVENTS = []
        for _ in range(4):
            step_clips(clips)(m["batch"])
        sync_all()
        att = ops.ATTN_EVENTS
        ops.ATTN_EVENTS = None
        lit.use_cuda_graph = use_graph
        att_times = sorted(a.elapsed_time(b) for a, b, _ in att)
        att_ms = sum(att_times) / max(len(att_times), 1)             # mean over 160 launches (reported as `achieved`)
        att_ms_median = att_times[len(att_times) // 2] if att_times else 0.0
        att_flops = sum(f for _, _, f in att) / max(len(att), 1)
        # the same kernel on the step's own Q / K / V buffers, 40 launches back to back between ONE event pair (no event
        # record between launches, programmatic dependent launch intact): the kernel's duration without the measuring gaps
        ws = next(iter(model.video_encoder._ws.values()))
        hd = model.video_encoder.embed_dim // model.video_encoder.num_heads
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.attention(ws["q"], ws["k"], ws["vt"], ws["att"], hd, hd ** -0.5)
        b0.record()
        for _ in range(40):
            ops.attention(ws["q"], ws["k"], ws["vt"], ws["att"], hd, hd ** -0.5)
        b1.record()
        sync_all()
        att_ms_b2b = b0.elapsed_time(b1) / 40

        extra_cfgs = {}
        if not args.skip_configs:
            # ------------------------------------------------------------ cfg 3: 4 clips per GPU as one batch
            c3 = 4
            host3 = {k: v.pin_memory() for k, v in synth_batch(c3).items()}
            m3 = measure(step_clips(c3), host3, steps=4, warmup=2)
            f3 = 16 * c3 * world
            extra_cfgs["cfg3"] = {
                "workload": f"BASELINE.json configs[2] scaled to N GPUs: {c3} clips per GPU as one batch (encoder + DPT heads "
                            f"batched over clips, tracker per clip, 128 queries each), all heads; {c3 * world} clips total",
                "value": f3 / (m3["ms"] * 1e-3), "unit": "frames/s", "ms_per_step": m3["ms"], "steps": 4, "warmup": 2,
                "e2e": {"value": f3 / (m3["ms_e2e"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": m3["h2d"],
                        "d2h_bytes_per_step": m3["d2h"]},
                "gpu_launches": m3["launches"], "scaling": "weak"}
            del m3
            lit.enable_cuda_graph(False)     # drop the captured graphs (and their memory pools) of the clip legs
            torch.cuda.empty_cache()
            # ------------------------------------------------------------ cfg 4: long video, depth + camray, joint alignment
            keys4 = ["depth_est_b1thw", "traj3d_est_b16t", "traj3d_intrinsics_est_b16t"]
            tasks4 = ["depth", "camray"]
            if world > 1:
                model.enable_window_sharding(True)

            from l4p_b200.graph import StepGraph

            def step_video_eager(b):
                out = model.forward(b if b["rgb_b3thw"].is_cuda else {k: v.to(dev, non_blocking=True) for k, v in b.items()}, tasks4)
                return pack_outputs(out, keys4, 1)

            for T in (264, 512):
                hostv = {k: v.pin_memory() for k, v in synth_batch(1, T, queries=False).items()}
                step_video = step_video_eager
                # window-sharded runs are captured too when every rank holds at least one window (the exchange is then ONE NCCL
                # all-gather with no host-side metadata round, l4p_b200/parallel.py); L4P_CFG4_GRAPH=0 keeps them eager
                nW_ = (T - 16) // 8 + 1
                graph_ok = world == 1 or (nW_ >= world and -(-nW_ // world) * (world - 1) < nW_ and os.environ.get("L4P_CFG4_GRAPH", "1") != "0")
                if use_graph and graph_ok:
                    vg = StepGraph(lambda b: {"packed": pack_outputs(model.forward(b, tasks4), keys4, 1)}, hostv, dev, warmup=1)
                    step_video = lambda b, vg=vg: vg(b)["packed"]
                mv = measure(step_video, hostv, steps=2, warmup=1)
                step_video = None
                vg = None
                torch.cuda.empty_cache()
                nW = (T - 16) // 8 + 1
                extra_cfgs[f"cfg4_T{T}"] = {
                    "workload": f"BASELINE.json configs[3]: one {T}-frame video = {nW} overlapping 16-frame windows, depth + camray "
                                f"heads, joint Sim(3) window alignment on the GPU; windows sharded over {world} GPU(s), one "
                                f"all-gather of per-window depth + ray maps, alignment chain on every rank",
                    "value": T / (mv["ms"] * 1e-3), "unit": "video frames/s", "windows_per_s": nW / (mv["ms"] * 1e-3),
                    "ms_per_step": mv["ms"], "steps": 2, "warmup": 1,
                    "e2e": {"value": T / (mv["ms_e2e"] * 1e-3), "unit": "video frames/s", "h2d_bytes_per_step": mv["h2d"],
                            "d2h_bytes_per_step": mv["d2h"]},
                    "gpu_launches": mv["launches"], "scaling": "strong"}
                del mv, hostv
            if world > 1:
                # sharded == unsharded on real NCCL ranks: a 5-window clip through both paths on every rank, compared on the
                # heads whose windows are only STITCHED (flow, dyn-mask). With RANDOM weights the windows do not describe one
                # scene, so every overlap ALIGNMENT (depth's inverse-depth affine fit on near-constant maps, the joint Sim(3)
                # consensus, the pose fit on ray maps that are no camera) is ill-posed and amplifies the ~1e-4 round-off
                # differences between a 3 + 2 window and a 5 window batch arbitrarily; the aligned paths' sharded ==
                # unsharded equality is asserted on a consistent scene by tests/test_dist_gpu.py instead.
                b5 = {k: v.to(dev) for k, v in synth_batch(1, 48, queries=False).items()}
                tasks5 = ["flow_2d_backward", "dyn_mask"]
                keys5 = ["flow_2d_backward_est_b2thw", "dyn_mask_est_b1thw"]
                sh = model.forward(b5, tasks5)
                model.enable_window_sharding(False)
                un = model.forward(b5, tasks5)
                diffs = []
                for k in keys5:
                    a, b = sh[k].float(), un[k].float()
                    diffs.append(float((a - b).norm() / (b.norm() + 1e-30)))
                d = torch.tensor(diffs, device=dev, dtype=torch.float64)
                dist.all_reduce(d, op=dist.ReduceOp.MAX)
                extra_cfgs["cfg4_sharded_vs_unsharded_rel_l2"] = dict(zip(keys5, d.tolist()))
            torch.cuda.empty_cache()

    frames = 16 * clips * world
    if rank == 0:
        pk = peaks()
        ach = att_flops / (att_ms * 1e-3) / 1e12
        ncu = {}
        pj = ROOT / "profiles" / "attention_ncu_r2.json"
        if pj.exists():
            ncu = json.loads(pj.read_text())
        line = {
            "metric": "frames/sec (16x224x224, all heads)", "value": frames / (m["ms"] * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD + "; N>1: one clip per GPU + one all-gather of head outputs",
                       "clips_per_gpu": clips, "track_queries": NQ, "weights": "random (seeded), reference architecture",
                       "api": "L4PLitModule.predict_step(batch, 0)" + (" replaying the captured CUDA graph of the step" if use_graph else ""),
                       "cuda_graph": use_graph,
                       "l2": "per-step working set (2.8 GB weights + >1 GB activations) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": frames / (m["ms_e2e"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"],
                    "note": "per rank: pinned host batch -> predict_step (H2D inside) -> D2H of this rank's packed head outputs "
                            "(double-buffered pinned memory, copy stream); the all-gathered buffer stays on the device"},
            "gpu_launches": m["launches"],
            "clocks": m["clocks"],
            # the whole step against the tensor roofline: FLOPs of the REFERENCE's order of operations for this workload
            # (cfg 2: 23 316.6 GF per window; the folded track-head attention executes fewer)
            "step_roofline": {"reference_algorithmic_gflop": STEP_GFLOP * clips,
                              "tflops": STEP_GFLOP * clips / m["ms"], "peak": pk["tflops"],
                              "frac": STEP_GFLOP * clips / m["ms"] / pk["tflops"]},
            "roofline": {"kernel": "attention_kernel (fused QK^T+softmax+PV, tcgen05)", "bound": "tensor", "achieved": ach,
                         "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                         # dram__bytes_read + write per launch of the shipped kernel, from the committed ncu --set full capture
                         "traffic": ncu.get("dram_bytes_per_launch"), "traffic_source": ncu.get("source"),
                         "tensor_pipe_pct_ncu": ncu.get("tensor_pipe_pct"),
                         "peak_source": pk["src"], "launch_us": att_ms * 1e3, "launch_us_median": att_ms_median * 1e3,
                         "launch_us_back_to_back": att_ms_b2b * 1e3,
                         "frac_back_to_back": att_flops / (att_ms_b2b * 1e-3) / 1e12 / pk["tflops"],
                         "how": "achieved / frac: mean CUDA-event time of every attention launch of 4 instrumented steps (an "
                                "event pair around each launch: includes the launch gap the events themselves create); "
                                "*_back_to_back: 40 launches on the step's own Q/K/V between one event pair",
                         "algorithmic_flops_per_launch": att_flops},
        }
        if extra_cfgs:
            line["configs"] = extra_cfgs
        pp = ROOT / "profiles" / "parity_r2.json"
        if pp.exists():
            line["parity"] = json.loads(pp.read_text())
        if world == 1 and not args.no_cpu_baseline:
            del m
            torch.cuda.empty_cache()
            prior = next((p for p in REF_ARM_FILES if p.exists()), None)
            cb, extra = reference_cpu(steps=1, warmup=0, budget_s=1.0, with_cuda=False)
            line["cpu_baseline"] = dict(cb, window_ms=extra["window_ms"])
            if prior is not None:   # the --impl reference run of this box (the driver runs it right before this arm)
                try:
                    pr = json.loads(prior.read_text())
                    line["reference_arm"] = {k: pr[k] for k in ("value", "unit", "steps_timed", "cpu_baseline", "reference_eager_cuda")
                                             if k in pr}
                    line["reference_arm"]["source"] = str(prior)
                except Exception:  # noqa: BLE001
                    pass
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
