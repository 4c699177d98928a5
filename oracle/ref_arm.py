"""TEST/BENCH INFRASTRUCTURE ONLY (never imported by l4p_b200/). Runs the UNMODIFIED reference - its own modules, its own
`configs/model.yaml`, its own `L4PLitModule.forward` - as the comparator arm of `bench.py --impl reference`.

The reference is pure Python with no setup.py, so "installing" it is a copy of its package: `stage()` (called by
`__graft_entry__.build()` in the build container, where /root/reference exists) copies `l4p/` and `configs/` into the
git-ignored `baseline/_ref/`, which travels to the GPU box with the snapshot exactly like the built `.so` does. Nothing is
edited: missing third-party packages (timm, lightning, scikit-image, kornia) are answered by the `sys.modules` shims of
`oracle/ref_loader.py`, jsonargparse by its ~20-line `instantiate`.

Two timings on the SAME full BASELINE.json configs[1] workload (one 16x224x224 clip, all five tasks, 128 track queries,
the reference's shipped config: windowed path + joint alignment), no sampling, no extrapolation:
  * `time_cpu`        - fp32 eager on all host threads (the headline comparator: cpu_baseline.kind = "reference")
  * `time_cuda_eager` - the same modules moved to the B200, eager ATen kernels under `torch.autocast(fp16)`, i.e. what the
                        reference's demo does (`demo/demo.py:22-23`, `l4p/models/utils.py:57-58`); informational
                        (`reference_eager_cuda`): this is the bar SURVEY.md section 2.1 names.
Weights are the reference's own random initialisation (no checkpoint is available offline); speed does not depend on them.
"""
from __future__ import annotations

import os
import shutil
import time
from pathlib import Path
from typing import Dict, List, Optional

ROOT = Path(__file__).resolve().parents[1]
STAGED = ROOT / "baseline" / "_ref"
SOURCE = Path("/root/reference")
TASKS = ["flow_2d_backward", "track_2d", "depth", "dyn_mask", "camray"]


def stage(force: bool = False) -> Optional[Path]:
    """Copy the reference's Python package + config (unmodified) to baseline/_ref. No-op without /root/reference."""
    if not (SOURCE / "l4p").is_dir():
        return STAGED if (STAGED / "l4p").is_dir() else None
    if (STAGED / "l4p").is_dir() and not force:
        return STAGED
    STAGED.mkdir(parents=True, exist_ok=True)
    for sub in ("l4p", "configs"):
        dst = STAGED / sub
        if dst.exists():
            shutil.rmtree(dst)
        shutil.copytree(SOURCE / sub, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("LICENSE",):
        if (SOURCE / f).exists():
            shutil.copy2(SOURCE / f, STAGED / f)
    return STAGED


def root() -> Optional[Path]:
    for cand in (SOURCE, STAGED):
        if (cand / "l4p").is_dir():
            return cand
    return None


def build_model(max_queries: int = 128):
    """The reference's LightningModule from the reference's own configs/model.yaml (l4p/models/utils.py:37-49 semantics:
    yaml -> class_path/init_args instantiation, `max_queries` override at the same config node)."""
    import yaml

    from . import ref_loader

    r = root()
    if r is None:
        raise RuntimeError("reference sources not found (neither /root/reference nor baseline/_ref)")
    ref_loader.REFERENCE_ROOT = str(r)
    ref_loader.load()
    with open(r / "configs" / "model.yaml") as f:
        cfg = yaml.safe_load(f)
    cfg["init_args"]["l4p_model"]["init_args"]["task_heads"]["init_args"]["modules"]["track_2d"]["init_args"]["max_queries"] = max_queries
    model = ref_loader.instantiate(cfg)
    return model.eval()


def time_cpu(model, batch: Dict, tasks: List[str], steps: int, warmup: int, budget_s: float = 200.0) -> Dict:
    """Full-window fp32 eager forward on all host threads; stops early when `budget_s` of timed work is spent (at least one
    timed step is always completed). Returns per-step seconds and how many steps were actually timed."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    times: List[float] = []
    with torch.no_grad():
        t_w = time.perf_counter()
        for _ in range(min(warmup, 1)):   # one full window is already > 10^13 flop: a single warm-up pass
            model.forward(dict(batch), tasks)
        warm_s = time.perf_counter() - t_w
        spent = 0.0
        for i in range(max(steps, 1)):
            t0 = time.perf_counter()
            out = model.forward(dict(batch), tasks)
            dt = time.perf_counter() - t0
            times.append(dt)
            spent += dt
            if spent + dt > budget_s:
                break
    keys = sorted(k for k in out if not k.startswith("enc_features"))
    return {"step_s": times, "mean_s": sum(times) / len(times), "steps_timed": len(times), "warmup_s": warm_s,
            "threads": torch.get_num_threads(), "out_keys": keys}


def time_cuda_eager(model, batch: Dict, tasks: List[str], steps: int = 5, warmup: int = 2, dtype: str = "fp16") -> Dict:
    """The reference's eager GPU path: same modules on cuda:0, `torch.autocast` like Lightning Fabric's 16-mixed."""
    import torch

    dev = torch.device("cuda", 0)
    model = model.to(dev)
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    adt = torch.float16 if dtype == "fp16" else torch.bfloat16
    with torch.no_grad(), torch.autocast("cuda", dtype=adt):
        for _ in range(warmup):
            model.forward(dict(b), tasks)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model.forward(dict(b), tasks)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "frames_per_s": 16.0 * b["rgb_b3thw"].shape[0] / (ms * 1e-3), "steps": steps, "warmup": warmup,
            "autocast": dtype, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
