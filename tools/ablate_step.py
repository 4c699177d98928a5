"""Marginal in-graph cost of every kernel family of the all-heads step (the number an optimisation can actually win).

Per-op CUDA events (tools/profile_step.py) add a launch gap to every kernel and break the programmatic-dependent-launch
overlap; the ncu launch list is cold-cache and serialised. This tool measures what matters for bench.py instead: the step is
captured as a CUDA graph (as bench.py replays it) once per ablation, each time with ONE kernel family replaced by a no-op
(its outputs stay whatever the allocator left there - timing only, results are garbage), and the difference to the full
graph is that family's marginal cost with the head streams running concurrently and PDL intact.

  python tools/ablate_step.py [--only encoder] [--steps 10]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from l4p_b200 import ops, weights  # noqa: E402
from l4p_b200.config import load_model  # noqa: E402
from l4p_b200.graph import StepGraph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--only", default="", help="substring filter on the ablation names")
ap.add_argument("--serial", action="store_true", help="all heads on one stream (no inter-stream concurrency)")
args = ap.parse_args()

dev = torch.device("cuda")
lit = load_model(device=dev, max_queries=bench.NQ + 1, compute_dtype=torch.float16)
model = lit.l4p_model
weights.fill_module_fast_(model, seed=0)
lit.enable_cuda_graph(False)
batch = {k: v.to(dev) for k, v in bench.synth_batch(1).items()}

ORIG = {n: getattr(ops, n) for n in ("layernorm", "linear", "linear_qkv", "attention", "conv3d", "conv_transpose3d", "upsample3d",
                                     "conv_transpose3d_hyper", "token_attention", "image_attention", "layernorm16", "track_readout",
                                     "cast16", "im2col3", "patchify", "token_weighted_sum", "row_softmax16", "group_softmax_t16",
                                     "head_expand", "head_diag_gather")}


def rows(t):
    return t.numel() // t.shape[-1]


def lin(pred):
    def f(a, w, **kw):
        return pred(rows(a), w.shape[0], w.shape[1])
    return f


# name -> (ops function, predicate over its arguments: True = skip this call)
ABL = {
    "enc.layernorm(2048x1408)": ("layernorm", lambda x, *a, **k: rows(x) == 2048),
    "enc.qkv": ("linear_qkv", lambda *a, **k: True),
    "enc.attention": ("attention", lambda *a, **k: True),
    "enc.proj(2048,1408,1408)": ("linear", lin(lambda M, N, K: (M, N, K) == (2048, 1408, 1408))),
    "enc.fc1(2048,6144,1408)": ("linear", lin(lambda M, N, K: (M, N, K) == (2048, 6144, 1408))),
    "enc.fc2(2048,1408,6144)": ("linear", lin(lambda M, N, K: (M, N, K) == (2048, 1408, 6144))),
    "cast16(2048 rows: encoder taps)": ("cast16", lambda x, *a, **k: rows(x) == 2048),
    "cast16(other: track token side)": ("cast16", lambda x, *a, **k: rows(x) != 2048),
    "trk.layernorm(262144x1408)": ("layernorm", lambda x, *a, **k: rows(x) == 262144),
    "trk.kvq(262144,704,1408)": ("linear", lin(lambda M, N, K: (M, N, K) == (262144, 704, 1408))),
    "trk.outproj(262144,1408,704)": ("linear", lin(lambda M, N, K: (M, N, K) == (262144, 1408, 704))),
    "trk.fold.scores(6144,2048,1408)": ("linear", lin(lambda M, N, K: M == 6144 and K == 1408 and N in (2048, 128 * 2048))),
    "trk.fold.small_linear(M=6144)": ("linear", lin(lambda M, N, K: M == 6144 and not (K == 1408 and N in (2048, 128 * 2048)))),
    "trk.fold.out(262144,1408,48)": ("linear", lin(lambda M, N, K: M == 262144 and K == 48)),
    "trk.fold.token_weighted_sum": ("token_weighted_sum", lambda *a, **k: True),
    "trk.fold.softmaxes": ("row_softmax16", lambda *a, **k: True),
    "trk.fold.group_softmax": ("group_softmax_t16", lambda *a, **k: True),
    "trk.fold.expand+gather": ("head_expand", lambda *a, **k: True),
    "trk.other_linear(M<=768)": ("linear", lin(lambda M, N, K: M <= 768)),
    "trk.token_attention": ("token_attention", lambda *a, **k: True),
    "trk.image_attention": ("image_attention", lambda *a, **k: True),
    "trk.layernorm16": ("layernorm16", lambda *a, **k: True),
    "trk.convT": ("conv_transpose3d", lambda x, w, *a, **k: x.shape[-1] == 1408),
    "trk.convT_hyper": ("conv_transpose3d_hyper", lambda *a, **k: True),
    "dpt.conv3d(224^2)": ("conv3d", lambda x, *a, **k: x.shape[2] == 224),
    "dpt.conv3d(128^2)": ("conv3d", lambda x, *a, **k: x.shape[2] == 128),
    "dpt.conv3d(64^2)": ("conv3d", lambda x, *a, **k: x.shape[2] == 64),
    "dpt.conv3d(<=32^2)": ("conv3d", lambda x, *a, **k: x.shape[2] <= 32),
    "dpt.upsample3d": ("upsample3d", lambda *a, **k: True),
    "dpt.convT": ("conv_transpose3d", lambda x, w, *a, **k: x.shape[-1] != 1408),
    "dpt.linear(2048 rows, heads' 1x1)": ("linear", lin(lambda M, N, K: M == 2048 and N <= 1024)),
    "dpt.linear(other)": ("linear", lin(lambda M, N, K: M not in (2048, 262144, 6144) and M > 768)),
}


def time_graph(skip=None):
    for n, f in ORIG.items():
        setattr(ops, n, f)
    saved = []
    if skip is not None:
        name, pred = skip
        orig = ORIG[name]

        def wrapped(*a, **k):
            if pred(*a, **k):
                saved.append((a, k))
                return None
            return orig(*a, **k)
        setattr(ops, name, wrapped)
    g = StepGraph(lambda b: lit.l4p_model.forward(b, bench.TASKS), batch, dev, warmup=1)
    for n, f in ORIG.items():
        setattr(ops, n, f)
    # The skipped kernels' outputs would stay never-written graph-pool memory (zeros): GEMMs on zeros draw less power and the
    # power-capped clock rises, which credits the ablated kernel with time it never used. Fill them once with realistic data:
    # replay, run the skipped calls of the CAPTURE pass eagerly on the (now populated) graph buffers, replay again.
    if skip is not None:
        ncap = len(saved) // 2          # the warm-up pass recorded the same calls on eager-pool tensors first
        g(batch)
        for a, k in saved[ncap:]:
            ORIG[skip[0]](*a, **k)
        torch.cuda.synchronize()
    for _ in range(3):
        g(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g(batch)
    e1.record()
    torch.cuda.synchronize()
    ms, n = e0.elapsed_time(e1) / args.steps, g.launches
    del g, saved
    return ms, n


with torch.no_grad():
    model.parallel_heads = False
    ser, _ = time_graph()
    model.parallel_heads = True
    par, _ = time_graph()
    print(f"heads on one stream: {ser:.3f} ms; heads on concurrent streams: {par:.3f} ms")
    model.parallel_heads = not args.serial
    full, nfull = time_graph()
    full2, _ = time_graph()
    print(f"full step: {full:.3f} ms ({nfull} launches); repeat {full2:.3f} ms")
    res = []
    for name, spec in ABL.items():
        if args.only and args.only not in name:
            continue
        try:
            ms, n = time_graph(spec)
        except Exception as e:  # noqa: BLE001 - a skipped producer can break a consumer's argument checks
            print(f"{name}: failed ({type(e).__name__}: {str(e)[:100]})")
            continue
        res.append((full - ms, name, nfull - n))
    res.sort(reverse=True)
    tot = 0.0
    for d, name, n in res:
        tot += d
        print(f"  {d:7.3f} ms  {100 * d / full:5.1f}%  n={n:4d}  {name}")
    print(f"sum of marginal costs {tot:.3f} ms of {full:.3f} ms")
