"""Per-call CUDA-event timing of one all-heads step, aggregated by (op, shape) -> where the time goes."""
import collections
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from l4p_b200 import ops, weights  # noqa: E402
from l4p_b200.config import load_model  # noqa: E402

dev = torch.device("cuda")
lit = load_model(device=dev, max_queries=bench.NQ + 1)
model = lit.l4p_model
weights.fill_module_fast_(model, seed=0)
batch = {k: v.to(dev) for k, v in bench.synth_batch(1).items()}
records = []


def wrap(name):
    fn = getattr(ops, name)

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        t = [x for x in a if torch.is_tensor(x)]
        if name in ("linear", "linear_qkv"):
            M, K, N = t[0].numel() // t[0].shape[-1], t[0].shape[-1], t[1].shape[0]
            desc, fl = f"M={M} N={N} K={K} act={k.get('act', 0)} o16={k.get('out_16') is not None}", 2 * M * N * K
        elif name == "conv3d":
            B, T, H, W, C = t[0].shape
            N = t[1].shape[0]
            desc, fl = f"{T}x{H}x{W} {C}->{N} head={k.get('head_w2') is not None}", 2 * B * T * H * W * 27 * C * N
        elif name in ("conv_transpose3d", "conv_transpose3d_hyper"):
            M, K, N = t[0].numel() // t[0].shape[-1], t[0].shape[-1], t[1].shape[0]
            desc, fl = f"M={M} N={N} K={K}", 2 * M * N * K
        else:
            desc, fl = "x".join(str(s) for s in t[0].shape), 0
        records.append((name, desc, fl, e0, e1))
        return r

    setattr(ops, name, w)


for n in ("linear", "linear_qkv", "conv3d", "conv_transpose3d", "conv_transpose3d_hyper", "layernorm", "layernorm16",
          "attention", "token_attention", "image_attention", "upsample3d", "track_readout", "cast16", "im2col3", "patchify"):
    wrap(n)
import l4p_b200.models.videomae, l4p_b200.models.task_heads.dpt, l4p_b200.models.task_heads.sparse_heads  # noqa
with torch.no_grad():
    for _ in range(2):
        records.clear()
        for c in range(1):
            lit.predict_step(dict(batch), 0)
        torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, desc, fl, e0, e1 in records:
    k = (name, desc)
    a = agg.setdefault(k, [0, 0.0, 0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += fl
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.2f} ms")
for (name, desc), (n, ms, fl) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    tf = f"{fl/ms/1e9:7.1f} TF/s" if fl else ""
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}% n={n:3d} {name:22s} {desc:50s} {tf}")
