"""A/B of a track-head switch at the bench size (default L4P_TRACK_FOLD_T2I: reference-order K / V projections vs the folded
token -> video-token attention; `python tools/track_ab.py L4P_TRACK_STREAM16` for the fp32 vs 16-bit token stream): time of
one 128-query window and the difference of the two arms' tracks (each arm in its own subprocess writes its outputs, the
parent compares)."""
import os
import subprocess
import sys
import tempfile

if os.environ.get("_TRK_ARM") is None:
    d = tempfile.mkdtemp()
    var = sys.argv[1] if len(sys.argv) > 1 else "L4P_TRACK_FOLD_T2I"
    for arm in ("0", "1"):
        r = subprocess.run([sys.executable, __file__], env=dict(os.environ, _TRK_ARM=arm, _TRK_OUT=d, **{var: arm}),
                           capture_output=True, text=True, timeout=280)
        print(f"--- {var}={arm} (exit {r.returncode})\n{r.stdout}{r.stderr[-1500:]}")
    import torch

    a, b = torch.load(os.path.join(d, "0.pt")), torch.load(os.path.join(d, "1.pt"))
    for k in a:
        print(f"{k}: max abs diff {(a[k] - b[k]).abs().max():.4g} (max |value| {a[k].abs().max():.4g})")
    sys.exit(0)

import torch  # noqa: E402

sys.path.insert(0, ".")
import bench  # noqa: E402
from l4p_b200 import weights  # noqa: E402
from l4p_b200.models.task_heads.sparse_heads import VideoMAETrack2DSamHead  # noqa: E402

h = VideoMAETrack2DSamHead(task_name="track_2d", estimate_vis=True, estimate_depth=True, sam_head_depth=2, num_point_embeddings=2,
                           prompt_using_features=True, attend_to_past=True, modify_pointlabels_for_windowing=True,
                           estimation_directions=[1], depth_fn="exp", vis_fn="linear", max_queries=129).cuda()
weights.fill_module_fast_(h, seed=3)
g = torch.Generator(device="cuda").manual_seed(5)
feat = torch.randn(1, 2048, 1408, device="cuda", generator=g)
b = bench.synth_batch(1)
q, lab = b["track_2d_pointquerries_bn3"].cuda(), b["track_2d_pointlabels_bn"].cuda()
feats = [None] * 40 + [feat]
with torch.no_grad():
    for _ in range(3):
        out = h.forward_windowed([feats], q, lab, time_strides=torch.tensor([0]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = h.forward_windowed([feats], q, lab, time_strides=torch.tensor([0]))
    e1.record()
    torch.cuda.synchronize()
print(f"track head, 128 queries, one window: {e0.elapsed_time(e1) / 10:.2f} ms")
torch.save({k: v.float().cpu() for k, v in out.items()}, os.path.join(os.environ["_TRK_OUT"], os.environ["_TRK_ARM"] + ".pt"))
