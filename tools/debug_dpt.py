"""Stage-by-stage comparison of the DPT depth head against the oracle on random token features (GPU box)."""
import sys

import torch

sys.path.insert(0, ".")
from l4p_b200 import weights  # noqa: E402
from l4p_b200.models.task_heads.dense_heads import VideoMAEDepthDPTHead  # noqa: E402
from oracle import l4p_oracle as O  # noqa: E402

HOOKS = [14, 21, 28, 36]
head = VideoMAEDepthDPTHead("depth", depth_fn="linear", hooks_idx=HOOKS)
weights.fill_module_(head, seed=0)
sd = head.state_dict()
g = torch.Generator().manual_seed(1)
feats = [None] * 41
for h in HOOKS:
    feats[h] = torch.randn(1, 2048, 1408, generator=g) * (3.0 if h < 36 else 1.0)
dbg_ref = {}
with torch.no_grad():
    ref = O.dpt_forward(sd, "task_head.dpt.", feats, HOOKS, debug=dbg_ref)
head.task_head.dpt.debug = {}
out = head.forward([None if f is None else f.cuda() for f in feats])["depth_est_b1thw"]
torch.cuda.synchronize()
for k, v in head.task_head.dpt.debug.items():
    r = dbg_ref[k]
    e = (v.cpu() - r)
    print(f"{k}: shape {tuple(r.shape)} ref-rms {r.pow(2).mean().sqrt():.4f} rel-L2 {(e.norm()/r.norm()):.3e} max-abs {e.abs().max():.3e}")
e = out.cpu() - ref
print(f"out: ref-rms {ref.pow(2).mean().sqrt():.4f} ref-absmax {ref.abs().max():.3f} rel-L2 {(e.norm()/ref.norm()):.3e} max-abs {e.abs().max():.3e}")
