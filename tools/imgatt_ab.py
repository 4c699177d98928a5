"""A/B at the real size (128 queries x 2048 video tokens, 8 heads of 88, 6 keys): round-1 image-attention kernel
(L4P_IMGATT_STREAM=0) vs the round-2 streaming kernel (default: bulk-TMA ring + tensor-core formulation with K / V in
registers), each in its own subprocess (the selector is
read once per process): parity of 4096 sampled rows against fp32 torch, time per launch, achieved HBM GB/s
(algorithmic bytes = read + write of the [G*Np, 704] 16-bit stream = 738 MB)."""
import os
import subprocess
import sys

if os.environ.get("_IA_ARM") is None:
    for arm in (sys.argv[1:] or ["0", "1"]):
        r = subprocess.run([sys.executable, __file__], env=dict(os.environ, _IA_ARM=arm, L4P_IMGATT_STREAM=arm),
                           capture_output=True, text=True, timeout=240)
        print(f"--- L4P_IMGATT_STREAM={arm} (exit {r.returncode})\n{r.stdout}{r.stderr[-1500:]}")
    sys.exit(0)

import torch  # noqa: E402

sys.path.insert(0, ".")
from l4p_b200 import ops  # noqa: E402

G, Np, nk, H, d = 128, 2048, 6, 8, 88
for dt in (torch.float16, torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(1)
    q16 = torch.randn(G * Np, H * d, device="cuda", generator=g).to(dt)
    k = torch.randn(G, nk, H * d, device="cuda", generator=g)
    v = torch.randn(G, nk, H * d, device="cuda", generator=g)
    out = torch.empty_like(q16)
    ops.image_attention(q16, k, v, out, G, H, d ** -0.5)
    torch.cuda.synchronize()
    worst = 0.0
    for gi in (0, 63, 127):
        rows = slice(gi * Np + 512, gi * Np + 512 + 1365)
        qf = q16[rows].float().view(-1, H, d)
        kf, vf = k[gi].view(nk, H, d), v[gi].view(nk, H, d)
        a = torch.softmax(torch.einsum("rhd,jhd->rhj", qf, kf) * d ** -0.5, dim=-1)
        ref = torch.einsum("rhj,jhd->rhd", a, vf).reshape(-1, H * d)
        worst = max(worst, ((out[rows].float() - ref).abs().max() / ref.abs().max()).item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.image_attention(q16, k, v, out, G, H, d ** -0.5)
    e0.record()
    for _ in range(10):
        ops.image_attention(q16, k, v, out, G, H, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"{str(dt)[6:]}: max rel err {worst:.2e} | {us:.0f} us / launch, {2 * q16.numel() * 2 / us / 1e3:.0f} GB/s algorithmic")
