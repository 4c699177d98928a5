"""A/B of two arms of the attention kernel selected by an environment variable (default: L4P_ATT_SPLIT=0 -> the round-1
kernel with two full-row softmax warpgroups, 1 -> four half-row warpgroups): parity vs fp32 torch on the same rounded
operands, then back-to-back timing. The variable is read once per process, so each arm runs in its own subprocess; wrap the
whole call in `timeout` on the GPU box (every device-side wait is bounded and traps).

    timeout 300 python tools/att_pair_ab.py [ENVVAR [value ...]]      # e.g. L4P_ATT_POLY 0 1 2 3
"""
import os
import subprocess
import sys

ARM = os.environ.get("_ATT_ARM")
if ARM is None:
    var = sys.argv[1] if len(sys.argv) > 1 else "L4P_ATT_SPLIT"
    for arm in (sys.argv[2:] or ["0", "1"]):
        env = dict(os.environ, _ATT_ARM=arm, **{var: arm})
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=240)
        print(f"--- {var}={arm} (exit {r.returncode})\n{r.stdout}{r.stderr[-2000:]}")
    sys.exit(0)

import torch  # noqa: E402

sys.path.insert(0, ".")
from l4p_b200 import ops  # noqa: E402


def run(B, dtype, peaky=1.0):
    H, N, d, dp = 16, 2048, 88, 96
    g = torch.Generator().manual_seed(1)
    q = torch.zeros(B, H, N, dp, device="cuda", dtype=dtype)
    k = torch.zeros_like(q)
    vt = torch.zeros(B, H, dp, N, device="cuda", dtype=dtype)
    q[..., :d] = (torch.randn(B, H, N, d, generator=g) * peaky).to(dtype).cuda()
    k[..., :d] = torch.randn(B, H, N, d, generator=g).to(dtype).cuda()
    v = torch.randn(B, H, N, d, generator=g).to(dtype).cuda()
    vt[:, :, :d] = v.transpose(-1, -2)
    out = torch.empty(B * N, H * d, device="cuda", dtype=dtype)
    ops.attention(q, k, vt, out, d, d ** -0.5)
    torch.cuda.synchronize()
    bsel = 0
    s = (q[bsel, :, :, :d].float() @ k[bsel, :, :, :d].float().transpose(-1, -2)) * d ** -0.5
    ref = (torch.softmax(s, -1) @ v[bsel].float()).permute(1, 0, 2).reshape(N, H * d)
    got = out[:N].float()
    rel = ((got - ref).norm() / ref.norm()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        ops.attention(q, k, vt, out, d, d ** -0.5)
    e0.record()
    for _ in range(50):
        ops.attention(q, k, vt, out, d, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    tf = 4 * N * N * d * H * B / (us * 1e-6) / 1e12
    print(f"B={B} {str(dtype)[6:]} peaky={peaky}: rel-L2 {rel:.2e} | {us:.1f} us / launch, {tf:.0f} TFLOP/s algorithmic")


for B in (1, 8):
    for dt in (torch.float16, torch.bfloat16):
        run(B, dt)
run(1, torch.float16, peaky=6.0)   # exercises the lazy TMEM rescale
