"""A/B of the 2-CTA conv kernel's line-halo stages (L4P_CONV_HALO=0/1, read once per process: one subprocess per arm) on the
narrow DPT convolutions: the 224^2 head conv (128 -> 128 + ReLU + 1x1x1) and the 128^2 conv 256 -> 128."""
import os
import subprocess
import sys

if os.environ.get("_ARM") is None:
    for arm in ("0", "1"):
        r = subprocess.run([sys.executable, __file__], env=dict(os.environ, _ARM=arm, L4P_CONV_HALO=arm), capture_output=True, text=True, timeout=280)
        print(f"--- L4P_CONV_HALO={arm} (exit {r.returncode})\n{r.stdout}{r.stderr[-1500:]}")
    sys.exit(0)

import torch  # noqa: E402

sys.path.insert(0, ".")
from l4p_b200 import ops  # noqa: E402

dt = torch.float16
dev = "cuda"


def timeit(name, f, flops, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:52s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TF/s")


x = torch.randn(1, 16, 224, 224, 128, device=dev, dtype=dt)
w = torch.randn(128, 27 * 128, device=dev, dtype=dt) * 0.02
b = torch.zeros(128, device=dev)
w2 = torch.randn(2, 128, device=dev) * 0.1
b2 = torch.zeros(2, device=dev)
of = torch.empty(1, 2, 16, 224, 224, device=dev)
timeit("conv 16x224x224 128->128 head (ReLU + 1x1x1)", lambda: ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, head_w2=w2, head_b2=b2, out_f32=of),
       2.0 * 16 * 224 * 224 * 128 * 27 * 128)
x2 = torch.randn(1, 16, 128, 128, 256, device=dev, dtype=dt)
w3 = torch.randn(128, 27 * 256, device=dev, dtype=dt) * 0.02
o2 = torch.empty(1, 16, 128, 128, 128, device=dev, dtype=dt)
timeit("conv 16x128x128 256->128", lambda: ops.conv3d(x2, w3, ksize=(3, 3, 3), bias=b, out_16=o2), 2.0 * 16 * 128 * 128 * 128 * 27 * 256)
x3 = torch.randn(1, 16, 64, 64, 256, device=dev, dtype=dt)
w4 = torch.randn(256, 27 * 256, device=dev, dtype=dt) * 0.02
b4 = torch.zeros(256, device=dev)
o3 = torch.empty(1, 16, 64, 64, 256, device=dev, dtype=dt)
timeit("conv 16x64x64 256->256", lambda: ops.conv3d(x3, w4, ksize=(3, 3, 3), bias=b4, out_16=o3), 2.0 * 16 * 64 * 64 * 256 * 27 * 256)
x4 = torch.randn(1, 16, 32, 32, 256, device=dev, dtype=dt)
o4 = torch.empty(1, 16, 32, 32, 256, device=dev, dtype=dt)
timeit("conv 16x32x32 256->256", lambda: ops.conv3d(x4, w4, ksize=(3, 3, 3), bias=b4, out_16=o4), 2.0 * 16 * 32 * 32 * 256 * 27 * 256)
x5 = torch.randn(1, 16, 32, 32, 512, device=dev, dtype=dt)
w5 = torch.randn(256, 27 * 512, device=dev, dtype=dt) * 0.02
timeit("conv 16x32x32 512->256", lambda: ops.conv3d(x5, w5, ksize=(3, 3, 3), bias=b4, out_16=o4), 2.0 * 16 * 32 * 32 * 256 * 27 * 512)
