"""Mask-decoder hyper ConvT (ConvT(1,2,2) 352 -> 176 + GELU + hyper-network dot, 128 queries) and the fc1 GELU GEMM, back to back."""
import sys, torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt, dev = torch.float16, "cuda"
G = 128
x = torch.randn(G, 16, 32, 32, 352, device=dev, dtype=dt)
w = torch.randn(4 * 176, 352, device=dev, dtype=dt) * 0.05
b = torch.zeros(4 * 176, device=dev)
hy = torch.randn(G, 3, 176, device=dev)
om = torch.empty(G, 3, 16, 64, 64, device=dev)
a2 = torch.randn(2048, 1408, device=dev, dtype=dt); w3 = torch.randn(6144, 1408, device=dev, dtype=dt) * 0.03
b3 = torch.zeros(6144, device=dev); o3 = torch.empty(2048, 6144, device=dev, dtype=dt)


def t(name, f, flops, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:40s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TF/s")


t("hyper ConvT (G=128)", lambda: ops.conv_transpose3d_hyper(x, w, b, (1, 2, 2), hy, om, act=lib.ACT_GELU), 2.0 * G * 16384 * 704 * 352)
t("fc1 (2048 x 6144 x 1408, GELU)", lambda: ops.linear(a2, w3, bias=b3, act=lib.ACT_GELU, out_16=o3), 2.0 * 2048 * 6144 * 1408, n=40)
