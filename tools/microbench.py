"""Kernel micro-benchmarks on the encoder shapes (CUDA-event timed, L2 flushed between iterations).
Usage: python tools/microbench.py [B]   (B = number of 16x224x224 windows)."""
import sys

import torch

sys.path.insert(0, ".")
from l4p_b200 import lib, ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dt = torch.float16
    M, D, Hd = B * 2048, 1408, 6144
    a = torch.randn(M, D, device="cuda", dtype=dt)
    for name, N, K, kw in [("qkv", 4224, D, {}), ("proj", D, D, {}), ("fc1", Hd, D, dict(act=lib.ACT_GELU)), ("fc2", D, Hd, {})]:
        x = torch.randn(M, K, device="cuda", dtype=dt)
        w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
        bias = torch.zeros(N, device="cuda")
        o = torch.empty(M, N, device="cuda", dtype=dt)
        ms = timeit(lambda: ops.linear(x, w, bias=bias, out_16=o, **kw))
        print(f"gemm {name:5s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s")
    q = torch.randn(B, 16, 2048, 96, device="cuda", dtype=dt)
    k = torch.randn_like(q)
    vt = torch.randn(B, 16, 96, 2048, device="cuda", dtype=dt)
    o = torch.empty(B * 2048, D, device="cuda", dtype=dt)
    ms = timeit(lambda: ops.attention(q, k, vt, o, 88, 88 ** -0.5))
    print(f"attention B={B}: {ms*1e3:8.1f} us  {4*2048*2048*88*16*B/ms/1e9:8.1f} TFLOP/s (unpadded d=88)")
    x = torch.randn(M, D, device="cuda")
    g = torch.ones(D, device="cuda")
    ms = timeit(lambda: ops.layernorm(x, g, g, 1e-6, out16=a))
    print(f"layernorm M={M}: {ms*1e3:8.1f} us  {M*D*6/ms/1e6:8.1f} GB/s")
    for (T, H, W, Ci, Co) in [(16, 64, 64, 256, 256), (16, 128, 128, 256, 128), (16, 224, 224, 128, 128)]:
        xx = torch.randn(B, T, H, W, Ci, device="cuda", dtype=dt)
        ww = torch.randn(Co, 27 * Ci, device="cuda", dtype=dt) * (27 * Ci) ** -0.5
        oo = torch.empty(B, T, H, W, Co, device="cuda", dtype=dt)
        ms = timeit(lambda: ops.conv3d(xx, ww, ksize=(3, 3, 3), out_16=oo), iters=5)
        print(f"conv3d {T}x{H}x{W} {Ci}->{Co}: {ms*1e3:8.1f} us  {2*B*T*H*W*27*Ci*Co/ms/1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
