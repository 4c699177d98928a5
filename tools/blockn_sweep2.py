"""Tile-width sweep of the matrix-mode GEMM after the two-warp producer / 128-wide K stages: encoder shapes at 1, 4 and 8 windows and
the track head's M = 262144 GEMMs, block_n forced from 96 to 256 against the automatic choice (back-to-back launches, CUDA events)."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


BNS = (0, 96, 112, 128, 144, 160, 176, 192, 208, 224, 240, 256)
shapes = []
for B in (1, 4, 8):
    M = 2048 * B
    shapes += [(f"proj B={B}", M, 1408, 1408, "res"), (f"fc2 B={B}", M, 1408, 6144, "res"), (f"fc1 B={B}", M, 6144, 1408, "gelu"),
               (f"qkv B={B}", M, 4224, 1408, "qkv")]
shapes += [("trk out-proj 704", 32 * 2048, 1408, 704, "o16"), ("trk kvq 1408->704", 32 * 2048, 704, 1408, "o16"),
           ("dpt token linear", 2048, 1024, 1408, "o16"), ("trk M=6144 small", 6144, 1408, 1408, "o16")]
only = sys.argv[1:] 
for name, M, N, K, kind in shapes:
    if only and not any(o in name for o in only): continue
    x = torch.randn(M, K, device="cuda", dtype=dt)
    w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.zeros(N, device="cuda")
    if kind == "res":
        r32 = torch.randn(M, N, device="cuda")
        run = lambda bn: ops.linear(x, w, bias=b, res_f32=r32, out_f32=r32, block_n=bn)
    elif kind == "gelu":
        o = torch.empty(M, N, device="cuda", dtype=dt)
        run = lambda bn: ops.linear(x, w, bias=b, act=lib.ACT_GELU, out_16=o, block_n=bn)
    elif kind == "o16":
        o = torch.empty(M, N, device="cuda", dtype=dt)
        run = lambda bn: ops.linear(x, w, bias=b, out_16=o, block_n=bn)
    else:
        Bc = M // 2048
        q5 = torch.zeros(Bc, 16, 2048, 96, device="cuda", dtype=dt); k5 = torch.zeros_like(q5); v5 = torch.zeros(Bc, 16, 96, 2048, device="cuda", dtype=dt)
        run = lambda bn: ops.linear_qkv(x, w, b, q5, k5, v5, 16, 88, 2048, block_n=bn)
    res = []
    for bn in BNS:
        if bn > N: continue
        try:
            us = timeit(lambda: run(bn), 20 if M <= 16384 else 5)
        except Exception as e:  # a forced width the store mode cannot take
            continue
        res.append((bn, us))
    auto = res[0][1]
    best = min(res[1:], key=lambda t: t[1])
    print(f"{name:22s} M={M:6d} N={N:5d} K={K:5d}: auto {auto:7.1f} us ({2*M*N*K/auto/1e6:5.0f} TF/s); best block_n={best[0]} {best[1]:7.1f} us; "
          + " ".join(f"{bn}:{us:.1f}" for bn, us in res[1:]))
