"""Per-k-block cost of the 2-CTA (M = 256) GEMM mainloop against the tile width N: clock64 timeline of CTA 0 (tools/gemm_prof.py) on the
fc2 shape (2048 x 1408 x 6144) with block_n forced. Shows what one k-block (4 UMMAs of M256 x N x K16) costs inside the kernel."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import ops
dt = torch.float16
M, N, K = 2048, 1408, 6144
x = torch.randn(M, K, device="cuda", dtype=dt); w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
bias = torch.zeros(N, device="cuda"); o = torch.empty(M, N, device="cuda", dtype=dt)
for pair in (1, -1):
    for bn in (64, 96, 128, 144, 160, 176, 192, 208, 224, 240, 256):
        prof = torch.zeros(3 * 512, device="cuda", dtype=torch.int64)
        for _ in range(3):
            ops.linear(x, w, bias=bias, out_16=o, cta_pair=pair, block_n=bn, prof=prof)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.linear(x, w, bias=bias, out_16=o, cta_pair=pair, block_n=bn)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        p = prof.cpu().view(3, 512)
        t0 = int(p[2, 511])
        mma = [int(v) - t0 for v in p[1] if v > 0][:96]
        d = sorted(b - a for a, b in zip(mma, mma[1:]))
        print(f"pair={pair:2d} block_n={bn:3d}: per-kb median {d[len(d)//2]:4d} cycles = {d[len(d)//2]/4:6.1f} per UMMA (nominal {bn/2 if pair==1 else bn/2:5.1f}); "
              f"{us:6.1f} us back to back")
