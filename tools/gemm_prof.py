"""clock64 timeline of CTA 0 of the GEMM kernel on the encoder shapes: producer issue, MMA full-wait, epilogue."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16
M = 2048
def run(name, N, K, pair, **kw):
    x = torch.randn(M, K, device="cuda", dtype=dt); w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    bias = torch.zeros(N, device="cuda"); o = torch.empty(M, N, device="cuda", dtype=dt)
    prof = torch.zeros(3 * 512, device="cuda", dtype=torch.int64)
    for _ in range(3):
        ops.linear(x, w, bias=bias, out_16=o, cta_pair=pair, prof=prof, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.linear(x, w, bias=bias, out_16=o, cta_pair=pair, **kw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    p = prof.cpu().view(3, 512)
    t0 = int(p[2, 511])
    prod = [int(v) - t0 for v in p[0] if v > 0]
    mma = [int(v) - t0 for v in p[1] if v > 0]
    epi = [int(v) - t0 for v in p[2, :16] if v > 0]
    print(f"== {name} M={M} N={N} K={K} pair={pair}: {us:.1f} us back-to-back, {2*M*N*K/us/1e6:.0f} TF/s; kb events {len(mma)}")
    print("  producer issue (first 12):", prod[:12])
    print("  mma full-wait done (first 12):", mma[:12])
    nkb = (K + 63) // 64
    for t in range(len(mma) // nkb):
        seg = mma[t * nkb:(t + 1) * nkb]
        d = [b - a for a, b in zip(seg, seg[1:])]
        print(f"  tile {t}: mma first {seg[0]} last {seg[-1]} span {seg[-1]-seg[0]} per-kb median {sorted(d)[len(d)//2]} max {max(d)}")
    print("  epilogue (release, end) per tile:", epi)
run("fc1", 6144, 1408, 0, act=lib.ACT_GELU)
run("fc1-nogelu", 6144, 1408, 0)
run("fc1-1cta", 6144, 1408, -1, act=lib.ACT_GELU)
run("proj", 1408, 1408, 0)
run("qkv-like", 4224, 1408, 0)
run("fc2", 1408, 6144, 0)
