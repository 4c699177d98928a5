"""clock64 timeline of CTA 0 for the fused-dot GEMM modes (mask-decoder hyper ConvT, DPT head conv)."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16
dev = "cuda"
def t(*s, dtype=dt): return torch.randn(*s, device=dev, dtype=dtype)
def report(name, fn, flops, nkb):
    prof = torch.zeros(3 * 512, device=dev, dtype=torch.int64)
    fn(None); fn(None)
    prof.zero_()
    fn(prof)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn(None)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    p = prof.cpu().view(3, 512)
    t0 = int(p[2, 511])
    mma = [int(v) - t0 for v in p[1] if v > 0]
    epi = [int(v) - t0 for v in p[2, :40] if v > 0]
    prod = [int(v) - t0 for v in p[0, :256] if v > 0]
    print(f"== {name}: {us:.1f} us, {flops/us/1e6:.0f} TF/s")
    print("  producer issue:", prod[:24])
    for k in range(min(6, len(mma) // nkb)):
        seg = mma[k * nkb:(k + 1) * nkb]
        print(f"  tile {k}: mma first {seg[0]} last {seg[-1]} span {seg[-1]-seg[0]}")
    print("  epilogue (release, end):", epi[:24])
    fine = [int(v) - t0 for v in p[2, 64:64 + 60] if v > 0]
    hs = [int(v) - t0 for v in p[0, 300:300 + 40] if v > 0]
    if hs:
        print("  helper stamps [stage: start, sempty ok, staged | finalize: pfull ok, done]:", hs)
    if fine:
        print("  fine stamps [tile: pre-stage, staged, acc-ready, chunk..., drained, end]:", fine)
G = 4
x = t(G, 16, 32, 32, 352); w = t(4 * 176, 352) * 0.05; b = torch.zeros(4 * 176, device=dev); hy = t(G, 3, 176, dtype=torch.float32)
om = torch.empty(G, 3, 16, 64, 64, device=dev)
report("hyper", lambda pr: ops.conv_transpose3d_hyper(x, w, b, (1, 2, 2), hy, om, prof=pr), 2 * G * 16384 * 704 * 352, 6)
xx = t(1, 16, 224, 224, 128); ww = t(128, 27 * 128) * 0.02; bb = torch.zeros(128, device=dev)
w2 = t(2, 128, dtype=torch.float32); b2 = torch.zeros(2, device=dev); oo = torch.empty(1, 2, 16, 224, 224, device=dev)
report("head conv 224", lambda pr: ops.conv3d(xx, ww, ksize=(3, 3, 3), bias=bb, head_w2=w2, head_b2=b2, out_f32=oo, prof=pr), 2 * 16 * 224 * 224 * 27 * 128 * 128, 54)
x3 = t(1, 16, 128, 128, 256); w3 = t(128, 27 * 256) * 0.02; o3 = torch.empty(1, 16, 128, 128, 128, device=dev, dtype=dt)
report("conv 128^2 256->128", lambda pr: ops.conv3d(x3, w3, ksize=(3, 3, 3), bias=bb, out_16=o3, prof=pr), 2 * 16 * 128 * 128 * 27 * 256 * 128, 108)
