"""EXPERIMENT: start the epilogue warpgroups / warps of a tile staggered (L4P_EPI_STAGGER_G / _W cycles) so that their TMEM-read,
smem-transposition and global-store phases overlap instead of colliding. Epilogue-bound cases: K = 48 output GEMM (whole kernel),
fc1 / qkv / proj / fc2 (exposed last tile), big-M out16 GEMM."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16


def t(name, f, n=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us")


G, P, N, K = 128, 2048, 1408, 48
a = torch.softmax(torch.randn(G * P, K, device="cuda"), -1).to(dt)
w = (torch.randn(G * N, K, device="cuda") * 0.1).to(dt)
b = torch.zeros(N, device="cuda")
r16 = torch.randn(G * P, N, device="cuda", dtype=dt)
o = torch.empty(G * P, N, device="cuda", dtype=dt)
t("K=48 grouped GEMM res_16 -> out_16", lambda: ops.linear(a, w, bias=b, res_16=r16, out_16=o, group_rows=P))
del a, w, r16, o
x = torch.randn(32 * 2048, 1408, device="cuda", dtype=dt); w2 = torch.randn(1408, 1408, device="cuda", dtype=dt) * 0.03
o2 = torch.empty(32 * 2048, 1408, device="cuda", dtype=dt)
t("M=65536 N=1408 K=1408 out_16", lambda: ops.linear(x, w2, bias=b, out_16=o2))
# exposed last-tile epilogues: clock64 timeline of CTA 0
M = 2048
def tail(name, N, K, **kw):
    xx = torch.randn(M, K, device="cuda", dtype=dt); ww = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    bb = torch.zeros(N, device="cuda")
    r32 = torch.randn(M, N, device="cuda")
    out = dict(res_f32=r32, out_f32=r32) if kw.pop("res", False) else dict(out_16=torch.empty(M, N, device="cuda", dtype=dt))
    prof = torch.zeros(3 * 512, device="cuda", dtype=torch.int64)
    for _ in range(3):
        ops.linear(xx, ww, bias=bb, prof=prof, **out, **kw)
    torch.cuda.synchronize()
    p = prof.cpu().view(3, 512)
    t0 = int(p[2, 511])
    mma = [int(v) - t0 for v in p[1] if v > 0]
    epi = [int(v) - t0 for v in p[2, :16] if v > 0]
    print(f"{name:12s} last stage issued {mma[-1]:6d}  kernel end {epi[-1]:6d}  tail {epi[-1] - mma[-1]:6d} cycles")
tail("fc1 gelu", 6144, 1408, act=lib.ACT_GELU)
tail("qkv-like", 4224, 1408)
tail("proj res32", 1408, 1408, res=True)
tail("fc2 res32", 1408, 6144, res=True)
