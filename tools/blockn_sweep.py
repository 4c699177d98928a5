"""N-tile width sweep for the big track-head projections (M = G*2048 rows, N = 704) and the QKV GEMM."""
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
M, K = 32 * 2048, 1408
x = torch.randn(M, K, device="cuda", dtype=dt)
for N in (704, 1408):
    w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.zeros(N, device="cuda"); o = torch.empty(M, N, device="cuda", dtype=dt)
    for bn in (0, 176, 192, 224, 240, 256):
        if bn and bn > N: continue
        us = timeit(lambda: ops.linear(x, w, bias=b, out_16=o, block_n=bn))
        print(f"linear M={M} N={N} K={K} block_n={bn or 'auto'}: {us:.1f} us  {2*M*N*K/us/1e6:.0f} TF/s")
a2 = torch.randn(2048, 1408, device="cuda", dtype=dt); w5 = torch.randn(4224, 1408, device="cuda", dtype=dt) * 0.03; b5 = torch.zeros(4224, device="cuda")
q5 = torch.zeros(1, 16, 2048, 96, device="cuda", dtype=dt); k5 = torch.zeros_like(q5); v5 = torch.zeros(1, 16, 96, 2048, device="cuda", dtype=dt)
for bn in (0, 192, 256):
    us = timeit(lambda: ops.linear_qkv(a2, w5, b5, q5, k5, v5, 16, 88, 2048, block_n=bn), 20)
    print(f"qkv M=2048 N=4224 block_n={bn or 'auto'}: {us:.1f} us  {2*2048*4224*1408/us/1e6:.0f} TF/s")
# encoder shapes at one clip (M = 2048)
xs = {1408: torch.randn(2048, 1408, device="cuda", dtype=dt), 6144: torch.randn(2048, 6144, device="cuda", dtype=dt)}
r32 = torch.randn(2048, 1408, device="cuda")
for name, N, K, kw, bns in (("proj", 1408, 1408, dict(res_f32=r32, out_f32=r32), (0, 160, 176, 240, 256)),
                            ("fc2", 1408, 6144, dict(res_f32=r32, out_f32=r32), (0, 160, 176, 240, 256)),
                            ("fc1", 6144, 1408, dict(act=lib.ACT_GELU, out_16=torch.empty(2048, 6144, device="cuda", dtype=dt)), (0, 192, 240, 256))):
    w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.zeros(N, device="cuda")
    for bn in bns:
        us = timeit(lambda: ops.linear(xs[K], w, bias=b, block_n=bn, **kw), 20)
        print(f"{name} M=2048 N={N} K={K} block_n={bn or 'auto'}: {us:.1f} us  {2*2048*N*K/us/1e6:.0f} TF/s")
