"""In-run A/B: 1-CTA vs CTA-pair kernel for the M=2048 encoder GEMMs with N=1408 (64 pair tiles < 74 pairs)."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
r32 = torch.randn(2048, 1408, device="cuda")
for name, K in (("proj", 1408), ("fc2", 6144)):
    x = torch.randn(2048, K, device="cuda", dtype=dt); w = torch.randn(1408, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.zeros(1408, device="cuda")
    for rep in range(2):
        for pair in (-1, 1):
            us = timeit(lambda: ops.linear(x, w, bias=b, res_f32=r32, out_f32=r32, cta_pair=pair))
            print(f"{name} K={K} cta_pair={pair}: {us:.1f} us  {2*2048*1408*K/us/1e6:.0f} TF/s")
