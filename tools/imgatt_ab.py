"""A/B at the real size (128 queries x 2048 video tokens): CUDA-core image attention vs the tcgen05 formulation."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import ops
G, Np, nk, H, d = 128, 2048, 6, 8, 88
dt = torch.float16
q16 = torch.randn(G * Np, H * d, device="cuda", dtype=dt)
k = torch.randn(G, nk, H * d, device="cuda"); v = torch.randn(G, nk, H * d, device="cuda")
o0 = torch.empty_like(q16); o1 = torch.empty_like(q16)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
ops.IMGATT_TC = False
t0 = timeit(lambda: ops.image_attention(q16, k, v, o0, G, H, d ** -0.5))
ops.IMGATT_TC = True
t1 = timeit(lambda: ops.image_attention(q16, k, v, o1, G, H, d ** -0.5))
print(f"cuda-core {t0:.0f} us, tcgen05 {t1:.0f} us, max abs diff {(o0.float() - o1.float()).abs().max().item():.3e}")
