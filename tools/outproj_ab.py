"""Track-head token stream: the out-projection (M = 128 queries x 2048 tokens, N = 1408, K = 704) under every residual /
output combination, and the LayerNorm that follows it, timed back to back (CUDA events around 10 launches).
Bytes: fp32 stream = 1.48 GB per tensor, 16-bit stream = 0.74 GB."""
import sys

import torch

sys.path.insert(0, ".")
from l4p_b200 import ops  # noqa: E402

dev = "cuda"
dt = torch.float16
M, N, K = 128 * 2048, 1408, 704
a = torch.randn(M, K, device=dev, dtype=dt)
w = torch.randn(N, K, device=dev, dtype=dt) * 0.03
b = torch.zeros(N, device=dev)
r32 = torch.randn(M, N, device=dev)
r16 = torch.randn(M, N, device=dev, dtype=dt)
rtab = torch.randn(2048, N, device=dev)
o32 = torch.empty(M, N, device=dev)
o16 = torch.empty(M, N, device=dev, dtype=dt)
g = torch.ones(N, device=dev)
be = torch.zeros(N, device=dev)
y16 = torch.empty(M, N, device=dev, dtype=dt)


def timeit(name, f, gb, flops=0.0, n=10):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:58s} {ms * 1e3:8.1f} us  {gb / ms:7.2f} TB/s" + (f"  {flops / ms / 1e9:7.1f} TF/s" if flops else ""))


F = 2.0 * M * N * K
A = M * K * 2 / 1e9
S32, S16 = M * N * 4 / 1e9, M * N * 2 / 1e9
for bn in (0, 176, 256, 128):
    tag = f"[block_n={bn or 'auto'}] "
    kw = dict(block_n=bn) if bn else {}
    try:
        timeit(tag + "res_f32 -> out_f32", lambda: ops.linear(a, w, bias=b, res_f32=r32, out_f32=o32, **kw), A + 2 * S32, F)
        timeit(tag + "res_f32(table) -> out_f32", lambda: ops.linear(a, w, bias=b, res_f32=rtab, res_row_mod=2048, out_f32=o32, **kw), A + S32, F)
        timeit(tag + "res_f32 -> out_16", lambda: ops.linear(a, w, bias=b, res_f32=r32, out_16=o16, **kw), A + S32 + S16, F)
        timeit(tag + "res_f32(table) -> out_16", lambda: ops.linear(a, w, bias=b, res_f32=rtab, res_row_mod=2048, out_16=o16, **kw), A + S16, F)
        timeit(tag + "res_16 -> out_16", lambda: ops.linear(a, w, bias=b, res_16=r16, out_16=o16, **kw), A + 2 * S16, F)
        timeit(tag + "none -> out_16", lambda: ops.linear(a, w, bias=b, out_16=o16, **kw), A + S16, F)
        timeit(tag + "none -> out_f32", lambda: ops.linear(a, w, bias=b, out_f32=o32, **kw), A + S32, F)
    except TypeError as e:
        print("block_n override not supported by ops.linear:", e)
        break
timeit("layernorm fp32 -> 16", lambda: ops.layernorm(r32, g, be, 1e-5, out16=y16), S32 + S16)
timeit("layernorm16 16 -> 16", lambda: ops.layernorm16(r16, g, be, 1e-5, y16), 2 * S16)
x2 = torch.randn(2048, N, device=dev)
y2 = torch.empty(2048, N, device=dev, dtype=dt)
timeit("layernorm 2048 rows fp32 -> 16 (back to back, PDL)", lambda: ops.layernorm(x2, g, be, 1e-6, out16=y2), 2048 * N * 6 / 1e9, n=40)
