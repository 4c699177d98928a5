"""Timeline of CTA 0 for the track head's per-query K = 48 output GEMM (M = 262144, N = 1408, grouped weights, 16-bit
residual and output): where does a tile's time go when the main loop is one k-block?"""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import ops
dt = torch.float16
G, P, N, K = 128, 2048, 1408, 48
a = torch.softmax(torch.randn(G * P, K, device="cuda"), -1).to(dt)
w = (torch.randn(G * N, K, device="cuda") * 0.1).to(dt)
b = torch.zeros(N, device="cuda")
r16 = torch.randn(G * P, N, device="cuda", dtype=dt)
tab = torch.randn(P, N, device="cuda")
o = torch.empty(G * P, N, device="cuda", dtype=dt)
prof = torch.zeros(3 * 512, device="cuda", dtype=torch.int64)


def t(name, f, n=10):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us")


for bn in (0, 256, 128):
    kw = dict(block_n=bn) if bn else {}
    t(f"[bn={bn or 'auto'}] res_16 -> out_16", lambda: ops.linear(a, w, bias=b, res_16=r16, out_16=o, group_rows=P, **kw))
    t(f"[bn={bn or 'auto'}] res_f32 table -> out_16", lambda: ops.linear(a, w, bias=b, res_f32=tab, res_row_mod=P, out_16=o, group_rows=P, **kw))
    t(f"[bn={bn or 'auto'}] none -> out_16", lambda: ops.linear(a, w, bias=b, out_16=o, group_rows=P, **kw))
ops.linear(a, w, bias=b, res_16=r16, out_16=o, group_rows=P, prof=prof)
torch.cuda.synchronize()
p = prof.cpu().view(3, 512)
t0 = int(p[2, 511])
print("producer issue:", [int(v) - t0 for v in p[0, :12] if v > 0])
print("mma full-wait done:", [int(v) - t0 for v in p[1, :12] if v > 0])
print("epilogue (release, end) per tile:", [int(v) - t0 for v in p[2, :24] if v > 0])
