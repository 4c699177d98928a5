"""In-run A/B of the QKV GEMM N-tile width (auto vs 240 vs 256), two repetitions."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import ops
dt = torch.float16
def timeit(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
a2 = torch.randn(2048, 1408, device="cuda", dtype=dt); w5 = torch.randn(4224, 1408, device="cuda", dtype=dt) * 0.03; b5 = torch.zeros(4224, device="cuda")
q5 = torch.zeros(1, 16, 2048, 96, device="cuda", dtype=dt); k5 = torch.zeros_like(q5); v5 = torch.zeros(1, 16, 96, 2048, device="cuda", dtype=dt)
import ctypes as C
from l4p_b200 import lib
d = lib.GemmDesc(); d.a = d.w = 0x10000; d.M, d.N, d.K, d.lda, d.ldw = 2048, 4224, 1408, 1408, 1408
d.store_mode = lib.STORE_QKV; d.q = d.k = d.vt = 0x10000; d.heads, d.head_dim, d.head_dim_pad, d.tokens = 16, 88, 96, 2048
out = (C.c_int * 6)(); print('plan rc', lib.load().l4p_gemm_plan(C.byref(d), out), list(out))
for rep in range(2):
    for bn in (256, 240, 0, 240, 0, 256):
        us = timeit(lambda: ops.linear_qkv(a2, w5, b5, q5, k5, v5, 16, 88, 2048, block_n=bn))
        print(f"qkv block_n={bn or 'auto'}: {us:.1f} us  {2*2048*4224*1408/us/1e6:.0f} TF/s")
