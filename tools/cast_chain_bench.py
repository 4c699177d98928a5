import sys, torch
sys.path.insert(0, ".")
from l4p_b200 import ops
dev="cuda"; dt=torch.float16
x32 = torch.randn(768, 1408, device=dev); x16 = torch.empty(768, 1408, device=dev, dtype=dt)
w = (torch.randn(1408, 1408, device=dev)*0.03).to(dt); b = torch.zeros(1408, device=dev)
y32 = torch.empty(768, 1408, device=dev)
def chain(with_cast, n=40):
    for _ in range(n):
        if with_cast: ops.cast16(x32, x16)
        ops.linear(x16, w, bias=b, out_f32=y32)
def timeit(name, f):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f(); f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        f()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/10/40*1e3:.2f} us per iteration")
timeit("linear only", lambda: chain(False))
timeit("cast16 + linear", lambda: chain(True))
def casts(n=40):
    for _ in range(n): ops.cast16(x32, x16)
timeit("cast16 only", casts)
