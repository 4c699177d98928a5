"""Run a few representative GEMM configurations inside a cudaProfiler range (for `ncu --profile-from-start off`)."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16
dev = "cuda"
def t(*s, dtype=dt): return torch.randn(*s, device=dev, dtype=dtype)
G = 16
cases = {}
# hyper convT: [G,16,32,32,352] -> taps (1,2,2) x 176
x = t(G, 16, 32, 32, 352); w = t(4 * 176, 352) * 0.05; b = torch.zeros(4 * 176, device=dev); hy = t(G, 3, 176, dtype=torch.float32)
om = torch.empty(G, 3, 16, 64, 64, device=dev)
cases["hyper"] = lambda: ops.conv_transpose3d_hyper(x, w, b, (1, 2, 2), hy, om)
a1 = t(G * 2048, 704); w1 = t(1408, 704) * 0.03; b1 = torch.zeros(1408, device=dev); r1 = t(G * 2048, 1408, dtype=torch.float32); o1 = torch.empty_like(r1)
cases["outproj_res32"] = lambda: ops.linear(a1, w1, bias=b1, res_f32=r1, out_f32=o1)
a2 = t(2048, 1408); w2 = t(1408, 1408) * 0.03; r2 = t(2048, 1408, dtype=torch.float32)
cases["proj"] = lambda: ops.linear(a2, w2, bias=b1, res_f32=r2, out_f32=r2)
w3 = t(6144, 1408) * 0.03; b3 = torch.zeros(6144, device=dev); o3 = torch.empty(2048, 6144, device=dev, dtype=dt)
cases["fc1"] = lambda: ops.linear(a2, w3, bias=b3, act=lib.ACT_GELU, out_16=o3)
a4 = t(G * 2048, 1408); w4 = t(704, 1408) * 0.03; b4 = torch.zeros(704, device=dev); o4 = torch.empty(G * 2048, 704, device=dev, dtype=dt)
cases["kproj"] = lambda: ops.linear(a4, w4, bias=b4, out_16=o4)
w5 = t(4224, 1408) * 0.03; b5 = torch.zeros(4224, device=dev)
q5 = torch.empty(1, 16, 2048, 96, device=dev, dtype=dt); k5 = torch.empty_like(q5); v5 = torch.empty(1, 16, 96, 2048, device=dev, dtype=dt)
cases["qkv"] = lambda: ops.linear_qkv(a2, w5, b5, q5, k5, v5, 16, 88, 2048)
a6 = t(2048, 6144); w6 = t(1408, 6144) * 0.02
cases["fc2"] = lambda: ops.linear(a6, w6, bias=b1, res_f32=r2, out_f32=r2)
for f in cases.values():
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for n, f in cases.items():
    f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("order:", list(cases))
