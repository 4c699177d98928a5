import sys, torch
sys.path.insert(0, ".")
from l4p_b200 import ops
x = torch.randn(128 * 16 * 32 * 32, 352, device="cuda", dtype=torch.float16)
y = torch.empty_like(x)
g = torch.ones(352, device="cuda"); b = torch.zeros(352, device="cuda")
for _ in range(3): ops.layernorm16(x, g, b, 1e-6, y, gelu=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.layernorm16(x, g, b, 1e-6, y, gelu=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"LN3d+GELU 2M x 352: {ms*1e3:.1f} us, {x.numel()*4/ms/1e9:.2f} TB/s")
