"""Two launches of the encoder-shaped attention kernel inside a cudaProfiler range (ncu --profile-from-start off)."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dt = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float16
q = torch.randn(B, 16, 2048, 96, device="cuda", dtype=dt); q[..., 88:] = 0
k = torch.randn_like(q); k[..., 88:] = 0
vt = torch.randn(B, 16, 96, 2048, device="cuda", dtype=dt); vt[:, :, 88:] = 0
o = torch.empty(B * 2048, 1408, device="cuda", dtype=dt)
for _ in range(3):
    ops.attention(q, k, vt, o, 88, 88 ** -0.5)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    ops.attention(q, k, vt, o, 88, 88 ** -0.5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
