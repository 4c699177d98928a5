import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import ops
dt = torch.float16
x = torch.randn(1, 16, 224, 224, 128, device="cuda", dtype=dt)
w = torch.randn(128, 27 * 128, device="cuda", dtype=dt) * 0.02
o = torch.empty(1, 16, 224, 224, 128, device="cuda", dtype=dt)
for _ in range(2):
    ops.conv3d(x, w, ksize=(3, 3, 3), out_16=o)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.conv3d(x, w, ksize=(3, 3, 3), out_16=o)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
