"""One launch each of the four encoder GEMM shapes at one window (M = 2048) and at eight (M = 16384) inside a cudaProfiler range
(ncu --set full --clock-control none --profile-from-start off): tensor-pipe share after the two-warp producer / 128-wide K stages."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt, dev = torch.float16, "cuda"
cases = []
for M in (2048, 16384):
    x14 = torch.randn(M, 1408, device=dev, dtype=dt); x61 = torch.randn(M, 6144, device=dev, dtype=dt)
    w_fc1 = torch.randn(6144, 1408, device=dev, dtype=dt) * 0.03; w_fc2 = torch.randn(1408, 6144, device=dev, dtype=dt) * 0.01
    w_pr = torch.randn(1408, 1408, device=dev, dtype=dt) * 0.03; w_qkv = torch.randn(4224, 1408, device=dev, dtype=dt) * 0.03
    b61 = torch.zeros(6144, device=dev); b14 = torch.zeros(1408, device=dev); b42 = torch.zeros(4224, device=dev)
    r32 = torch.randn(M, 1408, device=dev); h = torch.empty(M, 6144, device=dev, dtype=dt)
    B = M // 2048
    q = torch.zeros(B, 16, 2048, 96, device=dev, dtype=dt); k = torch.zeros_like(q); vt = torch.zeros(B, 16, 96, 2048, device=dev, dtype=dt)
    cases.append((x14, x61, w_fc1, w_fc2, w_pr, w_qkv, b61, b14, b42, r32, h, q, k, vt))


def run():
    for x14, x61, w_fc1, w_fc2, w_pr, w_qkv, b61, b14, b42, r32, h, q, k, vt in cases:
        ops.linear_qkv(x14, w_qkv, b42, q, k, vt, 16, 88, 2048)
        ops.linear(x14, w_pr, bias=b14, res_f32=r32, out_f32=r32)
        ops.linear(x14, w_fc1, bias=b61, act=lib.ACT_GELU, out_16=h)
        ops.linear(x61, w_fc2, bias=b14, res_f32=r32, out_f32=r32)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
