"""Phase timeline of the attention kernel (CTA 0): clock64 stamps per KV block for softmax WG0/WG1 and the MMA thread."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import ops
B = 1
q = torch.randn(B, 16, 2048, 96, device="cuda", dtype=torch.float16)
k = torch.randn_like(q); vt = torch.randn(B, 16, 96, 2048, device="cuda", dtype=torch.float16)
o = torch.empty(B * 2048, 1408, device="cuda", dtype=torch.float16)
GRID = B * 16 * 8
prof = torch.zeros(3 * 64 * 8 + 5 * GRID, device="cuda", dtype=torch.int64)
for _ in range(3):
    ops.attention(q, k, vt, o, 88, 88 ** -0.5, prof=prof)
torch.cuda.synchronize()
pc = prof.cpu()
p = pc[:1536].view(3, 64, 8)
cta = pc[1536:].view(GRID, 5)
t0 = int(p[0, 0, 0])
names = ["wait_s", "got_s", "ldtm", "max", "exp", "pvfree", "pstored"]
for role in (0, 1):
    print(f"softmax WG{role}: iteration start and phase durations (cycles)")
    for j in range(16):
        st = [int(x) - t0 for x in p[role, j, :7]]
        print(f"  j={j:2d} start {st[0]:7d} | wait_s {st[1]-st[0]:5d} ldtm {st[2]-st[1]:5d} max {st[3]-st[2]:5d} exp {st[4]-st[3]:5d} wait_pv {st[5]-st[4]:5d} store {st[6]-st[5]:5d} | total {st[6]-st[0]:6d}")
print("MMA thread: per j: [t0: before pfull wait, after, after issue_pv] [t1: ...]")
for j in range(16):
    st = [int(x) - t0 for x in p[2, j, :7]]
    print(f"  j={j:2d} t0: wait@{st[0]:7d} +{st[1]-st[0]:5d} issue +{st[2]-st[1]:4d} | t1: wait@{st[4]:7d} +{st[5]-st[4]:5d} issue +{st[6]-st[5]:4d}")

g0 = int(cta[:, 0].min())
dur_ns = (cta[:, 1] - cta[:, 0]).float()
dur_cy = (cta[:, 3] - cta[:, 2]).float()
print(f"per-CTA: wall ns min/med/max {dur_ns.min():.0f}/{dur_ns.median():.0f}/{dur_ns.max():.0f}; cycles min/med/max {dur_cy.min():.0f}/{dur_cy.median():.0f}/{dur_cy.max():.0f}; "
      f"MHz med {(dur_cy / dur_ns * 1e3).median():.0f}; start spread ns {int(cta[:, 0].max()) - g0}; kernel span ns {int(cta[:, 1].max()) - g0}; distinct SMs {len(set(cta[:, 4].tolist()))}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.attention(q, k, vt, o, 88, 88 ** -0.5)
e1.record(); torch.cuda.synchronize()
print(f"back-to-back launches: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us each")
