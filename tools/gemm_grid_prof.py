"""Whole-grid timeline of one GEMM launch inside a back-to-back chain (experiment build with -DL4P_GEMM_GRID_PROF=1:
L4P_BUILD_TAG=gprof L4P_NVCC_EXTRA=-DL4P_GEMM_GRID_PROF=1 python -m l4p_b200.build; L4P_LIB=l4p_b200/libl4p_b200_gprof.so python tools/gemm_grid_prof.py).
Every CTA stamps the global timer at kernel entry, after the PDL wait and at its end."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16


def run(name, M, N, K, **kw):
    x = torch.randn(M, K, device="cuda", dtype=dt); w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    bias = torch.zeros(N, device="cuda")
    r32 = torch.randn(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=dt)
    out = dict(res_f32=r32, out_f32=r32) if kw.pop("res", False) else dict(out_16=o16)
    profs = [torch.zeros(2048, device="cuda", dtype=torch.int64) for _ in range(6)]
    for _ in range(3):
        ops.linear(x, w, bias=bias, **out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for pr in profs:
        ops.linear(x, w, bias=bias, prof=pr, **out, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"== {name} M={M} N={N} K={K}: {e0.elapsed_time(e1) / len(profs) * 1e3:.1f} us per launch (6 back to back)")
    g = [p.cpu()[1536:1536 + 3 * 148].view(-1, 3) for p in profs]
    t00 = None
    for i, t in enumerate(g):
        t = t[t[:, 2] > 0]
        if t00 is None: t00 = int(t[:, 0].min())
        entry, start, end = t[:, 0] - t00, t[:, 1] - t00, t[:, 2] - t00
        print(f"  launch {i}: CTAs {t.shape[0]:3d}  entry {int(entry.min()):7d}..{int(entry.max()):7d} ns  after PDL wait {int(start.min()):7d}..{int(start.max()):7d}"
              f"  end {int(end.min()):7d}..{int(end.max()):7d}  (CTA busy median {int((end - start).median())} ns, max {int((end - start).max())})")


run("proj (res32 in place)", 2048, 1408, 1408, res=True)
run("fc2 (res32 in place)", 2048, 1408, 6144, res=True)
run("fc1 gelu", 2048, 6144, 1408, act=lib.ACT_GELU)
run("qkv-like", 2048, 4224, 1408)
