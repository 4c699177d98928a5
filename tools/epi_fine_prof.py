"""Per-phase clock64 stamps of epilogue thread 0 (CTA 0) in the store-mode epilogue (tuning build: L4P_BUILD_TAG=fine
L4P_NVCC_EXTRA=-DL4P_GEMM_FINE_PROF=1 python -m l4p_b200.build; L4P_LIB=l4p_b200/libl4p_b200_fine.so python tools/epi_fine_prof.py)."""
import sys
import torch
sys.path.insert(0, ".")
from l4p_b200 import lib, ops
dt = torch.float16
M = 2048


def run(name, N, K, **kw):
    x = torch.randn(M, K, device="cuda", dtype=dt); w = torch.randn(N, K, device="cuda", dtype=dt) * K ** -0.5
    b = torch.zeros(N, device="cuda")
    r32 = torch.randn(M, N, device="cuda")
    out = dict(res_f32=r32, out_f32=r32) if kw.pop("res", False) else dict(out_16=torch.empty(M, N, device="cuda", dtype=dt))
    prof = torch.zeros(3 * 512, device="cuda", dtype=torch.int64)
    for _ in range(3):
        prof.zero_()
        ops.linear(x, w, bias=b, prof=prof, **out, **kw)
    torch.cuda.synchronize()
    p = prof.cpu().view(3, 512)
    t0 = int(p[2, 511])
    fine = [int(v) - t0 for v in p[2, 64:64 + 200] if v > 0]
    mma = [int(v) - t0 for v in p[1] if v > 0]
    print(f"== {name} N={N} K={K}: last stage issued {mma[-1]}")
    # stamps: acc-ready, then per chunk (tmem loaded, staged, staging read, stored)
    i = 0
    while i < len(fine):
        ready = fine[i]; i += 1
        chunks = []
        while i + 3 < len(fine) + 1 and len(chunks) < 8 and i + 3 <= len(fine):
            c = fine[i:i + 4]
            if len(c) < 4: break
            chunks.append(c); i += 4
            if i < len(fine) and fine[i] - c[3] > 3000: break   # next tile
        print(f"  acc ready {ready}: " + " | ".join(f"ld+{c[0] - (prev if prev else ready)} st+{c[1] - c[0]} rd+{c[2] - c[1]} out+{c[3] - c[2]}"
                                                  for prev, c in zip([None] + [cc[3] for cc in chunks[:-1]], chunks)))


run("proj res32", 1408, 1408, res=True)
run("fc2 res32", 1408, 6144, res=True)
run("fc1 gelu", 6144, 1408, act=lib.ACT_GELU)
run("plain out16", 1408, 1408)
