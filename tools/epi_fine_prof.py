"""Per-phase clock64 stamps of epilogue thread 0 (CTA 0) in the store-mode epilogue (tuning build: L4P_BUILD_TAG=fine
L4P_NVCC_EXTRA=-DL4P_GEMM_FINE_PROF=1 python -m l4p_b200.build; L4P_LIB=l4p_b200/libl4p_b200_fine.so python tools/epi_fine_prof.py)."""
import ctypes as C
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
    # stamps of epilogue thread 0 (warpgroup 0): per tile one "acc ready", then four per 32-column chunk it owns (every second chunk)
    d = lib.GemmDesc()
    d.a = d.w = 0x10000
    d.M, d.N, d.K, d.lda, d.ldw, d.ld_out, d.ld_res = M, N, K, K, K, N, N
    d.a_mode, d.store_mode, d.out_16 = lib.A_MATRIX, lib.STORE_ROWMAJOR, 0x10000
    plan = (C.c_int * 6)()
    assert lib.load().l4p_gemm_plan(C.byref(d), plan) == 0
    bn = plan[0]
    per_tile = ((bn + 31) // 32 + 1) // 2
    i = 0
    while i + 1 + 4 * per_tile <= len(fine):
        ready = fine[i]
        chunks = [fine[i + 1 + 4 * c:i + 5 + 4 * c] for c in range(per_tile)]
        i += 1 + 4 * per_tile
        prev = [ready] + [c[3] for c in chunks[:-1]]
        print(f"  acc ready {ready} (block_n {bn}): " + " | ".join(f"ld+{c[0] - p0} st+{c[1] - c[0]} rd+{c[2] - c[1]} out+{c[3] - c[2]}"
                                                                  for p0, c in zip(prev, chunks)))


run("proj res32", 1408, 1408, res=True)
run("fc2 res32", 1408, 6144, res=True)
run("fc1 gelu", 6144, 1408, act=lib.ACT_GELU)
run("plain out16", 1408, 1408)
