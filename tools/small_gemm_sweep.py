"""Token-side GEMMs of the track head (M = 768 rows = 128 queries x 6 tokens): time per launch inside a dependent chain
(CUDA graph of 40 launches) for the automatic tiling (split-K + finalize where few tiles) and explicit N-tile widths."""
import sys, torch
sys.path.insert(0, ".")
from l4p_b200 import ops
dev, dt = "cuda", torch.float16


def timeit(f, n=40):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f(); f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 / n * 1e3


for (M, N, K) in [(768, 1408, 1408), (768, 704, 1408), (768, 1408, 704), (768, 2048, 1408), (768, 1408, 2048), (128, 1408, 1408),
                  (6144, 1408, 704), (6144, 2048, 704), (6144, 704, 1408),
                  (2048, 256, 1408), (2048, 512, 1408), (2048, 1024, 1408), (2048, 704, 1408), (4096, 256, 256), (16384, 256, 256)]:
    x = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) * 0.03).to(dt)
    b = torch.zeros(N, device=dev); y = torch.empty(M, N, device=dev)
    row = [f"M={M} N={N} K={K}:"]
    for bn in (0, 32, 64, 96, 128, 176, 256):
        if bn and (bn > N):
            continue
        try:
            t = timeit(lambda: ops.linear(x, w, bias=b, out_f32=y, **(dict(block_n=bn) if bn else {})))
            row.append(f"bn={bn or 'auto'} {t:.1f}us")
        except Exception as e:  # noqa: BLE001
            row.append(f"bn={bn} err")
    print("  ".join(row))
