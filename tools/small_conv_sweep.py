"""Low-resolution DPT convolutions (grouped x3 as the dense heads run them): automatic tiling (split-K + finalize) vs explicit
N-tile widths without split-K, time per launch inside a dependent chain (CUDA graph of 20 launches)."""
import sys, torch
sys.path.insert(0, ".")
from l4p_b200 import ops
dev, dt = "cuda", torch.float16


def timeit(f, n=20):
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


G = 3
for (T, H, W, Cin, Cout) in [(4, 8, 8, 256, 256), (4, 8, 8, 1024, 256), (8, 16, 16, 256, 256), (8, 16, 16, 1024, 256), (16, 16, 16, 256, 256),
                             (16, 32, 32, 256, 256), (16, 32, 32, 512, 256)]:
    x = torch.randn(G, T, H, W, Cin, device=dev).to(dt)
    w = (torch.randn(G * Cout, 27 * Cin, device=dev) * 0.02).to(dt)
    b = torch.zeros(G * Cout, device=dev)
    y = torch.empty(G, T, H, W, Cout, device=dev, dtype=dt)
    fl = 2.0 * G * T * H * W * Cout * 27 * Cin
    row = [f"{T}x{H}x{W} {Cin}->{Cout} (x{G}):"]
    for bn in (0, 64, 128, 256):
        t = timeit(lambda: ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, out_16=y, groups=G, **(dict(block_n=bn) if bn else {})))
        row.append(f"bn={bn or 'auto'} {t:.1f}us ({fl / t / 1e6:.0f} TF/s)")
    for bn in (128, 256):   # the 2-CTA kernel forced (line-halo stages) where the tile count allows it
        try:
            t = timeit(lambda: ops.conv3d(x, w, ksize=(3, 3, 3), bias=b, out_16=y, groups=G, block_n=bn, cta_pair=1))
            row.append(f"pair bn={bn} {t:.1f}us")
        except Exception:  # noqa: BLE001
            row.append(f"pair bn={bn} n/a")
    print("  ".join(row))
