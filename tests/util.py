"""Shared helpers for parity tests: synthetic inputs and error metrics."""
import torch

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def synth_rgb(B: int, T: int, H: int = 224, W: int = 224, seed: int = 0) -> torch.Tensor:
    """torch.rand(B,3,T,H,W; seed) ImageNet-normalised (l4p_dataset_mini.py:103-104,576-580)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, T, H, W, generator=g)
    m = torch.tensor(MEAN).view(1, 3, 1, 1, 1)
    s = torch.tensor(STD).view(1, 3, 1, 1, 1)
    return (x - m) / s


def synth_intrinsics(B: int, T: int, H: int = 224, W: int = 224) -> torch.Tensor:
    """Dummy pinhole fx=fy=min(H,W), cx=W/2, cy=H/2 repeated over T (video_dataset.py:113-127) -> [B,4,4,T]."""
    k = torch.eye(4)
    k[0, 0] = k[1, 1] = float(min(H, W))
    k[0, 2], k[1, 2] = W / 2.0, H / 2.0
    return k[None, :, :, None].repeat(B, 1, 1, T)


def grid_queries(n_side: int, t: float = 0.5, H: int = 224, W: int = 224) -> torch.Tensor:
    """Uniform grid of track queries (t+0.5, x+0.5, y+0.5) at frame 0 (l4p_dataset_mini.py:440-490) -> [1,N,3]."""
    xs = torch.linspace(8, W - 8, n_side)
    ys = torch.linspace(8, H - 8, n_side)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    q = torch.stack([torch.full_like(gx, t), gx + 0.5, gy + 0.5], dim=-1).reshape(1, -1, 3)
    return q


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def max_rel(a: torch.Tensor, b: torch.Tensor, floor: float = 1e-6) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs() / (b.abs() + floor)).max().item()
