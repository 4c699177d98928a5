"""Parameter containers named exactly like the torch modules the reference uses, so that
`state_dict()` / `load_state_dict(strict=True)` are key-compatible with reference checkpoints
(SURVEY.md §5 "checkpoint / resume"). They hold weights only: their forward is never the compute path."""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn


def _empty(*shape, device=None):
    return nn.Parameter(torch.empty(*shape, device=device, dtype=torch.float32), requires_grad=False)


class Linear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = _empty(out_features, in_features, device=device)
        if bias:
            self.bias = _empty(out_features, device=device)
        else:
            self.register_parameter("bias", None)


class LayerNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5, device=None):
        super().__init__()
        self.eps = eps
        self.weight = _empty(dim, device=device)
        self.bias = _empty(dim, device=device)


class Conv3d(nn.Module):
    def __init__(self, cin: int, cout: int, kernel: Sequence[int], stride: Sequence[int] = (1, 1, 1),
                 padding: Sequence[int] = (0, 0, 0), bias: bool = True, device=None):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding = tuple(kernel), tuple(stride), tuple(padding)
        self.weight = _empty(cout, cin, *kernel, device=device)
        if bias:
            self.bias = _empty(cout, device=device)
        else:
            self.register_parameter("bias", None)


class ConvTranspose3d(nn.Module):
    def __init__(self, cin: int, cout: int, kernel: Sequence[int], stride: Sequence[int], device=None):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride = tuple(kernel), tuple(stride)
        self.weight = _empty(cin, cout, *kernel, device=device)
        self.bias = _empty(cout, device=device)


class Embedding(nn.Module):
    def __init__(self, num: int, dim: int, device=None):
        super().__init__()
        self.weight = _empty(num, dim, device=device)


class Identity(nn.Module):
    pass
