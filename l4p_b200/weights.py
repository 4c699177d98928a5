"""Deterministic synthetic weights, `w = f(seed, state_dict_key, shape)`.

There is no network for the released checkpoint, so benchmarks and parity tests use random weights of the
reference architecture. The generator is keyed on the *state-dict key*, so the reference model, the oracle
and this package get bit-identical tensors without sharing any initialisation code.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import torch


PEAKY_HEATMAP = 8.0


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int = 0, peaky: bool = False) -> torch.Tensor:
    g = _gen(seed, key)
    shape = tuple(shape)
    if "positional_encoding_gaussian_matrix" in key:
        return torch.randn(shape, generator=g)
    if len(shape) >= 2:
        recept = 1
        for s in shape[2:]:
            recept *= s
        fan_in, fan_out = shape[1] * recept, shape[0] * recept
        if len(shape) == 2:
            # Linear / Embedding: Xavier-uniform, the encoder's own init (modeling_pretrain.py:106-113)
            a = math.sqrt(6.0 / (fan_in + fan_out))
        else:
            # Conv3d / ConvTranspose3d: torch's default kaiming_uniform(a=sqrt(5)) bound = 1/sqrt(fan_in), which is
            # what the reference's DPT / SAM convolutions are constructed with (no custom init in dpt_block.py)
            a = 1.0 / math.sqrt(fan_in)
        w = (torch.rand(shape, generator=g) * 2 - 1) * a
        if peaky and key.endswith("attn.qkv.weight"):
            w *= 4.0
        if peaky and key.endswith("output_hypernetworks_mlps.0.layers.2.weight"):
            # the track head's trajectory heat-map logits x PEAKY_HEATMAP: with reference-style initialisation the logits have
            # a standard deviation of ~1.3 over 16x64x64 cells, i.e. a near-uniform soft-argmax that sits at the image centre
            # for every query; x8 makes the expectation follow the logit maxima across the image
            w *= PEAKY_HEATMAP
        return w
    u = torch.rand(shape, generator=g) * 2 - 1
    if key.endswith("weight"):  # norm scales
        return 1.0 + 0.1 * u
    return 0.02 * u  # biases


def synth_state_dict(manifest: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0,
                     peaky: bool = False) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed, peaky) for k, s in manifest}


def fill_module_(module: torch.nn.Module, seed: int = 0, peaky: bool = False, prefix: str = "") -> None:
    """In-place synthetic init of every state-dict entry (CPU generator -> copied to the tensor's device)."""
    with torch.no_grad():
        for k, t in module.state_dict().items():
            t.copy_(synth_tensor(prefix + k, tuple(t.shape), seed, peaky).to(t.device))


def fill_module_fast_(module: torch.nn.Module, seed: int = 0) -> None:
    """Same distributions as `fill_module_`, generated on the tensor's own device (benchmarks: the values
    need not be reproducible across devices, 1.4 B parameters should not take a minute of CPU RNG)."""
    with torch.no_grad():
        for i, (k, t) in enumerate(module.state_dict().items()):
            g = torch.Generator(device=t.device)
            g.manual_seed((zlib.crc32(k.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
            shape = tuple(t.shape)
            if "positional_encoding_gaussian_matrix" in k:
                t.copy_(torch.randn(shape, generator=g, device=t.device))
            elif len(shape) >= 2:
                recept = 1
                for s in shape[2:]:
                    recept *= s
                fan_in, fan_out = shape[1] * recept, shape[0] * recept
                a = math.sqrt(6.0 / (fan_in + fan_out)) if len(shape) == 2 else 1.0 / math.sqrt(fan_in)
                t.copy_((torch.rand(shape, generator=g, device=t.device) * 2 - 1) * a)
            else:
                u = torch.rand(shape, generator=g, device=t.device) * 2 - 1
                t.copy_(1.0 + 0.1 * u if k.endswith("weight") else 0.02 * u)
