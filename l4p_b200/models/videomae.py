"""VideoMAE-v2 ViT-giant video encoder on the B200 kernels.

Mirrors the constructor arguments, parameter names and `forward` contract of the reference's
`PretrainVisionTransformerEncoder` / `Block` / `Attention` / `Mlp` / `PatchEmbed`
(l4p/models/VideoMAEv2/models/modeling_pretrain.py:32-104, modeling_finetune.py:51-69,137-190,193-283)
as wrapped by `VideoMAEEncoder` (l4p/models/l4p_videomae.py:17-122). The modules below only *hold*
parameters; the compute is the fixed kernel sequence in `VideoMAEEncoder.forward`:

    patchify -> GEMM(+bias +pos-embed)                                   K1
    per block:  LN -> GEMM(qkv, head-major scatter) -> fused attention   K2 K3 K4
                -> GEMM(proj +bias +residual) -> LN                      K5 K2
                -> GEMM(fc1 +bias +erf-GELU) -> GEMM(fc2 +bias +residual)  K6
    final LN on the last block output                                    K2

The residual stream and all LayerNorm/softmax statistics are fp32; GEMM operands are fp16 (default) or bf16.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import numpy as np
import torch
from torch import nn

from .. import lib as _l
from .. import ops
from . import params as P


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """tab[p, j] = sin(p / 10000^(2*(j//2)/d)) for even j, cos(.) for odd j, in float64 then cast to fp32
    (modeling_finetune.py:288-299)."""
    j = np.arange(d_hid)
    denom = np.power(10000.0, 2.0 * (j // 2) / d_hid)
    ang = np.arange(n_position, dtype=np.float64)[:, None] / denom[None, :]
    tab = np.empty_like(ang)
    tab[:, 0::2] = np.sin(ang[:, 0::2])
    tab[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.tensor(tab, dtype=torch.float32).unsqueeze(0)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2,
                 device=None):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.tubelet_size = tubelet_size
        self.num_patches = (img_size // patch_size) ** 2 * (num_frames // tubelet_size)
        k = (tubelet_size, patch_size, patch_size)
        self.proj = P.Conv3d(in_chans, embed_dim, k, stride=k, device=device)


class Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias, qk_scale=None, device=None):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.qkv = P.Linear(dim, dim * 3, bias=False, device=device)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.empty(dim, device=device), requires_grad=False)
            self.v_bias = nn.Parameter(torch.empty(dim, device=device), requires_grad=False)
        else:
            self.q_bias = None
            self.v_bias = None
        self.proj = P.Linear(dim, dim, device=device)


class Mlp(nn.Module):
    def __init__(self, dim, hidden, device=None):
        super().__init__()
        self.fc1 = P.Linear(dim, hidden, device=device)
        self.fc2 = P.Linear(hidden, dim, device=device)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, qk_scale, eps, init_values, device=None):
        super().__init__()
        self.norm1 = P.LayerNorm(dim, eps, device=device)
        self.attn = Attention(dim, num_heads, qkv_bias, qk_scale, device=device)
        self.norm2 = P.LayerNorm(dim, eps, device=device)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), device=device)
        if init_values is not None and init_values > 0:
            raise NotImplementedError("layer-scale (init_values > 0) is not used by L4P (l4p_videomae.py:178)")
        self.gamma_1, self.gamma_2 = None, None


class FeatureList(list):
    """The reference's `features_list` plus the kernel-ready 16-bit copies of the materialised entries
    (`taps16[i]`: [B*tokens, C] fp16/bf16) so heads do not re-cast what the encoder already produced."""

    taps16: Dict[int, torch.Tensor]

    def __init__(self, items, taps16):
        super().__init__(items)
        self.taps16 = taps16


def slice_features(feats: "FeatureList", start: int, stop: int) -> "FeatureList":
    """Rows [start, stop) of the batch dimension of a batched FeatureList (views, no copies): lets the heads decode a long
    video's windows in bounded chunks while the encoder result stays one batch."""
    items = [None if f is None else f[start:stop] for f in feats]
    taps = {}
    for k, t in getattr(feats, "taps16", {}).items():
        ntok = feats[k].shape[1]
        taps[k] = t[start * ntok:stop * ntok]
    return FeatureList(items, taps)


def _require_device(x: torch.Tensor) -> None:
    if not x.is_cuda:
        raise _l.L4PError("VideoMAEEncoder.forward needs a CUDA tensor: the l4p_b200 hot path has no CPU fallback")


def _norm_eps(norm_layer) -> float:
    kw = getattr(norm_layer, "keywords", None) or {}
    return float(kw.get("eps", 1e-5))


class VideoMAEEncoder(nn.Module):
    """Drop-in for l4p.models.l4p_videomae.VideoMAEEncoder (same ctor signature, same state-dict keys).

    forward(x[B,3,T,H,W]) returns the reference's `depth+1`-long feature list. Only the entries listed in
    `self.keep_features` (default: every index a shipped head reads, {0,14,21,28,36,40} for depth 40) are
    materialised as fp32 tensors; the other entries are None placeholders so indices stay valid. Set
    `keep_features = "all"` to materialise all 41 like the reference (costs 41 x 11.5 MB per window).
    """

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4, qkv_bias=False, qk_scale=None, drop_rate=0, attn_drop_rate=0, drop_path_rate=0,
                 norm_layer=torch.nn.LayerNorm, init_values=None, tubelet_size=2, use_learnable_pos_emb=False,
                 with_cp=False, all_frames=16, cos_attn=False, cam_emb_placed_at=None, cam_emb_type="add",
                 compute_dtype: torch.dtype = torch.float16, device=None):
        super().__init__()
        if cos_attn or use_learnable_pos_emb or num_classes or cam_emb_placed_at is not None:
            raise NotImplementedError("cos_attn / learnable pos-emb / classifier head / camera embedding are outside "
                                      "the L4P inference path (SURVEY.md §2 rows 4, 11)")
        self.embed_dim, self.depth, self.num_heads = embed_dim, depth, num_heads
        self.patch_size, self.tubelet_size = patch_size, tubelet_size
        self.compute_dtype = compute_dtype
        eps = _norm_eps(norm_layer)
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, all_frames, tubelet_size, device=device)
        self.pos_embed = sinusoid_table(self.patch_embed.num_patches, embed_dim)  # plain tensor, not in the ckpt
        self.blocks = nn.ModuleList(
            [Block(embed_dim, num_heads, mlp_ratio, qkv_bias, qk_scale, eps, init_values, device=device)
             for _ in range(depth)])
        self.norm = P.LayerNorm(embed_dim, eps, device=device)
        self.head = P.Identity()
        hooks = {0, depth * 14 // 40, depth * 21 // 40, depth * 28 // 40, depth * 36 // 40, depth}
        self.keep_features: Iterable[int] | str = sorted(hooks)
        self._packed: Optional[Dict[str, object]] = None
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate())

    # ------------------------------------------------------------------ weights
    def invalidate(self) -> None:
        self._packed = None

    def prepare(self, device: torch.device) -> Dict[str, object]:
        """Pack kernel-ready weights on `device`: 16-bit GEMM operands, fp32 biases / norm params."""
        if self._packed is not None and self._packed["device"] == device and self._packed["dtype"] == self.compute_dtype:
            return self._packed
        dt = self.compute_dtype
        f32 = dict(device=device, dtype=torch.float32)

        def w16(t):
            return t.detach().to(device=device, dtype=dt).contiguous()

        def f(t):
            return t.detach().to(**f32).contiguous()

        pk: Dict[str, object] = {"device": device, "dtype": dt}
        D = self.embed_dim
        pk["pe_w"] = w16(self.patch_embed.proj.weight.reshape(D, -1))
        pk["pe_b"] = f(self.patch_embed.proj.bias)
        pk["pos"] = f(self.pos_embed[0])
        blocks = []
        for blk in self.blocks:
            a = blk.attn
            if a.q_bias is not None:
                qkv_b = torch.cat([a.q_bias.detach(), torch.zeros_like(a.v_bias), a.v_bias.detach()])
            else:
                qkv_b = torch.zeros(3 * D)
            blocks.append(dict(
                n1w=f(blk.norm1.weight), n1b=f(blk.norm1.bias), eps1=blk.norm1.eps,
                qkv_w=w16(a.qkv.weight), qkv_b=f(qkv_b),
                proj_w=w16(a.proj.weight), proj_b=f(a.proj.bias),
                n2w=f(blk.norm2.weight), n2b=f(blk.norm2.bias), eps2=blk.norm2.eps,
                fc1_w=w16(blk.mlp.fc1.weight), fc1_b=f(blk.mlp.fc1.bias),
                fc2_w=w16(blk.mlp.fc2.weight), fc2_b=f(blk.mlp.fc2.bias),
                scale=float(a.scale)))
        pk["blocks"] = blocks
        pk["nw"], pk["nb"], pk["neps"] = f(self.norm.weight), f(self.norm.bias), self.norm.eps
        self._packed = pk
        return pk

    def _workspace(self, B: int, ntok: int, device: torch.device) -> Dict[str, torch.Tensor]:
        key = (B, ntok, device, self.compute_dtype)
        ws = self._ws.get(key)
        if ws is None:
            D, H = self.embed_dim, self.num_heads
            hd = D // H
            dpad = 96
            dt = self.compute_dtype
            hidden = self.blocks[0].mlp.fc1.out_features   # module attributes, not parameter shapes: the fp32 masters may
            # have been released after packing (l4p_b200/arena.py)
            kdim = self.patch_embed.proj.in_channels * self.tubelet_size * self.patch_size * self.patch_size
            M = B * ntok
            e = lambda *s, dtype=dt: torch.empty(*s, device=device, dtype=dtype)
            ws = dict(
                patches=e(M, kdim), x=e(M, D, dtype=torch.float32), ln=e(M, D), att=e(M, D), hid=e(M, hidden),
                # pad lanes of q/k/vt are never written by the QKV epilogue: zero them once
                q=torch.zeros(B, H, ntok, dpad, device=device, dtype=dt),
                k=torch.zeros(B, H, ntok, dpad, device=device, dtype=dt),
                vt=torch.zeros(B, H, dpad, ntok, device=device, dtype=dt))
            assert hd <= dpad
            self._ws = {key: ws}  # keep one shape resident
        return ws

    # ------------------------------------------------------------------ compute
    @torch.no_grad()
    def forward(self, x: torch.Tensor, intrinsics_b44t=None, extrinsics_b44t=None) -> FeatureList:
        _require_device(x)
        B, Cin, T, H, W = x.shape
        assert H == self.patch_embed.img_size[0] and W == self.patch_embed.img_size[1], (
            f"Input image size ({H}*{W}) doesn't match model "
            f"({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]}).")
        pk = self.prepare(x.device)
        ntok = (T // self.tubelet_size) * (H // self.patch_size) * (W // self.patch_size)
        assert ntok == pk["pos"].shape[0], "window length does not match the position table"
        ws = self._workspace(B, ntok, x.device)
        D, Hh = self.embed_dim, self.num_heads
        hd = D // Hh
        keep = range(self.depth + 1) if self.keep_features == "all" else set(self.keep_features)

        xin = x.contiguous().float()
        ops.patchify(xin, ws["patches"], (self.tubelet_size, self.patch_size, self.patch_size))
        xs = ws["x"]
        ops.linear(ws["patches"], pk["pe_w"], bias=pk["pe_b"], res_f32=pk["pos"], res_row_mod=ntok, out_f32=xs)

        feats: List[Optional[torch.Tensor]] = [None] * (self.depth + 1)
        taps16: Dict[int, torch.Tensor] = {}

        def tap(i: int, src32: torch.Tensor) -> None:
            if i in keep:
                feats[i] = src32.view(B, ntok, D).clone()
                t16 = torch.empty(B * ntok, D, device=x.device, dtype=self.compute_dtype)
                ops.cast16(src32, t16)
                taps16[i] = t16

        tap(0, xs)
        for i, w in enumerate(pk["blocks"], start=1):
            ops.layernorm(xs, w["n1w"], w["n1b"], w["eps1"], out16=ws["ln"])
            ops.linear_qkv(ws["ln"], w["qkv_w"], w["qkv_b"], ws["q"], ws["k"], ws["vt"], Hh, hd, ntok)
            ops.attention(ws["q"], ws["k"], ws["vt"], ws["att"], hd, w["scale"])
            ops.linear(ws["att"], w["proj_w"], bias=w["proj_b"], res_f32=xs, out_f32=xs)
            ops.layernorm(xs, w["n2w"], w["n2b"], w["eps2"], out16=ws["ln"])
            ops.linear(ws["ln"], w["fc1_w"], bias=w["fc1_b"], act=_l.ACT_GELU, out_16=ws["hid"])
            ops.linear(ws["hid"], w["fc2_w"], bias=w["fc2_b"], res_f32=xs, out_f32=xs)
            if i < self.depth:
                tap(i, xs)
        # features_list[-1] = head(norm(features_list[-1]))  (l4p_videomae.py:115)
        last32 = torch.empty(B, ntok, D, device=x.device, dtype=torch.float32)
        last16 = torch.empty(B * ntok, D, device=x.device, dtype=self.compute_dtype)
        ops.layernorm(xs, pk["nw"], pk["nb"], pk["neps"], out16=last16, out32=last32)
        feats[self.depth] = last32
        taps16[self.depth] = last16
        return FeatureList(feats, taps16)
