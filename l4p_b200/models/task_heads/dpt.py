"""DPT dense decoder (3-D conv pyramid) on the B200 kernels.

State-dict compatible with the reference's `PixelwiseTaskWithDPT` / `DPTOutputAdapter_fix`
(l4p/models/task_heads/dpt/dust3r/dpt_head.py:26-115, l4p/models/task_heads/dpt/croco/dpt_block.py:29-278,
344-509), including the aliased duplicate keys `scratch.layerN_rn.*` == `scratch.layer_rn.{N-1}.*` and the
unused `refinenet4.resConfUnit1.*`.

Compute (all activations channels-last [B,T,H,W,C] 16-bit, fp32 accumulation, one kernel per conv):
    tap i  -> GEMM 1x1x1 -> {ConvT(k==s) GEMM + pixel-shuffle | identity | im2col + GEMM (stride 2)}    K7
           -> implicit-GEMM 3x3x3 conv to 256 ch, epilogue also emits ReLU(x)                           K8
    refinenet: RCU = conv(ReLU) -> conv(ReLU) + skip(s) fused in the conv epilogue                        K8
               out_conv (1x1x1) is applied BEFORE the trilinear upsample (both linear, interpolation
               weights sum to 1, so conv1x1(up(x)) == up(conv1x1(x))): 4-8x fewer FLOPs                 K9
    head1 conv -> trilinear resize -> head2 conv + ReLU + 1x1x1 conv (+exp) fused epilogue -> fp32 NCTHW  K8 K10
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import nn

from ... import lib as _l
from ... import ops
from .. import params as P


class _RCU(nn.Module):
    def __init__(self, features: int, device=None):
        super().__init__()
        self.conv1 = P.Conv3d(features, features, (3, 3, 3), padding=(1, 1, 1), device=device)
        self.conv2 = P.Conv3d(features, features, (3, 3, 3), padding=(1, 1, 1), device=device)


class _Fusion(nn.Module):
    def __init__(self, features: int, scale_factor, device=None):
        super().__init__()
        self.scale_factor = tuple(scale_factor)
        self.out_conv = P.Conv3d(features, features, (1, 1, 1), device=device)
        self.resConfUnit1 = _RCU(features, device)
        self.resConfUnit2 = _RCU(features, device)


class _Seq(nn.Module):
    """nn.Sequential-like container with integer child names (keys '0.weight', '2.bias', ...)."""

    def __init__(self, items: Dict[int, nn.Module]):
        super().__init__()
        for i, m in items.items():
            self.add_module(str(i), m)


def _reassemble_op(cin: int, cout: int, sf: Sequence[int], device=None) -> nn.Module:
    assert all(s >= 0 for s in sf) or all(s <= 0 for s in sf)
    if any(s > 0 for s in sf):
        stride = tuple(2 ** s for s in sf)
        return P.ConvTranspose3d(cin, cout, stride, stride, device=device)
    if any(s < 0 for s in sf):
        stride = tuple(2 ** (-s) for s in sf)
        k = tuple((s // 2) * 2 + 1 for s in stride)
        return P.Conv3d(cin, cout, k, stride=stride, padding=tuple(s // 2 for s in stride), device=device)
    return P.Identity()


class DPTOutputAdapter_fix(nn.Module):
    def __init__(self, num_channels=1, hooks=(2, 5, 8, 11), layer_dims=(96, 192, 384, 768), feature_dim=256, last_dim=32,
                 dim_tokens_enc=(1408,) * 4, patch_size=(2, 14, 14),
                 actpost_scale_factors=((1, 2, 2), (1, 1, 1), (0, 0, 0), (-1, -1, -1)),
                 fusion_scale_factors=((1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)), output_size=None, device=None):
        super().__init__()
        self.num_channels, self.hooks = num_channels, list(hooks)
        self.layer_dims, self.feature_dim, self.last_dim = list(layer_dims), feature_dim, last_dim
        self.patch_size = tuple(patch_size)
        self.actpost_scale_factors = tuple(tuple(s) for s in actpost_scale_factors)
        self.fusion_scale_factors = tuple(tuple(s) for s in fusion_scale_factors)
        self.output_size = None if output_size is None else tuple(output_size)

        scratch = nn.Module()
        rn = [P.Conv3d(layer_dims[i], feature_dim, (3, 3, 3), padding=(1, 1, 1), bias=False, device=device)
              for i in range(4)]
        scratch.layer1_rn, scratch.layer2_rn, scratch.layer3_rn, scratch.layer4_rn = rn
        scratch.layer_rn = nn.ModuleList(rn)  # aliases of the four modules above (duplicate state-dict keys)
        scratch.refinenet1 = _Fusion(feature_dim, self.fusion_scale_factors[0], device)
        scratch.refinenet2 = _Fusion(feature_dim, self.fusion_scale_factors[1], device)
        scratch.refinenet3 = _Fusion(feature_dim, self.fusion_scale_factors[2], device)
        scratch.refinenet4 = _Fusion(feature_dim, self.fusion_scale_factors[3], device)
        self.scratch = scratch
        self.head1 = _Seq({0: P.Conv3d(feature_dim, feature_dim // 2, (3, 3, 3), padding=(1, 1, 1), device=device)})
        self.head2 = _Seq({0: P.Conv3d(feature_dim // 2, last_dim, (3, 3, 3), padding=(1, 1, 1), device=device),
                           2: P.Conv3d(last_dim, num_channels, (1, 1, 1), device=device)})
        self.act_postprocess = nn.ModuleList([
            _Seq({0: P.Conv3d(dim_tokens_enc[i], layer_dims[i], (1, 1, 1), device=device),
                  1: _reassemble_op(layer_dims[i], layer_dims[i], self.actpost_scale_factors[i], device)})
            for i in range(4)])
        self._packed: Optional[Dict[str, object]] = None
        self.debug: Optional[Dict[str, torch.Tensor]] = None  # set to {} to capture intermediates (tests)
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate())

    def invalidate(self) -> None:
        self._packed = None

    # ------------------------------------------------------------------ weight packing
    @staticmethod
    def _pack_conv3(w: torch.Tensor, device, dt) -> torch.Tensor:
        """[Cout,Cin,kt,kh,kw] -> [Cout, (kt,kh,kw,Cin)] 16-bit (K order of the implicit-GEMM conv)."""
        co = w.shape[0]
        return w.detach().permute(0, 2, 3, 4, 1).reshape(co, -1).to(device=device, dtype=dt).contiguous()

    def prepare(self, device, dt) -> Dict[str, object]:
        if self._packed is not None and self._packed["device"] == device and self._packed["dtype"] == dt:
            return self._packed
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        pk: Dict[str, object] = {"device": device, "dtype": dt}
        act = []
        for i in range(4):
            seq = self.act_postprocess[i]
            c1, op = getattr(seq, "0"), getattr(seq, "1")
            e = dict(w1=c1.weight.detach().reshape(c1.out_channels, -1).to(device=device, dtype=dt).contiguous(),
                     b1=f(c1.bias))
            if isinstance(op, P.ConvTranspose3d):
                st, sh, sw = op.stride
                # [Cin,Cout,kt,kh,kw] -> rows (kt,kh,kw,co), cols Cin
                e["kind"] = "convT"
                e["stride"] = op.stride
                e["w2"] = op.weight.detach().permute(2, 3, 4, 1, 0).reshape(st * sh * sw * op.out_channels, -1) \
                    .to(device=device, dtype=dt).contiguous()
                e["b2"] = f(op.bias).repeat(st * sh * sw).contiguous()
            elif isinstance(op, P.Conv3d):
                e["kind"] = "conv_s"
                e["stride"] = op.stride
                assert op.kernel_size == (3, 3, 3), "strided reassemble conv: only 3x3x3 (scale factor -1) is built"
                e["w2"] = self._pack_conv3(op.weight, device, dt)
                e["b2"] = f(op.bias)
            else:
                e["kind"] = "id"
            e["rn"] = self._pack_conv3(self.scratch.layer_rn[i].weight, device, dt)
            act.append(e)
        pk["act"] = act
        fus = []
        for name in ("refinenet1", "refinenet2", "refinenet3", "refinenet4"):
            m = getattr(self.scratch, name)
            d = dict(scale=m.scale_factor,
                     ow=m.out_conv.weight.detach().reshape(m.out_conv.out_channels, -1).to(device=device, dtype=dt).contiguous(),
                     ob=f(m.out_conv.bias))
            for r in ("resConfUnit1", "resConfUnit2"):
                u = getattr(m, r)
                d[r] = dict(w1=self._pack_conv3(u.conv1.weight, device, dt), b1=f(u.conv1.bias),
                            w2=self._pack_conv3(u.conv2.weight, device, dt), b2=f(u.conv2.bias))
            fus.append(d)
        pk["fusion"] = fus
        h1, h20, h22 = getattr(self.head1, "0"), getattr(self.head2, "0"), getattr(self.head2, "2")
        pk["h1w"], pk["h1b"] = self._pack_conv3(h1.weight, device, dt), f(h1.bias)
        pk["h2w"], pk["h2b"] = self._pack_conv3(h20.weight, device, dt), f(h20.bias)
        pk["h3w"], pk["h3b"] = f(h22.weight.reshape(h22.out_channels, -1)), f(h22.bias)
        self._packed = pk
        return pk

    # ------------------------------------------------------------------ compute
    def _rcu(self, w, x, x_relu, extra_res=None, want_relu=True, groups=1):
        """RCU(x) (+ extra_res): conv2(relu(conv1(relu(x)))) + x. Returns (out, relu(out) or None)."""
        mid = torch.empty_like(x)
        ops.conv3d(x_relu, w["w1"], ksize=(3, 3, 3), bias=w["b1"], act=_l.ACT_RELU, out_16=mid, groups=groups)
        out = torch.empty_like(x)
        out_relu = torch.empty_like(x) if want_relu else None
        ops.conv3d(mid, w["w2"], ksize=(3, 3, 3), bias=w["b2"], res_16=x, res2_16=extra_res, out_16=out,
                   out_16_relu=out_relu, groups=groups)
        return out, out_relu

    def _fuse(self, w, x0, x0_relu, x1=None, x1_relu=None, crop=None, groups=1):
        """FeatureFusionBlock_custom.forward (dpt_block.py:210-238) with out_conv commuted before the upsample."""
        if x1 is not None:
            s, s_relu = self._rcu(w["resConfUnit1"], x1, x1_relu, extra_res=x0, groups=groups)  # s = x0 + RCU1(x1)
        else:
            s, s_relu = x0, x0_relu
        y, _ = self._rcu(w["resConfUnit2"], s, s_relu, want_relu=False, groups=groups)
        B, T, H, W, C = y.shape
        z = torch.empty_like(y)
        if groups > 1:   # per-head 1x1x1 weights and biases: the conv path carries the group of every tile
            ops.conv3d(y, w["ow"], ksize=(1, 1, 1), bias=w["ob"], out_16=z, groups=groups)
        else:
            ops.linear(y.view(-1, C), w["ow"], bias=w["ob"], out_16=z.view(-1, C))
        st, sh, sw = w["scale"]
        osz = (T * st, H * sh, W * sw)
        if osz == (T, H, W):
            up = z
        else:
            up = torch.empty(B, *osz, C, device=y.device, dtype=y.dtype)
            ops.upsample3d(z, osz, align_corners=True, y=up)
        if crop is not None and (up.shape[1] > crop[0] or up.shape[2] > crop[1]):
            up = up[:, : crop[0], : crop[1]].contiguous()
        return up

    @torch.no_grad()
    def forward(self, taps16: Sequence[torch.Tensor], batch: int, image_size: Tuple[int, int, int], *,
                exp_out: bool = False) -> torch.Tensor:
        """taps16: the four hooked token tensors [B*tokens, C] 16-bit (hook order). Returns fp32 [B,Cout,T',H',W']."""
        return forward_grouped([self], taps16, batch, image_size, exp_outs=[exp_out])[0]

    # ------------------------------------------------------------------ several adapters of identical architecture at once
    def group_signature(self):
        """Adapters with equal signatures can run as one group (forward_grouped): everything but the output channel count."""
        return (tuple(self.hooks), tuple(self.layer_dims), self.feature_dim, self.last_dim, self.patch_size,
                self.actpost_scale_factors, self.fusion_scale_factors, self.output_size,
                tuple(getattr(self.act_postprocess[i], "0").in_channels for i in range(4)))


def _stack_packed(pks):
    """Weights of the group's adapters stacked along the output-row axis (one block per adapter), biases concatenated, for
    the layers that run as ONE grouped launch: layer_rn, the four fusion blocks, head1."""
    cat = lambda key_fn: torch.cat([key_fn(pk) for pk in pks], dim=0).contiguous()
    g = {"rn": [cat(lambda pk, i=i: pk["act"][i]["rn"]) for i in range(4)], "fusion": []}
    for j in range(4):
        d = dict(scale=pks[0]["fusion"][j]["scale"], ow=cat(lambda pk: pk["fusion"][j]["ow"]), ob=cat(lambda pk: pk["fusion"][j]["ob"]))
        for r in ("resConfUnit1", "resConfUnit2"):
            d[r] = {k: cat(lambda pk, k=k: pk["fusion"][j][r][k]) for k in ("w1", "b1", "w2", "b2")}
        g["fusion"].append(d)
    g["h1w"], g["h1b"] = cat(lambda pk: pk["h1w"]), cat(lambda pk: pk["h1b"])
    return g


@torch.no_grad()
def forward_grouped(adapters: Sequence[DPTOutputAdapter_fix], taps16: Sequence[torch.Tensor], batch: int,
                    image_size: Tuple[int, int, int], *, exp_outs: Sequence[bool]):
    """The DPT decoders of several heads that read the SAME taps and share one architecture (the flow, depth and motion-mask
    heads of the shipped config: dense_heads.py:20-217) as one launch sequence: every layer between the per-head token
    projections and the per-head final convolution runs ONCE on the heads' activations stacked along the batch axis, with
    grouped weights (`ops.conv3d(..., groups=G)`: the tile's batch entry selects the head's weight block and bias). Three heads
    cost 3 x 38 launches one by one; grouped 38 + 2 x 10, and the low-resolution pyramid levels (2048-16384 voxels per head,
    split-K territory) become three times larger problems. A single adapter (`forward`) is the G = 1 case of the same code.
    Returns one fp32 [B, Cout_g, T', H', W'] tensor per adapter."""
    G = len(adapters)
    a0 = adapters[0]
    assert all(a.group_signature() == a0.group_signature() for a in adapters), "forward_grouped: different DPT architectures"
    dev, dt = taps16[0].device, taps16[0].dtype
    pks = [a.prepare(dev, dt) for a in adapters]
    if G == 1:
        pg = dict(rn=[pks[0]["act"][i]["rn"] for i in range(4)], fusion=pks[0]["fusion"], h1w=pks[0]["h1w"], h1b=pks[0]["h1b"])
    else:
        key = tuple(id(pk) for pk in pks)
        cache = a0.__dict__.setdefault("_group_packed", {})
        if key not in cache:
            cache.clear()
            cache[key] = _stack_packed(pks)
        pg = cache[key]
    T, H, W = image_size
    nt, nh, nw = T // a0.patch_size[0], H // a0.patch_size[1], W // a0.patch_size[2]
    B = batch
    GB = G * B
    layers, layers_relu = [], []
    for i in range(4):
        z_all = None
        for g, pk in enumerate(pks):          # per-head token projection + reassemble op, written into the head's batch slice
            a = pk["act"][i]
            tok = taps16[i]
            c1 = a["w1"].shape[0]
            y = torch.empty(B * nt * nh * nw, c1, device=dev, dtype=dt)
            ops.linear(tok, a["w1"], bias=a["b1"], out_16=y)
            y = y.view(B, nt, nh, nw, c1)
            if a["kind"] == "convT":
                st, sh, sw = a["stride"]
                shape = (nt * st, nh * sh, nw * sw)
            elif a["kind"] == "conv_s":
                st, sh, sw = a["stride"]
                shape = ((nt - 1) // st + 1, (nh - 1) // sh + 1, (nw - 1) // sw + 1)
            else:
                shape = (nt, nh, nw)
            if a["kind"] == "id" and G == 1:
                z_all = y
                continue
            if z_all is None:
                z_all = torch.empty(GB, *shape, c1, device=dev, dtype=dt)
            z = z_all[g * B:(g + 1) * B]
            if a["kind"] == "convT":
                ops.conv_transpose3d(y, a["w2"], a["b2"], a["stride"], z)
            elif a["kind"] == "conv_s":
                col = torch.empty(B * shape[0] * shape[1] * shape[2], 27 * c1, device=dev, dtype=dt)
                ops.im2col3(y, col, a["stride"])
                ops.linear(col, a["w2"], bias=a["b2"], out_16=z.view(-1, c1))
            else:
                z.copy_(y)
        l = torch.empty(*z_all.shape[:4], a0.feature_dim, device=dev, dtype=dt)
        lr = torch.empty_like(l)
        ops.conv3d(z_all, pg["rn"][i], ksize=(3, 3, 3), out_16=l, out_16_relu=lr, groups=G)
        layers.append(l)
        layers_relu.append(lr)
    f1, f2, f3, f4 = pg["fusion"]
    p4 = a0._fuse(f4, layers[3], layers_relu[3], crop=(layers[2].shape[1], layers[2].shape[2]), groups=G)
    p3 = a0._fuse(f3, p4, None, layers[2], layers_relu[2], groups=G)
    p2 = a0._fuse(f2, p3, None, layers[1], layers_relu[1], groups=G)
    p1 = a0._fuse(f1, p2, None, layers[0], layers_relu[0], groups=G)
    h1 = torch.empty(*p1.shape[:4], a0.feature_dim // 2, device=dev, dtype=dt)
    ops.conv3d(p1, pg["h1w"], ksize=(3, 3, 3), bias=pg["h1b"], out_16=h1, groups=G)
    for g, ad in enumerate(adapters):
        if ad.debug is not None:
            cf = lambda t: t[g * B:(g + 1) * B].float().permute(0, 4, 1, 2, 3)
            ad.debug.update(l0=cf(layers[0]), l1=cf(layers[1]), l2=cf(layers[2]), l3=cf(layers[3]), p4=cf(p4), p3=cf(p3),
                            p2=cf(p2), p1=cf(p1), h1=cf(h1))
    osz = tuple(image_size) if a0.output_size is None else a0.output_size
    if tuple(h1.shape[1:4]) != osz:
        r = torch.empty(GB, *osz, h1.shape[-1], device=dev, dtype=dt)
        ops.upsample3d(h1, osz, align_corners=True, y=r)
        h1 = r
    outs = []
    for g, (ad, pk) in enumerate(zip(adapters, pks)):   # per-head final convolution: fused ReLU + 1x1x1 (+exp) epilogue
        out = torch.empty(B, ad.num_channels, *osz, device=dev, dtype=torch.float32)
        ops.conv3d(h1[g * B:(g + 1) * B], pk["h2w"], ksize=(3, 3, 3), bias=pk["h2b"], head_w2=pk["h3w"], head_b2=pk["h3b"],
                   head_exp=bool(exp_outs[g]), out_f32=out)
        outs.append(out)
    return outs


class PixelwiseTaskWithDPT(nn.Module):
    """Drop-in for dust3r's PixelwiseTaskWithDPT (dpt_head.py:89-115): holds `.dpt`."""

    def __init__(self, *, n_cls_token=0, hooks_idx=None, dim_tokens=None, output_width_ratio=1, num_channels=1,
                 device=None, **kwargs):
        super().__init__()
        assert n_cls_token == 0, "Not implemented"
        kwargs.pop("is_use_conv3d", None)
        kwargs.pop("head_type", None)
        args = dict(num_channels=num_channels, device=device, **kwargs)
        if hooks_idx is not None:
            args["hooks"] = hooks_idx
        if dim_tokens is not None:
            args["dim_tokens_enc"] = dim_tokens
        self.dpt = DPTOutputAdapter_fix(**args)
